#!/bin/bash
# usage: scripts/build_variant.sh <name> "<extra nvcc -D flags>"   -> hsmc_b200/csrc/variants/<name>.so
# Kernel-variant experiments (loaded with HSMC_GPU_LIB=...); same flags as hsmc_b200/build.py otherwise.
set -e
cd "$(dirname "$0")/.."
mkdir -p hsmc_b200/csrc/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -DHSMC_FAST_U01 --extended-lambda \
  -Xcompiler -fPIC -shared $2 -Xptxas=-v -o hsmc_b200/csrc/variants/$1.so hsmc_b200/csrc/hsmc_gpu.cu -lnccl 2>&1 \
  | grep -A2 "k_sweep_leanILb0" | grep -E "registers|spill" | tr '\n' ' '
echo " <- $1 ($2)"
