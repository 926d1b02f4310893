#!/bin/bash
# A/B of kernel variants built by scripts/build_variant.sh at the benchmark size; usage: gpurun -- 'bash scripts/r02_variants.sh v_a v_b ...'
mkdir -p gpurun_out
touch gpurun_out/r02_variants.txt
for v in "$@"; do
  HSMC_GPU_LIB=$PWD/hsmc_b200/csrc/variants/$v.so timeout 300 python scripts/lean_bench.py --cells 162 162 162 --sweeps 40 2>&1 | grep -v "^ *$" >> gpurun_out/r02_variants.txt
done
tail -n $# gpurun_out/r02_variants.txt
