#!/bin/bash
# 1 GPU, final: suite, bench + reference arm, stamps, ncu of the final sweep kernel, launch list
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_1gpu.txt; cat gpurun_out/r02_pytest_gpu_1gpu.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 400 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_n1_ref.json 2> gpurun_out/r02_bench_n1_ref.err; tail -c 300 gpurun_out/r02_bench_n1_ref.json
( export HSMC_BLOCK_STAMPS=1; python scripts/lean_bench.py --cells 162 162 162 --sweeps 20 2>&1 | tail -2; python scripts/lean_bench.py --cells 20 162 162 --sweeps 20 2>&1 | tail -2; python scripts/lean_bench.py --cells 20 20 20 --sweeps 200 2>&1 | tail -2 ) > gpurun_out/r02_block_stamps.txt 2>&1; cat gpurun_out/r02_block_stamps.txt | cut -c1-300
bash scripts/r02_evidence.sh "ncu list" 2>&1 | tail -30
