#!/bin/bash
# 1 GPU: the whole -m gpu suite, then the default bench line and its reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_1gpu.txt; cat gpurun_out/r02_pytest_gpu_1gpu.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 2500 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_n1_ref.json 2> gpurun_out/r02_bench_n1_ref.err; tail -c 900 gpurun_out/r02_bench_n1_ref.json
