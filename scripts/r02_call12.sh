#!/bin/bash
# round 2, GPU call 12: whole GPU suite + the default bench line + the reference arm (1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -25 > gpurun_out/r02m_pytest_gpu.log; tail -6 gpurun_out/r02m_pytest_gpu.log
HSMC_DEBUG_TILES=1 timeout 900 python bench.py > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err; tail -c 3000 gpurun_out/r02m_bench_n1.json; tail -5 gpurun_out/r02m_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02m_bench_ref.json 2> gpurun_out/r02m_bench_ref.err; tail -c 600 gpurun_out/r02m_bench_ref.json
