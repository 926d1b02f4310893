#!/bin/bash
# round 2, GPU call 1: parity of the lean sweep, memcheck, first timings (1 GPU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.log
timeout 200 python scripts/lean_debug.py > gpurun_out/r02a_debug.log 2>&1; tail -20 gpurun_out/r02a_debug.log
timeout 1200 python -m pytest tests/test_gpu_sweep.py -q --tb=short 2>&1 | tail -40 > gpurun_out/r02a_pytest_sweep.log; tail -5 gpurun_out/r02a_pytest_sweep.log
timeout 400 compute-sanitizer --tool memcheck python scripts/lean_debug.py 8 6 7 0.85 0.15 2 > gpurun_out/r02a_memcheck.log 2>&1; tail -5 gpurun_out/r02a_memcheck.log
timeout 600 python scripts/lean_bench.py --sweeps 20 --check > gpurun_out/r02a_bench.log 2>&1
timeout 200 python scripts/lean_bench.py --sweeps 20 --impl 6 >> gpurun_out/r02a_bench.log 2>&1
for v in t_192_4 t_256_3 t_128_6 t_256_4 t_128_8; do
  HSMC_GPU_LIB=$PWD/hsmc_b200/csrc/variants/$v.so timeout 200 python scripts/lean_bench.py --sweeps 20 >> gpurun_out/r02a_bench.log 2>&1
done
for b in 8,8,12 8,8,28 6,6,24 4,4,12; do
  HSMC_BLOCK=$b timeout 200 python scripts/lean_bench.py --sweeps 20 >> gpurun_out/r02a_bench.log 2>&1
done
cat gpurun_out/r02a_bench.log
