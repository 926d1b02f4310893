#!/bin/bash
# round 2, GPU call 4: k_propose + self-contained lean kernel: parity, timings, ncu
mkdir -p gpurun_out
timeout 200 python scripts/lean_debug.py > gpurun_out/r02j_debug.log 2>&1; tail -4 gpurun_out/r02j_debug.log
timeout 1200 python -m pytest tests/test_gpu_sweep.py -q --tb=short -x 2>&1 | tail -30 > gpurun_out/r02j_pytest_sweep.log; tail -5 gpurun_out/r02j_pytest_sweep.log
timeout 400 compute-sanitizer --tool memcheck python scripts/lean_debug.py 8 6 7 0.85 0.15 2 > gpurun_out/r02j_memcheck.log 2>&1; tail -2 gpurun_out/r02j_memcheck.log
timeout 600 python scripts/lean_bench.py --sweeps 20 --check > gpurun_out/r02j_bench.log 2>&1
timeout 300 python scripts/lean_bench.py --sweeps 20 --cells 162 162 162 >> gpurun_out/r02j_bench.log 2>&1
for v in t_192_4 t_128_6 t_256_4 t_256_3; do
  HSMC_GPU_LIB=$PWD/hsmc_b200/csrc/variants/$v.so timeout 200 python scripts/lean_bench.py --sweeps 20 >> gpurun_out/r02j_bench.log 2>&1
done
for b in 8,8,28 8,8,16 6,6,24; do
  HSMC_BLOCK=$b timeout 200 python scripts/lean_bench.py --sweeps 20 >> gpurun_out/r02j_bench.log 2>&1
done
cat gpurun_out/r02j_bench.log
for k in k_propose k_sweep_lean; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r02j_$k \
    python scripts/lean_bench.py --sweeps 4 > gpurun_out/r02j_ncu_$k.log 2>&1
  python profiles/ncu_summary.py gpurun_out/r02j_$k.ncu-rep $k > gpurun_out/r02j_$k.txt 2>&1
  python profiles/ncu_source_hot.py gpurun_out/r02j_$k.ncu-rep 45 > gpurun_out/r02j_${k}_source.txt 2>&1
done
