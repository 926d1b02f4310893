#!/bin/bash
mkdir -p gpurun_out
HSMC_DEBUG_TILES=1 timeout 300 python scripts/lean_bench.py --sweeps 20 > gpurun_out/r02l_bench.log 2>&1
timeout 600 python -m pytest tests/test_gpu_sweep.py -q --tb=short -x 2>&1 | tail -3 >> gpurun_out/r02l_bench.log
timeout 300 python scripts/lean_bench.py --sweeps 20 --cells 162 162 162 >> gpurun_out/r02l_bench.log 2>&1
for v in t_192_4; do
  HSMC_GPU_LIB=$PWD/hsmc_b200/csrc/variants/$v.so timeout 200 python scripts/lean_bench.py --sweeps 20 >> gpurun_out/r02l_bench.log 2>&1
done
HSMC_BLOCK_CAPF=1.08 timeout 300 python scripts/lean_bench.py --sweeps 20 >> gpurun_out/r02l_bench.log 2>&1
cat gpurun_out/r02l_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_lean -s 4 -c 1 -f -o gpurun_out/r02l_k_sweep_lean python scripts/lean_bench.py --sweeps 4 > gpurun_out/r02l_ncu.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02l_k_sweep_lean.ncu-rep k_sweep_lean > gpurun_out/r02l_k_sweep_lean.txt 2>&1
grep -E "occupancy|duration|shared_mem" gpurun_out/r02l_k_sweep_lean.txt
