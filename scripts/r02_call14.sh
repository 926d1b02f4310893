#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep.py -q --tb=short -x 2>&1 | tail -12 > gpurun_out/r02o_pytest.log; tail -4 gpurun_out/r02o_pytest.log
timeout 300 compute-sanitizer --tool memcheck python -c "
import sys; sys.path.insert(0,'.')
import numpy as np, hsmc_b200, bench
box, conf = bench.fcc_lattice(8,6,7,0.85)
with hsmc_b200.HsmcGpu(conf.shape[0], box, seed=3, sweep_impl=7) as h:
    h.upload(conf); h.sweep_nvt(3, 0.15); print('min_r2', h.min_dist2(), h.counters())
" 2>&1 | tail -4 > gpurun_out/r02o_memcheck.log; tail -2 gpurun_out/r02o_memcheck.log
timeout 600 python scripts/lean_bench.py --sweeps 20 --impl 7 > gpurun_out/r02o_bench.log 2>&1
timeout 300 python scripts/lean_bench.py --sweeps 20 --impl 7 --cells 162 162 162 >> gpurun_out/r02o_bench.log 2>&1
timeout 300 python scripts/lean_bench.py --sweeps 20 --impl 8 >> gpurun_out/r02o_bench.log 2>&1
cat gpurun_out/r02o_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_gather -s 40 -c 4 -f -o gpurun_out/r02o_k_sweep_gather python scripts/lean_bench.py --sweeps 3 --impl 7 > gpurun_out/r02o_ncu.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02o_k_sweep_gather.ncu-rep k_sweep_gather > gpurun_out/r02o_k_sweep_gather.txt 2>&1
grep -E "duration|dram__bytes|inst_executed.sum|issue_active|warps_active|registers|l1tex__t_sector_hit|lts__t_sector_hit" gpurun_out/r02o_k_sweep_gather.txt | head -40
