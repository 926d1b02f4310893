#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/lean_bench.py --sweeps 5 > gpurun_out/r02_dbg.log 2>&1
HSMC_BLOCK=8,8,16 timeout 300 python scripts/lean_bench.py --sweeps 5 >> gpurun_out/r02_dbg.log 2>&1
cat gpurun_out/r02_dbg.log
