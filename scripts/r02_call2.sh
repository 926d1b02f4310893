#!/bin/bash
# round 2, GPU call 2: ncu --set full of k_block_plan and k_sweep_lean at the benchmark size (1 GPU)
mkdir -p gpurun_out
for k in k_block_plan k_sweep_lean; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r02b_$k \
    python scripts/lean_bench.py --sweeps 4 > gpurun_out/r02b_ncu_$k.log 2>&1
  python profiles/ncu_summary.py gpurun_out/r02b_$k.ncu-rep $k > gpurun_out/r02b_$k.txt 2>&1
  python profiles/ncu_source_hot.py gpurun_out/r02b_$k.ncu-rep 45 > gpurun_out/r02b_${k}_source.txt 2>&1
done
head -50 gpurun_out/r02b_k_block_plan.txt
ls -la gpurun_out/*.ncu-rep
