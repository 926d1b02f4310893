#!/bin/bash
# 1 GPU: suite + bench with the 256 x 4 default, ncu of the fp32-prefiltered observable kernels, c3 bench lines
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_1gpu.txt; cat gpurun_out/r02_pytest_gpu_1gpu.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 600 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 ncu --set full --clock-control none -k regex:"k_overlap_scaled_f32|k_contact_hist_f32" -f -o /tmp/r02_obs python scripts/profile_all.py > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/r02_obs.ncu-rep > gpurun_out/r02_ncu_obs_f32.txt 2>&1; grep -E "^###|gpu__time_duration|issue_active|dram_throughput" gpurun_out/r02_ncu_obs_f32.txt
for w in c3; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; tail -c 700 gpurun_out/r02_bench_$w.json; tail -2 gpurun_out/r02_bench_$w.err
  timeout 900 python bench.py --workload $w --impl reference > gpurun_out/r02_bench_${w}_ref.json 2> gpurun_out/r02_bench_${w}_ref.err; tail -c 500 gpurun_out/r02_bench_${w}_ref.json
done
