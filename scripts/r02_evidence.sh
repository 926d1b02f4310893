#!/bin/bash
# round 2 evidence call (1 GPU): ncu --set full of every kernel, the launch list of the bench command, bench lines of
# BASELINE configs 2, 3 and 5
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_all python scripts/profile_all.py > gpurun_out/r02_all.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02_all.ncu-rep > gpurun_out/r02_ncu_all_kernels.txt 2>&1
grep -E "^###|duration" gpurun_out/r02_ncu_all_kernels.txt | paste - - | awk '{print $2, $(NF-1), $NF}' | sort | uniq -c | sort -rn | head -40
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 3 --warmup 3 --sweeps-per-step 10 --no-cpu-baseline --no-secondary --e2e-steps 1 > /dev/null 2>&1
python profiles/launch_shares.py gpurun_out/r02_launches.csv > gpurun_out/r02_launch_shares.txt 2>&1; head -20 gpurun_out/r02_launch_shares.txt
for w in c2 c3; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; tail -c 700 gpurun_out/r02_bench_$w.json; tail -2 gpurun_out/r02_bench_$w.err
  timeout 900 python bench.py --workload $w --impl reference > gpurun_out/r02_bench_${w}_ref.json 2> gpurun_out/r02_bench_${w}_ref.err; tail -c 300 gpurun_out/r02_bench_${w}_ref.json
done
timeout 900 python bench.py --workload widom --steps 3 --warmup 3 > gpurun_out/r02_bench_widom_n1.json 2> gpurun_out/r02_bench_widom_n1.err; tail -c 1200 gpurun_out/r02_bench_widom_n1.json; tail -2 gpurun_out/r02_bench_widom_n1.err
