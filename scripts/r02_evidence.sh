#!/bin/bash
# round 2 evidence call (1 GPU): ncu --set full of every kernel (summaries made on the box: the report itself is too
# large to bring back), the launch list of the bench command, bench lines of BASELINE configs 2, 3 and 5, and the
# patched-reference byte comparison.  usage: gpurun -- 'bash scripts/r02_evidence.sh [parts]'   parts: ncu list cfg widom patched
PARTS=${1:-"ncu list cfg widom patched"}
mkdir -p gpurun_out
REP=/tmp/r02_all.ncu-rep
if [[ "$PARTS" == *ncu* ]]; then
  timeout 900 ncu --set full --clock-control none --import-source on -f -o /tmp/r02_all python scripts/profile_all.py > gpurun_out/r02_all.log 2>&1
  ls -la $REP >> gpurun_out/r02_all.log
  python profiles/ncu_summary.py $REP > gpurun_out/r02_ncu_all_kernels.txt 2>&1
  # the sweep kernel alone, with source: one launch
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_lean -s 1 -c 1 -f -o /tmp/r02_lean python scripts/profile_all.py > /dev/null 2>&1
  python profiles/ncu_summary.py /tmp/r02_lean.ncu-rep > gpurun_out/r02_ncu_k_sweep_lean.txt 2>&1
  python profiles/ncu_source_hot.py /tmp/r02_lean.ncu-rep 60 > gpurun_out/r02_ncu_k_sweep_lean_source.txt 2>&1
  python profiles/ncu_source_regions.py /tmp/r02_lean.ncu-rep "{'prologue+geometry':(300,343),'CSR rows+shadow prefetch+flag wait':(344,420),'row scan':(421,441),'chunk build':(442,494),'cell index pass+prefetch':(495,526),'staging':(527,560),'trial decode+record':(561,603),'stencil scan':(180,290),'scan call+mates':(604,614),'verdict rounds':(615,652),'log+colour barrier':(653,672),'commit':(673,719),'global path+epilogue+ghost delivery':(720,800)}" > gpurun_out/r02_ncu_k_sweep_lean_regions.txt 2>&1
  sz=$(stat -c %s /tmp/r02_lean.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -gt 0 ] && [ "$sz" -lt 25000000 ]; then cp /tmp/r02_lean.ncu-rep gpurun_out/; fi
  grep -E "^###|gpu__time_duration" gpurun_out/r02_ncu_all_kernels.txt | paste - - | awk '{print $2, $(NF-1), $NF}' | sort | uniq -c | sort -rn | head -30
  cat gpurun_out/r02_ncu_k_sweep_lean_regions.txt
fi
if [[ "$PARTS" == *list* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 3 --warmup 3 --sweeps-per-step 10 --no-cpu-baseline --no-secondary --e2e-steps 1 > /dev/null 2>&1
  python profiles/launch_shares.py gpurun_out/r02_launches.csv > gpurun_out/r02_launch_shares.txt 2>&1; head -14 gpurun_out/r02_launch_shares.txt
fi
if [[ "$PARTS" == *cfg* ]]; then
  for w in c2 c3; do
    timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; tail -c 700 gpurun_out/r02_bench_$w.json; tail -2 gpurun_out/r02_bench_$w.err
    timeout 900 python bench.py --workload $w --impl reference > gpurun_out/r02_bench_${w}_ref.json 2> gpurun_out/r02_bench_${w}_ref.err; tail -c 300 gpurun_out/r02_bench_${w}_ref.json
  done
fi
if [[ "$PARTS" == *widom* ]]; then
  timeout 900 python bench.py --workload widom --steps 3 --warmup 3 > gpurun_out/r02_bench_widom_n1.json 2> gpurun_out/r02_bench_widom_n1.err; tail -c 1500 gpurun_out/r02_bench_widom_n1.json; tail -2 gpurun_out/r02_bench_widom_n1.err
fi
if [[ "$PARTS" == *patched* ]]; then
  HSMC_TEST_PATCHED_REF=1 timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -k patched 2>&1 | tail -15 > gpurun_out/r02_patched_reference.txt; cat gpurun_out/r02_patched_reference.txt
fi
du -sh gpurun_out
