#!/bin/bash
# First GPU call of the next round (1 GPU, ~4 min of box time):
#   gpurun --timeout 600 -- 'bash scripts/round2_first.sh'
# 1. the opt-in byte comparison of the patched reference (oracle/_ref/hsmc_gpu_patched) with the drop-in
#    driver -- not yet green on hardware;  2. the whole GPU suite;  3. the headline bench and
#    the Widom workload;  4. an ncu --set full capture of the observables' kernels (not profiled yet).
mkdir -p gpurun_out
HSMC_TEST_PATCHED_REF=1 timeout 200 python -m pytest tests/test_gpu_configs.py -q -k patched 2>&1 | tail -25 | tee gpurun_out/r02_patched.log
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest.log
timeout 200 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 600 gpurun_out/r02_bench_n1.json
timeout 100 python bench.py --workload widom --steps 5 > gpurun_out/r02_bench_widom.json 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_contact_hist|k_overlap_scaled|k_widom|k_cell_scatter|k_cell_count' -c 8 -f \
  -o gpurun_out/r02_observables python bench.py --steps 1 --warmup 3 --sweeps-per-step 2 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
ls -la gpurun_out | tail -8
# multi-GPU (separate calls, charged N x):  gpurun --gpus 2 -- 'python -m pytest tests/test_gpu_multi.py tests/test_gpu_configs.py -q -k "2- or two_gpus"'
#                                           gpurun --gpus 8 -- 'bash scripts/bench2.sh 8 -'
# experiment prepared in round 1: the split colour barrier
#   scripts/build_variant.sh v_split "-DBLK_SPLIT_BARRIER"      (here, before the call)
#   gpurun -- 'HSMC_GPU_LIB=$PWD/hsmc_b200/csrc/variants/v_split.so python -m pytest tests/test_gpu_sweep.py -x -q -m gpu | tail -3;
#              BENCH_EXTRA=--no-secondary bash scripts/block_sweep.sh "default 8,8,24" "v_split.so 8,8,24"'
