# fused block phases: correctness first (bounded), then timing against separate launches
timeout 300 python -m pytest tests/test_gpu_sweep.py -x -q -k "fused or variants or replay" 2>&1 | tail -5
export BENCH_EXTRA="--no-secondary"
for f in 1 0; do echo "HSMC_FUSE=$f"; HSMC_FUSE=$f timeout 200 bash scripts/block_sweep.sh "default 8,8,24" "default 8,8,16" "default 8,8,12" "default 6,6,16" "default 4,4,12"; done
