export BENCH_EXTRA="--no-secondary"
bash scripts/block_sweep.sh "c5.so 8,8,16" "c6.so 8,8,12" "t96c6.so 8,7,12" "t96c6.so 8,7,16"
