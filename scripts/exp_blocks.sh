export BENCH_EXTRA="--no-secondary"
bash scripts/block_sweep.sh "default 8,8,24" "default 8,8,16" "default 8,8,12" "default 8,8,20" "default 8,8,28" "default 8,6,16" "default 6,6,20" \
  "t160.so 8,8,24" "t160.so 8,8,20" "t160c4.so 8,8,24" "t192.so 8,8,24" "t192.so 8,8,28" "t96.so 8,7,16" "t96.so 8,7,12"
