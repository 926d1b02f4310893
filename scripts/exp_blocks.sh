export BENCH_EXTRA="--no-secondary"
bash scripts/block_sweep.sh "default 8,8,24" "v_old.so 8,8,24" "v_halves.so 8,8,24" "v_pipe.so 8,8,24" "v_halves_pipe.so 8,8,24" "default 8,8,24" "v_old.so 8,8,24"
