for cfg in "12 1.15" "12 1.04" "11 1.06" "10 1.06" "8 1.08" "6 1.10"; do
  set -- $cfg
  echo "AZ=$1 CAPF=$2: $(HSMC_TILE_AZ=$1 HSMC_TILE_CAPF=$2 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["build_share_of_step"])')"
done
