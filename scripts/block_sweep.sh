# usage: scripts/block_sweep.sh "<lib-or-default> <bx,by,bz>" ...   (one bench line per configuration)
for cfg in "$@"; do
  set -- $cfg
  lib=$1; shape=$2
  if [ "$lib" != "default" ]; then export HSMC_GPU_LIB=$PWD/hsmc_b200/csrc/variants/$lib; else unset HSMC_GPU_LIB; fi
  echo "$lib $shape: $(HSMC_DEBUG_TILES=1 HSMC_BLOCK=$shape python bench.py --steps 3 --warmup 3 --sweeps-per-step 5 --no-cpu-baseline --e2e-steps 1 $BENCH_EXTRA 2>&1 | python -c '
import sys,json
blk=""
for ln in sys.stdin:
    if "blocks" in ln and not blk: blk=ln.strip().split(": ",1)[1]
    if ln.startswith("{"):
        d=json.loads(ln); print("%.3e moves/s  phase %.1f us  build %.2f |" % (d["value"], 1e3*d["roofline"]["avg_launch_ms"], d["roofline"]["build_share_of_step"]), blk)
')"
done
