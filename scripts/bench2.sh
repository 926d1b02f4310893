# usage: scripts/bench2.sh NGPU "<shape or ->" ...   one torchrun bench line per block shape
n=$1; shift
for shape in "$@"; do
  if [ "$shape" != "-" ]; then export HSMC_BLOCK=$shape; else unset HSMC_BLOCK; fi
  HSMC_DEBUG_TILES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 4 --warmup 3 --e2e-steps 1 > /tmp/b.json 2> /tmp/b.err
  echo "n=$n shape=$shape: $(grep -m1 blocks /tmp/b.err | cut -d: -f2-) | $(tail -1 /tmp/b.json | python -c '
import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("%.3e moves/s  phase %.1f us  build %.2f halo %.2f e2e %.3e" % (d["value"], 1e3*r["avg_launch_ms"], r["build_share_of_step"], r["halo_share_of_step"], d["e2e"]["value"]))')"
done
