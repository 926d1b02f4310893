#!/bin/bash
# block-shape experiment for slab runs: 2 GPUs on fcc 40x162x162 (slabs of ~33 cell layers, what each of 8 GPUs holds at
# the benchmark size); usage  gpurun --gpus 2 -- 'bash scripts/r02_slab_shapes.sh'
mkdir -p gpurun_out
: > gpurun_out/r02_slab_shapes.txt
port=29700
for shape in ${SHAPES:-auto 8,8,27 6,8,27 8,8,19 6,8,19 8,8,14 6,8,14 4,8,14 6,8,10}; do
  port=$((port+1))
  if [ "$shape" = "auto" ]; then unset HSMC_BLOCK; else export HSMC_BLOCK=$shape; fi
  HSMC_DEBUG_TILES=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus 2 --cells 40 162 162 --steps 3 --warmup 3 --sweeps-per-step 150 --e2e-steps 1 --chain-check-sweeps 0 \
    --no-cpu-baseline --no-secondary > gpurun_out/slab_shape.json 2> gpurun_out/slab_shape.err
  python - "$shape" <<'PY' >> gpurun_out/r02_slab_shapes.txt
import json, sys, re
try:
    d = json.loads(open("gpurun_out/slab_shape.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    S = d["config"]["sweeps_per_step"]
    tiles = [l for l in open("gpurun_out/slab_shape.err") if "rank 0: blocks" in l]
    print("%-8s %.3e moves/s  sweep %.3f ms  kernel/plan/build/halo %.2f %.2f %.2f %.2f  launch %.4f ms | %s" % (
        sys.argv[1], d["value"], d["ms_per_step"] / S, r["kernel_share_of_step"], r["plan_share_of_step"], r["build_share_of_step"],
        r["halo_share_of_step"], r["avg_launch_ms"], tiles[-1].strip()[-110:] if tiles else ""))
except Exception as e:
    print(sys.argv[1], "failed:", e, open("gpurun_out/slab_shape.err").read()[-400:])
PY
done
cat gpurun_out/r02_slab_shapes.txt
