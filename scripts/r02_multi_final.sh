#!/bin/bash
# final multi-GPU call: usage  gpurun --gpus N -- 'bash scripts/r02_multi_final.sh N'
# identity check with its per-quantity report (NVLink-window and NCCL halo paths), strong-scaling bench line with
# chain_identical, and the same with two launches per sweep (HSMC_SLAB_LINK=0) for comparison
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/r02_multi${N}_gpus.log
: > gpurun_out/r02_multi_gpu_check_world${N}.txt
for P2P in 1 0; do
  HSMC_CHECK_CELLS=$([ "$N" = "2" ] && echo 24,10,12 || echo 40,10,12) HSMC_CHECK_P2P=$P2P timeout 300 python -m torch.distributed.run --nnodes=1 \
    --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29620 + P2P)) tests/multi_gpu_check.py 2>/dev/null | grep -v "^\*\|OMP_NUM" >> gpurun_out/r02_multi_gpu_check_world${N}.txt
done
grep "MULTI_GPU_CHECK" gpurun_out/r02_multi_gpu_check_world${N}.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
HSMC_SLAB_LINK=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
  bench.py --gpus $N --steps 5 --warmup 3 --chain-check-sweeps 0 --e2e-steps 1 > gpurun_out/r02_bench_n${N}_link.json 2> gpurun_out/r02_bench_n${N}_link.err
python - <<PY
import json
for f in ("gpurun_out/r02_bench_n${N}.json", "gpurun_out/r02_bench_n${N}_link.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, "%.3e moves/s, e2e %.3e, S=%d, kernel/plan/build/halo share %.2f %.2f %.2f %.2f, chain_identical %s" % (
            d["value"], d["e2e"]["value"], d["config"]["sweeps_per_step"], r["kernel_share_of_step"], r["plan_share_of_step"],
            r["build_share_of_step"], r["halo_share_of_step"], d.get("chain_identical")))
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/r02_bench_n${N}.err
