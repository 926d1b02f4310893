#!/bin/bash
# round 2, multi-GPU call: usage  gpurun --gpus N -- 'bash scripts/r02_multi.sh N'
# slab-vs-single-GPU identity tests at world N (both halo paths), then the strong-scaling bench line with its
# chain_identical check, then the NCCL-halo bench for comparison
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/r02_multi${N}_gpus.log
timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=short -k "${N}-" 2>&1 | tail -15 > gpurun_out/r02_multi${N}_pytest.log; tail -3 gpurun_out/r02_multi${N}_pytest.log
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_configs.py -q --tb=short -k "two_gpus" 2>&1 | tail -8 >> gpurun_out/r02_multi${N}_pytest.log; tail -3 gpurun_out/r02_multi${N}_pytest.log
fi
# the identity check itself, with its per-quantity report (what the tests above assert on)
for P2P in 1 0; do
  HSMC_CHECK_CELLS=$([ "$N" = "2" ] && echo 24,10,12 || echo 40,10,12) HSMC_CHECK_P2P=$P2P timeout 600 python -m torch.distributed.run --nnodes=1 \
    --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29620 + P2P)) tests/multi_gpu_check.py 2>/dev/null | grep -v "^\*\|OMP_NUM" >> gpurun_out/r02_multi_gpu_check_world${N}.txt
done
tail -4 gpurun_out/r02_multi_gpu_check_world${N}.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
tail -c 1500 gpurun_out/r02_bench_n${N}.json; tail -3 gpurun_out/r02_bench_n${N}.err
HSMC_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
  bench.py --gpus $N --steps 5 --warmup 3 --chain-check-sweeps 0 > gpurun_out/r02_bench_n${N}_nccl.json 2> gpurun_out/r02_bench_n${N}_nccl.err
python - <<PY
import json
for f in ("gpurun_out/r02_bench_n${N}.json", "gpurun_out/r02_bench_n${N}_nccl.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, "%.3e moves/s, e2e %.3e, S=%d, kernel/plan/build/halo share %.2f %.2f %.2f %.2f, chain_identical %s" % (
            d["value"], d["e2e"]["value"], d["config"]["sweeps_per_step"], r["kernel_share_of_step"], r["plan_share_of_step"],
            r["build_share_of_step"], r["halo_share_of_step"], d.get("chain_identical")))
    except Exception as e:
        print(f, "unreadable:", e)
PY
