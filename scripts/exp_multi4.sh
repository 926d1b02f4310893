run() { echo "=== P2P=$1 FUSE=$2"; HSMC_DEBUG_TILES=1 HSMC_CHECK_CELLS=40,10,12 HSMC_CHECK_P2P=$1 HSMC_FUSE=$2 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $3 tests/multi_gpu_check.py 2>&1 | grep -E "ok$|MISMATCH|MULTI_GPU|slabs|nccl calls|Error|error|blocks" | sort | uniq -c | head -40; }
run 1 1 29701
run 0 1 29702
