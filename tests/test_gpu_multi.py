"""Multi-GPU slab decomposition (needs >= 2 CUDA devices; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("halo", ["p2p", "nccl", "p2p-link"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_run_is_bitwise_identical_to_single_gpu(lib_built, world, halo):
    """halo: boundary layers through NVLink windows, through NCCL, or -- p2p-link, HSMC_SLAB_LINK=1 -- stored by the
    sweep kernel itself into the neighbour's ghost layer (one launch per sweep)."""
    n = lib_built.load_library().hsmc_gpu_device_count()
    if n < world:
        pytest.skip(f"needs {world} GPUs, {n} visible")
    env = dict(os.environ, HSMC_CHECK_CELLS="40,10,12" if world > 2 else "24,10,12",
               HSMC_CHECK_P2P="0" if halo == "nccl" else "1", HSMC_SLAB_LINK="1" if halo == "p2p-link" else "0")
    halo = halo.split("-")[0]
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", str(29530 + world + (10 if halo == "p2p" else 0) + (20 if env["HSMC_SLAB_LINK"] == "1" else 0)),
         os.path.join(ROOT, "tests", "multi_gpu_check.py")],
        capture_output=True, text=True, env=env, timeout=600)
    assert "MULTI_GPU_CHECK PASS" in out.stdout and f"halo={halo}" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
