"""CPU tests of the host-side logic: slab planning (single process and world_size 2 / 4 over
gloo), the blocking error analysis against the reference's own python/hsmc_stat.py, the
output-file parsers against the reference executable's files, and the benchmark's
reference-arm rank gating."""
import os
import subprocess
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def test_plan_single(lib_built):
    from hsmc_b200 import gpu as G
    p = G.plan([32.8829, 32.8829, 32.8829], 1.0)
    assert p["cells"] == (32, 32, 32) and p["own_x"] == (0, 32)
    assert all(c % 2 == 0 for c in p["cells"]) and min(p["cell_size"]) >= 1.0
    p = G.plan([12.5992, 12.5992, 12.5992], 1.05)            # config-1 shape: 11.99 -> 10 even cells
    assert p["cells"] == (10, 10, 10) and min(p["cell_size"]) >= 1.05
    with pytest.raises(G.HsmcError, match="too small"):
        G.plan([3.5, 20, 20], 1.0)
    with pytest.raises(G.HsmcError, match="too few cell layers"):
        G.plan([10.0, 10.0, 10.0], 1.0, world=4, rank=0)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_plan_slabs_tile_the_box(lib_built, world):
    from hsmc_b200 import gpu as G
    box = [420.9, 210.45, 210.45]
    plans = [G.plan(box, 1.0, world, r) for r in range(world)]
    nx = plans[0]["cells"][0]
    edges = [p["own_x"] for p in plans]
    assert edges[0][0] == 0 and edges[-1][1] == nx
    for a, b in zip(edges[:-1], edges[1:]):
        assert a[1] == b[0]
    assert all((e[1] - e[0]) % 2 == 0 and e[0] % 2 == 0 and e[1] - e[0] >= 4 for e in edges)
    assert max(e[1] - e[0] for e in edges) - min(e[1] - e[0] for e in edges) <= 2


_WORKER = r'''
import os, sys, json
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["HSMC_ROOT"])
from hsmc_b200 import gpu as G
from bench import fcc_lattice
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
# the 128-byte communicator id travels rank 0 -> all exactly as in bench.py
ids = [bytes(range(128)) if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
assert ids[0] == bytes(range(128))
box, conf = fcc_lattice(12, 4, 4, 0.9)
p = G.plan(box, 1.0, world, rank)
nx = p["cells"][0]
w = p["cell_size"][0]
ix = np.floor(conf[:, 1] / w).astype(int) % nx            # unshifted grid: what upload() keeps
mine = conf[(ix >= p["own_x"][0]) & (ix < p["own_x"][1])]
got = [None] * world
dist.all_gather_object(got, (p["own_x"], mine[:, 0].astype(int).tolist()))
if rank == 0:
    ids_all = sorted(i for _, l in got for i in l)
    ok = ids_all == list(range(conf.shape[0])) and [g[0] for g in got] == sorted(g[0] for g in got)
    print("HOST_LOGIC", "PASS" if ok else "FAIL", [g[0] for g in got], [len(g[1]) for g in got])
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 4])
def test_slab_ownership_over_gloo(lib_built, tmp_path, world):
    """world_size-2/4 processes on CPU (gloo): every particle of a full table is claimed by
    exactly one rank's slab, slabs are ordered, the communicator id broadcast works."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, HSMC_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert "HOST_LOGIC PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_blocking_matches_reference_hsmc_stat():
    ref_py = "/root/reference/python"
    if not os.path.isdir(ref_py):
        pytest.skip("reference python/ not present on this machine")
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    sys.path.insert(0, ref_py)
    try:
        import hsmc_stat
    finally:
        sys.path.remove(ref_py)
    from blocking import blocking_std
    rng = np.random.default_rng(0)
    x = np.cumsum(rng.normal(size=5000)) * 0.01 + rng.normal(size=5000)     # correlated series
    mine = blocking_std(x)
    ref = hsmc_stat.blocking_std(x.copy(), plt_flag=False, print_flag=False)
    assert np.allclose(mine, ref, rtol=0, atol=0)


def test_output_parsers_on_reference_fixtures():
    from blocking import std_error
    s1 = dict(np.load(os.path.join(GOLDEN, "stat", "S1_nvt_rho08_ref.npz")))
    assert s1["g_contact"].shape == (2048,) and s1["pressv_rr"].shape == (25,)
    eta = np.pi * 0.8 / 6
    assert abs(s1["g_contact"].mean() - (1 - eta / 2) / (1 - eta) ** 3) < 0.1       # Carnahan-Starling
    assert 0 < std_error(s1["g_contact"]) < 0.05
    s3 = dict(np.load(os.path.join(GOLDEN, "stat", "S3_npt_p3471_ref.npz")))
    assert abs(np.pi * s3["density"].mean() / 6 - 0.349) < 0.01                     # README.md:151-158


_MP_TEST_C = r"""
#include <stdio.h>
#include <string.h>
#include "hs_mp.h"
int main(void) {
  hs_mp mp;
  const int W = 3, N = 3000;
  int rank = hs_mp_start(&mp, W, N);
  /* id broadcast from rank 0 */
  unsigned char id[128];
  for (int i = 0; i < 128; i++) id[i] = rank == 0 ? (unsigned char)(i * 7 + 1) : 0;
  hs_mp_bcast_id(&mp, id, 128);
  int ok = 1;
  for (int i = 0; i < 128; i++) ok &= id[i] == (unsigned char)(i * 7 + 1);
  /* neighbour blobs on the ring */
  unsigned char blob[64];
  memset(blob, 100 + rank, 64);
  const void *l, *r;
  hs_mp_exchange_blobs(&mp, blob, &l, &r);
  ok &= ((const unsigned char *)l)[5] == 100 + (rank + W - 1) % W && ((const unsigned char *)r)[63] == 100 + (rank + 1) % W;
  hs_mp_barrier(&mp);
  /* every rank files the rows it "owns" in the shared table; everybody sees all of them afterwards */
  for (int i = rank; i < N; i += W) { mp.table[i][0] = i; mp.table[i][1] = rank; }
  hs_mp_barrier(&mp);
  for (int i = 0; i < N; i++) ok &= mp.table[i][0] == i && mp.table[i][1] == i % W;
  hs_mp_scratch(&mp, rank)[0] = (unsigned long)ok;
  hs_mp_barrier(&mp);
  if (rank == 0) {
    int all = 1;
    for (int q = 0; q < W; q++) all &= (int)hs_mp_scratch(&mp, q)[0];
    printf("MP_PLUMBING %s\n", all ? "PASS" : "FAIL");
  } else {
    printf("this line must not appear: children are silent\n");
  }
  return hs_mp_finish(&mp);
}
"""


_MP_DEATH_C = r"""
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include "hs_mp.h"
int main(void) {
  hs_mp mp;
  int rank = hs_mp_start(&mp, 3, 16);
  if (rank == 2) { usleep(100000); abort(); }      /* a rank crashes while the others wait for it */
  hs_mp_barrier(&mp);
  printf("must not get here\n");
  return hs_mp_finish(&mp);
}
"""


def test_a_dying_rank_takes_the_run_down_instead_of_hanging_it(tmp_path):
    host = os.path.join(ROOT, "hsmc_b200", "host")
    src = tmp_path / "d.c"
    src.write_text(_MP_DEATH_C)
    exe = tmp_path / "d"
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-I", host, str(src), os.path.join(host, "hs_mp.c"), "-o", str(exe), "-lpthread"],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=30)
    assert out.returncode != 0 and "died" in out.stderr and "must not get here" not in out.stdout


def test_host_driver_process_per_gpu_plumbing(tmp_path):
    """hs_mp.c (fork before CUDA, shared table, process-shared barrier, id / blob exchange) with
    three ranks on CPU; rank 0 alone reports."""
    host = os.path.join(ROOT, "hsmc_b200", "host")
    src = tmp_path / "t.c"
    src.write_text(_MP_TEST_C)
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-I", host, str(src), os.path.join(host, "hs_mp.c"), "-o", str(exe), "-lpthread"],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == "MP_PLUMBING PASS", out.stdout + out.stderr


@pytest.mark.parametrize("total,world", [(100_000_000, 8), (1000, 3), (7, 8), (0, 2)])
def test_widom_insertion_shares_tile_the_range(total, world):
    """bench.py --workload widom: the shares of the insertion index range are disjoint and cover it"""
    from bench import shard_range
    nxt = 0
    for r in range(world):
        first, count = shard_range(total, r, world)
        assert first == nxt and count >= 0
        nxt = first + count
    assert nxt == total


# ---- the drop-in driver's input handling against the reference executable (no GPU needed) --------
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "hsmc_ref")
DRV_EXE = os.path.join(ROOT, "hsmc_b200", "host", "hsmc_b200")


@pytest.fixture(scope="module")
def driver_built(lib_built):
    from hsmc_b200 import build
    build.build_host()
    if not os.path.exists(REF_EXE):
        pytest.skip("oracle/_ref/hsmc_ref not built on this machine")
    return DRV_EXE


@pytest.mark.parametrize("text", ["rho 0.5\nbogus_key 3\n", "rho\n", "opt 1 1000\n", "neigh_list 1.0\n", "widom 100\n",
                                  "rho 0.5\nnpt 10\n", "rdf 0.01 5.0 10\n", "restart_read 1 " + "x" * 120 + "\n"],
                         ids=["unknown_key", "no_value", "opt_short", "neigh_short", "widom_short", "npt_short", "rdf_short",
                              "restart_name_too_long"])
def test_driver_rejects_bad_input_exactly_like_the_reference(driver_built, tmp_path, text):
    """same messages on stdout, same exit code (read_input.c:138-441, 477-500)"""
    (tmp_path / "in.dat").write_text(text)
    ref = subprocess.run([REF_EXE, "-i", "in.dat"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    mine = subprocess.run([driver_built, "-i", "in.dat"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert ref.returncode == mine.returncode == 1
    assert mine.stdout == ref.stdout


def test_driver_reads_the_reference_example_and_refuses_to_run_without_a_gpu(driver_built, tmp_path):
    """The reference's own `-e` example goes through the drop-in parser and set-up (same box and particle
    lines as the reference prints); on a machine without a CUDA device the run then stops with an ERROR --
    there is no CPU fallback.  (On the GPU box this test only checks the set-up lines.)"""
    ex = subprocess.run([REF_EXE, "-e"], capture_output=True, text=True, timeout=60).stdout
    ex = ex.replace("sweep_eq 1000000", "sweep_eq 2").replace("sweep_stat 1000000", "sweep_stat 2").replace(
        "opt 1 1000 10 0.5 0.5", "opt 0 1000 10 0.5 0.5").replace("out 10000", "out 1")
    (tmp_path / "in.dat").write_text(ex)
    ref = subprocess.run([REF_EXE, "-i", "in.dat"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    (tmp_path / "mine").mkdir()
    (tmp_path / "mine" / "in.dat").write_text(ex)
    mine = subprocess.run([driver_built, "-i", "in.dat"], cwd=tmp_path / "mine", capture_output=True, text=True, timeout=120)
    head = lambda out: [ln for ln in out.splitlines() if ln.startswith(("Reading input", "Done", "Simulation box", "Number of particles"))]
    assert ref.returncode == 0 and head(ref.stdout) == head(mine.stdout) and len(head(ref.stdout)) == 4
    import hsmc_b200
    if hsmc_b200.load_library().hsmc_gpu_device_count() < 1:
        assert mine.returncode == 1 and "ERROR: no CUDA device" in mine.stdout and "no CPU fallback" in mine.stdout
    else:
        assert mine.returncode == 0 and "Simulation complete!" in mine.stdout


def test_our_example_input_is_accepted_by_the_reference(driver_built, tmp_path):
    ex = subprocess.run([driver_built, "-e"], capture_output=True, text=True, timeout=60).stdout
    ex = ex.replace("sweep_eq 1000000", "sweep_eq 2").replace("sweep_stat 1000000", "sweep_stat 2").replace(
        "opt 1 1000 10 0.5 0.5", "opt 0 1000 10 0.5 0.5").replace("out 10000", "out 1")
    (tmp_path / "in.dat").write_text(ex)
    ref = subprocess.run([REF_EXE, "-i", "in.dat"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert ref.returncode == 0 and "Number of particles: 1000" in ref.stdout and "Production completed." in ref.stdout


def test_driver_reads_a_restart_file_written_by_the_reference(driver_built, tmp_path):
    """restart_%d.bin of the unmodified reference (io_config.c:28-74: dr_max, dv_max, box_info, p_info, the
    {id,x,y,z} table, the MT19937 state) is accepted by the drop-in driver: same box, same particle count."""
    text = ("rho 0.7\ncells_x 4\ncells_y 5\ncells_z 6\ntype 2\nneigh_list 1.0 10\ndr_max 0.1\nopt 0 10 2 0.5 0.5\n"
            "seed 3\nrestart_write 4\nsweep_eq 4\nsweep_stat 4\nout 2\n")
    (tmp_path / "in.dat").write_text(text)
    ref = subprocess.run([REF_EXE, "-i", "in.dat"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert ref.returncode == 0, ref.stdout[-2000:]
    rs = sorted(f for f in os.listdir(tmp_path) if f.startswith("restart_"))
    assert rs and os.path.getsize(tmp_path / rs[-1]) == 16 + 64 + 8 + 480 * 32 + 5000
    d2 = tmp_path / "again"
    d2.mkdir()
    (d2 / "in.dat").write_text(text + f"restart_read 1 {tmp_path / rs[-1]}\n")
    mine = subprocess.run([driver_built, "-i", "in.dat"], cwd=d2, capture_output=True, text=True, timeout=120)
    ref2 = subprocess.run([REF_EXE, "-i", "in.dat"], cwd=d2, capture_output=True, text=True, timeout=120)
    pick = lambda out: [ln for ln in out.splitlines() if ln.startswith(("Reading data from restart", "Simulation box", "Number of particles"))]
    assert ref2.returncode == 0 and pick(mine.stdout) == pick(ref2.stdout) and len(pick(mine.stdout)) == 3
    assert "Number of particles: 480" in mine.stdout


# ---- block partition of the two-level checkerboard: the same on every rank (host-only planning) -------
def _fcc_box(nx, ny, nz, rho):
    a = (4.0 / rho) ** (1.0 / 3.0)
    return [nx * a, ny * a, nz * a], 4 * nx * ny * nz


@pytest.mark.parametrize("cells,rho", [((256, 128, 128), 0.9), ((40, 10, 12), 0.85), ((24, 10, 12), 0.85), ((64, 64, 64), 0.9),
                                       ((30, 30, 30), 0.94), ((100, 8, 6), 0.5), ((37, 11, 9), 0.7)])
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_block_shape_is_the_same_on_every_rank_and_on_the_single_gpu_mimic(lib_built, cells, rho, world):
    """The block shape is part of the chain's definition.  Choosing it from the rank's own slab (slabs differ
    by two layers) once made a 4-GPU run a different chain from the single-GPU run; it must be a function of
    the grid and the world size only, and the x cuts of the ranks must tile the single-GPU partition."""
    from hsmc_b200 import gpu as G
    box, n = _fcc_box(*cells, rho)
    try:
        plans = [G.plan_blocks(box, n, world=world, rank=r) for r in range(world)]
    except G.HsmcError as e:
        assert "too few cell layers" in str(e)
        return
    single = G.plan_blocks(box, n, xpart_world=world)
    assert all(p["ok"] for p in plans) and single["ok"]
    assert len({(p["max_extent"], p["blocks"][1:], p["staged_capacity"], p["smem_bytes"]) for p in plans}) == 1
    assert plans[0]["max_extent"] == single["max_extent"] and plans[0]["blocks"][1:] == single["blocks"][1:]
    cuts = []
    for r, p in enumerate(plans):
        own = G.plan(box, 1.0, world, r)["own_x"]
        assert p["xcuts"][0] == own[0] and p["xcuts"][-1] == own[1] and p["blocks"][0] % 2 == 0
        assert all(b > a for a, b in zip(p["xcuts"], p["xcuts"][1:]))
        cuts += p["xcuts"][:-1]
    assert cuts + [plans[-1]["xcuts"][-1]] == single["xcuts"]
    # every block with its one-cell halo fits the staging limits the kernel was compiled with
    mx, my, mz = single["max_extent"]
    assert (mx + 2) * (my + 2) <= 128 and mz + 3 <= 32 and single["smem_bytes"] <= 100 * 1024


def test_block_plan_of_the_benchmark_box(lib_built):
    from hsmc_b200 import gpu as G
    box, n = _fcc_box(256, 128, 128, 0.9)
    p = G.plan_blocks(box, n)
    # 256 threads x 4 CTAs per SM leave 56 KB of shared memory per CTA: 8 x 8 x 27-cell blocks
    assert p["blocks"] == (54, 28, 8) and p["max_extent"] == (8, 8, 27) and p["ctas_per_phase"] == 1512
    assert p["xcuts"][0] == 0 and p["xcuts"][-1] == 420
    box, n = _fcc_box(162, 162, 162, 0.9)               # bench.py's default box
    p = G.plan_blocks(box, n)
    assert p["blocks"] == (34, 34, 10) and p["max_extent"] == (8, 8, 27) and p["smem_bytes"] <= 56 * 1024


def test_optimizer_step_guard(tmp_path):
    """the reference's secant step (optimizer.c:45-58, 115-139) divides by the difference of two
    acceptance ratios unguarded; equal ratios give inf/NaN and poison the run.  hs_opt.h keeps the previous step
    instead, and otherwise reproduces the reference's update (cap at 1.0, sign flip, halving) exactly."""
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <math.h>
#include <stdio.h>
#include "hs_opt.h"
static double ref_dr(double x1, double y1, double x2, double y2, double t) {   /* optimizer.c:45-58 */
  double v = x2 - (y2 - t) * (x2 - x1) / (y2 - y1);
  if (v > 1.0) v = 1.0; else if (v <= 0.0) { v = -v; if (v > 1.0) v = x2 / 2; }
  return v;
}
static double ref_dv(double x1, double y1, double x2, double y2, double t) {   /* optimizer.c:115-139 */
  double v = x2 - (y2 - t) * (x2 - x1) / (y2 - y1);
  if (v <= 0.0) { v = -v; if (v > 0.1) v = x2 / 2; }
  return v;
}
int main(void) {
  int bad = 0;
  /* regular steps: identical to the reference's arithmetic */
  double c[][5] = {{0.05, 0.8, 0.1, 0.6, 0.5}, {0.1, 0.6, 0.15, 0.45, 0.5}, {0.4, 0.52, 0.8, 0.3, 0.5},
                   {0.2, 0.3, 0.1, 0.9, 0.5}, {0.001, 0.7, 0.002, 0.65, 0.5}, {0.9, 0.55, 1.0, 0.54, 0.5}};
  for (int i = 0; i < 6; i++) {
    if (hs_opt_next_dr(c[i][0], c[i][1], c[i][2], c[i][3], c[i][4]) != ref_dr(c[i][0], c[i][1], c[i][2], c[i][3], c[i][4])) bad |= 1;
    if (hs_opt_next_dv(c[i][0], c[i][1], c[i][2], c[i][3], c[i][4]) != ref_dv(c[i][0], c[i][1], c[i][2], c[i][3], c[i][4])) bad |= 2;
  }
  /* equal acceptance ratios: the reference produces inf / NaN, the guard keeps the previous step */
  if (isfinite(ref_dr(0.05, 0.5, 0.1, 0.5, 0.5))) bad |= 4;        /* 0/0 */
  if (hs_opt_next_dr(0.05, 0.5, 0.1, 0.5, 0.5) != 0.1) bad |= 8;
  if (hs_opt_next_dr(0.05, 0.7, 0.1, 0.7, 0.5) != 0.05) bad |= 16;   /* -inf: the reference's own clamp (sign flip, then x2/2) already yields a usable step */
  if (hs_opt_next_dv(0.001, 0.3, 0.002, 0.3, 0.5) != 0.002) bad |= 32;
  if (hs_opt_next_dv(0.001, 0.5, 0.002, 0.5, 0.5) != 0.002) bad |= 64;
  /* a NaN sample (0 volume moves attempted: 0/0 acceptance) never reaches the step */
  if (hs_opt_next_dv(0.001, NAN, 0.002, 0.4, 0.5) != 0.002) bad |= 128;
  if (hs_opt_next_dr(0.05, 0.6, 0.1, NAN, 0.5) != 0.1) bad |= 256;
  /* results are always usable steps */
  for (double y1 = 0.0; y1 <= 1.0; y1 += 0.125) for (double y2 = 0.0; y2 <= 1.0; y2 += 0.125) {
    double v = hs_opt_next_dr(0.05, y1, 0.1, y2, 0.5), w = hs_opt_next_dv(0.001, y1, 0.002, y2, 0.5);
    if (!(v > 0.0 && v <= 1.0) || !(w > 0.0 && w < 1e300)) bad |= 512;
  }
  printf("%d\n", bad);
  return bad != 0;
}
''')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-I", os.path.join(ROOT, "hsmc_b200", "host"), "-o", str(exe), str(src), "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "0", r.stdout
