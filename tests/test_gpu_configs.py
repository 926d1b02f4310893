"""BASELINE.json configurations 1-3 run end to end through the C host driver on the GPU
(short versions: the point is that every shape, keyword and code path works and keeps the
hard-sphere invariants; speed is bench.py's business, statistics test_gpu_stat.py's)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT
from hsmc_outputs import collect

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "hsmc_b200", "host", "hsmc_b200")

CONFIG1 = """# config 1: the `hsmc -e` example shape (SC, N=1000, rho 0.5), neigh_list 1.05 (SURVEY 0.10)
rho 0.5
cells_x 10
cells_y 10
cells_z 10
type 1
neigh_list 1.05 10
dr_max 0.05
opt 1 200 10 0.5 0.5
press_virial 0.002 10
seed 124787
restart_write 200
config_write 200 100
sweep_eq 300
sweep_stat 400
out 100
"""

CONFIG2 = """# config 2: NVT rho 0.9, N = 32000, pressure + rdf + widom
rho 0.9
cells_x 20
cells_y 20
cells_z 20
type 2
neigh_list 1.05 10
dr_max 0.1
opt 1 100 10 0.5 0.5
press_virial 0.002 20
press_thermo 0.0001 0.002 20
rdf 0.01 5.0 50 100
widom 1000 20
seed 7
sweep_eq 100
sweep_stat 200
out 50
"""

CONFIG3 = """# config 3: NpT P = 10 from rho 0.94, N = 108000
npt 10 0.001
rho 0.94
cells_x 30
cells_y 30
cells_z 30
type 2
neigh_list 1.1 12
dr_max 0.05
opt 0 100 10 0.5 0.5
press_thermo 0.0001 0.002 20
seed 99
sweep_eq 60
sweep_stat 100
out 20
"""


def _run(text, args=(), env=None):
    d = tempfile.mkdtemp(prefix="hsmc_b200_cfg_")
    with open(os.path.join(d, "in.dat"), "w") as f:
        f.write(text)
    r = subprocess.run([EXE, "-o", "out.txt", *args], cwd=d, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, **(env or {})))
    log = open(os.path.join(d, "out.txt")).read() if os.path.exists(os.path.join(d, "out.txt")) else ""
    assert r.returncode == 0 and "Simulation complete!" in log, (r.stdout + r.stderr + log)[-3000:]
    _run.stderr = r.stderr
    return d, log


@pytest.fixture(scope="module", autouse=True)
def _built(lib_built):
    if lib_built.load_library().hsmc_gpu_device_count() < 1:
        pytest.fail("no CUDA device visible")
    from hsmc_b200 import build
    build.build_host()


def test_config1_example_shape():
    d, log = _run(CONFIG1)
    assert "Number of particles: 1000" in log and "Optimal maximum displacement:" in log
    obs = collect(d)
    assert obs["g_contact"].shape[0] == 40                       # sweeps 300..699 sampled every 10
    assert abs(obs["g_contact"].mean() - 2.16) < 0.25            # Carnahan-Starling at rho 0.5
    # restart file: the reference's byte layout (16 + 64 + 8 + N*32 + 5000) + 16 bytes of Philox state
    rs = [f for f in os.listdir(d) if f.startswith("restart_")]
    assert rs and os.path.getsize(os.path.join(d, sorted(rs)[-1])) == 16 + 64 + 8 + 1000 * 32 + 5000 + 16
    assert os.path.exists(os.path.join(d, "config_000000.dat.gz"))


def test_config2_nvt_32000():
    d, log = _run(CONFIG2)
    assert "Number of particles: 32000" in log
    obs = collect(d)
    assert obs["g_contact"].shape[0] == 10 and obs["presst_h"].shape == (10, 20)
    assert obs["rdf_g"].shape[0] == 4 and obs["widom_frac"].shape[0] == 10
    assert np.all(obs["rdf_g"] >= 0) and obs["rdf_g"][:, 0].mean() > 3.0      # g(1+) of a dense fluid/solid
    moves = float(log.split("-- Particle moves:")[1].split()[0])
    assert moves == 300 * 32000


def test_config3_npt_108000():
    d, log = _run(CONFIG3)
    assert "Number of particles: 108000" in log and "Pressure: 10.00000000" in log
    obs = collect(d)
    assert obs["density"].shape[0] == 5
    assert np.all((obs["density"] > 0.85) & (obs["density"] < 1.0))
    vol_moves = float(log.split("-- Volume moves:")[1].split()[0])
    assert 100 < vol_moves < 250                                  # about one per sweep


CONFIG_SNAP = """# N = 1,048,576: two snapshots of a GPU-evolved configuration
rho 0.9
cells_x 64
cells_y 64
cells_z 64
type 2
neigh_list 1.0 1
dr_max 0.1
opt 0 100 10 0.5 0.5
seed 3
config_write 5 100
sweep_eq 4
sweep_stat 10
out 5
"""


def test_streamed_snapshot_equals_the_whole_table_path():
    """hs_write_config on one GPU: the id-ordered table is fetched in slices (hsmc_gpu_pack_table +
    hsmc_gpu_fetch_rows) while earlier slices are being formatted and deflated.  Same chain, same bytes as the
    path that downloads the whole table first and as the serial gzprintf writer (the reference's own loop,
    io_config.c:134-191); and what was written is the configuration the device holds."""
    import gzip
    d1, _ = _run(CONFIG_SNAP)
    d2, _ = _run(CONFIG_SNAP, env={"HSMC_IO_NO_STREAM": "1"})
    d3, _ = _run(CONFIG_SNAP, env={"HSMC_IO_SERIAL": "1"})
    names = sorted(f for f in os.listdir(d1) if f.startswith("config_"))
    assert names and names == sorted(f for f in os.listdir(d2) if f.startswith("config_"))
    for f in names:
        a = gzip.open(os.path.join(d1, f)).read()
        assert a == gzip.open(os.path.join(d2, f)).read() == gzip.open(os.path.join(d3, f)).read()
        parts = a.split(b"# Sweep number\n")[1:]
        assert len(parts) == 2                                      # sweeps 5 and 10 appended to one file
        for part in parts:
            rows = np.loadtxt(part.split(b"# Configuration\n")[1].decode().splitlines())
            assert rows.shape == (4 * 64 ** 3, 4) and np.array_equal(rows[:, 0], np.arange(rows.shape[0]))
            assert np.all((rows[:, 1:] >= 0) & (rows[:, 1:] <= 64 * (4 / 0.9) ** (1 / 3) + 1e-6))
        first = np.loadtxt(parts[0].split(b"# Configuration\n")[1].decode().splitlines())
        second = np.loadtxt(parts[1].split(b"# Configuration\n")[1].decode().splitlines())
        assert 0.3 < np.mean(np.any(first[:, 1:] != second[:, 1:], axis=1)) <= 1.0     # the chain moved in between


def test_restart_round_trip():
    d, log = _run(CONFIG1)
    rs = sorted(f for f in os.listdir(d) if f.startswith("restart_"))
    text = CONFIG1.replace("opt 1 200 10 0.5 0.5", "opt 0 200 10 0.5 0.5") + f"restart_read 1 {os.path.join(d, rs[-1])}\n"
    d2, log2 = _run(text)
    assert "Reading data from restart file" in log2 and "Number of particles: 1000" in log2
    # ... and the unmodified reference (CPU) continues from the same file: same byte layout, the trailing
    # Philox counter is ignored
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "hsmc_ref")
    if os.path.exists(ref_exe):
        d3 = tempfile.mkdtemp(prefix="hsmc_b200_cfg_")
        short = text.replace("sweep_eq 300", "sweep_eq 20").replace("sweep_stat 400", "sweep_stat 20").replace("out 100", "out 10")
        with open(os.path.join(d3, "in.dat"), "w") as f:
            f.write(short)
        r = subprocess.run([ref_exe, "-i", "in.dat"], cwd=d3, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "Number of particles: 1000" in r.stdout and "Production completed." in r.stdout, r.stdout[-2000:]


CONFIG_SLAB = """# slab-decomposed driver run: every observable, snapshots, a restart file
rho 0.85
cells_x 16
cells_y 8
cells_z 8
type 2
neigh_list 1.05 10
dr_max 0.12
opt 1 60 6 0.5 0.5
press_virial 0.002 10
press_thermo 0.0001 0.002 10
rdf 0.02 4.0 20 100
widom 20000 10
ql 6 1.5 10
seed 31
restart_write 40
config_write 40 100
sweep_eq 40
sweep_stat 60
out 20
"""

CONFIG_SLAB_NPT = """# NpT on slabs: volume moves with all-reduced verdicts; Lx starts 0.02 % above 26 cells of 1.1, so the first
# accepted compressions change the cell grid (26 -> 24 layers) and the slabs have to be redistributed
npt 8 0.004
rho 0.6999
cells_x 16
cells_y 8
cells_z 8
type 2
neigh_list 1.1 12
dr_max 0.1
opt 0 60 6 0.5 0.5
press_thermo 0.0001 0.002 10
seed 5
config_write 50 100
sweep_eq 50
sweep_stat 50
out 25
"""


@pytest.mark.parametrize("text", [CONFIG_SLAB, CONFIG_SLAB_NPT], ids=["nvt", "npt"])
@pytest.mark.parametrize("halo", ["p2p", "nccl"])
def test_driver_on_two_gpus_equals_single_gpu_chain(lib_built, text, halo):
    """`hsmc_b200 -g 2` (one forked process per GPU, x-slabs, rank 0 writes the files) against the
    single-GPU driver told to use the block partition of a 2-slab run: the same Markov chain, so
    every output file must be byte-identical (snapshots, restart, histograms, mu, q_l, density)."""
    if lib_built.load_library().hsmc_gpu_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d2, log2 = _run(text, args=("-g", "2"), env={"HSMC_P2P": "1" if halo == "p2p" else "0", "HSMC_DEBUG_MP": "1"})
    if "\nnpt " in text:
        # the box shrinks through a change of the cell grid: the slabs must have been redistributed
        assert "slabs redistributed" in _run.stderr, _run.stderr[-2000:]
    d1, log1 = _run(text, env={"HSMC_XPART_WORLD": "2"})
    names = sorted(f for f in os.listdir(d1) if f not in ("in.dat", "out.txt"))
    assert names == sorted(f for f in os.listdir(d2) if f not in ("in.dat", "out.txt")) and len(names) >= 3
    import gzip
    for f in names:
        rd = (lambda p: gzip.open(p).read()) if f.endswith(".gz") else (lambda p: open(p, "rb").read())
        assert rd(os.path.join(d1, f)) == rd(os.path.join(d2, f)), f
    # (NCCL announces its version on stdout when NCCL_DEBUG asks for it)
    strip = lambda log: [ln for ln in log.splitlines() if not ln.startswith(("Elapsed time", "NCCL version"))]
    assert strip(log1) == strip(log2)


PATCHED_REF = os.path.join(ROOT, "oracle", "_ref", "hsmc_gpu_patched")


@pytest.mark.skipif(os.environ.get("HSMC_TEST_PATCHED_REF") == "0" or not os.path.exists(PATCHED_REF),
                    reason="needs oracle/_ref/hsmc_gpu_patched (the reference's sources patched per INTEGRATION.md, built by "
                           "oracle/Makefile where /root/reference exists; the binary travels to the GPU box)")
@pytest.mark.parametrize("text", [CONFIG_SLAB, CONFIG3.replace("cells_x 30", "cells_x 12").replace("cells_y 30", "cells_y 12")
                                  .replace("cells_z 30", "cells_z 12")], ids=["nvt_all_observables", "npt"])
def test_patched_reference_and_drop_in_driver_write_the_same_files(text):
    """The reference's OWN drivers, optimizer and output writers on top of libhsmc_gpu.so (the seam patch of
    INTEGRATION.md, integration/patch_reference.py) against this repository's host driver: same seed, same
    library, hence the same chain -- every output file must carry the same bytes (snapshots after gunzip,
    restart files up to the trailing Philox counter)."""
    import gzip
    d_mine, _ = _run(text)
    d_ref = tempfile.mkdtemp(prefix="hsmc_b200_patched_")
    with open(os.path.join(d_ref, "in.dat"), "w") as f:
        f.write(text)
    r = subprocess.run([PATCHED_REF, "-i", "in.dat"], cwd=d_ref, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "Simulation complete!" in r.stdout, (r.stdout + r.stderr)[-3000:]
    names = sorted(f for f in os.listdir(d_ref) if f != "in.dat")
    assert names == sorted(f for f in os.listdir(d_mine) if f not in ("in.dat", "out.txt")) and names
    for f in names:
        a, b = os.path.join(d_ref, f), os.path.join(d_mine, f)
        if f.endswith(".gz"):
            assert gzip.open(a).read() == gzip.open(b).read(), f
        elif f.startswith("restart_"):
            ra, rb = open(a, "rb").read(), open(b, "rb").read()
            assert rb[: len(ra)] == ra and len(rb) == len(ra) + 16, f
        else:
            assert open(a, "rb").read() == open(b, "rb").read(), f
