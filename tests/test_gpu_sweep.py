"""GPU tests of the checkerboard sweep (K2): every trial move's verdict replayed through
the reference's part_move arithmetic, invariants, determinism, counters."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

GOLDEN_FILES = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN_FILES]


@pytest.fixture(scope="module")
def hs(lib_built):
    import hsmc_b200
    if hsmc_b200.load_library().hsmc_gpu_device_count() < 1:
        pytest.fail("no CUDA device visible: GPU tests cannot run (there is no CPU fallback)")
    return hsmc_b200


def _replay(checker, log, dr_max):
    """Feed the GPU's trial log, in its serial order, to part_move() of the checker.

    Trials the checkerboard chain rejected for leaving their cell (verdict 2) never reach
    check_overlap on the GPU and are skipped; every other trial must get the same verdict
    from the reference arithmetic, and the final coordinates must agree bit for bit."""
    keep = log[log["verdict"] != 2]
    acc = checker.replay_moves(keep["id"], keep["raw"], dr_max)
    return keep, acc


@pytest.mark.parametrize("impl", [0, 8, 5, 7, 1], ids=["auto", "lean", "block_global", "gather", "cell_global"])
@pytest.mark.parametrize("path", GOLDEN_FILES, ids=IDS)
def test_every_trial_verdict_replays_through_oracle(hs, path, impl, oracle_built):
    g = dict(np.load(path))
    dr_max = float(g["dr_max"])
    N = g["conf"].shape[0]
    p = oracle_built.Port(g["conf"], g["box"], neigh_dr=float(g["neigh_dr"]), max_part=12)
    with hs.HsmcGpu(N, g["box"][:3], seed=2024, sweep_impl=impl) as h:
        h.upload(g["conf"])
        for sweep in range(3):
            log = h.sweep_nvt_logged(dr_max)
            assert len(log) == N                                  # one trial per particle
            assert np.array_equal(np.sort(log["id"]), np.arange(N))
            keep, acc = _replay(p, log, dr_max)
            assert np.array_equal(acc, (keep["verdict"] == 0).astype(np.int32))
            assert np.array_equal(h.download(), p.get_conf())
        c = h.counters()
        assert c[0] == 3 * N and c[1] + c[2] == c[0]
        assert c[1] == p.counters()[1]
        assert h.cell_rejects() == c[0] - p.counters()[0]


@pytest.mark.parametrize("name,lat,rho,neigh,dr_max", [
    ("C2", (2, 20, 20, 20), 0.9, 1.05, 0.1),        # BASELINE configs[1]: N = 32 000, rho 0.9
    ("C3", (2, 30, 30, 30), 0.94, 1.1, 0.06),       # BASELINE configs[2]: N = 108 000, rho 0.94 (NpT start)
])
def test_replay_at_baseline_sizes(hs, oracle_built, name, lat, rho, neigh, dr_max):
    """Every trial of three sweeps at the sizes BASELINE.json names for configs 2 and 3, replayed through the
    oracle's part_move (one C call per sweep): verdicts and final coordinates bit for bit.  At these sizes a
    box holds several blocks per axis, so interior blocks, wrapped blocks and the fused-phase protocol are all
    on the path (the small fixtures have two blocks per axis)."""
    box, conf = oracle_built.Port.lattice(*lat, rho)
    N = conf.shape[0]
    p = oracle_built.Port(conf, box, neigh_dr=neigh, max_part=16)
    with hs.HsmcGpu(N, box[:3], seed=424242, cell_min=neigh) as h:
        h.upload(conf)
        h.sweep_nvt(5, dr_max)                       # leave the perfect lattice first
        p.set_conf(h.download())
        for sweep in range(3):
            log = h.sweep_nvt_logged(dr_max)
            assert len(log) == N and np.array_equal(np.sort(log["id"]), np.arange(N))
            keep, acc = _replay(p, log, dr_max)
            assert np.array_equal(acc, (keep["verdict"] == 0).astype(np.int32)), name
            assert np.array_equal(h.download(), p.get_conf()), name
        assert h.min_dist2() >= 1.0


def test_replay_through_unmodified_reference(hs, oracle_built):
    """Same replay, but through the reference's own part_move() (oracle/_ref)."""
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref not present")
    g = dict(np.load(os.path.join(GOLDEN, "fcc6_rho09.npz")))
    dr_max = float(g["dr_max"])
    N = g["conf"].shape[0]
    with oracle_built.Ref(conf=g["conf"], box=g["box"], neigh_dr=float(g["neigh_dr"]), max_part=12) as r, \
            hs.HsmcGpu(N, g["box"][:3], seed=99) as h:
        h.upload(g["conf"])
        for sweep in range(2):
            log = h.sweep_nvt_logged(dr_max)
            keep = log[log["verdict"] != 2]
            cnt = r.replay_moves(keep["id"], keep["raw"], dr_max)
            assert cnt[0] == len(keep)
            assert cnt[1] == int((keep["verdict"] == 0).sum())
            assert cnt[2] == int((keep["verdict"] == 1).sum())
            assert np.array_equal(h.download(), r.get_conf())


def test_logged_and_plain_sweeps_are_the_same_chain(hs):
    g = dict(np.load(os.path.join(GOLDEN, "sc10_rho05.npz")))
    N = g["conf"].shape[0]
    with hs.HsmcGpu(N, g["box"][:3], seed=5) as a, hs.HsmcGpu(N, g["box"][:3], seed=5) as b:
        a.upload(g["conf"])
        b.upload(g["conf"])
        for _ in range(4):
            a.sweep_nvt_logged(0.2)
        b.sweep_nvt(4, 0.2)
        assert np.array_equal(a.download(), b.download())
        assert np.array_equal(a.counters(), b.counters())


def test_sweeps_keep_hard_sphere_invariants(hs, oracle_built):
    """Config-2 shape scaled down: fcc start, rho 0.9, many sweeps: no pair < 1, N conserved,
    acceptance in a sane band, ids a permutation, coordinates inside the closed box."""
    box, conf = oracle_built.Port.lattice(2, 10, 10, 10, 0.9)
    N = conf.shape[0]
    with hs.HsmcGpu(N, box[:3], seed=1) as h:
        h.upload(conf)
        h.sweep_nvt(200, 0.1)
        out = h.download()
        assert np.array_equal(out[:, 0], np.arange(N))
        assert (out[:, 1:] >= 0).all() and (out[:, 1:] <= box[None, :3]).all()
        assert h.min_dist2() >= 1.0
        assert h.overlap_scaled(1.0) == 0
        p = oracle_built.Port(out, box, neigh_dr=1.0, max_part=12)
        assert p.any_overlap(1.0) == 0
        c = h.counters()
        assert c[0] == 200 * N
        assert 0.3 < c[1] / c[0] < 0.9
        assert not np.array_equal(out, conf)
        # particles do cross cell walls thanks to the per-sweep grid shift
        assert np.abs(out[:, 1:] - conf[:, 1:]).max() > 0.5


def test_determinism_and_seed_sensitivity(hs, oracle_built):
    box, conf = oracle_built.Port.lattice(2, 6, 6, 6, 0.7)
    N = conf.shape[0]
    outs = []
    for seed in (42, 42, 43):
        with hs.HsmcGpu(N, box[:3], seed=seed) as h:
            h.upload(conf)
            h.sweep_nvt(25, 0.2)
            outs.append(h.download())
    assert np.array_equal(outs[0], outs[1])
    assert not np.array_equal(outs[0], outs[2])


def test_counters_reset_and_64bit(hs, oracle_built):
    box, conf = oracle_built.Port.lattice(1, 8, 8, 8, 0.4)
    with hs.HsmcGpu(512, box[:3], seed=1) as h:
        h.upload(conf)
        h.sweep_nvt(3, 0.3)
        assert h.counters()[0] == 3 * 512
        h.add_vol_move(True)
        h.add_vol_move(False)
        assert list(h.counters()[3:]) == [2, 1, 1]
        h.reset_counters()
        assert not h.counters().any()
        assert h.counters().dtype == np.int64


@pytest.mark.parametrize("block", [None, "2,3,4", "4,4,8", "8,8,24", "6,12,28"])
def test_kernel_variants_produce_the_same_chain(hs, oracle_built, block, monkeypatch):
    """The default path (proposals generated up front, packed-fp32 filter over the staged block,
    exact re-check) and the all-double global-memory evaluation of the same update order are the
    same Markov chain, bit for bit, on a box large enough to have interior cells, boundary cells,
    wrapped regions and ragged blocks -- and running the default twice gives the same bits.  (The
    one-launch-per-cell-colour kernel, sweep_impl 1, orders the updates differently and is a
    different -- equally valid -- chain.)"""
    impls = (8, 5, 8)
    if block is not None:
        monkeypatch.setenv("HSMC_BLOCK", block)
    box, conf = oracle_built.Port.lattice(2, 14, 9, 11, 0.85)
    N = conf.shape[0]
    outs, cnts = [], []
    for impl in impls:
        with hs.HsmcGpu(N, box[:3], seed=77, sweep_impl=impl) as h:
            h.upload(conf)
            h.sweep_nvt(12, 0.15)
            outs.append(h.download())
            cnts.append(h.counters())
            assert h.min_dist2() >= 1.0
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert np.array_equal(cnts[0], cnts[1]) and np.array_equal(cnts[0], cnts[2])
    assert cnts[0][0] == 12 * N


def test_gather_kernel_is_the_cell_colour_chain(hs, oracle_built):
    """k_sweep_gather (one thread per trial, one launch per cell colour and trial index, fp32 filter over the shadow
    table + exact re-check) against the one-launch-per-colour all-double reference kernel (sweep_impl 1): the same
    Markov chain, bit for bit -- on a box with wrapped columns, boundary cells and cells holding several particles."""
    box, conf = oracle_built.Port.lattice(2, 14, 9, 11, 0.85)
    N = conf.shape[0]
    outs, cnts = [], []
    for impl in (7, 1, 7):
        with hs.HsmcGpu(N, box[:3], seed=77, sweep_impl=impl) as h:
            h.upload(conf)
            h.sweep_nvt(12, 0.15)
            outs.append(h.download())
            cnts.append(h.counters())
            assert h.min_dist2() >= 1.0
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert np.array_equal(cnts[0], cnts[1]) and cnts[0][0] == 12 * N
    # denser cells (cell edge 1.3: up to five particles per cell, the tail launch runs)
    box, conf = oracle_built.Port.lattice(2, 9, 9, 9, 0.95)
    N = conf.shape[0]
    outs = []
    for impl in (7, 1):
        with hs.HsmcGpu(N, box[:3], seed=5, sweep_impl=impl, cell_min=1.3) as h:
            h.upload(conf)
            h.sweep_nvt(10, 0.05)
            outs.append(h.download())
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("block", [None, "2,2,2", "3,2,4", "8,8,24", "10,10,6"])      # the last: 144 staging rows per block
@pytest.mark.parametrize("impl", [8, 5], ids=["staged", "global"])
def test_fused_phase_launch_is_the_same_chain(hs, oracle_built, block, impl, monkeypatch):
    """All eight block phases in ONE launch, ordered by per-block completion flags, against eight
    separate launches (HSMC_FUSE=0): identical coordinates and counters.  Small blocks give
    thousands of CTAs per launch (far more than fit the GPU at once), so the ticket / flag protocol is
    exercised with waiting CTAs; impl 5 reads neighbours straight from global memory, which would
    expose a stale (non-coherent) read of a block finished earlier in the same launch."""
    if block is not None:
        monkeypatch.setenv("HSMC_BLOCK", block)
    box, conf = oracle_built.Port.lattice(2, 26, 22, 24, 0.88)
    N = conf.shape[0]
    outs, cnts = [], []
    for fuse in ("1", "0", "1"):
        monkeypatch.setenv("HSMC_FUSE", fuse)
        with hs.HsmcGpu(N, box[:3], seed=4711, sweep_impl=impl) as h:
            h.upload(conf)
            h.sweep_nvt(15, 0.12)
            outs.append(h.download())
            cnts.append(h.counters())
            assert h.min_dist2() >= 1.0
            launches = h.info()["kernel_launches"]
        outs[-1] = (outs[-1], launches)
    (a, la), (b, lb), (c, lc) = outs
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert np.array_equal(cnts[0], cnts[1]) and cnts[0][0] == 15 * N
    assert lb - la == 15 * 7 and la == lc          # seven launches fewer per sweep


def test_interior_fast_path_replays_through_oracle(hs, oracle_built):
    """A 16^3-cell box has interior cells (no minimum-image branches evaluated on the GPU);
    their verdicts must still equal the reference arithmetic, which always evaluates them."""
    box, conf = oracle_built.Port.lattice(2, 10, 10, 10, 0.9)
    N = conf.shape[0]
    p = oracle_built.Port(conf, box, neigh_dr=1.0, max_part=12)
    with hs.HsmcGpu(N, box[:3], seed=31337) as h:
        assert min(h.info()["cells"]) >= 8
        h.upload(conf)
        for _ in range(4):
            log = h.sweep_nvt_logged(0.12)
            keep = log[log["verdict"] != 2]
            acc = p.replay_moves(keep["id"], keep["raw"], 0.12)
            assert np.array_equal(acc, (keep["verdict"] == 0).astype(np.int32))
            assert np.array_equal(h.download(), p.get_conf())


def _near_contact_system(oracle_built):
    """Simple-cubic crystal with lattice constant 1 + 1e-7 and dr_max = 1e-6: every trial
    lands within ~1e-6 of contact with its six neighbours, far below fp32 resolution, so
    every verdict has to come from the exact double-precision re-evaluation."""
    a = 1.0 + 1e-7
    box, conf = oracle_built.Port.lattice(1, 8, 8, 8, 1.0 / a**3)
    return box, conf, 1e-6


def test_near_contact_trials_take_the_exact_path(hs, oracle_built):
    box, conf, dr = _near_contact_system(oracle_built)
    N = conf.shape[0]
    p = oracle_built.Port(conf, box, neigh_dr=1.0, max_part=12)
    with hs.HsmcGpu(N, box[:3], seed=8) as h:
        h.upload(conf)
        n_acc = n_rej = 0
        for _ in range(6):
            log = h.sweep_nvt_logged(dr)
            keep = log[log["verdict"] != 2]
            acc = p.replay_moves(keep["id"], keep["raw"], dr)
            assert np.array_equal(acc, (keep["verdict"] == 0).astype(np.int32))
            assert np.array_equal(h.download(), p.get_conf())
            n_acc += int((keep["verdict"] == 0).sum()); n_rej += int((keep["verdict"] == 1).sum())
        assert n_acc > 20 and n_rej > 20          # both verdicts, all decided inside the fp32 band
        assert h.min_dist2() >= 1.0


def test_fp32_filter_without_error_band_is_caught(hs, oracle_built):
    """Negative control: with the uncertainty band forced to zero (sweep_impl=3) the fp32
    filter decides near-contact pairs on its own and the replay through the oracle must
    disagree -- i.e. the parity test really is sensitive to the last bits."""
    box, conf, dr = _near_contact_system(oracle_built)
    N = conf.shape[0]
    p = oracle_built.Port(conf, box, neigh_dr=1.0, max_part=12)
    with hs.HsmcGpu(N, box[:3], seed=8, sweep_impl=3) as h:      # (3 = the block-resident kernel with eps = 0)
        h.upload(conf)
        mismatch = False
        for _ in range(8):
            log = h.sweep_nvt_logged(dr)
            keep = log[log["verdict"] != 2]
            acc = p.replay_moves(keep["id"], keep["raw"], dr)
            if not np.array_equal(acc, (keep["verdict"] == 0).astype(np.int32)):
                mismatch = True
                break
        assert mismatch


def test_logged_trial_draws_are_philox_of_cell_trial_and_sweep(hs, oracle_built):
    """The device RNG contract (BASELINE north star: Philox keyed by sweep, colour/cell): the three raw draws
    of every logged trial are Philox4x32-10(counter = (global cell, stream 0 | trial index in the cell, sweep),
    key = seed), recomputed here with the oracle's independent Philox; per sweep every particle has exactly one
    trial, the trials of a cell are numbered 0..n-1 in ascending particle id, and (cell, trial) is unique."""
    g = dict(np.load(os.path.join(GOLDEN, "fcc6_rho09.npz")))
    N = g["conf"].shape[0]
    seed = 0x1234ABCD5
    key = [seed & 0xFFFFFFFF, seed >> 32]
    with hs.HsmcGpu(N, g["box"][:3], seed=seed) as h:
        h.upload(g["conf"])
        for sweep in range(3):
            assert h.info()["sweeps_done"] == sweep
            log = h.sweep_nvt_logged(0.1)
            assert log.shape[0] == N and np.array_equal(np.sort(log["id"]), np.arange(N))
            gcell = (log["seq"] >> np.uint64(8)) & np.uint64((1 << 48) - 1)
            j = (log["seq"] & np.uint64(0xFF)).astype(np.int64)
            assert np.unique(log["seq"] & np.uint64((1 << 56) - 1)).shape[0] == N
            for c in np.unique(gcell):
                sel = np.flatnonzero(gcell == c)
                o = sel[np.argsort(j[sel])]
                assert np.array_equal(j[o], np.arange(o.shape[0])) and np.all(np.diff(log["id"][o]) > 0)
            for k in range(0, N, 7):                          # every 7th trial, all three draws
                r = oracle_built.Port.philox([int(gcell[k]) & 0xFFFFFFFF, int(j[k]), sweep & 0xFFFFFFFF, sweep >> 32], key)
                assert tuple(int(x) for x in r[:3]) == tuple(int(x) for x in log["raw"][k]), (sweep, k)
