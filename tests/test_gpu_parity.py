"""GPU parity tests (run with -m gpu on the B200 box), all through the C ABI.

Checker: the CPU oracle (oracle/liboracle_port.so, validated against the unmodified
reference in test_oracle_cpu.py) and the golden fixtures produced by the reference itself.
Bar: bit-exact -- overlap verdicts, accept/reject decisions, Widom counts and integer
histogram counts are integers; coordinates are compared as raw doubles."""
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


def _gpu(hs, g, **kw):
    h = hs.HsmcGpu(g["conf"].shape[0], g["box"][:3], **kw)
    h.upload(g["conf"])
    return h


def _philox_widom_raw(oracle, seed, sample, first, count):
    out = np.zeros((count, 3), dtype=np.uint32)
    key = [seed & 0xFFFFFFFF, seed >> 32]
    for i in range(count):
        m = first + i
        r = oracle.Port.philox([m & 0xFFFFFFFF, (1 << 24) | (m >> 32), sample & 0xFFFFFFFF, sample >> 32], key)
        out[i] = r[:3]
    return out


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=IDS)
def test_upload_download_roundtrip(hs, path):
    g = dict(np.load(path))
    with _gpu(hs, g) as h:
        back = h.download()
        assert np.array_equal(back, g["conf"])
        info = h.info()
        assert all(c % 2 == 0 and c >= 4 for c in info["cells"])
        assert all(s >= 1.0 for s in info["cell_size"])


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=IDS)
def test_trial_verdicts_match_reference_golden(hs, path):
    g = dict(np.load(path))
    with _gpu(hs, g) as h:
        f = h.trial_verdicts(g["trial_idx"], g["trial_xyz"], 1.0)
        assert np.array_equal(f, g["trial_flags"])
        fs = h.trial_verdicts(g["trial_idx"], g["trial_xyz"], float(g["sf"]))
        assert np.array_equal(fs, g["trial_flags_sf"])


def test_adversarial_tail_exercises_both_verdicts():
    """The last 1000 trial points of each fixture sit at |r-1| <= 4 ulp from a particle;
    across the fixtures (dilute ones in particular) both verdicts must occur there."""
    tails = np.concatenate([np.load(p)["trial_flags"][-1000:] for p in GOLDEN_FILES])
    assert 0 < tails.sum() < tails.size
    dilute = np.load(os.path.join(GOLDEN, "fcc5_rho03.npz"))["trial_flags"][-1000:]
    assert 100 < dilute.sum() < 900


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=IDS)
def test_global_overlap_verdict(hs, path, oracle_built):
    g = dict(np.load(path))
    with _gpu(hs, g) as h:
        assert h.overlap_scaled(1.0) == int(g["overlap_all_1"].any()) == 0
        assert h.overlap_scaled(float(g["sf"])) == int(g["overlap_all_sf"].any())
        if "presst_hist" in g:
            xi = g["presst_xi"]
            sf = np.array([pow(1 - x, 1.0 / 3.0) for x in xi])
            p = oracle_built.Port(g["conf"], g["box"], neigh_dr=float(g["neigh_dr"]), max_part=12)
            _, sf_o = p.presst_flags(0.0001, 0.002)
            assert np.array_equal(sf, sf_o)
            f = h.presst_flags(sf)
            assert np.array_equal(f.astype(float), g["presst_hist"])
        assert h.min_dist2() >= 1.0


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=IDS)
def test_accepted_volume_move_rescales_like_the_reference(hs, path, oracle_built):
    """K7 `hsmc_gpu_rescale` = the accepted branch of vol_move (moves.c:135-142): every coordinate times sf,
    new box, cell list rebuilt.  Coordinates bit for bit against the oracle's rescale (itself checked against
    the unmodified reference, test_oracle_cpu.py), for an expansion, a compression and a factor that changes the
    number of cells per axis; the rebuilt cell list must give the oracle's verdicts afterwards."""
    g = dict(np.load(path))
    box = np.array(g["box"][:3], dtype=np.float64)
    N = g["conf"].shape[0]
    sf_big = (np.floor(box[0]) + 2.5) / box[0]          # enough to change the number of (even) cells per axis
    for sf in (1.0 + 3.0e-4, pow(1.0 - 1.0e-3, 1.0 / 3.0), sf_big):
        p = oracle_built.Port(g["conf"], g["box"], neigh_dr=1.0, max_part=16)
        new_box = np.array([box[0] * sf, box[1] * sf, box[2] * sf])      # moves.c:129-131 order: edge * sf
        with _gpu(hs, g) as h:
            cells0 = h.info()["cells"]
            h.rescale(sf, new_box)
            p.rescale(sf, new_box)
            out = h.download()
            assert np.array_equal(out, p.get_conf()), f"sf={sf}"
            assert np.allclose(h.info()["box"], new_box, rtol=0, atol=0)
            if sf == sf_big:
                assert h.info()["cells"] != cells0
            # the rebuilt device cell list answers like the oracle on the rescaled configuration
            rng = np.random.default_rng(5)
            idx = rng.integers(0, N, 2000).astype(np.int32)
            xyz = (out[idx, 1:] + (rng.random((2000, 3)) - 0.5) * 0.4) % new_box[None, :]
            q = oracle_built.Port(out, new_box, neigh_dr=1.0, max_part=16)
            assert np.array_equal(h.trial_verdicts(idx, xyz), q.trial_verdicts(idx, xyz))
            assert h.overlap_scaled(1.0) == q.any_overlap(1.0)
            # and sweeps go on from there as the same chain the oracle replays
            log = h.sweep_nvt_logged(0.1)
            keep = log[log["verdict"] != 2]
            acc = q.replay_moves(keep["id"], keep["raw"], 0.1)
            assert np.array_equal(acc, (keep["verdict"] == 0).astype(np.int32))
            assert np.array_equal(h.download(), q.get_conf())


def test_compression_below_the_cell_edge_is_evaluated_not_refused(hs, oracle_built):
    """with the reference default `neigh_list 1.0` the cell edge L/n can sit anywhere above 1.0, and a
    volume-move or press_thermo compression with cell*sf < 1 used to be refused (the host driver died).  The
    verdict is now evaluated over a two-cell stencil; checked against brute-force all-pairs arithmetic in the
    reference's operation order (the reference's own 27-cell scan misses such pairs, moves.c:108)."""
    box, conf = oracle_built.Port.lattice(2, 6, 6, 6, 0.9)
    L = float(box[0])
    N = conf.shape[0]
    with hs.HsmcGpu(N, box[:3], seed=11, cell_min=1.0) as h:
        h.upload(conf)
        h.sweep_nvt(30, 0.1)
        out = h.download()
        w = min(h.info()["cell_size"])
        # smallest pair distance decides which compressions overlap
        d = out[:, None, 1:] - out[None, :, 1:]
        d -= L * np.round(d / L)
        r = np.sqrt((d ** 2).sum(-1))
        r[np.arange(N), np.arange(N)] = 9.0
        rmin = r.min()
        for sf in (0.999 / w, 0.97 / w, 0.8):
            assert w * sf < 1.0
            assert abs(rmin * sf - 1.0) > 1e-9          # (no verdict within rounding of the threshold)
            assert h.overlap_scaled(sf) == int(rmin * sf < 1.0)
        sfs = np.array([1.0, 0.9999, 0.999 / w, 0.9 / w])
        f = h.presst_flags(sfs)
        assert list(f) == [int(not (rmin * s < 1.0)) for s in sfs]


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=IDS)
def test_widom_verdicts_and_counts(hs, path, oracle_built):
    g = dict(np.load(path))
    box = g["box"]
    with _gpu(hs, g, seed=0xC0FFEE1234) as h:
        raw = g["widom_raw"]
        xyz = (raw.astype(np.float64) / 4294967295.0) * box[None, :3]
        assert np.array_equal(h.widom_verdicts(xyz), g["widom_flags"])
        # device-generated points: regenerate the same Philox draws on the host and count
        # with the oracle
        M = 20000
        n_gpu = h.widom(sample_id=7, count=M)
        raw_d = _philox_widom_raw(oracle_built, 0xC0FFEE1234, 7, 0, M)
        p = oracle_built.Port(g["conf"], box, neigh_dr=float(g["neigh_dr"]), max_part=12)
        assert n_gpu == p.widom_count_raw(raw_d)
        # range splitting is additive (multi-GPU sharding of insertions)
        assert h.widom(7, 5000, first=0) + h.widom(7, M - 5000, first=5000) == n_gpu


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=IDS)
def test_rdf_counts(hs, path):
    g = dict(np.load(path))
    with _gpu(hs, g) as h:
        dr = float(g["rdf_dr"])
        nn = int((float(g["rdf_rmax_eff"]) - 1.0) / dr)
        c = h.rdf_counts(dr, nn)
        assert np.array_equal(2.0 * c.astype(np.float64), g["rdf_hist"])


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=IDS)
def test_contact_counts(hs, path, oracle_built):
    g = dict(np.load(path))
    with _gpu(hs, g) as h:
        info = h.info()
        if min(info["cell_size"]) < 1.05:
            with pytest.raises(hs.HsmcError, match="size of the cells"):
                h.contact_counts(0.002, 25)
            # a narrower histogram that fits the cells still has to agree with the oracle
            nn = int((min(info["cell_size"]) - 1.0) / 0.002)
            if nn < 1:
                return
        else:
            nn = int((1.05 - 1.0) / 0.002)
            if "pressv_hist" in g:
                assert np.array_equal(2.0 * h.contact_counts(0.002, nn).astype(float), g["pressv_hist"])
        # oracle all-pairs histogram restricted to the first nn bins is the same quantity
        p = oracle_built.Port(g["conf"], g["box"], neigh_dr=float(g["neigh_dr"]), max_part=12)
        full = p.rdf_counts(0.002, 1.0 + 0.002 * nn)
        assert np.array_equal(h.contact_counts(0.002, nn), full[:nn])


@pytest.mark.parametrize("cells,rho,cell_min", [((12, 10, 8), 0.9, 1.0), ((9, 9, 9), 0.6, 1.3), ((24, 24, 24), 0.94, 1.0)])
def test_prefiltered_pair_kernels_equal_the_all_double_ones(hs, monkeypatch, cells, rho, cell_min):
    """K3 / K6 walk the fp32 shadow table first (thread per particle) and evaluate in double only what the filter
    lets through; HSMC_OBS_DOUBLE=1 selects the all-double thread-per-cell kernels.  Same verdicts for a ladder of
    compressions that straddles the closest pair, same contact histogram, on configurations evolved on shifted
    grids (offsets of particles in the cell that straddles the periodic edge included)."""
    import bench
    box, conf = bench.fcc_lattice(*cells, rho)
    with hs.HsmcGpu(conf.shape[0], box, seed=5, cell_min=cell_min) as h:
        h.upload(conf)
        for rounds in range(3):
            h.sweep_nvt(7, 0.1)
            rmin = float(np.sqrt(h.min_dist2()))
            w = min(h.info()["cell_size"])
            lo = max(1.0 / w, 0.98 / rmin) + 1e-9
            sfs = np.concatenate([np.linspace(lo, 1.0 / rmin, 9), [np.nextafter(1.0 / rmin, 0), np.nextafter(1.0 / rmin, 2), 1.0]])
            nn, dr = 20, min(0.002, 0.9 * (w - 1.0) / 20)
            res = {}
            for mode in ("0", "1"):
                monkeypatch.setenv("HSMC_OBS_DOUBLE", mode)
                res[mode] = (h.presst_flags(sfs), [h.overlap_scaled(float(x)) for x in sfs[-4:]], h.contact_counts(dr, nn))
            assert np.array_equal(res["0"][0], res["1"][0]) and res["0"][1] == res["1"][1]
            assert np.array_equal(res["0"][2], res["1"][2])
            assert res["0"][0].any() and not res["0"][0].all()          # the ladder straddles the closest pair
            assert res["0"][2].sum() > 0


def test_random_fluid_against_oracle(hs, oracle_built):
    """Seeded non-lattice input: dilute random configuration, larger box, ragged cells."""
    rng = np.random.default_rng(11)
    L = 14.3
    pts = []
    while len(pts) < 700:
        c = rng.random(3) * L
        if all(np.linalg.norm((c - q + L / 2) % L - L / 2) >= 1.0 for q in pts):
            pts.append(c)
    conf = oracle_built.conf_from_xyz(np.array(pts))
    box = [L, L, L]
    p = oracle_built.Port(conf, box, neigh_dr=1.0, max_part=12)
    with hs.HsmcGpu(700, box, seed=3) as h:
        h.upload(conf)
        idx = rng.integers(0, 700, 5000).astype(np.int32)
        xyz = (conf[idx, 1:] + (rng.random((5000, 3)) - 0.5) * 0.8) % L
        assert np.array_equal(h.trial_verdicts(idx, xyz), p.trial_verdicts(idx, xyz))
        w = rng.random((5000, 3)) * L
        assert np.array_equal(h.widom_verdicts(w), p.widom_verdicts(w))
        nn = int((L / 2 - 1.0) / 0.05)
        assert np.array_equal(h.rdf_counts(0.05, nn), p.rdf_counts(0.05, L / 2))


def test_errors_follow_reference_convention(hs):
    with pytest.raises(hs.HsmcError, match="too small"):
        hs.HsmcGpu(10, [3.0, 3.0, 3.0])
    with hs.HsmcGpu(500, [8.5, 8.5, 8.5]) as h:
        with pytest.raises(hs.HsmcError, match="no configuration"):
            h.sweep_nvt(1, 0.1)
        with pytest.raises(hs.HsmcError, match="n_rows"):
            h.upload(np.zeros((10, 4)))


def test_division_free_u01_is_exact_for_all_raw_values(hs):
    """u = raw/0xffffffff (rng.c:29-31) evaluated without a division must equal the IEEE
    division for every one of the 2^32 possible draws."""
    with hs.HsmcGpu(500, [8.5, 8.5, 8.5]) as h:
        n_bad, first = h.selftest_u01()
        assert n_bad == 0, f"{n_bad} mismatches, first at raw={first}"


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=IDS)
def test_order_parameter_matches_reference_and_oracle(hs, path, oracle_built):
    """q_l (compute_order_parameter.c:84-229): GPU value against (i) the committed outputs of the
    unmodified reference (tests/golden/ql/ql_ref.json, made by tests/golden/make_ql_golden.py) and
    (ii) the oracle restatement on a configuration the GPU evolved itself.  Floating point:
    1e-12 relative (different but equivalent summation / recurrence order)."""
    import json
    ref = json.load(open(os.path.join(GOLDEN, "ql", "ql_ref.json")))[os.path.basename(path)[:-4]]
    g = dict(np.load(path))
    N = g["conf"].shape[0]
    with hs.HsmcGpu(N, g["box"][:3], seed=3, cell_min=float(g["neigh_dr"])) as h:
        h.upload(g["conf"])
        edge = min(h.info()["cell_size"])
        assert ref["rmax"] <= edge * (1 + 1e-12)
        for l in (4, 6):
            assert h.order_parameter(l, ref["rmax"]) == pytest.approx(ref["ql"][str(l)], rel=1e-12, abs=1e-13)
        h.sweep_nvt(5, float(g["dr_max"]))
        out = h.download()
        p = oracle_built.Port(out, g["box"], neigh_dr=float(g["neigh_dr"]), max_part=12)
        for l in (2, 6, 12):
            assert h.order_parameter(l, ref["rmax"]) == pytest.approx(p.order_param(l, ref["rmax"]), rel=1e-12, abs=1e-13)
        with pytest.raises(hs.HsmcError):
            h.order_parameter(6, edge * 1.01)          # cutoff beyond the 27-cell stencil
        with pytest.raises(hs.HsmcError):
            h.order_parameter(13, ref["rmax"])
