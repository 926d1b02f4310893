"""Full-size properties (BASELINE configs[3] shape: fcc 256x128x128, N = 16 777 216, rho = 0.9).

The oracle cannot replay 16.8M-particle sweeps in seconds, so at this size the checks are the
size-independent ones: the chain does not depend on how it is launched (eight block phases
in one launch vs eight launches -- bitwise), particle identities are conserved, no pair is
closer than sigma afterwards, the counters add up, and sharded observables sum to the whole.
Bit-exact replay against the reference arithmetic is done at small sizes in test_gpu_sweep.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CELLS = (256, 128, 128)
RHO, DR_MAX, SWEEPS = 0.9, 0.1, 20


@pytest.fixture(scope="module")
def big(lib_built):
    import hsmc_b200
    if hsmc_b200.load_library().hsmc_gpu_device_count() < 1:
        pytest.fail("no CUDA device visible: GPU tests cannot run (there is no CPU fallback)")
    from bench import fcc_lattice
    box, conf = fcc_lattice(*CELLS, RHO)
    return hsmc_b200, box, conf


def _run(hs, box, conf, monkeypatch, fuse):
    monkeypatch.setenv("HSMC_FUSE", fuse)
    N = conf.shape[0]
    with hs.HsmcGpu(N, box, seed=99) as h:
        h.upload(conf)
        h.sweep_nvt(SWEEPS, DR_MAX)
        out = h.download()
        extra = {
            "counters": h.counters(), "min_r2": h.min_dist2(), "cell_rejects": h.cell_rejects(),
            "launches": h.info()["kernel_launches"],
        }
        if fuse == "1":
            M = 4_000_000
            whole = h.widom(11, M)
            parts = [h.widom(11, M // 4, first=k * (M // 4), reduce=False) for k in range(4)]
            extra["widom"] = (whole, parts)
            extra["overlap"] = (h.overlap_scaled(1.0), h.overlap_scaled(0.999))
    return out, extra


def test_full_size_chain_properties(big, monkeypatch):
    hs, box, conf = big
    N = conf.shape[0]
    a, ea = _run(hs, box, conf, monkeypatch, "1")
    b, eb = _run(hs, box, conf, monkeypatch, "0")
    # same chain whichever way the phases are launched
    assert np.array_equal(a, b)
    assert np.array_equal(ea["counters"], eb["counters"]) and ea["min_r2"] == eb["min_r2"]
    assert eb["launches"] - ea["launches"] == 7 * SWEEPS
    # identities conserved, row i is particle i (the reference's host layout), everything inside the box
    assert np.array_equal(a[:, 0], np.arange(N, dtype=np.float64))
    assert (a[:, 1:] >= 0.0).all() and (a[:, 1:] <= np.asarray(box)).all()
    # hard-sphere invariant and bookkeeping
    assert ea["min_r2"] >= 1.0
    c = ea["counters"]
    assert c[0] == SWEEPS * N and c[1] + c[2] == c[0] and 0 < ea["cell_rejects"] < c[2]
    moved = (a[:, 1:] != conf[:, 1:]).any(axis=1).sum()
    assert 0.3 * N < moved <= c[1]                   # accepted moves move particles; a particle can move twice
    # sharded Widom ranges (BASELINE config 5) add up to the unsharded count
    whole, parts = ea["widom"]
    assert sum(parts) == whole and 0 <= whole < 4_000_000
    # no overlap as is; compressing a rho = 0.9 fluid/crystal by 0.1 % in length must create one
    assert ea["overlap"] == (0, 1)
