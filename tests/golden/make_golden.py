"""Generate the golden fixtures under tests/golden/ FROM THE UNMODIFIED REFERENCE.

Run in the build container (needs /root/reference and oracle/_ref/libhsmc_ref.so, built by
`make -C oracle ref`).  Each fixture is a small .npz holding an equilibrated configuration
produced by the reference's own sweep_nvt() and the outputs of the reference's own hot-path
routines on it:

  conf, box                      restart-precision configuration (full doubles)
  trial_idx/xyz/flags[_sf]       check_overlap() verdicts for explicit trial points
                                 (uniform (u-0.5)*dr_max cubes + adversarial |r-1| <= 4 ulp)
  widom_raw, widom_flags         widom_rand_pos()/widom_check_overlap() per raw draw
  rdf_hist, pressv_hist          rdf_hist_compute() / pressv_compute_hist() (un-normalised)
  presst_hist, presst_xi         presst_compute_hist()
  replay_*                       part_move() driven by scripted draws: counters + final conf

The reference has no tests or golden vectors of its own; these files
are outputs of the reference run here, which is what pins the oracle.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (type, nx, ny, nz, rho, neigh_dr, dr_max, sweeps, seed)
    "fcc5_rho08": (2, 5, 5, 5, 0.8, 1.0, 0.1, 200, 4357),
    "sc10_rho05": (1, 10, 10, 10, 0.5, 1.05, 0.2, 100, 124787),
    "fcc6_rho09": (2, 6, 6, 6, 0.9, 1.05, 0.08, 150, 99),
    "fcc8_rho094": (2, 8, 8, 8, 0.94, 1.02, 0.05, 60, 7),
    "fcc5_rho03": (2, 5, 5, 5, 0.3, 1.1, 0.5, 100, 31),
}


def adversarial_points(conf, box, rng, n):
    """Trial points at distance 1 +- k ulp from an existing particle (k in -4..4)."""
    N = conf.shape[0]
    idx = rng.integers(0, N, n).astype(np.int32)
    other = (idx + 1 + rng.integers(0, N - 1, n)) % N
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    k = rng.integers(-4, 5, n)
    r = 1.0 + k * np.finfo(np.float64).eps
    xyz = conf[other, 1:] + d * r[:, None]
    xyz = np.where(xyz > box[None, :3], xyz - box[None, :3], xyz)
    xyz = np.where(xyz < 0.0, xyz + box[None, :3], xyz)
    return idx, xyz


def main():
    for name, (typ, nx, ny, nz, rho, ndr, dr_max, sweeps, seed) in CASES.items():
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        r = Ref(lattice=(typ, nx, ny, nz, rho), neigh_dr=ndr, max_part=12, seed=seed)
        r.set_moves(dr_max=dr_max)
        r.sweep_nvt(sweeps)
        conf = r.get_conf()
        box = r.box4()
        N = r.N
        out = {"conf": conf, "box": box, "neigh_dr": ndr, "dr_max": dr_max, "cells": r.cells()[0]}

        # explicit trial points
        n = 3000
        idx = rng.integers(0, N, n).astype(np.int32)
        xyz = conf[idx, 1:] + (rng.random((n, 3)) - 0.5) * (4 * dr_max)
        xyz = np.where(xyz > box[None, :3], xyz - box[None, :3], xyz)
        xyz = np.where(xyz < 0.0, xyz + box[None, :3], xyz)
        ai, axyz = adversarial_points(conf, box, rng, 1000)
        idx = np.concatenate([idx, ai])
        xyz = np.concatenate([xyz, axyz])
        out["trial_idx"], out["trial_xyz"] = idx, xyz
        out["trial_flags"] = r.trial_verdicts(idx, xyz, 1.0)
        cell_min = float(min(r.cells()[1]))
        sf = 0.9995 if cell_min * 0.9995 >= 1.0 else 1.0 - 0.5 * (1.0 - 1.0 / cell_min)
        out["sf"] = sf
        out["trial_flags_sf"] = r.trial_verdicts(idx, xyz, sf)
        out["overlap_all_sf"] = r.overlap_all(sf)
        out["overlap_all_1"] = r.overlap_all(1.0)

        # widom
        raw = rng.integers(0, 2**32, (4000, 3), dtype=np.uint64).astype(np.uint32)
        out["widom_raw"] = raw
        out["widom_flags"] = r.widom_verdicts_raw(raw)
        assert r.widom_count_raw(raw) == int((out["widom_flags"] == 0).sum())

        # histograms
        out["rdf_dr"], out["rdf_rmax"] = 0.01, 10.0
        # compute_rdf clamps rmax to half the box (compute_rdf.c:39-52) before allocating
        rmax = min(10.0, box[0] / 2.0)
        out["rdf_rmax_eff"] = rmax
        out["rdf_hist"] = r.rdf_hist(0.01, rmax)
        if cell_min >= 1.05:
            out["pressv_dr"] = 0.002
            out["pressv_hist"] = r.pressv_hist(0.002)
        xi_max = 0.002 if cell_min * (1 - 0.002) ** (1 / 3.0) >= 1.0 else 0.0
        if xi_max > 0:
            h, xi = r.presst_hist(0.0001, xi_max)
            out["presst_hist"], out["presst_xi"] = h, xi

        # scripted part_move replay: 2 sweeps' worth of explicit draws
        m = 2 * N
        ids = rng.integers(0, N, m)
        draws = rng.integers(0, 2**32, (m, 3), dtype=np.uint64).astype(np.uint32)
        cnt = r.replay_moves(ids, draws, dr_max)
        out["replay_ids"], out["replay_raw"], out["replay_counters"] = ids.astype(np.int32), draws, cnt
        out["replay_conf"] = r.get_conf()
        r.close()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, N, box[:3], "cells", out["cells"], "acc", cnt[1], "/", cnt[0],
              "widom ok", int((out["widom_flags"] == 0).sum()), "rdf pairs", out["rdf_hist"].sum() / 2)


if __name__ == "__main__":
    main()
