#!/usr/bin/env python3
"""Diagnostics for the lean sweep: run the same logged sweeps with sweep_impl 0 and 5 (same chain by construction)
and report the first trials whose verdicts differ."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
import hsmc_b200

cells = [int(x) for x in (sys.argv[1:4] or [14, 9, 11])]
rho = float(sys.argv[4]) if len(sys.argv) > 4 else 0.85
dr = float(sys.argv[5]) if len(sys.argv) > 5 else 0.15
nsweep = int(sys.argv[6]) if len(sys.argv) > 6 else 4
box, conf = bench.fcc_lattice(*cells, rho)
N = conf.shape[0]
hs = [hsmc_b200.HsmcGpu(N, box, seed=77, sweep_impl=i) for i in (0, 5)]
for h in hs:
    h.upload(conf)
print("cells", hs[0].info()["cells"], "N", N)
for s in range(nsweep):
    logs = [h.sweep_nvt_logged(dr) for h in hs]
    outs = [h.download() for h in hs]
    ok = np.array_equal(outs[0], outs[1])
    a, b = (np.sort(l, order="seq") for l in logs)
    same_set = len(a) == len(b) and np.array_equal(a["seq"], b["seq"])
    print(f"sweep {s}: tables equal {ok}; logged {len(a)} / {len(b)}; same (phase,cell,j) set {same_set}")
    if same_set:
        bad = np.flatnonzero((a["verdict"] != b["verdict"]) | (a["id"] != b["id"]) | (a["raw"] != b["raw"]).any(axis=1))
        print("  differing trials:", len(bad))
        for k in bad[:12]:
            seq = int(a["seq"][k])
            print(f"   phase/colour {seq >> 56} gcell {(seq >> 8) & ((1 << 48) - 1)} j {seq & 255}: lean id {a['id'][k]} verdict {a['verdict'][k]} raw {a['raw'][k]} | ref id {b['id'][k]} verdict {b['verdict'][k]} raw {b['raw'][k]}")
    else:
        sa, sb = set(a["seq"].tolist()), set(b["seq"].tolist())
        print("  only lean:", len(sa - sb), "only ref:", len(sb - sa), "dup lean:", len(a) - len(sa))
        for q in sorted(sb - sa)[:8]:
            print(f"   missing in lean: phase/colour {q >> 56} gcell {(q >> 8) & ((1 << 48) - 1)} j {q & 255}")
    if not ok:
        d = np.flatnonzero((outs[0] != outs[1]).any(axis=1))
        print("  rows differing:", len(d), d[:10])
        break
