"""Golden values of the order parameter from the UNMODIFIED reference (oracle/_ref), run in the
build container: global_ql_compute() (compute_order_parameter.c:84-97) on the committed golden
configurations, l = 4 and 6, bond cutoff = the smallest neighbour-list cell edge.

    python tests/golden/make_ql_golden.py        # rewrites tests/golden/ql/ql_ref.json
"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

out = {}
for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
    g = dict(np.load(f))
    nd = float(g["neigh_dr"])
    with pyoracle.Ref(conf=g["conf"], box=g["box"], neigh_dr=nd, max_part=12) as r:
        _, size = r.cells()
        rmax = float(min(size))
        out[os.path.basename(f)[:-4]] = {"rmax": rmax, "ql": {str(l): r.order_param(l, rmax) for l in (4, 6)}}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ql", "ql_ref.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
