#!/usr/bin/env python3
"""One invocation of every kernel of the hot path, in a fixed order, for an ncu capture (profiles/):
  ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_all python scripts/profile_all.py
At the benchmark size (cubic fcc 162^3, N = 17 006 112) except RDF (C2 shape, N = 32 000: O(N^2)) and q_l (fcc 64^3 on
cells of edge 1.5, which hold the first neighbour shell)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench, hsmc_b200

cells = [int(x) for x in (sys.argv[1:4] or [162, 162, 162])]
box, conf = bench.fcc_lattice(*cells, 0.9)
N = conf.shape[0]
with hsmc_b200.HsmcGpu(N, box, seed=20261017) as h:
    h.upload(conf)                                   # k_unpack_rows, k_cell_count, k_scan_*, k_cell_scatter
    h.sweep_nvt(2, 0.1)                              # per sweep: rebuild kernels, k_propose, k_sweep_lean
    print("widom", h.widom(7, 20_000_000))           # k_widom
    print("overlap", h.overlap_scaled(1.0))          # k_overlap_scaled (one factor)
    sf = (1.0 - 0.0001 * (np.arange(20) + 1.0)) ** (1.0 / 3.0)
    print("presst", h.presst_flags(sf).sum())        # k_overlap_scaled (20 factors)
    dr_c = min(0.002, 0.9 * (min(h.info()["cell_size"]) - 1.0))
    print("contact", h.contact_counts(dr_c, 1))      # k_contact_hist
    print("min r2", h.min_dist2())                   # k_min_r2
    s = 1.0 + 2.0e-5
    h.rescale(s, [b * s for b in box])               # k_rescale + rebuild
    out = h.download()                               # k_pack_by_id
box2, conf2 = bench.fcc_lattice(20, 20, 20, 0.9)
with hsmc_b200.HsmcGpu(conf2.shape[0], box2, seed=3) as h2:
    h2.upload(conf2)
    h2.sweep_nvt(20, 0.1)
    print("rdf", h2.rdf_counts(0.01, 400)[:3])       # k_rdf_pairs
box3, conf3 = bench.fcc_lattice(64, 64, 64, 0.9)
with hsmc_b200.HsmcGpu(conf3.shape[0], box3, seed=4, cell_min=1.5) as h3:
    h3.upload(conf3)
    h3.sweep_nvt(2, 0.1)
    print("q6", h3.order_parameter(6, min(1.5, min(h3.info()["cell_size"]))))   # k_order_param, k_sum_partials
