#!/usr/bin/env python3
"""Quick A/B timing of the sweep path at the benchmark size (one process per configuration: the library is
chosen by HSMC_GPU_LIB, the block shape by HSMC_BLOCK).  Prints ms per sweep by profile bucket.
  python scripts/lean_bench.py [--cells 256 128 128] [--sweeps 20] [--impl 0] [--check]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, nargs=3, default=[256, 128, 128])
ap.add_argument("--sweeps", type=int, default=20)
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--rho", type=float, default=0.9)
ap.add_argument("--dr", type=float, default=0.1)
ap.add_argument("--check", action="store_true", help="compare the final table with sweep_impl 5 (all-double global path)")
a = ap.parse_args()
import bench
import hsmc_b200
nx, ny, nz = a.cells
N = 4 * nx * ny * nz
box, conf = bench.fcc_lattice(nx, ny, nz, a.rho)
outs = []
for impl in ([a.impl, 5] if a.check else [a.impl]):
    h = hsmc_b200.HsmcGpu(N, box, seed=20261017, device=0, sweep_impl=impl)
    h.upload(conf)
    h.sweep_nvt(3, a.dr)
    h.sync()
    h.profile(True); h.profile_read()
    t0 = time.perf_counter()
    h.sweep_nvt(a.sweeps, a.dr)
    h.sync()
    dt = time.perf_counter() - t0
    p = h.profile_read()
    c = h.counters()
    print(f"impl {impl} lib {os.path.basename(os.environ.get('HSMC_GPU_LIB', 'default'))} block {os.environ.get('HSMC_BLOCK', '-')}: "
          f"{N * a.sweeps / dt:.3e} moves/s wall | per sweep ms: " +
          " ".join(f"{k} {v[0] / a.sweeps:.3f}" for k, v in p.items()) + f" | acc {c[1] / c[0]:.3f} min_r2 {h.min_dist2():.6f}", flush=True)
    try:
        import ctypes
        raw = (ctypes.c_uint64 * 8)()
        h.L.hsmc_gpu_debug_counters.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        h.L.hsmc_gpu_debug_counters(h.h, raw)
        if raw[4] or raw[5] or raw[6]:
            print(f"   blocks off the staged path: capacity {raw[4]} deep cell {raw[5]} trial lists {raw[6]}", flush=True)
    except Exception as e:
        print("   (no debug counters:", e, ")")
    try:
        st = (ctypes.c_uint64 * 32)()
        h.L.hsmc_gpu_debug_stamps.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        h.L.hsmc_gpu_debug_stamps(h.h, st)
        if st[31]:
            names = ["geometry", "rows+flag wait", "row scan+chunks", "cell index+staging"] + [f"colour {c}" for c in range(8)] + ["commit", "counters+publish"]
            tot = sum(st[i] for i in range(14))
            print(f"   cycles per block (thread 0's clock, {st[31]} blocks, {tot / st[31]:.0f} total): " +
                  ", ".join(f"{n} {st[i] / st[31]:.0f}" for i, n in enumerate(names)) +
                  f" | warp 0 inside staging: cell index pass {st[14] / st[31]:.0f}, its rows {st[15] / st[31]:.0f}", flush=True)
    except Exception as e:
        print("   (no stamps:", e, ")")
    if a.check:
        outs.append(h.download())
    h.close()
if a.check:
    print("identical to the all-double chain:", bool(np.array_equal(outs[0], outs[1])), flush=True)
