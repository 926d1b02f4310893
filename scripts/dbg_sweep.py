"""Timing ablations of the block kernel (HSMC_BLOCK_DBG): prints us per sweep phase."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import fcc_lattice
import hsmc_b200
box, conf = fcc_lattice(256, 128, 128, 0.9)
N = conf.shape[0]
for dbg in sys.argv[1:]:
    os.environ["HSMC_BLOCK_DBG"] = dbg
    with hsmc_b200.HsmcGpu(N, box, seed=1) as h:
        h.upload(conf)
        if dbg == "0":
            h.sweep_nvt(3, 0.1)
        h.sync(); h.profile(True); h.profile_read()
        h.sweep_nvt(3, 0.1)
        pr = h.profile_read()
        if dbg == "10":
            import ctypes as C
            out = (C.c_uint64 * 16)()
            h.L.hsmc_gpu_debug_block_cycles.argtypes = [C.c_void_p, C.c_void_p]
            h.L.hsmc_gpu_debug_block_cycles(h.h, out)
            names = ["row ends", "row scan", "issue copies + CSR", "wait data", "convert + census", "trial slots", "own chunks residual", "colour barrier wait", "  ticket + stage_a(next)", "  master + trial + hide", "  scan", "  mates + verdict + commit", "", "", "  loop top"]
            tot = sum(out[:16])
            for k, nm in enumerate(names):
                if nm: print(f"   {nm:28s} {out[k] / 24 / 3024:10.0f} cycles/CTA  {100.0 * out[k] / tot:5.1f} %")
        print(f"dbg={dbg}: phase {1e3 * pr['sweep'][0] / pr['sweep'][1]:.1f} us, build {1e3 * pr['build'][0] / max(pr['build'][1],1):.1f} us", flush=True)
