"""Run the UNMODIFIED reference executable (oracle/_ref/hsmc_ref) on the statistical-test
inputs tests/golden/stat/*.in and store the per-sample observable series it writes
(press_virial.dat, press_thermo.dat, chem_pot.dat, rdf, order_param.dat, density.dat) as
compact .npz fixtures.  The GPU statistical tests run this repo's host driver on the same
inputs and require agreement within 3 sigma (blocking analysis, tests/blocking.py)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hsmc_outputs import collect  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "hsmc_ref")


def main():
    for name in sorted(os.listdir(os.path.join(HERE, "stat"))):
        if not name.endswith(".in"):
            continue
        case = name[:-3]
        pre = os.path.join("/tmp/refrun", case)
        if os.path.exists(os.path.join(pre, "out.txt")) and "Simulation complete" in open(os.path.join(pre, "out.txt")).read():
            d = pre
        else:
            d = tempfile.mkdtemp(prefix="hsmc_ref_")
            subprocess.run([EXE, "-i", os.path.join(HERE, "stat", name), "-o", "out.txt"], cwd=d, check=True)
        obs = collect(d)
        keep = {}
        for k, v in obs.items():
            if k == "rdf_g":           # keep the sample mean and a per-sample series at the first peak region
                keep["rdf_g_mean"] = v.mean(axis=0)
                keep["rdf_g_samples_first8"] = v[:, :8]
            elif k == "pressv_g":
                continue
            else:
                keep[k] = v
        np.savez_compressed(os.path.join(HERE, "stat", case + "_ref.npz"), **keep)
        print(case, {k: getattr(v, "shape", None) for k, v in keep.items()})
        for k in ("g_contact", "widom_frac", "ql", "density"):
            if k in keep:
                print("   ", k, "mean", keep[k].mean())


if __name__ == "__main__":
    main()
