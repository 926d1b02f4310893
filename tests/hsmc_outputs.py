"""Parsers for the reference's output files (the drop-in contract: the host driver of this
repo writes the same formats).  Formats: compute_press.c:277-359, compute_widom_chem_pot.c:164-182,
compute_rdf.c:155-205, compute_order_parameter.c:232-253."""
import gzip
import os

import numpy as np


def _blocks(lines):
    """Samples of the '7-line header + nn lines' files (press_virial, press_thermo, rdf)."""
    out, i = [], 0
    while i < len(lines):
        assert lines[i].startswith("#"), lines[i]
        hdr = lines[i + 3].split()
        nn = int(hdr[-3]) if len(hdr) == 4 else int(hdr[0])
        vals = np.array([[float(x) for x in ln.split()] for ln in lines[i + 7:i + 7 + nn]])
        out.append((hdr, vals))
        i += 7 + nn
    return out


def read_hist_file(path):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip()]
    blocks = _blocks(lines)
    x = blocks[0][1][:, 0]
    y = np.array([b[1][:, 1] for b in blocks])
    return x, y, [b[0] for b in blocks]


def read_column_file(path):
    with open(path) as f:
        rows = [[float(x) for x in ln.split()] for ln in f if ln.strip() and not ln.startswith("#")]
    return np.array(rows)


def contact_value(rr, g_samples):
    """g(1+) per sample by the linear extrapolation of the near-contact histogram
    (hsmc_pressure.py:11-41 fits a line to the bins and evaluates it at r = 1)."""
    A = np.vstack([rr - 1.0, np.ones_like(rr)]).T
    coef, *_ = np.linalg.lstsq(A, g_samples.T, rcond=None)
    return coef[1]


def virial_pressure(rho, g_contact):
    """beta P / rho = 1 + (2 pi / 3) rho g(1+)   (hard spheres, sigma = 1)."""
    return rho * (1.0 + 2.0 * np.pi / 3.0 * rho * g_contact)


def collect(run_dir, rho=None):
    """All observables found in a run directory, as per-sample series."""
    out = {}
    p = os.path.join(run_dir, "press_virial.dat")
    if os.path.exists(p):
        rr, g, _ = read_hist_file(p)
        out["pressv_rr"], out["pressv_g"] = rr, g
        out["g_contact"] = contact_value(rr, g)
    p = os.path.join(run_dir, "press_thermo.dat")
    if os.path.exists(p):
        xi, h, _ = read_hist_file(p)
        out["presst_xi"], out["presst_h"] = xi, h
    p = os.path.join(run_dir, "chem_pot.dat")
    if os.path.exists(p):
        c = read_column_file(p)
        out["mu"], out["widom_frac"] = c[:, 0], c[:, 1]
    p = os.path.join(run_dir, "order_param.dat")
    if os.path.exists(p):
        out["ql"] = read_column_file(p)[:, 0]
    p = os.path.join(run_dir, "density.dat")
    if os.path.exists(p):
        out["density"] = read_column_file(p)[:, 0]
    p = os.path.join(run_dir, "rdf_000000.dat.gz")
    if os.path.exists(p):
        rr, g, _ = read_hist_file(p)
        out["rdf_rr"], out["rdf_g"] = rr, g
    p = os.path.join(run_dir, "out.txt")
    if os.path.exists(p):
        txt = open(p).read()
        out["stdout"] = np.array(txt)
    return out
