"""The seam-by-seam patch of INTEGRATION.md is real: integration/patch_reference.py applies it to the
reference's sources (found by function name, nothing of the reference is stored here), the result compiles
with the author's flags and links against libhsmc_gpu.so, and the patched reference runs its own set-up
until it reaches the device -- where, on a machine without a GPU, it stops with the library's error
through the reference's own error convention (there is no CPU fallback).  Needs /root/reference."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF_SRC = "/root/reference/src"
PATCHER = os.path.join(ROOT, "integration", "patch_reference.py")

INPUT = ("rho 0.5\ncells_x 6\ncells_y 6\ncells_z 6\ntype 1\nneigh_list 1.05 10\ndr_max 0.05\nopt 0 10 2 0.5 0.5\n"
         "press_virial 0.002 1\nwidom 100 1\nseed 11\nsweep_eq 2\nsweep_stat 2\nout 1\n")


@pytest.fixture(scope="module")
def patched(lib_built, tmp_path_factory):
    if not os.path.isdir(REF_SRC):
        pytest.skip("the reference sources are not on this machine")
    dst = str(tmp_path_factory.mktemp("patched") / "src")
    out = subprocess.run([sys.executable, PATCHER, REF_SRC, dst, "--build"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    return dst, out.stdout


def test_patch_touches_only_the_seams(patched):
    dst, log = patched
    changed = set(log.split("patched:")[1].split("\n")[0].replace(",", " ").split())
    assert changed == {"cell_list.c", "cell_list.h", "nvt.c", "npt.c", "moves.c", "compute_widom_chem_pot.c", "compute_rdf.c",
                       "compute_press.c", "compute_order_parameter.c", "io_config.c"}
    for f in sorted(os.listdir(REF_SRC)):
        same = open(os.path.join(REF_SRC, f), "rb").read() == open(os.path.join(dst, f), "rb").read()
        assert same == (f not in changed), f
    # the serial loops are gone from the patched seams, the calls into the ABI are there
    moves = open(os.path.join(dst, "moves.c")).read()
    assert "hsmc_gpu_overlap_scaled" in moves and "hsmc_gpu_rescale" in moves and "hsmc_gpu_counters" in moves
    assert "hsmc_gpu_sweep_nvt" in open(os.path.join(dst, "nvt.c")).read()
    assert "hsmc_gpu_widom" in open(os.path.join(dst, "compute_widom_chem_pot.c")).read()


def test_patched_reference_runs_to_the_device_boundary(patched, tmp_path):
    dst, _ = patched
    (tmp_path / "in.dat").write_text(INPUT)
    out = subprocess.run([os.path.join(dst, "hsmc_gpu_patched"), "-i", "in.dat"], cwd=tmp_path, capture_output=True, text=True,
                         timeout=300)
    assert "Number of particles: 216" in out.stdout and "Simulation box size (x, y, z): 7.55953 7.55953 7.55953" in out.stdout
    import hsmc_b200
    if hsmc_b200.load_library().hsmc_gpu_device_count() < 1:
        assert out.returncode == 1 and "ERROR: no CUDA device" in out.stdout
    else:
        assert out.returncode == 0 and "Simulation complete!" in out.stdout
