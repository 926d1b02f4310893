"""Golden snapshot written by the UNMODIFIED reference (write_config, io_config.c:134-191).

Run in the build container (needs oracle/_ref): equilibrates a small fcc box with the
reference's own sweep_nvt(), then lets the reference write two samples into one
config_%06d.dat.gz (second sample appended).  Commits the gz file as the reference wrote it
and the two double-precision tables it was written from.

    python tests/golden/make_config_golden.py
"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle  # noqa: E402

pyoracle.build()
r = pyoracle.Ref(lattice=(2, 5, 5, 5, 0.8), neigh_dr=1.0, max_part=10, seed=77)
r.set_moves(dr_max=0.2)
out = os.path.join(HERE, "io")
os.makedirs(out, exist_ok=True)
cwd = os.getcwd()
tmp = tempfile.mkdtemp()
os.chdir(tmp)
try:
    r.sweep_nvt(50)
    c0 = r.get_conf().copy()
    r.write_config(50, 2)
    r.sweep_nvt(50)
    c1 = r.get_conf().copy()
    r.write_config(100, 2)
    box = r.box4()
finally:
    os.chdir(cwd)
files = sorted(os.listdir(tmp))
assert len(files) == 1, files
shutil.copy(os.path.join(tmp, files[0]), os.path.join(out, "config_ref.dat.gz"))
np.savez_compressed(os.path.join(out, "config_ref_tables.npz"), conf0=c0, conf1=c1, box=np.asarray(box[:3]),
                    sweeps=np.array([50, 100]))
print("wrote", os.path.join(out, "config_ref.dat.gz"), os.path.getsize(os.path.join(out, "config_ref.dat.gz")), "bytes")
