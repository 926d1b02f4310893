"""Snapshot writer timing: the reference's write_config() (serial gzprintf, io_config.c:165-178)
against hs_fastio_write_config() on the same table.  CPU only.  usage: io_bench.py [fcc cells] [threads]"""
import ctypes as C, gzip, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pyoracle
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 64
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
HOST = os.path.join(ROOT, "hsmc_b200", "host")
subprocess.run(["make", "-C", HOST, os.path.join(HOST, "libhs_fastio.so")], check=True, capture_output=True)
L = C.CDLL(os.path.join(HOST, "libhs_fastio.so"))
L.hs_fastio_write_config.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_int]
r = pyoracle.Ref(lattice=(2, cells, cells, cells, 0.9), neigh_dr=1.0, max_part=10, seed=1)
r.set_moves(dr_max=0.1)
r.sweep_nvt(1)
conf, box = r.get_conf().copy(), r.box4()
tmp = tempfile.mkdtemp()
os.chdir(tmp)
t0 = time.perf_counter(); r.write_config(1, 1); t_ref = time.perf_counter() - t0
ref_file = os.listdir(".")[0]
b = (C.c_double * 3)(*box[:3])
t0 = time.perf_counter()
assert L.hs_fastio_write_config(b"mine.gz", 0, 1, conf.shape[0], b, C.c_void_p(conf.ctypes.data), threads) == 0
t_new = time.perf_counter() - t0
same = gzip.open("mine.gz").read() == gzip.open(ref_file).read()
print(f"N={conf.shape[0]}: reference write_config {t_ref:.2f} s ({os.path.getsize(ref_file)/1e6:.1f} MB), "
      f"hs_fastio {t_new:.2f} s ({os.path.getsize('mine.gz')/1e6:.1f} MB) on {threads or os.cpu_count()} threads, "
      f"x{t_ref/t_new:.1f}, identical bytes after gunzip: {same}")
