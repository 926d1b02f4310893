"""CPU tests of the parallel snapshot writer (hsmc_b200/host/hs_fastio.c):
the decompressed bytes must equal what the reference's write_config() (io_config.c:134-191)
produces -- checked against a golden file written by the unmodified reference, against the
reference itself where it is compiled, and against printf-formatted text."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

HOST = os.path.join(ROOT, "hsmc_b200", "host")


@pytest.fixture(scope="module")
def fastio():
    subprocess.run(["make", "-C", HOST, os.path.join(HOST, "libhs_fastio.so")], check=True, capture_output=True)
    L = C.CDLL(os.path.join(HOST, "libhs_fastio.so"))
    L.hs_fmt_f8.argtypes = [C.c_char_p, C.c_double]
    L.hs_fastio_write_config.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_int]
    return L


def _write(L, path, append, sweep, box, conf, threads):
    conf = np.ascontiguousarray(conf, dtype=np.float64)
    b = (C.c_double * 3)(*[float(x) for x in box[:3]])
    rc = L.hs_fastio_write_config(os.fsencode(path), append, sweep, conf.shape[0], b, C.c_void_p(conf.ctypes.data), threads)
    assert rc == 0


def _expected(sweep, box, conf):
    head = "# Sweep number\n%d\n# Number of particles\n%d\n# Simulation box size\n%.8f\n%.8f\n%.8f\n# Configuration\n" % (
        sweep, conf.shape[0], box[0], box[1], box[2])
    body = "".join("%d %.8f %.8f %.8f\n" % (int(r[0]), r[1], r[2], r[3]) for r in conf)
    return (head + body).encode()


def test_fmt_f8_is_printf(fastio):
    rng = np.random.default_rng(1)
    xs = np.concatenate([
        rng.random(20000) * 300.0,                                   # coordinates
        rng.random(2000) * 1e-7,                                     # below the last digit
        np.arange(1, 4000) / 512.0,                                  # exact ties at the 9th decimal (k/512)
        np.arange(1, 2000, 2) / 1024.0 + 17.0,
        np.array([0.0, -0.0, 0.999999995, 0.9999999949999999, 0.99999999500000001, 9.999999996, 1e-300, 5e-324,
                  0.5e-8, 1.5e-8, 2.5e-8, 123456789.123456789, 4503599627370496.5, -3.25, -1e-12, 1e15, 2.0 ** 52,
                  1e300, float("inf")]),
        np.nextafter(np.arange(1, 200) / 512.0, 0.0), np.nextafter(np.arange(1, 200) / 512.0, 1.0),
    ])
    buf = C.create_string_buffer(512)
    for x in xs:
        n = fastio.hs_fmt_f8(buf, float(x))
        assert buf.raw[:n].decode() == "%.8f" % x, repr(float(x))


def test_golden_reference_snapshot(fastio, tmp_path):
    """the file the unmodified reference wrote (two samples, the second appended)"""
    t = np.load(os.path.join(GOLDEN, "io", "config_ref_tables.npz"))
    want = gzip.open(os.path.join(GOLDEN, "io", "config_ref.dat.gz")).read()
    p = str(tmp_path / "config_000000.dat.gz")
    _write(fastio, p, 0, int(t["sweeps"][0]), t["box"], t["conf0"], 3)
    _write(fastio, p, 1, int(t["sweeps"][1]), t["box"], t["conf1"], 1)
    assert gzip.open(p).read() == want
    # zlib's own multi-member reader (what the reference's tooling uses) agrees
    assert subprocess.run(["gzip", "-dc", p], capture_output=True, check=True).stdout == want


def test_against_live_reference(fastio, oracle_built, tmp_path):
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref not built here")
    r = oracle_built.Ref(lattice=(2, 4, 4, 4, 0.9), neigh_dr=1.0, max_part=10, seed=5)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        r.set_moves(dr_max=0.1)
        r.sweep_nvt(10)
        before = set(os.listdir("."))
        r.write_config(7, 1)
        ref_file = (set(os.listdir(".")) - before).pop()
        conf, box = r.get_conf().copy(), r.box4()
    finally:
        os.chdir(cwd)
        r.close()
    p = str(tmp_path / "mine.gz")
    _write(fastio, p, 0, 7, box, conf, 2)
    assert gzip.open(p).read() == gzip.open(str(tmp_path / ref_file)).read()


@pytest.mark.parametrize("n", [0, 1, 32768, 32769, 100_003])
def test_chunking_and_thread_count_do_not_change_the_bytes(fastio, tmp_path, n):
    rng = np.random.default_rng(n)
    conf = np.empty((n, 4))
    conf[:, 0] = rng.permutation(n)
    conf[:, 1:] = rng.random((n, 3)) * np.array([210.4, 105.2, 105.2])
    if n > 10:
        conf[3, 1] = 210.4 - 1e-12          # rounds up across the decimal point
        conf[5, 2] = 7.0 / 512.0            # exact tie
    box = (210.4, 105.2, 105.2)
    want = _expected(12, box, conf)
    for threads in (1, 8):
        p = str(tmp_path / f"c{threads}.gz")
        _write(fastio, p, 0, 12, box, conf, threads)
        assert gzip.open(p).read() == want


FETCH = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_longlong, C.c_longlong)


@pytest.mark.parametrize("n,slice_rows", [(0, 1000), (1, 1000), (100_003, 7001), (100_003, 1 << 20), (250_000, 32768)])
def test_streamed_table_gives_the_same_bytes(fastio, tmp_path, n, slice_rows):
    """hs_fastio_write_config_stream: the table arrives slice by slice (on the GPU: D2H copies of the id-ordered
    table, hs_sim.c:hs_write_config) while earlier slices are formatted and deflated; a chunk never reads rows
    that have not arrived -- the destination starts as NaN, which would show in the text."""
    import time
    fastio.hs_fastio_write_config_stream.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                                     C.c_void_p, C.c_int, FETCH, C.c_void_p, C.c_int]
    rng = np.random.default_rng(n + slice_rows)
    src = np.empty((n, 4))
    src[:, 0] = rng.permutation(n)
    src[:, 1:] = rng.random((n, 3)) * 97.3
    dst = np.full((n, 4), np.nan)
    calls = []

    def fetch(ctx, first, m):
        time.sleep(0.002)                       # the writer threads are ahead of the producer
        dst[first:first + m] = src[first:first + m]
        calls.append((first, m))
        return 0

    box = (97.3, 97.3, 97.3)
    b = (C.c_double * 3)(*box)
    p = str(tmp_path / "s.gz")
    rc = fastio.hs_fastio_write_config_stream(os.fsencode(p), 0, 3, n, b, C.c_void_p(dst.ctypes.data), 4, FETCH(fetch), None,
                                              slice_rows)
    assert rc == 0
    assert gzip.open(p).read() == _expected(3, box, src)
    assert calls == [(f, min(slice_rows, n - f)) for f in range(0, n, slice_rows)]


def test_streamed_table_fetch_failure_is_reported(fastio, tmp_path):
    fastio.hs_fastio_write_config_stream.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                                     C.c_void_p, C.c_int, FETCH, C.c_void_p, C.c_int]
    n = 200_000
    dst = np.zeros((n, 4))
    b = (C.c_double * 3)(1.0, 1.0, 1.0)
    rc = fastio.hs_fastio_write_config_stream(os.fsencode(str(tmp_path / "f.gz")), 0, 0, n, b, C.c_void_p(dst.ctypes.data), 4,
                                              FETCH(lambda ctx, first, m: -1 if first else 0), None, 50_000)
    assert rc != 0


def test_unwritable_path_reports_failure(fastio, tmp_path):
    conf = np.zeros((4, 4))
    b = (C.c_double * 3)(1.0, 1.0, 1.0)
    rc = fastio.hs_fastio_write_config(os.fsencode(str(tmp_path / "no" / "such" / "dir.gz")), 0, 0, 4, b,
                                       C.c_void_p(conf.ctypes.data), 2)
    assert rc != 0


def test_fmt_f8_property(fastio):
    """hypothesis: any finite double formats exactly as printf("%.8f") does (CPython's repr-exact formatting
    is correctly rounded, ties to even, like glibc's)."""
    from hypothesis import given, settings, strategies as st
    buf = C.create_string_buffer(512)

    @settings(max_examples=3000, deadline=None)
    @given(st.one_of(st.floats(allow_nan=False, allow_infinity=False, width=64),
                     st.floats(min_value=0.0, max_value=500.0),
                     st.integers(min_value=0, max_value=10 ** 11).map(lambda k: k / 2.0 ** 9),       # k/512: ties
                     st.integers(min_value=0, max_value=10 ** 12).map(lambda k: (2 * k + 1) * 5e-9)))  # near ...5e-9
    def check(x):
        n = fastio.hs_fmt_f8(buf, float(x))
        assert buf.raw[:n].decode() == "%.8f" % x

    check()
