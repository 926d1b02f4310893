"""ctypes bindings for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Two checkers with the same surface:

* ``Port``  -- oracle/liboracle_port.so, the plain-C restatement (hsmc_oracle.c).
* ``Ref``   -- oracle/_ref/libhsmc_ref.so, the UNMODIFIED reference sources compiled
  against the GSL shim, driven through ref_harness.c.  The reference keeps its state
  in globals, so only one ``Ref`` may be live per process.

Neither is ever used by the product path (hsmc_b200); see DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle_port.so")
REF_SO = os.path.join(HERE, "_ref", "libhsmc_ref.so")
REF_EXE = os.path.join(HERE, "_ref", "hsmc_ref")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)


def build(verbose: bool = False) -> None:
    """Compile the port (always) and oracle/_ref (only where /root/reference exists)."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _p(a, t):
    return a.ctypes.data_as(t)


def _c4(conf):
    conf = np.ascontiguousarray(conf, dtype=np.float64)
    assert conf.ndim == 2 and conf.shape[1] == 4
    return conf


def conf_from_xyz(xyz):
    xyz = np.asarray(xyz, dtype=np.float64)
    conf = np.empty((xyz.shape[0], 4))
    conf[:, 0] = np.arange(xyz.shape[0])
    conf[:, 1:] = xyz
    return conf


class Port:
    """Restated oracle (explicit state; many instances allowed)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(PORT_SO):
                build()
            L = C.CDLL(PORT_SO)
            L.orc_create.restype = C.c_void_p
            L.orc_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, _dp, C.c_double, C.c_int]
            L.orc_destroy.argtypes = [C.c_void_p]
            L.orc_N.argtypes = [C.c_void_p]
            L.orc_status.argtypes = [C.c_void_p]
            L.orc_cells.argtypes = [C.c_void_p, _ip, _dp]
            L.orc_get_conf.argtypes = [C.c_void_p, _dp]
            L.orc_set_conf.argtypes = [C.c_void_p, _dp]
            L.orc_cell_of.argtypes = [C.c_void_p, C.c_int]
            L.orc_compute_dist.restype = C.c_double
            L.orc_compute_dist.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
            L.orc_check_overlap.argtypes = [C.c_void_p, C.c_int, C.c_double]
            L.orc_trial_verdicts.argtypes = [C.c_void_p, C.c_int, _ip, _dp, C.c_double, _ip]
            L.orc_overlap_all.argtypes = [C.c_void_p, C.c_double, _ip]
            L.orc_any_overlap.argtypes = [C.c_void_p, C.c_double]
            L.orc_part_move_raw.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
            L.orc_replay_moves.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
            L.orc_replay_moves.restype = None
            L.orc_counters.argtypes = [C.c_void_p, _i64p]
            L.orc_reset_counters.argtypes = [C.c_void_p]
            L.orc_rescale.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
            L.orc_mt_create.restype = C.c_void_p
            L.orc_mt_create.argtypes = [C.c_ulong]
            L.orc_mt_destroy.argtypes = [C.c_void_p]
            L.orc_mt_raw.restype = C.c_uint32
            L.orc_mt_raw.argtypes = [C.c_void_p]
            L.orc_mt_double.restype = C.c_double
            L.orc_mt_double.argtypes = [C.c_void_p]
            L.orc_mt_int.argtypes = [C.c_void_p, C.c_int]
            L.orc_sweep_nvt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double]
            L.orc_widom_verdicts.argtypes = [C.c_void_p, C.c_int, _dp, _ip]
            L.orc_widom_count_raw.restype = C.c_int64
            L.orc_widom_count_raw.argtypes = [C.c_void_p, C.c_int64, _u32p]
            L.orc_rdf_nn.argtypes = [C.c_double, C.c_double]
            L.orc_rdf_counts.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, _u64p]
            L.orc_pressv_nn.argtypes = [C.c_double]
            L.orc_pressv_counts.argtypes = [C.c_void_p, C.c_double, C.c_int, _u64p]
            L.orc_presst_nn.argtypes = [C.c_double, C.c_double]
            L.orc_presst_flags.argtypes = [C.c_void_p, C.c_double, C.c_int, _ip, _dp]
            L.orc_philox4x32.argtypes = [_u32p, _u32p, _u32p]
            L.orc_order_param.argtypes = [C.c_void_p, C.c_int, C.c_double]
            L.orc_order_param.restype = C.c_double
            L.orc_box_from_lattice.argtypes = [C.c_int] * 4 + [C.c_double, _dp]
            L.orc_lattice_count.argtypes = [C.c_int] * 4
            L.orc_lattice_fill.argtypes = [C.c_int] * 4 + [C.c_double, _dp]
            cls._lib = L
        return cls._lib

    # ---- lattice helpers -------------------------------------------------
    @classmethod
    def lattice(cls, type_, nx, ny, nz, rho):
        L = cls.lib()
        box = np.zeros(4)
        L.orc_box_from_lattice(type_, nx, ny, nz, rho, _p(box, _dp))
        n = L.orc_lattice_count(type_, nx, ny, nz)
        conf = np.zeros((n, 4))
        L.orc_lattice_fill(type_, nx, ny, nz, rho, _p(conf, _dp))
        return box, conf

    @classmethod
    def philox(cls, ctr, key):
        c = np.asarray(ctr, dtype=np.uint32)
        k = np.asarray(key, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        cls.lib().orc_philox4x32(_p(c, _u32p), _p(k, _u32p), _p(o, _u32p))
        return o

    def __init__(self, conf, box, neigh_dr=1.0, max_part=10):
        self.L = self.lib()
        conf = _c4(conf)
        self.N = conf.shape[0]
        self.box = tuple(float(b) for b in box[:3])
        self.h = self.L.orc_create(self.N, *self.box, _p(conf, _dp), neigh_dr, max_part)
        if self.L.orc_status(self.h):
            st = self.L.orc_status(self.h)
            self.close()
            raise RuntimeError(f"oracle port: reference would exit (status {st})")

    def close(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def status(self):
        return self.L.orc_status(self.h)

    def cells(self):
        n = np.zeros(3, dtype=np.int32)
        s = np.zeros(3)
        self.L.orc_cells(self.h, _p(n, _ip), _p(s, _dp))
        return n, s

    def get_conf(self):
        c = np.zeros((self.N, 4))
        self.L.orc_get_conf(self.h, _p(c, _dp))
        return c

    def set_conf(self, conf):
        self.L.orc_set_conf(self.h, _p(_c4(conf), _dp))

    def compute_dist(self, i, j, sf=1.0):
        return self.L.orc_compute_dist(self.h, i, j, sf)

    def trial_verdicts(self, idx, xyz, sf=1.0):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        f = np.zeros(idx.shape[0], dtype=np.int32)
        self.L.orc_trial_verdicts(self.h, idx.shape[0], _p(idx, _ip), _p(xyz, _dp), sf, _p(f, _ip))
        return f

    def overlap_all(self, sf=1.0):
        f = np.zeros(self.N, dtype=np.int32)
        self.L.orc_overlap_all(self.h, sf, _p(f, _ip))
        return f

    def any_overlap(self, sf=1.0):
        return self.L.orc_any_overlap(self.h, sf)

    def part_move_raw(self, idx, rx, ry, rz, dr_max):
        return self.L.orc_part_move_raw(self.h, int(idx), int(rx), int(ry), int(rz), dr_max)

    def replay_moves(self, ids, raw, dr_max):
        """Drive part_move with explicit (id, rx, ry, rz) draws; returns accept flags."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        raw = np.ascontiguousarray(raw, dtype=np.uint32).reshape(-1, 3)
        out = np.zeros(len(ids), dtype=np.int32)
        self.L.orc_replay_moves(self.h, len(ids), ids.ctypes.data, raw.ctypes.data, float(dr_max), out.ctypes.data)
        return out

    def counters(self):
        o = np.zeros(6, dtype=np.int64)
        self.L.orc_counters(self.h, _p(o, _i64p))
        return o

    def reset_counters(self):
        self.L.orc_reset_counters(self.h)

    def rescale(self, sf, box):
        self.box = tuple(float(b) for b in box[:3])
        self.L.orc_rescale(self.h, sf, *self.box)

    def sweep_nvt(self, n_sweeps, dr_max, seed):
        m = self.L.orc_mt_create(seed)
        self.L.orc_sweep_nvt(self.h, m, n_sweeps, dr_max)
        self.L.orc_mt_destroy(m)

    def widom_verdicts(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        f = np.zeros(xyz.shape[0], dtype=np.int32)
        self.L.orc_widom_verdicts(self.h, xyz.shape[0], _p(xyz, _dp), _p(f, _ip))
        return f

    def widom_count_raw(self, raw3):
        raw3 = np.ascontiguousarray(raw3, dtype=np.uint32)
        return int(self.L.orc_widom_count_raw(self.h, raw3.shape[0], _p(raw3, _u32p)))

    def rdf_counts(self, dr, rmax):
        nn = self.L.orc_rdf_nn(dr, rmax)
        c = np.zeros(nn, dtype=np.uint64)
        self.L.orc_rdf_counts(self.h, dr, rmax, nn, _p(c, _u64p))
        return c

    def pressv_counts(self, dr):
        nn = self.L.orc_pressv_nn(dr)
        c = np.zeros(nn, dtype=np.uint64)
        self.L.orc_pressv_counts(self.h, dr, nn, _p(c, _u64p))
        return c

    def order_param(self, l, rmax):
        """Average Steinhardt q_l with bond cutoff rmax (<= cell edge of the neighbour list)."""
        return float(self.L.orc_order_param(self.h, int(l), float(rmax)))

    def presst_flags(self, dxi, xi_max):
        nn = self.L.orc_presst_nn(dxi, xi_max)
        f = np.zeros(nn, dtype=np.int32)
        sf = np.zeros(nn)
        self.L.orc_presst_flags(self.h, dxi, nn, _p(f, _ip), _p(sf, _dp))
        return f, sf


class MT:
    """MT19937 stream with the reference's rng.c call semantics (port side)."""

    def __init__(self, seed):
        self.L = Port.lib()
        self.h = self.L.orc_mt_create(seed)

    def raw(self):
        return self.L.orc_mt_raw(self.h)

    def double(self):
        return self.L.orc_mt_double(self.h)

    def int(self, n):
        return self.L.orc_mt_int(self.h, n)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_mt_destroy(self.h)
            self.h = None


class Ref:
    """The unmodified reference, through oracle/ref_harness.c.  One live instance per process."""

    _lib = None
    _live = False

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO)
            L = C.CDLL(REF_SO)
            L.ref_setup_lattice.argtypes = [C.c_int] * 4 + [C.c_double, C.c_double, C.c_int, C.c_ulong]
            L.ref_setup_box.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, _dp, C.c_double, C.c_int, C.c_ulong]
            L.ref_box.argtypes = [_dp]
            L.ref_cells.argtypes = [_ip, _dp]
            L.ref_get_conf.argtypes = [_dp]
            L.ref_set_conf.argtypes = [_dp]
            L.ref_set_moves.argtypes = [C.c_double] * 3
            L.ref_get_rho.restype = C.c_double
            L.ref_script.argtypes = [_u32p, C.c_size_t]
            L.ref_script_pos.restype = C.c_size_t
            L.ref_trial_verdicts.argtypes = [C.c_int, _ip, _dp, C.c_double, _ip]
            L.ref_overlap_all.argtypes = [C.c_double, _ip]
            L.ref_any_overlap.argtypes = [C.c_double]
            L.ref_compute_dist.restype = C.c_double
            L.ref_compute_dist.argtypes = [C.c_int, C.c_int, C.c_double]
            L.ref_part_moves.argtypes = [C.c_long]
            L.ref_counters.argtypes = [_ip]
            L.ref_widom.argtypes = [C.c_int, _dp]
            L.ref_widom_verdicts.argtypes = [C.c_int, _ip]
            L.ref_rdf_hist.argtypes = [C.c_double, C.c_double, _dp, C.c_int]
            L.ref_pressv_hist.argtypes = [C.c_double, _dp, C.c_int]
            L.ref_order_param.argtypes = [C.c_int, C.c_double]
            L.ref_order_param.restype = C.c_double
            L.ref_write_config.argtypes = [C.c_int, C.c_int]
            L.ref_presst_hist.argtypes = [C.c_double, C.c_double, _dp, _dp, C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, conf=None, box=None, neigh_dr=1.0, max_part=10, seed=4357, lattice=None):
        if Ref._live:
            raise RuntimeError("only one Ref may be live (the reference uses globals)")
        self.L = self.lib()
        if lattice is not None:
            type_, nx, ny, nz, rho = lattice
            rc = self.L.ref_setup_lattice(type_, nx, ny, nz, rho, neigh_dr, max_part, seed)
        else:
            conf = _c4(conf)
            rc = self.L.ref_setup_box(conf.shape[0], float(box[0]), float(box[1]), float(box[2]),
                                      _p(conf, _dp), neigh_dr, max_part, seed)
        assert rc == 0
        Ref._live = True
        self.N = self.L.ref_N()
        self._script = None
        if conf is not None and lattice is not None:
            self.set_conf(conf)

    def close(self):
        if Ref._live and getattr(self, "L", None) is not None:
            self.L.ref_teardown()
            Ref._live = False
            self.L = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def box4(self):
        b = np.zeros(4)
        self.L.ref_box(_p(b, _dp))
        return b

    def cells(self):
        n = np.zeros(3, dtype=np.int32)
        s = np.zeros(3)
        self.L.ref_cells(_p(n, _ip), _p(s, _dp))
        return n, s

    def get_conf(self):
        c = np.zeros((self.N, 4))
        self.L.ref_get_conf(_p(c, _dp))
        return c

    def set_conf(self, conf):
        self.L.ref_set_conf(_p(_c4(conf), _dp))

    def set_moves(self, dr_max=0.05, dv_max=0.001, press=0.0):
        self.L.ref_set_moves(dr_max, dv_max, press)

    def rho(self):
        return self.L.ref_get_rho()

    def script(self, raw):
        """Replay these raw 32-bit outputs through gsl_rng_get from now on (None = real MT)."""
        if raw is None:
            self._script = None
            self.L.ref_script(None, 0)
        else:
            self._script = np.ascontiguousarray(raw, dtype=np.uint32).ravel()
            self.L.ref_script(_p(self._script, _u32p), self._script.shape[0])

    def script_pos(self):
        return self.L.ref_script_pos()

    def compute_dist(self, i, j, sf=1.0):
        return self.L.ref_compute_dist(i, j, sf)

    def trial_verdicts(self, idx, xyz, sf=1.0):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        f = np.zeros(idx.shape[0], dtype=np.int32)
        self.L.ref_trial_verdicts(idx.shape[0], _p(idx, _ip), _p(xyz, _dp), sf, _p(f, _ip))
        return f

    def overlap_all(self, sf=1.0):
        f = np.zeros(self.N, dtype=np.int32)
        self.L.ref_overlap_all(sf, _p(f, _ip))
        return f

    def any_overlap(self, sf=1.0):
        return self.L.ref_any_overlap(sf)

    def part_moves(self, n):
        self.L.ref_part_moves(n)

    def replay_moves(self, ids, raw, dr_max):
        """Drive the reference's own part_move() with scripted draws.

        gsl_rng_uniform_int(N) computes k = raw/scale with scale = 0xffffffff/N, so the
        raw value id*scale selects particle id (rng.c:34-36)."""
        ids = np.asarray(ids, dtype=np.uint64)
        raw = np.asarray(raw, dtype=np.uint32)
        scale = np.uint64(0xFFFFFFFF // self.N)
        s = np.empty((ids.shape[0], 4), dtype=np.uint32)
        s[:, 0] = (ids * scale).astype(np.uint32)
        s[:, 1:] = raw
        self.set_moves(dr_max=dr_max)
        before = self.counters()
        self.script(s)
        self.L.ref_part_moves(ids.shape[0])
        assert self.script_pos() == s.size
        self.script(None)
        return self.counters() - before

    def vol_move(self):
        self.L.ref_vol_move()

    def sweep_nvt(self, n):
        self.L.ref_sweep_nvt(n)

    def sweep_npt(self, n):
        self.L.ref_sweep_npt(n)

    def reset_counters(self):
        self.L.ref_reset_counters()

    def counters(self):
        o = np.zeros(6, dtype=np.int32)
        self.L.ref_counters(_p(o, _ip))
        return o.astype(np.int64)

    def widom(self, M):
        mu = C.c_double(0)
        w = self.L.ref_widom(M, C.byref(mu))
        return w, mu.value

    def widom_count_raw(self, raw3):
        raw3 = np.ascontiguousarray(raw3, dtype=np.uint32)
        self.script(raw3)
        w, _ = self.widom(raw3.shape[0])
        self.script(None)
        return w

    def widom_verdicts_raw(self, raw3):
        raw3 = np.ascontiguousarray(raw3, dtype=np.uint32)
        f = np.zeros(raw3.shape[0], dtype=np.int32)
        self.script(raw3)
        self.L.ref_widom_verdicts(raw3.shape[0], _p(f, _ip))
        self.script(None)
        return f

    def order_param(self, l, rmax):
        """global_ql_compute() of the unmodified reference (compute_order_parameter.c:84-97)."""
        return float(self.L.ref_order_param(int(l), float(rmax)))

    def write_config(self, sweep, samples_per_file=1):
        """write_config() of the unmodified reference (io_config.c:134-191) into the cwd."""
        self.L.ref_write_config(int(sweep), int(samples_per_file))

    def rdf_hist(self, dr, rmax):
        buf = np.zeros(1 << 16)
        nn = self.L.ref_rdf_hist(dr, rmax, _p(buf, _dp), buf.shape[0])
        return buf[:nn].copy()

    def pressv_hist(self, dr):
        buf = np.zeros(4096)
        nn = self.L.ref_pressv_hist(dr, _p(buf, _dp), buf.shape[0])
        return buf[:nn].copy()

    def presst_hist(self, dxi, xi_max):
        buf = np.zeros(4096)
        xi = np.zeros(4096)
        nn = self.L.ref_presst_hist(dxi, xi_max, _p(buf, _dp), _p(xi, _dp), buf.shape[0])
        return buf[:nn].copy(), xi[:nn].copy()
