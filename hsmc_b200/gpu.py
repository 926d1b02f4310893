"""ctypes binding of include/hsmc_gpu.h.  Mirrors the reference's seams one to one
(sweep_nvt, check_overlap under scaling, widom_insertion, rdf_hist_compute,
pressv_compute_hist, presst_compute_hist, get/reset_moves_counters); see the header for
the reference file:line of each."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("HSMC_GPU_LIB") or os.path.join(_HERE, "csrc", "libhsmc_gpu.so")   # env: kernel-variant experiments
_lib = None

NCCL_ID_BYTES = 128

# every symbol include/hsmc_gpu.h declares
ABI_SYMBOLS = [
    "hsmc_gpu_last_error", "hsmc_gpu_device_count", "hsmc_gpu_nccl_id", "hsmc_gpu_create",
    "hsmc_gpu_destroy", "hsmc_gpu_ipc_export", "hsmc_gpu_ipc_attach", "hsmc_gpu_get_info", "hsmc_gpu_plan", "hsmc_gpu_plan_blocks", "hsmc_gpu_stream", "hsmc_gpu_sync", "hsmc_gpu_upload",
    "hsmc_gpu_download", "hsmc_gpu_download_owned", "hsmc_gpu_pack_table", "hsmc_gpu_fetch_rows", "hsmc_gpu_pin_host", "hsmc_gpu_sweep_nvt", "hsmc_gpu_overlap_scaled",
    "hsmc_gpu_rescale", "hsmc_gpu_widom", "hsmc_gpu_rdf_counts", "hsmc_gpu_rdf_counts_part", "hsmc_gpu_contact_counts",
    "hsmc_gpu_presst_flags", "hsmc_gpu_order_parameter", "hsmc_gpu_counters", "hsmc_gpu_reset_counters", "hsmc_gpu_add_vol_move",
    "hsmc_gpu_cell_rejects", "hsmc_gpu_profile", "hsmc_gpu_profile_read", "hsmc_gpu_set_sweep_counter", "hsmc_gpu_trial_verdicts",
    "hsmc_gpu_widom_verdicts", "hsmc_gpu_sweep_nvt_logged", "hsmc_gpu_selftest_u01", "hsmc_gpu_min_dist2",
]


class HsmcError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("device", C.c_int), ("rank", C.c_int), ("world", C.c_int), ("nccl_id", C.c_void_p),
                ("seed", C.c_uint64), ("cell_min", C.c_double), ("regrid_interval", C.c_int),
                ("sweep_impl", C.c_int)]


class _Info(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("rank", C.c_int), ("world", C.c_int), ("n_total", C.c_int64),
                ("n_owned", C.c_int64), ("n_local", C.c_int64), ("cells", C.c_int * 3), ("own_x0", C.c_int),
                ("own_x1", C.c_int), ("cell_size", C.c_double * 3), ("box", C.c_double * 3),
                ("sweeps_done", C.c_uint64), ("kernel_launches", C.c_uint64), ("nccl_calls", C.c_uint64)]


PLAN_MAX_XCUTS = 512


class _BlockPlan(C.Structure):
    _fields_ = [("ok", C.c_int), ("blocks", C.c_int * 3), ("max_extent", C.c_int * 3), ("ctas_per_phase", C.c_int),
                ("staged_capacity", C.c_int), ("smem_bytes", C.c_int), ("n_xcuts", C.c_int),
                ("xcuts", C.c_int * PLAN_MAX_XCUTS)]


class Trial(C.Structure):
    _fields_ = [("seq", C.c_uint64), ("id", C.c_int32), ("verdict", C.c_int32), ("raw", C.c_uint32 * 3),
                ("pad", C.c_uint32)]


TRIAL_DTYPE = np.dtype([("seq", "<u8"), ("id", "<i4"), ("verdict", "<i4"), ("raw", "<u4", (3,)), ("pad", "<u4")])


def library_path() -> str:
    return _LIB_PATH


def load_library():
    """Load libhsmc_gpu.so; fail loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise HsmcError(f"{_LIB_PATH} is missing: run `python -m hsmc_b200.build` "
                        "(the B200 path has no CPU fallback)")
    L = C.CDLL(_LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.hsmc_gpu_last_error.restype = C.c_char_p
    L.hsmc_gpu_nccl_id.argtypes = [vp]
    L.hsmc_gpu_create.argtypes = [C.POINTER(vp), C.POINTER(_Config), C.c_int64, dp]
    L.hsmc_gpu_destroy.argtypes = [vp]
    L.hsmc_gpu_ipc_export.argtypes = [vp, vp]
    L.hsmc_gpu_ipc_attach.argtypes = [vp, vp, vp]
    L.hsmc_gpu_get_info.argtypes = [vp, C.POINTER(_Info)]
    L.hsmc_gpu_plan.argtypes = [dp, C.c_double, C.c_int, C.c_int, C.POINTER(_Info)]
    L.hsmc_gpu_plan_blocks.argtypes = [dp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(_BlockPlan)]
    L.hsmc_gpu_stream.restype = vp
    L.hsmc_gpu_stream.argtypes = [vp]
    L.hsmc_gpu_sync.argtypes = [vp]
    L.hsmc_gpu_upload.argtypes = [vp, vp, C.c_int64]
    L.hsmc_gpu_download.argtypes = [vp, vp]
    L.hsmc_gpu_download_owned.argtypes = [vp, vp, C.c_int64, C.POINTER(C.c_int64)]
    L.hsmc_gpu_pack_table.argtypes = [vp]
    L.hsmc_gpu_fetch_rows.argtypes = [vp, C.c_int64, C.c_int64, vp]
    L.hsmc_gpu_pin_host.argtypes = [vp, C.c_size_t, C.c_int]
    L.hsmc_gpu_sweep_nvt.argtypes = [vp, C.c_int, C.c_double]
    L.hsmc_gpu_overlap_scaled.argtypes = [vp, C.c_double, ip]
    L.hsmc_gpu_rescale.argtypes = [vp, C.c_double, dp]
    L.hsmc_gpu_widom.argtypes = [vp, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
    L.hsmc_gpu_rdf_counts.argtypes = [vp, C.c_double, C.c_int, vp]
    L.hsmc_gpu_order_parameter.argtypes = [vp, C.c_int, C.c_double, dp]
    L.hsmc_gpu_rdf_counts_part.argtypes = [vp, C.c_double, C.c_int, C.c_int, C.c_int, vp]
    L.hsmc_gpu_contact_counts.argtypes = [vp, C.c_double, C.c_int, vp]
    L.hsmc_gpu_presst_flags.argtypes = [vp, vp, C.c_int, vp]
    L.hsmc_gpu_counters.argtypes = [vp, vp]
    L.hsmc_gpu_reset_counters.argtypes = [vp]
    L.hsmc_gpu_add_vol_move.argtypes = [vp, C.c_int]
    L.hsmc_gpu_cell_rejects.argtypes = [vp, C.POINTER(C.c_int64)]
    L.hsmc_gpu_set_sweep_counter.argtypes = [vp, C.c_uint64]
    L.hsmc_gpu_profile.argtypes = [vp, C.c_int]
    L.hsmc_gpu_profile_read.argtypes = [vp, vp, vp]
    L.hsmc_gpu_trial_verdicts.argtypes = [vp, C.c_int, vp, vp, C.c_double, vp]
    L.hsmc_gpu_widom_verdicts.argtypes = [vp, C.c_int, vp, vp]
    L.hsmc_gpu_sweep_nvt_logged.argtypes = [vp, C.c_double, vp, C.c_int64, C.POINTER(C.c_int64)]
    L.hsmc_gpu_min_dist2.argtypes = [vp, dp]
    L.hsmc_gpu_selftest_u01.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    _lib = L
    return L


def nccl_unique_id() -> bytes:
    L = load_library()
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    if L.hsmc_gpu_nccl_id(buf):
        raise HsmcError(L.hsmc_gpu_last_error().decode())
    return buf.raw


def plan(box, cell_min=1.0, world=1, rank=0):
    """Cell grid and x-slab of `rank` for this box (host only, no GPU needed)."""
    L = load_library()
    b = (C.c_double * 3)(*[float(x) for x in box[:3]])
    i = _Info()
    if L.hsmc_gpu_plan(b, float(cell_min), int(world), int(rank), C.byref(i)):
        raise HsmcError(L.hsmc_gpu_last_error().decode())
    return {"cells": tuple(i.cells), "cell_size": tuple(i.cell_size), "own_x": (i.own_x0, i.own_x1),
            "rank": i.rank, "world": i.world}


def plan_blocks(box, n_particles, cell_min=1.0, world=1, rank=0, xpart_world=0):
    """Block partition of the two-level checkerboard for this box (host only, no GPU needed)."""
    L = load_library()
    b = (C.c_double * 3)(*[float(x) for x in box[:3]])
    o = _BlockPlan()
    if L.hsmc_gpu_plan_blocks(b, float(cell_min), int(world), int(rank), int(xpart_world), int(n_particles), C.byref(o)):
        raise HsmcError(L.hsmc_gpu_last_error().decode())
    return {"ok": bool(o.ok), "blocks": tuple(o.blocks), "max_extent": tuple(o.max_extent),
            "ctas_per_phase": o.ctas_per_phase, "staged_capacity": o.staged_capacity, "smem_bytes": o.smem_bytes,
            "xcuts": list(o.xcuts[: o.n_xcuts])}


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class HsmcGpu:
    """One GPU's (slab of the) hard-sphere system.  Thin, stateful, not thread-safe."""

    def __init__(self, n_particles, box, seed=0, device=0, rank=0, world=1, nccl_id=None, cell_min=1.0,
                 regrid_interval=1, sweep_impl=0, xpart_world=0):
        # sweep_impl: 0 default (k_sweep_gather for N >= 2e6, k_sweep_lean below); 7 / 8 force them.
        # 1 / 5: the same two chains evaluated all in double from global memory; 3: the lean kernel
        # with the fp32 error band forced to zero (negative control).  xpart_world: a single-GPU run uses the x block
        # partition of a run on that many slabs (bitwise identity checks).
        sweep_impl = int(sweep_impl) | (int(xpart_world) << 8)
        self.L = load_library()
        self.h = C.c_void_p()
        self._idbuf = C.create_string_buffer(nccl_id, NCCL_ID_BYTES) if nccl_id is not None else None
        cfg = _Config(device, rank, world, C.cast(self._idbuf, C.c_void_p) if self._idbuf is not None else None,
                      seed, cell_min, regrid_interval, sweep_impl)
        b = (C.c_double * 3)(*[float(x) for x in box[:3]])
        self.N = int(n_particles)
        self.world = world
        self._ck(self.L.hsmc_gpu_create(C.byref(self.h), C.byref(cfg), self.N, b))

    def _ck(self, rc):
        if rc:
            raise HsmcError(self.L.hsmc_gpu_last_error().decode())

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.hsmc_gpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- plumbing ----
    def ipc_export(self):
        """Opaque 64-byte description of this rank's NVLink receive window."""
        buf = C.create_string_buffer(64)
        self._ck(self.L.hsmc_gpu_ipc_export(self.h, buf))
        return buf.raw

    def ipc_attach(self, left_blob, right_blob):
        """Map the neighbours' windows: halo traffic then goes peer-to-peer over NVLink."""
        lb, rb = C.create_string_buffer(left_blob, 64), C.create_string_buffer(right_blob, 64)
        self._ck(self.L.hsmc_gpu_ipc_attach(self.h, lb, rb))

    def info(self):
        i = _Info()
        self._ck(self.L.hsmc_gpu_get_info(self.h, C.byref(i)))
        return {
            "abi_version": i.abi_version, "rank": i.rank, "world": i.world, "n_total": i.n_total,
            "n_owned": i.n_owned, "n_local": i.n_local, "cells": tuple(i.cells),
            "own_x": (i.own_x0, i.own_x1), "cell_size": tuple(i.cell_size), "box": tuple(i.box),
            "sweeps_done": i.sweeps_done, "kernel_launches": i.kernel_launches, "nccl_calls": i.nccl_calls,
        }

    def stream_ptr(self):
        return self.L.hsmc_gpu_stream(self.h)

    def sync(self):
        self._ck(self.L.hsmc_gpu_sync(self.h))

    def upload(self, conf):
        conf = np.ascontiguousarray(conf, dtype=np.float64)
        assert conf.ndim == 2 and conf.shape[1] == 4
        self._ck(self.L.hsmc_gpu_upload(self.h, _ptr(conf), conf.shape[0]))

    def upload_ptr(self, ptr, n_rows):
        self._ck(self.L.hsmc_gpu_upload(self.h, C.c_void_p(ptr), n_rows))

    def download(self, out=None):
        if out is None:
            out = np.empty((self.N, 4))
        self._ck(self.L.hsmc_gpu_download(self.h, _ptr(out)))
        return out

    def download_ptr(self, ptr):
        self._ck(self.L.hsmc_gpu_download(self.h, C.c_void_p(ptr)))

    def download_owned(self, out=None):
        if out is None:
            out = np.empty((self.info()["n_owned"], 4))
        n = C.c_int64(0)
        self._ck(self.L.hsmc_gpu_download_owned(self.h, _ptr(out), out.shape[0], C.byref(n)))
        return out[: n.value]

    def pack_table(self):
        self._ck(self.L.hsmc_gpu_pack_table(self.h))

    def fetch_rows(self, first, n, out=None):
        if out is None:
            out = np.empty((n, 4))
        self._ck(self.L.hsmc_gpu_fetch_rows(self.h, C.c_int64(first), C.c_int64(n), _ptr(out)))
        return out

    def download_owned_ptr(self, ptr, cap):
        n = C.c_int64(0)
        self._ck(self.L.hsmc_gpu_download_owned(self.h, C.c_void_p(ptr), cap, C.byref(n)))
        return n.value

    # ---- the hot path ----
    def sweep_nvt(self, n_sweeps, dr_max):
        self._ck(self.L.hsmc_gpu_sweep_nvt(self.h, int(n_sweeps), float(dr_max)))

    def overlap_scaled(self, sf):
        f = C.c_int(0)
        self._ck(self.L.hsmc_gpu_overlap_scaled(self.h, float(sf), C.byref(f)))
        return f.value

    def rescale(self, sf, new_box):
        b = (C.c_double * 3)(*[float(x) for x in new_box[:3]])
        self._ck(self.L.hsmc_gpu_rescale(self.h, float(sf), b))

    def widom(self, sample_id, count, first=0, reduce=True):
        acc = C.c_int64(0)
        self._ck(self.L.hsmc_gpu_widom(self.h, int(sample_id), int(first), int(count), int(bool(reduce)), C.byref(acc)))
        return acc.value

    def rdf_counts(self, dr, nn):
        c = np.zeros(nn, dtype=np.uint64)
        self._ck(self.L.hsmc_gpu_rdf_counts(self.h, float(dr), int(nn), _ptr(c)))
        return c

    def rdf_counts_part(self, dr, nn, part, nparts):
        """Share `part` of `nparts` of the pair triangle (replicated configuration; sum the parts)."""
        c = np.zeros(nn, dtype=np.uint64)
        self._ck(self.L.hsmc_gpu_rdf_counts_part(self.h, float(dr), int(nn), int(part), int(nparts), _ptr(c)))
        return c

    def order_parameter(self, l, rmax):
        """Average Steinhardt q_l (compute_order_parameter.c:84-229); rmax <= cell edge."""
        o = C.c_double(0)
        self._ck(self.L.hsmc_gpu_order_parameter(self.h, int(l), float(rmax), C.byref(o)))
        return o.value

    def contact_counts(self, dr, nn):
        c = np.zeros(nn, dtype=np.uint64)
        self._ck(self.L.hsmc_gpu_contact_counts(self.h, float(dr), int(nn), _ptr(c)))
        return c

    def presst_flags(self, sf):
        sf = np.ascontiguousarray(sf, dtype=np.float64)
        f = np.zeros(sf.shape[0], dtype=np.int32)
        self._ck(self.L.hsmc_gpu_presst_flags(self.h, _ptr(sf), sf.shape[0], _ptr(f)))
        return f

    def counters(self):
        o = np.zeros(6, dtype=np.int64)
        self._ck(self.L.hsmc_gpu_counters(self.h, _ptr(o)))
        return o

    def reset_counters(self):
        self._ck(self.L.hsmc_gpu_reset_counters(self.h))

    def add_vol_move(self, accepted):
        self._ck(self.L.hsmc_gpu_add_vol_move(self.h, int(bool(accepted))))

    def cell_rejects(self):
        o = C.c_int64(0)
        self._ck(self.L.hsmc_gpu_cell_rejects(self.h, C.byref(o)))
        return o.value

    def profile(self, enable=True):
        self._ck(self.L.hsmc_gpu_profile(self.h, int(bool(enable))))

    def profile_read(self):
        """{bucket: (milliseconds, launch groups)} since the last read; buckets sweep/build/halo/other."""
        ms = np.zeros(4)
        n = np.zeros(4, dtype=np.int64)
        self._ck(self.L.hsmc_gpu_profile_read(self.h, _ptr(ms), _ptr(n)))
        names = ("sweep", "build", "halo", "other")
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(names)}

    def set_sweep_counter(self, n):
        self._ck(self.L.hsmc_gpu_set_sweep_counter(self.h, int(n)))

    # ---- parity entry points ----
    def trial_verdicts(self, idx, xyz, sf=1.0):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        f = np.zeros(idx.shape[0], dtype=np.int32)
        self._ck(self.L.hsmc_gpu_trial_verdicts(self.h, idx.shape[0], _ptr(idx), _ptr(xyz), float(sf), _ptr(f)))
        return f

    def widom_verdicts(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        f = np.zeros(xyz.shape[0], dtype=np.int32)
        self._ck(self.L.hsmc_gpu_widom_verdicts(self.h, xyz.shape[0], _ptr(xyz), _ptr(f)))
        return f

    def sweep_nvt_logged(self, dr_max):
        """One sweep; returns the trial log sorted into a serial order the reference can replay."""
        cap = 2 * self.N + 1024
        log = np.zeros(cap, dtype=TRIAL_DTYPE)
        n = C.c_int64(0)
        self._ck(self.L.hsmc_gpu_sweep_nvt_logged(self.h, float(dr_max), _ptr(log), cap, C.byref(n)))
        log = log[: n.value]
        return log[np.argsort(log["seq"], kind="stable")]

    def selftest_u01(self):
        n, first = C.c_uint64(0), C.c_uint32(0)
        self._ck(self.L.hsmc_gpu_selftest_u01(self.h, C.byref(n), C.byref(first)))
        return n.value, first.value

    def min_dist2(self):
        o = C.c_double(0)
        self._ck(self.L.hsmc_gpu_min_dist2(self.h, C.byref(o)))
        return o.value
