#!/usr/bin/env python
"""Benchmark of the hard-sphere MC hot path (BASELINE.json: trial moves/s and % of the
cell-list HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Workload (config.workload): NVT hard spheres, fcc start, rho = 0.9, N = 16 777 216
(fcc 256 x 128 x 128 unit cells; the long axis is x, the slab axis), i.e. BASELINE.json
configs[3] -- the configuration the headline metric is quoted on; it fits one B200, and
the same total system is slab-decomposed over N GPUs ("scaling": "strong").

A step = one hsmc_gpu_sweep_nvt() call of S sweeps (each sweep = one grid shift + cell-list
rebuild, one k_propose launch that generates the sweep's proposals,
and N trial moves = 8 checkerboard block phases, each running the 8 cell colours of its blocks
inside the CTA).  The block phases that need no halo exchange between them run as ONE
k_sweep_lean launch (all 8 on one GPU, 0-3 and 4-7 on slabs), so a "launch" of the roofline
object is a whole sweep (N moves) at N=1 and half a sweep on slabs.  S (--sweeps-per-step,
default: chosen after the warm-up so that the timed region lasts >= 2.5 s) is reported in
config.  `value` is timed on the device (CUDA events on the handle's stream) with the
configuration resident in HBM; `e2e` is the same call driven from pinned HOST buffers: upload
of the {id,x,y,z} table, the sweeps, download of the table and the move counters, wall-clock.
At N > 1 the line also carries `chain_identical`: after the timed region the slabs' state is
gathered, rank 0 runs the same sweeps on ONE GPU (told to use the N-slab block partition) and
the coordinates (checksum of every row) and counters must be equal.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RHO = 0.9
DR_MAX = 0.1
FALLBACK_HBM_GBS = 6650.0
L2_BYTES = 126e6


# --------------------------------------------------------------------------------------
# synthetic input: the reference's own lattice generator (sim_info.c:125-166), vectorised
# --------------------------------------------------------------------------------------
def fcc_lattice(nx, ny, nz, rho, out=None):
    a = (4.0 / rho) ** (1.0 / 3.0)   # pow(cell_vol, 1./3.)
    n = 4 * nx * ny * nz
    conf = out if out is not None else np.empty((n, 4))
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ii, jj, kk = ii.ravel().astype(np.float64), jj.ravel().astype(np.float64), kk.ravel().astype(np.float64)
    v = conf.reshape(nx * ny * nz, 4, 4)
    for b, (ox, oy, oz) in enumerate(((0, 0, 0), (0.5, 0.5, 0), (0.5, 0, 0.5), (0, 0.5, 0.5))):
        v[:, b, 1] = (ii + ox) * a
        v[:, b, 2] = (jj + oy) * a
        v[:, b, 3] = (kk + oz) * a
    conf[:, 0] = np.arange(n)
    return np.array([nx * a, ny * a, nz * a]), conf


# --------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per sweep-phase launch from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "sweep_phase_dram_bytes.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


# --------------------------------------------------------------------------------------
# CPU arm: the reference's own sweep_nvt()/part_move() on host cores
# --------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One replica: the reference's own lattice generator + cell list for a cubic fcc `cells`^3 box, then
    `moves` part_move() calls per step (sweep_nvt() is N of them, nvt.c:201-209)."""
    kind, cells, seed, steps, warmup, moves, conn = args
    from oracle import pyoracle
    if kind == "reference":
        r = pyoracle.Ref(lattice=(2, cells, cells, cells, RHO), neigh_dr=1.0, max_part=10, seed=seed)
        r.set_moves(dr_max=DR_MAX)
        run = lambda: r.part_moves(moves)
        n = r.N
    else:
        box, conf = pyoracle.Port.lattice(2, cells, cells, cells, RHO)
        p = pyoracle.Port(conf, box, neigh_dr=1.0, max_part=10)
        n = conf.shape[0]
        state = {"k": 0}

        def run():
            # the restatement only exposes whole sweeps: scale the count afterwards
            p.sweep_nvt(max(1, int(round(moves / n))), DR_MAX, seed + state["k"])
            state["k"] += 1
        moves = max(1, int(round(moves / n))) * n
    times = []
    for s in range(warmup + steps):
        conn.send("ready")
        conn.recv()               # lock-step start so replicas overlap
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
    conn.send(("done", n, times[warmup:], moves))


def cpu_replicas_that_fit(cells, cores):
    """Replicas of the reference that fit the host's memory: its cell matrices cost ~250 bytes per particle at
    rho 0.9 / neigh_list 1.0 (one malloc per cell row, cell_list.c:33-58), 4.2 GB at N = 17 M."""
    need = 4 * cells ** 3 * 260.0
    avail = None
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = float(ln.split()[1]) * 1024.0
        for pth in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
            if os.path.exists(pth):
                v = open(pth).read().strip()
                if v.isdigit():
                    avail = min(avail, float(v)) if avail else float(v)
    except Exception:
        pass
    if not avail:
        return max(1, min(cores, 4))
    return int(max(1, min(cores, (0.4 * avail) // need)))


def cpu_reference_arm(steps, warmup, cells, moves_per_step, replicas):
    """K steps; in each, every replica (one per host core used) makes `moves_per_step` trial moves on its own
    cubic fcc `cells`^3 box at rho 0.9 through the reference's own part_move()."""
    import multiprocessing as mp
    from oracle import pyoracle
    kind = "reference" if pyoracle.have_ref() else "port"
    if kind == "port":
        pyoracle.build()
    ctx = mp.get_context("spawn")
    procs, conns = [], []
    for c in range(replicas):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_cpu_worker, args=((kind, cells, 1000 + c, steps, warmup, moves_per_step, b),))
        p.start()
        procs.append(p); conns.append(a)
    for s in range(warmup + steps):
        for c in conns:
            assert c.recv() == "ready"
        for c in conns:
            c.send("go")
    results = [c.recv() for c in conns]
    for p in procs:
        p.join()
    n, moves = results[0][1], results[0][3]
    per_step = np.max(np.array([r[2] for r in results]), axis=0)   # slowest replica per step
    total = float(per_step.sum())
    return {
        "value": moves * replicas * steps / total, "ms_per_step": 1e3 * total / steps, "kind": kind, "cores": replicas,
        "N": n, "moves_per_step_per_replica": moves,
        "sample": (f"{replicas} independent replica(s), one per host core, each {moves} part_move() calls per step "
                   f"({moves / n:.3g} sweep) on a cubic fcc {cells}^3 box, N={n}, rho={RHO}, dr_max={DR_MAX}, neigh_list 1.0, "
                   f"through the reference's own part_move()/check_overlap() "
                   f"({'unmodified sources, oracle/_ref' if kind == 'reference' else 'oracle C restatement'}); "
                   f"{steps} timed step(s) after {warmup} warm-up step(s), slowest replica per step"),
    }


# --------------------------------------------------------------------------------------
# secondary kernels (insertions/s, pairs/s, particles/s), device-timed
# --------------------------------------------------------------------------------------
def _timed(stream, fn, reps=3):
    """best-of-`reps` device time (ms) of one ABI call, CUDA events on the handle's stream"""
    import torch
    fn()                                      # warm-up
    best = float("inf")
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def secondary_observables(h, stream, N, nbar, peak, device):
    """Throughput of the observables' kernels on the benchmark configuration (resident in HBM,
    larger than L2), each with SURVEY 8(d)'s algorithmic bytes per unit (16 -> 32: these kernels
    read the double4 table; the pair kernels walk the forward HALF stencil, own cell + 13
    neighbours, so 14 cells per particle, not 27) against the same HBM peak.  RDF is O(N^2) on shared-memory tiles (ALU / shared
    atomics bound): measured on the C2 shape (fcc 20^3, N = 32 000) in pairs/s, no HBM
    fraction.  q_l is measured on a fcc 64^3 box whose cells are wide enough (1.5) to hold
    the first neighbour shell, as the reference's `ql` keyword needs."""
    import torch
    import hsmc_b200
    out = {}

    def entry(name, unit, units, ms, bytes_per_unit, what):
        rate = units / (ms * 1e-3)
        e = {"value": rate, "unit": unit, "ms": ms, "units_per_call": units, "what": what}
        if bytes_per_unit:
            gbs = rate * bytes_per_unit / 1e9
            e.update(algorithmic_bytes_per_unit=bytes_per_unit, achieved_gbs=gbs, frac_of_hbm_peak=gbs / peak)
        out[name] = e

    M = 100_000_000
    ms = _timed(stream, lambda: h.widom(7, M))
    entry("widom", "insertions/s", M, ms, 32.0 * 27.0 * nbar, "hsmc_gpu_widom, 1e8 insertion points (k_widom)")
    ms = _timed(stream, lambda: h.overlap_scaled(1.0))
    entry("overlap_scaled", "particles/s", N, ms, 32.0 * (14.0 * nbar + 1.0),
          "hsmc_gpu_overlap_scaled(sf=1.0, no overlap found = every pair visited): the NpT volume-move verdict (k_overlap_scaled_f32)")
    sf = (1.0 - 0.0001 * (np.arange(20) + 1.0)) ** (1.0 / 3.0)
    ms = _timed(stream, lambda: h.presst_flags(sf))
    entry("presst_flags", "particles/s", N, ms, 32.0 * (14.0 * nbar + 1.0),
          "hsmc_gpu_presst_flags, 20 compressions in one pass (k_overlap_scaled_f32)")
    dr_c = min(0.002, 0.9 * (min(h.info()["cell_size"]) - 1.0))       # one bin that still fits the cell edge
    ms = _timed(stream, lambda: h.contact_counts(dr_c, 1))
    entry("contact_counts", "particles/s", N, ms, 32.0 * (14.0 * nbar + 1.0),
          f"hsmc_gpu_contact_counts(dr={dr_c:.5f}, one bin below the cell edge) (k_contact_hist_f32)")

    # RDF on the C2 shape
    box2, conf2 = fcc_lattice(20, 20, 20, RHO)
    with hsmc_b200.HsmcGpu(conf2.shape[0], box2, seed=3, device=device) as h2:
        h2.upload(conf2)
        h2.sweep_nvt(50, DR_MAX)
        s2 = torch.cuda.ExternalStream(h2.stream_ptr(), device=torch.device("cuda", device))
        nn = int((5.0 - 1.0) / 0.01)
        ms = _timed(s2, lambda: h2.rdf_counts(0.01, nn))
        n2 = conf2.shape[0]
        entry("rdf", "pairs/s", n2 * (n2 - 1) // 2, ms, None,
              "hsmc_gpu_rdf_counts(dr=0.01, rmax=5.0) at N=32000 (k_rdf_pairs); ALU/shared-atomic bound, no HBM fraction")
    # q_l on cells that hold the first shell
    box3, conf3 = fcc_lattice(64, 64, 64, RHO)
    with hsmc_b200.HsmcGpu(conf3.shape[0], box3, seed=4, device=device, cell_min=1.5) as h3:
        h3.upload(conf3)
        h3.sweep_nvt(5, DR_MAX)
        s3 = torch.cuda.ExternalStream(h3.stream_ptr(), device=torch.device("cuda", device))
        i3 = h3.info()
        nb3 = conf3.shape[0] / (i3["cells"][0] * i3["cells"][1] * i3["cells"][2])
        rmax = min(1.5, min(i3["cell_size"]))
        ms = _timed(s3, lambda: h3.order_parameter(6, rmax))
        entry("order_parameter", "particles/s", conf3.shape[0], ms, 32.0 * (27.0 * nb3 + 1.0),
              f"hsmc_gpu_order_parameter(l=6, rmax={rmax:.3f}) at N={conf3.shape[0]}, cells >= 1.5 (k_order_param; fp64 ALU heavy)")
    return out


# --------------------------------------------------------------------------------------
# BASELINE configs[4]: Widom chemical-potential sweep, insertions sharded over the GPUs
# --------------------------------------------------------------------------------------
def shard_range(total, rank, world):
    """[first, first + count) of `rank`: equal shares, the last rank takes the remainder"""
    share = total // world
    first = rank * share
    return first, (share if rank < world - 1 else total - first)


def widom_workload(args, rank, world, local_rank):
    """`--workload widom`: the C2 shape (fcc 20^3, N = 32 000) at rho = 0.3 ... 0.9, equilibrated on the
    device; a step = one sample of --insertions (1e8) trial insertions at EACH density.  The
    configuration is replicated (1 MB), the insertion index range is cut into WORLD_SIZE shares
    (hsmc_gpu_widom(first, count, reduce = 0)), the accepted counts are summed by the caller -- no
    data-path collective.  One JSON line: insertions/s over all GPUs, device-timed, max over
    ranks, with mu_ex per density."""
    import torch
    import torch.distributed as dist
    import hsmc_b200
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
    rhos = [0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]
    M = int(args.insertions)
    first, count = shard_range(M, rank, world)
    handles, prep = [], []
    for rho in rhos:
        box, conf = fcc_lattice(20, 20, 20, rho)
        h = hsmc_b200.HsmcGpu(conf.shape[0], box, seed=20261017, device=local_rank)      # same chain on every rank
        h.upload(conf)
        # melt the fcc start and equilibrate: sweeps until the Bragg peaks of the start lattice are gone -- the
        # structure factor S(G)/N = |sum_j exp(i G r_j)|^2 / N^2 at the {111} and {200} reciprocal vectors of the
        # 20^3 fcc cells is ~0.8 for the crystal and O(1/N) for a fluid (the reference's q_l is an average of
        # per-particle values and stays near 0.35 in a dense fluid, so it cannot tell).  Every rank runs the same
        # deterministic chain, so every rank holds the same fluid
        dr_eq = min(0.5, 0.08 / rho ** 2)
        gvecs = 2.0 * np.pi * 20.0 / np.asarray(box[:3]) * np.array(
            [[1, 1, 1], [-1, 1, 1], [1, -1, 1], [1, 1, -1], [2, 0, 0], [0, 2, 0], [0, 0, 2]], dtype=np.float64)

        def bragg(rows):
            ph = rows[:, 1:4] @ gvecs.T
            return float(np.max(np.cos(ph).sum(axis=0) ** 2 + np.sin(ph).sum(axis=0) ** 2) / rows.shape[0] ** 2)

        sg, eq_sweeps = 1.0, 0
        while sg > 0.005 and eq_sweeps < args.widom_max_eq_sweeps:
            h.sweep_nvt(1000, dr_eq)
            eq_sweeps += 1000
            sg = bragg(h.download())
        h.sweep_nvt(2000, dr_eq)
        prep.append({"rho": rho, "bragg_peak_S_over_N": sg, "equilibration_sweeps": eq_sweeps + 2000, "melted": bool(sg <= 0.005)})
        info = h.info()
        nbar = conf.shape[0] / (info["cells"][0] * info["cells"][1] * info["cells"][2])
        handles.append((rho, h, nbar, torch.cuda.ExternalStream(h.stream_ptr(), device=torch.device("cuda", local_rank))))
    sampler = ClockSampler(local_rank)
    sampler.start()
    for w in range(args.warmup):
        for rho, h, nbar, st in handles:
            h.widom(w, count, first=first, reduce=False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    acc = np.zeros((args.steps, len(rhos)), dtype=np.int64)
    ms = 0.0
    launches0 = sum(h.info()["kernel_launches"] for _, h, _, _ in handles)
    for s in range(args.steps):
        for k, (rho, h, nbar, st) in enumerate(handles):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            acc[s, k] = h.widom(1000 + s, count, first=first, reduce=False)
            b.record(st)
            b.synchronize()
            ms += a.elapsed_time(b)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = sum(h.info()["kernel_launches"] for _, h, _, _ in handles) - launches0
    t_acc, t_ms, t_l = torch.from_numpy(acc.copy()), torch.tensor([ms], dtype=torch.float64), torch.tensor([launches])
    if world > 1:
        dist.barrier()
        dist.all_reduce(t_acc)
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_l)
    acc, ms = t_acc.numpy(), float(t_ms[0])
    total = M * len(rhos) * args.steps
    value = total / (ms * 1e-3)
    peak, peak_src = measured_peak()
    # k_widom reads up to 27 cells of double4 per insertion (early exit on the first overlap)
    bytes_per = float(np.mean([32.0 * 27.0 * nb for _, _, nb, _ in handles]))
    frac_acc = acc.mean(axis=0) / M
    with np.errstate(divide="ignore"):
        mu = np.where(frac_acc > 0, -np.log(frac_acc), 0.0)                 # compute_widom_chem_pot.c:62-68
    for _, h, _, _ in handles:
        h.close()
    if rank == 0:
        line = {
            "metric": "widom_insertions_per_sec", "value": value, "unit": "insertions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Widom mu_ex sweep, fcc 20^3 start (N=32000), rho=0.3..0.9, {M} insertions per sample "
                                   f"and density, insertion range sharded over {world} GPU(s), configuration replicated",
                       "rhos": rhos, "insertions_per_sample": M, "l2": "table fits L2 (1 MB): L2-resident by construction"},
            "clocks": clocks, "gpu_launches": int(t_l[0]),
            "roofline": {"bound": "hbm", "kernel": "k_widom", "achieved": value / world * bytes_per / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": value / world * bytes_per / 1e9 / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_insertion": bytes_per, "traffic": None,
                         "note": "no-reuse figure of SURVEY 8(d); the 1 MB table is cache-resident, so the kernel is "
                                 "bound by issue/latency, not HBM, and frac can exceed 1"},
            "mu_ex": {f"{r:.1f}": float(m) for r, m in zip(rhos, mu)},
            "mu_ex_carnahan_starling": {f"{r:.1f}": float((8 * e - 9 * e * e + 3 * e ** 3) / (1 - e) ** 3)
                                        for r, e in ((r, np.pi * r / 6) for r in rhos)},
            "preparation": prep,
            "accepted_fraction": {f"{r:.1f}": float(f) for r, f in zip(rhos, frac_acc)},
        }
        args.emit(line)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------
# BASELINE configs[1] / configs[2]: whole runs of the input file through the host driver (and, as the reference arm,
# through the reference executable on the IDENTICAL file): moves/s by the reference's own accounting
# (`Particle moves` / `Elapsed time`, nvt.c:65-95), sampling hooks included
# --------------------------------------------------------------------------------------
CONFIG_INPUTS = {
    "c2": ("NVT rho=0.9, N=32000 (fcc 20^3), pressure (virial + thermodynamic), RDF and Widom sampling", """rho 0.9
cells_x 20
cells_y 20
cells_z 20
type 2
neigh_list 1.05 10
dr_max 0.1
opt 1 100 10 0.5 0.5
press_virial 0.002 20
press_thermo 0.0001 0.002 20
rdf 0.01 5.0 50 100
widom 1000 20
seed 7
sweep_eq 100
sweep_stat 200
out 50
"""),
    "c3": ("NpT P=10 from rho=0.94, N=108000 (fcc 30^3), optimizer-tuned moves, thermodynamic pressure sampling", """npt 10 0.001
rho 0.94
cells_x 30
cells_y 30
cells_z 30
type 2
neigh_list 1.1 12
dr_max 0.05
opt 1 100 10 0.5 0.5
press_thermo 0.0001 0.002 20
seed 99
sweep_eq 60
sweep_stat 100
out 20
"""),
}


def _run_driver(exe, text, extra=()):
    import tempfile
    d = tempfile.mkdtemp(prefix="hsmc_bench_cfg_")
    with open(os.path.join(d, "in.dat"), "w") as f:
        f.write(text)
    t0 = time.perf_counter()
    r = subprocess.run([exe, "-o", "out.txt", *extra], cwd=d, capture_output=True, text=True, timeout=1800,
                       env=dict(os.environ, HSMC_REPORT_LAUNCHES="1"))
    wall = time.perf_counter() - t0
    log = open(os.path.join(d, "out.txt")).read() if os.path.exists(os.path.join(d, "out.txt")) else ""
    if r.returncode != 0 or "Simulation complete!" not in log:
        raise RuntimeError((r.stdout + r.stderr + log)[-1500:])
    moves = float(log.split("-- Particle moves:")[1].split()[0])
    elapsed = float(log.split("Elapsed time:")[1].split()[0])
    return moves, elapsed, wall, log + r.stderr


def _run_reference(ref_exe, text):
    """The unmodified reference on the identical input; if it crashes there (its NpT optimizer segfaults at N = 108 000,
    input `opt 1`, in this container and on the GPU box alike), the same input without the optimizer stage -- the
    sweeps, the sampling and the sizes are unchanged, the step sizes stay at the input's initial values."""
    try:
        return _run_driver(ref_exe, text) + (None,)
    except RuntimeError as e:
        if "opt 1" not in text:
            raise
        import re
        alt = re.sub(r"^opt 1 ", "opt 0 ", text, flags=re.M)
        note = ("the unmodified reference crashes on the identical input (NpT optimizer, `opt 1`, at this N); timed on the same "
                "input with `opt 0`: same sweeps, sampling and sizes, step sizes left at the input's initial values")
        return _run_driver(ref_exe, alt) + (note,)


def config_workload(args, rank, world, local_rank):
    if rank != 0:
        return
    name, text = CONFIG_INPUTS[args.workload]
    cfg = {"workload": f"BASELINE configs[{1 if args.workload == 'c2' else 2}]: {name}; whole run of the input file below, sampling hooks at "
                       "the input's intervals", "input": text, "same_config": True}
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "hsmc_ref")
    if args.impl == "reference":
        if not os.path.exists(ref_exe):
            args.emit({"impl": "reference", "unavailable": "oracle/_ref/hsmc_ref not built (no /root/reference on this box and no prebuilt copy)"})
            return
        moves, elapsed, wall, _, note = _run_reference(ref_exe, text)
        if note:
            cfg = dict(cfg, same_config=False, reference_note=note)
        args.emit({"impl": "reference", "metric": "hard_sphere_trial_moves_per_sec", "value": moves / elapsed, "unit": "moves/s",
                   "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": 1e3 * elapsed, "higher_is_better": True,
                   "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                   "cpu_baseline": {"value": moves / elapsed, "unit": "moves/s", "cores": 1, "kind": "reference",
                                    "sample": "the unmodified reference executable (oracle/_ref/hsmc_ref, serial) on the identical input, once"},
                   "e2e": {"value": moves / wall, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   "gpu_launches": 0})
        return
    from hsmc_b200 import build
    exe = build.build_host()
    extra = ("-g", str(args.gpus)) if args.gpus > 1 else ()
    sampler = ClockSampler(local_rank)
    for _ in range(max(1, args.warmup // 3)):
        _run_driver(exe, text, extra)
    sampler.start()
    runs = [_run_driver(exe, text, extra) for _ in range(args.steps)]
    clocks = sampler.stop()
    moves = runs[0][0]
    el = sum(r[1] for r in runs) / len(runs)
    wall = sum(r[2] for r in runs) / len(runs)
    launches = None
    for ln in runs[-1][3].splitlines():
        if "kernel launches" in ln.lower():
            launches = int(float(ln.split()[-1]))
    cpu = None
    if not args.no_cpu_baseline and os.path.exists(ref_exe):
        m_r, e_r, w_r, _, note = _run_reference(ref_exe, text)
        cpu = {"value": m_r / e_r, "unit": "moves/s", "cores": 1, "kind": "reference", "same_config": note is None,
               "sample": "the unmodified reference executable (oracle/_ref/hsmc_ref, serial) on the identical input file, once"
                         + ("" if note is None else "; " + note),
               "elapsed_s": e_r, "moves": m_r}
    # the sweep kernels at this shape, timed in-process (launch-latency-bound at these sizes)
    roofline = config_roofline(args, text, local_rank)
    args.emit({"metric": "hard_sphere_trial_moves_per_sec", "value": moves / el, "unit": "moves/s", "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": dict(cfg, value_is="Particle moves / Elapsed time as the driver prints them (equilibration + production, "
                                            "sampling and output files included; the reference's own accounting)"),
               "clocks": clocks,
               "e2e": {"value": moves / wall, "unit": "moves/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                       "what": "the same run by the wall clock of the whole process (CUDA context, upload, output files)",
                       "wall_s": wall},
               "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu})


def config_roofline(args, text, device):
    """Sweep path at the shape of a config input, in-process: ms per sweep by profile bucket and the SURVEY 8(d) figure."""
    try:
        import hsmc_b200
        kv = dict(ln.split(None, 1) for ln in text.splitlines() if ln.strip() and not ln.startswith("#"))
        nx = int(kv["cells_x"]); rho = float(kv["rho"]); cell_min = float(kv["neigh_list"].split()[0]); dr = float(kv["dr_max"])
        box, conf = fcc_lattice(nx, nx, nx, rho)
        N = conf.shape[0]
        with hsmc_b200.HsmcGpu(N, box, seed=1, device=device, cell_min=cell_min) as h:
            h.upload(conf)
            h.sweep_nvt(200, dr)
            h.sync()
            h.profile(True); h.profile_read()
            S = 2000
            t0 = time.perf_counter()
            h.sweep_nvt(S, dr)
            h.sync()
            wall = time.perf_counter() - t0
            p = h.profile_read()
            info = h.info()
        nbar = N / (info["cells"][0] * info["cells"][1] * info["cells"][2])
        b_move = 16.0 * (27.0 * nbar + 2.0)
        peak, peak_src = measured_peak()
        k_ms = p["sweep"][0] / S
        return {"bound": "hbm (launch latency at this size: the whole table fits L2)", "kernel": "k_sweep_lean",
                "achieved": N * b_move / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": N * b_move / (k_ms * 1e-3) / 1e9 / peak,
                "peak_source": peak_src, "algorithmic_bytes_per_move": b_move, "nbar": nbar, "traffic": None,
                "ms_per_sweep": {k: v[0] / S for k, v in p.items()}, "wall_ms_per_sweep": 1e3 * wall / S,
                "moves_per_s_sweeps_only": N * S / wall,
                "what": f"{S} plain sweeps of the same shape (N={N}) after 200 warm-up sweeps, CUDA events per launch group"}
    except Exception as e:      # reported, never required
        return {"error": repr(e)}


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sweeps-per-step", type=int, default=0, help="0: chosen so that the timed region lasts >= --min-seconds")
    ap.add_argument("--min-seconds", type=float, default=2.5)
    ap.add_argument("--chain-check-sweeps", type=int, default=8, help="N > 1: sweeps of the 1-GPU identity check (0 = skip)")
    ap.add_argument("--cells", type=int, nargs=3, default=[162, 162, 162],
                    help="fcc unit cells (x is the slab axis); the default cubic 162^3 (N = 17 006 112) is the C4 shape the reference can run too (SURVEY 8d)")
    ap.add_argument("--regrid", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-cells", type=int, default=64, help="CPU arm box when --cells is not cubic (the reference mis-indexes non-cubic boxes)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--cpu-moves-per-step", type=int, default=200000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep-impl", type=int, default=0)
    ap.add_argument("--no-secondary", action="store_true", help="skip the observables' kernel timings (N=1 only)")
    ap.add_argument("--workload", default="sweep", choices=["sweep", "widom", "c2", "c3"],
                    help="sweep: the headline metric (default, BASELINE configs[3]); widom: configs[4], insertions sharded over the "
                         "GPUs; c2 / c3: configs[1] / configs[2] run through the drop-in host driver on their input files "
                         "(the reference arm runs the reference executable on the identical file)")
    ap.add_argument("--insertions", type=float, default=1e8, help="--workload widom: insertions per sample and density")
    ap.add_argument("--widom-max-eq-sweeps", type=int, default=60000, help="--workload widom: give up melting a density after this many sweeps (it is then flagged)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: whatever native libraries print on fd 1 while the benchmark runs
    # (NCCL's version banner under NCCL_DEBUG=VERSION, say) is sent to stderr; emit() restores fd 1 for the line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    args.emit = emit
    if args.workload == "widom" and args.impl == "b200":
        return widom_workload(args, rank, world, local_rank)
    if args.workload in ("c2", "c3"):
        return config_workload(args, rank, world, local_rank)
    nx, ny, nz = args.cells
    N = 4 * nx * ny * nz
    workload = {
        "workload": f"NVT hard spheres, fcc {nx}x{ny}x{nz} start, N={N}, rho={RHO}, dr_max={DR_MAX}",
        "N": N, "rho": RHO, "dr_max": DR_MAX, "sweeps_per_step": args.sweeps_per_step,
        "regrid_interval": args.regrid, "positions": "double4 {x,y,z,id}, cell-ordered",
    }
    workload_extra = {}

    if args.impl == "reference":
        if rank != 0:
            return
        cores = len(os.sched_getaffinity(0))
        same = nx == ny == nz
        ccells = nx if same else args.cpu_cells
        reps = cpu_replicas_that_fit(ccells, cores)
        res = cpu_reference_arm(args.steps, args.warmup, ccells, args.cpu_moves_per_step, reps)
        cfg = dict(workload, same_config=same, replicas=reps, host_cores=cores,
                   moves_per_step_per_replica=res["moves_per_step_per_replica"])
        for k in ("positions", "sweeps_per_step", "regrid_interval"):
            cfg.pop(k, None)
        if same:
            cfg["workload"] += (f" -- reference arm: {reps} independent replica(s) of this same system, one per host core "
                                f"(as many as fit the host memory), {res['moves_per_step_per_replica']} part_move() calls each per step")
        else:
            # the reference mis-indexes non-cubic boxes: say what is run instead
            cfg["b200_arm_workload"] = workload["workload"]
            cfg["workload"] = (f"NVT hard spheres, {reps} independent replicas of a CUBIC fcc {ccells}^3 start, N={res['N']} each, "
                               f"rho={RHO}, dr_max={DR_MAX}, neigh_list 1.0 (NOT the b200 arm's non-cubic box)")
            cfg["N"] = res["N"]
        line = {
            "impl": "reference", "metric": "hard_sphere_trial_moves_per_sec", "value": res["value"], "unit": "moves/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": res["value"], "unit": "moves/s", "cores": res["cores"], "kind": res["kind"],
                             "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit(line)
        return

    import torch
    import torch.distributed as dist
    import hsmc_b200
    from hsmc_b200 import gpu as G

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
        ids = [G.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]
    else:
        nccl_id = None

    def barrier():
        if world > 1:
            dist.barrier()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    def allsum_u64(x, dist, torch):
        """sum modulo 2^64 over the ranks"""
        parts = [None] * world
        dist.all_gather_object(parts, int(x))
        return sum(parts) % (1 << 64)

    box, conf = fcc_lattice(nx, ny, nz, RHO)
    h = hsmc_b200.HsmcGpu(N, box, seed=20261017, device=local_rank, rank=rank, world=world, nccl_id=nccl_id,
                          cell_min=1.0, regrid_interval=args.regrid, sweep_impl=args.sweep_impl)
    halo = "single GPU"
    if world > 1:
        halo = "nccl send/recv"
        if os.environ.get("HSMC_P2P", "1") == "1":
            # NVLink peer-to-peer halo path: every rank attaches its neighbours' receive windows
            blobs = [None] * world
            dist.all_gather_object(blobs, h.ipc_export())
            h.ipc_attach(blobs[(rank - 1) % world], blobs[(rank + 1) % world])
            dist.barrier()
            halo = "NVLink peer-to-peer windows (kernels store into the neighbour's HBM, sequence flags)"
    h.upload(conf)
    info = h.info()
    del conf
    n_owned0 = info["n_owned"]
    # pinned host mirror of this rank's rows, used by the end-to-end leg
    cap_rows = int(n_owned0 * 1.3) + 4096 if world > 1 else N
    host = torch.empty((cap_rows, 4), dtype=torch.float64, pin_memory=True)
    stream = torch.cuda.ExternalStream(h.stream_ptr(), device=torch.device("cuda", local_rank))
    S = args.sweeps_per_step
    if S <= 0:
        # sweeps per step: long enough a timed region for the clock sampler (>= 20 samples at 100 ms) at every N
        h.sweep_nvt(5, DR_MAX)
        h.sync()
        barrier()
        t0 = time.perf_counter()
        h.sweep_nvt(10, DR_MAX)
        h.sync()
        t_sweep = allmax(time.perf_counter() - t0) / 10.0
        S = int(min(5000, max(10, np.ceil(args.min_seconds / (args.steps * t_sweep)))))
    workload["sweeps_per_step"] = S

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- warm-up ----
    for _ in range(args.warmup):
        h.sweep_nvt(S, DR_MAX)
    h.sync()
    h.reset_counters()
    h.profile(True)
    h.profile_read()
    l0 = h.info()["kernel_launches"]

    # ---- timed region: K steps, device-timed per step on the handle's stream ----
    resident = 2 * info["n_local"] * 32
    flush = None
    if resident < 2 * L2_BYTES:
        flush = torch.empty(int(3 * L2_BYTES), dtype=torch.uint8, device=f"cuda:{local_rank}")
    barrier()
    torch.cuda.synchronize()
    ev = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        if flush is not None:
            with torch.cuda.stream(stream):
                flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        h.sweep_nvt(S, DR_MAX)
        b.record(stream)
        ev.append((a, b))
    h.sync()
    torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    ms_local = sum(a.elapsed_time(b) for a, b in ev)
    ms_total = allmax(ms_local)
    prof = h.profile_read()
    h.profile(False)
    launches = allsum(h.info()["kernel_launches"] - l0)
    cnt = h.counters()           # all-reduced over ranks by the library
    moves = int(cnt[0])
    assert moves == N * S * args.steps, (moves, N * S * args.steps)
    value = moves / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (one launch = the fused block phases: N trial moves on one
    #      GPU, n_owned/2 on a slab; HSMC_FUSE=0 makes it one block phase = N/8) ----
    ncell = info["cells"][0] * info["cells"][1] * info["cells"][2]
    nbar = N / ncell
    # Algorithmic bytes per trial move, no reuse credit: SURVEY 8(d)'s figure, 16*(27*nbar + 2) (full 27-cell
    # stencil gather of 16-byte entries + own read + own write) -- the headline `frac`.  What this layout moves
    # per move without reuse credit is reported beside it: the stencil is gathered from 12-byte fp32 pair-records,
    # the trial slot is a 16-byte record, an accepted move copies a 32-byte proposal into the master table
    # (24 bytes written) and 12 bytes of shadow.
    b_move = 16.0 * (27.0 * nbar + 2.0)
    acc_now = float(cnt[1]) / float(cnt[0])
    b_move_layout = 12.0 * 27.0 * nbar + 16.0 + acc_now * (32.0 + 24.0 + 12.0)
    sweep_ms, sweep_groups = prof["sweep"]
    plan_ms, plan_groups = prof["other"]
    moves_local = info["n_owned"] * S * args.steps     # this rank's trial moves (N/world up to migration)
    per_launch_s = (sweep_ms * 1e-3) / max(sweep_groups, 1)
    achieved = (moves_local / max(sweep_groups, 1)) * b_move / per_launch_s / 1e9
    peak, peak_src = measured_peak()
    traffic = ncu_traffic()
    kname = {0: "k_sweep_lean", 3: "k_sweep_lean", 5: "k_sweep_lean (global-memory path)"}.get(args.sweep_impl & 0xff, "k_sweep_phase")
    roofline = {
        "bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "peak_source": peak_src,
        "algorithmic_bytes_per_move": b_move, "nbar": nbar, "moves_per_launch": moves_local / max(sweep_groups, 1),
        "algorithmic_bytes_definition": "SURVEY 8(d): 16*(27*nbar + 2), no reuse credit",
        "layout_bytes_per_move": b_move_layout,
        "layout_bytes_definition": "12*27*nbar (stencil from fp32 pair-records) + 16 (trial record) + acceptance*(32 proposal read + 24 master + 12 shadow written)",
        "frac_with_layout_bytes": achieved / b_move * b_move_layout / peak,
        # the whole step (rebuild + plan + sweep + halo) against the same figure: value * bytes / (n_gpus * peak)
        "step_frac": value * b_move / 1e9 / (world * peak),
        "avg_launch_ms": per_launch_s * 1e3, "launches_timed": sweep_groups,
        "kernel_share_of_step": sweep_ms / ms_local if ms_local > 0 else None,
        "plan_share_of_step": plan_ms / ms_local if ms_local > 0 else None,
        "build_share_of_step": prof["build"][0] / ms_local if ms_local > 0 else None,
        "halo_share_of_step": prof["halo"][0] / ms_local if ms_local > 0 else None,
        "traffic": None, "traffic_source": None,
    }
    if traffic and world == 1 and os.environ.get("HSMC_FUSE", "1") != "0":
        # dram__bytes of one launch of the same kernel from the committed ncu capture; marked stale when the capture
        # is of another kernel than the one timed here
        if traffic.get("kernel", "") == kname.split(" ")[0] and traffic.get("N", N) == N:
            roofline["traffic"] = traffic.get("dram_bytes_per_launch")
            roofline["traffic_source"] = traffic.get("source")
        else:
            roofline["traffic_source"] = "stale: committed capture is of " + str(traffic.get("kernel", "k_sweep_block (round 1)"))

    if roofline["traffic"]:
        # what actually crossed the HBM interface (ncu capture of the same launch), for comparison with the
        # no-reuse algorithmic figure above: the kernel is latency-bound, not bandwidth-bound
        roofline["dram_gbs"] = roofline["traffic"] / per_launch_s / 1e9
        roofline["dram_frac"] = roofline["dram_gbs"] / peak
        roofline["dram_bytes_per_move"] = roofline["traffic"] / max(roofline["moves_per_launch"], 1)

    # ---- end to end: host buffers in, host buffers out, every step ----
    def pull():
        if world > 1:
            return h.download_owned_ptr(host.data_ptr(), cap_rows)
        h.download_ptr(host.data_ptr())
        return N
    n_rows = pull()
    h2d = d2h = 0
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        h.upload_ptr(host.data_ptr(), n_rows)
        h2d += n_rows * 32
        h.sweep_nvt(S, DR_MAX)
        n_rows = pull()
        d2h += n_rows * 32 + 48
        c2 = h.counters()
    h.sync()
    barrier()
    t_e2e = allmax(time.perf_counter() - t0)
    e2e_val = N * S * args.e2e_steps / t_e2e
    e2e = {"value": e2e_val, "unit": "moves/s", "h2d_bytes_per_step": int(allsum(h2d) / args.e2e_steps),
           "d2h_bytes_per_step": int(allsum(d2h) / args.e2e_steps), "steps": args.e2e_steps,
           "ms_per_step": 1e3 * t_e2e / args.e2e_steps,
           "what": "hsmc_gpu_upload(pinned host rows) + hsmc_gpu_sweep_nvt + hsmc_gpu_download + hsmc_gpu_counters per step"}
    # ---- N > 1: is the slab run the same Markov chain as one GPU?  (the Philox draws are keyed by global
    #      cell, the block partition is that of the N-slab run.)  The slabs' table is gathered on rank 0, which runs
    #      the same sweeps on ONE GPU; every row (order-independent 64-bit checksum) and the counters must agree.
    chain = None
    if world > 1 and args.chain_check_sweeps > 0:
        C_ = args.chain_check_sweeps
        K64 = np.array([0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0x27D4EB2F165667C5], dtype=np.uint64)

        def checksum(rows):
            with np.errstate(over="ignore"):
                return int((rows.view(np.uint64).reshape(-1, 4) * K64[None, :]).sum(dtype=np.uint64))
        n_rows = pull()
        mine = host[:n_rows].numpy()
        counts = [None] * world
        dist.all_gather_object(counts, int(n_rows))
        mx = max(counts)
        pad = torch.zeros((mx, 4), dtype=torch.float64)
        pad[:n_rows] = host[:n_rows]
        gathered = [torch.empty((mx, 4), dtype=torch.float64) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, gathered, dst=0)
        sweeps_done = h.info()["sweeps_done"]
        h.upload_ptr(host.data_ptr(), n_rows)            # (both sides restart their regrid phase from an upload)
        h.reset_counters()
        h.sweep_nvt(C_, DR_MAX)
        n_rows = pull()
        cs_slabs = int(allsum_u64(checksum(host[:n_rows].numpy()), dist, torch))
        cnt_slabs = [int(x) for x in h.counters()]
        ok = None
        if rank == 0:
            table = np.empty((N, 4))
            for r_ in range(world):
                rows = gathered[r_][:counts[r_]].numpy()
                table[rows[:, 0].astype(np.int64)] = rows
            del gathered
            with hsmc_b200.HsmcGpu(N, box, seed=20261017, device=local_rank, cell_min=1.0, regrid_interval=args.regrid,
                                   sweep_impl=args.sweep_impl, xpart_world=world) as m1:
                m1.upload(table)
                m1.set_sweep_counter(sweeps_done)
                m1.sweep_nvt(C_, DR_MAX)
                out1 = m1.download()
                cnt1 = [int(x) for x in m1.counters()]
            cs1 = checksum(out1)
            ok = bool(cs1 == cs_slabs and cnt1 == cnt_slabs)
            chain = {"identical": ok, "sweeps": C_, "rows_checksum_slabs": f"{cs_slabs:016x}", "rows_checksum_one_gpu": f"{cs1:016x}",
                     "counters_slabs": cnt_slabs[:3], "counters_one_gpu": cnt1[:3],
                     "what": f"{C_} sweeps from the state after the timed region: {world} slabs vs one GPU with the {world}-slab block partition"}
        barrier()
    acc = acc_now
    min_r2 = h.min_dist2()
    assert min_r2 >= 1.0, f"overlap after benchmark: min r^2 = {min_r2}"
    # build (cell-list) kernels: one rebuild per sweep, bytes per particle from DESIGN.md K1
    build_ms, build_groups = prof["build"]
    secondary = None
    if world == 1 and not args.no_secondary:
        secondary = {}
        if build_groups:
            rate = N / (build_ms * 1e-3 / build_groups)
            bpp = 88.0 + 12.0 / nbar
            secondary["cell_list_build"] = {
                "value": rate, "unit": "particles/s", "ms": build_ms / build_groups, "units_per_call": N,
                "algorithmic_bytes_per_unit": bpp, "achieved_gbs": rate * bpp / 1e9,
                "frac_of_hbm_peak": rate * bpp / 1e9 / peak,
                "what": "counting-sort rebuild inside the timed sweeps (k_cell_count, k_scan_*, k_cell_scatter)"}
        try:
            secondary.update(secondary_observables(h, stream, N, nbar, peak, local_rank))
        except Exception as e:      # reported, never required for the headline line
            secondary["error"] = repr(e)
    h.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        same = nx == ny == nz
        ccells = nx if same else args.cpu_cells
        # bounded sample: ~cpu_seconds of CPU work per replica (6.5 us per move per core at N = 17 M, 1.1 us at N = 1 M)
        per_move = 7.0e-6 if ccells > 100 else 1.2e-6
        steps_cpu = max(2, int(args.cpu_seconds / (args.cpu_moves_per_step * per_move)))
        try:
            reps = cpu_replicas_that_fit(ccells, cores)
            r = cpu_reference_arm(steps_cpu, 1, ccells, args.cpu_moves_per_step, reps)
            cpu = {"value": r["value"], "unit": "moves/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                   "per_core": r["value"] / max(r["cores"], 1), "same_config": same, "host_cores": cores}
            if same and reps > 1:
                # one replica alone (no contention for the host's memory bandwidth): the serial reference as its
                # author runs it, on the same system as the GPU arm
                r1 = cpu_reference_arm(max(2, steps_cpu // 2), 1, ccells, args.cpu_moves_per_step, 1)
                cpu["c4_cubic_1core"] = {"value": r1["value"], "unit": "moves/s", "cores": 1, "kind": r1["kind"],
                                         "sample": r1["sample"], "same_config": True,
                                         "sweep_seconds_extrapolated": r1["N"] / r1["value"]}
        except Exception as e:   # the baseline is reported, never required
            cpu = {"value": None, "unit": "moves/s", "cores": cores, "kind": "unavailable", "sample": repr(e)}

    if rank == 0:
        line = {
            "metric": "hard_sphere_trial_moves_per_sec", "value": value, "unit": "moves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload, cells=list(info["cells"]), cell_size=list(info["cell_size"]),
                           acceptance=acc, l2="inputs larger than L2" if flush is None else "L2 flushed between steps",
                           resident_bytes_per_rank=resident, wall_s_timed_region=t_wall, min_r2_after=min_r2, halo=halo),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        if chain is not None:
            line["chain_identical"] = chain["identical"]
            line["chain_check"] = chain
        if secondary is not None:
            line["secondary"] = secondary
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
