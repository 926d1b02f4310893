"""Slab-decomposed run vs single-GPU run: bitwise identity (launch with torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29531 tests/multi_gpu_check.py

Every rank builds the same fcc start, keeps its x-slab (+ ghost layers), runs the same
number of sweeps with NCCL halo exchange, and returns its owned rows; rank 0 repeats the
run on one GPU with the same seed and compares coordinates, counters and observables bit
for bit (the Philox stream is keyed by global cell, so the chain cannot depend on the
decomposition -- SURVEY.md section 4, item 4; the single-GPU run is told to use the
x block partition of the slab run, which is part of the chain's definition)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import hsmc_b200
    from hsmc_b200 import gpu as G
    from bench import fcc_lattice

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = [G.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    cells = [int(x) for x in os.environ.get("HSMC_CHECK_CELLS", "24,10,12").split(",")]
    sweeps = int(os.environ.get("HSMC_CHECK_SWEEPS", "20"))
    rho, dr_max, seed = 0.85, 0.15, 4242
    box, conf = fcc_lattice(*cells, rho)
    N = conf.shape[0]
    ok = True
    with hsmc_b200.HsmcGpu(N, box, seed=seed, device=lr, rank=rank, world=world, nccl_id=ids[0]) as h:
        if os.environ.get("HSMC_CHECK_P2P", "1") == "1":
            # NVLink peer-to-peer halo path: gather every rank's window blob, attach the neighbours'
            blobs = [None] * world
            dist.all_gather_object(blobs, h.ipc_export())
            h.ipc_attach(blobs[(rank - 1) % world], blobs[(rank + 1) % world])
            dist.barrier()
        h.upload(conf)
        info0 = h.info()
        h.sweep_nvt(sweeps, dr_max)
        rows = h.download_owned().copy()
        cnt = h.counters()
        wid = h.widom(3, 200000)
        con = h.contact_counts(0.002, 10) if min(h.info()["cell_size"]) >= 1.02 else None
        ovl = [h.overlap_scaled(sf) for sf in (1.0, 0.9995, 0.98)] if min(h.info()["cell_size"]) * 0.98 >= 1.0 else \
            [h.overlap_scaled(1.0)]
        mind = h.min_dist2()
        q6 = h.order_parameter(6, min(h.info()["cell_size"]))
        # second leg: re-upload only what this rank owns (the e2e pattern) and continue
        h.upload(rows)
        h.sweep_nvt(3, dr_max)
        rows2 = h.download_owned().copy()
        # third leg: an accepted NpT volume move in slab mode (cell counts unchanged), then more sweeps
        sf_v = 1.0004
        h.rescale(sf_v, np.asarray(box) * sf_v)
        ovl_v = h.overlap_scaled(1.0)
        h.sweep_nvt(2, dr_max)
        rows3 = h.download_owned().copy()
        info = h.info()
        # RDF sharded over the ranks: replicated configuration on a world-1 handle per rank
        with hsmc_b200.HsmcGpu(N, box, seed=seed, device=lr) as rep:
            rep.upload(conf)
            nn_rdf = int((min(box) / 2 - 1.0) / 0.02)
            rdf_part = torch.from_numpy(rep.rdf_counts_part(0.02, nn_rdf, rank, world).astype(np.int64))
            dist.all_reduce(rdf_part)
            rdf_full = rep.rdf_counts(0.02, nn_rdf).astype(np.int64) if rank == 0 else None
    gathered = [None] * world
    dist.gather_object((rows, rows2, info0["own_x"], rows3), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        allrows = np.concatenate([g[0] for g in gathered])
        allrows2 = np.concatenate([g[1] for g in gathered])
        allrows3 = np.concatenate([g[3] for g in gathered])
        multi3 = allrows3[np.argsort(allrows3[:, 0])]
        print("slabs:", [g[2] for g in gathered], "rows:", [len(g[0]) for g in gathered], flush=True)
        ok &= allrows.shape[0] == N and np.array_equal(np.sort(allrows[:, 0]), np.arange(N))
        multi = allrows[np.argsort(allrows[:, 0])]
        multi2 = allrows2[np.argsort(allrows2[:, 0])]
        with hsmc_b200.HsmcGpu(N, box, seed=seed, device=lr, xpart_world=world) as s:
            s.upload(conf)
            s.sweep_nvt(sweeps, dr_max)
            single = s.download()
            checks = {
                "coordinates": np.array_equal(single, multi),
                "counters": np.array_equal(s.counters(), cnt),
                "widom": s.widom(3, 200000) == wid,
                "min_dist2": s.min_dist2() == mind and mind >= 1.0,
                "order_parameter": abs(s.order_parameter(6, min(s.info()["cell_size"])) - q6) < 1e-12 and 0.0 < q6 <= 1.0,
                "overlap": [s.overlap_scaled(sf) for sf in ((1.0, 0.9995, 0.98) if len(ovl) == 3 else (1.0,))] == ovl,
            }
            if con is not None:
                checks["contact"] = np.array_equal(s.contact_counts(0.002, 10), con)
            s.upload(single)
            s.set_sweep_counter(sweeps)
            s.sweep_nvt(3, dr_max)
            checks["reupload_continue"] = np.array_equal(s.download(), multi2)
            s.rescale(sf_v, np.asarray(box) * sf_v)
            checks["volume_move_verdict"] = s.overlap_scaled(1.0) == ovl_v
            s.sweep_nvt(2, dr_max)
            checks["volume_move_continue"] = np.array_equal(s.download(), multi3)
            checks["rdf_sharded"] = np.array_equal(rdf_part.numpy(), rdf_full)
        print("nccl calls per rank:", info["nccl_calls"], "acceptance:", cnt[1] / cnt[0], "q6:", q6, flush=True)
        for k, v in checks.items():
            print(f"{k}: {'ok' if v else 'MISMATCH'}", flush=True)
            ok &= bool(v)
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", f"world={world} N={N}",
              "halo=" + ("p2p" if os.environ.get("HSMC_CHECK_P2P", "1") == "1" else "nccl"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
