/* include/hsmc_gpu.h -- C ABI of the B200-native hard-sphere Monte Carlo hot path.
 *
 * The reference (fedluc/HSMC) has no plugin/FFI interface: its "boundary" is the set
 * of internal C seams its host driver calls.  Every entry
 * point below names the reference routine it replaces (paths relative to the
 * reference's src/).  Plain pointers and sizes only; all functions return 0 on
 * success and nonzero on failure, in which case hsmc_gpu_last_error() describes
 * the problem -- the host driver prints "ERROR: ..." and exit(EXIT_FAILURE)s, which
 * preserves the reference's error convention (cell_list.c:127-128, 226-231).
 *
 * The handle is not thread-safe (neither is the reference: file-scope globals).
 * There is NO CPU fallback: without a CUDA device hsmc_gpu_create() fails.
 *
 * Particle tables cross the boundary in the reference's own host layout,
 * `double (*)[4]` = {id, x, y, z} (sim_info.h:20), so that write_config /
 * write_restart / compute_op keep working on the host mirror untouched.
 */
#ifndef HSMC_GPU_H
#define HSMC_GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSMC_GPU_ABI_VERSION 1
#define HSMC_GPU_NCCL_ID_BYTES 128

typedef struct hsmc_gpu hsmc_gpu;

typedef struct hsmc_gpu_config {
  int device;           /* CUDA device ordinal of this rank */
  int rank;             /* slab index along x, 0 <= rank < world */
  int world;            /* number of slabs/GPUs; 1 = whole box on one GPU */
  const void *nccl_id;  /* HSMC_GPU_NCCL_ID_BYTES from hsmc_gpu_nccl_id(), same on all ranks; NULL if world == 1 */
  uint64_t seed;        /* device Philox key (the host MT19937 `seed` keyword is reused) */
  double cell_min;      /* minimum cell edge, >= 1.0 (the reference's neigh_list value); 0 = 1.0 */
  int regrid_interval;  /* sweeps between random grid shifts + cell-list rebuilds; 0 = 1 */
  int sweep_impl;       /* 0 = default kernel choice; nonzero values select a variant (bench/ablation) */
} hsmc_gpu_config;

typedef struct hsmc_gpu_info {
  int abi_version;
  int rank, world;
  int64_t n_total;          /* particles in the whole system */
  int64_t n_owned;          /* particles owned by this rank */
  int64_t n_local;          /* owned + ghost-layer particles resident on this rank */
  int cells[3];             /* global cell grid (even per axis, >= 4) */
  int own_x0, own_x1;       /* global x-layers owned by this rank: [x0, x1) */
  double cell_size[3];
  double box[3];
  uint64_t sweeps_done;     /* Philox sweep counter */
  uint64_t kernel_launches; /* kernels launched by this handle so far */
  uint64_t nccl_calls;      /* NCCL operations issued so far */
} hsmc_gpu_info;

const char *hsmc_gpu_last_error(void);

/* Number of CUDA devices visible (0 or negative => nothing can run). */
int hsmc_gpu_device_count(void);

/* ncclGetUniqueId wrapper: call on rank 0, broadcast the bytes with the host-side
   plumbing (torch.distributed / MPI / files), pass to every rank's create(). */
int hsmc_gpu_nccl_id(void *out_id /* HSMC_GPU_NCCL_ID_BYTES */);

/* Replaces cell_list_init(true) (cell_list.c:33-58) + device allocation.
   box = {lx, ly, lz} (sim_info.c:61-63).  Called where hs_nvt/hs_npt call
   cell_list_init (nvt.c:55, npt.c:50). */
int hsmc_gpu_create(hsmc_gpu **out, const hsmc_gpu_config *cfg, int64_t n_particles,
                    const double box[3]);

/* Optional NVLink peer-to-peer halo path (world > 1, one process per GPU on one node).
   Every rank exports an opaque blob (HSMC_GPU_IPC_BYTES) describing its receive window, the
   launcher gathers the blobs, and each rank attaches its left and right neighbours' blobs.
   After that the migration/ghost messages of the cell-list rebuild and the boundary-layer
   refreshes are written by the producing kernels straight into the neighbour's HBM over
   NVLink (exact sizes, no NCCL launch), ordered by sequence flags; NCCL is only used for
   the all-reduces.  Without attach the same exchanges go through ncclSend/ncclRecv. */
#define HSMC_GPU_IPC_BYTES 64
int hsmc_gpu_ipc_export(hsmc_gpu *h, void *out_blob);
int hsmc_gpu_ipc_attach(hsmc_gpu *h, const void *left_blob, const void *right_blob);

/* Replaces cell_list_free (cell_list.c:83-90; nvt.c:99, npt.c:101). */
int hsmc_gpu_destroy(hsmc_gpu *h);

int hsmc_gpu_get_info(hsmc_gpu *h, hsmc_gpu_info *out);

/* Host-only planning (no CUDA call): the cell grid and the x-slab this rank would own for
   a box, minimum cell edge and world size; fills cells, cell_size, box, own_x0/own_x1,
   rank, world.  Lets launchers and CPU tests reason about the decomposition. */
int hsmc_gpu_plan(const double box[3], double cell_min, int world, int rank, hsmc_gpu_info *out);

/* Host-only (no CUDA call): the two-level checkerboard's block partition hsmc_gpu_sweep_nvt would use
   for this box / particle count on rank `rank` of `world` (xpart_world > 1: a single-GPU handle told to
   mimic that many slabs, hsmc_gpu_config.sweep_impl bits 8..15).  The block SHAPE must be the same on
   every rank of a run and on the single-GPU run that mimics it -- it is part of the chain's definition;
   CPU tests check exactly that. */
#define HSMC_GPU_PLAN_MAX_XCUTS 512
typedef struct hsmc_gpu_block_plan {
  int ok;                 /* 0: the block-resident kernel cannot be used on this grid (generic kernel instead) */
  int blocks[3];          /* blocks per axis (even); x counts the blocks of this rank's layers */
  int max_extent[3];      /* largest block extent per axis, in cells */
  int ctas_per_phase;
  int staged_capacity;    /* particles (pad included) one CTA can stage */
  int smem_bytes;         /* dynamic shared memory per CTA */
  int n_xcuts;            /* blocks[0] + 1 */
  int xcuts[HSMC_GPU_PLAN_MAX_XCUTS];  /* global x-layer boundaries of this rank's blocks, ascending */
} hsmc_gpu_block_plan;
int hsmc_gpu_plan_blocks(const double box[3], double cell_min, int world, int rank, int xpart_world,
                         int64_t n_particles, hsmc_gpu_block_plan *out);

/* CUDA stream (cudaStream_t) all work of this handle is enqueued on. */
void *hsmc_gpu_stream(hsmc_gpu *h);

/* Block the host until all enqueued work of the handle is complete. */
int hsmc_gpu_sync(hsmc_gpu *h);

/* part_conf -> device + cell_list_new() (sim_info.c:159-166, cell_list.c:142-175).
   `rows` is n_rows x {id,x,y,z}.  world == 1: n_rows must equal n_particles.
   world > 1: rows may be the full table or any superset of this rank's slab;
   rows outside the slab are ignored, ghost layers come from the neighbours. */
int hsmc_gpu_upload(hsmc_gpu *h, const double *rows, int64_t n_rows);

/* device -> part_conf, id-ordered full table (world == 1 only).  Needed before
   write_config / write_restart / compute_op (io_config.c:28-74,134-191). */
int hsmc_gpu_download(hsmc_gpu *h, double *conf /* n_particles x 4 */);

/* Owned rows of this rank, cell-ordered, {id,x,y,z}; any world size. */
int hsmc_gpu_download_owned(hsmc_gpu *h, double *rows, int64_t capacity_rows, int64_t *n_rows);

/* The same table as hsmc_gpu_download, piecewise, for writers that work while the copy goes on (write_config,
   io_config.c:134-191: formatting + deflate of rows [0, k) overlaps the transfer of rows [k, ...)).
   pack_table orders the table by id on the device once (the device copy of what download would return);
   fetch_rows copies rows [first_row, first_row + n_rows) of it to the host and returns when they have arrived.
   The packed copy is valid until the next call that changes the configuration.  world == 1. */
int hsmc_gpu_pack_table(hsmc_gpu *h);
int hsmc_gpu_fetch_rows(hsmc_gpu *h, int64_t first_row, int64_t n_rows, double *rows);

/* Page-lock (or release: pin = 0) a host buffer the caller owns -- the particle table part_conf of
   sim_info.c:99 -- so that uploads and downloads run at the full host-link rate.  Optional; failure is not fatal. */
int hsmc_gpu_pin_host(void *ptr, size_t bytes, int pin);

/* Replaces n_sweeps x sweep_nvt() (nvt.c:201-209): each sweep = N single-particle
   trial displacements (moves.c:27-80) executed as 8 checkerboard colour phases.
   dr_max is passed every call because the optimizer mutates it (optimizer.c:34-48). */
int hsmc_gpu_sweep_nvt(hsmc_gpu *h, int n_sweeps, double dr_max);

/* The all-particle scaled overlap test inside vol_move() (moves.c:106-112) and
   presst (compute_press.c:250-271): *overlap = 1 iff any pair is closer than 1.0
   after scaling coordinates and box by sf.  Collective over all ranks. */
int hsmc_gpu_overlap_scaled(hsmc_gpu *h, double sf, int *overlap);

/* Accepted volume move (moves.c:129-142): coordinates *= sf, PBC, new box
   (the host recomputes it with sim_box_init), cell list rebuilt.  world > 1: supported while
   the number of cells per axis does not change (slab ownership then scales with the box). */
int hsmc_gpu_rescale(hsmc_gpu *h, double sf, const double new_box[3]);

/* widom_insertion() (compute_widom_chem_pot.c:44-71): insertion points
   [first, first+count) of sample `sample_id`; *accepted = non-overlapping ones.
   Point m is r = u*L with u = philox32/0xffffffff (the reference's u = mt/max).
   world > 1: every rank passes the same range and tests the points that fall in its
   slab; the result is all-reduced.  reduce = 0 skips the all-reduce (replicated
   configurations with the range sharded by the caller). */
int hsmc_gpu_widom(hsmc_gpu *h, uint64_t sample_id, int64_t first, int64_t count, int reduce,
                   int64_t *accepted);

/* rdf_hist_compute() (compute_rdf.c:110-128): pair counts per bin, bin =
   (int)((dr-1.0)/dr_bin) for dr < dr_bin*nn + 1.0.  rdf_hist[k] = 2.0 * counts[k]. */
int hsmc_gpu_rdf_counts(hsmc_gpu *h, double dr_bin, int nn, uint64_t *counts);

/* The same histogram sharded over GPUs: every process holds the WHOLE configuration
   (world == 1 handles, replicated upload) and counts the tile pairs of share `part` of `nparts`;
   the caller sums the nparts results (NCCL / torch.distributed all-reduce).  The sum over all
   parts equals hsmc_gpu_rdf_counts bin for bin. */
int hsmc_gpu_rdf_counts_part(hsmc_gpu *h, double dr_bin, int nn, int part, int nparts, uint64_t *counts);

/* pressv_compute_hist() (compute_press.c:123-165): same binning restricted to
   r < dr_bin*nn + 1.0 <= cell edge, through the cell list.  pressv_hist[k] = 2.0*counts[k]. */
int hsmc_gpu_contact_counts(hsmc_gpu *h, double dr_bin, int nn, uint64_t *counts);

/* global_ql_compute() (compute_order_parameter.c:84-97, per particle :99-229): the average over
   all particles of the Steinhardt bond-order parameter q_l, bonds = neighbours within rmax.
   rmax must not exceed the cell edge (the reference clips it to its neighbour-list cell size,
   compute_order_parameter.c:29-40; the host driver does the same before calling).  l <= 12.
   Floating point (1e-12 relative against the reference); collective over all ranks. */
int hsmc_gpu_order_parameter(hsmc_gpu *h, int l, double rmax, double *ql_ave);

/* presst_compute_hist() (compute_press.c:239-273): for each scale factor sf[k]
   (host computes pow(1-xi_k, 1./3.) as the reference does) no_overlap[k] = 1 iff the
   compressed system has no overlapping pair.  One pass over the pairs for all k. */
int hsmc_gpu_presst_flags(hsmc_gpu *h, const double *sf, int nn, int *no_overlap);

/* get_moves_counters / reset_moves_counters (moves.c:231-251), 64-bit:
   {part_moves, acc_part_moves, rej_part_moves, vol_moves, acc_vol_moves, rej_vol_moves}.
   The host driver owns the three volume counters and adds them with add_vol_move. */
int hsmc_gpu_counters(hsmc_gpu *h, int64_t out[6]);
int hsmc_gpu_reset_counters(hsmc_gpu *h);
int hsmc_gpu_add_vol_move(hsmc_gpu *h, int accepted);
/* trial moves rejected because the displacement left the particle's cell (a rule of
   the checkerboard chain, not of the reference); included in rej_part_moves. */
int hsmc_gpu_cell_rejects(hsmc_gpu *h, int64_t *out);

/* Optional device-side timing of the handle's own kernels (CUDA events recorded on the
   handle's stream around each launch group).  Buckets: 0 = sweep colour phases (K2),
   1 = cell-list rebuild (K1), 2 = halo exchange / NCCL, 3 = everything else.
   profile_read() synchronises, returns accumulated milliseconds + launch groups per
   bucket since the last read, and resets them. */
#define HSMC_GPU_PROFILE_BUCKETS 4
int hsmc_gpu_profile(hsmc_gpu *h, int enable);
int hsmc_gpu_profile_read(hsmc_gpu *h, double ms[HSMC_GPU_PROFILE_BUCKETS],
                          int64_t groups[HSMC_GPU_PROFILE_BUCKETS]);

/* Philox sweep counter (restart support: append to the restart file). */
int hsmc_gpu_set_sweep_counter(hsmc_gpu *h, uint64_t sweeps_done);

/* ---- parity-test entry points (not used by the host driver) ---- */

/* check_overlap(idx, sf, sf, sf) for particle idx[i] placed at xyz[i] with everything
   else fixed, exactly as part_move evaluates a trial (moves.c:52-60).  world == 1. */
int hsmc_gpu_trial_verdicts(hsmc_gpu *h, int n, const int *idx, const double *xyz, double sf,
                            int *flags);

/* widom_check_overlap() for explicit points (compute_widom_chem_pot.c:82-120). */
int hsmc_gpu_widom_verdicts(hsmc_gpu *h, int n, const double *xyz, int *flags);

/* One sweep with every trial recorded: id, the three raw 32-bit draws, the verdict
   (0 accepted, 1 rejected: overlap, 2 rejected: left cell) and a sequence key
   (phase << 56 | global cell << 8 | order within cell).  Used to replay the GPU's
   trial moves through the reference's own part_move(). */
typedef struct hsmc_gpu_trial {
  uint64_t seq;
  int32_t id;
  int32_t verdict;
  uint32_t raw[3];
  uint32_t pad;
} hsmc_gpu_trial;
int hsmc_gpu_sweep_nvt_logged(hsmc_gpu *h, double dr_max, hsmc_gpu_trial *log, int64_t capacity,
                              int64_t *n_logged);

/* Exhaustive self-test of the division-free evaluation of u = raw/0xffffffff used for the
   trial displacements: counts, over all 2^32 raw values, those for which it differs from the
   IEEE double division the reference performs (rng.c:29-31).  Must report 0. */
int hsmc_gpu_selftest_u01(hsmc_gpu *h, uint64_t *n_mismatch, uint32_t *first_bad);

/* min over all stencil pairs of the pair distance squared (invariant checks). */
int hsmc_gpu_min_dist2(hsmc_gpu *h, double *out);

#ifdef __cplusplus
}
#endif
#endif /* HSMC_GPU_H */
