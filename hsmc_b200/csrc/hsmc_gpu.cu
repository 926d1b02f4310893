// hsmc_gpu.cu -- B200 (sm_100a) hard-sphere Monte Carlo hot path behind the C ABI of
// include/hsmc_gpu.h.  Hand-written CUDA, no CPU fallback, no library kernels.
//
// Data layout in HBM (per rank):
//   pos[2][cap]   double4 {x, y, z, id}, cell-ordered (counting sort), ping-pong
//   cell_start    int[ncell_local + 1]   CSR offsets into pos
//   key, rnk      int[cap]               scratch of the counting sort
// One translation unit; the kernels live in the headers included below, this file keeps the handle, the
// host-side orchestration and the ABI entry points ("new kernel" table):
//   cell_list.cuh      K1  k_cell_count / k_scan_* / k_cell_scatter   cell_list_new (cell_list.c:142-175)
//   sweep_generic.cuh  K2  k_sweep_phase (global memory)  part_move + check_overlap (moves.c:27-80,157-212)
//   sweep_lean.cuh     K2  k_propose + k_sweep_lean       all proposals of a sweep up front; block-resident fp32x2 stencil
//                                                         filter, fused block phases (the sweep of small systems)
//   sweep_gather.cuh   K2  k_sweep_gather                 one thread per trial, one launch per (colour, trial index): the
//                                                         sweep of large systems
//   observables.cuh    K3  k_overlap_scaled   vol_move / presst verdict   (moves.c:106-112)
//                      K4  k_widom            widom_insertion             (compute_widom_chem_pot.c:44-71)
//                      K5  k_rdf_pairs        rdf_hist_compute            (compute_rdf.c:110-128)
//                      K6  k_contact_hist     pressv_compute_hist         (compute_press.c:123-165)
//                      K7  k_rescale          accepted volume move        (moves.c:135-142)
//                      K8  k_order_param      global_ql_compute           (compute_order_parameter.c:84-229)
//   slab.cuh           slab decomposition: classification, halo messages, flags (world > 1)
#include <cuda_runtime.h>
#include <nccl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/hsmc_gpu.h"
#include "geom.cuh"
#include "philox.cuh"

// ----------------------------------------------------------------------------------
// error plumbing
// ----------------------------------------------------------------------------------
static thread_local std::string g_err;

static thread_local std::string g_dbg;

static int fail(const std::string& m) {
  g_err = m + g_dbg;
  return 1;
}

#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return fail(std::string("CUDA: ") + cudaGetErrorString(e__) + " at " #call);     \
  } while (0)

#define NC(call)                                                                       \
  do {                                                                                 \
    ncclResult_t r__ = (call);                                                         \
    if (r__ != ncclSuccess)                                                            \
      return fail(std::string("NCCL: ") + ncclGetErrorString(r__) + " at " #call);     \
  } while (0)

#define TRY(call)            \
  do {                       \
    int rc__ = (call);       \
    if (rc__) return rc__;   \
  } while (0)

extern "C" const char* hsmc_gpu_last_error(void) { return g_err.c_str(); }

// ----------------------------------------------------------------------------------
// handle
// ----------------------------------------------------------------------------------
enum { CNT_TRIALS = 0, CNT_ACC = 1, CNT_REJ_OVERLAP = 2, CNT_REJ_CELL = 3, CNT_N = 8 };

// block-resident sweep (sweep_lean.cuh)
struct BlockCfg {
  int nbx, nby, nbz;      // blocks per axis (even); x: over the layers this rank owns
  int mbx, mby, mbz;      // largest block extent per axis (cells)
  int cap;                // staged shadow capacity (float4 entries, pad included)
  int cs_stride;          // (unused)
  int cz_stride;          // ushorts per staged CSR row (multiple of 8)
  int nslots;             // chunks (32 trial slots each) the block's trial table holds (< 256)
  int max_rows;           // (mbx+2)*(mby+2)
  int max_cells;          // mbx*mby*mbz
  int use_tma;
  int force_global;       // ablation: every block takes the global-memory path (same chain)
  int dbg;                // timing ablations (env HSMC_BLOCK_DBG), 0 in production
  unsigned int* ticket;   // fused launches: CTA ticket counter (never reset; SweepArgs::ticket_base)
  unsigned int* done;     // fused launches: [nbx][nby][nbz] epoch of the launch that last finished the block
  unsigned long long* stamps;   // tuning aid (HSMC_BLOCK_STAMPS=1): [0..23] clock cycles summed per stage of a block, [31] blocks
};

// per-row staging record: global slots of the row's one or two pieces, staged offset
struct BlockRow { int gbA, gbB, cntA, off; };

// trial lists of k_sweep_gather (sweep_gather.cuh), filled by k_propose
struct GatherLists {
  int* list;                    // [GATHER_LISTS][stride] local cell indices
  int* count;                   // [GATHER_LISTS]
  long long stride;
};

// Slab runs over NVLink windows: the right ghost layer is delivered INSIDE the sweep launch.  The blocks of a rank's
// first owned layer (block column 0, phases 0-3) store that layer's cells straight into the left neighbour's right
// ghost layer (same cells, same slot order: both are kept id-sorted) and then raise a per-(y,z)-block flag in the
// neighbour's window; the neighbour's last block column (phases 4-7), the only reader of that ghost layer, waits for
// the nine flags around it.  So all eight block phases of a slab are ONE launch, with no halo exchange in between.
struct SlabLink {
  double4* peer_pos[2];                 // left neighbour's master table (both halves), nullptr = link off
  float4* peer_rel;
  unsigned int* peer_done;              // left neighbour's window: [nby][nbz] number of the last sweep that block finished
  const volatile unsigned int* info;    // my window: {right-ghost base slot, rebuild number, current half} of the LEFT neighbour
  const volatile unsigned int* my_done; // my window: the same flags, raised by the right neighbour
  unsigned int seq;                     // this sweep (same count on every rank)
  unsigned int rebuild;                 // the rebuild the layout must come from
};

// cfg.sweep_impl: low byte = kernel variant, next byte = virtual world of the x block partition
// 0 (= 8) default: k_propose + k_sweep_lean (block-resident, one launch per sweep).  7: k_propose + k_sweep_gather (one
// thread per trial, one launch per cell colour and trial index; measured slower, kept as the second implementation
// of its chain).  Reference evaluations of the same two update orders, all in double from global memory: 5 the
// block order (the chain of 0), 1 one thread per cell, one launch per cell colour (the chain of 7).  3: the default
// with the filter's error band forced to zero (negative control of the parity tests).
enum { IMPL_AUTO = 0, IMPL_CELL_GLOBAL = 1, IMPL_EPS0 = 3, IMPL_BLOCK_GLOBAL = 5, IMPL_GATHER = 7, IMPL_LEAN = 8 };

struct hsmc_gpu {
  hsmc_gpu_config cfg;
  int impl = 0;            // cfg.sweep_impl & 0xff
  int xpart_world = 0;     // (cfg.sweep_impl >> 8) & 0xff; 0 = cfg.world
  int64_t N = 0;           // global particle count
  int64_t n_local = 0;     // resident particles (owned + ghosts)
  int64_t n_owned = 0;
  int64_t own_first = 0;   // slot of the first owned particle
  int64_t cap = 0;         // slots in pos[*]
  double box[3];
  Grid g;
  int64_t ncell = 0;       // local cells
  int64_t cap_cells = 0;
  cudaStream_t st = nullptr;
  cudaStream_t st2 = nullptr;            // copy stream (chunked upload / download)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  double4* pos[2] = {nullptr, nullptr};
  float4* rel = nullptr;                 // fp32 shadow {cell-relative offset, id} of pos[cur]
  int cur = 0;
  int *key = nullptr, *rnk = nullptr;    // [cap_keys]
  int64_t cap_keys = 0;
  int *key_halo = nullptr, *rnk_halo = nullptr;   // [2*cap_halo]
  int *cell_count = nullptr, *cell_start = nullptr, *bsum = nullptr;
  unsigned long long* d_cnt = nullptr;       // CNT_N counters
  unsigned long long* d_scratch = nullptr;   // SCRATCH_N x u64 general scratch (flags, hist, min)
  int* d_slot_of_id = nullptr;           // parity entry points only
  void* h_stage = nullptr;               // pinned staging for small results
  double* d_io = nullptr;                // staging for upload/download
  int64_t cap_io = 0;
  bool io_packed = false;                // d_io holds the current table by id (hsmc_gpu_pack_table)
  uint64_t sweeps_done = 0;
  uint64_t launches = 0, nccl_calls = 0;
  int64_t vol_cnt[3] = {0, 0, 0};
  bool have_conf = false;
  int since_regrid = 0;
  ncclComm_t comm = nullptr;
  // slab mode
  double4 *send_l = nullptr, *send_r = nullptr, *recv_l = nullptr, *recv_r = nullptr;
  int64_t cap_halo = 0;                  // slots per halo buffer, slot 0 = header
  int* d_halo_cnt = nullptr;             // [0]=send_l count [1]=send_r count [2]=error flags
  int lay[6] = {0, 0, 0, 0, 0, 0};       // slot offsets of layers 0,1,2,nlx-2,nlx-1,nlx (host mirror)
  int* d_lay = nullptr;                  // the same + error flags on the device (slab rebuilds never sync)
  bool lay_valid = true;                 // host mirror current?
  bool ghost1_stale = false;
  // NVLink peer-to-peer windows (optional, see hsmc_gpu_ipc_attach)
  unsigned char* win = nullptr;          // my receive window: [flags | A_l x2 | A_r x2 | B | C]
  unsigned char *win_left = nullptr, *win_right = nullptr;   // neighbours' windows, mapped
  bool p2p = false;
  bool pos_in_win = false;               // pos[*] / rel live inside the window allocation (slab runs)
  size_t peer_off[4] = {0, 0, 0, 0};     // left neighbour's window: offsets of its block flags, pos[0], pos[1], rel
  uint32_t seqG = 0, seqR = 0;           // sweeps with the in-kernel ghost delivery / slab rebuilds so far (same on every rank)
  uint32_t seqA = 0, seqB = 0, seqC = 0; // exchanges issued so far (same on every rank)             // left ghost layer not refreshed since the last odd-x phases
  void* d_sfargs = nullptr;
  BlockCfg blk;
  size_t blk_smem = 0;                   // dynamic shared memory of k_sweep_lean
  bool blk_ok = false;                   // the block-resident kernel runs the sweeps
  uint4* trec = nullptr;                 // [cap] trial records of the current sweep, trial order inside each cell (k_propose)
  uint4* traw = nullptr;                 // [cap] logged sweeps only: {raw draws, id}
  bool gather = false;                   // the sweeps run k_sweep_gather (else k_sweep_lean if blk_ok, else k_sweep_phase)
  GatherLists glists = {nullptr, nullptr, 0};
  float blk_eps = 0.f;
  std::vector<int> xoff;                 // x block boundaries (local layers), blk.nbx + 1 entries
  int* d_xoff = nullptr;
  unsigned int* d_fuse = nullptr;      // fused block phases: [0] ticket counter, [64..] completion flags per block
  int64_t cap_fuse = 0;
  unsigned int fuse_epoch = 0, fuse_tickets = 0;
  int fuse_mode = -1;                  // -1: not decided yet; 0 off (HSMC_FUSE=0); 1 on
  int64_t cap_xoff = 0;
  bool xoff_dirty = true;
  unsigned long long* d_stamps = nullptr;   // HSMC_BLOCK_STAMPS
  hsmc_gpu_trial* d_log = nullptr;
  int64_t cap_log = 0;
  // optional event timing
  bool prof = false;
  std::vector<cudaEvent_t> ev_pool;
  struct Span { cudaEvent_t a, b; int bucket; };
  std::vector<Span> spans;
  double prof_ms[HSMC_GPU_PROFILE_BUCKETS] = {0, 0, 0, 0};
  int64_t prof_n[HSMC_GPU_PROFILE_BUCKETS] = {0, 0, 0, 0};
};

static cudaEvent_t prof_event(hsmc_gpu* h) {
  cudaEvent_t e;
  if (!h->ev_pool.empty()) { e = h->ev_pool.back(); h->ev_pool.pop_back(); }
  else cudaEventCreate(&e);
  return e;
}
struct ProfSpan {
  hsmc_gpu* h; cudaEvent_t a; int bucket; bool on;
  ProfSpan(hsmc_gpu* h_, int bucket_) : h(h_), bucket(bucket_), on(h_->prof) {
    if (on) { a = prof_event(h); cudaEventRecord(a, h->st); }
  }
  ~ProfSpan() {
    if (on) { cudaEvent_t b = prof_event(h); cudaEventRecord(b, h->st); h->spans.push_back({a, b, bucket}); }
  }
};

#define SCRATCH_N 16384

static inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

#include "cell_list.cuh"
#include "async_copy.cuh"
#include "sweep_generic.cuh"
#include "sweep_lean.cuh"
#include "sweep_gather.cuh"

#include "observables.cuh"
#include "slab.cuh"
// ----------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------
static int even_cells(double L, double cell_min) {
  int n = (int)floor(L / cell_min);
  n &= ~1;
  // keep a rounding margin between the cell edge and the requested minimum
  while (n >= 2 && L / n < cell_min * (1.0 + 1e-9)) n -= 2;
  return n;
}

static void setup_blocks(hsmc_gpu* h);
static int sync_layout(hsmc_gpu* h);

static int setup_grid(hsmc_gpu* h) {
  double cm = h->cfg.cell_min;
  Grid& g = h->g;
  g.nx = even_cells(h->box[0], cm);
  g.ny = even_cells(h->box[1], cm);
  g.nz = even_cells(h->box[2], cm);
  if (g.nx < 4 || g.ny < 4 || g.nz < 4)
    return fail("simulation box too small for the checkerboard decomposition: every edge must hold at least 4 cells of size >= cell_min");
  g.Lx = h->box[0]; g.Ly = h->box[1]; g.Lz = h->box[2];
  g.wx = g.Lx / g.nx; g.wy = g.Ly / g.ny; g.wz = g.Lz / g.nz;
  g.iwx = 1.0 / g.wx; g.iwy = 1.0 / g.wy; g.iwz = 1.0 / g.wz;
  g.sx = g.sy = g.sz = 0.0;
  int W = h->cfg.world;
  if (W == 1) {
    g.gx0 = 0; g.nlx = g.nx; g.own_lo = 0; g.own_hi = g.nx; g.wrap_x = 1;
  } else {
    int pairs = g.nx / 2;
    if (pairs < 2 * W) return fail("too few cell layers along x for this many slabs (need >= 4 layers per rank)");
    int x0 = 2 * (int)(((long long)pairs * h->cfg.rank) / W);
    int x1 = 2 * (int)(((long long)pairs * (h->cfg.rank + 1)) / W);
    g.gx0 = (x0 - 1 + g.nx) % g.nx;
    g.nlx = (x1 - x0) + 2;
    g.own_lo = 1; g.own_hi = 1 + (x1 - x0);
    g.wrap_x = 0;
  }
  h->ncell = (int64_t)g.nlx * g.ny * g.nz;
  if (h->ncell + 1 > (int64_t)INT32_MAX) return fail("too many cells for 32-bit cell indices");
  setup_blocks(h);
  return 0;
}

static int ensure_cell_arrays(hsmc_gpu* h) {
  if (h->ncell + 1 <= h->cap_cells) return 0;
  if (h->cell_start) cudaFree(h->cell_start);
  if (h->cell_count) cudaFree(h->cell_count);
  if (h->bsum) cudaFree(h->bsum);
  h->cap_cells = h->ncell + 1 + h->ncell / 8;
  CU(cudaMalloc(&h->cell_start, sizeof(int) * (size_t)h->cap_cells));
  CU(cudaMalloc(&h->cell_count, sizeof(int) * (size_t)h->cap_cells));
  CU(cudaMalloc(&h->bsum, sizeof(int) * (size_t)((h->cap_cells + SCAN_CHUNK - 1) / SCAN_CHUNK + 1)));
  return 0;
}

static int ensure_keys(hsmc_gpu* h, int64_t n) {
  if (n <= h->cap_keys) return 0;
  if (h->key) cudaFree(h->key);
  if (h->rnk) cudaFree(h->rnk);
  h->cap_keys = n;
  CU(cudaMalloc(&h->key, sizeof(int) * (size_t)n));
  CU(cudaMalloc(&h->rnk, sizeof(int) * (size_t)n));
  return 0;
}

static int ensure_io(hsmc_gpu* h, int64_t rows) {
  if (rows <= h->cap_io) return 0;
  if (h->d_io) cudaFree(h->d_io);
  h->cap_io = rows;
  CU(cudaMalloc(&h->d_io, sizeof(double) * 4 * (size_t)rows));
  return 0;
}

// grid shift of the current sweep counter: same on every rank, no communication
static void draw_shift(hsmc_gpu* h) {
  Philox4 r = philox4x32_10(0u, HSMC_STREAM_SHIFT << 24, (uint32_t)h->sweeps_done,
                            (uint32_t)(h->sweeps_done >> 32), (uint32_t)h->cfg.seed,
                            (uint32_t)(h->cfg.seed >> 32));
  Grid& g = h->g;
  g.sx = ((double)r.v[0] / 4294967296.0) * g.wx;
  g.sy = ((double)r.v[1] / 4294967296.0) * g.wy;
  g.sz = ((double)r.v[2] / 4294967296.0) * g.wz;
}

static int exclusive_scan(hsmc_gpu* h, const int* in, int64_t n, int* out) {
  int nb = (int)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
  k_scan_blocksum<<<nb, SCAN_T, 0, h->st>>>(in, n, h->bsum);
  k_scan_top<<<1, SCAN_T, 0, h->st>>>(h->bsum, nb);
  k_scan_final<<<nb, SCAN_T, 0, h->st>>>(in, n, h->bsum, out);
  h->launches += 3;
  CU(cudaGetLastError());
  return 0;
}


// ---- block-resident sweep: block shape, x partition, error band ------------------------
// Even number of blocks per axis, extents differing by at most one cell.  Along x the
// partition is made slab by slab (every slab gets an even number of blocks) so that a
// single-GPU run told to use the partition of a W-slab run (sweep_impl bits 8..15) is the
// same Markov chain as the W-GPU run.
static int even_blocks(int n_cells, int b) { return 2 * std::max(1, (n_cells + 2 * b - 1) / (2 * b)); }

static void setup_blocks(hsmc_gpu* h) {
  Grid& g = h->g;
  BlockCfg& b = h->blk;
  h->blk_ok = false;
  h->xoff_dirty = true;
  const int W = h->cfg.world;
  const int Wv = (W > 1) ? W : std::max(1, h->xpart_world);
  const double nbar = (double)h->N / ((double)g.nx * g.ny * g.nz);
  if ((long long)g.nz * 16 > 65535) return;        // 16-bit row-relative CSR
  // x slabs of the (virtual) world: [lo, hi) in global layers
  // `all_slabs`: every slab of the (virtual) world -- the block SHAPE is chosen from these, so that
  // it is a function of the grid and the world size only: every rank of a slab run and the single-GPU
  // run that mimics it (xpart_world) pick the same shape, hence the same chain.  `slabs`: the slabs
  // whose blocks this handle runs (its own one in slab mode, all of them on a single GPU).
  std::vector<std::pair<int, int>> slabs, all_slabs;
  {
    const int pairs = g.nx / 2;
    if (Wv > 1 && pairs < 2 * Wv) return;            // cannot honour the requested partition
    for (int r = 0; r < Wv; r++) {
      int x0 = 2 * (int)(((long long)pairs * r) / Wv), x1 = 2 * (int)(((long long)pairs * (r + 1)) / Wv);
      all_slabs.push_back({x0, x1});
      if (W > 1 && r != h->cfg.rank) continue;
      slabs.push_back({x0, x1});
    }
  }
  double capf = 1.12;        // (a block of a crystal can hold 10 % more than the mean: lattice planes beat against the cells)
  if (const char* e = getenv("HSMC_BLOCK_CAPF")) capf = atof(e);
  int want[3] = {0, 0, 0};
  if (const char* e = getenv("HSMC_BLOCK")) sscanf(e, "%d,%d,%d", &want[0], &want[1], &want[2]);
  struct Shape { int bx, by, bz, mx, my, mz, cap, nslots; size_t smem; };
  auto eval = [&](int bx, int by, int bz, Shape& s) -> bool {
    // a block and its halo must not cover a cell twice: extent + 2 <= cells of the axis
    int mx = 0;
    for (auto& sl : all_slabs) {
      int n = sl.second - sl.first, nb = even_blocks(n, bx);
      if (nb > n) return false;
      mx = std::max(mx, (n + nb - 1) / nb);
    }
    int nby = even_blocks(g.ny, by), nbz = even_blocks(g.nz, bz);
    if (nby > g.ny || nbz > g.nz) return false;
    int my = (g.ny + nby - 1) / nby, mz = (g.nz + nbz - 1) / nbz;
    // (x: with Wv > 1 slabs every block is at most its own slab long, and a slab + 2 never exceeds
    //  the local layer count of a rank nor the box, so only the one-slab case needs the test)
    if ((Wv == 1 && mx + 2 > g.nx) || my + 2 > g.ny || mz + 2 > g.nz) return false;
    // k_sweep_lean: 4-bit row coordinates, one lane per cell of a staged row (mz + 3 entries), 12-bit staged index
    if ((mx + 2) * (my + 2) > LEAN_MAX_ROWS || mz + 3 > 32 || mx + 2 > 16 || my + 2 > 16) return false;
    double region = (double)(mx + 2) * (my + 2) * (mz + 2);
    int cap = ((int)(region * nbar * capf) + 48 + LEAN_PAD + 31) & ~31;
    if (cap > 4096) return false;
    int cz_stride = (mz + 3 + 7) & ~7;
    const size_t rows = (size_t)(mx + 2) * (my + 2);
    // trial table: chunks of at most 32 cells and 32 trials, handed out one by one; a chunk is cut at a cell boundary
    // and balanced over the warps, so it holds ~27 trials on average; head-room for dense blocks
    const double interior = (double)mx * my * mz;
    const int nslots = std::min(250, (int)std::ceil(std::max(interior / 30.0, interior * std::max(nbar, 0.2) * 1.25 / 26.0)) + 16);
    size_t smem = (size_t)cap * 12 + rows * cz_stride * 2 + (size_t)nslots * (64 + 4 + 1) + (size_t)cap / 8 + 8 * LEAN_COL_CHUNKS;
    smem = (smem + 15) & ~(size_t)15;
    if (smem > 100 * 1024) return false;
    s = {bx, by, bz, mx, my, mz, cap, nslots, smem};
    return true;
  };
  Shape best{};
  bool have = false;
  if (want[0] > 0 && want[1] > 0 && want[2] > 0) have = eval(want[0], want[1], want[2], best);
  if (!have) {
    // Measured on B200 (profiles/): the cost per trial is flat once a CTA holds >= ~1000 trials and
    // five CTAs fit an SM; smaller blocks pay the fixed prologue more often, larger ones lose
    // residency (8x8x23 cells: four CTAs per SM, 12 % slower per trial than 8x8x21 with five).
    // Score = (fraction of the 5 x 148 CTA slots a phase can fill) x (prologue amortisation) x
    // (residency), ties to the larger block.
    double best_score = -1.0;
    const int cand_xy[] = {2, 3, 4, 5, 6, 8}, cand_z[] = {2, 4, 6, 8, 10, 12, 16, 18, 20, 22, 24, 28};
    for (int bx : cand_xy) for (int by : cand_xy) for (int bz : cand_z) {
      Shape s;
      if (!eval(bx, by, bz, s)) continue;
      // CTAs of one phase on one GPU: all slabs on a single GPU, the largest slab in a slab run
      // (the same number for a real and for a mimicked slab run)
      long long ctas = 0;
      for (auto& sl : all_slabs) {
        const long long c = even_blocks(sl.second - sl.first, bx) / 2;
        ctas = (Wv > 1) ? std::max(ctas, c) : ctas + c;
      }
      ctas *= (long long)(even_blocks(g.ny, by) / 2) * (even_blocks(g.nz, bz) / 2);
      // CTAs of k_sweep_lean per SM: dynamic + static shared memory + the 1 KB the hardware reserves per CTA; the
      // register file holds LEAN_MIN_CTAS CTAs of LEAN_THREADS threads at 64 registers
      const int per_sm = (int)std::min<size_t>(LEAN_MIN_CTAS, (size_t)(227 * 1024) / (s.smem + sizeof(BlockRow) * LEAN_MAX_ROWS + 4 * LEAN_MAX_ROWS + 1200));
      if (per_sm < 1) continue;
      const double interior = (double)s.mx * s.my * s.mz;
      // below two waves of CTA slots the ragged last wave costs a whole CTA latency
      const double waves = (double)ctas / (148.0 * per_sm);   // (register file: LEAN_MIN_CTAS CTAs of LEAN_THREADS threads)
      const double fill = waves < 2.0 ? waves / std::ceil(waves) : 1.0;
      const double score = fill * (interior / (interior + 400.0)) * (per_sm / (double)LEAN_MIN_CTAS) + 1e-9 * interior;
      if (score > best_score) { best_score = score; best = s; have = true; }
    }
  }
  if (!have) return;
  b.mbx = best.mx; b.mby = best.my; b.mbz = best.mz; b.cap = best.cap;
  b.nby = even_blocks(g.ny, best.by); b.nbz = even_blocks(g.nz, best.bz);
  b.cs_stride = 0; b.cz_stride = (best.mz + 3 + 7) & ~7;
  b.max_rows = (best.mx + 2) * (best.my + 2);
  b.max_cells = best.mx * best.my * best.mz;
  b.nslots = best.nslots;
  b.use_tma = 0;
  b.force_global = (h->impl == IMPL_BLOCK_GLOBAL) ? 1 : 0;
  b.dbg = getenv("HSMC_BLOCK_DBG") ? atoi(getenv("HSMC_BLOCK_DBG")) : 0;
  b.stamps = nullptr;
  if (getenv("HSMC_BLOCK_STAMPS") && atoi(getenv("HSMC_BLOCK_STAMPS"))) {
    if (!h->d_stamps && cudaMalloc(&h->d_stamps, sizeof(unsigned long long) * 32) == cudaSuccess)
      cudaMemset(h->d_stamps, 0, sizeof(unsigned long long) * 32);
    b.stamps = h->d_stamps;
  }
  h->xoff.clear();
  for (auto& sl : slabs) {
    int n = sl.second - sl.first, nb = even_blocks(n, best.bx);
    int base = (W > 1) ? g.own_lo : sl.first;      // local layer of the slab's first owned layer
    for (int j = 0; j < nb; j++) h->xoff.push_back(base + (int)(((long long)j * n) / nb));
  }
  h->xoff.push_back((W > 1) ? g.own_hi : g.nx);
  b.nbx = (int)h->xoff.size() - 1;
  h->blk_smem = best.smem;
  if (const char* e = getenv("HSMC_BLOCK_PADSMEM")) h->blk_smem += (size_t)atoi(e);    // occupancy experiments
  // fp32 filter error bound: staged coordinates are block-relative,
  // |X| <= (m/2 + 2) cells; per pair and axis: two final roundings at that magnitude, the
  // rounding of the cell edge times the <= 2 cells between a stencil pair, two offset
  // roundings; r2 error <= 2 |d| sum(delta) with |d| <= ~1, plus the fp32 sum itself
  auto half_ulp = [](double m) { int e; frexp(m, &e); return ldexp(1.0, e - 25); };
  double mag[3] = {(0.5 * (best.mx + 2) + 1.0) * g.wx, (0.5 * (best.my + 2) + 1.0) * g.wy, (0.5 * (best.mz + 2) + 1.0) * g.wz};
  double wv[3] = {g.wx, g.wy, g.wz};
  double sum = 0.0;
  for (int k = 0; k < 3; k++) sum += 2.0 * half_ulp(mag[k]) + 2.0 * wv[k] * ldexp(1.0, -24) + 2.0 * wv[k] * ldexp(1.0, -25);
  double r2err = 2.0 * 1.01 * sum + 8.0 * ldexp(1.0, -24);
  h->blk_eps = (float)(2.0 * r2err);
  h->blk_ok = (h->impl == IMPL_AUTO || h->impl == IMPL_LEAN || h->impl == IMPL_EPS0 || h->impl == IMPL_BLOCK_GLOBAL);
  if (getenv("HSMC_DEBUG_TILES"))
    fprintf(stderr, "[hsmc_gpu] rank %d: blocks %dx%dx%d of up to %dx%dx%d cells, %d CTAs/phase, cap %d, smem %zu B, eps %.3g\n",
            h->cfg.rank, b.nbx, b.nby, b.nbz, b.mbx, b.mby, b.mbz, (b.nbx / 2) * (b.nby / 2) * (b.nbz / 2), b.cap,
            h->blk_smem, (double)h->blk_eps);
}

// ---- NVLink peer-to-peer receive window layout (identical on every rank) ----
static inline size_t win_off_A(const hsmc_gpu* h, int from_right, int parity) {
  return 256 + ((size_t)(from_right * 2 + parity) * (size_t)h->cap_halo) * sizeof(double4);
}
static inline size_t win_off_B(const hsmc_gpu* h) { return 256 + (size_t)4 * h->cap_halo * sizeof(double4); }
static inline size_t win_off_C(const hsmc_gpu* h) { return win_off_B(h) + (size_t)(h->cap_halo / 2) * sizeof(double4); }
// ... followed by the block flags of the in-kernel ghost delivery and, so that a neighbour can store into them, the
// master table and its shadow themselves.  Header: uint32 0-3 message flags, 8-10 SlabLink::info; byte 64: four
// uint64 offsets {block flags, pos[0], pos[1], rel} (cap differs from rank to rank, so a neighbour reads them here)
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline size_t win_off_gdone(const hsmc_gpu* h) { return align256(win_off_C(h) + (size_t)(h->cap_halo / 2) * sizeof(double4)); }
static inline size_t win_gdone_n(const hsmc_gpu* h) { return (size_t)(h->g.ny / 2) * (size_t)(h->g.nz / 2) + 64; }
static inline size_t win_off_pos(const hsmc_gpu* h, int k) {
  return align256(win_off_gdone(h) + win_gdone_n(h) * sizeof(unsigned int)) + (size_t)k * align256(sizeof(double4) * (size_t)h->cap);
}
static inline size_t win_off_rel(const hsmc_gpu* h) { return win_off_pos(h, 2); }
static inline size_t win_bytes(const hsmc_gpu* h) { return win_off_rel(h) + align256(sizeof(float4) * (size_t)h->cap); }
enum { WIN_INFO = 8, WIN_OFFSETS_BYTE = 64 };
static inline volatile uint32_t* win_flag(unsigned char* w, int k) { return reinterpret_cast<volatile uint32_t*>(w) + k; }
enum { FLAG_A_FROM_LEFT = 0, FLAG_A_FROM_RIGHT = 1, FLAG_B = 2, FLAG_C = 3 };

static inline int left_of(const hsmc_gpu* h) { return (h->cfg.rank + h->cfg.world - 1) % h->cfg.world; }
static inline int right_of(const hsmc_gpu* h) { return (h->cfg.rank + 1) % h->cfg.world; }

// Rebuild the cell-ordered table from `src` (n_in particles; rows_layout: {id,x,y,z}
// instead of {x,y,z,id}) under the current grid.  Output goes to pos[cur^1]; cur flips.
static int rebuild(hsmc_gpu* h, const double4* src, int64_t n_in, int rows_layout, int upload_mode, bool dev_range = false) {
  ProfSpan span(h, 1);
  TRY(ensure_cell_arrays(h));
  TRY(ensure_keys(h, std::max<int64_t>(n_in, h->cap)));
  if (dev_range) n_in = 0;   // source range read from d_lay on the device
  Grid& g = h->g;
  const int T = 256;
  double4* dst = h->pos[h->cur ^ 1];
  CU(cudaMemsetAsync(h->cell_count, 0, sizeof(int) * (size_t)h->ncell, h->st));
  if (h->cfg.world == 1) {
    if (n_in != h->N) return fail("internal: particle count mismatch in rebuild");
    if (rows_layout) {
      // host rows {id,x,y,z} -> slots {x,y,z,id} in the (free) current buffer first
      k_unpack_rows<<<nblk(n_in, T), T, 0, h->st>>>(reinterpret_cast<const double*>(src), (int)n_in, h->pos[h->cur]);
      h->launches++;
      src = h->pos[h->cur];
    }
    k_cell_count<<<nblk(n_in, T), T, 0, h->st>>>(g, src, (int)n_in, h->key, h->rnk, h->cell_count);
    h->launches++;
    TRY(exclusive_scan(h, h->cell_count, h->ncell, h->cell_start));
    k_cell_scatter<<<nblk(n_in, T), T, 0, h->st>>>(g, src, (int)n_in, h->key, h->rnk, h->cell_start, dst, h->rel);
    h->launches++;
    CU(cudaGetLastError());
    h->cur ^= 1;
    h->n_local = h->n_owned = h->N;
    h->own_first = 0;
    return 0;
  }
  // ---- slab mode: fully asynchronous (no host round trip); sync_layout() fetches the layer
  //      offsets and the error flags when the host needs them ----
  const int* d_range = dev_range ? h->d_lay : nullptr;
  const int64_t n_src = dev_range ? h->cap : n_in;
  CU(cudaMemsetAsync(h->d_halo_cnt, 0, sizeof(int) * 4, h->st));
  double4 *out_l = h->send_l, *out_r = h->send_r, *in_l = h->recv_l, *in_r = h->recv_r;
  if (h->p2p) {
    // messages are written straight into the neighbours' windows (double-buffered by parity)
    h->seqA++;
    const int par = h->seqA & 1;
    out_l = reinterpret_cast<double4*>(h->win_left + win_off_A(h, 1, par));    // I am my left neighbour's right
    out_r = reinterpret_cast<double4*>(h->win_right + win_off_A(h, 0, par));
    in_l = reinterpret_cast<double4*>(h->win + win_off_A(h, 0, par));
    in_r = reinterpret_cast<double4*>(h->win + win_off_A(h, 1, par));
  }
  if (n_src > 0) {
    k_slab_classify<<<nblk(n_src, T), T, 0, h->st>>>(g, src, (int)n_in, rows_layout, upload_mode, h->key,
                                                     h->rnk, h->cell_count, out_l, out_r,
                                                     h->d_halo_cnt, (int)h->cap_halo, d_range);
    h->launches++;
  }
  k_halo_headers<<<1, 1, 0, h->st>>>(out_l, out_r, h->d_halo_cnt);
  h->launches++;
  if (h->p2p) {
    k_flag_post<<<1, 1, 0, h->st>>>(win_flag(h->win_left, FLAG_A_FROM_RIGHT), win_flag(h->win_right, FLAG_A_FROM_LEFT), h->seqA);
    k_flag_wait<<<1, 1, 0, h->st>>>(win_flag(h->win, FLAG_A_FROM_LEFT), win_flag(h->win, FLAG_A_FROM_RIGHT), h->seqA);
    h->launches += 2;
  } else {
    size_t cnt = (size_t)h->cap_halo * 4;
    NC(ncclGroupStart());
    NC(ncclSend(h->send_l, cnt, ncclDouble, left_of(h), h->comm, h->st));
    NC(ncclRecv(h->recv_r, cnt, ncclDouble, right_of(h), h->comm, h->st));
    NC(ncclSend(h->send_r, cnt, ncclDouble, right_of(h), h->comm, h->st));
    NC(ncclRecv(h->recv_l, cnt, ncclDouble, left_of(h), h->comm, h->st));
    NC(ncclGroupEnd());
    h->nccl_calls += 4;
  }
  dim3 gb2(148 * 2, 2);
  k_recv_count<<<gb2, T, 0, h->st>>>(g, in_l, in_r, (int)h->cap_halo, h->key_halo, h->rnk_halo,
                                     h->cell_count, h->d_halo_cnt);
  h->launches++;
  TRY(exclusive_scan(h, h->cell_count, h->ncell, h->cell_start));
  if (n_src > 0) {
    k_scatter_layout<<<nblk(n_src, T), T, 0, h->st>>>(g, src, (int)n_in, rows_layout, h->key, h->rnk,
                                                      h->cell_start, dst, h->rel, d_range, (int)h->cap,
                                                      h->d_halo_cnt + 2);
    h->launches++;
  }
  k_recv_scatter<<<gb2, T, 0, h->st>>>(g, in_l, in_r, (int)h->cap_halo, h->key_halo, h->rnk_halo,
                                       h->cell_start, dst, h->rel, (int)h->cap, h->d_halo_cnt + 2);
  h->launches++;
  long long per = (long long)g.ny * g.nz;
  k_sort_cells_by_id<<<dim3(nblk(2 * per, 128), 2), 128, 0, h->st>>>(g, dst, h->rel, h->cell_start, 0, 1);
  h->launches++;
  // (+ for the in-kernel ghost delivery: where my right ghost layer starts, in which half, after which rebuild -- into
  //  the window of the rank that writes it, my right neighbour)
  h->seqR++;
  k_gather_layout<<<1, 32, 0, h->st>>>(g, h->cell_start, h->d_halo_cnt, h->d_lay,
                                       h->p2p ? win_flag(h->win_right, WIN_INFO) : nullptr, h->seqR, h->cur ^ 1);
  h->launches++;
  CU(cudaGetLastError());
  h->cur ^= 1;
  h->lay_valid = false;
  if (getenv("HSMC_DEBUG_SYNC") && sync_layout(h)) return fail(std::string(g_err) + " (after rebuild)");
  return 0;
}

// slab mode: bring the layer offsets and the accumulated error flags to the host
static int sync_layout(hsmc_gpu* h) {
  if (h->cfg.world == 1 || h->lay_valid) return 0;
  int* hs = (int*)h->h_stage;
  CU(cudaMemcpyAsync(hs, h->d_lay, sizeof(int) * 9, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  for (int k = 0; k < 6; k++) h->lay[k] = hs[k];
  h->n_local = h->lay[5];
  h->own_first = h->lay[1];
  h->n_owned = h->lay[4] - h->lay[1];
  h->lay_valid = true;
  if (hs[8]) {
    char dbg[256];
    snprintf(dbg, sizeof(dbg), " [layers %d %d %d %d %d %d, sent %d %d, flags %d, cap %lld, halo cap %lld]", hs[0], hs[1],
             hs[2], hs[3], hs[4], hs[5], hs[6], hs[7], hs[8], (long long)h->cap, (long long)h->cap_halo);
    g_dbg = dbg;
  } else g_dbg.clear();
  if (hs[8] & 1) return fail("slab decomposition: a particle moved more than one cell layer between regrids");
  if (hs[8] & 2) return fail("slab decomposition: halo buffer overflow");
  if (hs[8] & 4) return fail("slab decomposition: received a particle outside the local layers");
  if (hs[8] & 16) return fail("slab decomposition: boundary-layer message does not match the ghost layer");
  if (hs[8] & 64) return fail("cell list: more than 65535 particles in one row of cells");
  if ((hs[8] & 32) || hs[5] > h->cap) return fail("slab decomposition: local particle capacity exceeded");
  return 0;
}

extern "C" int hsmc_gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" int hsmc_gpu_nccl_id(void* out_id) {
  ncclUniqueId id;
  NC(ncclGetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == HSMC_GPU_NCCL_ID_BYTES, "nccl id size");
  memcpy(out_id, &id, sizeof(id));
  return 0;
}

extern "C" int hsmc_gpu_ipc_export(hsmc_gpu* h, void* out_blob) {
  if (!h || !out_blob) return fail("null argument");
  if (h->cfg.world < 2) return fail("ipc_export: only meaningful with world > 1");
  CU(cudaSetDevice(h->cfg.device));
  if (!h->win) return fail("ipc_export: no window (internal)");
  cudaIpcMemHandle_t mh;
  CU(cudaIpcGetMemHandle(&mh, h->win));
  static_assert(sizeof(cudaIpcMemHandle_t) == HSMC_GPU_IPC_BYTES, "ipc handle size");
  memcpy(out_blob, &mh, sizeof(mh));
  return 0;
}

extern "C" int hsmc_gpu_ipc_attach(hsmc_gpu* h, const void* left_blob, const void* right_blob) {
  if (!h || !left_blob || !right_blob) return fail("null argument");
  if (!h->win) return fail("ipc_attach: call ipc_export first");
  if (h->p2p) return fail("ipc_attach: already attached");
  CU(cudaSetDevice(h->cfg.device));
  cudaIpcMemHandle_t ml, mr;
  memcpy(&ml, left_blob, sizeof(ml));
  memcpy(&mr, right_blob, sizeof(mr));
  void* pl = nullptr; void* pr = nullptr;
  CU(cudaIpcOpenMemHandle(&pl, ml, cudaIpcMemLazyEnablePeerAccess));
  if (memcmp(&ml, &mr, sizeof(ml)) == 0) pr = pl;        // two ranks: both neighbours are the same process
  else CU(cudaIpcOpenMemHandle(&pr, mr, cudaIpcMemLazyEnablePeerAccess));
  h->win_left = (unsigned char*)pl;
  h->win_right = (unsigned char*)pr;
  unsigned long long offs[4];
  CU(cudaMemcpy(offs, h->win_left + WIN_OFFSETS_BYTE, sizeof(offs), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 4; k++) h->peer_off[k] = (size_t)offs[k];
  h->p2p = true;
  return 0;
}

extern "C" int hsmc_gpu_destroy(hsmc_gpu* h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  if (h->st) cudaStreamSynchronize(h->st);
  if (h->win_left) cudaIpcCloseMemHandle(h->win_left);
  if (h->win_right && h->win_right != h->win_left) cudaIpcCloseMemHandle(h->win_right);
  if (h->win) cudaFree(h->win);
  if (h->comm) ncclCommDestroy(h->comm);
  if (h->pos_in_win) h->pos[0] = h->pos[1] = nullptr, h->rel = nullptr;      // part of the window, freed above
  void* ptrs[] = {h->pos[0], h->pos[1], h->rel, h->key, h->rnk, h->cell_count, h->cell_start, h->bsum, h->d_cnt,
                  h->d_scratch, h->d_slot_of_id, h->d_io, h->send_l, h->send_r, h->recv_l, h->recv_r,
                  h->d_halo_cnt, h->d_sfargs, h->d_log, h->key_halo, h->rnk_halo, h->d_lay,
                  h->d_xoff, h->d_fuse, h->trec, h->traw, h->glists.list, h->glists.count, h->d_stamps};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->st2) cudaStreamDestroy(h->st2);
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
  return 0;
}

extern "C" int hsmc_gpu_create(hsmc_gpu** out, const hsmc_gpu_config* cfg, int64_t n_particles,
                               const double box[3]) {
  if (!out || !cfg || !box) return fail("null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail("no CUDA device: the B200 path has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail("invalid CUDA device ordinal");
  if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return fail("invalid rank/world");
  if (cfg->world > 1 && !cfg->nccl_id) return fail("world > 1 needs an NCCL unique id");
  if (n_particles <= 0 || n_particles >= (int64_t)INT32_MAX) return fail("invalid particle count");
  hsmc_gpu* h = new hsmc_gpu();
  h->cfg = *cfg;
  h->impl = cfg->sweep_impl & 0xff;
  h->xpart_world = (cfg->sweep_impl >> 8) & 0xff;
  if (h->impl != IMPL_AUTO && h->impl != IMPL_LEAN && h->impl != IMPL_GATHER && h->impl != IMPL_CELL_GLOBAL &&
      h->impl != IMPL_EPS0 && h->impl != IMPL_BLOCK_GLOBAL) {
    delete h;
    return fail("unknown sweep_impl variant");
  }
  if (h->cfg.cell_min == 0.0) h->cfg.cell_min = 1.0;
  if (h->cfg.cell_min < 1.0) { delete h; return fail("cell_min must be >= 1.0 (the particle diameter)"); }
  if (h->cfg.regrid_interval <= 0) h->cfg.regrid_interval = 1;
  h->N = n_particles;
  h->box[0] = box[0]; h->box[1] = box[1]; h->box[2] = box[2];
#define CUD(call)                                                                              \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      fail(std::string("CUDA: ") + cudaGetErrorString(e__) + " at " #call);                    \
      hsmc_gpu_destroy(h);                                                                     \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)
  CUD(cudaSetDevice(cfg->device));
  CUD(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
  CUD(cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking));
  CUD(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CUD(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  if (setup_grid(h)) { hsmc_gpu_destroy(h); return 1; }
  int W = cfg->world;
  if (W == 1) {
    h->cap = h->N;
    h->cap_halo = 0;
  } else {
    // owned share + two ghost layers, with head-room for density fluctuations
    double own_frac = (double)(h->g.own_hi - h->g.own_lo) / h->g.nx;
    double lay_frac = 1.0 / h->g.nx;
    h->cap = (int64_t)((own_frac + 2 * 2.5 * lay_frac) * h->N * 1.25) + 4096;
    // a cell layer of a crystal can hold 1.6x the mean (lattice planes beat against the cell grid)
    h->cap_halo = (int64_t)(2 * lay_frac * h->N * 2.5) + 4096;
  }
  if (W == 1) {
    CUD(cudaMalloc(&h->pos[0], sizeof(double4) * (size_t)h->cap));
    CUD(cudaMalloc(&h->pos[1], sizeof(double4) * (size_t)h->cap));
    CUD(cudaMalloc(&h->rel, sizeof(float4) * (size_t)h->cap));
  } else {
    // one allocation = the NVLink window (hsmc_gpu_ipc_export hands out its handle): message areas, block flags,
    // and the tables a neighbour stores its boundary layer into
    CUD(cudaMalloc(&h->win, win_bytes(h)));
    CUD(cudaMemset(h->win, 0, win_off_pos(h, 0)));
    h->pos[0] = reinterpret_cast<double4*>(h->win + win_off_pos(h, 0));
    h->pos[1] = reinterpret_cast<double4*>(h->win + win_off_pos(h, 1));
    h->rel = reinterpret_cast<float4*>(h->win + win_off_rel(h));
    h->pos_in_win = true;
    const unsigned long long offs[4] = {win_off_gdone(h), win_off_pos(h, 0), win_off_pos(h, 1), win_off_rel(h)};
    CUD(cudaMemcpy(h->win + WIN_OFFSETS_BYTE, offs, sizeof(offs), cudaMemcpyHostToDevice));
  }
  if (h->cap_halo) {
    CUD(cudaMalloc(&h->key_halo, sizeof(int) * (size_t)(2 * h->cap_halo)));
    CUD(cudaMalloc(&h->rnk_halo, sizeof(int) * (size_t)(2 * h->cap_halo)));
  }
  CUD(cudaMalloc(&h->d_cnt, sizeof(unsigned long long) * CNT_N));
  CUD(cudaMemset(h->d_cnt, 0, sizeof(unsigned long long) * CNT_N));
  CUD(cudaMalloc(&h->d_scratch, sizeof(unsigned long long) * SCRATCH_N));
  CUD(cudaMalloc(&h->d_halo_cnt, sizeof(int) * 4));
  CUD(cudaMemset(h->d_halo_cnt, 0, sizeof(int) * 4));
  CUD(cudaMalloc(&h->d_sfargs, sizeof(SfArgs)));
  CUD(cudaMalloc(&h->d_lay, sizeof(int) * 16));
  CUD(cudaMemset(h->d_lay, 0, sizeof(int) * 16));
  CUD(cudaMallocHost(&h->h_stage, sizeof(unsigned long long) * SCRATCH_N));
  if (ensure_cell_arrays(h)) { hsmc_gpu_destroy(h); return 1; }
  CUD(cudaFuncSetAttribute(k_sweep_lean<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  CUD(cudaFuncSetAttribute(k_sweep_lean<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  CUD(cudaFuncSetAttribute(k_sweep_lean<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  if (W > 1) {
    CUD(cudaMalloc(&h->send_l, sizeof(double4) * (size_t)h->cap_halo));
    CUD(cudaMalloc(&h->send_r, sizeof(double4) * (size_t)h->cap_halo));
    CUD(cudaMalloc(&h->recv_l, sizeof(double4) * (size_t)h->cap_halo));
    CUD(cudaMalloc(&h->recv_r, sizeof(double4) * (size_t)h->cap_halo));
    ncclUniqueId id;
    memcpy(&id, cfg->nccl_id, sizeof(id));
    ncclResult_t r = ncclCommInitRank(&h->comm, W, id, cfg->rank);
    if (r != ncclSuccess) {
      fail(std::string("NCCL: ") + ncclGetErrorString(r) + " at ncclCommInitRank");
      h->comm = nullptr;
      hsmc_gpu_destroy(h);
      return 1;
    }
  }
#undef CUD
  *out = h;
  return 0;
}

extern "C" int hsmc_gpu_get_info(hsmc_gpu* h, hsmc_gpu_info* o) {
  if (!h || !o) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  TRY(sync_layout(h));
  memset(o, 0, sizeof(*o));
  o->abi_version = HSMC_GPU_ABI_VERSION;
  o->rank = h->cfg.rank; o->world = h->cfg.world;
  o->n_total = h->N; o->n_owned = h->n_owned; o->n_local = h->n_local;
  o->cells[0] = h->g.nx; o->cells[1] = h->g.ny; o->cells[2] = h->g.nz;
  int x0 = (h->g.gx0 + h->g.own_lo) % h->g.nx;
  o->own_x0 = x0; o->own_x1 = x0 + (h->g.own_hi - h->g.own_lo);
  o->cell_size[0] = h->g.wx; o->cell_size[1] = h->g.wy; o->cell_size[2] = h->g.wz;
  o->box[0] = h->box[0]; o->box[1] = h->box[1]; o->box[2] = h->box[2];
  o->sweeps_done = h->sweeps_done;
  o->kernel_launches = h->launches;
  o->nccl_calls = h->nccl_calls;
  return 0;
}

extern "C" int hsmc_gpu_plan(const double box[3], double cell_min, int world, int rank, hsmc_gpu_info* o) {
  if (!box || !o) return fail("null argument");
  if (world < 1 || rank < 0 || rank >= world) return fail("invalid rank/world");
  hsmc_gpu tmp;
  tmp.cfg.device = 0; tmp.cfg.rank = rank; tmp.cfg.world = world; tmp.cfg.nccl_id = nullptr; tmp.cfg.seed = 0;
  tmp.cfg.cell_min = cell_min == 0.0 ? 1.0 : cell_min; tmp.cfg.regrid_interval = 1; tmp.cfg.sweep_impl = 0;
  tmp.impl = 0; tmp.xpart_world = 0;
  if (tmp.cfg.cell_min < 1.0) return fail("cell_min must be >= 1.0 (the particle diameter)");
  tmp.N = 1;
  tmp.box[0] = box[0]; tmp.box[1] = box[1]; tmp.box[2] = box[2];
  TRY(setup_grid(&tmp));
  memset(o, 0, sizeof(*o));
  o->abi_version = HSMC_GPU_ABI_VERSION;
  o->rank = rank; o->world = world;
  o->cells[0] = tmp.g.nx; o->cells[1] = tmp.g.ny; o->cells[2] = tmp.g.nz;
  int x0 = (tmp.g.gx0 + tmp.g.own_lo) % tmp.g.nx;
  o->own_x0 = x0; o->own_x1 = x0 + (tmp.g.own_hi - tmp.g.own_lo);
  o->cell_size[0] = tmp.g.wx; o->cell_size[1] = tmp.g.wy; o->cell_size[2] = tmp.g.wz;
  o->box[0] = box[0]; o->box[1] = box[1]; o->box[2] = box[2];
  return 0;
}

extern "C" int hsmc_gpu_plan_blocks(const double box[3], double cell_min, int world, int rank, int xpart_world,
                                    int64_t n_particles, hsmc_gpu_block_plan* o) {
  if (!box || !o) return fail("null argument");
  if (world < 1 || rank < 0 || rank >= world) return fail("invalid rank/world");
  if (n_particles < 1) return fail("invalid particle count");
  hsmc_gpu tmp;
  tmp.cfg.device = 0; tmp.cfg.rank = rank; tmp.cfg.world = world; tmp.cfg.nccl_id = nullptr; tmp.cfg.seed = 0;
  tmp.cfg.cell_min = cell_min == 0.0 ? 1.0 : cell_min; tmp.cfg.regrid_interval = 1; tmp.cfg.sweep_impl = 0;
  tmp.impl = IMPL_LEAN; tmp.xpart_world = (world == 1 && xpart_world > 1) ? xpart_world : 0;
  if (tmp.cfg.cell_min < 1.0) return fail("cell_min must be >= 1.0 (the particle diameter)");
  tmp.N = n_particles;
  tmp.box[0] = box[0]; tmp.box[1] = box[1]; tmp.box[2] = box[2];
  TRY(setup_grid(&tmp));
  setup_blocks(&tmp);
  memset(o, 0, sizeof(*o));
  o->ok = tmp.blk_ok ? 1 : 0;
  if (!tmp.blk_ok) return 0;
  const BlockCfg& b = tmp.blk;
  o->blocks[0] = b.nbx; o->blocks[1] = b.nby; o->blocks[2] = b.nbz;
  o->max_extent[0] = b.mbx; o->max_extent[1] = b.mby; o->max_extent[2] = b.mbz;
  o->ctas_per_phase = (b.nbx / 2) * (b.nby / 2) * (b.nbz / 2);
  o->staged_capacity = b.cap;
  o->smem_bytes = (int)tmp.blk_smem;
  if ((int)tmp.xoff.size() > HSMC_GPU_PLAN_MAX_XCUTS) return fail("plan_blocks: too many x blocks to report");
  o->n_xcuts = (int)tmp.xoff.size();
  // local layers -> global layers, counted from the first layer this rank owns (ascending, no wrap)
  const int x0 = (tmp.g.gx0 + tmp.g.own_lo) % tmp.g.nx;
  for (int k = 0; k < o->n_xcuts; k++) o->xcuts[k] = x0 + (tmp.xoff[k] - tmp.g.own_lo);
  return 0;
}

extern "C" void* hsmc_gpu_stream(hsmc_gpu* h) { return h ? (void*)h->st : nullptr; }

extern "C" int hsmc_gpu_sync(hsmc_gpu* h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->st));
  TRY(sync_layout(h));
  return 0;
}

extern "C" int hsmc_gpu_upload(hsmc_gpu* h, const double* rows, int64_t n_rows) {
  if (!h || !rows) return fail("null argument");
  h->io_packed = false;
  CU(cudaSetDevice(h->cfg.device));
  if (h->cfg.world == 1 && n_rows != h->N) return fail("upload: n_rows must equal the particle count");
  if (n_rows < 0 || n_rows > h->N) return fail("upload: bad row count");
  TRY(ensure_io(h, std::max<int64_t>(n_rows, 1)));
  CU(cudaMemcpyAsync(h->d_io, rows, sizeof(double) * 4 * (size_t)n_rows, cudaMemcpyHostToDevice, h->st));
  h->ghost1_stale = false;
  TRY(rebuild(h, reinterpret_cast<const double4*>(h->d_io), n_rows, 1, 1));
  TRY(sync_layout(h));
  h->have_conf = true;
  h->since_regrid = 0;
  return 0;
}

extern "C" int hsmc_gpu_download(hsmc_gpu* h, double* conf) {
  if (!h || !conf) return fail("null argument");
  if (!h->have_conf) return fail("download: no configuration uploaded");
  if (h->cfg.world != 1) return fail("download: full-table download needs world == 1; use download_owned");
  CU(cudaSetDevice(h->cfg.device));
  TRY(ensure_io(h, h->N));
  k_pack_by_id<<<nblk(h->N, 256), 256, 0, h->st>>>(h->pos[h->cur], 0, (int)h->N, h->d_io);
  h->launches++;
  CU(cudaGetLastError());
  h->io_packed = true;
  CU(cudaMemcpyAsync(conf, h->d_io, sizeof(double) * 4 * (size_t)h->N, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  return 0;
}

extern "C" int hsmc_gpu_pack_table(hsmc_gpu* h) {
  if (!h) return fail("null handle");
  if (!h->have_conf) return fail("download: no configuration uploaded");
  if (h->cfg.world != 1) return fail("pack_table: full-table download needs world == 1; use download_owned");
  CU(cudaSetDevice(h->cfg.device));
  TRY(ensure_io(h, h->N));
  k_pack_by_id<<<nblk(h->N, 256), 256, 0, h->st>>>(h->pos[h->cur], 0, (int)h->N, h->d_io);
  h->launches++;
  CU(cudaGetLastError());
  h->io_packed = true;
  return 0;
}

extern "C" int hsmc_gpu_fetch_rows(hsmc_gpu* h, int64_t first_row, int64_t n_rows, double* rows) {
  if (!h || !rows) return fail("null argument");
  if (!h->io_packed) return fail("fetch_rows: call pack_table first (and again after the configuration changed)");
  if (first_row < 0 || n_rows < 0 || first_row + n_rows > h->N) return fail("fetch_rows: row range out of bounds");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(rows, h->d_io + 4 * first_row, sizeof(double) * 4 * (size_t)n_rows, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  return 0;
}

extern "C" int hsmc_gpu_pin_host(void* ptr, size_t bytes, int pin) {
  if (!ptr) return fail("null argument");
  if (pin) CU(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  else CU(cudaHostUnregister(ptr));
  return 0;
}

extern "C" int hsmc_gpu_download_owned(hsmc_gpu* h, double* rows, int64_t capacity_rows, int64_t* n_rows) {
  if (!h || !rows || !n_rows) return fail("null argument");
  if (!h->have_conf) return fail("download: no configuration uploaded");
  CU(cudaSetDevice(h->cfg.device));
  TRY(sync_layout(h));
  if (capacity_rows < h->n_owned) return fail("download_owned: buffer too small");
  TRY(ensure_io(h, std::max<int64_t>(h->n_owned, 1)));
  h->io_packed = false;
  if (h->n_owned > 0) {
    k_pack_rows<<<nblk(h->n_owned, 256), 256, 0, h->st>>>(h->pos[h->cur], (int)h->own_first, (int)h->n_owned, h->d_io);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(rows, h->d_io, sizeof(double) * 4 * (size_t)h->n_owned, cudaMemcpyDeviceToHost, h->st));
  }
  CU(cudaStreamSynchronize(h->st));
  *n_rows = h->n_owned;
  return 0;
}

// boundary-layer refresh after the colour phases of x-parity `cx` (slab mode):
// the layer of that parity at the slab edge was updated; its ghost copy lives on the
// neighbour, slot-for-slot in the same order.
static int halo_refresh(hsmc_gpu* h, int cx) {
  ProfSpan span(h, 2);
  Grid& g = h->g;
  double4* p = h->pos[h->cur];
  const int cap_msg = (int)(h->cap_halo / 2);          // one layer with head-room
  const long long per = (long long)g.ny * g.nz;
  // x-parity 0: first owned layer -> left neighbour's right ghost; parity 1: last owned layer ->
  // right neighbour's left ghost
  const int src_layer = (cx == 0) ? g.own_lo : g.own_hi - 1;
  const int dst_layer = (cx == 0) ? g.nlx - 1 : 0;
  double4* sbuf = (cx == 0) ? h->send_l : h->send_r;
  double4* rbuf = (cx == 0) ? h->recv_r : h->recv_l;
  const int to = (cx == 0) ? left_of(h) : right_of(h), from = (cx == 0) ? right_of(h) : left_of(h);
  if (h->p2p) {
    // pack straight into the neighbour's window over NVLink, then flag; consume from my own window
    uint32_t seq = (cx == 0) ? ++h->seqB : ++h->seqC;
    unsigned char* peer = (cx == 0) ? h->win_left : h->win_right;
    sbuf = reinterpret_cast<double4*>(peer + (cx == 0 ? win_off_B(h) : win_off_C(h)));
    rbuf = reinterpret_cast<double4*>(h->win + (cx == 0 ? win_off_B(h) : win_off_C(h)));
    k_halo_pack<<<148, 256, 0, h->st>>>(g, p, h->cell_start, src_layer, sbuf, cap_msg, h->d_lay + 8);
    k_flag_post<<<1, 1, 0, h->st>>>(win_flag(peer, cx == 0 ? FLAG_B : FLAG_C), nullptr, seq);
    k_flag_wait<<<1, 1, 0, h->st>>>(win_flag(h->win, cx == 0 ? FLAG_B : FLAG_C), nullptr, seq);
    h->launches += 2;
  } else {
    k_halo_pack<<<148, 256, 0, h->st>>>(g, p, h->cell_start, src_layer, sbuf, cap_msg, h->d_lay + 8);
    NC(ncclGroupStart());
    NC(ncclSend(sbuf, (size_t)cap_msg * 4, ncclDouble, to, h->comm, h->st));
    NC(ncclRecv(rbuf, (size_t)cap_msg * 4, ncclDouble, from, h->comm, h->st));
    NC(ncclGroupEnd());
    h->nccl_calls += 2;
  }
  k_halo_unpack<<<nblk(per, 128), 128, 0, h->st>>>(g, p, h->rel, h->cell_start, dst_layer, rbuf, h->d_lay + 8);
  h->launches += 2;
  CU(cudaGetLastError());
  if (getenv("HSMC_DEBUG_SYNC")) {
    h->lay_valid = false;
    if (sync_layout(h)) return fail(std::string(g_err) + (cx ? " (after halo refresh 1)" : " (after halo refresh 0)"));
  }
  return 0;
}

// make the ghost layers current (no-op unless a refresh was deferred)
static int fresh_ghosts(hsmc_gpu* h) {
  if (h->cfg.world > 1 && h->ghost1_stale) {
    h->ghost1_stale = false;
    TRY(halo_refresh(h, 1));
  }
  return 0;
}

static int do_regrid(hsmc_gpu* h) {
  h->ghost1_stale = false;   // the rebuild's exchange refreshes both ghost layers
  draw_shift(h);
  if (h->cfg.world > 1) return rebuild(h, h->pos[h->cur], 0, 0, 0, true);
  return rebuild(h, h->pos[h->cur], h->N, 0, 0);
}

// trial tables of k_propose: one 16-byte record per particle slot
static int ensure_trial_tables(hsmc_gpu* h, bool logged) {
  if (!h->trec) CU(cudaMalloc(&h->trec, sizeof(uint4) * (size_t)h->cap));
  if (logged && !h->traw) CU(cudaMalloc(&h->traw, sizeof(uint4) * (size_t)h->cap));
  if (h->gather) {
    // a list holds at most one entry per cell of its colour
    const long long need = (long long)(h->g.own_hi - h->g.own_lo) * h->g.ny * h->g.nz / 8 + 1024;
    if (need > h->glists.stride) {
      if (h->glists.list) cudaFree(h->glists.list);
      h->glists.stride = need + need / 8;
      CU(cudaMalloc(&h->glists.list, sizeof(int) * (size_t)GATHER_LISTS * (size_t)h->glists.stride));
    }
    if (!h->glists.count) CU(cudaMalloc(&h->glists.count, sizeof(int) * GATHER_LISTS));
  }
  return 0;
}

static int sweep_once(hsmc_gpu* h, double dr_max, bool logged) {
  h->io_packed = false;
  if (h->since_regrid % h->cfg.regrid_interval == 0) TRY(do_regrid(h));
  h->since_regrid++;
  Grid& g = h->g;
  SweepArgs a;
  a.g = g;
  a.box = make_box(g.Lx, g.Ly, g.Lz, 1.0);
  a.dr_max = dr_max;
  a.key0 = (uint32_t)h->cfg.seed; a.key1 = (uint32_t)(h->cfg.seed >> 32);
  a.sweep_lo = (uint32_t)h->sweeps_done; a.sweep_hi = (uint32_t)(h->sweeps_done >> 32);
  a.eps = 1.0e-5f * (float)std::max(1.0, std::max(g.wx, std::max(g.wy, g.wz)));
  if (h->blk_ok) {
    a.eps = (h->impl == IMPL_EPS0) ? 0.0f : h->blk_eps;
    if (h->xoff_dirty) {
      if ((int64_t)h->xoff.size() > h->cap_xoff) {
        if (h->d_xoff) cudaFree(h->d_xoff);
        h->cap_xoff = (int64_t)h->xoff.size() + 64;
        CU(cudaMalloc(&h->d_xoff, sizeof(int) * (size_t)h->cap_xoff));
      }
      CU(cudaMemcpyAsync(h->d_xoff, h->xoff.data(), sizeof(int) * h->xoff.size(), cudaMemcpyHostToDevice, h->st));
      CU(cudaStreamSynchronize(h->st));        // the source is pageable host memory
      h->xoff_dirty = false;
    }
  }
  long long total = (long long)((g.own_hi - g.own_lo) / 2) * (g.ny / 2) * (g.nz / 2);
  const int T = 128;
  // Fused block phases: phases that need no halo exchange between them go into ONE launch whose
  // CTAs order themselves by per-block completion flags (see k_sweep_block): all eight on a single
  // GPU, 0-3 and 4-7 in slab mode.  HSMC_FUSE=0 launches the phases one by one (same chain).
  int fuse = 1;
  bool link = false;
  if (h->blk_ok) {
    if (h->fuse_mode < 0) { const char* e = getenv("HSMC_FUSE"); h->fuse_mode = (e && atoi(e) == 0) ? 0 : 1; }
    if (h->fuse_mode == 1 && h->blk.dbg == 0) fuse = (h->cfg.world > 1) ? 4 : 8;
    // slabs over NVLink windows, HSMC_SLAB_LINK=1: the right ghost layer arrives inside the launch (SlabLink), so all
    // eight phases fuse into one launch per sweep (needs a rebuild before every sweep: the ghost layout is published by
    // the rebuild).  Same chain, tested at 2 and 8 GPUs; NOT the default: measured at 8 GPUs it is 2 % slower than two
    // launches with the boundary-layer message in between (profiles/r02_bench_n8*.json) -- a slab's sweep is a chain of
    // eight block latencies either way.
    if (fuse == 4 && h->p2p && h->cfg.regrid_interval == 1 && h->impl != IMPL_GATHER) {
      const char* e = getenv("HSMC_SLAB_LINK");
      if (e && atoi(e) == 1) { fuse = 8; link = true; }
    }
    if (fuse > 1) {
      const int64_t need = 64 + (int64_t)h->blk.nbx * h->blk.nby * h->blk.nbz;
      if (need > h->cap_fuse) {
        if (h->d_fuse) cudaFree(h->d_fuse);
        h->cap_fuse = need + 1024;
        CU(cudaMalloc(&h->d_fuse, sizeof(unsigned int) * (size_t)h->cap_fuse));
        CU(cudaMemsetAsync(h->d_fuse, 0, sizeof(unsigned int) * (size_t)h->cap_fuse, h->st));
        h->fuse_tickets = 0;
      }
      h->blk.ticket = h->d_fuse;
      h->blk.done = h->d_fuse + 64;
    }
  }
  SlabLink sl;
  memset(&sl, 0, sizeof(sl));
  if (link) {
    sl.peer_pos[0] = reinterpret_cast<double4*>(h->win_left + h->peer_off[1]);
    sl.peer_pos[1] = reinterpret_cast<double4*>(h->win_left + h->peer_off[2]);
    sl.peer_rel = reinterpret_cast<float4*>(h->win_left + h->peer_off[3]);
    sl.peer_done = reinterpret_cast<unsigned int*>(h->win_left + h->peer_off[0]);
    sl.info = win_flag(h->win, WIN_INFO);
    sl.my_done = reinterpret_cast<const volatile unsigned int*>(h->win + win_off_gdone(h));
    sl.seq = ++h->seqG;
    sl.rebuild = h->seqR;
  }
  a.fuse = fuse; a.epoch = 0; a.ticket_base = 0;
  a.cx = a.cy = a.cz = a.phase = 0;
  h->gather = h->impl == IMPL_GATHER;
  if (h->gather) a.eps = 1.0e-5f * (float)std::max(1.0, std::max(g.wx, std::max(g.wy, g.wz)));
  if (h->blk_ok || h->gather) {
    // all proposals of the sweep, element-wise (they only depend on the particles' own positions at this point)
    ProfSpan span(h, 3);
    TRY(ensure_trial_tables(h, logged));
    int* lists = h->gather ? h->glists.list : nullptr;
    if (h->gather) CU(cudaMemsetAsync(h->glists.count, 0, sizeof(int) * GATHER_LISTS, h->st));
    if (logged)
      k_propose<true><<<nblk(h->cap, PROPOSE_THREADS), PROPOSE_THREADS, 0, h->st>>>(a, h->pos[h->cur], h->cell_start, h->ncell,
                                                                                   h->pos[h->cur ^ 1], h->trec, h->traw, lists,
                                                                                   h->glists.count, h->glists.stride);
    else
      k_propose<false><<<nblk(h->cap, PROPOSE_THREADS), PROPOSE_THREADS, 0, h->st>>>(a, h->pos[h->cur], h->cell_start, h->ncell,
                                                                                    h->pos[h->cur ^ 1], h->trec, nullptr, lists,
                                                                                    h->glists.count, h->glists.stride);
    h->launches++;
  }
  if (h->gather) fuse = 1;
  for (int ph = 0; ph < 8; ph++) {
    a.cx = (ph >> 2) & 1; a.cy = (ph >> 1) & 1; a.cz = ph & 1; a.phase = ph;
    if (fuse <= 1 || ph % fuse == 0) {     // else: launched together with phase ph - ph % fuse
    ProfSpan span(h, 0);
    if (h->gather) {
      // cell colour ph: first, second, third trial of every cell, then the cells holding more (4 launches)
      const long long cells_c = (long long)((g.own_hi - g.own_lo) / 2) * (g.ny / 2) * (g.nz / 2);
      for (int jj = 0; jj < 4; jj++) {
        // (second and later trials: fewer than a third of the cells have them)
        const int nb = (int)std::min<long long>(148LL * 8, std::max<long long>(1, nblk(jj == 0 ? cells_c : cells_c / 3, GATHER_THREADS)));
        if (logged)
          k_sweep_gather<true><<<nb, GATHER_THREADS, 0, h->st>>>(a, h->glists, ph, jj, h->pos[h->cur], h->rel, h->pos[h->cur ^ 1], h->trec,
                                                                 h->traw, h->cell_start, h->d_cnt, h->d_log, h->d_scratch,
                                                                 (long long)h->cap_log);
        else
          k_sweep_gather<false><<<nb, GATHER_THREADS, 0, h->st>>>(a, h->glists, ph, jj, h->pos[h->cur], h->rel, h->pos[h->cur ^ 1], h->trec,
                                                                  nullptr, h->cell_start, h->d_cnt, nullptr, nullptr, 0);
        h->launches++;
      }
      h->launches--;
    } else if (h->blk_ok) {
      // block phase ph (or phases ph .. ph+fuse-1): all blocks of block-index parity (cx,cy,cz);
      // each CTA runs the eight cell colours of its block
      const int nb = (h->blk.nbx / 2) * (h->blk.nby / 2) * (h->blk.nbz / 2) * fuse;
      if (fuse > 1) {
        a.epoch = ++h->fuse_epoch;
        if (a.epoch == 0) a.epoch = ++h->fuse_epoch;      // 0 means "never finished"
        a.ticket_base = h->fuse_tickets;
        h->fuse_tickets += (unsigned int)nb;
      }
      if (logged)
        k_sweep_lean<true><<<nb, LEAN_THREADS, h->blk_smem, h->st>>>(a, h->blk, sl, h->d_xoff, h->pos[h->cur], h->rel, h->pos[h->cur ^ 1],
                                                                     h->trec, h->traw, h->cell_start, h->d_cnt, h->d_log,
                                                                     h->d_scratch, (long long)h->cap_log);
      else
        k_sweep_lean<false><<<nb, LEAN_THREADS, h->blk_smem, h->st>>>(a, h->blk, sl, h->d_xoff, h->pos[h->cur], h->rel, h->pos[h->cur ^ 1],
                                                                      h->trec, nullptr, h->cell_start, h->d_cnt, nullptr, nullptr, 0);
    } else if (logged)
      k_sweep_phase<true><<<nblk(total, T), T, 0, h->st>>>(a, h->pos[h->cur], h->rel, h->cell_start, h->d_cnt, h->d_log,
                                                            h->d_scratch, (long long)h->cap_log);
    else
      k_sweep_phase<false><<<nblk(total, T), T, 0, h->st>>>(a, h->pos[h->cur], h->rel, h->cell_start, h->d_cnt, nullptr,
                                                             nullptr, 0);
    h->launches++;
    }
    if (h->cfg.world > 1) {
      // ghosts of parity cx are read only by phases of the other parity: one refresh
      // after the last phase of each parity is enough
      if (ph == 3 && !link) TRY(halo_refresh(h, 0));        // (link: delivered inside the launch)
      // the odd-parity boundary layer is read next by phases 0-3 of the FOLLOWING sweep; when
      // that sweep regrids first, its exchange rebuilds the ghost layers anyway, so the refresh
      // is deferred until somebody needs current ghosts (observables, end of the call)
      if (ph == 7) {
        if (h->since_regrid % h->cfg.regrid_interval == 0) h->ghost1_stale = true;
        else TRY(halo_refresh(h, 1));
      }
    }
  }
  CU(cudaGetLastError());
  h->sweeps_done++;
  return 0;
}

extern "C" int hsmc_gpu_sweep_nvt(hsmc_gpu* h, int n_sweeps, double dr_max) {
  if (!h) return fail("null handle");
  if (!h->have_conf) return fail("sweep: no configuration uploaded");
  if (!(dr_max > 0.0) || dr_max > 2.0 * std::min(h->g.wx, std::min(h->g.wy, h->g.wz)))
    return fail("sweep: dr_max out of range");
  CU(cudaSetDevice(h->cfg.device));
  for (int s = 0; s < n_sweeps; s++) TRY(sweep_once(h, dr_max, false));
  TRY(fresh_ghosts(h));
  return sync_layout(h);
}

extern "C" int hsmc_gpu_sweep_nvt_logged(hsmc_gpu* h, double dr_max, hsmc_gpu_trial* log, int64_t capacity,
                                         int64_t* n_logged) {
  if (!h || !log || !n_logged) return fail("null argument");
  if (!h->have_conf) return fail("sweep: no configuration uploaded");
  CU(cudaSetDevice(h->cfg.device));
  if (capacity > h->cap_log) {
    if (h->d_log) cudaFree(h->d_log);
    h->cap_log = capacity;
    CU(cudaMalloc(&h->d_log, sizeof(hsmc_gpu_trial) * (size_t)capacity));
  }
  CU(cudaMemsetAsync(h->d_scratch, 0, sizeof(unsigned long long), h->st));
  TRY(sweep_once(h, dr_max, true));
  TRY(fresh_ghosts(h));
  unsigned long long* hs = (unsigned long long*)h->h_stage;
  CU(cudaMemcpyAsync(hs, h->d_scratch, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  int64_t n = (int64_t)hs[0];
  if (n > capacity) return fail("sweep log capacity exceeded");
  CU(cudaMemcpyAsync(log, h->d_log, sizeof(hsmc_gpu_trial) * (size_t)n, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  *n_logged = n;
  return 0;
}

// HSMC_OBS_DOUBLE=1: the all-double pair kernels (thread per cell) instead of the fp32-prefiltered ones -- cross-check
static bool obs_all_double() {
  const char* e = getenv("HSMC_OBS_DOUBLE");          // (read per call: tests switch it inside one process)
  return e && atoi(e) != 0;
}

// ---- scaled overlap verdicts -----------------------------------------------------
static int overlap_flags(hsmc_gpu* h, const double* sf, int nn, int* flags_out) {
  if (nn < 1 || nn > MAX_SF) return fail("scaled overlap: number of scale factors must be in [1, 64]");
  Grid& g = h->g;
  SfArgs sa;
  memset(&sa, 0, sizeof(sa));
  sa.n = nn;
  double sfmin = sf[0];
  for (int k = 0; k < nn; k++) {
    if (!(sf[k] > 0.0)) return fail("scaled overlap: scale factor must be positive");
    sa.sf[k] = sf[k];
    sa.box[k] = make_box(g.Lx, g.Ly, g.Lz, sf[k]);
    sfmin = std::min(sfmin, sf[k]);
  }
  double wmin = std::min(g.wx, std::min(g.wy, g.wz));
  // cell * sf < 1: pairs two cells apart can overlap after the compression -- wider stencil (the reference
  // would silently miss them, moves.c:108); exact as long as two compressed cells still span a diameter
  const bool wide = wmin * sfmin < 1.0;
  if (wide && h->cfg.world > 1)
    return fail("scaled overlap: cell size too small for this compression on slabs (cell*sf < 1, one ghost layer); increase neigh_list");
  if (wide && 2.0 * wmin * sfmin < 1.0)
    return fail("scaled overlap: compression by more than a factor of two of the cell edge is not supported");
  sa.r2_skip = (1.0 / (sfmin * sfmin)) * (1.0 + 1e-6);
  SfArgs* hsa = (SfArgs*)h->h_stage;
  *hsa = sa;
  CU(cudaMemcpyAsync(h->d_sfargs, hsa, sizeof(SfArgs), cudaMemcpyHostToDevice, h->st));
  int* d_flags = (int*)h->d_scratch;
  CU(cudaMemsetAsync(d_flags, 0, sizeof(int) * MAX_SF, h->st));
  long long total = (long long)(g.own_hi - g.own_lo) * g.ny * g.nz;
  if (wide)
    k_overlap_scaled_wide<<<nblk(total, 128), 128, 0, h->st>>>(g, make_box(g.Lx, g.Ly, g.Lz, 1.0), (const SfArgs*)h->d_sfargs,
                                                                h->pos[h->cur], h->cell_start, 2, d_flags);
  else if (obs_all_double())
    k_overlap_scaled<<<nblk(total, 128), 128, 0, h->st>>>(g, make_box(g.Lx, g.Ly, g.Lz, 1.0), (const SfArgs*)h->d_sfargs,
                                                           h->pos[h->cur], h->cell_start, d_flags);
  else {
    TRY(sync_layout(h));
    if (h->n_owned > 0)
      k_overlap_scaled_f32<<<nblk(h->n_owned, 256), 256, 0, h->st>>>(g, make_box(g.Lx, g.Ly, g.Lz, 1.0), (const SfArgs*)h->d_sfargs,
                                                                    h->pos[h->cur], h->rel, h->cell_start, (int)h->own_first,
                                                                    (int)h->n_owned, d_flags);
  }
  h->launches++;
  CU(cudaGetLastError());
  if (h->cfg.world > 1) {
    NC(ncclAllReduce(d_flags, d_flags, nn, ncclInt, ncclMax, h->comm, h->st));
    h->nccl_calls++;
  }
  // h_stage currently holds the SfArgs copy source; wait for it before reuse
  CU(cudaStreamSynchronize(h->st));
  int* hf = (int*)h->h_stage;
  CU(cudaMemcpyAsync(hf, d_flags, sizeof(int) * nn, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  for (int k = 0; k < nn; k++) flags_out[k] = hf[k];
  return 0;
}

extern "C" int hsmc_gpu_overlap_scaled(hsmc_gpu* h, double sf, int* overlap) {
  if (!h || !overlap) return fail("null argument");
  if (!h->have_conf) return fail("no configuration uploaded");
  CU(cudaSetDevice(h->cfg.device));
  int f = 0;
  TRY(overlap_flags(h, &sf, 1, &f));
  *overlap = f;
  return 0;
}

extern "C" int hsmc_gpu_presst_flags(hsmc_gpu* h, const double* sf, int nn, int* no_overlap) {
  if (!h || !sf || !no_overlap) return fail("null argument");
  if (!h->have_conf) return fail("no configuration uploaded");
  CU(cudaSetDevice(h->cfg.device));
  for (int k0 = 0; k0 < nn; k0 += MAX_SF) {
    int m = std::min(MAX_SF, nn - k0);
    int f[MAX_SF];
    TRY(overlap_flags(h, sf + k0, m, f));
    for (int k = 0; k < m; k++) no_overlap[k0 + k] = f[k] ? 0 : 1;
  }
  return 0;
}

extern "C" int hsmc_gpu_rescale(hsmc_gpu* h, double sf, const double new_box[3]) {
  if (!h || !new_box) return fail("null argument");
  h->io_packed = false;
  if (!h->have_conf) return fail("no configuration uploaded");
  CU(cudaSetDevice(h->cfg.device));
  if (h->cfg.world > 1) {
    // slab mode: every resident particle (owned + ghosts) is rescaled in place; as long as the
    // cell COUNT per axis is unchanged the layers a rank owns scale with the box, so ownership
    // is preserved up to rounding and the ordinary regrid exchange repairs the rest
    TRY(fresh_ghosts(h));
    TRY(sync_layout(h));
    const int onx = h->g.nx, ony = h->g.ny, onz = h->g.nz;
    if (even_cells(new_box[0], h->cfg.cell_min) != onx || even_cells(new_box[1], h->cfg.cell_min) != ony ||
        even_cells(new_box[2], h->cfg.cell_min) != onz)
      return fail("rescale: this volume change alters the cell grid, which needs a redistribution of the slabs; "
                  "download_owned / upload the configuration (or run NpT on one GPU)");
  }
  k_rescale<<<nblk(h->n_local, 256), 256, 0, h->st>>>(h->pos[h->cur], (int)h->n_local, sf, new_box[0],
                                                       new_box[1], new_box[2]);
  h->launches++;
  CU(cudaGetLastError());
  h->box[0] = new_box[0]; h->box[1] = new_box[1]; h->box[2] = new_box[2];
  double sx = h->g.sx / h->g.wx, sy = h->g.sy / h->g.wy, sz = h->g.sz / h->g.wz;
  TRY(setup_grid(h));
  h->g.sx = sx * h->g.wx; h->g.sy = sy * h->g.wy; h->g.sz = sz * h->g.wz;
  if (h->cfg.world > 1) {
    h->ghost1_stale = false;
    TRY(rebuild(h, h->pos[h->cur], 0, 0, 0, true));
    return sync_layout(h);
  }
  TRY(rebuild(h, h->pos[h->cur], h->n_local, 0, 0));
  return 0;
}

// ---- observables ------------------------------------------------------------------
extern "C" int hsmc_gpu_widom(hsmc_gpu* h, uint64_t sample_id, int64_t first, int64_t count, int reduce,
                              int64_t* accepted) {
  if (!h || !accepted) return fail("null argument");
  if (!h->have_conf) return fail("no configuration uploaded");
  if (count < 0 || first < 0) return fail("widom: bad range");
  CU(cudaSetDevice(h->cfg.device));
  Grid& g = h->g;
  unsigned long long* d_acc = h->d_scratch;
  CU(cudaMemsetAsync(d_acc, 0, sizeof(unsigned long long), h->st));
  const long long chunk = 1LL << 30;
  for (long long off = 0; off < count; off += chunk) {
    long long c = std::min<long long>(chunk, count - off);
    k_widom<<<nblk(c, 256), 256, 0, h->st>>>(g, make_box(g.Lx, g.Ly, g.Lz, 1.0), h->pos[h->cur], h->cell_start,
                                             (uint32_t)h->cfg.seed, (uint32_t)(h->cfg.seed >> 32),
                                             (uint32_t)sample_id, (uint32_t)(sample_id >> 32), first + off, c, d_acc);
    h->launches++;
  }
  CU(cudaGetLastError());
  if (h->cfg.world > 1 && reduce) {
    NC(ncclAllReduce(d_acc, d_acc, 1, ncclUint64, ncclSum, h->comm, h->st));
    h->nccl_calls++;
  }
  unsigned long long* hs = (unsigned long long*)h->h_stage;
  CU(cudaMemcpyAsync(hs, d_acc, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  *accepted = (int64_t)hs[0];
  return 0;
}

extern "C" int hsmc_gpu_rdf_counts(hsmc_gpu* h, double dr_bin, int nn, uint64_t* counts) {
  return hsmc_gpu_rdf_counts_part(h, dr_bin, nn, 0, 1, counts);
}

extern "C" int hsmc_gpu_rdf_counts_part(hsmc_gpu* h, double dr_bin, int nn, int part, int nparts, uint64_t* counts) {
  if (!h || !counts) return fail("null argument");
  if (nparts < 1 || part < 0 || part >= nparts) return fail("rdf: bad part/nparts");
  if (!h->have_conf) return fail("no configuration uploaded");
  if (h->cfg.world > 1) return fail("rdf: the all-pairs histogram needs the whole configuration on one GPU (world == 1)");
  if (nn < 1 || nn > SCRATCH_N) return fail("rdf: bin count out of range");
  if (!(dr_bin > 0.0)) return fail("rdf: bin width must be positive");
  CU(cudaSetDevice(h->cfg.device));
  Grid& g = h->g;
  double rmax = dr_bin * nn + 1.0;   // compute_rdf.c:104
  double r2_pre = rmax * rmax * (1.0 + 1e-9);
  CU(cudaMemsetAsync(h->d_scratch, 0, sizeof(unsigned long long) * nn, h->st));
  int ntile = (int)((h->N + RDF_T - 1) / RDF_T);
  long long nblocks = (long long)ntile * (ntile + 1) / 2;
  if (nblocks > 0x7fffffffLL) return fail("rdf: system too large for the all-pairs histogram");
  size_t smem = sizeof(double) * 3 * RDF_T + sizeof(unsigned int) * (nn <= RDF_MAX_SMEM_BINS ? nn : 0);
  // this caller's share of the tile pairs (replicated configuration, pair triangle sharded k ways)
  const long long b0 = nblocks * part / nparts, b1 = nblocks * (part + 1) / nparts;
  if (b1 > b0) {
    k_rdf_pairs<<<(unsigned)(b1 - b0), RDF_T, smem, h->st>>>(h->pos[h->cur], (int)h->N, make_box(g.Lx, g.Ly, g.Lz, 1.0),
                                                             rmax, r2_pre, dr_bin, nn, ntile, b0, h->d_scratch);
    h->launches++;
  }
  CU(cudaGetLastError());
  unsigned long long* hs = (unsigned long long*)h->h_stage;
  CU(cudaMemcpyAsync(hs, h->d_scratch, sizeof(unsigned long long) * nn, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  for (int k = 0; k < nn; k++) counts[k] = hs[k];
  return 0;
}

extern "C" int hsmc_gpu_contact_counts(hsmc_gpu* h, double dr_bin, int nn, uint64_t* counts) {
  if (!h || !counts) return fail("null argument");
  if (!h->have_conf) return fail("no configuration uploaded");
  if (nn < 1 || nn > CONTACT_MAX_BINS) return fail("contact histogram: bin count out of range");
  if (!(dr_bin > 0.0)) return fail("contact histogram: bin width must be positive");
  CU(cudaSetDevice(h->cfg.device));
  Grid& g = h->g;
  double rmax = dr_bin * nn + 1.0;   // compute_press.c:116
  if (std::min(g.wx, std::min(g.wy, g.wz)) < rmax)
    return fail("The size of the cells in the neighbor list does not allow a correct calculation of the pressure, increase neigh_list");
  CU(cudaMemsetAsync(h->d_scratch, 0, sizeof(unsigned long long) * nn, h->st));
  long long total = (long long)(g.own_hi - g.own_lo) * g.ny * g.nz;
  if (obs_all_double())
    k_contact_hist<<<nblk(total, 128), 128, 0, h->st>>>(g, make_box(g.Lx, g.Ly, g.Lz, 1.0), h->pos[h->cur],
                                                         h->cell_start, rmax, dr_bin, nn, h->d_scratch);
  else {
    TRY(sync_layout(h));
    if (h->n_owned > 0)
      k_contact_hist_f32<<<nblk(h->n_owned, 256), 256, 0, h->st>>>(g, make_box(g.Lx, g.Ly, g.Lz, 1.0), h->pos[h->cur], h->rel,
                                                                  h->cell_start, (int)h->own_first, (int)h->n_owned, rmax,
                                                                  dr_bin, nn, h->d_scratch);
  }
  h->launches++;
  CU(cudaGetLastError());
  if (h->cfg.world > 1) {
    NC(ncclAllReduce(h->d_scratch, h->d_scratch, nn, ncclUint64, ncclSum, h->comm, h->st));
    h->nccl_calls++;
  }
  unsigned long long* hs = (unsigned long long*)h->h_stage;
  CU(cudaMemcpyAsync(hs, h->d_scratch, sizeof(unsigned long long) * nn, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  for (int k = 0; k < nn; k++) counts[k] = hs[k];
  return 0;
}


extern "C" int hsmc_gpu_order_parameter(hsmc_gpu* h, int l, double rmax, double* ql_ave) {
  if (!h || !ql_ave) return fail("null argument");
  if (!h->have_conf) return fail("no configuration uploaded");
  if (l < 0 || l > QL_MAX_L) return fail("order parameter: order l must be in [0, 12]");
  CU(cudaSetDevice(h->cfg.device));
  Grid& g = h->g;
  if (!(rmax > 0.0) || rmax > std::min(g.wx, std::min(g.wy, g.wz)))
    return fail("order parameter: the bond cutoff must not exceed the cell edge (27-cell stencil); reduce it to the neighbour-list size as the reference does");
  TRY(sync_layout(h));
  static bool coef_ready[64] = {false};
  if (!coef_ready[h->cfg.device & 63]) {
    double A[(QL_MAX_L + 1) * (QL_MAX_L + 1)] = {0}, B[(QL_MAX_L + 1) * (QL_MAX_L + 1)] = {0}, D[QL_MAX_L + 2] = {0};
    for (int k = 1; k <= QL_MAX_L; k++)
      for (int m = 0; m < k; m++) {
        A[k * (QL_MAX_L + 1) + m] = sqrt((4.0 * k * k - 1.0) / ((double)k * k - (double)m * m));
        B[k * (QL_MAX_L + 1) + m] = sqrt((((double)k - 1.0) * (k - 1.0) - (double)m * m) / (4.0 * (k - 1.0) * (k - 1.0) - 1.0));
      }
    for (int m = 1; m <= QL_MAX_L + 1; m++) D[m] = sqrt((2.0 * m + 1.0) / (2.0 * m));
    CU(cudaMemcpyToSymbol(c_ql_A, A, sizeof(A)));
    CU(cudaMemcpyToSymbol(c_ql_B, B, sizeof(B)));
    CU(cudaMemcpyToSymbol(c_ql_D, D, sizeof(D)));
    coef_ready[h->cfg.device & 63] = true;
  }
  const int64_t n = h->n_owned;
  const int nb = nblk(std::max<int64_t>(n, 1), QL_T);
  // block partials live in the (free) ping-pong table; the final sum goes to the scratch block
  double* d_partial = reinterpret_cast<double*>(h->pos[h->cur ^ 1]);
  if ((int64_t)nb * (int64_t)sizeof(double) > h->cap * (int64_t)sizeof(double4)) return fail("order parameter: scratch too small");
  double* d_out = reinterpret_cast<double*>(h->d_scratch);
  k_order_param<<<nb, QL_T, 0, h->st>>>(g, make_box(g.Lx, g.Ly, g.Lz, 1.0), h->pos[h->cur], h->cell_start,
                                        (int)h->own_first, (int)n, l, rmax, d_partial);
  k_sum_partials<<<1, 256, 0, h->st>>>(d_partial, nb, d_out);
  h->launches += 2;
  CU(cudaGetLastError());
  if (h->cfg.world > 1) {
    NC(ncclAllReduce(d_out, d_out, 1, ncclDouble, ncclSum, h->comm, h->st));
    h->nccl_calls++;
  }
  double* hs = (double*)h->h_stage;
  CU(cudaMemcpyAsync(hs, d_out, sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  *ql_ave = hs[0] / (double)h->N;     // compute_order_parameter.c:92-96
  return 0;
}

extern "C" int hsmc_gpu_min_dist2(hsmc_gpu* h, double* out) {
  if (!h || !out) return fail("null argument");
  if (!h->have_conf) return fail("no configuration uploaded");
  CU(cudaSetDevice(h->cfg.device));
  Grid& g = h->g;
  unsigned long long* hs = (unsigned long long*)h->h_stage;
  double big = 1e300;
  memcpy(&hs[0], &big, 8);
  CU(cudaMemcpyAsync(h->d_scratch, hs, 8, cudaMemcpyHostToDevice, h->st));
  long long total = (long long)(g.own_hi - g.own_lo) * g.ny * g.nz;
  k_min_r2<<<(int)std::min<long long>(nblk(total, 128), 148 * 16), 128, 0, h->st>>>(g, make_box(g.Lx, g.Ly, g.Lz, 1.0), h->pos[h->cur],
                                                                                     h->cell_start, h->d_scratch);
  h->launches++;
  CU(cudaGetLastError());
  if (h->cfg.world > 1) {
    NC(ncclAllReduce(h->d_scratch, h->d_scratch, 1, ncclDouble, ncclMin, h->comm, h->st));
    h->nccl_calls++;
  }
  CU(cudaStreamSynchronize(h->st));
  CU(cudaMemcpyAsync(hs, h->d_scratch, 8, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  memcpy(out, &hs[0], 8);
  return 0;
}

// ---- counters ----------------------------------------------------------------------
static int read_counters(hsmc_gpu* h, unsigned long long* c) {
  unsigned long long* hs = (unsigned long long*)h->h_stage;
  if (h->cfg.world > 1) {
    unsigned long long* tmp = h->d_scratch;
    NC(ncclAllReduce(h->d_cnt, tmp, CNT_N, ncclUint64, ncclSum, h->comm, h->st));
    h->nccl_calls++;
    CU(cudaMemcpyAsync(hs, tmp, sizeof(unsigned long long) * CNT_N, cudaMemcpyDeviceToHost, h->st));
  } else {
    CU(cudaMemcpyAsync(hs, h->d_cnt, sizeof(unsigned long long) * CNT_N, cudaMemcpyDeviceToHost, h->st));
  }
  CU(cudaStreamSynchronize(h->st));
  for (int k = 0; k < CNT_N; k++) c[k] = hs[k];
  return 0;
}

extern "C" int hsmc_gpu_counters(hsmc_gpu* h, int64_t out[6]) {
  if (!h || !out) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  unsigned long long c[CNT_N];
  TRY(read_counters(h, c));
  out[0] = (int64_t)c[CNT_TRIALS];
  out[1] = (int64_t)c[CNT_ACC];
  out[2] = (int64_t)(c[CNT_REJ_OVERLAP] + c[CNT_REJ_CELL]);
  out[3] = h->vol_cnt[0]; out[4] = h->vol_cnt[1]; out[5] = h->vol_cnt[2];
  return 0;
}

extern "C" int hsmc_gpu_cell_rejects(hsmc_gpu* h, int64_t* out) {
  if (!h || !out) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  unsigned long long c[CNT_N];
  TRY(read_counters(h, c));
  *out = (int64_t)c[CNT_REJ_CELL];
  return 0;
}

extern "C" int hsmc_gpu_reset_counters(hsmc_gpu* h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemsetAsync(h->d_cnt, 0, sizeof(unsigned long long) * CNT_N, h->st));
  h->vol_cnt[0] = h->vol_cnt[1] = h->vol_cnt[2] = 0;
  return 0;
}

extern "C" int hsmc_gpu_add_vol_move(hsmc_gpu* h, int accepted) {
  if (!h) return fail("null handle");
  h->vol_cnt[0]++;
  if (accepted) h->vol_cnt[1]++; else h->vol_cnt[2]++;
  return 0;
}

extern "C" int hsmc_gpu_profile(hsmc_gpu* h, int enable) {
  if (!h) return fail("null handle");
  h->prof = enable != 0;
  return 0;
}

extern "C" int hsmc_gpu_profile_read(hsmc_gpu* h, double ms[HSMC_GPU_PROFILE_BUCKETS],
                                     int64_t groups[HSMC_GPU_PROFILE_BUCKETS]) {
  if (!h || !ms || !groups) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->st));
  for (auto& sp : h->spans) {
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, sp.a, sp.b));
    h->prof_ms[sp.bucket] += t;
    h->prof_n[sp.bucket] += 1;
    h->ev_pool.push_back(sp.a);
    h->ev_pool.push_back(sp.b);
  }
  h->spans.clear();
  for (int k = 0; k < HSMC_GPU_PROFILE_BUCKETS; k++) {
    ms[k] = h->prof_ms[k]; groups[k] = h->prof_n[k];
    h->prof_ms[k] = 0; h->prof_n[k] = 0;
  }
  return 0;
}

// tuning aid (not part of the drop-in surface): the raw device counters; [4..6] count the blocks that left the
// staged path of k_sweep_lean (staging capacity / a cell with more than 8 particles / trial-list capacity)
extern "C" int hsmc_gpu_debug_counters(hsmc_gpu* h, uint64_t out[8]) {
  if (!h || !out) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  unsigned long long* hs = (unsigned long long*)h->h_stage;
  CU(cudaMemcpyAsync(hs, h->d_cnt, sizeof(unsigned long long) * CNT_N, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  for (int k = 0; k < CNT_N; k++) out[k] = hs[k];
  return 0;
}

// tuning aid: clock cycles per stage of k_sweep_lean summed over blocks (HSMC_BLOCK_STAMPS=1), [31] = blocks; read + reset
extern "C" int hsmc_gpu_debug_stamps(hsmc_gpu* h, uint64_t out[32]) {
  if (!h || !out) return fail("null argument");
  memset(out, 0, sizeof(uint64_t) * 32);
  if (!h->d_stamps) return 0;
  CU(cudaSetDevice(h->cfg.device));
  unsigned long long* hs = (unsigned long long*)h->h_stage;
  CU(cudaMemcpyAsync(hs, h->d_stamps, sizeof(unsigned long long) * 32, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemsetAsync(h->d_stamps, 0, sizeof(unsigned long long) * 32, h->st));
  CU(cudaStreamSynchronize(h->st));
  for (int k = 0; k < 32; k++) out[k] = hs[k];
  return 0;
}

extern "C" int hsmc_gpu_set_sweep_counter(hsmc_gpu* h, uint64_t sweeps_done) {
  if (!h) return fail("null handle");
  h->sweeps_done = sweeps_done;
  return 0;
}

__global__ void k_selftest_u01(unsigned long long* __restrict__ out) {
  unsigned long long bad = 0;
  unsigned int first = 0xffffffffu;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < (1ull << 32); r += stride) {
    uint32_t raw = (uint32_t)r;
    if (hsmc_u01_fast(raw) != hsmc_u01_div(raw)) { bad++; first = min(first, raw); }
  }
  if (bad) { atomicAdd(&out[0], bad); atomicMin(&out[1], (unsigned long long)first); }
}

extern "C" int hsmc_gpu_selftest_u01(hsmc_gpu* h, uint64_t* n_mismatch, uint32_t* first_bad) {
  if (!h || !n_mismatch || !first_bad) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  unsigned long long* hs = (unsigned long long*)h->h_stage;
  hs[0] = 0; hs[1] = 0xffffffffull;
  CU(cudaMemcpyAsync(h->d_scratch, hs, 16, cudaMemcpyHostToDevice, h->st));
  k_selftest_u01<<<148 * 8, 256, 0, h->st>>>(h->d_scratch);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->st));
  CU(cudaMemcpyAsync(hs, h->d_scratch, 16, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  *n_mismatch = hs[0];
  *first_bad = (uint32_t)hs[1];
  return 0;
}

// ---- parity entry points ---------------------------------------------------------
extern "C" int hsmc_gpu_trial_verdicts(hsmc_gpu* h, int n, const int* idx, const double* xyz, double sf, int* flags) {
  if (!h || !idx || !xyz || !flags) return fail("null argument");
  if (!h->have_conf) return fail("no configuration uploaded");
  if (h->cfg.world != 1) return fail("trial_verdicts: world == 1 only");
  CU(cudaSetDevice(h->cfg.device));
  Grid& g = h->g;
  if (sf < 1.0 && std::min(g.wx, std::min(g.wy, g.wz)) * sf < 1.0)
    return fail("scaled overlap: cell size too small for this compression (cell*sf < 1); increase neigh_list");
  double* d_xyz; int *d_idx, *d_fl;
  CU(cudaMalloc(&d_xyz, sizeof(double) * 3 * (size_t)n));
  CU(cudaMalloc(&d_idx, sizeof(int) * (size_t)n));
  CU(cudaMalloc(&d_fl, sizeof(int) * (size_t)n));
  CU(cudaMemcpyAsync(d_xyz, xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(d_idx, idx, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->st));
  k_trial_points<<<nblk(n, 128), 128, 0, h->st>>>(g, make_box(g.Lx, g.Ly, g.Lz, sf), sf, h->pos[h->cur],
                                                   h->cell_start, d_idx, d_xyz, n, d_fl);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(flags, d_fl, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  cudaFree(d_xyz); cudaFree(d_idx); cudaFree(d_fl);
  return 0;
}

extern "C" int hsmc_gpu_widom_verdicts(hsmc_gpu* h, int n, const double* xyz, int* flags) {
  if (!h || !xyz || !flags) return fail("null argument");
  if (!h->have_conf) return fail("no configuration uploaded");
  if (h->cfg.world != 1) return fail("widom_verdicts: world == 1 only");
  CU(cudaSetDevice(h->cfg.device));
  Grid& g = h->g;
  double* d_xyz; int* d_fl;
  CU(cudaMalloc(&d_xyz, sizeof(double) * 3 * (size_t)n));
  CU(cudaMalloc(&d_fl, sizeof(int) * (size_t)n));
  CU(cudaMemcpyAsync(d_xyz, xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, h->st));
  k_widom_points<<<nblk(n, 128), 128, 0, h->st>>>(g, make_box(g.Lx, g.Ly, g.Lz, 1.0), h->pos[h->cur],
                                                   h->cell_start, d_xyz, n, d_fl);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(flags, d_fl, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  cudaFree(d_xyz); cudaFree(d_fl);
  return 0;
}
