// sweep_lean.cuh -- K2, the default sweep: k_propose (all proposals of a sweep, up front) and k_sweep_lean
// (block-resident packed-fp32 stencil filter).  Two-level checkerboard, same Markov chain as the all-double
// global-memory evaluation (sweep_impl 5): every particle gets one trial per sweep, each trial is the reference's
// part_move() (moves.c:27-80), a trial that leaves its cell is rejected.
//
// Design notes (profiles/r01_ncu_k_sweep_block_fused.txt: the round-1 kernel generated, staged, scanned and committed
// in one 128-register kernel: 16 warps/SM, 38 % issue, 55 % shared-memory bank conflicts, prologue 31 % of warp time;
// profiles/r02_ncu_k_block_plan_v1.txt: a per-block "plan" kernel that pre-built the trial lists cost as many
// instructions as the sweep itself -- dropped):
//
//  * A trial point depends on nothing but the particle's own position at the start of the sweep (a particle only ever
//    moves by its own trial) and on Philox(global cell, trial index, sweep).  k_propose, an element-wise throughput
//    kernel (one thread per particle), generates them all: new position in double (moves.c:52-57, 215-226) into the
//    idle half of the ping-pong master table, the cell test, and a 16-byte trial record {fp32 shadow of the new
//    position, slot | accept-able} stored in TRIAL ORDER inside the cell's slot range (ascending particle id).
//    k_sweep_lean has no Philox and no double arithmetic on its hot path: 64 registers, 32 warps per SM.
//  * The sweep kernel derives everything else itself: one thread per staged (x,y) row reads the row's CSR entries, a
//    warp scan places the rows, one warp per row stages the fp32 shadow as block-relative coordinates in PAIRS
//    {x0,x1,y0,y1} + {z0,z1}, so the stencil filter runs on packed fp32x2 instructions (FADD2 / FMUL2 / FFMA2 +
//    FMNMX3: 7 math instructions and two loads per two neighbours instead of 14 + 2).
//  * Per cell colour one warp cuts the colour's cells (z fastest, so that the stencil rows of neighbouring lanes start at
//    nearby shared-memory addresses) into chunks of at most 32 cells and 32 trials; a chunk is one warp's work between
//    two colour barriers, lane = trial.  The trials of one cell are always in one chunk, in adjacent lanes, in order.
//  * Where the time goes (profiles/r02_block_stamps.txt, r02_ncu_k_sweep_lean*.txt): the kernel is bound by the
//    latency of a block's dependent stages, not by arithmetic; DESIGN.md section 5 lists what was measured and dropped.
//
// Exactness is unchanged: the fp32 minimum r^2 is a filter with a rigorous error bound eps (setup_blocks); min < 1 - eps
// is a certain overlap, min > 1 + eps a certain miss, anything in between is re-evaluated from the master table with
// the reference's double arithmetic (moves.c:400-431).  Trials of one cell that fall into the same 32-lane chunk are
// ordered with warp shuffles: a later trial sees its earlier mates at the positions their own trials left them in.
#pragma once

// 256 threads x 4 CTAs per SM (64 registers): 2.02 ms per sweep at the benchmark against 2.09 for 192 x 5, 2.04 for
// 128 x 7 and 224 x 4, 2.18 for 320 x 3 (profiles/r02_variants_rejected.txt); the block shape follows (setup_blocks)
#ifndef LEAN_THREADS
#define LEAN_THREADS 256
#endif
#ifndef LEAN_MIN_CTAS
#define LEAN_MIN_CTAS 4
#endif
#ifndef LEAN_COMMIT_B
#define LEAN_COMMIT_B 2         // items a thread commits per round (the loads of all of them before any store);
                                // measured at the benchmark: 1 -> 1.80, 2 -> 1.71, 3 -> 1.73, 4 -> 1.76 ms per sweep
#endif
#ifndef LEAN_NP
#define LEAN_NP 3               // pair-records (2 entries each) per stencil row in straight-line code; longer rows: loop
                                // (3: -3 % against 4 at the benchmark, rows of three cells hold 2.7 particles; 2: -2 %)
#endif
#define LEAN_PAD 12             // far-away entries after the last staged particle (covers the over-scan)
#define LEAN_FAR 1.0e15f
#define LEAN_MAX_OCC 8          // most particles per cell on the staged path (3-bit trial index)
#define LEAN_MAX_ROWS 144       // (mbx+2)*(mby+2): blocks of up to 10 x 10 cells across
#define LEAN_COL_CHUNKS 32       // most chunks one cell colour of a block may need
#define LEAN_MAX_CHUNKS 30       // chunks per cell colour of a block
#define PROPOSE_THREADS 256
#define GATHER_LISTS_N 32        // trial lists of k_sweep_gather: 8 colours x {first, second, third trial, cells with more}

// trial record code: slot of the particle inside its cell (4 bits) | bit 4: the trial stays inside its cell
#define TREC_ACT 16u

// item of a warp's trial queue: first staged index of the cell 12 | rx 4 | ry 4 | rz 5 | j 3 | n-1 3
#define LEAN_ITEM(ob, rx, ry, rz, j, n1) ((unsigned)(ob) | ((unsigned)(rx) << 12) | ((unsigned)(ry) << 16) | ((unsigned)(rz) << 20) | ((unsigned)(j) << 25) | ((unsigned)(n1) << 28))

struct BlkGeom {
  int xa, ya, za;            // first interior cell (local x layer, y, z)
  int ex, ey, ez;            // interior extent
  int x0, y0, z0;            // region origin (may be -1: periodic wrap)
  int nrx, nry, lenz, nrows;
  int zs;                    // wrapped z of the region's first cell
  bool zwrap;
  int bxi, byi, bzi;         // block index (k_sweep_lean)
};

__device__ __forceinline__ BlkGeom blk_geom(const Grid& g, const BlockCfg& bc, const int* __restrict__ xoff, int bxi,
                                            int byi, int bzi) {
  BlkGeom q;
  q.xa = xoff[bxi];
  const int xb = xoff[bxi + 1];
  // (32-bit: block index x cells per axis stays far below 2^31)
  q.ya = (byi * g.ny) / bc.nby;
  const int yb = ((byi + 1) * g.ny) / bc.nby;
  q.za = (bzi * g.nz) / bc.nbz;
  const int zb = ((bzi + 1) * g.nz) / bc.nbz;
  q.ex = xb - q.xa; q.ey = yb - q.ya; q.ez = zb - q.za;
  q.x0 = q.xa - 1; q.y0 = q.ya - 1; q.z0 = q.za - 1;
  q.nrx = q.ex + 2; q.nry = q.ey + 2; q.lenz = q.ez + 2;
  q.nrows = q.nrx * q.nry;
  q.zs = (q.z0 < 0) ? q.z0 + g.nz : q.z0;
  q.zwrap = q.zs + q.lenz > g.nz;
  return q;
}

// global (x,y) row index of staged row (rx, ry)
__device__ __forceinline__ long long blk_global_row(const Grid& g, const BlkGeom& q, int rx, int ry) {
  int lx = q.x0 + rx;
  if (g.wrap_x) { if (lx < 0) lx += g.nlx; else if (lx >= g.nlx) lx -= g.nlx; }
  int y = q.y0 + ry;
  if (y < 0) y += g.ny; else if (y >= g.ny) y -= g.ny;
  return (long long)lx * g.ny + y;
}

// exact re-evaluation of a whole stencil from the master table (moves.c:157-212, 400-431).  The block commits its
// accepted moves to the master table at its end (one parallel pass): until then a particle of this block whose own
// trial was accepted (bit set in s_pacc) is at its proposal, prop[slot].
// (`pos` deliberately not const __restrict__: neighbouring blocks wrote it earlier in this launch)
__device__ __noinline__ bool block_exact_rescan(const double4* pos, const double4* __restrict__ prop, const BlockRow* s_row,
                                                const unsigned short* s_cz, const unsigned int* s_pacc, int cz_stride,
                                                int nry, int rxc, int ryc, int rz, int sel, double xn, double yn, double zn,
                                                const Box box) {
  for (int dx = -1; dx <= 1; dx++)
    for (int dy = -1; dy <= 1; dy++) {
      const int row = (rxc + dx) * nry + ryc + dy;
      const BlockRow rw = s_row[row];
      const unsigned short* cp = s_cz + row * cz_stride + rz;
      const int b = cp[-1], e = cp[2];
      for (int k = b; k < e; k++) {
        if (k == sel) continue;
        const int o = k - rw.off;
        const int gs = (o < rw.cntA) ? rw.gbA + o : rw.gbB + o - rw.cntA;
        const double4 q = ((s_pacc[k >> 5] >> (k & 31)) & 1u) ? prop[gs] : pos[gs];
        if (pair_r2(xn, yn, zn, q.x, q.y, q.z, box) < 1.0) return true;
      }
    }
  return false;
}

// ---------------------------------------------------------------------------------------------------
// k_propose: one thread per resident particle.  Trial index j = number of particles of the same cell with a
// smaller id; the trial record is stored at slot (cell start + j), i.e. in trial order, the proposal at the particle's slot.
// ---------------------------------------------------------------------------------------------------
template <bool LOG>
__global__ void __launch_bounds__(PROPOSE_THREADS)
k_propose(SweepArgs a, const double4* __restrict__ pos, const int* __restrict__ cs, long long ncell,
          double4* __restrict__ prop, uint4* __restrict__ trec, uint4* __restrict__ traw, int* __restrict__ lists,
          int* __restrict__ lcount, long long lstride) {
  const Grid& g = a.g;
  int gs = blockIdx.x * PROPOSE_THREADS + threadIdx.x;
  const bool live = gs < cs[ncell];
  if (!live) {
    if (!lists) return;
    gs = 0;                      // (keeps the CTA's barriers of the list building; nothing is written for this thread)
  }
  const double4 p = pos[gs];
  const long long c = local_cell(g, p.x, p.y, p.z);       // the cell the counting sort put it in
  const int b = cs[c], n = cs[c + 1] - b;
  int j = 0;
  for (int k = 0; k < n; k++) j += pos[b + k].w < p.w;
  const int iz = (int)(c % g.nz);
  const long long r = c / g.nz;
  const int iy = (int)(r % g.ny), l = (int)(r / g.ny);
  const int gx = (g.gx0 + l >= g.nx) ? g.gx0 + l - g.nx : g.gx0 + l;
  const long long gcell = ((long long)gx * g.ny + iy) * g.nz + iz;
  // the reference's trial point (moves.c:52-57), apply_pbc (moves.c:215-226), then the cell test
  const Philox4 rn = philox4x32_10((uint32_t)gcell, (HSMC_STREAM_MOVE << 24) | (uint32_t)j, a.sweep_lo, a.sweep_hi, a.key0, a.key1);
  double xn = p.x + (hsmc_u01(rn.v[0]) - 0.5) * a.dr_max;
  double yn = p.y + (hsmc_u01(rn.v[1]) - 0.5) * a.dr_max;
  double zn = p.z + (hsmc_u01(rn.v[2]) - 0.5) * a.dr_max;
  if (xn > g.Lx) xn -= g.Lx; else if (xn < 0.0) xn += g.Lx;
  if (yn > g.Ly) yn -= g.Ly; else if (yn < 0.0) yn += g.Ly;
  if (zn > g.Lz) zn -= g.Lz; else if (zn < 0.0) zn += g.Lz;
  const bool act = axis_cell(xn, g.sx, g.iwx, g.nx) == gx && axis_cell(yn, g.sy, g.iwy, g.ny) == iy &&
                   axis_cell(zn, g.sz, g.iwz, g.nz) == iz;
  float4 nrel = make_float4(0.f, 0.f, 0.f, 0.f);
  if (act) nrel = make_rel(g, gx, iy, iz, xn, yn, zn);
  const unsigned int code = (unsigned int)min(gs - b, 15) | (act ? TREC_ACT : 0u);
  if (live) {
    trec[b + j] = make_uint4(__float_as_uint(nrel.x), __float_as_uint(nrel.y), __float_as_uint(nrel.z), code);
    prop[gs] = make_double4(xn, yn, zn, p.w);
    if (LOG) traw[b + j] = make_uint4(rn.v[0], rn.v[1], rn.v[2], (unsigned int)(int)p.w);
  }
  // k_sweep_gather: per cell colour, the cells holding a first / second / third trial and those holding more
  // (owned cells only).  The 32 list counters are bumped once per CTA: ranks inside the CTA come from shared-memory
  // counters, then one global atomic per list reserves the CTA's range.
  if (lists) {
    __shared__ int s_n[GATHER_LISTS_N], s_base[GATHER_LISTS_N];
    if (threadIdx.x < GATHER_LISTS_N) s_n[threadIdx.x] = 0;
    __syncthreads();
    const bool mine = live && j <= 3 && l >= g.own_lo && l < g.own_hi;
    const int li = ((((gx & 1) << 2) | ((iy & 1) << 1) | (iz & 1)) << 2) | (j & 3);
    int rank = 0;
    if (mine) rank = atomicAdd(&s_n[li], 1);
    __syncthreads();
    if (threadIdx.x < GATHER_LISTS_N && s_n[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&lcount[threadIdx.x], s_n[threadIdx.x]);
    __syncthreads();
    if (mine) {
      const int at = s_base[li] + rank;
      if (at < lstride) lists[(long long)li * lstride + at] = (l << 20) | (iy << 10) | iz;     // (fewer than 1024 cells per axis and rank)
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2); a pair lives in a 64-bit register
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float f_min3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// squared distances of the trial point to the two neighbours of one pair-record, folded into the minimum
// (same order of operations as the scalar filter: fma(dz, dz, fma(dy, dy, dx * dx)))
__device__ __forceinline__ float lean_pair(const ulonglong2 xy, const unsigned long long zz, unsigned long long TX,
                                           unsigned long long TY, unsigned long long TZ, float r2min) {
  const unsigned long long dx = f2_sub(TX, xy.x), dy = f2_sub(TY, xy.y), dz = f2_sub(TZ, zz);
  const unsigned long long r2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
  float lo, hi;
  f2_unpack(r2, lo, hi);
  return f_min3(r2min, lo, hi);
}

// The fp32 filter over the 27-cell stencil of one trial: nine (x,y) rows; in each, the three z-cells are one
// contiguous staged range [b, e), read as pair-records from the pair-aligned start of cell rz-1.  A lane reads only the
// records its own row needs (predicated loads: the kernel is bound by shared-memory wavefronts, and a row holds 2.7
// particles on average but up to 8 when the lattice planes of a crystal beat against the cell grid), four in
// straight-line code, the rare longer rows in a loop.  The entry before b in an odd-aligned first record and the one
// after e in the last are real particles of the neighbouring cells (true positions: harmless) or the far-away pad.
// (predicated loads in inline PTX: left to the compiler, every conditional record becomes a branch region of its own,
//  which serialises the shared-memory latencies of a row)
__device__ __forceinline__ void lds_pair_if(unsigned long long& x, unsigned long long& y, unsigned long long& z,
                                            uint32_t a_xy, uint32_t a_z, int on) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p ld.shared.v2.u64 {%0, %1}, [%3];\n\t@p ld.shared.u64 %2, [%4];\n\t}"
               : "+l"(x), "+l"(y), "+l"(z) : "r"(a_xy), "r"(a_z), "r"(on));
}

__device__ __forceinline__ float lean_scan(uint32_t a_xy0, uint32_t a_z0, const unsigned short* __restrict__ cp0, int nry, int czs,
                                           float tx, float ty, float tz) {
  const unsigned long long TX = f2_pack(tx, tx), TY = f2_pack(ty, ty), TZ = f2_pack(tz, tz);
  float r2min = 3.0e38f;
  unsigned long long x0 = 0, y0 = 0, z0 = 0, x1 = 0, y1 = 0, z1 = 0;     // (a record not read keeps stale values: its minimum is not taken)
#pragma unroll
  for (int r = 0; r < 9; r++) {
    const unsigned short* cp = cp0 + ((r / 3) * nry + (r % 3)) * czs;
    const int p0 = (int)cp[0] >> 1, e = cp[3];
    const int np = (e + 1 - 2 * p0) >> 1;              // pair-records holding entries below e
    const uint32_t axy = a_xy0 + 16u * (uint32_t)p0, az = a_z0 + 8u * (uint32_t)p0;
#pragma unroll
    for (int s0 = 0; s0 < LEAN_NP; s0++) {               // (two register sets in turn, so that a load can run ahead)
      if (s0 & 1) {
        lds_pair_if(x1, y1, z1, axy + 16u * s0, az + 8u * s0, s0 < np);
        const unsigned long long dx = f2_sub(TX, x1), dy = f2_sub(TY, y1), dz = f2_sub(TZ, z1);
        float lo, hi;
        f2_unpack(f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx))), lo, hi);
        if (s0 < np) r2min = f_min3(r2min, lo, hi);
      } else {
        lds_pair_if(x0, y0, z0, axy + 16u * s0, az + 8u * s0, s0 < np);
        const unsigned long long dx = f2_sub(TX, x0), dy = f2_sub(TY, y0), dz = f2_sub(TZ, z0);
        float lo, hi;
        f2_unpack(f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx))), lo, hi);
        if (s0 < np) r2min = f_min3(r2min, lo, hi);
      }
    }
    if (np > LEAN_NP) {
#pragma unroll 1
      for (int s = LEAN_NP; s < np; s++) {
        lds_pair_if(x0, y0, z0, axy + 16u * s, az + 8u * s, 1);
        const unsigned long long dx = f2_sub(TX, x0), dy = f2_sub(TY, y0), dz = f2_sub(TZ, z0);
        float lo, hi;
        f2_unpack(f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx))), lo, hi);
        r2min = f_min3(r2min, lo, hi);
      }
    }
  }
  return r2min;
}

// ---------------------------------------------------------------------------------------------------
// k_sweep_lean: one CTA per block; with a.fuse > 1 the block phases [a.phase, a.phase + a.fuse) are one launch
// ordered by per-block completion flags (tickets enumerate (phase, block) in phase order; a block waits for its
// neighbouring blocks of earlier phases only -- see DESIGN.md "Fused phases").
// ---------------------------------------------------------------------------------------------------
template <bool LOG>
__global__ void __launch_bounds__(LEAN_THREADS, LEAN_MIN_CTAS)
k_sweep_lean(SweepArgs a, BlockCfg bc, SlabLink sl, const int* __restrict__ xoff, double4* pos, float4* rel,
             const double4* __restrict__ prop, const uint4* __restrict__ trec, const uint4* __restrict__ traw,
             const int* __restrict__ cs, unsigned long long* __restrict__ cnt,
             hsmc_gpu_trial* __restrict__ log, unsigned long long* __restrict__ nlog, long long logcap) {
  const Grid& g = a.g;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ulonglong2* s_xy = reinterpret_cast<ulonglong2*>(smem_raw);                              // [cap/2] {x0,x1},{y0,y1}
  unsigned long long* s_z2 = reinterpret_cast<unsigned long long*>(s_xy + (bc.cap >> 1));   // [cap/2] {z0,z1}
  unsigned short* s_cz = reinterpret_cast<unsigned short*>(s_z2 + (bc.cap >> 1));           // [max_rows][cz_stride]
  unsigned short* s_items = s_cz + bc.max_rows * bc.cz_stride;                              // [nslots][32] trial items: rx 4 | ry 4 | rz 5 | j 3
  unsigned int* s_iacc = reinterpret_cast<unsigned int*>(s_items + bc.nslots * 32);         // [nslots] accepted trials of a chunk, one bit per lane
  unsigned int* s_pacc = s_iacc + bc.nslots;                                                // [cap/32] staged particles whose trial was accepted
  unsigned char* s_cht = reinterpret_cast<unsigned char*>(s_pacc + (bc.cap >> 5));          // [nslots] trials of each chunk
  unsigned char* s_cslot = s_cht + bc.nslots;                                               // [8][LEAN_COL_CHUNKS] chunk slots of each colour
  float* s_xyf = reinterpret_cast<float*>(s_xy);
  float* s_zf = reinterpret_cast<float*>(s_z2);
  __shared__ BlockRow s_row[LEAN_MAX_ROWS];
  __shared__ int s_cnt[LEAN_MAX_ROWS + 1];
  __shared__ int s_done_idx, s_bad, s_bad2, s_nch[8], s_ph, s_nslot, s_next_row;
  __shared__ BlkGeom s_geom;
  __shared__ long long s_t[24];                 // tuning aid (bc.stamps): clock at the stage boundaries, thread 0
#define LEAN_STAMP(i) do { if (bc.stamps && threadIdx.x == 0) s_t[i] = clock64(); } while (0)
  LEAN_STAMP(0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = LEAN_THREADS / 32;
  const int czs = bc.cz_stride;
  const unsigned FULL = 0xffffffffu;

  // ---- which block (fused launches: ticket) -----------------------------------------------------------------
  // (thread 0 alone does the index arithmetic -- a dozen integer divisions -- and shares the geometry)
  if (tid == 0) {
    const int hbz = bc.nbz >> 1, hby = bc.nby >> 1, hbx = bc.nbx >> 1;
    int ph0 = a.phase, bid = blockIdx.x;
    if (a.fuse > 1) {
      const int per = hbx * hby * hbz, t = (int)(atomicAdd(bc.ticket, 1u) - a.ticket_base);
      ph0 = a.phase + t / per;
      bid = t - (t / per) * per;
    }
    const int bz0 = 2 * (bid % hbz) + (ph0 & 1);
    const int by0 = 2 * ((bid / hbz) % hby) + ((ph0 >> 1) & 1);
    const int bx0 = 2 * (bid / (hbz * hby)) + ((ph0 >> 2) & 1);
    s_geom = blk_geom(g, bc, xoff, bx0, by0, bz0);
    s_geom.bxi = bx0; s_geom.byi = by0; s_geom.bzi = bz0;
    s_ph = ph0;
    s_done_idx = (bx0 * bc.nby + by0) * bc.nbz + bz0;
    s_bad = 0; s_bad2 = 0; s_nslot = 0; s_next_row = 0;
  }
  __syncthreads();
  LEAN_STAMP(1);
  const BlkGeom q = s_geom;
  const int ph = s_ph, bxi = q.bxi, byi = q.byi, bzi = q.bzi;
  const int nrows = q.nrows, nry = q.nry, lenz = q.lenz;

  // ---- staging rows (static while cell membership is: no need to wait for the neighbours): one THREAD per row
  //      (a row is some 25 cells: a short serial loop costs 30x fewer issue slots than a warp per row);
  //      row-relative index of the first particle of every cell, row populations ---------------------------------
  const float inv_nry = 1.0f / (float)nry;
  for (int r = tid; r < nrows; r += LEAN_THREADS) {
    const int rx = (int)(((float)r + 0.5f) * inv_nry), ry = r - rx * nry;       // (exact: r < 256, nry <= 16)
    const int* row = cs + blk_global_row(g, q, rx, ry) * g.nz;
    const int gbA = row[q.zs];
    int gbB = 0, geA;
    if (q.zwrap) { gbB = row[0]; geA = row[g.nz]; }
    else geA = row[q.zs + lenz];
    unsigned short* cz = s_cz + r * czs;
    const int nA = min(lenz, g.nz - q.zs);          // cells zi <= nA read piece A
#pragma unroll 4
    for (int zi = 0; zi <= nA; zi++) cz[zi] = (unsigned short)(row[q.zs + zi] - gbA);
    const int cA = geA - gbA;
#pragma unroll 4
    for (int zi = nA + 1; zi <= lenz; zi++) cz[zi] = (unsigned short)(cA + row[q.zs + zi - g.nz] - gbB);
    s_row[r].gbA = gbA; s_row[r].gbB = gbB; s_row[r].cntA = cA;
    const int cntr = q.zwrap ? cA + (row[q.zs + lenz - g.nz] - gbB) : cA;
    s_cnt[r] = cntr;
    // the row's shadow entries towards L2 now; they are staged two barriers later (a neighbour that is still writing
    // some of them writes into L2 as well)
    for (int k = 0; k < cA; k += 8) prefetch_l2(rel + gbA + k);
    for (int k = 0; k < cntr - cA; k += 8) prefetch_l2(rel + gbB + k);
  }
  for (int i = tid; i < bc.nslots; i += LEAN_THREADS) s_cht[i] = 0;
  // ---- fused launches: wait for the neighbouring blocks of earlier phases -----------------------------------------
  if (a.fuse > 1) {
    if (tid < 27 && tid != 13) {
      int nx = bxi + tid / 9 - 1, ny = byi + (tid / 3) % 3 - 1, nz = bzi + tid % 3 - 1;
      bool have = true;
      if (g.wrap_x) { if (nx < 0) nx += bc.nbx; else if (nx >= bc.nbx) nx -= bc.nbx; }
      else have = nx >= 0 && nx < bc.nbx;          // slab edge: that neighbour lives on another rank (halo exchange)
      if (ny < 0) ny += bc.nby; else if (ny >= bc.nby) ny -= bc.nby;
      if (nz < 0) nz += bc.nbz; else if (nz >= bc.nbz) nz -= bc.nbz;
      const int qp = ((nx & 1) << 2) | ((ny & 1) << 1) | (nz & 1);
      if (have && qp >= a.phase && qp < ph) {
        const unsigned int* f = bc.done + ((size_t)nx * bc.nby + ny) * bc.nbz + nz;
        // (bounded: a protocol error must surface as a CUDA error, not as a GPU that spins for ever)
        unsigned long long t0 = 0;
        unsigned int spins = 0;
        while (ld_relaxed_gpu(f) != a.epoch) {
          __nanosleep(64);
          if ((++spins & 4095u) == 0) {
            const unsigned long long t = hsmc_globaltimer_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 60ull * 1000000000ull) __trap();
          }
        }
      }
      __threadfence();                     // acquire: everything those blocks wrote is visible from here on
    }
    // slab with in-kernel ghost delivery: the last block column (odd, phases 4-7) is the only reader of the right ghost
    // layer; the right neighbour's first-column blocks behind the nine (y,z) blocks around this one must have stored
    // their cells of this sweep (SlabLink)
    if (sl.my_done && bxi == bc.nbx - 1 && tid >= 32 && tid < 41) {
      const int k = tid - 32;
      int ny = byi + k / 3 - 1, nz = bzi + k % 3 - 1;
      if (ny < 0) ny += bc.nby; else if (ny >= bc.nby) ny -= bc.nby;
      if (nz < 0) nz += bc.nbz; else if (nz >= bc.nbz) nz -= bc.nbz;
      const volatile unsigned int* f = sl.my_done + ny * bc.nbz + nz;
      unsigned long long t0 = 0;
      unsigned int spins = 0;
      while ((int)(*f - sl.seq) < 0) {
        __nanosleep(100);
        if ((++spins & 4095u) == 0) {
          const unsigned long long t = hsmc_globaltimer_ns();
          if (t0 == 0) t0 = t;
          else if (t - t0 > 60ull * 1000000000ull) __trap();
        }
      }
      __threadfence_system();
    }
  }
  __syncthreads();
  LEAN_STAMP(2);
  if (tid < 32) {                               // exclusive scan of the row populations
    int carry = 0;
    for (int base = 0; base < nrows; base += 32) {
      const int r = base + tid;
      const int v = (r < nrows) ? s_cnt[r] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (tid >= o) inc += t;
      }
      if (r < nrows) s_row[r].off = carry + inc - v;
      carry += __shfl_sync(FULL, inc, 31);
    }
    if (tid == 0) {
      s_cnt[nrows] = carry;
      if (carry + LEAN_PAD > bc.cap) s_bad = 1;
    }
  }
  const int parx = (g.gx0 + q.x0) & 1, pary = q.y0 & 1, parz = q.z0 & 1;   // parity of region cell (0,0,0); grids are even
  if (!bc.force_global) {
  // ---- chunks (same barrier interval as the row scan: they only need the row-relative cell indices): per colour, the interior cells of the colour (x slowest, z fastest) are cut into runs of at most 32
  //      cells holding at most 32 trials; a run is one warp's work between two colour barriers (lane = trial, a
  //      cell's trials in adjacent lanes).  One warp per colour lists the trials of its chunks, once per block.
  for (int col = warp; col < 8; col += NW) {
    const int fx = 1 + ((parx + 1 + (col >> 2)) & 1), fy = 1 + ((pary + 1 + (col >> 1)) & 1), fz = 1 + ((parz + 1 + col) & 1);
    const int nxc = (q.ex - fx + 2) >> 1, nyc = (q.ey - fy + 2) >> 1, nzc = (q.ez - fz + 2) >> 1;
    const int ncell = nxc * nyc * nzc;
    const float inz = 1.0f / (float)max(nzc, 1), iny = 1.0f / (float)max(nyc, 1);
    const int tgt = 32;
    int at = 0, k = 0;
    while (at < ncell) {
      // a free chunk slot (any 32 consecutive trial slots of the block's table)
      int slot = 0;
      if (lane == 0) slot = atomicAdd(&s_nslot, 1);
      slot = __shfl_sync(FULL, slot, 0);
      if (slot >= bc.nslots || k >= LEAN_COL_CHUNKS) {          // table full: global-memory path
        if (lane == 0) atomicOr(&s_bad2, 4);
        break;
      }
      const int cq = at + lane;
      int n = 0;
      unsigned int cell = 0;
      if (cq < ncell) {
        const int t2 = (int)(((float)cq + 0.5f) * inz), izc = cq - t2 * nzc;      // (exact: cq < 4096, divisors <= 16)
        const int ixc = (int)(((float)t2 + 0.5f) * iny), iyc = t2 - ixc * nyc;
        const int rx = fx + 2 * ixc, ry = fy + 2 * iyc, rz = fz + 2 * izc;
        const unsigned short* cz = s_cz + (rx * nry + ry) * czs + rz;
        n = (int)cz[1] - (int)cz[0];
        cell = ((unsigned)rx << 12) | ((unsigned)ry << 8) | ((unsigned)rz << 3);
        if (n > LEAN_MAX_OCC) { n = 0; atomicOr(&s_bad2, 2); }        // unusually full cell: the block takes the global-memory path
      }
      int inc = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
      }
      const bool in = cq < ncell && inc <= tgt;
      const unsigned m = __ballot_sync(FULL, in);
      unsigned short* items = s_items + slot * 32;
      if (in)
        for (int j = 0; j < n; j++) items[inc - n + j] = (unsigned short)(cell | (unsigned)j);
      const int ncl = __popc(m);                 // cells of this chunk (>= 1: a cell holds at most 8 trials)
      if (lane == ncl - 1) s_cht[slot] = (unsigned char)inc;
      if (lane == 0) s_cslot[col * LEAN_COL_CHUNKS + k] = (unsigned char)slot;
      at += ncl;
      k++;
    }
    if (lane == 0) s_nch[col] = k;
  }
  }
  __syncthreads();
  LEAN_STAMP(3);
  const int total = s_cnt[nrows];

  int n_acc = 0, n_ov = 0, n_cell = 0;
  const float wxf = (float)g.wx, wyf = (float)g.wy, wzf = (float)g.wz;
  const float hxr = 0.5f * (float)q.nrx, hyr = 0.5f * (float)nry, hzr = 0.5f * (float)lenz;
  if (!s_bad && !bc.force_global) {
    // ---- staged indices of the cells; trial records of the interior towards L2 ----
    {
      for (int r = tid; r < nrows; r += LEAN_THREADS) {
        const int rx = (int)(((float)r + 0.5f) * inv_nry), ry = r - rx * nry;
        const BlockRow rw = s_row[r];
        unsigned short* cz = s_cz + r * czs;
        const bool rint = rx >= 1 && rx <= q.ex && ry >= 1 && ry <= q.ey;
        int vb = rw.off, v = (int)cz[1] + rw.off, vn = (int)cz[2] + rw.off;      // cells zi-1, zi, zi+1 (zi = 1)
        cz[0] = (unsigned short)vb;
        for (int zi = 1; zi <= q.ez; zi++) {
          const int ve = (int)cz[zi + 2] + rw.off;                               // start of cell zi+2 = end of cell zi+1
          cz[zi] = (unsigned short)v;
          vb = v; v = vn; vn = ve;
        }
        cz[q.ez + 1] = (unsigned short)v;
        cz[q.ez + 2] = (unsigned short)vn;
        if (rint) {
          // interior particles of the row: slots [first, first + m) of the (trial-ordered) tables
          const int i1 = (int)cz[1] - rw.off, m = (int)cz[q.ez + 1] - (int)cz[1];
          const int first = (i1 < rw.cntA) ? rw.gbA + i1 : rw.gbB + i1 - rw.cntA;
          for (int k = 0; k < m; k += 8) prefetch_l2(trec + first + k);      // (the proposals are not prefetched: measured, no gain)
        }
      }
    }
    LEAN_STAMP(14);
    // ---- stage the fp32 shadow as block-relative coordinates fma(cell index - centre, edge, offset) -----------------
    {
      // one WARP per staged row, lane = particle of the row (a row holds ~26 particles): coalesced 16-byte loads of the
      // shadow (it carries the z cell of its particle), already on their way to L2 since the row pass.  (Measured: more
      // rows in flight per warp, rows handed out on request, or a thread per row piece are all slower.)
#pragma unroll 1
      for (int r = warp; r < nrows; r += NW) {
        const BlockRow rw = s_row[r];
        const int cntr = s_cnt[r];
        const int rx = (int)(((float)r + 0.5f) * inv_nry), ry = r - rx * nry;
        const float cxw = (float)rx - hxr, cyw = (float)ry - hyr;
        for (int k = lane; k < cntr; k += 32) {
          const float4 w = __ldcg(rel + ((k < rw.cntA) ? rw.gbA + k : rw.gbB + (k - rw.cntA)));   // L2: neighbours' blocks wrote these earlier in this launch
          const int i = rw.off + k;
          int zi = __float_as_int(w.w) - q.z0;                    // z cell of the particle inside the region
          if (zi < 0) zi += g.nz; else if (zi >= g.nz) zi -= g.nz;
          float* xy = s_xyf + ((i >> 1) << 2) + (i & 1);
          xy[0] = __fmaf_rn(cxw, wxf, w.x);
          xy[2] = __fmaf_rn(cyw, wyf, w.y);
          s_zf[i] = __fmaf_rn((float)zi - hzr, wzf, w.z);
        }
      }
      LEAN_STAMP(15);
      if (tid < LEAN_PAD) {
        const int i = total + tid;
        float* xy = s_xyf + ((i >> 1) << 2) + (i & 1);
        xy[0] = LEAN_FAR; xy[2] = LEAN_FAR; s_zf[i] = LEAN_FAR;
      }
      for (int i = tid; i < (bc.cap >> 5); i += LEAN_THREADS) s_pacc[i] = 0u;
      for (int i = tid; i < bc.nslots; i += LEAN_THREADS) s_iacc[i] = 0u;
    }
    __syncthreads();
  }
  if (!s_bad && !s_bad2 && !bc.force_global) {
    // ---- trials -----------------------------------------------------------------------------------------------
    const float lo = 1.0f - a.eps, hi = 1.0f + a.eps;
    LEAN_STAMP(4);
    if (s_bad2) goto global_path;                // (block-uniform)
#pragma unroll 1
    for (int col = 0; col < 8; col++) {
      const int nch = s_nch[col];
#pragma unroll 1
      for (int kc = warp; kc < nch; kc += NW) {
        const int slot = s_cslot[col * LEAN_COL_CHUNKS + kc];
        const int T = s_cht[slot];
        const bool valid = lane < T;
        const unsigned int item = valid ? (unsigned int)s_items[slot * 32 + lane] : 0x1108u;    // (padding lanes: cell (1,1,1))
        const int rxc = (item >> 12) & 15, ryc = (item >> 8) & 15, rz = (item >> 3) & 31, j = item & 7;
        const unsigned short* czc = s_cz + (rxc * nry + ryc) * czs + rz;
        const int cob = czc[0], n1 = (int)czc[1] - cob - 1;
        // mates of the same cell: all in this chunk, in adjacent lanes
        const int nprev = valid ? j : 0;
        const int nnext = valid ? n1 - j : 0;
        // trial record: slot (cell start + j) of the trial-ordered table
        const BlockRow rwc = s_row[rxc * nry + ryc];
        const int ro = cob - rwc.off;
        const int gcell0 = (ro < rwc.cntA) ? rwc.gbA + ro : rwc.gbB + ro - rwc.cntA;    // global slot of the cell's first particle
        uint4 rec = make_uint4(0u, 0u, 0u, 0u);
        if (valid) rec = __ldg(trec + gcell0 + j);
        const bool act = valid && (rec.w & TREC_ACT);
        const int koff = rec.w & 15;
        const int sel = cob + koff;
        // trial point, block-relative
        const float tx = __fmaf_rn((float)rxc - hxr, wxf, __uint_as_float(rec.x));
        const float ty = __fmaf_rn((float)ryc - hyr, wyf, __uint_as_float(rec.y));
        const float tz = __fmaf_rn((float)rz - hzr, wzf, __uint_as_float(rec.z));
        // every trial particle of the chunk is hidden while the chunk is scanned
        float* mxy = s_xyf + ((sel >> 1) << 2) + (sel & 1);
        float kx = 0.f, ky = 0.f, kz = 0.f;
        if (valid) {
          kx = mxy[0]; ky = mxy[2]; kz = s_zf[sel];
          mxy[0] = LEAN_FAR; mxy[2] = LEAN_FAR; s_zf[sel] = LEAN_FAR;
        }
        __syncwarp();
        float r2min = 3.0e38f;
        if (act) {
          const unsigned short* cp0 = s_cz + ((rxc - 1) * nry + ryc - 1) * czs + rz - 1;
          r2min = lean_scan(smem_u32(s_xy), smem_u32(s_z2), cp0, nry, czs, tx, ty, tz);
        }
        // mates with a LATER trial in this chunk: still at their old positions
        const int maxnext = __reduce_max_sync(FULL, nnext);
        for (int s2 = 1; s2 <= maxnext; s2++) {
          const float qx = __shfl_down_sync(FULL, kx, s2), qy = __shfl_down_sync(FULL, ky, s2), qz = __shfl_down_sync(FULL, kz, s2);
          if (act && s2 <= nnext) {
            const float ddx = tx - qx, ddy = ty - qy, ddz = tz - qz;
            r2min = fminf(r2min, __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, ddx * ddx)));
          }
        }
        // verdicts in trial order: round s decides the trials with s earlier mates in the chunk; the later
        // trials of those cells then see where the decided particle ended up.  An accepted move updates the
        // staged coordinates now; the master table follows at the end of the block.
        const int maxprev = __reduce_max_sync(FULL, nprev);
        float fx2 = kx, fy2 = ky, fz2 = kz;           // where this lane's particle is after its own trial
        int verdict = 2;
        bool accepted = false;
#pragma unroll 1
        for (int s2 = 0; s2 <= maxprev; s2++) {
          if (valid && nprev == s2) {
            if (act) {
              bool ov = r2min < lo;
              if (!ov && r2min <= hi) {
                const double4 pr = prop[gcell0 + koff];
                ov = block_exact_rescan(pos, prop, s_row, s_cz, s_pacc, czs, nry, rxc, ryc, rz, sel, pr.x, pr.y, pr.z, a.box);
              }
              if (ov) { verdict = 1; n_ov++; }
              else {
                verdict = 0; n_acc++;
                accepted = true;
                fx2 = tx; fy2 = ty; fz2 = tz;
                atomicOr(&s_pacc[sel >> 5], 1u << (sel & 31));
              }
            } else n_cell++;
            mxy[0] = fx2; mxy[2] = fy2; s_zf[sel] = fz2;
          }
          if (s2 < maxprev) {
            __syncwarp();
            const int src = (nprev > s2) ? lane - (nprev - s2) : lane;
            const float qx = __shfl_sync(FULL, fx2, src), qy = __shfl_sync(FULL, fy2, src), qz = __shfl_sync(FULL, fz2, src);
            if (act && nprev > s2) {
              const float ddx = tx - qx, ddy = ty - qy, ddz = tz - qz;
              r2min = fminf(r2min, __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, ddx * ddx)));
            }
          }
        }
        const unsigned accm = __ballot_sync(FULL, accepted);
        if (lane == 0) s_iacc[slot] = accm;
        if (LOG && valid) {
          const uint4 rw4 = __ldg(traw + gcell0 + j);
          const int iy = q.y0 + ryc, iz = q.z0 + rz;
          const int gxl = g.gx0 + q.x0 + rxc;
          const int gx = (gxl >= g.nx) ? gxl - g.nx : gxl;          // interior cells never wrap inside the block
          const long long gcell = ((long long)gx * g.ny + iy) * g.nz + iz;
          const unsigned long long sl = atomicAdd(nlog, 1ull);
          if ((long long)sl < logcap) {
            hsmc_gpu_trial tr;
            tr.seq = ((unsigned long long)(ph * 8 + col) << 56) | ((unsigned long long)gcell << 8) | (unsigned)j;
            tr.id = (int)rw4.w; tr.verdict = verdict;
            tr.raw[0] = rw4.x; tr.raw[1] = rw4.y; tr.raw[2] = rw4.z; tr.pad = 0;
            log[sl] = tr;
          }
        }
        __syncwarp();
      }
      __syncthreads();                       // colour barrier
      LEAN_STAMP(5 + col);
    }
    // ---- commit: the accepted moves of the block go to the master table and its shadow; two items per thread and round,
    //      the loads of both (trial record, then proposal) issued before either is stored -----------------------------
    {
      const int nidx = min(s_nslot, bc.nslots) * 32;
#pragma unroll 1
      for (int idx = tid; idx < nidx; idx += LEAN_COMMIT_B * LEAN_THREADS) {
        int g0[LEAN_COMMIT_B];
        uint4 rec[LEAN_COMMIT_B];
        double2 pxy[LEAN_COMMIT_B];
        double2 pzw[LEAN_COMMIT_B];
#pragma unroll
        for (int u = 0; u < LEAN_COMMIT_B; u++) {
          const int id2 = idx + u * LEAN_THREADS;
          g0[u] = -1;
          if (id2 < nidx && ((s_iacc[id2 >> 5] >> (id2 & 31)) & 1u)) {
            const unsigned int item = s_items[id2];
            const int rxc = (item >> 12) & 15, ryc = (item >> 8) & 15, rz = (item >> 3) & 31;
            const int cob = s_cz[(rxc * nry + ryc) * czs + rz];
            const BlockRow rwc = s_row[rxc * nry + ryc];
            const int ro = cob - rwc.off;
            g0[u] = (ro < rwc.cntA) ? rwc.gbA + ro : rwc.gbB + ro - rwc.cntA;
            rec[u] = __ldg(trec + g0[u] + (int)(item & 7u));
          }
        }
#pragma unroll
        for (int u = 0; u < LEAN_COMMIT_B; u++) {
          if (g0[u] >= 0) {
            g0[u] += (int)(rec[u].w & 15);
            const double* pp = reinterpret_cast<const double*>(prop + g0[u]);
            pxy[u] = *reinterpret_cast<const double2*>(pp);
            pzw[u] = *reinterpret_cast<const double2*>(pp + 2);
          }
        }
#pragma unroll
        for (int u = 0; u < LEAN_COMMIT_B; u++) {
          if (g0[u] >= 0) {
            double* pd = reinterpret_cast<double*>(pos + g0[u]);
            *reinterpret_cast<double2*>(pd) = pxy[u];
            *reinterpret_cast<double2*>(pd + 2) = pzw[u];          // (the whole 32-byte sector {x,y,z,id}: nothing to read-fill on eviction)
            float* rl = reinterpret_cast<float*>(rel + g0[u]);
            *reinterpret_cast<float2*>(rl) = make_float2(__uint_as_float(rec[u].x), __uint_as_float(rec[u].y));
            rl[2] = __uint_as_float(rec[u].z);
          }
        }
      }
    }
  } else {
global_path:
    // ---- the block does not fit the staged scheme (unusually dense) or ablation: global-memory path,
    //      same order of updates ----------------------------------------------------------------------------------
    const int ncell_b = q.ex * q.ey * q.ez;
#pragma unroll 1
    for (int col = 0; col < 8; col++) {
      for (int c = tid; c < ncell_b; c += LEAN_THREADS) {
        const int qz = c % q.ez, qy = (c / q.ez) % q.ey, qx = c / (q.ez * q.ey);
        const int l = q.xa + qx, iy = q.ya + qy, iz = q.za + qz;
        const int cc = (((g.gx0 + l) & 1) << 2) | ((iy & 1) << 1) | (iz & 1);
        if (cc == col)
          cell_update_global_noinline<LOG>(a, ph * 8 + col, pos, rel, cs, l, iy, iz, 0, 1 << 30, n_acc, n_ov, n_cell,
                                           log, nlog, logcap);
      }
      __syncthreads();
    }
  }

  LEAN_STAMP(13);
  // (tuning aid: how many blocks left the staged path, and why -- hsmc_gpu_debug_counters)
  if (tid == 0 && (s_bad | s_bad2) && !bc.force_global) atomicAdd(&cnt[4 + (s_bad ? 0 : (s_bad2 & 2) ? 1 : 2)], 1ull);
  // ---- counters: warp reduce, then straight to the global counters ------------------------------------------
  n_acc = __reduce_add_sync(FULL, n_acc);
  n_ov = __reduce_add_sync(FULL, n_ov);
  n_cell = __reduce_add_sync(FULL, n_cell);
  if (lane == 0 && (n_acc | n_ov | n_cell)) {
    atomicAdd(&cnt[CNT_TRIALS], (unsigned long long)(n_acc + n_ov + n_cell));
    if (n_acc) atomicAdd(&cnt[CNT_ACC], (unsigned long long)n_acc);
    if (n_ov) atomicAdd(&cnt[CNT_REJ_OVERLAP], (unsigned long long)n_ov);
    if (n_cell) atomicAdd(&cnt[CNT_REJ_CELL], (unsigned long long)n_cell);
  }
  const bool deliver = sl.peer_rel != nullptr && bxi == 0;
  if (deliver) {
    // ---- slab: this block's cells of the first owned layer go to the left neighbour's right ghost layer, slot for
    //      slot (SlabLink); the neighbour has published where that layer starts after its rebuild ------------------
    __syncthreads();                                   // the commit above is complete
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while ((int)(sl.info[1] - sl.rebuild) < 0) {
      __nanosleep(100);
      if ((++spins & 4095u) == 0) {
        const unsigned long long t = hsmc_globaltimer_ns();
        if (t0 == 0) t0 = t;
        else if (t - t0 > 60ull * 1000000000ull) __trap();
      }
    }
    __threadfence_system();
    const long long per = (long long)g.ny * g.nz;
    const int shift = (int)sl.info[0] - cs[(long long)g.own_lo * per];
    double4* ppos = sl.peer_pos[sl.info[2] & 1u];
    for (int y = warp; y < q.ey; y += NW) {
      const long long c0 = (long long)g.own_lo * per + (long long)(q.ya + y) * g.nz + q.za;
      const int e = cs[c0 + q.ez];
      for (int k = cs[c0] + lane; k < e; k += 32) {
        ppos[k + shift] = pos[k];
        sl.peer_rel[k + shift] = rel[k];
      }
    }
  }
  if (a.fuse > 1) {
    // every thread's stores to pos / rel precede the barrier; thread 0 then publishes the block
    __syncthreads();
    if (tid == 0) {
      if (deliver) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int*>(sl.peer_done + byi * bc.nbz + bzi) = sl.seq;
      }
      __threadfence();
      st_release_gpu(bc.done + s_done_idx, a.epoch);
    }
  }
  if (bc.stamps && tid == 0 && !s_bad && !s_bad2) {
    const long long t14 = clock64();
    for (int i = 0; i < 13; i++) atomicAdd(bc.stamps + i, (unsigned long long)(s_t[i + 1] - s_t[i]));
    atomicAdd(bc.stamps + 13, (unsigned long long)(t14 - s_t[13]));
    atomicAdd(bc.stamps + 14, (unsigned long long)(s_t[14] - s_t[3]));       // warp 0: cell index pass + prefetches
    atomicAdd(bc.stamps + 15, (unsigned long long)(s_t[15] - s_t[14]));      // warp 0: its staging rows
    atomicAdd(bc.stamps + 31, 1ull);
  }
}
