// sweep_lean.cuh -- K2, the default sweep: k_block_plan (trial planning + proposals) and k_sweep_lean
// (block-resident fp32x2 stencil filter).  Same two-level checkerboard and the same Markov chain as
// k_sweep_block (sweep_block.cuh, kept as an ablation) and as the all-double global-memory evaluation
// (sweep_impl 5): every particle gets one trial per sweep, each trial is the reference's part_move()
// (moves.c:27-80), a trial that leaves its cell is rejected.
//
// What moved out of the latency-critical kernel, and why (profiles/r01_ncu_k_sweep_block_fused.txt: 128
// registers, 16 warps/SM, 38 % issue, 55 % shared-memory bank conflicts, prologue 31 % of warp time):
//
//  * A trial point depends on nothing but the particle's own position at the start of the sweep (a particle
//    only ever moves by its own trial) and on Philox(global cell, trial index, sweep).  So all proposals of a
//    sweep are generated up front by k_block_plan, a throughput kernel: new position in double
//    (moves.c:52-57, 215-226) into the idle half of the ping-pong master table, cell test, fp32 shadow of the
//    new position.  k_sweep_lean has no Philox and no double arithmetic on its hot path.
//  * k_block_plan also lays out, per block, everything that is static during a sweep (cell membership is):
//    the staging rows, the staged index of every region cell, and per cell colour the list of trials in
//    (row, z, trial index) order.  The sweep kernel's prologue is three bulk-TMA copies plus one pass that
//    stages the shadow coordinates; its trial slots are read coalesced, 16 B per lane.
//  * Staged coordinates are block-relative fp32, stored as PAIRS {x0,x1,y0,y1} + {z0,z1}, so the stencil
//    filter runs on packed fp32x2 instructions (FADD2 / FMUL2 / FFMA2 + FMNMX3: 7 math instructions and two
//    loads per two neighbours instead of 14 + 2).  Consecutive lanes of a warp hold consecutive trials along
//    z of one (x,y) row: their stencil rows start at monotonically increasing shared-memory addresses about
//    one pair-record apart, which is (nearly) conflict-free for the 16-byte and the 8-byte loads alike.
//
// Exactness is unchanged: the fp32 minimum r^2 is a filter with a rigorous error bound eps
// (setup_blocks); min < 1 - eps is a certain overlap, min > 1 + eps a certain miss, anything in between is
// re-evaluated from the master table with the reference's double arithmetic (moves.c:400-431).  Trials of
// one cell that fall into the same 32-lane chunk are ordered with warp shuffles: a later trial sees its
// earlier mates at the positions their own trials left them in.
#pragma once

#ifndef LEAN_THREADS
#define LEAN_THREADS 192
#endif
#ifndef LEAN_MIN_CTAS
#define LEAN_MIN_CTAS 5
#endif
#define LEAN_NP_MIN 3           // pair-records (2 entries each) scanned per stencil row in straight-line code:
#define LEAN_NP_MAX 5           //   chosen per block by k_block_plan from its longest stencil row; beyond MAX: deep loop
#define LEAN_PAD 10             // far-away entries after the last staged particle (covers the over-scan)
#define LEAN_FAR 1.0e15f
#define LEAN_MAX_OCC 8          // most particles per cell (3-bit trial index)
#define LEAN_MAX_ROWS 256       // (mbx+2)*(mby+2)
#define PLAN_THREADS 256

// LeanPlan and the plan header layout: see hsmc_gpu.cu

// trial code: sel (staged index) 12 | rx 4 | ry 4 | rz 5 | j 3 | n-1 3 | act 1
#define LEAN_INVALID 0xffffffffu     // (rx = 15 never occurs: at most 16 staged rows along x, the last one halo)
#define LEAN_CODE(sel, rx, ry, rz, j, n1) ((unsigned)(sel) | ((unsigned)(rx) << 12) | ((unsigned)(ry) << 16) | ((unsigned)(rz) << 20) | ((unsigned)(j) << 25) | ((unsigned)(n1) << 28))

struct BlkGeom {
  int xa, ya, za;            // first interior cell (local x layer, y, z)
  int ex, ey, ez;            // interior extent
  int x0, y0, z0;            // region origin (may be -1: periodic wrap)
  int nrx, nry, lenz, nrows;
  int zs;                    // wrapped z of the region's first cell
  bool zwrap;
};

__device__ __forceinline__ BlkGeom blk_geom(const Grid& g, const BlockCfg& bc, const int* __restrict__ xoff, int bxi,
                                            int byi, int bzi) {
  BlkGeom q;
  q.xa = xoff[bxi];
  const int xb = xoff[bxi + 1];
  // (32-bit: block index x cells per axis stays far below 2^31; same values as the 64-bit form)
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

// global (x,y) row index of staged row r
__device__ __forceinline__ long long blk_global_row(const Grid& g, const BlkGeom& q, int r) {
  const int rx = r / q.nry, ry = r - rx * q.nry;
  int lx = q.x0 + rx;
  if (g.wrap_x) { if (lx < 0) lx += g.nlx; else if (lx >= g.nlx) lx -= g.nlx; }
  int y = q.y0 + ry;
  if (y < 0) y += g.ny; else if (y >= g.ny) y -= g.ny;
  return (long long)lx * g.ny + y;
}

// ---------------------------------------------------------------------------------------------------
// k_block_plan: one CTA per block (all eight block phases at once; nothing here depends on the order
// of the updates).  Writes the block's plan and generates the proposals of its interior particles.
// ---------------------------------------------------------------------------------------------------
template <bool LOG>
__global__ void __launch_bounds__(PLAN_THREADS)
k_block_plan(SweepArgs a, BlockCfg bc, LeanPlan pl, const int* __restrict__ xoff, const double4* __restrict__ pos,
             const int* __restrict__ cs, double4* __restrict__ prop) {
  const Grid& g = a.g;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned short* s_cz = reinterpret_cast<unsigned short*>(smem_raw);                          // [max_rows][cz_stride]
  unsigned int* s_desc = reinterpret_cast<unsigned int*>(s_cz + bc.max_rows * bc.cz_stride);   // [desc_cap]
  unsigned short* s_pre = reinterpret_cast<unsigned short*>(s_desc + bc.desc_cap);             // [interior rows][32]
  __shared__ BlockRow s_row[LEAN_MAX_ROWS];
  __shared__ int s_cnt[LEAN_MAX_ROWS + 1];
  __shared__ int s_grow[LEAN_MAX_ROWS];
  __shared__ int s_rc[2 * LEAN_MAX_ROWS];      // per interior row and z parity (a "run"): trials, then offset in its colour's list
  __shared__ int s_ntr[8], s_cbase[9], s_flags, s_need, s_tbase;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWP = PLAN_THREADS / 32;
  const int czs = bc.cz_stride;
  const unsigned FULL = 0xffffffffu;

  const int bl = blockIdx.x;
  const int bzi = bl % bc.nbz, byi = (bl / bc.nbz) % bc.nby, bxi = bl / (bc.nbz * bc.nby);
  const BlkGeom q = blk_geom(g, bc, xoff, bxi, byi, bzi);
  const int nrows = q.nrows, nry = q.nry, lenz = q.lenz;
  if (tid == 0) { s_flags = 0; s_need = 0; }

  // ---- staging rows: one or two contiguous slot ranges of the cell-ordered table each ----------------
  for (int r = tid; r < nrows; r += PLAN_THREADS) {
    const long long grow = blk_global_row(g, q, r);
    const long long rbase = grow * g.nz;
    const int gbA = cs[rbase + q.zs], geA = cs[rbase + min(q.zs + lenz, g.nz)];
    int gbB = 0, geB = 0;
    if (q.zwrap) { gbB = cs[rbase]; geB = cs[rbase + (q.zs + lenz - g.nz)]; }
    s_grow[r] = (int)grow;
    s_row[r].gbA = gbA; s_row[r].gbB = gbB; s_row[r].cntA = geA - gbA;
    s_cnt[r] = (geA - gbA) + (geB - gbB);
  }
  __syncthreads();
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
      if (carry + LEAN_PAD > bc.cap) atomicOr(&s_flags, PLAN_BAD);
    }
  }
  __syncthreads();
  const int total = s_cnt[nrows];

  // ---- one warp per staged row, one lane per cell of the row (at most 32): staged index of the first particle
  //      of every cell; longest stencil row; for interior rows the trials of each z parity ("run") and every
  //      cell's offset inside its run ----------------------------------------------------------------------
  const int parx = (g.gx0 + q.x0) & 1, pary = q.y0 & 1, parz = q.z0 & 1;   // parity of region cell (0,0,0); grids are even
  unsigned short* gcz = pl.cz + (size_t)bl * bc.max_rows * czs;
  for (int r = warp; r < nrows; r += NWP) {
    const int rx = r / nry, ry = r - rx * nry;
    const long long rbase = (long long)s_grow[r] * g.nz;
    const BlockRow rw = s_row[r];
    const int zi = lane, z = q.zs + zi;
    int v = 0;
    if (zi <= lenz) {
      v = (z <= g.nz) ? rw.off + (cs[rbase + z] - rw.gbA) : rw.off + rw.cntA + (cs[rbase + (z - g.nz)] - rw.gbB);
      s_cz[r * czs + zi] = (unsigned short)v;
      gcz[r * czs + zi] = (unsigned short)v;
    }
    const int vn = __shfl_down_sync(FULL, v, 1), vb = __shfl_up_sync(FULL, v, 1), ve = __shfl_down_sync(FULL, v, 2);
    const bool zint = zi >= 1 && zi <= q.ez;
    const int n = zint ? vn - v : 0;
    // entries a trial in cell zi scans in this row, from the pair-aligned start of cell zi-1 to the end of cell zi+1
    const int need = __reduce_max_sync(FULL, zint ? ve - (vb & ~1) : 0);
    if (lane == 0) atomicMax(&s_need, need);
    if (rx >= 1 && rx <= q.ex && ry >= 1 && ry <= q.ey) {
      const int ri = (rx - 1) * q.ey + (ry - 1);
      if (__any_sync(FULL, n > LEAN_MAX_OCC) && lane == 0) atomicOr(&s_flags, PLAN_BAD);
      const int par = (parz + zi) & 1;
      int inc = par ? (n << 16) : n;                 // both parities in one scan
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
      }
      const int tot = __shfl_sync(FULL, inc, 31);
      const int pre = par ? (inc >> 16) - n : (inc & 0xffff) - n;
      if (zint) s_pre[ri * 32 + zi] = (unsigned short)pre;
      if (lane == 0) { s_rc[2 * ri] = tot & 0xffff; s_rc[2 * ri + 1] = tot >> 16; }
    }
  }
  __syncthreads();
  // ---- per colour (one thread each): place the runs, rows in (rx, ry) order, in the colour's list.  A colour's
  //      trials are split evenly over its chunks (each chunk is one warp's work between two colour barriers), and
  //      the trials of one cell never straddle a chunk (chunks of a colour run concurrently, the trials of a cell
  //      are ordered): a cell that would is moved to the start of the next chunk; skipped slots stay invalid.
  if (tid < 8) {
    const int p = tid & 1;
    int sum = 0;
    for (int rx = 1, ri = 0; rx <= q.ex; rx++)
      for (int ry = 1; ry <= q.ey; ry++, ri++)
        if (((((parx + rx) & 1) << 2) | (((pary + ry) & 1) << 1)) == (tid & 6)) sum += s_rc[2 * ri + p];
    const int nch = max(1, (sum + 31) >> 5);
    const int ccap = min(32, (sum + nch - 1) / nch + 2);        // chunk capacity (+2: room for the moved cells)
    int at = 0;
    for (int rx = 1, ri = 0; rx <= q.ex; rx++)
      for (int ry = 1; ry <= q.ey; ry++, ri++) {
        if (((((parx + rx) & 1) << 2) | (((pary + ry) & 1) << 1)) != (tid & 6)) continue;
        const int len = s_rc[2 * ri + p];
        if ((at & 31) >= ccap) at = (at + 31) & ~31;
        s_rc[2 * ri + p] = at;
        if ((at & 31) + len <= ccap) { at += len; continue; }
        // the run crosses a chunk boundary: walk its cells
        const unsigned short* cz = s_cz + (rx * nry + ry) * czs;
        unsigned short* pre = s_pre + ri * 32;
        const int start = at;
        for (int zi = 1 + ((parz + 1 + p) & 1); zi <= q.ez; zi += 2) {
          const int n = (int)cz[zi + 1] - (int)cz[zi];
          if (n == 0) continue;
          if ((at & 31) + n > ccap) at = (at + 31) & ~31;
          pre[zi] = (unsigned short)(at - start);
          at += n;
        }
      }
    s_ntr[tid] = at;
  }
  __syncthreads();
  if (tid == 0) {
    int base = 0;
    for (int c = 0; c < 8; c++) { s_cbase[c] = base; base += (s_ntr[c] + 31) & ~31; }
    s_cbase[8] = base;
    if (base > bc.desc_cap) atomicOr(&s_flags, PLAN_BAD);
    unsigned int tb = 0;
    if (!(s_flags & PLAN_BAD)) {
      tb = atomicAdd(pl.cursor, (unsigned int)base);
      if ((long long)tb + base > pl.cap_trials) atomicOr(&s_flags, PLAN_BAD);
    }
    s_tbase = (int)tb;
    // pair-records the straight-line scan must cover; beyond LEAN_NP_MAX the sweep runs its deep loop too
    if (s_need > 2 * LEAN_NP_MAX) atomicOr(&s_flags, PLAN_DEEP);
  }
  __syncthreads();
  const int flags = s_flags;
  // ---- header and rows go to global memory (the cell indices went out above) ---------------------------------
  {
    int* hdr = pl.hdr + (size_t)bl * PLAN_HDR_INTS;
    if (tid < PLAN_HDR_INTS) {
      int v = 0;
      if (tid == PLAN_TOTAL) v = total;
      else if (tid >= PLAN_NTR && tid < PLAN_NTR + 8) v = s_ntr[tid - PLAN_NTR];
      else if (tid == PLAN_FLAGS) v = flags;
      else if (tid == PLAN_TBASE) v = s_tbase;
      else if (tid == PLAN_NPAIRS) v = min(LEAN_NP_MAX, max(LEAN_NP_MIN, (s_need + 1) >> 1));
      hdr[tid] = v;
    }
    int4* grow = reinterpret_cast<int4*>(pl.row + (size_t)bl * bc.max_rows);
    for (int r = tid; r < nrows; r += PLAN_THREADS) grow[r] = reinterpret_cast<const int4*>(s_row)[r];
  }
  if (flags & PLAN_BAD) return;              // the sweep runs this block from global memory (no proposals needed)

  // ---- trial descriptors ----------------------------------------------------------------------------------
  const int nslots = s_cbase[8];
  for (int u = tid; u < nslots; u += PLAN_THREADS) s_desc[u] = LEAN_INVALID;
  __syncthreads();
  for (int ri = warp; ri < q.ex * q.ey; ri += NWP) {
    const int rx = ri / q.ey + 1, ry = ri - (rx - 1) * q.ey + 1;
    const int zi = lane;
    if (zi >= 1 && zi <= q.ez) {
      const int p = (parz + zi) & 1;
      const int col = (((parx + rx) & 1) << 2) | (((pary + ry) & 1) << 1) | p;
      const unsigned short* cz = s_cz + (rx * nry + ry) * czs;
      const int b = cz[zi], n = (int)cz[zi + 1] - b;
      unsigned int* out = s_desc + s_cbase[col] + s_rc[2 * ri + p] + s_pre[ri * 32 + zi];
      for (int j = 0; j < n; j++) out[j] = LEAN_CODE(b, rx, ry, zi, j, n - 1);
    }
  }
  __syncthreads();

  // ---- proposals: the reference's trial point (moves.c:52-57), apply_pbc (moves.c:215-226), cell test ---------
  const long long tbase = s_tbase;
#pragma unroll 2
  for (int u = tid; u < nslots; u += PLAN_THREADS) {
    const unsigned int d = s_desc[u];
    if (d == LEAN_INVALID) { pl.trial[tbase + u] = make_uint4(0u, 0u, 0u, LEAN_INVALID); continue; }
    const int ob = d & 0xfff, rx = (d >> 12) & 15, ry = (d >> 16) & 15, rz = (d >> 20) & 31;
    const int j = (d >> 25) & 7, n1 = (d >> 28) & 7;
    const BlockRow rw = s_row[rx * nry + ry];
    const int o = ob - rw.off;
    const int gs0 = (o < rw.cntA) ? rw.gbA + o : rw.gbB + o - rw.cntA;
    // the particle with exactly j smaller ids in its cell (cells rarely hold more than three)
    int k = 0;
    if (n1 > 0) {
      for (int m = 0; m <= n1; m++) {
        const double idm = pos[gs0 + m].w;
        int c2 = 0;
        for (int m2 = 0; m2 <= n1; m2++) c2 += pos[gs0 + m2].w < idm;
        if (c2 == j) k = m;
      }
    }
    const int gs = gs0 + k;
    const double4 p = pos[gs];
    const int iy = q.y0 + ry, iz = q.z0 + rz;
    const int gxl = g.gx0 + q.x0 + rx;
    const int gx = (gxl >= g.nx) ? gxl - g.nx : gxl;          // interior cells never wrap inside the block
    const long long gcell = ((long long)gx * g.ny + iy) * g.nz + iz;
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
    if (act) nrel = make_rel(g, gx, iy, iz, xn, yn, zn, p.w);
    const unsigned int code = LEAN_CODE(ob + k, rx, ry, rz, j, n1) | (act ? 0x80000000u : 0u);
    pl.trial[tbase + u] = make_uint4(__float_as_uint(nrel.x), __float_as_uint(nrel.y), __float_as_uint(nrel.z), code);
    prop[gs] = make_double4(xn, yn, zn, p.w);
    if (LOG) pl.raw[tbase + u] = make_uint4(rn.v[0], rn.v[1], rn.v[2], (unsigned int)(int)p.w);
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

// the fp32 filter over the 27-cell stencil of one trial: nine (x,y) rows, NP pair-records each, read at fixed offsets
// from the pair-aligned start of cell rz-1, unmasked (entries past the row's three cells are real particles farther
// along -- true positions, so harmless -- or the far-away pad after the last staged particle)
template <int NP>
__device__ __forceinline__ float lean_scan(const ulonglong2* __restrict__ s_xy, const unsigned long long* __restrict__ s_z2,
                                           const unsigned short* __restrict__ cp0, int nry, int czs, float tx, float ty,
                                           float tz) {
  const unsigned long long TX = f2_pack(tx, tx), TY = f2_pack(ty, ty), TZ = f2_pack(tz, tz);
  float r2min = 3.0e38f;
#pragma unroll
  for (int r = 0; r < 9; r++) {
    const int p0 = (int)cp0[((r / 3) * nry + (r % 3)) * czs] >> 1;
    const ulonglong2* pxy = s_xy + p0;
    const unsigned long long* pz = s_z2 + p0;
#pragma unroll
    for (int s = 0; s < NP; s++) r2min = lean_pair(pxy[s], pz[s], TX, TY, TZ, r2min);
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
k_sweep_lean(SweepArgs a, BlockCfg bc, LeanPlan pl, const int* __restrict__ xoff, double4* pos, float4* rel,
             const double4* __restrict__ prop, const int* __restrict__ cs, unsigned long long* __restrict__ cnt,
             hsmc_gpu_trial* __restrict__ log, unsigned long long* __restrict__ nlog, long long logcap) {
  const Grid& g = a.g;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ulonglong2* s_xy = reinterpret_cast<ulonglong2*>(smem_raw);                              // [cap/2] {x0,x1},{y0,y1}
  unsigned long long* s_z2 = reinterpret_cast<unsigned long long*>(s_xy + (bc.cap >> 1));   // [cap/2] {z0,z1}
  unsigned short* s_cz = reinterpret_cast<unsigned short*>(s_z2 + (bc.cap >> 1));           // [max_rows][cz_stride]
  BlockRow* s_row = reinterpret_cast<BlockRow*>(s_cz + bc.max_rows * bc.cz_stride);         // [max_rows]
  int* s_hdr = reinterpret_cast<int*>(s_row + bc.max_rows);                                 // [PLAN_HDR_INTS]
  float* s_xyf = reinterpret_cast<float*>(s_xy);
  float* s_zf = reinterpret_cast<float*>(s_z2);
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_tick, s_done_idx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = LEAN_THREADS / 32;
  const int czs = bc.cz_stride;

  // ---- which block (fused launches: ticket) -----------------------------------------------------------------
  const int hbz = bc.nbz >> 1, hby = bc.nby >> 1, hbx = bc.nbx >> 1;
  int ph = a.phase, bid = blockIdx.x;
  if (tid == 0) mbar_init(&s_bar, 1);
  if (a.fuse > 1) {
    if (tid == 0) s_tick = (int)(atomicAdd(bc.ticket, 1u) - a.ticket_base);
    __syncthreads();
    const int per = hbx * hby * hbz, t = s_tick;
    ph = a.phase + t / per;
    bid = t - (t / per) * per;
  }
  const int pcx = (ph >> 2) & 1, pcy = (ph >> 1) & 1, pcz = ph & 1;
  const int bzi = 2 * (bid % hbz) + pcz;
  const int byi = 2 * ((bid / hbz) % hby) + pcy;
  const int bxi = 2 * (bid / (hbz * hby)) + pcx;
  const int bl = (bxi * bc.nby + byi) * bc.nbz + bzi;
  // the plan of this block is static data of an earlier kernel: fetch it while waiting for the neighbours
  if (tid == 0) {
    s_done_idx = bl;
    const uint32_t b_hdr = PLAN_HDR_INTS * 4, b_row = (uint32_t)bc.max_rows * 16u, b_cz = (uint32_t)(bc.max_rows * czs) * 2u;
    mbar_arrive_tx(&s_bar, b_hdr + b_row + b_cz);
    tma_bulk_g2s(s_hdr, pl.hdr + (size_t)bl * PLAN_HDR_INTS, b_hdr, &s_bar);
    tma_bulk_g2s(s_row, pl.row + (size_t)bl * bc.max_rows, b_row, &s_bar);
    tma_bulk_g2s(s_cz, pl.cz + (size_t)bl * bc.max_rows * czs, b_cz, &s_bar);
  }
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
  }
  __syncthreads();                          // (also: the mbarrier initialisation is visible to every waiter)
  const BlkGeom q = blk_geom(g, bc, xoff, bxi, byi, bzi);
  const int nrows = q.nrows, nry = q.nry, lenz = q.lenz;
  mbar_wait_bounded(&s_bar, 0);
  const int total = s_hdr[PLAN_TOTAL], pflags = s_hdr[PLAN_FLAGS];
  const long long tbase = s_hdr[PLAN_TBASE];

  int n_acc = 0, n_ov = 0, n_cell = 0;
  if (!(pflags & PLAN_BAD) && !bc.force_global) {
    // ---- stage the fp32 shadow as block-relative coordinates fma(cell index - centre, edge, offset) -----------------
    const float wxf = (float)g.wx, wyf = (float)g.wy, wzf = (float)g.wz;
    const float hxr = 0.5f * (float)q.nrx, hyr = 0.5f * (float)nry, hzr = 0.5f * (float)lenz;
    {
      // each staged row is cut into P runs of particles, one (row, run) item per thread and round; a thread first
      // issues all the loads of a batch, then converts them (the cell of a particle follows from the row's cell index)
      const int P = (2 * nrows <= LEAN_THREADS) ? 2 : 1;
#pragma unroll 1
      for (int idx = tid; idx < P * nrows; idx += LEAN_THREADS) {
        const int r = (P == 2) ? idx >> 1 : idx, piece = (P == 2) ? idx & 1 : 0;
        const BlockRow rw = s_row[r];
        const int cntr = ((r + 1 < nrows) ? s_row[r + 1].off : total) - rw.off;
        const int k0 = (cntr * piece) / P, k1 = (cntr * (piece + 1)) / P;
        const int rx = r / nry, ry = r - rx * nry;
        const float cxw = ((float)rx - hxr), cyw = ((float)ry - hyr);
        const unsigned short* cz = s_cz + r * czs;
        int zi = 0, nextb = cz[1];                       // cell of the particle being converted; start of the next cell
#pragma unroll 1
        for (int kb = k0; kb < k1; kb += 8) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const int k = kb + u;
            if (k < k1) v[u] = __ldcg(rel + ((k < rw.cntA) ? rw.gbA + k : rw.gbB + (k - rw.cntA)));   // L2: neighbours' blocks wrote these earlier in this launch
          }
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const int k = kb + u;
            if (k < k1) {
              const int i = rw.off + k;
              while (i >= nextb && zi < lenz - 1) { zi++; nextb = cz[zi + 1]; }
              float* xy = s_xyf + ((i >> 1) << 2) + (i & 1);
              xy[0] = __fmaf_rn(cxw, wxf, v[u].x);
              xy[2] = __fmaf_rn(cyw, wyf, v[u].y);
              s_zf[i] = __fmaf_rn((float)zi - hzr, wzf, v[u].z);
            }
          }
        }
      }
      if (tid < LEAN_PAD) {
        const int i = total + tid;
        float* xy = s_xyf + ((i >> 1) << 2) + (i & 1);
        xy[0] = LEAN_FAR; xy[2] = LEAN_FAR; s_zf[i] = LEAN_FAR;
      }
    }
    __syncthreads();

    // ---- trials -----------------------------------------------------------------------------------------------
    const unsigned FULL = 0xffffffffu;
    const float lo = 1.0f - a.eps, hi = 1.0f + a.eps;
    const bool deep_rows = (pflags & PLAN_DEEP) != 0;
    const int npairs = s_hdr[PLAN_NPAIRS];
    int cbase = 0;
#pragma unroll 1
    for (int col = 0; col < 8; col++) {
      const int ntr = s_hdr[PLAN_NTR + col];
#pragma unroll 1
      for (int chunk = warp; chunk * 32 < ntr; chunk += NW) {
        const int t = chunk * 32 + lane;
        bool valid = t < ntr;
        uint4 rec = make_uint4(0u, 0u, 0u, LEAN_INVALID);
        if (valid) rec = __ldg(pl.trial + tbase + cbase + t);
        // the next records of this warp (next chunk of this colour, else its first chunk of the next colour): towards L2
        {
          const bool same = (chunk + NW) * 32 < ntr;
          const long long nt = same ? (long long)cbase + t + NW * 32 : (long long)cbase + ((ntr + 31) & ~31) + warp * 32 + lane;
          if (same || col < 7) prefetch_l1(pl.trial + tbase + nt);
        }
        const unsigned int code = rec.w;
        valid = valid && code != LEAN_INVALID;
        const bool act = valid && (code >> 31);
        const int sel = code & 0xfff, rxc = (code >> 12) & 15, ryc = (code >> 16) & 15, rz = (code >> 20) & 31;
        const int j = (code >> 25) & 7, n1 = (code >> 28) & 7;
        // mates: the trials of a cell sit in adjacent lanes of one chunk (k_block_plan never lets a cell straddle)
        const int nprev = valid ? j : 0;
        const int nnext = valid ? n1 - j : 0;
        if (act) {                                 // the proposal an accepted trial copies: towards L1 now
          const BlockRow rwc = s_row[rxc * nry + ryc];
          const int ro = sel - rwc.off;
          prefetch_l1(prop + ((ro < rwc.cntA) ? rwc.gbA + ro : rwc.gbB + ro - rwc.cntA));
        }
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
          if (npairs <= 3) r2min = lean_scan<3>(s_xy, s_z2, cp0, nry, czs, tx, ty, tz);
          else if (npairs == 4) r2min = lean_scan<4>(s_xy, s_z2, cp0, nry, czs, tx, ty, tz);
          else r2min = lean_scan<5>(s_xy, s_z2, cp0, nry, czs, tx, ty, tz);
          if (deep_rows) {                         // some stencil row of this block is longer than the straight-line scan
            const unsigned long long TX = f2_pack(tx, tx), TY = f2_pack(ty, ty), TZ = f2_pack(tz, tz);
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
              const unsigned short* cp = cp0 + ((r / 3) * nry + (r % 3)) * czs;
              const int e = cp[3];
#pragma unroll 1
              for (int p = ((int)cp[0] >> 1) + LEAN_NP_MAX; 2 * p < e; p++) r2min = lean_pair(s_xy[p], s_z2[p], TX, TY, TZ, r2min);
            }
          }
        }
        // mates with a LATER trial in this chunk: still at their old positions
        const int maxnext = __reduce_max_sync(FULL, nnext);
        for (int s = 1; s <= maxnext; s++) {
          const float qx = __shfl_down_sync(FULL, kx, s), qy = __shfl_down_sync(FULL, ky, s), qz = __shfl_down_sync(FULL, kz, s);
          if (act && s <= nnext) {
            const float ddx = tx - qx, ddy = ty - qy, ddz = tz - qz;
            r2min = fminf(r2min, __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, ddx * ddx)));
          }
        }
        // verdicts in trial order: round s decides the trials with s earlier mates in the chunk; the later
        // trials of those cells then see where the decided particle ended up
        const int maxprev = __reduce_max_sync(FULL, nprev);
        float fx = kx, fy = ky, fz = kz;              // where this lane's particle is after its own trial
        int verdict = 2;
#pragma unroll 1
        for (int s = 0; s <= maxprev; s++) {
          if (valid && nprev == s) {
            if (act) {
              const BlockRow rwc = s_row[rxc * nry + ryc];
              const int ro = sel - rwc.off;
              const int gs = (ro < rwc.cntA) ? rwc.gbA + ro : rwc.gbB + ro - rwc.cntA;
              bool ov = r2min < lo;
              if (!ov && r2min <= hi) {
                const double4 pr = prop[gs];
                ov = block_exact_rescan(pos, s_row, s_cz, czs, nry, rxc, ryc, rz, sel, pr.x, pr.y, pr.z, a.box);
              }
              if (ov) { verdict = 1; n_ov++; }
              else {
                verdict = 0; n_acc++;
                fx = tx; fy = ty; fz = tz;
                const double4 pr = prop[gs];
                double* pd = reinterpret_cast<double*>(pos + gs);
                *reinterpret_cast<double2*>(pd) = make_double2(pr.x, pr.y);
                pd[2] = pr.z;
                float* rl = reinterpret_cast<float*>(rel + gs);
                *reinterpret_cast<float2*>(rl) = make_float2(__uint_as_float(rec.x), __uint_as_float(rec.y));
                rl[2] = __uint_as_float(rec.z);
              }
            } else n_cell++;
            mxy[0] = fx; mxy[2] = fy; s_zf[sel] = fz;
          }
          if (s < maxprev) {
            __syncwarp();
            const int src = (nprev > s) ? lane - (nprev - s) : lane;
            const float qx = __shfl_sync(FULL, fx, src), qy = __shfl_sync(FULL, fy, src), qz = __shfl_sync(FULL, fz, src);
            if (act && nprev > s) {
              const float ddx = tx - qx, ddy = ty - qy, ddz = tz - qz;
              r2min = fminf(r2min, __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, ddx * ddx)));
            }
          }
        }
        if (LOG && valid) {
          const uint4 rw4 = __ldg(pl.raw + tbase + cbase + t);
          const int iy = q.y0 + ryc, iz = q.z0 + rz;
          const int gxl = g.gx0 + q.x0 + rxc;
          const int gx = (gxl >= g.nx) ? gxl - g.nx : gxl;
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
      cbase += (ntr + 31) & ~31;
      __syncthreads();                       // colour barrier
    }
  } else {
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

  // ---- counters: warp reduce, then straight to the global counters ------------------------------------------
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  n_ov = __reduce_add_sync(0xffffffffu, n_ov);
  n_cell = __reduce_add_sync(0xffffffffu, n_cell);
  if (lane == 0 && (n_acc | n_ov | n_cell)) {
    atomicAdd(&cnt[CNT_TRIALS], (unsigned long long)(n_acc + n_ov + n_cell));
    if (n_acc) atomicAdd(&cnt[CNT_ACC], (unsigned long long)n_acc);
    if (n_ov) atomicAdd(&cnt[CNT_REJ_OVERLAP], (unsigned long long)n_ov);
    if (n_cell) atomicAdd(&cnt[CNT_REJ_CELL], (unsigned long long)n_cell);
  }
  if (a.fuse > 1) {
    // every thread's stores to pos / rel precede the barrier; thread 0 then publishes the block
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      st_release_gpu(bc.done + s_done_idx, a.epoch);
    }
  }
}
