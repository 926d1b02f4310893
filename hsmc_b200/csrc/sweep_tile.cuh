// sweep_tile.cuh -- K2, tile-staged colour phase of the checkerboard sweep.
//
// One CTA owns a tile of AX x AY x AZ active cells of the current colour.  What it stages
// is not the double-precision master table but its float4 shadow `rel` = {x,y,z offset of
// the particle from its own cell's origin, id}: 16 B per particle, magnitudes < one cell
// edge, so fp32 carries ~1e-7 absolute accuracy.  The shadow rows of the tile's stencil
// union -- (2AX+1) x (2AY+1) cell rows of 2AZ+1 cells, each row one contiguous slot range of
// the cell-ordered table -- are pulled into shared memory by TMA bulk copies
// (cp.async.bulk, one per row piece, completion on an mbarrier); the CSR offsets of the
// staged cells are rebased to shared-memory indices; the tile's non-empty cells are sorted
// by occupancy so the lanes of a warp run equally many sequential trials.
//
// Each trial is generated in double from the master copy exactly as the reference's
// part_move() does (moves.c:52-57).  Its 27-cell stencil is then scanned out of shared
// memory in fp32 as a FILTER: pair separation = (cell-index difference) * edge +
// (offset difference), which is also the minimum image.  With a rigorous error bound
// eps on the fp32 r^2, r2f < 1 - eps is a certain overlap and r2f > 1 + eps a
// certain miss; the rare pairs in between are re-evaluated from the master table with
// the reference's exact double arithmetic (moves.c:400-431).  The verdict of every trial
// is therefore bit-identical to an all-double evaluation.
#pragma once

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool LOG>
__device__ __noinline__ void cell_update_global_noinline(const SweepArgs a, int phase, double4* __restrict__ pos,
                                                         float4* __restrict__ rel, const int* __restrict__ cs, int l,
                                                         int iy, int iz, int j0, int j1, int& n_acc, int& n_ov,
                                                         int& n_cell, hsmc_gpu_trial* __restrict__ log,
                                                         unsigned long long* __restrict__ nlog, long long logcap) {
  cell_update_global<LOG>(a, phase, pos, rel, cs, l, iy, iz, j0, j1, n_acc, n_ov, n_cell, log, nlog, logcap);
}

#define TILE_SLOTS 7     // padded stencil slots evaluated per row and pass (a z-column of 3 cells holds <= 6 in 99.7 %)

// per-row staging record
struct TileRow {
  int gbA, gbB;     // global slot of the first particle of piece A / B (B: wrapped part of the row)
  int cntA;         // particles in piece A
  int off;          // staged index of the row's first particle
  int shift;        // alignment shift of the TMA'd CSR row
  int delta;        // gbA - off: staged index = CSR value - delta
};

// exact re-evaluation of a whole stencil from the master table with the reference's
// arithmetic (moves.c:157-212, 400-431): taken only when some pair fell inside the fp32
// error band and no certain overlap was found (~3e-4 of the trials), kept out of line
__device__ __noinline__ bool tile_exact_rescan(const double4* __restrict__ pos, const TileRow* s_row,
                                               const int* s_cs, int cs_stride, int nry, int rxc, int ryc, int rz,
                                               int sel, double xn, double yn, double zn, const Box& box) {
  for (int dx = -1; dx <= 1; dx++)
    for (int dy = -1; dy <= 1; dy++) {
      int row = (rxc + dx) * nry + ryc + dy;
      TileRow rw = s_row[row];
      const int* cp = s_cs + row * cs_stride + rw.shift + rz;
      int b = cp[-1] - rw.delta, e = cp[2] - rw.delta;
      for (int k = b; k < e; k++) {
        if (k == sel) continue;
        int o = k - rw.off;
        int gs = (o < rw.cntA) ? rw.gbA + o : rw.gbB + o - rw.cntA;
        double4 q = pos[gs];
        if (pair_r2(xn, yn, zn, q.x, q.y, q.z, box) < 1.0) return true;
      }
    }
  return false;
}

template <bool LOG>
__global__ void __launch_bounds__(TILE_THREADS, 5)
k_sweep_tile(SweepArgs a, TileCfg tc, double4* __restrict__ pos, float4* __restrict__ rel,
             const int* __restrict__ cs, unsigned long long* __restrict__ cnt, hsmc_gpu_trial* __restrict__ log,
             unsigned long long* __restrict__ nlog, long long logcap) {
  const Grid& g = a.g;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* s_rel = reinterpret_cast<float4*>(smem_raw);
  int* s_cs = reinterpret_cast<int*>(s_rel + tc.cap);        // [rows][cs_stride] raw CSR values
  __shared__ TileRow s_row[TILE_MAX_ROWS];
  __shared__ int s_cnt[TILE_MAX_ROWS + 1];
  __shared__ unsigned short s_items[TILE_MAX_CELLS];
  __shared__ int s_n[4];                                   // [0] cells with >= 2, [1] cells with 1
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x;
  const int cs_stride = tc.cs_stride;

  // ---- which tile -------------------------------------------------------------------
  const int tz = blockIdx.x % tc.ntz;
  const int ty = (blockIdx.x / tc.ntz) % tc.nty;
  const int tx = blockIdx.x / (tc.ntz * tc.nty);
  const int hx = (g.own_hi - g.own_lo) >> 1, hy = g.ny >> 1, hz = g.nz >> 1;
  const int a0x = tx * tc.ax, a0y = ty * tc.ay, a0z = tz * tc.az;
  const int nax = min(tc.ax, hx - a0x), nay = min(tc.ay, hy - a0y), naz = min(tc.az, hz - a0z);
  const int par0 = (g.gx0 + g.own_lo) & 1;
  // region origin in (local layer, y, z) cell coordinates; -1 means periodic wrap
  const int x0 = g.own_lo + 2 * a0x + ((a.cx - par0) & 1) - 1;
  const int y0 = 2 * a0y + a.cy - 1;
  const int z0 = 2 * a0z + a.cz - 1;
  const int nry = 2 * nay + 1, lenz = 2 * naz + 1;
  const int nrows = (2 * nax + 1) * nry;
  const int zs = (z0 < 0) ? z0 + g.nz : z0;
  const bool zwrap = zs + lenz > g.nz;     // tile-uniform: the region crosses the periodic z edge

  if (tid == 0) {
    mbar_init(&s_bar, TILE_THREADS);
    s_n[0] = s_n[1] = 0;
  }

  // ---- row pieces: each (x,y) row of the region is one contiguous slot range, or two
  //      when it wraps in z
  // row `myrow` belongs to thread (warp, lane) with myrow = lane * W + warp: every warp
  // issues the same number of (serialised) TMA copies
  long long my_rbase = 0;
  int my_cntB = 0;
  const int myrow = ((tid & 31) < (TILE_MAX_ROWS + TILE_THREADS / 32 - 1) / (TILE_THREADS / 32))
                        ? (tid & 31) * (TILE_THREADS / 32) + (tid >> 5) : TILE_MAX_ROWS;
  if (myrow < nrows) {
    int rx = myrow / nry, ry = myrow - rx * nry;
    int lx = x0 + rx;
    if (g.wrap_x) { if (lx < 0) lx += g.nlx; else if (lx >= g.nlx) lx -= g.nlx; }
    int y = y0 + ry;
    if (y < 0) y += g.ny; else if (y >= g.ny) y -= g.ny;
    my_rbase = ((long long)lx * g.ny + y) * g.nz;
    int gbA = cs[my_rbase + zs], geA = cs[my_rbase + min(zs + lenz, g.nz)];
    int gbB = 0, geB = 0;
    if (zwrap) { gbB = cs[my_rbase]; geB = cs[my_rbase + (zs + lenz - g.nz)]; }
    my_cntB = geB - gbB;
    s_row[myrow].gbA = gbA; s_row[myrow].gbB = gbB; s_row[myrow].cntA = geA - gbA;
    s_cnt[myrow] = (geA - gbA) + my_cntB;
  }
  __syncthreads();
  // ---- exclusive scan of the row counts (<= 81 rows) by warp 0 ----------------------
  if (tid < 32) {
    int carry = 0;
    for (int base = 0; base < nrows; base += 32) {
      int r = base + tid;
      int v = (r < nrows) ? s_cnt[r] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (tid >= o) inc += t;
      }
      if (r < nrows) s_cnt[r] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (tid == 0) s_cnt[nrows] = carry;
  }
  __syncthreads();
  const int total = s_cnt[nrows];
  const bool staged = total <= tc.cap;

  int n_acc = 0, n_ov = 0, n_cell = 0;

  if (staged) {
    // ---- stage shadow rows and CSR rows: TMA bulk copies, completion on s_bar ----------
    const bool tma_cs = tc.use_tma && !zwrap;
    if (myrow < nrows) {
      TileRow& rw = s_row[myrow];
      const int off = s_cnt[myrow], cA = rw.cntA;
      const long long i0 = my_rbase + zs;
      const int shift = tma_cs ? (int)(i0 & 3) : 0;
      rw.off = off; rw.shift = shift; rw.delta = rw.gbA - off;
      if (tc.use_tma) {
        uint32_t bytes = (uint32_t)(cA + my_cntB) * 16u;
        uint32_t cs_bytes = tma_cs ? (uint32_t)((shift + lenz + 1 + 3) & ~3) * 4u : 0u;
        if (bytes + cs_bytes) mbar_arrive_tx(&s_bar, bytes + cs_bytes); else mbar_arrive(&s_bar);
        if (cA) tma_bulk_g2s(&s_rel[off], &rel[rw.gbA], (uint32_t)cA * 16u, &s_bar);
        if (my_cntB) tma_bulk_g2s(&s_rel[off + cA], &rel[rw.gbB], (uint32_t)my_cntB * 16u, &s_bar);
        if (cs_bytes) tma_bulk_g2s(&s_cs[myrow * cs_stride], &cs[i0 - shift], cs_bytes, &s_bar);
      }
    } else if (tc.use_tma) {
      mbar_arrive(&s_bar);
    }
    __syncthreads();
    if (!tc.use_tma) {
      for (int r = tid >> 5; r < nrows; r += TILE_THREADS / 32) {
        TileRow rw = s_row[r];
        int cT = s_cnt[r + 1] - s_cnt[r];
        for (int k = tid & 31; k < rw.cntA; k += 32) s_rel[rw.off + k] = rel[rw.gbA + k];
        for (int k = (tid & 31) + rw.cntA; k < cT; k += 32) s_rel[rw.off + k] = rel[rw.gbB + k - rw.cntA];
      }
    }
    if (!tma_cs) {
      // CSR rows by plain loads; entries of the wrapped part are renumbered so that
      // (value - delta) is the staged index for every entry of the row
#pragma unroll 1
      for (int r = tid >> 5; r < nrows; r += TILE_THREADS / 32) {
        int rx = r / nry, ry = r - rx * nry;
        int lx = x0 + rx;
        if (g.wrap_x) { if (lx < 0) lx += g.nlx; else if (lx >= g.nlx) lx -= g.nlx; }
        int y = y0 + ry;
        if (y < 0) y += g.ny; else if (y >= g.ny) y -= g.ny;
        long long rbase = ((long long)lx * g.ny + y) * g.nz;
        TileRow rw = s_row[r];
        for (int zi = tid & 31; zi <= lenz; zi += 32) {
          int z = zs + zi;
          s_cs[r * cs_stride + zi] = (z <= g.nz) ? cs[rbase + z] : cs[rbase + z - g.nz] - rw.gbB + rw.gbA + rw.cntA;
        }
      }
    }
    if (tc.use_tma) mbar_wait(&s_bar, 0);
    __syncthreads();

    // ---- the tile's non-empty cells: those with >= 2 particles first, then singles -----
    const int ncell_t = nax * nay * naz;
    constexpr int CELLS_PER_THREAD = (TILE_MAX_CELLS + TILE_THREADS - 1) / TILE_THREADS;
    int my_code[CELLS_PER_THREAD], my_rank[CELLS_PER_THREAD];
#pragma unroll
    for (int m = 0; m < CELLS_PER_THREAD; m++) {
      int q = tid + m * TILE_THREADS;
      my_code[m] = -1;
      my_rank[m] = 0;
      if (q < ncell_t) {
        int qz = q % naz, qy = (q / naz) % nay, qx = q / (naz * nay);
        int rxc = 2 * qx + 1, ryc = 2 * qy + 1, rz = 2 * qz + 1;
        int row = rxc * nry + ryc;
        const int* cp = s_cs + row * cs_stride + s_row[row].shift + rz;
        int n = cp[1] - cp[0];
        if (n > 0) {
          int cls = n >= 2 ? 0 : 1;
          my_rank[m] = atomicAdd(&s_n[cls], 1);
          my_code[m] = (cls << 15) | (rxc << 10) | (ryc << 6) | rz;      // rxc,ryc <= 9, rz <= 33
        }
      }
    }
    __syncthreads();
    const int nA = s_n[0], n_items = s_n[0] + s_n[1];
#pragma unroll
    for (int m = 0; m < CELLS_PER_THREAD; m++)
      if (my_code[m] >= 0)
        s_items[((my_code[m] >> 15) ? nA : 0) + my_rank[m]] = (unsigned short)(my_code[m] & 0x7fff);
    __syncthreads();

    // ---- trials: at most two per cell here (deeper cells finish in k_sweep_deep) ---------
    // lane t runs item t; lanes whose first item was a single take further singles, so
    // every lane does about two trials
    const float wxf = (float)g.wx, wyf = (float)g.wy, wzf = (float)g.wz;
    const float lo = 1.0f - a.eps, hi = 1.0f + a.eps;
    const int nB_lanes = TILE_THREADS - min(nA, TILE_THREADS);
    // One flat loop: every iteration is exactly one trial (item `it`, trial index `j`), so
    // lanes on their second trial of a cell and lanes that moved on to another cell stay
    // converged in the same code.  The master-table entry of the NEXT trial's particle is
    // requested one iteration ahead (its HBM/L2 latency is as long as a whole stencil scan).
    struct Cur {
      int rxc, ryc, rz, ob, oe, n, sel, gslot, gx, iy, iz, last_id;
      long long gcell;
    };
    auto decode = [&](int it, int j, int last_id) {
      Cur c;
      const int code = s_items[it];
      c.rxc = code >> 10; c.ryc = (code >> 6) & 15; c.rz = code & 63;
      const int rowc = c.rxc * nry + c.ryc;
      const TileRow rwc = s_row[rowc];
      const int* cpc = s_cs + rowc * cs_stride + rwc.shift + c.rz;
      c.ob = cpc[0] - rwc.delta; c.oe = cpc[1] - rwc.delta;
      c.n = c.oe - c.ob;
      const int l = x0 + c.rxc;
      c.iy = y0 + c.ryc; c.iz = z0 + c.rz;
      c.gx = (g.gx0 + l >= g.nx) ? g.gx0 + l - g.nx : g.gx0 + l;
      c.gcell = ((long long)c.gx * g.ny + c.iy) * g.nz + c.iz;
      // particle of trial j: ascending id inside the cell
      c.sel = c.ob;
      c.last_id = last_id;
      if (c.n > 1) {
        int best = 0x7fffffff;
        for (int k = c.ob; k < c.oe; k++) {
          int id = __float_as_int(s_rel[k].w);
          if (id > last_id && id < best) { best = id; c.sel = k; }
        }
        c.last_id = best;
      }
      const int ro = c.sel - rwc.off;
      c.gslot = (ro < rwc.cntA) ? rwc.gbA + ro : rwc.gbB + ro - rwc.cntA;
      return c;
    };
    int it = tid, pass = 0, j = 0;
    Cur cur;
    double4 p_next = make_double4(0, 0, 0, 0);
    if (it < n_items) { cur = decode(it, 0, -1); p_next = pos[cur.gslot]; }
#pragma unroll 1
    while (it < n_items) {
      const double4 p = p_next;
      const Cur c = cur;
      // where this lane goes next, and the early request for that particle
      int it2 = it, j2 = j + 1, pass2 = pass;
      if (j2 >= min(c.n, 2)) {
        j2 = 0;
        if (nB_lanes == 0) it2 = it + TILE_THREADS;
        else if (tid < nA) it2 = n_items;
        else { it2 = TILE_THREADS + pass * nB_lanes + (tid - nA); pass2 = pass + 1; }
      }
      if (it2 < n_items) { cur = decode(it2, j2, j2 ? c.last_id : -1); p_next = pos[cur.gslot]; }
      {
        const int rxc = c.rxc, ryc = c.ryc, rz = c.rz, sel = c.sel, gslot = c.gslot;
        const int gx = c.gx, iy = c.iy, iz = c.iz;
        const long long gcell = c.gcell;
        Philox4 rn = philox4x32_10((uint32_t)gcell, (HSMC_STREAM_MOVE << 24) | (uint32_t)j, a.sweep_lo, a.sweep_hi,
                                   a.key0, a.key1);
        double xn = p.x + (hsmc_u01(rn.v[0]) - 0.5) * a.dr_max;
        double yn = p.y + (hsmc_u01(rn.v[1]) - 0.5) * a.dr_max;
        double zn = p.z + (hsmc_u01(rn.v[2]) - 0.5) * a.dr_max;
        if (xn > g.Lx) xn -= g.Lx; else if (xn < 0.0) xn += g.Lx;
        if (yn > g.Ly) yn -= g.Ly; else if (yn < 0.0) yn += g.Ly;
        if (zn > g.Lz) zn -= g.Lz; else if (zn < 0.0) zn += g.Lz;
        int verdict;
        if (axis_cell(xn, g.sx, g.iwx, g.nx) != gx || axis_cell(yn, g.sy, g.iwy, g.ny) != iy ||
            axis_cell(zn, g.sz, g.iwz, g.nz) != iz) {
          verdict = 2;
          n_cell++;
        } else {
          // offsets of the trial position from its cell origin (the cell may straddle the box edge)
          const float4 nrel = make_rel(g, gx, iy, iz, xn, yn, zn, p.w);
          const float fzm = nrel.z + wzf, fzp = nrel.z - wzf;
          bool ov = false, inband = false;
#pragma unroll 1
          for (int dx = -1; dx <= 1; dx++) {
            const float fx = nrel.x - (float)dx * wxf;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++) {
              const int row = (rxc + dx) * nry + ryc + dy;
              const int shift = s_row[row].shift, delta = s_row[row].delta;
              const int* cp = s_cs + row * cs_stride + shift + rz;
              const int b = cp[-1] - delta, m1 = cp[0] - delta, m2 = cp[1] - delta, e = cp[2] - delta;
              const float fy = nrel.y - (float)dy * wyf;
#pragma unroll 1
              for (int k0 = b; k0 < e; k0 += TILE_SLOTS) {
#pragma unroll
                for (int s = 0; s < TILE_SLOTS; s++) {
                  int k = k0 + s;
                  bool v = (k < e) && (k != sel);
                  int kk = v ? k : sel;
                  float4 qv = s_rel[kk];
                  float fz = (k < m1) ? fzm : ((k < m2) ? nrel.z : fzp);
                  float ddx = fx - qv.x, ddy = fy - qv.y, ddz = fz - qv.z;
                  float r2 = __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, ddx * ddx));
                  ov |= v && (r2 < lo);
                  inband |= v && (r2 <= hi);
                }
              }
            }
          }
          if (!ov && inband)
            ov = tile_exact_rescan(pos, s_row, s_cs, cs_stride, nry, rxc, ryc, rz, sel, xn, yn, zn, a.box);
          if (ov) { verdict = 1; n_ov++; }
          else {
            verdict = 0; n_acc++;
            s_rel[sel] = nrel;
            rel[gslot] = nrel;
            pos[gslot] = make_double4(xn, yn, zn, p.w);
          }
        }
        if (LOG) {
          unsigned long long s = atomicAdd(nlog, 1ull);
          if ((long long)s < logcap) {
            hsmc_gpu_trial tr;
            tr.seq = ((unsigned long long)a.phase << 56) | ((unsigned long long)gcell << 8) | (unsigned)j;
            tr.id = (int)p.w; tr.verdict = verdict;
            tr.raw[0] = rn.v[0]; tr.raw[1] = rn.v[1]; tr.raw[2] = rn.v[2]; tr.pad = 0;
            log[s] = tr;
          }
        }
      }
      it = it2; j = j2; pass = pass2;
    }
  } else {
    // ---- staging capacity exceeded (unusually dense tile): global-memory path ---------
    const int ncell_t = nax * nay * naz;
    for (int q = tid; q < ncell_t; q += TILE_THREADS) {
      int qz = q % naz, qy = (q / naz) % nay, qx = q / (naz * nay);
      int l = x0 + 2 * qx + 1, iy = y0 + 2 * qy + 1, iz = z0 + 2 * qz + 1;
      cell_update_global_noinline<LOG>(a, a.phase, pos, rel, cs, l, iy, iz, 0, 2, n_acc, n_ov, n_cell, log, nlog, logcap);
    }
  }

  // ---- counters: warp reduce, then straight to the global counters (no CTA barrier) -----
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  n_ov = __reduce_add_sync(0xffffffffu, n_ov);
  n_cell = __reduce_add_sync(0xffffffffu, n_cell);
  if ((tid & 31) == 0 && (n_acc | n_ov | n_cell)) {
    atomicAdd(&cnt[CNT_TRIALS], (unsigned long long)(n_acc + n_ov + n_cell));
    if (n_acc) atomicAdd(&cnt[CNT_ACC], (unsigned long long)n_acc);
    if (n_ov) atomicAdd(&cnt[CNT_REJ_OVERLAP], (unsigned long long)n_ov);
    if (n_cell) atomicAdd(&cnt[CNT_REJ_CELL], (unsigned long long)n_cell);
  }
}

// ---- trials beyond the second in cells holding three or more particles ----------------
// (2-3 % of the cells at rho 0.9).  The per-colour lists are built with the cell list.
// One WARP per listed cell, two memory round trips per cell: (1) the CSR entries of the
// 27 stencil cells, lane c < 27 its own cell; (2) the cell's own particles (lane i the
// i-th) and up to three particles of each neighbour cell, cached in registers -- they do
// not change while this cell is updated.  Trials (index continues at j = 2) are generated
// redundantly by all lanes (uniform), each lane tests its cached particles with the
// reference's double arithmetic, the verdict is a ballot.
#define DEEP_CACHE 3
template <bool LOG>
__global__ void __launch_bounds__(256, 3)
k_sweep_deep(SweepArgs a, const int* __restrict__ deep_list, const int* __restrict__ deep_count, int colour,
             int list_stride, double4* __restrict__ pos, float4* __restrict__ rel, const int* __restrict__ cs,
             unsigned long long* __restrict__ cnt, hsmc_gpu_trial* __restrict__ log,
             unsigned long long* __restrict__ nlog, long long logcap) {
  const Grid& g = a.g;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  int n_acc = 0, n_ov = 0, n_cell = 0;
  const int nlist = min(deep_count[colour], list_stride);
  const int ddx = lane / 9 - 1, ddy = (lane / 3) % 3 - 1, ddz = lane % 3 - 1;   // this lane's stencil cell
  for (int i = warp; i < nlist; i += nwarps) {
    const int c = deep_list[(long long)colour * list_stride + i];
    const int iz = c % g.nz;
    const int r = c / g.nz;
    const int iy = r % g.ny, l = r / g.ny;
    const int gx = (g.gx0 + l >= g.nx) ? g.gx0 + l - g.nx : g.gx0 + l;
    const long long gcell = ((long long)gx * g.ny + iy) * g.nz + iz;
    // round trip 1: CSR of the stencil
    int nb = 0, ne = 0;
    if (lane < 27) {
      int ll = l + ddx, yy = iy + ddy, zz = iz + ddz;
      if (g.wrap_x) { if (ll < 0) ll += g.nlx; else if (ll >= g.nlx) ll -= g.nlx; }
      if (yy < 0) yy += g.ny; else if (yy >= g.ny) yy -= g.ny;
      if (zz < 0) zz += g.nz; else if (zz >= g.nz) zz -= g.nz;
      long long cc = ((long long)ll * g.ny + yy) * g.nz + zz;
      nb = cs[cc]; ne = cs[cc + 1];
    }
    const int beg = __shfl_sync(FULL, nb, 13), end = __shfl_sync(FULL, ne, 13);
    if (end - beg > 32) {      // more particles than lanes (only with very wide cells): generic path
      if (lane == 0) cell_update_global<LOG>(a, a.phase, pos, rel, cs, l, iy, iz, 2, 1 << 30, n_acc, n_ov, n_cell, log, nlog, logcap);
      __syncwarp();
      continue;
    }
    const int n = end - beg;
    if (lane == 13) ne = nb;                         // the own cell is handled through `own`
    // round trip 2: own particles + cached neighbours
    double4 own = make_double4(0, 0, 0, 1e300);
    if (lane < n) own = pos[beg + lane];
    double4 q[DEEP_CACHE];
#pragma unroll
    for (int s = 0; s < DEEP_CACHE; s++) q[s] = (nb + s < ne) ? pos[nb + s] : make_double4(1e30, 1e30, 1e30, 0);
    // rank of this lane's own particle in ascending-id order
    int rank = 0;
    for (int k = 0; k < n; k++) {
      double idk = __shfl_sync(FULL, own.w, k);
      rank += (idk < own.w) ? 1 : 0;
    }
    for (int j = 2; j < n; j++) {
      const int src = __ffs(__ballot_sync(FULL, lane < n && rank == j)) - 1;   // lane holding trial j's particle
      const double px = __shfl_sync(FULL, own.x, src), py = __shfl_sync(FULL, own.y, src);
      const double pz = __shfl_sync(FULL, own.z, src), pw = __shfl_sync(FULL, own.w, src);
      Philox4 rn = philox4x32_10((uint32_t)gcell, (HSMC_STREAM_MOVE << 24) | (uint32_t)j, a.sweep_lo, a.sweep_hi,
                                 a.key0, a.key1);
      double xn = px + (hsmc_u01(rn.v[0]) - 0.5) * a.dr_max;
      double yn = py + (hsmc_u01(rn.v[1]) - 0.5) * a.dr_max;
      double zn = pz + (hsmc_u01(rn.v[2]) - 0.5) * a.dr_max;
      if (xn > g.Lx) xn -= g.Lx; else if (xn < 0.0) xn += g.Lx;
      if (yn > g.Ly) yn -= g.Ly; else if (yn < 0.0) yn += g.Ly;
      if (zn > g.Lz) zn -= g.Lz; else if (zn < 0.0) zn += g.Lz;
      int verdict;
      if (axis_cell(xn, g.sx, g.iwx, g.nx) != gx || axis_cell(yn, g.sy, g.iwy, g.ny) != iy ||
          axis_cell(zn, g.sz, g.iwz, g.nz) != iz) {
        verdict = 2;
        if (lane == 0) n_cell++;
      } else {
        bool ov = false;
        if (lane < n && lane != src) ov = pair_r2(xn, yn, zn, own.x, own.y, own.z, a.box) < 1.0;
#pragma unroll
        for (int s = 0; s < DEEP_CACHE; s++)
          ov |= (nb + s < ne) && pair_r2(xn, yn, zn, q[s].x, q[s].y, q[s].z, a.box) < 1.0;
        for (int k = nb + DEEP_CACHE; k < ne; k++) {                 // rare: a neighbour cell with > 4 particles
          double4 qq = pos[k];
          ov |= pair_r2(xn, yn, zn, qq.x, qq.y, qq.z, a.box) < 1.0;
        }
        if (__any_sync(FULL, ov)) {
          verdict = 1;
          if (lane == 0) n_ov++;
        } else {
          verdict = 0;
          if (lane == 0) n_acc++;
          if (lane == src) {
            own.x = xn; own.y = yn; own.z = zn;
            pos[beg + src] = own;
            rel[beg + src] = make_rel(g, gx, iy, iz, xn, yn, zn, pw);
          }
        }
      }
      if (LOG && lane == 0) {
        unsigned long long s = atomicAdd(nlog, 1ull);
        if ((long long)s < logcap) {
          hsmc_gpu_trial tr;
          tr.seq = ((unsigned long long)a.phase << 56) | ((unsigned long long)gcell << 8) | (unsigned)j;
          tr.id = (int)pw; tr.verdict = verdict;
          tr.raw[0] = rn.v[0]; tr.raw[1] = rn.v[1]; tr.raw[2] = rn.v[2]; tr.pad = 0;
          log[s] = tr;
        }
      }
    }
  }
  if (lane == 0 && (n_acc | n_ov | n_cell)) {
    atomicAdd(&cnt[CNT_TRIALS], (unsigned long long)(n_acc + n_ov + n_cell));
    if (n_acc) atomicAdd(&cnt[CNT_ACC], (unsigned long long)n_acc);
    if (n_ov) atomicAdd(&cnt[CNT_REJ_OVERLAP], (unsigned long long)n_ov);
    if (n_cell) atomicAdd(&cnt[CNT_REJ_CELL], (unsigned long long)n_cell);
  }
}
