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
__device__ __noinline__ void cell_update_global_noinline(const SweepArgs& a, double4* __restrict__ pos,
                                                         float4* __restrict__ rel, const int* __restrict__ cs, int l,
                                                         int iy, int iz, int& n_acc, int& n_ov, int& n_cell,
                                                         hsmc_gpu_trial* __restrict__ log,
                                                         unsigned long long* __restrict__ nlog, long long logcap) {
  cell_update_global<LOG>(a, pos, rel, cs, l, iy, iz, n_acc, n_ov, n_cell, log, nlog, logcap);
}

// exact verdict for one pair inside the fp32 error band: staged index -> global slot ->
// reference arithmetic on the master table (rare, kept out of line)
__device__ __noinline__ bool tile_exact_overlap(const double4* __restrict__ pos, const int* s_gbA, const int* s_gbB,
                                                const int* s_cntA, const int* s_off, int row, int k, double xn,
                                                double yn, double zn, const Box& box) {
  int o = k - s_off[row];
  int gs = (o < s_cntA[row]) ? s_gbA[row] + o : s_gbB[row] + o - s_cntA[row];
  double4 q = pos[gs];
  return pair_r2(xn, yn, zn, q.x, q.y, q.z, box) < 1.0;
}

#define TILE_CS_STRIDE (2 * TILE_MAX_AZ + 2)

template <bool LOG>
__global__ void __launch_bounds__(TILE_THREADS, 4)
k_sweep_tile(SweepArgs a, TileCfg tc, double4* __restrict__ pos, float4* __restrict__ rel,
             const int* __restrict__ cs, unsigned long long* __restrict__ cnt, hsmc_gpu_trial* __restrict__ log,
             unsigned long long* __restrict__ nlog, long long logcap) {
  const Grid& g = a.g;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* s_rel = reinterpret_cast<float4*>(smem_raw);
  unsigned short* s_cs = reinterpret_cast<unsigned short*>(s_rel + tc.cap);   // [rows][TILE_CS_STRIDE]
  __shared__ int s_gbA[TILE_MAX_ROWS], s_gbB[TILE_MAX_ROWS], s_cntA[TILE_MAX_ROWS];
  __shared__ int s_off[TILE_MAX_ROWS + 1];
  __shared__ int s_items[TILE_MAX_CELLS];
  __shared__ int s_ccnt[8], s_coff[8];
  __shared__ int s_acc[3];
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x;

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

  if (tid == 0) {
    mbar_init(&s_bar, TILE_THREADS);
    s_acc[0] = s_acc[1] = s_acc[2] = 0;
  }
  if (tid < 8) s_ccnt[tid] = 0;

  // ---- one round trip to the CSR offsets: row pieces + every staged cell's offset ----
  // each (x,y) row of the region is one contiguous slot range, or two when it wraps in z
  int my_cntB = 0;
  if (tid < nrows) {
    int rx = tid / nry, ry = tid - rx * nry;
    int lx = x0 + rx;
    if (g.wrap_x) { if (lx < 0) lx += g.nlx; else if (lx >= g.nlx) lx -= g.nlx; }
    int y = y0 + ry;
    if (y < 0) y += g.ny; else if (y >= g.ny) y -= g.ny;
    long long rbase = ((long long)lx * g.ny + y) * g.nz;
    int gbA = cs[rbase + zs], geA = cs[rbase + min(zs + lenz, g.nz)];
    int gbB = 0, geB = 0;
    if (zs + lenz > g.nz) { gbB = cs[rbase]; geB = cs[rbase + (zs + lenz - g.nz)]; }
    s_gbA[tid] = gbA; s_cntA[tid] = geA - gbA;
    s_gbB[tid] = gbB; my_cntB = geB - gbB;
    s_off[tid] = (geA - gbA) + my_cntB;   // row count, scanned in place below
  }
  constexpr int RAW_PER_THREAD = (TILE_MAX_ROWS * TILE_CS_STRIDE + TILE_THREADS - 1) / TILE_THREADS;
  int raw[RAW_PER_THREAD];
  const int ncs = nrows * (lenz + 1);
#pragma unroll
  for (int m = 0; m < RAW_PER_THREAD; m++) {
    int idx = tid + m * TILE_THREADS;
    raw[m] = 0;
    if (idx < ncs) {
      int r = idx / (lenz + 1), zi = idx - r * (lenz + 1);
      int rx = r / nry, ry = r - rx * nry;
      int lx = x0 + rx;
      if (g.wrap_x) { if (lx < 0) lx += g.nlx; else if (lx >= g.nlx) lx -= g.nlx; }
      int y = y0 + ry;
      if (y < 0) y += g.ny; else if (y >= g.ny) y -= g.ny;
      long long rbase = ((long long)lx * g.ny + y) * g.nz;
      int z = zs + zi;
      raw[m] = cs[rbase + (z <= g.nz ? z : z - g.nz)];
    }
  }
  __syncthreads();
  // ---- exclusive scan of the row counts (<= 81 rows) by warp 0 ----------------------
  if (tid < 32) {
    int carry = 0;
    for (int base = 0; base < nrows; base += 32) {
      int r = base + tid;
      int v = (r < nrows) ? s_off[r] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (tid >= o) inc += t;
      }
      if (r < nrows) s_off[r] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (tid == 0) s_off[nrows] = carry;
  }
  __syncthreads();
  const int total = s_off[nrows];
  const bool staged = total <= tc.cap;

  int n_acc = 0, n_ov = 0, n_cell = 0;

  if (staged) {
    // ---- stage the shadow rows: TMA bulk copies (or plain loads), completion on s_bar --
    if (tc.use_tma) {
      if (tid < nrows) {
        int cA = s_cntA[tid], off = s_off[tid];
        uint32_t bytes = (uint32_t)(cA + my_cntB) * 16u;
        if (bytes) mbar_arrive_tx(&s_bar, bytes); else mbar_arrive(&s_bar);
        if (cA) tma_bulk_g2s(&s_rel[off], &rel[s_gbA[tid]], (uint32_t)cA * 16u, &s_bar);
        if (my_cntB) tma_bulk_g2s(&s_rel[off + cA], &rel[s_gbB[tid]], (uint32_t)my_cntB * 16u, &s_bar);
      } else {
        mbar_arrive(&s_bar);
      }
    } else {
      for (int r = tid >> 5; r < nrows; r += TILE_THREADS / 32) {
        int cA = s_cntA[r], cT = s_off[r + 1] - s_off[r], off = s_off[r], gA = s_gbA[r], gB = s_gbB[r];
        for (int k = tid & 31; k < cA; k += 32) s_rel[off + k] = rel[gA + k];
        for (int k = (tid & 31) + cA; k < cT; k += 32) s_rel[off + k] = rel[gB + k - cA];
      }
    }
    // ---- rebase the CSR offsets held in registers to shared-memory indices -------------
#pragma unroll
    for (int m = 0; m < RAW_PER_THREAD; m++) {
      int idx = tid + m * TILE_THREADS;
      if (idx < ncs) {
        int r = idx / (lenz + 1), zi = idx - r * (lenz + 1);
        int o = (zs + zi <= g.nz) ? raw[m] - s_gbA[r] : s_cntA[r] + raw[m] - s_gbB[r];
        s_cs[r * TILE_CS_STRIDE + zi] = (unsigned short)(s_off[r] + o);
      }
    }
    __syncthreads();
    // ---- the tile's non-empty cells, sorted by occupancy (>=4, 3, 2, 1) ----------------
    const int ncell_t = nax * nay * naz;
    constexpr int CELLS_PER_THREAD = (TILE_MAX_CELLS + TILE_THREADS - 1) / TILE_THREADS;
    int my_cls[CELLS_PER_THREAD], my_rank[CELLS_PER_THREAD];
#pragma unroll
    for (int m = 0; m < CELLS_PER_THREAD; m++) {
      int q = tid + m * TILE_THREADS;
      my_cls[m] = -1;
      my_rank[m] = 0;
      if (q < ncell_t) {
        int qz = q % naz, qy = (q / naz) % nay, qx = q / (naz * nay);
        int row = (2 * qx + 1) * nry + (2 * qy + 1), rz = 2 * qz + 1;
        int n = (int)s_cs[row * TILE_CS_STRIDE + rz + 1] - (int)s_cs[row * TILE_CS_STRIDE + rz];
        if (n > 0) {
          my_cls[m] = n >= 4 ? 0 : 4 - n;
          my_rank[m] = atomicAdd(&s_ccnt[my_cls[m]], 1);
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      int o = 0;
      for (int c = 0; c < 4; c++) { s_coff[c] = o; o += s_ccnt[c]; }
      s_coff[4] = o;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < CELLS_PER_THREAD; m++)
      if (my_cls[m] >= 0) s_items[s_coff[my_cls[m]] + my_rank[m]] = tid + m * TILE_THREADS;
    const int n_items = s_coff[4];
    if (tc.use_tma) mbar_wait(&s_bar, 0);
    __syncthreads();

    // ---- trials ----------------------------------------------------------------------
    const float wxf = (float)g.wx, wyf = (float)g.wy, wzf = (float)g.wz;
    const float lo = 1.0f - a.eps, hi = 1.0f + a.eps;
    for (int it = tid; it < n_items; it += TILE_THREADS) {
      const int q = s_items[it];
      const int qz = q % naz, qy = (q / naz) % nay, qx = q / (naz * nay);
      const int rxc = 2 * qx + 1, ryc = 2 * qy + 1, rz = 2 * qz + 1;
      const int rowc = rxc * nry + ryc;
      const int ob = s_cs[rowc * TILE_CS_STRIDE + rz], oe = s_cs[rowc * TILE_CS_STRIDE + rz + 1];
      const int n = oe - ob;
      // cell coordinates (local layer, y, z), its origin, global slot of its first particle
      const int l = x0 + rxc, iy = y0 + ryc, iz = z0 + rz;
      const int gx = (g.gx0 + l >= g.nx) ? g.gx0 + l - g.nx : g.gx0 + l;
      const long long gcell = ((long long)gx * g.ny + iy) * g.nz + iz;
      const int ro = ob - s_off[rowc];
      const int gslot0 = (ro < s_cntA[rowc]) ? s_gbA[rowc] + ro : s_gbB[rowc] + ro - s_cntA[rowc];
      int last_id = -1;
      for (int j = 0; j < n; j++) {
        int sel = ob;
        if (n > 1) {
          int best = 0x7fffffff;
          for (int k = ob; k < oe; k++) {
            int id = __float_as_int(s_rel[k].w);
            if (id > last_id && id < best) { best = id; sel = k; }
          }
          last_id = best;
        }
        const int gslot = gslot0 + (sel - ob);
        const double4 p = pos[gslot];
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
          bool ov = false;
#pragma unroll 1
          for (int dx = -1; dx <= 1; dx++) {
            const unsigned short* rowp = s_cs + ((rxc + dx) * nry + (ryc - 1)) * TILE_CS_STRIDE + rz;
            const float fx = nrel.x - (float)dx * wxf;
#pragma unroll
            for (int dy = 0; dy < 3; dy++) {
              const int b = rowp[dy * TILE_CS_STRIDE - 1], m1 = rowp[dy * TILE_CS_STRIDE];
              const int m2 = rowp[dy * TILE_CS_STRIDE + 1], e = rowp[dy * TILE_CS_STRIDE + 2];
              const float fy = nrel.y - (float)(dy - 1) * wyf;
#pragma unroll
              for (int s = 0; s < 4; s++) {
                int k = b + s;
                bool v = (k < e) && (k != sel);
                int kk = v ? k : sel;
                float4 qv = s_rel[kk];
                float fz = nrel.z - (float)((k >= m1) + (k >= m2) - 1) * wzf;
                float ddx = fx - qv.x, ddy = fy - qv.y, ddz = fz - qv.z;
                float r2 = __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, ddx * ddx));
                ov |= v && (r2 < lo);
                if (v && r2 >= lo && r2 <= hi)
                  ov |= tile_exact_overlap(pos, s_gbA, s_gbB, s_cntA, s_off, (rxc + dx) * nry + ryc - 1 + dy, kk, xn,
                                           yn, zn, a.box);
              }
              for (int k = b + 4; k < e; k++) {
                if (k == sel) continue;
                float4 qv = s_rel[k];
                float fz = nrel.z - (float)((k >= m1) + (k >= m2) - 1) * wzf;
                float ddx = fx - qv.x, ddy = fy - qv.y, ddz = fz - qv.z;
                float r2 = __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, ddx * ddx));
                ov |= (r2 < lo);
                if (r2 >= lo && r2 <= hi)
                  ov |= tile_exact_overlap(pos, s_gbA, s_gbB, s_cntA, s_off, (rxc + dx) * nry + ryc - 1 + dy, k, xn,
                                           yn, zn, a.box);
              }
            }
          }
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
    }
  } else {
    // ---- staging capacity exceeded (unusually dense tile): global-memory path ---------
    const int ncell_t = nax * nay * naz;
    for (int q = tid; q < ncell_t; q += TILE_THREADS) {
      int qz = q % naz, qy = (q / naz) % nay, qx = q / (naz * nay);
      int l = x0 + 2 * qx + 1, iy = y0 + 2 * qy + 1, iz = z0 + 2 * qz + 1;
      cell_update_global_noinline<LOG>(a, pos, rel, cs, l, iy, iz, n_acc, n_ov, n_cell, log, nlog, logcap);
    }
  }

  __syncthreads();
  if (n_acc) atomicAdd(&s_acc[0], n_acc);
  if (n_ov) atomicAdd(&s_acc[1], n_ov);
  if (n_cell) atomicAdd(&s_acc[2], n_cell);
  __syncthreads();
  if (tid == 0) {
    int tot = s_acc[0] + s_acc[1] + s_acc[2];
    if (tot) {
      atomicAdd(&cnt[CNT_TRIALS], (unsigned long long)tot);
      if (s_acc[0]) atomicAdd(&cnt[CNT_ACC], (unsigned long long)s_acc[0]);
      if (s_acc[1]) atomicAdd(&cnt[CNT_REJ_OVERLAP], (unsigned long long)s_acc[1]);
      if (s_acc[2]) atomicAdd(&cnt[CNT_REJ_CELL], (unsigned long long)s_acc[2]);
    }
  }
}
