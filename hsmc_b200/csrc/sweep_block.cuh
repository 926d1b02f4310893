// sweep_block.cuh -- K2, block-resident checkerboard sweep (the default sweep kernel).
//
// Two-level checkerboard.  The cell grid is cut into BLOCKS of up to MBX x MBY x MBZ cells
// (an even number of blocks per axis); blocks are coloured by the parity of their block
// index, 2x2x2 = 8 block phases = 8 kernel launches per sweep.  Two blocks of one phase are
// separated by a whole block (>= 1 cell >= sigma), so while a phase runs, the one-cell halo
// around each active block belongs to inactive blocks and is frozen: active blocks are
// independent and one CTA owns each.  The CTA stages its block plus halo ONCE in shared
// memory and then runs all eight CELL colours of the block back to back (cells of one
// colour are independent, as in the single-level scheme; a CTA barrier separates the
// colours).  Every particle still gets exactly one trial per sweep, each trial is the
// reference's part_move() (moves.c:27-80), and a trial that leaves its cell is rejected.
// Compared with one launch per cell colour this stages each particle 8x less often and
// amortises the staging prologue over ~8x more trials.
//
// Staged data: the float4 shadow `rel` = {offset from the own cell's origin, id}, pulled
// in by TMA bulk copies row by row ((x,y) rows are contiguous slot ranges of the
// cell-ordered table), then converted in place to BLOCK-relative coordinates
// fma(cell index - centre, edge, offset), so that the stencil scan needs no per-cell shifts
// and no minimum image: d = trial - neighbour directly.  Each of the 9 (x,y) rows of a
// trial's stencil is one contiguous range of three z-cells; it is scanned in groups of
// BLK_SLOTS entries read at fixed offsets WITHOUT masking: entries past the end of the
// range are real particles farther along (true positions, so harmless) or the far-away pad
// after the last staged particle.  The trial particle hides itself by parking a far-away
// position in its own slot during the scan.  The scan keeps min r^2 in fp32 as a FILTER with
// a rigorous error bound eps: min < 1 - eps is a certain overlap, min > 1 + eps a
// certain miss, anything in between is re-evaluated from the master table with the
// reference's exact double arithmetic (moves.c:400-431).  Verdicts are bit-identical to an
// all-double evaluation.
#pragma once

#ifndef BLK_THREADS
#define BLK_THREADS 128
#endif
#ifndef BLK_MIN_CTAS
#define BLK_MIN_CTAS 4
#endif
#define BLK_MAX_ROWS 128        // (MBX+2)*(MBY+2) <= BLK_MAX_ROWS: one staging row per thread
#define BLK_SLOTS 7
#define BLK_PAD 8               // far-away entries after the last staged particle
#define BLK_FAR 1.0e15f

// BlockCfg: see hsmc_gpu.cu (the handle keeps one)

// exact re-evaluation of a whole stencil from the master table (moves.c:157-212, 400-431)
// (`pos` deliberately not const __restrict__: entries of this block were written earlier in this
// launch by other threads of the CTA, so the loads must not take the non-coherent path)
__device__ __noinline__ bool block_exact_rescan(const double4* pos, const TileRow* s_row,
                                                const unsigned short* s_cz, int cz_stride, int nry, int rxc,
                                                int ryc, int rz, int sel, double xn, double yn, double zn,
                                                const Box& box) {
  for (int dx = -1; dx <= 1; dx++)
    for (int dy = -1; dy <= 1; dy++) {
      const int row = (rxc + dx) * nry + ryc + dy;
      const TileRow rw = s_row[row];
      const unsigned short* cp = s_cz + row * cz_stride + rz;
      const int b = cp[-1], e = cp[2];
      for (int k = b; k < e; k++) {
        if (k == sel) continue;
        const int o = k - rw.off;
        const int gs = (o < rw.cntA) ? rw.gbA + o : rw.gbB + o - rw.cntA;
        const double4 q = pos[gs];
        if (pair_r2(xn, yn, zn, q.x, q.y, q.z, box) < 1.0) return true;
      }
    }
  return false;
}

template <bool LOG>
__global__ void __launch_bounds__(BLK_THREADS, BLK_MIN_CTAS)
k_sweep_block(SweepArgs a, BlockCfg bc, const int* __restrict__ xoff, double4* __restrict__ pos,
              float4* __restrict__ rel, const int* __restrict__ cs, unsigned long long* __restrict__ cnt,
              hsmc_gpu_trial* __restrict__ log, unsigned long long* __restrict__ nlog, long long logcap) {
  const Grid& g = a.g;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* s_rel = reinterpret_cast<float4*>(smem_raw);
  int* s_raw = reinterpret_cast<int*>(s_rel + bc.cap);                         // [rows][cs_stride] raw CSR values
  unsigned short* s_items = reinterpret_cast<unsigned short*>(s_raw);          // aliases s_raw once the CSR is compact
  unsigned short* s_cz = reinterpret_cast<unsigned short*>(s_raw + bc.max_rows * bc.cs_stride);   // [rows][cz_stride]
  __shared__ TileRow s_row[BLK_MAX_ROWS];
  __shared__ int s_cnt[BLK_MAX_ROWS + 1];
  __shared__ int s_n[24], s_ibase[25], s_fill[24], s_next[8];
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = BLK_THREADS / 32;
  const int cs_stride = bc.cs_stride, czs = bc.cz_stride;

  // ---- which block ------------------------------------------------------------------
  const int hbz = bc.nbz >> 1, hby = bc.nby >> 1;
  const int bzi = 2 * (blockIdx.x % hbz) + a.cz;
  const int byi = 2 * ((blockIdx.x / hbz) % hby) + a.cy;
  const int bxi = 2 * (blockIdx.x / (hbz * hby)) + a.cx;
  const int xa = xoff[bxi], xb = xoff[bxi + 1];                                // local layers [xa, xb)
  const int ya = (int)((long long)byi * g.ny / bc.nby), yb = (int)((long long)(byi + 1) * g.ny / bc.nby);
  const int za = (int)((long long)bzi * g.nz / bc.nbz), zb = (int)((long long)(bzi + 1) * g.nz / bc.nbz);
  const int nbx = xb - xa, nby = yb - ya, nbz = zb - za;                       // interior extent
  const int x0 = xa - 1, y0 = ya - 1, z0 = za - 1;                             // region origin (-1: periodic wrap)
  const int nrx = nbx + 2, nry = nby + 2, lenz = nbz + 2;
  const int nrows = nrx * nry;
  const int zs = (z0 < 0) ? z0 + g.nz : z0;
  const bool zwrap = zs + lenz > g.nz;      // block-uniform: the region crosses the periodic z edge

  if (tid == 0) mbar_init(&s_bar, BLK_THREADS);
  if (tid < 24) { s_n[tid] = 0; }
  if (tid < 8) s_next[tid] = BLK_THREADS;

  // ---- row pieces: row r belongs to thread (lane, warp) with r = lane*NW + warp, so every
  //      warp issues the same number of (serialised) TMA copies
  long long my_rbase = 0;
  int my_cntB = 0;
  const int myrow = lane * NW + warp;
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
  // ---- exclusive scan of the row counts by warp 0 -----------------------------------
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
  const bool staged = !bc.force_global && total + BLK_PAD <= bc.cap;

  int n_acc = 0, n_ov = 0, n_cell = 0;

  if (staged) {
    // ---- stage shadow rows and CSR rows: TMA bulk copies, completion on s_bar ----------
    const bool tma_cs = bc.use_tma && !zwrap;
    if (myrow < nrows) {
      TileRow& rw = s_row[myrow];
      const int off = s_cnt[myrow], cA = rw.cntA;
      const long long i0 = my_rbase + zs;
      const int shift = tma_cs ? (int)(i0 & 3) : 0;
      rw.off = off; rw.shift = shift; rw.delta = rw.gbA - off;
      if (bc.use_tma) {
        uint32_t bytes = (uint32_t)(cA + my_cntB) * 16u;
        uint32_t cs_bytes = tma_cs ? (uint32_t)((shift + lenz + 1 + 3) & ~3) * 4u : 0u;
        if (bytes + cs_bytes) mbar_arrive_tx(&s_bar, bytes + cs_bytes); else mbar_arrive(&s_bar);
        if (cA) tma_bulk_g2s(&s_rel[off], &rel[rw.gbA], (uint32_t)cA * 16u, &s_bar);
        if (my_cntB) tma_bulk_g2s(&s_rel[off + cA], &rel[rw.gbB], (uint32_t)my_cntB * 16u, &s_bar);
        if (cs_bytes) tma_bulk_g2s(&s_raw[myrow * cs_stride], &cs[i0 - shift], cs_bytes, &s_bar);
      }
    } else if (bc.use_tma) {
      mbar_arrive(&s_bar);
    }
    if (tid < BLK_PAD) s_rel[total + tid] = make_float4(BLK_FAR, BLK_FAR, BLK_FAR, __int_as_float(-1));
    __syncthreads();
    if (!bc.use_tma) {
      for (int r = warp; r < nrows; r += NW) {
        TileRow rw = s_row[r];
        int cT = s_cnt[r + 1] - s_cnt[r];
        for (int k = lane; k < rw.cntA; k += 32) s_rel[rw.off + k] = rel[rw.gbA + k];
        for (int k = lane + rw.cntA; k < cT; k += 32) s_rel[rw.off + k] = rel[rw.gbB + k - rw.cntA];
      }
    }
    if (bc.use_tma) mbar_wait(&s_bar, 0);
    // ---- compact CSR: staged index of the first particle of every region cell ----------
#pragma unroll 1
    for (int r = warp; r < nrows; r += NW) {
      const TileRow rw = s_row[r];
      if (lane <= lenz) {
        int v;
        if (tma_cs) v = s_raw[r * cs_stride + rw.shift + lane];
        else {
          int rx = r / nry, ry = r - rx * nry;
          int lx = x0 + rx;
          if (g.wrap_x) { if (lx < 0) lx += g.nlx; else if (lx >= g.nlx) lx -= g.nlx; }
          int y = y0 + ry;
          if (y < 0) y += g.ny; else if (y >= g.ny) y -= g.ny;
          const long long rbase = ((long long)lx * g.ny + y) * g.nz;
          const int z = zs + lane;
          v = (z <= g.nz) ? cs[rbase + z] : cs[rbase + z - g.nz] - rw.gbB + rw.gbA + rw.cntA;
        }
        s_cz[r * czs + lane] = (unsigned short)(v - rw.delta);
      }
    }
    __syncthreads();

    // ---- block-relative coordinates; census of the interior cells by (colour, occupancy) ---
    const float wxf = (float)g.wx, wyf = (float)g.wy, wzf = (float)g.wz;
    const float hxr = 0.5f * (float)nrx, hyr = 0.5f * (float)nry, hzr = 0.5f * (float)lenz;
    const int parx = (g.gx0 + x0) & 1, pary = y0 & 1, parz = z0 & 1;    // parity of region cell (0,0,0); grids are even
#pragma unroll 1
    for (int idx = tid; idx < 2 * nrows; idx += BLK_THREADS) {
      const int r = idx >> 1, half = idx & 1;
      const int rx = r / nry, ry = r - rx * nry;
      const int zlo = half ? (lenz >> 1) : 0, zhi = half ? lenz : (lenz >> 1);
      const float cxf = (float)rx - hxr, cyf = (float)ry - hyr;
      const bool rowint = rx >= 1 && rx <= nbx && ry >= 1 && ry <= nby;
      const int colxy = (((parx + rx) & 1) << 2) | (((pary + ry) & 1) << 1);
      const unsigned short* cz = s_cz + r * czs;
      int b = cz[zlo];
      for (int zi = zlo; zi < zhi; zi++) {
        const int e = cz[zi + 1];
        const float czf = (float)zi - hzr;
        for (int k = b; k < e; k++) {
          float4 v = s_rel[k];
          v.x = __fmaf_rn(cxf, wxf, v.x); v.y = __fmaf_rn(cyf, wyf, v.y); v.z = __fmaf_rn(czf, wzf, v.z);
          s_rel[k] = v;
        }
        if (rowint && zi >= 1 && zi <= nbz && e > b) {
          const int n = e - b;
          atomicAdd(&s_n[(colxy | ((parz + zi) & 1)) * 3 + (n >= 3 ? 0 : (n == 2 ? 1 : 2))], 1);
        }
        b = e;
      }
    }
    __syncthreads();
    if (tid == 0) {
      int s = 0;
      for (int k = 0; k < 24; k++) { s_ibase[k] = s; s_fill[k] = s; s += s_n[k]; }
      s_ibase[24] = s;
    }
    __syncthreads();
    // ---- item lists: per colour, cells with >= 3 particles first, then 2, then 1 --------
#pragma unroll 1
    for (int idx = tid; idx < 2 * nbx * nby; idx += BLK_THREADS) {
      const int ri = idx >> 1, half = idx & 1;
      const int rx = ri / nby + 1, ry = ri - (rx - 1) * nby + 1;
      const int zlo = half ? (nbz >> 1) + 1 : 1, zhi = half ? nbz + 1 : (nbz >> 1) + 1;
      const int colxy = (((parx + rx) & 1) << 2) | (((pary + ry) & 1) << 1);
      const unsigned short* cz = s_cz + (rx * nry + ry) * czs;
      int b = cz[zlo];
      for (int zi = zlo; zi < zhi; zi++) {
        const int e = cz[zi + 1];
        const int n = e - b;
        if (n > 0) {
          const int p = atomicAdd(&s_fill[(colxy | ((parz + zi) & 1)) * 3 + (n >= 3 ? 0 : (n == 2 ? 1 : 2))], 1);
          s_items[p] = (unsigned short)((rx << 11) | (ry << 6) | zi);
        }
        b = e;
      }
    }
    __syncthreads();

    // ---- trials -------------------------------------------------------------------------
    const float lo = 1.0f - a.eps, hi = 1.0f + a.eps;
    struct Cur {
      int rxc, ryc, rz, n, sel, gslot, gx, iy, iz, last_id;
      long long gcell;
    };
    auto decode = [&](int item, int last_id) {
      Cur c;
      const int code = s_items[item];
      c.rxc = code >> 11; c.ryc = (code >> 6) & 31; c.rz = code & 63;
      const int rowc = c.rxc * nry + c.ryc;
      const unsigned short* cpc = s_cz + rowc * czs + c.rz;
      const int ob = cpc[0], oe = cpc[1];
      c.n = oe - ob;
      int l = x0 + c.rxc;
      c.iy = y0 + c.ryc; c.iz = z0 + c.rz;
      // interior cells never wrap: the block lies inside [0, n) on every axis
      c.gx = (g.gx0 + l >= g.nx) ? g.gx0 + l - g.nx : g.gx0 + l;
      c.gcell = ((long long)c.gx * g.ny + c.iy) * g.nz + c.iz;
      // particle of this trial: ascending id inside the cell
      c.sel = ob;
      c.last_id = last_id;
      if (c.n > 1) {
        int best = 0x7fffffff;
        for (int k = ob; k < oe; k++) {
          int id = __float_as_int(s_rel[k].w);
          if (id > last_id && id < best) { best = id; c.sel = k; }
        }
        c.last_id = best;
      }
      const TileRow rwc = s_row[rowc];
      const int ro = c.sel - rwc.off;
      c.gslot = (ro < rwc.cntA) ? rwc.gbA + ro : rwc.gbB + ro - rwc.cntA;
      return c;
    };
#pragma unroll 1
    for (int col = 0; col < 8; col++) {
      const int ib = s_ibase[3 * col], nit = s_ibase[3 * col + 3] - ib;
      // One flat loop: every iteration is exactly one trial (item `it`, trial index `j`).
      // Lanes start on item `tid` (deepest cells first) and fetch further items from a
      // shared ticket; the master-table entry of the NEXT trial's particle is requested one
      // iteration ahead (its latency is as long as a stencil scan).
      int it = tid, j = 0;
      Cur cur;
      double4 p_next = make_double4(0, 0, 0, 0);
      if (it < nit) { cur = decode(ib + it, -1); p_next = pos[cur.gslot]; }
#pragma unroll 1
      while (it < nit) {
        const double4 p = p_next;
        const Cur c = cur;
        int it2 = it, j2 = j + 1;
        if (j2 >= c.n) { j2 = 0; it2 = atomicAdd(&s_next[col], 1); }
        if (it2 < nit) { cur = decode(ib + it2, j2 ? c.last_id : -1); p_next = pos[cur.gslot]; }
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
            const float4 nrel = make_rel(g, gx, iy, iz, xn, yn, zn, p.w);
            const float tx = __fmaf_rn((float)rxc - hxr, wxf, nrel.x);
            const float ty = __fmaf_rn((float)ryc - hyr, wyf, nrel.y);
            const float tz = __fmaf_rn((float)rz - hzr, wzf, nrel.z);
            const float4 keep = s_rel[sel];
            s_rel[sel] = make_float4(BLK_FAR, BLK_FAR, BLK_FAR, keep.w);     // hide the trial particle from its own scan
            float r2min = 3.0e38f;
#pragma unroll 1
            for (int dx = -1; dx <= 1; dx++) {
#pragma unroll
              for (int dy = -1; dy <= 1; dy++) {
                const unsigned short* cp = s_cz + ((rxc + dx) * nry + ryc + dy) * czs + rz;
                const int b = cp[-1], e = cp[2];
#pragma unroll 1
                for (int k0 = b; k0 < e; k0 += BLK_SLOTS) {
                  const float4* q = s_rel + k0;
#pragma unroll
                  for (int s = 0; s < BLK_SLOTS; s++) {
                    const float4 qv = q[s];
                    const float ddx = tx - qv.x, ddy = ty - qv.y, ddz = tz - qv.z;
                    r2min = fminf(r2min, __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, ddx * ddx)));
                  }
                }
              }
            }
            bool ov = r2min < lo;
            if (!ov && r2min <= hi)
              ov = block_exact_rescan(pos, s_row, s_cz, czs, nry, rxc, ryc, rz, sel, xn, yn, zn, a.box);
            if (ov) {
              verdict = 1; n_ov++;
              s_rel[sel] = keep;
            } else {
              verdict = 0; n_acc++;
              s_rel[sel] = make_float4(tx, ty, tz, keep.w);
              rel[gslot] = nrel;
              pos[gslot] = make_double4(xn, yn, zn, p.w);
            }
          }
          if (LOG) {
            unsigned long long s = atomicAdd(nlog, 1ull);
            if ((long long)s < logcap) {
              hsmc_gpu_trial tr;
              tr.seq = ((unsigned long long)(a.phase * 8 + col) << 56) | ((unsigned long long)gcell << 8) | (unsigned)j;
              tr.id = (int)p.w; tr.verdict = verdict;
              tr.raw[0] = rn.v[0]; tr.raw[1] = rn.v[1]; tr.raw[2] = rn.v[2]; tr.pad = 0;
              log[s] = tr;
            }
          }
        }
        it = it2; j = j2;
      }
      __syncthreads();
    }
  } else {
    // ---- staging capacity exceeded (unusually dense block) or ablation: global-memory path,
    //      same order of updates ------------------------------------------------------------
    const int ncell_b = nbx * nby * nbz;
#pragma unroll 1
    for (int col = 0; col < 8; col++) {
      for (int q = tid; q < ncell_b; q += BLK_THREADS) {
        const int qz = q % nbz, qy = (q / nbz) % nby, qx = q / (nbz * nby);
        const int l = xa + qx, iy = ya + qy, iz = za + qz;
        const int c = (((g.gx0 + l) & 1) << 2) | ((iy & 1) << 1) | (iz & 1);
        if (c == col)
          cell_update_global_noinline<LOG>(a, a.phase * 8 + col, pos, rel, cs, l, iy, iz, 0, 1 << 30, n_acc, n_ov, n_cell,
                                           log, nlog, logcap);
      }
      __syncthreads();
    }
  }

  // ---- counters: warp reduce, then straight to the global counters ----------------------
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  n_ov = __reduce_add_sync(0xffffffffu, n_ov);
  n_cell = __reduce_add_sync(0xffffffffu, n_cell);
  if (lane == 0 && (n_acc | n_ov | n_cell)) {
    atomicAdd(&cnt[CNT_TRIALS], (unsigned long long)(n_acc + n_ov + n_cell));
    if (n_acc) atomicAdd(&cnt[CNT_ACC], (unsigned long long)n_acc);
    if (n_ov) atomicAdd(&cnt[CNT_REJ_OVERLAP], (unsigned long long)n_ov);
    if (n_cell) atomicAdd(&cnt[CNT_REJ_CELL], (unsigned long long)n_cell);
  }
}
