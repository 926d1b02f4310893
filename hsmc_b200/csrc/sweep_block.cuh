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
#ifndef BLK_SLOTS
#define BLK_SLOTS 7
#endif
#define BLK_PAD 8               // far-away entries after the last staged particle
#define BLK_FAR 1.0e15f
#define BLK_MAX_OCC 15           // most particles per cell the trial-slot scheme handles (denser block: global path)

// BlockCfg: see hsmc_gpu.cu (the handle keeps one)

// BlockRow: see hsmc_gpu.cu

// exact re-evaluation of a whole stencil from the master table (moves.c:157-212, 400-431)
// (`pos` deliberately not const __restrict__: entries of this block were written earlier in this
// launch by other threads of the CTA, so the loads must not take the non-coherent path)
__device__ __noinline__ bool block_exact_rescan(const double4* pos, const BlockRow* s_row,
                                                const unsigned short* s_cz, int cz_stride, int nry, int rxc,
                                                int ryc, int rz, int sel, double xn, double yn, double zn,
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
        const double4 q = pos[gs];
        if (pair_r2(xn, yn, zn, q.x, q.y, q.z, box) < 1.0) return true;
      }
    }
  return false;
}

// One (x,y) row of a trial's stencil: the three z-cells are one contiguous staged range,
// scanned in groups of BLK_SLOTS entries at fixed offsets, unmasked (see the file header).
// NT = 2 keeps the running minimum for two trial points at once (the two trials of a unit).
template <int NT>
__device__ __forceinline__ void block_scan_row(const float4* __restrict__ s_rel, int b, int e, float x0, float y0,
                                               float z0, float x1, float y1, float z1, float& r0, float& r1) {
#pragma unroll 1
  for (int k0 = b; k0 < e; k0 += BLK_SLOTS) {
    const float4* q = s_rel + k0;
#pragma unroll
    for (int s = 0; s < BLK_SLOTS; s++) {
      const float4 qv = q[s];
      const float ax = x0 - qv.x, ay = y0 - qv.y, az = z0 - qv.z;
      r0 = fminf(r0, __fmaf_rn(az, az, __fmaf_rn(ay, ay, ax * ax)));
      if (NT == 2) {
        const float bx = x1 - qv.x, by = y1 - qv.y, bz = z1 - qv.z;
        r1 = fminf(r1, __fmaf_rn(bz, bz, __fmaf_rn(by, by, bx * bx)));
      }
    }
  }
}

// slot-group class of a cell with n particles: groups of 16, 8, 4, 2, 1 slots
__device__ __forceinline__ int occ_class(int n) { return n > 8 ? 0 : (n > 4 ? 1 : (n > 2 ? 2 : (n == 2 ? 3 : 4))); }

// per-CTA cycle accounting for tuning (dbg == 10): accumulated clock64() deltas of thread 0
__device__ unsigned long long g_blk_t[16];
#define BLK_MARK(k)                                                         \
  if (bc.dbg == 10 && threadIdx.x == 0) {                                   \
    const long long t_now = clock64();                                      \
    atomicAdd(&g_blk_t[k], (unsigned long long)(t_now - t_mark));           \
    t_mark = t_now;                                                         \
  }

template <bool LOG>
__global__ void __launch_bounds__(BLK_THREADS, BLK_MIN_CTAS)
k_sweep_block(SweepArgs a, BlockCfg bc, const int* __restrict__ xoff, double4* __restrict__ pos,
              float4* __restrict__ rel, const int* __restrict__ cs, const unsigned short* __restrict__ cs16,
              unsigned long long* __restrict__ cnt, hsmc_gpu_trial* __restrict__ log,
              unsigned long long* __restrict__ nlog, long long logcap) {
  const Grid& g = a.g;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* s_rel = reinterpret_cast<float4*>(smem_raw);
  unsigned short* s_cz = reinterpret_cast<unsigned short*>(s_rel + bc.cap);    // [rows][cz_stride] staged index of each region cell
  unsigned int* s_trials = reinterpret_cast<unsigned int*>(s_cz + bc.max_rows * bc.cz_stride);   // [8][tr_cap] trial slots per colour: cell code | j << 16 | n << 20
  __shared__ BlockRow s_row[BLK_MAX_ROWS];
  __shared__ int s_cnt[BLK_MAX_ROWS + 1];
  __shared__ int s_n[40], s_cbase[40], s_fill[40], s_next[8], s_ntr[8], s_flag, s_done_idx;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = BLK_THREADS / 32;
  const int czs = bc.cz_stride;
  long long t_mark = clock64();

  // ---- which block ------------------------------------------------------------------
  // a.fuse <= 1: one launch = one block phase (a.phase), CTA = blockIdx.x.
  // a.fuse  > 1: one launch = the block phases [a.phase, a.phase + a.fuse).  CTAs draw a ticket;
  // tickets enumerate (phase, block) in phase order, so whenever a CTA holds ticket t, every
  // ticket < t is held by a CTA that is resident or finished.  A block of phase p only
  // depends on its (up to 26) neighbouring blocks of EARLIER phases of the launch: it waits
  // for their completion flags (release store / relaxed poll + acquire fence at gpu scope) instead of for a kernel
  // boundary, so the ragged last wave of one phase overlaps the first wave of the next.
  // The earliest unfinished phase never waits => no deadlock.  Same chain as separate launches.
  const int hbz = bc.nbz >> 1, hby = bc.nby >> 1, hbx = bc.nbx >> 1;
  int ph = a.phase, bid = blockIdx.x;
  if (a.fuse > 1) {
    if (tid == 0) s_flag = (int)(atomicAdd(bc.ticket, 1u) - a.ticket_base);
    __syncthreads();
    const int per = hbx * hby * hbz, t = s_flag;
    ph = a.phase + t / per;
    bid = t - (t / per) * per;
    __syncthreads();                      // s_flag is reused below
  }
  const int pcx = (ph >> 2) & 1, pcy = (ph >> 1) & 1, pcz = ph & 1;
  const int bzi = 2 * (bid % hbz) + pcz;
  const int byi = 2 * ((bid / hbz) % hby) + pcy;
  const int bxi = 2 * (bid / (hbz * hby)) + pcx;
  if (a.fuse > 1) {
    if (tid == 0) s_done_idx = (bxi * bc.nby + byi) * bc.nbz + bzi;
    if (tid < 27 && tid != 13) {
      int nx = bxi + tid / 9 - 1, ny = byi + (tid / 3) % 3 - 1, nz = bzi + tid % 3 - 1;
      bool have = true;
      if (g.wrap_x) { if (nx < 0) nx += bc.nbx; else if (nx >= bc.nbx) nx -= bc.nbx; }
      else have = nx >= 0 && nx < bc.nbx;          // slab edge: that neighbour lives on another rank (halo exchange)
      if (ny < 0) ny += bc.nby; else if (ny >= bc.nby) ny -= bc.nby;
      if (nz < 0) nz += bc.nbz; else if (nz >= bc.nbz) nz -= bc.nbz;
      const int q = ((nx & 1) << 2) | ((ny & 1) << 1) | (nz & 1);
      if (have && q >= a.phase && q < ph) {
        const unsigned int* f = bc.done + ((size_t)nx * bc.nby + ny) * bc.nbz + nz;
        // (bounded: a protocol error must surface as a CUDA error, not as a GPU that spins for ever;
        //  a whole launch lasts milliseconds, the limit is a minute)
        unsigned long long t0 = 0;
        unsigned int spins = 0;
        while (ld_relaxed_gpu(f) != a.epoch) {
          __nanosleep(100);
          if ((++spins & 4095u) == 0) {
            const unsigned long long t = hsmc_globaltimer_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 60ull * 1000000000ull) __trap();
          }
        }
      }
      __threadfence();                     // acquire: everything those blocks wrote is visible from here on
    }
    __syncthreads();
  }
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
  if (tid < 40) s_n[tid] = 0;
  if (tid < 8) s_next[tid] = BLK_THREADS / 32;
  if (tid == 0) s_flag = 0;

  // ---- row pieces: row r belongs to thread (lane, warp) with r = lane*NW + warp, so every
  //      warp issues the same number of (serialised) TMA copies
  long long my_row = 0;
  int my_cntB = 0;
  const int myrow = lane * NW + warp;
  if (myrow < nrows) {
    int rx = myrow / nry, ry = myrow - rx * nry;
    int lx = x0 + rx;
    if (g.wrap_x) { if (lx < 0) lx += g.nlx; else if (lx >= g.nlx) lx -= g.nlx; }
    int y = y0 + ry;
    if (y < 0) y += g.ny; else if (y >= g.ny) y -= g.ny;
    my_row = (long long)lx * g.ny + y;
    const long long rbase = my_row * g.nz;
    int gbA = cs[rbase + zs], geA = cs[rbase + min(zs + lenz, g.nz)];
    int gbB = 0, geB = 0;
    if (zwrap) { gbB = cs[rbase]; geB = cs[rbase + (zs + lenz - g.nz)]; }
    my_cntB = geB - gbB;
    s_row[myrow].gbA = gbA; s_row[myrow].gbB = gbB; s_row[myrow].cntA = geA - gbA;
    s_cnt[myrow] = (geA - gbA) + my_cntB;
  }
  __syncthreads();
  BLK_MARK(0)     // row ends read
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
  BLK_MARK(1)     // row scan
  if (bc.dbg == 1) return;
  const int total = s_cnt[nrows];
  const bool staged = !bc.force_global && total + BLK_PAD <= bc.cap;

  int n_acc = 0, n_ov = 0, n_cell = 0;
  bool ok = false;     // staged, and every interior cell / colour fits the trial-slot scheme

  if (staged) {
    // ---- stage the shadow rows --------------------------------------------------------------
    // Default: per-thread 16-byte async copies (cp.async / LDGSTS), one per particle, all in
    // flight at once -- the rows are short (~15 particles), and bulk-TMA copies that small are
    // bound by the per-copy cost of the TMA engine (measured: ~50 cycles of engine time per
    // copy, profiles/).  use_tma (ablation) issues one bulk copy per row piece instead.
    if (myrow < nrows) {
      BlockRow& rw = s_row[myrow];
      const int off = s_cnt[myrow], cA = rw.cntA;
      rw.off = off;
      if (bc.use_tma) {
        uint32_t bytes = (uint32_t)(cA + my_cntB) * 16u;
        if (bytes) mbar_arrive_tx(&s_bar, bytes); else mbar_arrive(&s_bar);
        if (cA) tma_bulk_g2s(&s_rel[off], &rel[rw.gbA], (uint32_t)cA * 16u, &s_bar);
        if (my_cntB) tma_bulk_g2s(&s_rel[off + cA], &rel[rw.gbB], (uint32_t)my_cntB * 16u, &s_bar);
      } else {
        // the row's own thread streams it in: back-to-back independent 16-byte async copies
        const float4* srcA = rel + rw.gbA;
        float4* dst = s_rel + off;
#pragma unroll 4
        for (int k = 0; k < cA; k++) cp_async16(dst + k, srcA + k);
        const float4* srcB = rel + rw.gbB;
        for (int k = 0; k < my_cntB; k++) cp_async16(dst + cA + k, srcB + k);
      }
      // staged index of the first particle of every cell of the row, from the 16-bit
      // row-relative CSR (a wrapped row continues at z = 0 of the same global row)
      const unsigned short* crow = cs16 + my_row * (g.nz + 1);
      unsigned short* out = s_cz + myrow * czs;
      const int v0 = crow[zs];
#pragma unroll 4
      for (int zi = 0; zi <= lenz; zi++) {
        const int z = zs + zi;
        const int v = (z <= g.nz) ? off + (int)crow[z] - v0 : off + cA + (int)crow[z - g.nz];
        out[zi] = (unsigned short)v;
      }
    } else if (bc.use_tma) {
      mbar_arrive(&s_bar);
    }
    if (tid < BLK_PAD) s_rel[total + tid] = make_float4(BLK_FAR, BLK_FAR, BLK_FAR, __int_as_float(-1));
    BLK_MARK(2)     // copies issued, CSR rows done
    if (bc.use_tma) mbar_wait(&s_bar, 0); else cp_async_wait_all();
    if (bc.dbg == 2) return;
    __syncthreads();
    BLK_MARK(3)     // staged data landed

    // ---- block-relative coordinates; census of the interior cells by colour ----------------
    const float wxf = (float)g.wx, wyf = (float)g.wy, wzf = (float)g.wz;
    const float hxr = 0.5f * (float)nrx, hyr = 0.5f * (float)nry, hzr = 0.5f * (float)lenz;
    const int parx = (g.gx0 + x0) & 1, pary = y0 & 1, parz = z0 & 1;    // parity of region cell (0,0,0); grids are even
    // each staged row is cut into P z-pieces, one (row, piece) item per thread and round; P is the
    // split that wastes the fewest thread-rounds (100 rows x 2 halves on 128 threads would leave the
    // second round 44 % empty)
    int P = 2;
    {
      int best = 1 << 30;
      for (int q = 2; q <= 8; q++) {
        const int cost = ((q * nrows + BLK_THREADS - 1) / BLK_THREADS) * ((lenz + q - 1) / q + 2);
        if (cost < best) { best = cost; P = q; }
      }
    }
#pragma unroll 1
    for (int idx = tid; idx < P * nrows; idx += BLK_THREADS) {
      const int r = idx / P, piece = idx - r * P;
      const int rx = r / nry, ry = r - rx * nry;
      const int zlo = (piece * lenz) / P, zhi = ((piece + 1) * lenz) / P;
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
          if (n > BLK_MAX_OCC) s_flag = 1;
          else atomicAdd(&s_n[(colxy | ((parz + zi) & 1)) * 5 + occ_class(n)], 1);
        }
        b = e;
      }
    }
    __syncthreads();
    BLK_MARK(4)     // block-relative conversion + census
    // ---- trial slots: every particle of a colour gets one lane-slot.  The slots of a cell are a
    //      group of 1, 2, 4, 8 or 16 consecutive slots (by occupancy), groups are laid out
    //      largest first, so a group never straddles a warp-sized chunk: the trials of one cell
    //      always sit in adjacent lanes of one warp.  Unused slots of a group hold 0xffffffff.
    if (tid < 8) {
      int base = 0;
      for (int c = 0; c < 5; c++) {
        s_cbase[tid * 5 + c] = base;
        s_fill[tid * 5 + c] = 0;
        base += s_n[tid * 5 + c] << (4 - c);
      }
      s_ntr[tid] = base;
      if (base > bc.tr_cap) s_flag = 1;
    }
    __syncthreads();
    if (s_flag == 0) {
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
            const int col = colxy | ((parz + zi) & 1), c = occ_class(n), G = 1 << (4 - c);
            const int q = atomicAdd(&s_fill[col * 5 + c], 1);
            unsigned int* tl = s_trials + col * bc.tr_cap + s_cbase[col * 5 + c] + q * G;
            const unsigned int code = (unsigned)((rx << 11) | (ry << 6) | zi) | ((unsigned)n << 20);
            for (int j = 0; j < G; j++) tl[j] = (j < n) ? (code | ((unsigned)j << 16)) : 0xffffffffu;
          }
          b = e;
        }
      }
    }
    __syncthreads();
    ok = s_flag == 0;
    BLK_MARK(5)     // trial slots built
    if (bc.dbg == 3) return;
  }
  if (ok) {
    // ---- trials -------------------------------------------------------------------------
    // One lane per trial; a warp takes a chunk of 32 trial slots at a time.  Every particle of
    // the cells in the chunk is hidden from the staged table while the chunk is scanned, so the
    // fp32 scan only sees particles that do not move during the chunk; the pairs INSIDE a cell
    // are then resolved exactly, in double, in ascending-id order with warp shuffles: trial j
    // is tested against its mates i > j at their old positions and against its mates i < j at
    // the positions their own trials left them in.
    const unsigned FULL = 0xffffffffu;
    const float wxf = (float)g.wx, wyf = (float)g.wy, wzf = (float)g.wz;
    const float hxr = 0.5f * (float)nrx, hyr = 0.5f * (float)nry, hzr = 0.5f * (float)lenz;
    const float lo = 1.0f - a.eps, hi = 1.0f + a.eps;
    struct Slot { int code, cell, sel, gs; };       // code < 0: padding / beyond the end
    auto stage_a = [&](int col, int chunk) {
      Slot u;
      const int t = chunk * 32 + lane;
      u.code = -1; u.cell = 0; u.sel = 0; u.gs = 0;
      if (t < s_ntr[col]) {
        const unsigned int c = s_trials[col * bc.tr_cap + t];
        if (c != 0xffffffffu) {
          u.code = (int)(c >> 16);                 // j | n << 4
          u.cell = (int)(c & 0xffffu);
          const int j = u.code & 15, n = u.code >> 4;
          const int rowc = (u.cell >> 11) * nry + ((u.cell >> 6) & 31);
          const int ob = s_cz[rowc * czs + (u.cell & 63)];
          // the particle with exactly j smaller ids in its cell
          u.sel = ob;
          if (n > 1) {
            for (int k = ob; k < ob + n; k++) {
              const int idk = __float_as_int(s_rel[k].w);
              int c2 = 0;
              for (int m = ob; m < ob + n; m++) c2 += __float_as_int(s_rel[m].w) < idk;
              if (c2 == j) u.sel = k;
            }
          }
          const BlockRow rwc = s_row[rowc];
          const int ro = u.sel - rwc.off;
          u.gs = (ro < rwc.cntA) ? rwc.gbA + ro : rwc.gbB + ro - rwc.cntA;
          prefetch_l1(&pos[u.gs]);
        }
      }
      return u;
    };
#pragma unroll 1
    for (int col = 0; col < 8; col++) {
      const int ntr = s_ntr[col];
      int chunk = warp;                            // warp-uniform
      Slot nx;
      if (chunk * 32 < ntr) nx = stage_a(col, chunk);
#pragma unroll 1
      while (chunk * 32 < ntr) {
        const Slot u = nx;
        const bool valid = u.code >= 0;
        const int j = u.code & 15, n = valid ? (u.code >> 4) : 0;
        BLK_MARK(14)    // loop top (residual)
        double4 p = make_double4(0, 0, 0, 0);
        if (valid) p = pos[u.gs];
        // next chunk of this warp: a ticket, and the early request for its master entries
        int nxt = 0;
        if (lane == 0) nxt = atomicAdd(&s_next[col], 1);
        nxt = __shfl_sync(FULL, nxt, 0);
        if (nxt * 32 < ntr) nx = stage_a(col, nxt);
        BLK_MARK(8)     // ticket + stage_a of the next chunk
        const int rxc = u.cell >> 11, ryc = (u.cell >> 6) & 31, rz = u.cell & 63;
        const int iy = y0 + ryc, iz = z0 + rz;
        const int gxl = g.gx0 + x0 + rxc;
        const int gx = (gxl >= g.nx) ? gxl - g.nx : gxl;          // interior cells never wrap inside the block
        const long long gcell = ((long long)gx * g.ny + iy) * g.nz + iz;
        // ---- the trial point (moves.c:52-57, 215-226) ----
        Philox4 rn;
        double xn = 0, yn = 0, zn = 0;
        bool act = false;
        float4 nrel = make_float4(0.f, 0.f, 0.f, 0.f), keep = nrel;
        float tx = 0.f, ty = 0.f, tz = 0.f;
        if (valid) {
          rn = philox4x32_10((uint32_t)gcell, (HSMC_STREAM_MOVE << 24) | (uint32_t)j, a.sweep_lo, a.sweep_hi, a.key0, a.key1);
          xn = p.x + (hsmc_u01(rn.v[0]) - 0.5) * a.dr_max;
          yn = p.y + (hsmc_u01(rn.v[1]) - 0.5) * a.dr_max;
          zn = p.z + (hsmc_u01(rn.v[2]) - 0.5) * a.dr_max;
          if (xn > g.Lx) xn -= g.Lx; else if (xn < 0.0) xn += g.Lx;
          if (yn > g.Ly) yn -= g.Ly; else if (yn < 0.0) yn += g.Ly;
          if (zn > g.Lz) zn -= g.Lz; else if (zn < 0.0) zn += g.Lz;
          act = axis_cell(xn, g.sx, g.iwx, g.nx) == gx && axis_cell(yn, g.sy, g.iwy, g.ny) == iy &&
                axis_cell(zn, g.sz, g.iwz, g.nz) == iz;
          if (act) {
            nrel = make_rel(g, gx, iy, iz, xn, yn, zn, p.w);
            tx = __fmaf_rn((float)rxc - hxr, wxf, nrel.x);
            ty = __fmaf_rn((float)ryc - hyr, wyf, nrel.y);
            tz = __fmaf_rn((float)rz - hzr, wzf, nrel.z);
          } else n_cell++;
          keep = s_rel[u.sel];
          s_rel[u.sel] = make_float4(BLK_FAR, BLK_FAR, BLK_FAR, keep.w);     // hidden while the chunk is scanned
        }
        __syncwarp();
        BLK_MARK(9)     // master load + trial point + hide
        float r2min = 3.0e38f;
        if (act && bc.dbg != 4) {
          // first BLK_SLOTS entries of all nine rows in straight-line code (independent loads the
          // scheduler can overlap with the arithmetic of the previous row) ...
          const unsigned short* cp0 = s_cz + ((rxc - 1) * nry + ryc - 1) * czs + rz;
          bool more = false;
#pragma unroll
          for (int r = 0; r < 9; r++) {
            const unsigned short* cp = cp0 + ((r / 3) * nry + (r % 3)) * czs;
            const int b = cp[-1];
            more |= (int)cp[2] - b > BLK_SLOTS;
            const float4* q = s_rel + b;
#pragma unroll
            for (int s = 0; s < BLK_SLOTS; s++) {
              const float4 qv = q[s];
              const float ddx = tx - qv.x, ddy = ty - qv.y, ddz = tz - qv.z;
              r2min = fminf(r2min, __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, ddx * ddx)));
            }
          }
          // ... and the rare rows holding more than BLK_SLOTS particles
          if (more) {
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
              const unsigned short* cp = cp0 + ((r / 3) * nry + (r % 3)) * czs;
              const int e = cp[2];
#pragma unroll 1
              for (int k0 = cp[-1] + BLK_SLOTS; k0 < e; k0 += BLK_SLOTS) {
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
        }
        BLK_MARK(10)    // stencil scan
        // ---- pairs inside a cell, exactly (moves.c:400-431), and the verdicts in trial order ----
        const int maxn = __reduce_max_sync(FULL, n);
        const int gb = lane - j;                       // lane of the cell's first trial
        bool mate_ov = false;
        for (int s = 1; s < maxn; s++) {               // mates with a LATER trial: still at their old positions
          const int src = (gb + s) & 31;
          const double qx = __shfl_sync(FULL, p.x, src), qy = __shfl_sync(FULL, p.y, src), qz = __shfl_sync(FULL, p.z, src);
          if (act && j < s && s < n && !mate_ov) mate_ov = pair_r2(xn, yn, zn, qx, qy, qz, a.box) < 1.0;
        }
        double cx = p.x, cy = p.y, cz = p.z;           // where this lane's particle is after its own trial
        int verdict = 2;
        for (int s = 0; s < maxn; s++) {
          if (valid && j == s) {
            bool acc = false;
            if (act) {
              bool ov = mate_ov || r2min < lo;
              if (!ov && r2min <= hi)
                ov = block_exact_rescan(pos, s_row, s_cz, czs, nry, rxc, ryc, rz, u.sel, xn, yn, zn, a.box);
              if (ov) { verdict = 1; n_ov++; }
              else {
                verdict = 0; n_acc++; acc = true;
                cx = xn; cy = yn; cz = zn;
                rel[u.gs] = nrel;
                pos[u.gs] = make_double4(xn, yn, zn, p.w);
              }
            }
            s_rel[u.sel] = acc ? make_float4(tx, ty, tz, keep.w) : keep;
          }
          if (s + 1 < maxn) {
            __syncwarp();
            const int src = (gb + s) & 31;
            const double qx = __shfl_sync(FULL, cx, src), qy = __shfl_sync(FULL, cy, src), qz = __shfl_sync(FULL, cz, src);
            if (act && j > s && !mate_ov) mate_ov = pair_r2(xn, yn, zn, qx, qy, qz, a.box) < 1.0;
          }
        }
        BLK_MARK(11)    // mates + verdicts + commit
        if (LOG && valid) {
          unsigned long long s = atomicAdd(nlog, 1ull);
          if ((long long)s < logcap) {
            hsmc_gpu_trial tr;
            tr.seq = ((unsigned long long)(ph * 8 + col) << 56) | ((unsigned long long)gcell << 8) | (unsigned)j;
            tr.id = (int)p.w; tr.verdict = verdict;
            tr.raw[0] = rn.v[0]; tr.raw[1] = rn.v[1]; tr.raw[2] = rn.v[2]; tr.pad = 0;
            log[s] = tr;
          }
        }
        chunk = nxt;
      }
      BLK_MARK(6)     // own chunks of a colour done (warp 0)
      __syncthreads();
      BLK_MARK(7)     // waiting for the other warps at the colour barrier
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
          cell_update_global_noinline<LOG>(a, ph * 8 + col, pos, rel, cs, l, iy, iz, 0, 1 << 30, n_acc, n_ov, n_cell,
                                           log, nlog, logcap);
      }
      __syncthreads();
    }
  }

  // ---- counters: warp reduce, then straight to the global counters ----------------------
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  n_ov = __reduce_add_sync(0xffffffffu, n_ov);
  n_cell = __reduce_add_sync(0xffffffffu, n_cell);
  if (lane == 0 && (n_acc | n_ov | n_cell) && bc.dbg != 5) {
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
