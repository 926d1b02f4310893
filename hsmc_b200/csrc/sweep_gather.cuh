// sweep_gather.cuh -- K2, experimental variant for large systems (sweep_impl 7): one launch per (cell colour, trial
// index), one THREAD per trial, the 27-cell stencil gathered straight from the cell-ordered fp32 shadow through L1/L2.
//
// Once the proposals of a sweep are generated up front (k_propose), a trial is a pure function of its record and of
// the current neighbour positions, so a cell colour is millions of independent trials: a compact list of the colour's
// trials (built by k_propose), a grid-stride loop, no shared memory, no barriers.  The price is that every colour
// phase streams the shadow table through L2 again (8 x 350 MB per sweep at N = 17 M) and that every thread walks its
// own nine short rows (little memory-level parallelism per thread).  Measured (profiles/r02_gather_*): slower than the
// block-resident kernel at N = 17 M, so it is NOT the default; it stays as the second implementation of the
// cell-colour chain (with k_sweep_phase, sweep_impl 1, the all-double reference of that chain).
//
// Order of updates: cell colours 0..7 (x parity slowest, so that slabs exchange their boundary layer twice per
// sweep), inside a colour every cell's trials in ascending particle id: launch (colour, 0) runs the first trial of
// every non-empty cell, (colour, 1) the second, (colour, 2) the third, and (colour, 3) -- one thread per cell -- the
// rest.  That is the chain of k_sweep_phase, bit for bit.
//
// Exactness as everywhere: fp32 minimum r^2 as a filter with error band eps, exact double re-evaluation
// (moves.c:400-431) from the master table for the rare trials inside the band.
#pragma once

#define GATHER_THREADS 256
#define GATHER_LISTS GATHER_LISTS_N

// GatherLists: see hsmc_gpu.cu

// one trial: record `rec` of the particle in slot `sel` of local cell (l, iy, iz); returns the verdict
// (0 accepted, 1 overlap, 2 left its cell)
__device__ __forceinline__ int gather_trial(const SweepArgs& a, double4* pos, float4* rel, const double4* __restrict__ prop,
                                            const int* __restrict__ cs, int l, int iy, int iz, int sel, const uint4 rec) {
  const Grid& g = a.g;
  if (!(rec.w & TREC_ACT)) return 2;
  const float tox = __uint_as_float(rec.x), toy = __uint_as_float(rec.y), toz = __uint_as_float(rec.z);
  const float wxf = (float)g.wx, wyf = (float)g.wy, wzf = (float)g.wz;
  // the z range of a stencil column: one slot range [za, za + na cells) and, when the column wraps around the box,
  // a second one
  const int za = (iz == 0) ? 0 : iz - 1, na = (iz == 0 || iz + 1 == g.nz) ? 2 : 3;
  const int zb = (iz == 0) ? g.nz - 1 : 0;
  const bool two = iz == 0 || iz + 1 == g.nz;
  // first all the range ends (18 independent loads), then the rows
  int kb[9], ke[9];
#pragma unroll
  for (int r = 0; r < 9; r++) {
    const int dx = r / 3 - 1, dy = r % 3 - 1;
    int ll = l + dx;
    if (g.wrap_x) { if (ll < 0) ll += g.nlx; else if (ll >= g.nlx) ll -= g.nlx; }
    int yy = iy + dy;
    if (yy < 0) yy += g.ny; else if (yy >= g.ny) yy -= g.ny;
    const int* row = cs + ((long long)ll * g.ny + yy) * g.nz;
    kb[r] = row[za]; ke[r] = row[za + na];
  }
  float r2min = 3.0e38f;
  auto scan = [&](int k0, int k1, float tx, float ty) {
#pragma unroll 1
    for (int k = k0; k < k1; k += 4) {
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; u++) q[u] = rel[min(k + u, k1 - 1)];       // (plain loads: the tail launch re-reads slots this thread wrote)
#pragma unroll
      for (int u = 0; u < 4; u++) {
        int dzc = __float_as_int(q[u].w) - iz;
        if (dzc > 1) dzc -= g.nz; else if (dzc < -1) dzc += g.nz;
        const float ddx = tx - q[u].x, ddy = ty - q[u].y, ddz = (toz - q[u].z) - (float)dzc * wzf;
        const float r2 = __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, ddx * ddx));
        if (min(k + u, k1 - 1) != sel) r2min = fminf(r2min, r2);
      }
    }
  };
#pragma unroll
  for (int r = 0; r < 9; r++) scan(kb[r], ke[r], tox - (float)(r / 3 - 1) * wxf, toy - (float)(r % 3 - 1) * wyf);
  if (two) {                       // the wrapped cell of every column (cells at the periodic z edge only)
#pragma unroll 1
    for (int r = 0; r < 9; r++) {
      const int dx = r / 3 - 1, dy = r % 3 - 1;
      int ll = l + dx;
      if (g.wrap_x) { if (ll < 0) ll += g.nlx; else if (ll >= g.nlx) ll -= g.nlx; }
      int yy = iy + dy;
      if (yy < 0) yy += g.ny; else if (yy >= g.ny) yy -= g.ny;
      const int* row = cs + ((long long)ll * g.ny + yy) * g.nz;
      scan(row[zb], row[zb + 1], tox - (float)dx * wxf, toy - (float)dy * wyf);
    }
  }
  bool ov = r2min < 1.0f - a.eps;
  double4 pr = make_double4(0, 0, 0, 0);
  if (!ov) {
    pr = prop[sel];
    if (r2min <= 1.0f + a.eps) {
      const Box& b = a.box;
      ov = stencil_any(g, cs, l, iy, iz, [&](int k) {
        if (k == sel) return false;
        const double4 q = pos[k];
        return pair_r2(pr.x, pr.y, pr.z, q.x, q.y, q.z, b) < 1.0;
      });
    }
  }
  if (ov) return 1;
  double* pd = reinterpret_cast<double*>(pos + sel);
  *reinterpret_cast<double2*>(pd) = make_double2(pr.x, pr.y);
  pd[2] = pr.z;
  float* rl = reinterpret_cast<float*>(rel + sel);
  *reinterpret_cast<float2*>(rl) = make_float2(tox, toy);
  rl[2] = toz;
  return 0;
}

template <bool LOG>
__global__ void __launch_bounds__(GATHER_THREADS, 4)
k_sweep_gather(SweepArgs a, GatherLists gl, int colour, int jj, double4* pos, float4* rel,
               const double4* __restrict__ prop, const uint4* __restrict__ trec, const uint4* __restrict__ traw,
               const int* __restrict__ cs, unsigned long long* __restrict__ cnt, hsmc_gpu_trial* __restrict__ log,
               unsigned long long* __restrict__ nlog, long long logcap) {
  const Grid& g = a.g;
  const int li = colour * 4 + jj;
  const int n_items = min(gl.count[li], (int)gl.stride);
  const int* list = gl.list + (long long)li * gl.stride;
  int n_acc = 0, n_ov = 0, n_cell = 0;
  for (int i = blockIdx.x * GATHER_THREADS + threadIdx.x; i < n_items; i += gridDim.x * GATHER_THREADS) {
    const int e = list[i];
    const int l = e >> 20, iy = (e >> 10) & 1023, iz = e & 1023;
    const long long c = ((long long)l * g.ny + iy) * g.nz + iz;
    const int b = cs[c];
    // (colour, 0..2): that one trial; (colour, 3): the fourth and later trials of the cell, in order
    const int j0 = jj, j1 = (jj < 3) ? jj + 1 : cs[c + 1] - b;
    for (int j = j0; j < j1; j++) {
      const uint4 rec = __ldg(trec + b + j);
      const int sel = b + (int)(rec.w & 15);
      const int verdict = gather_trial(a, pos, rel, prop, cs, l, iy, iz, sel, rec);
      n_acc += verdict == 0; n_ov += verdict == 1; n_cell += verdict == 2;
      if (LOG) {
        const uint4 rw4 = __ldg(traw + b + j);
        const long long gcell = global_cell_of_local(g, l, iy, iz);
        const unsigned long long sl = atomicAdd(nlog, 1ull);
        if ((long long)sl < logcap) {
          hsmc_gpu_trial tr;
          tr.seq = ((unsigned long long)colour << 56) | ((unsigned long long)gcell << 8) | (unsigned)j;
          tr.id = (int)rw4.w; tr.verdict = verdict;
          tr.raw[0] = rw4.x; tr.raw[1] = rw4.y; tr.raw[2] = rw4.z; tr.pad = 0;
          log[sl] = tr;
        }
      }
    }
  }
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  n_ov = __reduce_add_sync(0xffffffffu, n_ov);
  n_cell = __reduce_add_sync(0xffffffffu, n_cell);
  if ((threadIdx.x & 31) == 0 && (n_acc | n_ov | n_cell)) {
    atomicAdd(&cnt[CNT_TRIALS], (unsigned long long)(n_acc + n_ov + n_cell));
    if (n_acc) atomicAdd(&cnt[CNT_ACC], (unsigned long long)n_acc);
    if (n_ov) atomicAdd(&cnt[CNT_REJ_OVERLAP], (unsigned long long)n_ov);
    if (n_cell) atomicAdd(&cnt[CNT_REJ_CELL], (unsigned long long)n_cell);
  }
}
