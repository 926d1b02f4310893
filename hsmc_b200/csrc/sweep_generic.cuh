// sweep_generic.cuh -- K2 (generic form): SweepArgs and the one-thread-per-cell sweep from global memory; the staged kernel is sweep_lean.cuh
// (part of the single translation unit hsmc_gpu.cu; included there, in this order)
#pragma once

// ----------------------------------------------------------------------------------
// K2: one colour phase of the checkerboard sweep.
//
// Cells are coloured by the parity of their (global) indices, 2x2x2 = 8 colours.  Two
// cells of one colour are separated by a full cell (edge >= 1.0 = sigma), so particles
// in different active cells can never overlap whatever moves they make inside their
// cells: all active cells are independent and are processed concurrently, one thread
// per active cell, the particles of a cell sequentially in ascending-id order.  A trial
// that would leave its cell is rejected (membership is static within a sweep); the grid
// origin is redrawn between sweeps so that walls move (Anderson et al., J. Comput. Phys.
// 254 (2013) 27).  Each trial is the reference's part_move(): three uniforms,
// x += (u - 0.5)*dr_max, apply_pbc, accept iff check_overlap is false.
// ----------------------------------------------------------------------------------
struct SweepArgs {
  Grid g;
  Box box;
  double dr_max;
  uint32_t key0, key1;
  uint32_t sweep_lo, sweep_hi;
  int cx, cy, cz, phase;
  float eps;   // half-width of the fp32 filter's uncertainty band around r^2 = 1
  // k_sweep_block only: phases [phase, phase + fuse) in one launch (fuse <= 1: just `phase`)
  int fuse;
  unsigned int epoch, ticket_base;
};

// all trials of one active cell straight from global memory (generic path: any grid,
// minimum image always evaluated)
template <bool LOG>
__device__ __forceinline__ void cell_update_global(const SweepArgs& a, int phase, double4* __restrict__ pos,
                                                   float4* __restrict__ rel, const int* __restrict__ cs, int l,
                                                   int iy, int iz, int j0, int j1, int& n_acc,
                                                   int& n_ov, int& n_cell, hsmc_gpu_trial* __restrict__ log,
                                                   unsigned long long* __restrict__ nlog, long long logcap) {
  const Grid& g = a.g;
  long long c = ((long long)l * g.ny + iy) * g.nz + iz;
  int beg = cs[c], end = cs[c + 1];
  if (beg == end) return;
  if (j0 < 0) {                        // tile-kernel fallback: deep cells are not its business
    if (end - beg > j1) return;
    j0 = 0;
  }
  const int gx = (g.gx0 + l >= g.nx) ? g.gx0 + l - g.nx : g.gx0 + l;
  long long gcell = ((long long)gx * g.ny + iy) * g.nz + iz;
  double last_id = -1.0;
  for (int j = 0; j < end - beg; j++) {
    // next particle of this cell in ascending-id order (order is then independent of
    // how the counting sort happened to place them)
    int sel = beg;
    double best = 1e300;
    for (int k = beg; k < end; k++) {
      double id = pos[k].w;
      if (id > last_id && id < best) { best = id; sel = k; }
    }
    last_id = best;
    if (j < j0) continue;      // trials below j0 belong to the tile kernel
    if (j >= j1) break;
    double4 p = pos[sel];
    Philox4 rn = philox4x32_10((uint32_t)gcell, (HSMC_STREAM_MOVE << 24) | (uint32_t)j, a.sweep_lo, a.sweep_hi,
                               a.key0, a.key1);
    // moves.c:52-54
    double xn = p.x + (hsmc_u01(rn.v[0]) - 0.5) * a.dr_max;
    double yn = p.y + (hsmc_u01(rn.v[1]) - 0.5) * a.dr_max;
    double zn = p.z + (hsmc_u01(rn.v[2]) - 0.5) * a.dr_max;
    // moves.c:215-226
    if (xn > g.Lx) xn -= g.Lx; else if (xn < 0.0) xn += g.Lx;
    if (yn > g.Ly) yn -= g.Ly; else if (yn < 0.0) yn += g.Ly;
    if (zn > g.Lz) zn -= g.Lz; else if (zn < 0.0) zn += g.Lz;
    int verdict;
    if (axis_cell(xn, g.sx, g.iwx, g.nx) != gx || axis_cell(yn, g.sy, g.iwy, g.ny) != iy ||
        axis_cell(zn, g.sz, g.iwz, g.nz) != iz) {
      verdict = 2;
      n_cell++;
    } else {
      const Box& b = a.box;
      bool ov = stencil_any(g, cs, l, iy, iz, [&](int k) {
        if (k == sel) return false;
        double4 q = pos[k];
        return pair_r2(xn, yn, zn, q.x, q.y, q.z, b) < 1.0;
      });
      if (ov) { verdict = 1; n_ov++; }
      else {
        verdict = 0; n_acc++;
        pos[sel] = make_double4(xn, yn, zn, p.w);
        rel[sel] = make_rel(g, gx, iy, iz, xn, yn, zn);
      }
    }
    if (LOG) {
      unsigned long long s = atomicAdd(nlog, 1ull);
      if ((long long)s < logcap) {
        hsmc_gpu_trial tr;
        tr.seq = ((unsigned long long)phase << 56) | ((unsigned long long)gcell << 8) | (unsigned)j;
        tr.id = (int)p.w; tr.verdict = verdict;
        tr.raw[0] = rn.v[0]; tr.raw[1] = rn.v[1]; tr.raw[2] = rn.v[2]; tr.pad = 0;
        log[s] = tr;
      }
    }
  }
}

// generic kernel: one thread per active cell, everything from global memory
template <bool LOG>
__global__ void __launch_bounds__(128)
k_sweep_phase(SweepArgs a, double4* __restrict__ pos, float4* __restrict__ rel, const int* __restrict__ cs,
              unsigned long long* __restrict__ cnt, hsmc_gpu_trial* __restrict__ log,
              unsigned long long* __restrict__ nlog, long long logcap) {
  const Grid& g = a.g;
  const int hx = (g.own_hi - g.own_lo) >> 1, hy = g.ny >> 1, hz = g.nz >> 1;
  const long long total = (long long)hx * hy * hz;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int n_acc = 0, n_ov = 0, n_cell = 0;
  if (t < total) {
    int az = (int)(t % hz);
    long long r = t / hz;
    int ay = (int)(r % hy), ax = (int)(r / hy);
    int par0 = (g.gx0 + g.own_lo) & 1;
    int l = g.own_lo + 2 * ax + ((a.cx - par0) & 1);
    int iy = 2 * ay + a.cy, iz = 2 * az + a.cz;
    cell_update_global<LOG>(a, a.phase, pos, rel, cs, l, iy, iz, 0, 1 << 30, n_acc, n_ov, n_cell, log, nlog, logcap);
  }
  // block-aggregated counters
  __shared__ int s_cnt[3];
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  if (n_acc) atomicAdd(&s_cnt[0], n_acc);
  if (n_ov) atomicAdd(&s_cnt[1], n_ov);
  if (n_cell) atomicAdd(&s_cnt[2], n_cell);
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = s_cnt[0] + s_cnt[1] + s_cnt[2];
    if (tot) {
      atomicAdd(&cnt[CNT_TRIALS], (unsigned long long)tot);
      if (s_cnt[0]) atomicAdd(&cnt[CNT_ACC], (unsigned long long)s_cnt[0]);
      if (s_cnt[1]) atomicAdd(&cnt[CNT_REJ_OVERLAP], (unsigned long long)s_cnt[1]);
      if (s_cnt[2]) atomicAdd(&cnt[CNT_REJ_CELL], (unsigned long long)s_cnt[2]);
    }
  }
}


// out-of-line copy for the staged kernels' fallback path (blocks that do not fit the staged scheme)
template <bool LOG>
__device__ __noinline__ void cell_update_global_noinline(const SweepArgs a, int phase, double4* __restrict__ pos,
                                                         float4* __restrict__ rel, const int* __restrict__ cs, int l,
                                                         int iy, int iz, int j0, int j1, int& n_acc, int& n_ov,
                                                         int& n_cell, hsmc_gpu_trial* __restrict__ log,
                                                         unsigned long long* __restrict__ nlog, long long logcap) {
  cell_update_global<LOG>(a, phase, pos, rel, cs, l, iy, iz, j0, j1, n_acc, n_ov, n_cell, log, nlog, logcap);
}
