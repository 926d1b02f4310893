// observables.cuh -- K3-K8: scaled-overlap verdict, Widom insertions, RDF and contact histograms, minimum distance, order parameter q_l, volume rescale
// (part of the single translation unit hsmc_gpu.cu; included there, in this order)
#pragma once

// ----------------------------------------------------------------------------------
// K3: global scaled-overlap verdict for nsf scale factors in one pass over the pairs.
// One thread per owned cell over the forward half of its stencil (stencil_half): each unordered
// pair is loaded and visited once.
// A cheap unscaled pre-test skips pairs that cannot overlap under any of the factors
// (threshold carries a 1e-6 relative margin, far above rounding); pairs that pass are
// evaluated with the reference's exact scaled arithmetic for every factor.
// ----------------------------------------------------------------------------------
#define MAX_SF 64
struct SfArgs {
  int n;
  double r2_skip;       // unscaled r2 above which no factor can give an overlap
  double sf[MAX_SF];
  Box box[MAX_SF];
};

__global__ void __launch_bounds__(128)
k_overlap_scaled(Grid g, Box ubox, const SfArgs* __restrict__ sa, const double4* __restrict__ pos,
                 const int* __restrict__ cs, int* __restrict__ flags) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)(g.own_hi - g.own_lo) * g.ny * g.nz;
  if (t >= total) return;
  int nsf = sa->n;
  if (nsf == 1 && flags[0]) return;   // verdict already known
  int iz = (int)(t % g.nz);
  long long r = t / g.nz;
  int iy = (int)(r % g.ny), l = g.own_lo + (int)(r / g.ny);
  long long c = ((long long)l * g.ny + iy) * g.nz + iz;
  int beg = cs[c], end = cs[c + 1];
  double r2_skip = sa->r2_skip;
  for (int s = beg; s < end; s++) {
    double4 p = pos[s];
    stencil_half(g, cs, l, iy, iz, [&](int k, bool own) {
      double4 q = pos[k];
      if (own && !(q.w > p.w)) return false;
      if (pair_r2(p.x, p.y, p.z, q.x, q.y, q.z, ubox) > r2_skip) return false;
      bool all = true;
      for (int m = 0; m < nsf; m++) {
        if (pair_r2_scaled(p.x, p.y, p.z, q.x, q.y, q.z, sa->sf[m], sa->box[m]) < 1.0) flags[m] = 1;
        else all = false;
      }
      return nsf == 1 && all;
    });
  }
}

// The same verdict when a factor compresses the cells below the particle diameter (cell * sf < 1): pairs two
// cells apart can then overlap, which the 27-cell stencil cannot see (the reference scans its unscaled
// 27-stencil there and misses them, moves.c:108 / SURVEY section 7; it then exits in cell_list.c:127 if such
// a move is accepted).  Rare path (a compression proposal while the cell edge is within dv of 1.0): one thread
// per owned cell over the full (2R+1)^3 stencil, each unordered pair taken from the lower-id side.  Single
// GPU only (a slab keeps one ghost layer).
__global__ void __launch_bounds__(128)
k_overlap_scaled_wide(Grid g, Box ubox, const SfArgs* __restrict__ sa, const double4* __restrict__ pos,
                      const int* __restrict__ cs, int R, int* __restrict__ flags) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)g.nlx * g.ny * g.nz;
  if (t >= total) return;
  const int nsf = sa->n;
  int iz = (int)(t % g.nz);
  long long r = t / g.nz;
  int iy = (int)(r % g.ny), l = (int)(r / g.ny);
  const int beg = cs[t], end = cs[t + 1];
  const double r2_skip = sa->r2_skip;
  for (int s = beg; s < end; s++) {
    const double4 p = pos[s];
    for (int dx = -R; dx <= R; dx++)
      for (int dy = -R; dy <= R; dy++)
        for (int dz = -R; dz <= R; dz++) {
          const int ll = ((l + dx) % g.nlx + g.nlx) % g.nlx, yy = ((iy + dy) % g.ny + g.ny) % g.ny,
                    zz = ((iz + dz) % g.nz + g.nz) % g.nz;
          const long long c2 = ((long long)ll * g.ny + yy) * g.nz + zz;
          for (int k = cs[c2]; k < cs[c2 + 1]; k++) {
            const double4 q = pos[k];
            if (!(q.w > p.w)) continue;
            if (pair_r2(p.x, p.y, p.z, q.x, q.y, q.z, ubox) > r2_skip) continue;
            for (int m = 0; m < nsf; m++)
              if (pair_r2_scaled(p.x, p.y, p.z, q.x, q.y, q.z, sa->sf[m], sa->box[m]) < 1.0) flags[m] = 1;
          }
        }
  }
}

// ----------------------------------------------------------------------------------
// Near pairs through the fp32 shadow (K3 / K6, default): one thread per owned particle slot walks the forward half
// of its 27-cell stencil in the 16-byte shadow table (cell-relative offsets, z cell in .w) and computes r^2 in fp32
// from offsets + (cell difference) * edge; only pairs whose fp32 r^2 is below the caller's threshold (exact
// threshold + a margin far above the fp32 error: offsets are at most 1.5 cell edges, so r^2 near the threshold carries
// an error below ~2e-5 * max(1, edge)^2) are handed to f(k), which evaluates them from the master table in the
// reference's double arithmetic.  The pair set and the arithmetic behind every verdict / bin are those of the
// all-double kernels above; the filter only drops pairs that are certainly far.  Every unordered pair is visited once:
// inside the own cell from the lower slot, between cells from the cell that is lexicographically first (stencil_half's
// rule, so slabs see a pair across a face once, from the lower-x rank).
// ----------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void near_pairs_f32(const Grid& g, const float4* __restrict__ rel, const int* __restrict__ cs,
                                               int s, int l, int iy, int iz, float thr, F f) {
  const float4 me = rel[s];
  const float wxf = (float)g.wx, wyf = (float)g.wy, wzf = (float)g.wz;
  auto test = [&](int k, float ox, float oy, float oz) {
    const float4 q = rel[k];
    const float dx = (q.x + ox) - me.x, dy = (q.y + oy) - me.y, dz = (q.z + oz) - me.z;
    if (__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx)) < thr) f(k);
  };
  {
    // row (0,0): the later slots of the own cell, then the +z neighbour cell
    const long long rb = ((long long)l * g.ny + iy) * g.nz;
    const int m = cs[rb + iz + 1];
    for (int k = s + 1; k < m; k++) test(k, 0.f, 0.f, 0.f);
    int b2 = m, e2;
    if (iz + 1 < g.nz) e2 = cs[rb + iz + 2];
    else { b2 = cs[rb]; e2 = cs[rb + 1]; }
    for (int k = b2; k < e2; k++) test(k, 0.f, 0.f, wzf);
  }
#pragma unroll 1
  for (int r = 0; r < 4; r++) {            // rows (0,+1), (+1,-1), (+1,0), (+1,+1): three cells each
    const int dx = r == 0 ? 0 : 1, dy = r == 0 ? 1 : r - 2;
    int ll = l + dx;
    if (g.wrap_x && ll >= g.nlx) ll -= g.nlx;
    int yy = iy + dy;
    if (yy < 0) yy += g.ny; else if (yy >= g.ny) yy -= g.ny;
    const long long rb = ((long long)ll * g.ny + yy) * g.nz;
    const float ox = (float)dx * wxf, oy = (float)dy * wyf;
    auto range = [&](int b, int e) {
      for (int k = b; k < e; k++) {
        const float4 q = rel[k];
        int dzc = __float_as_int(q.w) - iz;
        if (dzc > 1) dzc -= g.nz; else if (dzc < -1) dzc += g.nz;
        const float ddx = (q.x + ox) - me.x, ddy = (q.y + oy) - me.y, ddz = (q.z + (float)dzc * wzf) - me.z;
        if (__fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, ddx * ddx)) < thr) f(k);
      }
    };
    const int zlo = iz - 1, zhi = iz + 1;
    if (zlo >= 0 && zhi < g.nz) range(cs[rb + zlo], cs[rb + zhi + 1]);
    else if (zlo < 0) { range(cs[rb + g.nz - 1], cs[rb + g.nz]); range(cs[rb], cs[rb + 2]); }
    else { range(cs[rb + g.nz - 2], cs[rb + g.nz]); range(cs[rb], cs[rb + 1]); }
  }
}

// margin of the fp32 pre-test (see above)
__host__ __device__ inline float near_pairs_margin(const Grid& g) {
  const double w = fmax(1.0, fmax(g.wx, fmax(g.wy, g.wz)));
  return (float)(1.0e-4 + 2.0e-5 * w * w);
}

__global__ void __launch_bounds__(256)
k_overlap_scaled_f32(Grid g, Box ubox, const SfArgs* __restrict__ sa, const double4* __restrict__ pos,
                     const float4* __restrict__ rel, const int* __restrict__ cs, int first, int n, int* __restrict__ flags) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nsf = sa->n;
  const double r2_skip = sa->r2_skip;
  const int s = first + min(t, n - 1);
  // pairs that pass the filter are not evaluated where they are found (a lane with one such pair would hold its
  // warp for nsf serial evaluations): each lane keeps up to four, and the warp then evaluates them together, one
  // scale factor per lane
  int c0 = -1, c1 = -1, c2 = -1, c3 = -1;
  if (t < n && !(nsf == 1 && flags[0])) {                    // (one factor: verdict already known -> nothing to do)
    const double4 p = pos[s];
    const long long c = local_cell(g, p.x, p.y, p.z);
    const int iz = (int)(c % g.nz);
    const long long r = c / g.nz;
    near_pairs_f32(g, rel, cs, s, (int)(r / g.ny), (int)(r % g.ny), iz, (float)r2_skip + near_pairs_margin(g), [&](int k) {
      if (c3 < 0) { c3 = c2; c2 = c1; c1 = c0; c0 = k; return; }
      const double4 q = pos[k];                               // a fifth one: here and now
      if (pair_r2(p.x, p.y, p.z, q.x, q.y, q.z, ubox) > r2_skip) return;
      for (int m = 0; m < nsf; m++)
        if (pair_r2_scaled(p.x, p.y, p.z, q.x, q.y, q.z, sa->sf[m], sa->box[m]) < 1.0) flags[m] = 1;
    });
  }
  const int lane = threadIdx.x & 31;
  for (;;) {
    const unsigned have = __ballot_sync(0xffffffffu, c0 >= 0);
    if (!have) break;
    const int src = __ffs(have) - 1;
    const int k = __shfl_sync(0xffffffffu, c0, src), si = __shfl_sync(0xffffffffu, s, src);
    if (lane == src) { c0 = c1; c1 = c2; c2 = c3; c3 = -1; }
    const double4 p = pos[si], q = pos[k];
    if (pair_r2(p.x, p.y, p.z, q.x, q.y, q.z, ubox) > r2_skip) continue;
    for (int m = lane; m < nsf; m += 32)
      if (pair_r2_scaled(p.x, p.y, p.z, q.x, q.y, q.z, sa->sf[m], sa->box[m]) < 1.0) flags[m] = 1;
  }
}

// ----------------------------------------------------------------------------------
// K4: Widom insertions.  One thread per insertion point; the point is
// r = u * L (compute_widom_chem_pot.c:73-80) with u from Philox(sample, index).
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_widom(Grid g, Box box, const double4* __restrict__ pos, const int* __restrict__ cs, uint32_t key0,
        uint32_t key1, uint32_t sample_lo, uint32_t sample_hi, long long first, long long count,
        unsigned long long* __restrict__ accepted) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ok = 0;
  if (t < count) {
    unsigned long long m = (unsigned long long)(first + t);
    Philox4 rn = philox4x32_10((uint32_t)m, (HSMC_STREAM_WIDOM << 24) | (uint32_t)(m >> 32), sample_lo,
                               sample_hi, key0, key1);
    double rx = hsmc_u01(rn.v[0]) * g.Lx, ry = hsmc_u01(rn.v[1]) * g.Ly, rz = hsmc_u01(rn.v[2]) * g.Lz;
    int l = local_layer(g, axis_cell(rx, g.sx, g.iwx, g.nx));
    if (l >= g.own_lo && l < g.own_hi) {
      int iy = axis_cell(ry, g.sy, g.iwy, g.ny), iz = axis_cell(rz, g.sz, g.iwz, g.nz);
      bool ov = stencil_any(g, cs, l, iy, iz, [&](int k) {
        double4 q = pos[k];
        return pair_r2(rx, ry, rz, q.x, q.y, q.z, box) < 1.0;
      });
      ok = ov ? 0 : 1;
    }
  }
  unsigned m = __ballot_sync(0xffffffffu, ok);
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_ok, __popc(m));
  __syncthreads();
  if (threadIdx.x == 0 && s_ok) atomicAdd(accepted, (unsigned long long)s_ok);
}

// explicit points (parity entry point)
__global__ void k_widom_points(Grid g, Box box, const double4* __restrict__ pos, const int* __restrict__ cs,
                               const double* __restrict__ xyz, int n, int* __restrict__ flags) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double rx = xyz[3 * t], ry = xyz[3 * t + 1], rz = xyz[3 * t + 2];
  int l = local_layer(g, axis_cell(rx, g.sx, g.iwx, g.nx));
  int iy = axis_cell(ry, g.sy, g.iwy, g.ny), iz = axis_cell(rz, g.sz, g.iwz, g.nz);
  bool ov = stencil_any(g, cs, l, iy, iz, [&](int k) {
    double4 q = pos[k];
    return pair_r2(rx, ry, rz, q.x, q.y, q.z, box) < 1.0;
  });
  flags[t] = ov ? 1 : 0;
}

// explicit trial moves (parity entry point): verdict of check_overlap for particle
// idx placed at xyz with everything else fixed
__global__ void k_trial_points(Grid g, Box sbox, double sf, const double4* __restrict__ pos,
                               const int* __restrict__ cs, const int* __restrict__ idx,
                               const double* __restrict__ xyz, int n, int* __restrict__ flags) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double rx = xyz[3 * t], ry = xyz[3 * t + 1], rz = xyz[3 * t + 2];
  double id = (double)idx[t];
  int l = local_layer(g, axis_cell(rx, g.sx, g.iwx, g.nx));
  int iy = axis_cell(ry, g.sy, g.iwy, g.ny), iz = axis_cell(rz, g.sz, g.iwz, g.nz);
  bool ov = stencil_any(g, cs, l, iy, iz, [&](int k) {
    double4 q = pos[k];
    if (q.w == id) return false;
    return pair_r2_scaled(rx, ry, rz, q.x, q.y, q.z, sf, sbox) < 1.0;
  });
  flags[t] = ov ? 1 : 0;
}

// ----------------------------------------------------------------------------------
// K5: RDF pair histogram, all pairs (compute_rdf.c:110-128), shared-memory privatised.
// Tiles of RDF_T x RDF_T pairs; the j tile is staged in shared memory; the square root
// and the division (needed for a bit-exact bin index) are only evaluated for pairs that
// pass a conservative r2 pre-test.
// ----------------------------------------------------------------------------------
#define RDF_T 256
#define RDF_MAX_SMEM_BINS 8192
__global__ void __launch_bounds__(RDF_T)
k_rdf_pairs(const double4* __restrict__ pos, int n, Box box, double rmax, double r2_pre, double dr_bin,
            int nn, int ntile, long long b0, unsigned long long* __restrict__ hist) {
  // linear block index -> (ti, tj) with tj >= ti
  long long b = b0 + blockIdx.x;
  int ti = 0;
  {
    // rows of the upper triangle have ntile - ti entries
    double nt = (double)ntile;
    ti = (int)floor(((2.0 * nt + 1.0) - sqrt((2.0 * nt + 1.0) * (2.0 * nt + 1.0) - 8.0 * (double)b)) * 0.5);
    while ((long long)ti * (2LL * ntile - ti + 1) / 2 > b) ti--;
    while ((long long)(ti + 1) * (2LL * ntile - ti) / 2 <= b) ti++;
  }
  int tj = ti + (int)(b - (long long)ti * (2LL * ntile - ti + 1) / 2);
  extern __shared__ unsigned char smem_raw[];
  double* sx = reinterpret_cast<double*>(smem_raw);
  double* sy = sx + RDF_T;
  double* sz = sy + RDF_T;
  unsigned int* sh = reinterpret_cast<unsigned int*>(sz + RDF_T);
  bool use_sh = nn <= RDF_MAX_SMEM_BINS;
  if (use_sh)
    for (int k = threadIdx.x; k < nn; k += RDF_T) sh[k] = 0;
  int j0 = tj * RDF_T;
  int jn = min(RDF_T, n - j0);
  if ((int)threadIdx.x < jn) {
    double4 q = pos[j0 + threadIdx.x];
    sx[threadIdx.x] = q.x; sy[threadIdx.x] = q.y; sz[threadIdx.x] = q.z;
  }
  __syncthreads();
  int i = ti * RDF_T + threadIdx.x;
  if (i < n) {
    double4 p = pos[i];
    int jb = (ti == tj) ? (int)threadIdx.x + 1 : 0;
    for (int j = jb; j < jn; j++) {
      double r2 = pair_r2(p.x, p.y, p.z, sx[j], sy[j], sz[j], box);
      if (r2 < r2_pre) {
        double dr = sqrt(r2);
        if (dr < rmax) {
          int bin = (int)((dr - 1.0) / dr_bin);
          if (bin >= 0 && bin < nn) {
            if (use_sh) atomicAdd(&sh[bin], 1u);
            else atomicAdd(&hist[bin], 1ull);
          }
        }
      }
    }
  }
  __syncthreads();
  if (use_sh)
    for (int k = threadIdx.x; k < nn; k += RDF_T)
      if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
}

// ----------------------------------------------------------------------------------
// K6: near-contact pair histogram through the cell list (compute_press.c:123-165).
// ----------------------------------------------------------------------------------
#define CONTACT_MAX_BINS 1024
__global__ void __launch_bounds__(128)
k_contact_hist(Grid g, Box box, const double4* __restrict__ pos, const int* __restrict__ cs, double rmax,
               double dr_bin, int nn, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[CONTACT_MAX_BINS];
  for (int k = threadIdx.x; k < nn; k += blockDim.x) sh[k] = 0;
  __syncthreads();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)(g.own_hi - g.own_lo) * g.ny * g.nz;
  if (t < total) {
    int iz = (int)(t % g.nz);
    long long r = t / g.nz;
    int iy = (int)(r % g.ny), l = g.own_lo + (int)(r / g.ny);
    long long c = ((long long)l * g.ny + iy) * g.nz + iz;
    int beg = cs[c], end = cs[c + 1];
    double r2_pre = rmax * rmax * (1.0 + 1e-9);
    for (int s = beg; s < end; s++) {
      double4 p = pos[s];
      stencil_half(g, cs, l, iy, iz, [&](int k, bool own) {
        double4 q = pos[k];
        if (own && !(q.w > p.w)) return false;
        double r2 = pair_r2(p.x, p.y, p.z, q.x, q.y, q.z, box);
        if (r2 < r2_pre) {
          double dr = sqrt(r2);
          if (dr < rmax) {
            int bin = (int)((dr - 1.0) / dr_bin);
            if (bin >= 0 && bin < nn) atomicAdd(&sh[bin], 1u);
          }
        }
        return false;
      });
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nn; k += blockDim.x)
    if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
}

// K6 through the fp32 shadow (default; see near_pairs_f32): same pairs, same bins
__global__ void __launch_bounds__(256)
k_contact_hist_f32(Grid g, Box box, const double4* __restrict__ pos, const float4* __restrict__ rel,
                   const int* __restrict__ cs, int first, int n, double rmax, double dr_bin, int nn,
                   unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[CONTACT_MAX_BINS];
  for (int k = threadIdx.x; k < nn; k += blockDim.x) sh[k] = 0;
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    const int s = first + t;
    const double4 p = pos[s];
    const long long c = local_cell(g, p.x, p.y, p.z);
    const int iz = (int)(c % g.nz);
    const long long r = c / g.nz;
    const double r2_pre = rmax * rmax * (1.0 + 1e-9);
    near_pairs_f32(g, rel, cs, s, (int)(r / g.ny), (int)(r % g.ny), iz, (float)r2_pre + near_pairs_margin(g), [&](int k) {
      const double4 q = pos[k];
      const double r2 = pair_r2(p.x, p.y, p.z, q.x, q.y, q.z, box);
      if (r2 < r2_pre) {
        const double dr = sqrt(r2);
        if (dr < rmax) {
          const int bin = (int)((dr - 1.0) / dr_bin);
          if (bin >= 0 && bin < nn) atomicAdd(&sh[bin], 1u);
        }
      }
    });
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nn; k += blockDim.x)
    if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
}

// min pair r2 over the stencil (invariant check: never below 1.0 in a valid run)
__global__ void k_min_r2(Grid g, Box box, const double4* __restrict__ pos, const int* __restrict__ cs,
                         unsigned long long* __restrict__ out) {
  long long total = (long long)(g.own_hi - g.own_lo) * g.ny * g.nz;
  double best = 1e300;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int iz = (int)(t % g.nz);
    long long r = t / g.nz;
    int iy = (int)(r % g.ny), l = g.own_lo + (int)(r / g.ny);
    long long c = ((long long)l * g.ny + iy) * g.nz + iz;
    int beg = cs[c], end = cs[c + 1];
    for (int s = beg; s < end; s++) {
      double4 p = pos[s];
      stencil_any(g, cs, l, iy, iz, [&](int k) {
        double4 q = pos[k];
        if (q.w > p.w) best = fmin(best, pair_r2(p.x, p.y, p.z, q.x, q.y, q.z, box));
        return false;
      });
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0 && best < 1e299) atomicMin(out, (unsigned long long)__double_as_longlong(best));
}


// ----------------------------------------------------------------------------------
// K8: Steinhardt bond-order parameter q_l (compute_order_parameter.c:84-229), SURVEY 8f #1.
// One thread per owned particle: bonds = stencil neighbours with r <= rmax (rmax <= cell edge),
// q_lm(i) = <Y_lm(r_ij)>_bonds, q_l(i) = sqrt(4 pi/(2l+1) sum_m |q_lm|^2) -- |q_l,-m| = |q_l,m|,
// so m runs over 0..l with weight 2 for m > 0.  Y_lm by the normalised three-term recurrence
// (coefficients in constant memory), e^{i m phi} by rotation; all in double.  A floating-point
// observable (3-sigma contract), not on the bit-exact surface; the sum over particles is reduced
// in a fixed order so results are reproducible run to run.
// ----------------------------------------------------------------------------------
#define QL_MAX_L 12
#define QL_T 128
__constant__ double c_ql_A[(QL_MAX_L + 1) * (QL_MAX_L + 1)];   // A(k,m) = sqrt((4k^2-1)/(k^2-m^2))
__constant__ double c_ql_B[(QL_MAX_L + 1) * (QL_MAX_L + 1)];   // B(k,m) = sqrt(((k-1)^2-m^2)/(4(k-1)^2-1))
__constant__ double c_ql_D[QL_MAX_L + 2];                       // D(m) = sqrt((2m+1)/(2m))

__global__ void __launch_bounds__(QL_T)
k_order_param(Grid g, Box box, const double4* __restrict__ pos, const int* __restrict__ cs, int first, int n,
              int l, double rmax, double* __restrict__ partial) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double q = 0.0;
  if (t < n) {
    const double4 p = pos[first + t];
    const int ll = local_layer(g, axis_cell(p.x, g.sx, g.iwx, g.nx));
    const int iy = axis_cell(p.y, g.sy, g.iwy, g.ny), iz = axis_cell(p.z, g.sz, g.iwz, g.nz);
    double re[QL_MAX_L + 1], im[QL_MAX_L + 1];
    for (int m = 0; m <= l; m++) re[m] = im[m] = 0.0;
    int bonds = 0;
    stencil_any(g, cs, ll, iy, iz, [&](int k) {
      const double4 o = pos[k];
      if (o.w == p.w) return false;
      double dx = p.x - o.x, dy = p.y - o.y, dz = p.z - o.z;
      if (dx > box.hx) dx -= box.Lx; else if (dx < -box.hx) dx += box.Lx;
      if (dy > box.hy) dy -= box.Ly; else if (dy < -box.hy) dy += box.Ly;
      if (dz > box.hz) dz -= box.Lz; else if (dz < -box.hz) dz += box.Lz;
      const double rho2 = dx * dx + dy * dy, dr = sqrt(rho2 + dz * dz);
      if (dr > rmax) return false;
      bonds++;
      const double rho = sqrt(rho2);
      const double x = dz / dr, sth = rho / dr;
      const double cph = rho > 0.0 ? dx / rho : 1.0, sph = rho > 0.0 ? dy / rho : 0.0;
      double pmm = 0.28209479177387814;                  // sqrt(1/(4 pi))
      double cm = 1.0, sm = 0.0;
      for (int m = 0; m <= l; m++) {
        double p0 = 0.0, p1 = pmm;
        for (int k2 = m + 1; k2 <= l; k2++) {
          const double p2 = c_ql_A[k2 * (QL_MAX_L + 1) + m] * (x * p1 - c_ql_B[k2 * (QL_MAX_L + 1) + m] * p0);
          p0 = p1; p1 = p2;
        }
        re[m] += p1 * cm;
        im[m] += p1 * sm;
        pmm = -c_ql_D[m + 1] * sth * pmm;
        const double c2 = cm * cph - sm * sph;
        sm = sm * cph + cm * sph;
        cm = c2;
      }
      return false;
    });
    double sum = 0.0;
    if (bonds) {
      const double inv = 1.0 / (double)bonds;
      for (int m = 0; m <= l; m++) {
        const double a = re[m] * inv, b = im[m] * inv;
        sum += (m == 0 ? 1.0 : 2.0) * (a * a + b * b);
      }
    }
    q = sqrt(sum * (4.0 * 3.14159265358979323846 / (double)(2 * l + 1)));
  }
  // fixed-order block reduction
  __shared__ double sh[QL_T];
  sh[threadIdx.x] = q;
  __syncthreads();
  for (int o = QL_T / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// sum of the block partials in a fixed order (one block)
__global__ void __launch_bounds__(256)
k_sum_partials(const double* __restrict__ partial, int nb, double* __restrict__ out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// ----------------------------------------------------------------------------------
// K7: accepted volume move (moves.c:135-141): x *= sf, then apply_pbc with the new box
// ----------------------------------------------------------------------------------
__global__ void k_rescale(double4* __restrict__ pos, int n, double sf, double lx, double ly, double lz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = pos[i];
  p.x *= sf; p.y *= sf; p.z *= sf;
  if (p.x > lx) p.x -= lx; else if (p.x < 0.0) p.x += lx;
  if (p.y > ly) p.y -= ly; else if (p.y < 0.0) p.y += ly;
  if (p.z > lz) p.z -= lz; else if (p.z < 0.0) p.z += lz;
  pos[i] = p;
}

