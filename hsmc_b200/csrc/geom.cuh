// Geometry shared by all kernels: box, cell grid, pair tests.
//
// Arithmetic contract (bit-exactness with the reference): double
// precision, no FMA contraction (the library is compiled with -fmad=false), the
// reference's operation order:
//   d = (a - b) * sf;  if (d > L*sf/2) d -= L*sf; else if (d < -L*sf/2) d += L*sf;
//   r2 = dx*dx + dy*dy + dz*dz;  overlap <=> sqrt(r2) < 1.0 <=> r2 < 1.0
// (moves.c:400-431; sqrt is correctly rounded and sqrt(1-2^-53) rounds below 1, so the
// square root is only taken where a bin index needs it).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct Box {
  double Lx, Ly, Lz;     // edges (already multiplied by sf for scaled tests)
  double hx, hy, hz;     // (L*sf)/2.0
};

__host__ __device__ inline Box make_box(double lx, double ly, double lz, double sf) {
  Box b;
  b.Lx = lx * sf; b.Ly = ly * sf; b.Lz = lz * sf;   // moves.c:405-407
  b.hx = b.Lx / 2.0; b.hy = b.Ly / 2.0; b.hz = b.Lz / 2.0;   // moves.c:408-410
  return b;
}

// Cell grid.  Cells are indexed c = (lx*ny + iy)*nz + iz (z fastest), a true bijection
// (the reference's ix*nx*nx + iy*ny + iz only is one on cubic grids).
// The grid origin is shifted by (sx,sy,sz) in [0,w) -- redrawn every regrid -- so that
// cell walls move although trial moves never leave a cell within a sweep.
// With world > 1 a rank stores the x-layers [gx0, gx0+nlx) (mod nx): one ghost layer,
// its owned layers [own_lo, own_hi), one ghost layer.  world == 1: gx0 = 0, nlx = nx,
// own = [0, nx), and x neighbours wrap.
struct Grid {
  int nx, ny, nz;
  int nlx, gx0, own_lo, own_hi, wrap_x;
  double wx, wy, wz;
  double iwx, iwy, iwz;
  double sx, sy, sz;
  double Lx, Ly, Lz;
};

__host__ __device__ inline int axis_cell(double x, double s, double iw, int n) {
  int i = (int)floor((x - s) * iw);
  if (i < 0) i += n;
  else if (i >= n) i -= n;
  return i;
}

// local layer of a global x cell index, or -1 when that layer is not resident here
__host__ __device__ inline int local_layer(const Grid& g, int gix) {
  int l = gix - g.gx0;
  if (l < 0) l += g.nx;
  return (l < g.nlx) ? l : -1;
}

__host__ __device__ inline long long local_cell(const Grid& g, double x, double y, double z) {
  int l = local_layer(g, axis_cell(x, g.sx, g.iwx, g.nx));
  if (l < 0) return -1;
  int iy = axis_cell(y, g.sy, g.iwy, g.ny);
  int iz = axis_cell(z, g.sz, g.iwz, g.nz);
  return ((long long)l * g.ny + iy) * g.nz + iz;
}

__host__ __device__ inline long long global_cell_of_local(const Grid& g, int l, int iy, int iz) {
  int gx = g.gx0 + l;
  if (gx >= g.nx) gx -= g.nx;
  return ((long long)gx * g.ny + iy) * g.nz + iz;
}

#if defined(__CUDACC__)

// float4 shadow entry of a particle: offset from the origin of its own cell (gx,iy,iz)
// and, in the fourth word, the z index of that cell (what the staging pass of k_sweep_lean needs to
// turn the offset into a block-relative coordinate).  The cell of index n-1 straddles the periodic box edge when the grid is
// shifted, hence the +-L repair.  Offsets lie in [0, edge]: fp32 keeps ~1e-7 absolute.
__device__ __forceinline__ float4 make_rel(const Grid& g, int gx, int iy, int iz, double x, double y, double z) {
  double ox = x - (g.sx + gx * g.wx), oy = y - (g.sy + iy * g.wy), oz = z - (g.sz + iz * g.wz);
  if (ox < -0.5 * g.wx) ox += g.Lx; else if (ox > 1.5 * g.wx) ox -= g.Lx;
  if (oy < -0.5 * g.wy) oy += g.Ly; else if (oy > 1.5 * g.wy) oy -= g.Ly;
  if (oz < -0.5 * g.wz) oz += g.Lz; else if (oz > 1.5 * g.wz) oz -= g.Lz;
  return make_float4((float)ox, (float)oy, (float)oz, __int_as_float(iz));
}

// shadow entry from a local cell index
__device__ __forceinline__ float4 make_rel_cell(const Grid& g, long long c, const double4& p) {
  int iz = (int)(c % g.nz);
  long long r = c / g.nz;
  int iy = (int)(r % g.ny), l = (int)(r / g.ny);
  int gx = g.gx0 + l;
  if (gx >= g.nx) gx -= g.nx;
  return make_rel(g, gx, iy, iz, p.x, p.y, p.z);
}

// unscaled squared distance, minimum image
__device__ __forceinline__ double pair_r2(double xi, double yi, double zi, double xj, double yj,
                                          double zj, const Box& b) {
  double dx = xi - xj, dy = yi - yj, dz = zi - zj;
  if (dx > b.hx) dx -= b.Lx; else if (dx < -b.hx) dx += b.Lx;
  if (dy > b.hy) dy -= b.Ly; else if (dy < -b.hy) dy += b.Ly;
  if (dz > b.hz) dz -= b.Lz; else if (dz < -b.hz) dz += b.Lz;
  return dx * dx + dy * dy + dz * dz;
}

// scaled squared distance, box already scaled (make_box(L, sf))
__device__ __forceinline__ double pair_r2_scaled(double xi, double yi, double zi, double xj,
                                                 double yj, double zj, double sf, const Box& b) {
  double dx = (xi - xj) * sf, dy = (yi - yj) * sf, dz = (zi - zj) * sf;
  if (dx > b.hx) dx -= b.Lx; else if (dx < -b.hx) dx += b.Lx;
  if (dy > b.hy) dy -= b.Ly; else if (dy < -b.hy) dy += b.Ly;
  if (dz > b.hz) dz -= b.Lz; else if (dz < -b.hz) dz += b.Lz;
  return dx * dx + dy * dy + dz * dz;
}

// Visit the particle slots of the 27-cell stencil around local cell (l, iy, iz) as at
// most 18 contiguous slot ranges (z is the fastest cell index, so the three z-cells of
// one (x,y) column are adjacent unless they wrap).  f(k) returns true to stop early.
template <class F>
__device__ __forceinline__ bool stencil_any(const Grid& g, const int* __restrict__ cs, int l, int iy,
                                            int iz, F f) {
#pragma unroll 1
  for (int dx = -1; dx <= 1; dx++) {
    int ll = l + dx;
    if (g.wrap_x) {
      if (ll < 0) ll += g.nlx; else if (ll >= g.nlx) ll -= g.nlx;
    }
#pragma unroll 1
    for (int dy = -1; dy <= 1; dy++) {
      int yy = iy + dy;
      if (yy < 0) yy += g.ny; else if (yy >= g.ny) yy -= g.ny;
      long long rb = ((long long)ll * g.ny + yy) * g.nz;
      int zlo = iz - 1, zhi = iz + 1;
      if (zlo >= 0 && zhi < g.nz) {
        int b = cs[rb + zlo], e = cs[rb + zhi + 1];
        for (int k = b; k < e; k++)
          if (f(k)) return true;
      } else {
        // wrapped column: cell nz-1 or cell 0 is the periodic neighbour
        int b1, e1, b2, e2;
        if (zlo < 0) { b1 = cs[rb + g.nz - 1]; e1 = cs[rb + g.nz]; b2 = cs[rb]; e2 = cs[rb + 2]; }
        else         { b1 = cs[rb + g.nz - 2]; e1 = cs[rb + g.nz]; b2 = cs[rb]; e2 = cs[rb + 1]; }
        for (int k = b1; k < e1; k++)
          if (f(k)) return true;
        for (int k = b2; k < e2; k++)
          if (f(k)) return true;
      }
    }
  }
  return false;
}

// The forward HALF of the stencil: the own cell plus the 13 neighbour cells (dx,dy,dz) that are
// lexicographically greater than (0,0,0).  With at least four cells per axis a cell and its
// mirror image are distinct, so looping over all cells visits every unordered pair of particles
// in adjacent cells exactly once (pairs inside the own cell: the caller orders them, own == true).
// In slab mode only owned cells loop; a pair across a slab face is seen from the lower-x cell's
// rank only.  Same contiguous-row trick as stencil_any.  f(k, own) returns true to stop early.
template <class F>
__device__ __forceinline__ bool stencil_half(const Grid& g, const int* __restrict__ cs, int l, int iy,
                                             int iz, F f) {
  // row (0,0): own cell and its +z neighbour
  {
    const long long rb = ((long long)l * g.ny + iy) * g.nz;
    const int b = cs[rb + iz], m = cs[rb + iz + 1];
    for (int k = b; k < m; k++)
      if (f(k, true)) return true;
    int b2 = m, e2;
    if (iz + 1 < g.nz) e2 = cs[rb + iz + 2];
    else { b2 = cs[rb]; e2 = cs[rb + 1]; }
    for (int k = b2; k < e2; k++)
      if (f(k, false)) return true;
  }
#pragma unroll 1
  for (int r = 0; r < 4; r++) {            // rows (0,+1), (+1,-1), (+1,0), (+1,+1)
    const int dx = r == 0 ? 0 : 1, dy = r == 0 ? 1 : r - 2;
    int ll = l + dx;
    if (g.wrap_x && ll >= g.nlx) ll -= g.nlx;
    int yy = iy + dy;
    if (yy < 0) yy += g.ny; else if (yy >= g.ny) yy -= g.ny;
    const long long rb = ((long long)ll * g.ny + yy) * g.nz;
    const int zlo = iz - 1, zhi = iz + 1;
    if (zlo >= 0 && zhi < g.nz) {
      const int b = cs[rb + zlo], e = cs[rb + zhi + 1];
      for (int k = b; k < e; k++)
        if (f(k, false)) return true;
    } else {
      int b1, e1, b2, e2;
      if (zlo < 0) { b1 = cs[rb + g.nz - 1]; e1 = cs[rb + g.nz]; b2 = cs[rb]; e2 = cs[rb + 2]; }
      else         { b1 = cs[rb + g.nz - 2]; e1 = cs[rb + g.nz]; b2 = cs[rb]; e2 = cs[rb + 1]; }
      for (int k = b1; k < e1; k++)
        if (f(k, false)) return true;
      for (int k = b2; k < e2; k++)
        if (f(k, false)) return true;
    }
  }
  return false;
}

#endif  // __CUDACC__
