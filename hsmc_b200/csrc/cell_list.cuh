// cell_list.cuh -- K1: counting-sort cell list (count / scan / scatter) and the host <-> device table conversion kernels
// (part of the single translation unit hsmc_gpu.cu; included there, in this order)
#pragma once

// ----------------------------------------------------------------------------------
// K1: cell list = counting sort
// ----------------------------------------------------------------------------------
// pass 1: cell key of every particle, rank inside its cell by atomic counter
__global__ void k_cell_count(Grid g, const double4* __restrict__ in, int n, int* __restrict__ key,
                             int* __restrict__ rnk, int* __restrict__ count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = in[i];
  long long c = local_cell(g, p.x, p.y, p.z);
  key[i] = (int)c;
  if (c >= 0) rnk[i] = atomicAdd(&count[c], 1);
}

// exclusive scan of count[0..n) into out[0..n], out[n] = total: three small kernels
#define SCAN_T 512
#define SCAN_V 8
#define SCAN_CHUNK (SCAN_T * SCAN_V)

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int wsum[SCAN_T / 32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = (lane < SCAN_T / 32) ? wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane < SCAN_T / 32) wsum[lane] = s;
  }
  __syncthreads();
  int base = (w > 0) ? wsum[w - 1] : 0;
  *total = wsum[SCAN_T / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

__global__ void k_scan_blocksum(const int* __restrict__ in, long long n, int* __restrict__ bsum) {
  long long base = (long long)blockIdx.x * SCAN_CHUNK;
  int s = 0;
  if (base + SCAN_CHUNK <= n) {
    const int4* p4 = reinterpret_cast<const int4*>(in + base);
#pragma unroll
    for (int j = 0; j < SCAN_V / 4; j++) {
      const int4 a = p4[j * SCAN_T + threadIdx.x];
      s += a.x + a.y + a.z + a.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_V; j++) {
      long long i = base + (long long)j * SCAN_T + threadIdx.x;
      if (i < n) s += in[i];
    }
  }
  int tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void k_scan_top(int* bsum, int nb) {
  // single block; sequential over chunks of SCAN_T with a running carry
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += SCAN_T) {
    int i = b0 + threadIdx.x;
    int v = (i < nb) ? bsum[i] : 0;
    int tot;
    int ex = block_exclusive_scan(v, &tot);
    if (i < nb) bsum[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
}

// last pass of the scan
__global__ void k_scan_final(const int* __restrict__ in, long long n, const int* __restrict__ bsum,
                             int* __restrict__ out) {
  long long base = (long long)blockIdx.x * SCAN_CHUNK + (long long)threadIdx.x * SCAN_V;
  int v[SCAN_V];
  int s = 0;
  static_assert(SCAN_V == 8, "two int4 per thread");
  if (base + SCAN_V <= n) {                       // 32-byte vector path (cudaMalloc'd arrays, base % 8 == 0)
    const int4 a0 = reinterpret_cast<const int4*>(in + base)[0], a1 = reinterpret_cast<const int4*>(in + base)[1];
    v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_V; j++) v[j] = (base + j < n) ? in[base + j] : 0;
  }
#pragma unroll
  for (int j = 0; j < SCAN_V; j++) s += v[j];
  int tot;
  int ex = block_exclusive_scan(s, &tot) + bsum[blockIdx.x];
  int o[SCAN_V];
#pragma unroll
  for (int j = 0; j < SCAN_V; j++) { o[j] = ex; ex += v[j]; }
  if (base + SCAN_V <= n) {
    reinterpret_cast<int4*>(out + base)[0] = make_int4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<int4*>(out + base)[1] = make_int4(o[4], o[5], o[6], o[7]);
    if (base + SCAN_V == n) out[n] = ex;
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_V; j++) {
      if (base + j < n) out[base + j] = o[j];
      if (base + j == n - 1) out[n] = o[j] + v[j];
    }
  }
}

// pass 3: scatter into cell order
__global__ void k_cell_scatter(Grid g, const double4* __restrict__ in, int n, const int* __restrict__ key,
                               const int* __restrict__ rnk, const int* __restrict__ cs,
                               double4* __restrict__ out, float4* __restrict__ rel) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = key[i];
  if (c < 0) return;
  double4 p = in[i];
  int d = cs[c] + rnk[i];
  out[d] = p;
  rel[d] = make_rel_cell(g, c, p);
}

// ----------------------------------------------------------------------------------
// host <-> device table conversion ({id,x,y,z} rows <-> {x,y,z,id} slots)
// ----------------------------------------------------------------------------------
__global__ void k_unpack_rows(const double* __restrict__ rows, int n, double4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 r = reinterpret_cast<const double4*>(rows)[i];
  out[i] = make_double4(r.y, r.z, r.w, r.x);
}

__global__ void k_pack_by_id(const double4* __restrict__ pos, int first, int n, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = pos[first + i];
  long long id = (long long)p.w;
  reinterpret_cast<double4*>(out)[id] = make_double4(p.w, p.x, p.y, p.z);
}

__global__ void k_pack_rows(const double4* __restrict__ pos, int first, int n, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = pos[first + i];
  reinterpret_cast<double4*>(out)[i] = make_double4(p.w, p.x, p.y, p.z);
}

__global__ void k_slot_of_id(const double4* __restrict__ pos, int n, int* __restrict__ slot) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  slot[(long long)pos[i].w] = i;
}

