// slab.cuh -- slab decomposition kernels: classification, halo messages, flags, layer bookkeeping
// (part of the single translation unit hsmc_gpu.cu; included there, in this order)
#pragma once

// ----------------------------------------------------------------------------------
// slab decomposition (world > 1): classify + halo buffers
// ----------------------------------------------------------------------------------
// Source particles are this rank's previously owned ones (or, for an upload, arbitrary
// rows).  Each is keyed into the local cell grid and, when it lies in one of the two
// layers at either end of the slab, also appended to the buffer bound for that
// neighbour: layers {0,1} -> left, {nlx-2,nlx-1} -> right (migrants + fresh ghosts in one
// message).  upload_mode keeps only owned layers and sends only boundary layers.
__global__ void k_slab_classify(Grid g, const double4* __restrict__ in, int n, int rows_layout,
                                int upload_mode, int* __restrict__ key, int* __restrict__ rnk,
                                int* __restrict__ count, double4* __restrict__ send_l,
                                double4* __restrict__ send_r, int* __restrict__ halo_cnt, int cap_halo,
                                const int* __restrict__ d_lay) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (d_lay) {                       // source = the previously owned slot range, known on the device only
    n = d_lay[4] - d_lay[1];
    in += d_lay[1];
  }
  if (i >= n) return;
  double4 p = in[i];
  if (rows_layout) p = make_double4(p.y, p.z, p.w, p.x);
  long long c = local_cell(g, p.x, p.y, p.z);
  int lyr = (c >= 0) ? (int)(c / ((long long)g.ny * g.nz)) : -1;
  bool keep, to_l, to_r;
  if (upload_mode) {
    keep = lyr >= g.own_lo && lyr < g.own_hi;
    to_l = lyr == g.own_lo;
    to_r = lyr == g.own_hi - 1;
  } else {
    if (c < 0) atomicOr(&halo_cnt[2], 1);   // moved more than one layer: impossible by construction
    keep = c >= 0;
    to_l = keep && lyr <= 1;
    to_r = keep && lyr >= g.nlx - 2;
  }
  key[i] = keep ? (int)c : -1;
  if (keep) rnk[i] = atomicAdd(&count[c], 1);
  if (to_l) {
    int s = atomicAdd(&halo_cnt[0], 1);
    if (s + 1 < cap_halo) send_l[s + 1] = p; else atomicOr(&halo_cnt[2], 2);
  }
  if (to_r) {
    int s = atomicAdd(&halo_cnt[1], 1);
    if (s + 1 < cap_halo) send_r[s + 1] = p; else atomicOr(&halo_cnt[2], 2);
  }
}

// p2p: publish "message seq is complete" to a neighbour's window / wait for a neighbour's
__global__ void k_flag_post(volatile uint32_t* flag_a, volatile uint32_t* flag_b, uint32_t seq) {
  __threadfence_system();
  if (flag_a) *flag_a = seq;
  if (flag_b) *flag_b = seq;
  __threadfence_system();
}
// A wait is bounded: a neighbour that never delivers (its process died, say) must turn into a loud CUDA
// error on this rank, not into a GPU that spins for ever.  HSMC_SPIN_LIMIT_NS is far beyond any legitimate
// delay (a rank writing a 16.8M-particle snapshot keeps its neighbours waiting for seconds).
#define HSMC_SPIN_LIMIT_NS (300ull * 1000000000ull)
__device__ __forceinline__ void hsmc_wait_flag(volatile uint32_t* flag, uint32_t seq) {
  unsigned long long t0 = 0;
  unsigned int spins = 0;
  while ((int32_t)(*flag - seq) < 0) {
    __nanosleep(200);
    if ((++spins & 4095u) == 0) {
      const unsigned long long t = hsmc_globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > HSMC_SPIN_LIMIT_NS) __trap();
    }
  }
}
__global__ void k_flag_wait(volatile uint32_t* flag_a, volatile uint32_t* flag_b, uint32_t seq) {
  if (flag_a) hsmc_wait_flag(flag_a, seq);
  if (flag_b) hsmc_wait_flag(flag_b, seq);
  __threadfence_system();
}

__global__ void k_halo_headers(double4* send_l, double4* send_r, const int* halo_cnt) {
  send_l[0] = make_double4((double)halo_cnt[0], 0, 0, 0);
  send_r[0] = make_double4((double)halo_cnt[1], 0, 0, 0);
}

// key the received particles (count in the header slot)
__global__ void k_recv_count(Grid g, const double4* __restrict__ buf0, const double4* __restrict__ buf1,
                             int cap_halo, int* __restrict__ key, int* __restrict__ rnk, int* __restrict__ count,
                             int* __restrict__ halo_cnt) {
  const double4* buf = blockIdx.y ? buf1 : buf0;
  key += (size_t)blockIdx.y * cap_halo;
  rnk += (size_t)blockIdx.y * cap_halo;
  int n = (int)buf[0].x;
  if (n > cap_halo - 1) n = cap_halo - 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double4 p = buf[i + 1];
    long long c = local_cell(g, p.x, p.y, p.z);
    key[i] = (int)c;
    if (c >= 0) rnk[i] = atomicAdd(&count[c], 1);
    else atomicOr(&halo_cnt[2], 4);
  }
}

__global__ void k_recv_scatter(Grid g, const double4* __restrict__ buf0, const double4* __restrict__ buf1,
                               int cap_halo, const int* __restrict__ key, const int* __restrict__ rnk,
                               const int* __restrict__ cs, double4* __restrict__ out, float4* __restrict__ rel,
                               int cap, int* __restrict__ flags) {
  const double4* buf = blockIdx.y ? buf1 : buf0;
  key += (size_t)blockIdx.y * cap_halo;
  rnk += (size_t)blockIdx.y * cap_halo;
  int n = (int)buf[0].x;
  if (n > cap_halo - 1) n = cap_halo - 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int c = key[i];
    if (c >= 0) {
      double4 p = buf[i + 1];
      int d = cs[c] + rnk[i];
      if (d >= cap) { atomicOr(flags, 32); continue; }
      out[d] = p;
      rel[d] = make_rel_cell(g, c, p);
    }
  }
}

// boundary layer -> message buffer (count in the header slot); the slot range of the layer
// is only known on the device
__global__ void k_halo_pack(Grid g, const double4* __restrict__ pos, const int* __restrict__ cs, int layer,
                            double4* __restrict__ buf, int cap_msg, int* __restrict__ flags) {
  long long per = (long long)g.ny * g.nz;
  int b = cs[(long long)layer * per], e = cs[(long long)(layer + 1) * per];
  int n = e - b;
  if (n > cap_msg - 1) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(flags, 2); n = cap_msg - 1; }
  if (blockIdx.x == 0 && threadIdx.x == 0) buf[0] = make_double4((double)n, 0, 0, 0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[1 + i] = pos[b + i];
}

// message buffer -> ghost layer, slot for slot (both sides keep these layers id-sorted),
// plus the fp32 shadow of the refreshed slots
__global__ void k_halo_unpack(Grid g, double4* __restrict__ pos, float4* __restrict__ rel, const int* __restrict__ cs,
                              int layer, const double4* __restrict__ buf, int* __restrict__ flags) {
  long long per = (long long)g.ny * g.nz;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int b0 = cs[(long long)layer * per];
  if (t == 0 && (int)buf[0].x != cs[(long long)(layer + 1) * per] - b0) atomicOr(flags, 16);
  if (t >= per) return;
  long long c = (long long)layer * per + t;
  for (int i = cs[c]; i < cs[c + 1]; i++) {
    double4 p = buf[1 + (i - b0)];
    pos[i] = p;
    rel[i] = make_rel_cell(g, c, p);
  }
}

__global__ void k_scatter_layout(Grid g, const double4* __restrict__ in, int n, int rows_layout,
                                 const int* __restrict__ key, const int* __restrict__ rnk,
                                 const int* __restrict__ cs, double4* __restrict__ out, float4* __restrict__ rel,
                                 const int* __restrict__ d_lay, int cap, int* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (d_lay) {
    n = d_lay[4] - d_lay[1];
    in += d_lay[1];
  }
  if (i >= n) return;
  int c = key[i];
  if (c < 0) return;
  double4 p = in[i];
  if (rows_layout) p = make_double4(p.y, p.z, p.w, p.x);
  int d = cs[c] + rnk[i];
  if (d >= cap) { atomicOr(flags, 32); return; }
  out[d] = p;
  rel[d] = make_rel_cell(g, c, p);
}

// canonical (ascending id) slot order inside every cell of the given layers, so that a
// boundary layer and its ghost copy on the neighbour are slot-for-slot identical
__global__ void k_sort_cells_by_id(Grid g, double4* __restrict__ pos, float4* __restrict__ rel,
                                   const int* __restrict__ cs, int layer_a, int layer_b) {
  long long per = (long long)g.ny * g.nz;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * per) return;
  if (blockIdx.y) { layer_a = g.nlx - 2; layer_b = g.nlx - 1; }     // second pair of layers
  long long c = (t < per) ? (long long)layer_a * per + t : (long long)layer_b * per + (t - per);
  int beg = cs[c], end = cs[c + 1];
  for (int i = beg + 1; i < end; i++) {
    double4 v = pos[i];
    int j = i - 1;
    while (j >= beg && pos[j].w > v.w) { pos[j + 1] = pos[j]; j--; }
    pos[j + 1] = v;
  }
  for (int i = beg; i < end; i++) rel[i] = make_rel_cell(g, c, pos[i]);
}

// shadow of one cell layer recomputed from the master table (ghost layers after a halo refresh)
__global__ void k_rel_layer(Grid g, const double4* __restrict__ pos, float4* __restrict__ rel,
                            const int* __restrict__ cs, int layer) {
  long long per = (long long)g.ny * g.nz;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= per) return;
  long long c = (long long)layer * per + t;
  for (int i = cs[c]; i < cs[c + 1]; i++) rel[i] = make_rel_cell(g, c, pos[i]);
}

// slab mode: the six layer offsets and the error flags in one staging vector (one D2H copy)
__global__ void k_gather_layout(Grid g, const int* __restrict__ cs, const int* __restrict__ halo_cnt,
                                int* __restrict__ out, volatile uint32_t* peer_info, uint32_t rebuild, int half) {
  long long per = (long long)g.ny * g.nz;
  int k = threadIdx.x;
  if (k == 9 && peer_info) {
    // SlabLink::info of the right neighbour: base slot of my right ghost layer, the half of the ping-pong table that
    // holds it, and -- last, behind a fence -- the number of the rebuild this comes from (everything this stream did
    // before, the rebuilt tables included, is then visible to whoever sees the number)
    peer_info[0] = (uint32_t)cs[(long long)(g.nlx - 1) * per];
    peer_info[2] = (uint32_t)half;
    __threadfence_system();
    peer_info[1] = rebuild;
    __threadfence_system();
  }
  if (k < 6) {
    long long offs = (k == 0) ? 0 : (k == 1) ? per : (k == 2) ? 2 * per : (k == 3) ? (long long)(g.nlx - 2) * per
                   : (k == 4) ? (long long)(g.nlx - 1) * per : (long long)g.nlx * per;
    out[k] = cs[offs];
  } else if (k < 8) {
    out[k] = halo_cnt[k - 6];
  } else if (k == 8) {
    out[8] |= halo_cnt[2];           // error bits are sticky until the host reads them
  }
}

