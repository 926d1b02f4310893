// Philox4x32-10 counter-based RNG (Salmon, Moraes, Dror, Shaw, SC'11), host + device.
// The device RNG of the B200 path: stateless, keyed by the run seed, counter =
// (global cell | insertion index, stream tag, sweep).  The reference draws from GSL's
// MT19937 (rng.c:21-36); this is a new generator, not a port -- parity is defined on
// identical trial points, which the oracle regenerates with its own Philox restatement.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HSMC_HD __host__ __device__ __forceinline__
#else
#define HSMC_HD inline
#endif

struct Philox4 {
  uint32_t v[4];
};

HSMC_HD uint32_t hsmc_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

HSMC_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                              uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = hsmc_mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = hsmc_mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

// stream tags (counter word 1, top byte)
#define HSMC_STREAM_MOVE 0u
#define HSMC_STREAM_WIDOM 1u
#define HSMC_STREAM_SHIFT 2u

// u in [0,1] with the reference's resolution: rng.c:29-31, u = raw / 0xffffffff
HSMC_HD double hsmc_u01_div(uint32_t raw) { return (double)raw / 4294967295.0; }

// The same value without a division: raw/(2^32-1) = raw*2^-32 * (1 + 2^-32 + 2^-64 + ...).
// q = raw*2^-32 is exact; the tail t = q*2^-32 + q*2^-64 carries everything that can
// influence the 53-bit rounding (the binary expansion of raw/(2^32-1) is the 32-bit
// pattern of raw repeated, so it never comes closer than 2^-33 ulp to a rounding
// boundary, while t is accurate to 2^-33 ulp of q... hence the EXHAUSTIVE device test
// hsmc_gpu_selftest_u01, which must report zero mismatches over all 2^32 inputs).
HSMC_HD double hsmc_u01_fast(uint32_t raw) {
  const double q = (double)raw * 2.3283064365386963e-10;                 // 2^-32, exact
  const double t = q * 2.3283064365386963e-10 + q * 5.421010862427522e-20;   // q*2^-32 + q*2^-64
  return q + t;
}

#ifdef HSMC_FAST_U01
HSMC_HD double hsmc_u01(uint32_t raw) { return hsmc_u01_fast(raw); }
#else
HSMC_HD double hsmc_u01(uint32_t raw) { return hsmc_u01_div(raw); }
#endif
