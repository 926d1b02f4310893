/* hs_rng.c -- MT19937 (Matsumoto & Nishimura), see hs_rng.h */
#include "hs_rng.h"

void hs_rng_seed(hs_rng *r, unsigned long seed) {
  if (seed == 0) seed = 4357;   /* GSL's substitution for seed 0 */
  r->mt[0] = seed & 0xffffffffUL;
  for (int i = 1; i < 624; i++)
    r->mt[i] = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
  r->mti = 624;
}

static void refill(hs_rng *r) {
  unsigned long *mt = r->mt;
  for (int k = 0; k < 624; k++) {
    unsigned long y = (mt[k] & 0x80000000UL) | (mt[(k + 1) % 624] & 0x7fffffffUL);
    mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
  }
  r->mti = 0;
}

uint32_t hs_rng_raw(hs_rng *r) {
  if (r->mti >= 624) refill(r);
  unsigned long k = r->mt[r->mti++];
  k ^= k >> 11;
  k ^= (k << 7) & 0x9d2c5680UL;
  k ^= (k << 15) & 0xefc60000UL;
  k ^= k >> 18;
  return (uint32_t)(k & 0xffffffffUL);
}

double hs_rng_double(hs_rng *r) { return (double)hs_rng_raw(r) / (double)0xffffffffUL; }

int hs_rng_int(hs_rng *r, int n) {
  unsigned long scale = 0xffffffffUL / (unsigned long)n, k;
  do k = hs_rng_raw(r) / scale; while (k >= (unsigned long)n);
  return (int)k;
}

int hs_rng_write(const hs_rng *r, FILE *f) { return fwrite(r, 1, sizeof(*r), f) == sizeof(*r) ? 0 : 1; }
int hs_rng_read(hs_rng *r, FILE *f) { return fread(r, 1, sizeof(*r), f) == sizeof(*r) ? 0 : 1; }
