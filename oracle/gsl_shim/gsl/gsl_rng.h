/* oracle/gsl_shim/gsl/gsl_rng.h -- TEST INFRASTRUCTURE, not product code.
 *
 * GSL is a third-party dependency of the reference (src/Makefile:6,22, unpinned)
 * that is absent from this image.  This header restates, from the published
 * MT19937 algorithm (Matsumoto & Nishimura 1998, 2002 initialisation), exactly
 * the nine gsl_rng_* entry points that /root/reference/src/rng.c:23-52 calls, so
 * that the UNMODIFIED reference sources compile into oracle/_ref/.
 *
 * Pinned by known-answer: seed 4357 -> 1000th raw output 1186927261 (GSL's own
 * rng/test.c value for mt19937; cross-checked against numpy's independent
 * MT19937 with _legacy_seeding in tests/test_oracle_rng.py).
 *
 * Extension used only by the parity harness: a "scripted" mode in which
 * gsl_rng_get() replays a caller-supplied array of raw 32-bit outputs, so the
 * reference's own part_move()/widom_insertion() can be driven with the exact
 * trial points the GPU path generated.
 */
#ifndef HSMC_ORACLE_GSL_RNG_SHIM_H
#define HSMC_ORACLE_GSL_RNG_SHIM_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define HSMC_MT_N 624
#define HSMC_MT_M 397

typedef struct {
  unsigned long mt[HSMC_MT_N];
  int mti;
} hsmc_mt_state_t;

typedef struct {
  const char *name;
  unsigned long int max;
  unsigned long int min;
  size_t size;
} gsl_rng_type;

typedef struct {
  const gsl_rng_type *type;
  void *state;
} gsl_rng;

static const gsl_rng_type hsmc_mt19937_type = {"mt19937", 0xffffffffUL, 0,
                                               sizeof(hsmc_mt_state_t)};
static const gsl_rng_type *gsl_rng_mt19937 = &hsmc_mt19937_type;

/* ---- scripted replay (shared across translation units through weak symbols) */
__attribute__((weak)) const unsigned int *hsmc_shim_script = NULL;
__attribute__((weak)) size_t hsmc_shim_script_len = 0;
__attribute__((weak)) size_t hsmc_shim_script_pos = 0;

static inline void hsmc_mt_seed(hsmc_mt_state_t *st, unsigned long int s) {
  if (s == 0) s = 4357; /* GSL's default seed for mt19937 */
  st->mt[0] = s & 0xffffffffUL;
  for (int i = 1; i < HSMC_MT_N; i++) {
    st->mt[i] = (1812433253UL * (st->mt[i - 1] ^ (st->mt[i - 1] >> 30)) + (unsigned long)i);
    st->mt[i] &= 0xffffffffUL;
  }
  st->mti = HSMC_MT_N;
}

static inline unsigned long int hsmc_mt_get(hsmc_mt_state_t *st) {
  unsigned long k;
  unsigned long *const mt = st->mt;
#define HSMC_MAGIC(y) (((y) & 0x1UL) ? 0x9908b0dfUL : 0UL)
  if (st->mti >= HSMC_MT_N) {
    int kk;
    for (kk = 0; kk < HSMC_MT_N - HSMC_MT_M; kk++) {
      unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
      mt[kk] = mt[kk + HSMC_MT_M] ^ (y >> 1) ^ HSMC_MAGIC(y);
    }
    for (; kk < HSMC_MT_N - 1; kk++) {
      unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
      mt[kk] = mt[kk + (HSMC_MT_M - HSMC_MT_N)] ^ (y >> 1) ^ HSMC_MAGIC(y);
    }
    {
      unsigned long y = (mt[HSMC_MT_N - 1] & 0x80000000UL) | (mt[0] & 0x7fffffffUL);
      mt[HSMC_MT_N - 1] = mt[HSMC_MT_M - 1] ^ (y >> 1) ^ HSMC_MAGIC(y);
    }
    st->mti = 0;
  }
#undef HSMC_MAGIC
  k = mt[st->mti];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680UL;
  k ^= (k << 15) & 0xefc60000UL;
  k ^= (k >> 18);
  st->mti++;
  return k & 0xffffffffUL;
}

static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *T) {
  gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
  r->type = T;
  r->state = calloc(1, T->size);
  hsmc_mt_seed((hsmc_mt_state_t *)r->state, 0);
  return r;
}

static inline void gsl_rng_set(const gsl_rng *r, unsigned long int seed) {
  hsmc_mt_seed((hsmc_mt_state_t *)r->state, seed);
}

static inline unsigned long int gsl_rng_get(const gsl_rng *r) {
  if (hsmc_shim_script != NULL) {
    if (hsmc_shim_script_pos >= hsmc_shim_script_len) {
      fprintf(stderr, "gsl shim: scripted RNG stream exhausted at %zu\n", hsmc_shim_script_pos);
      abort();
    }
    return (unsigned long int)hsmc_shim_script[hsmc_shim_script_pos++];
  }
  return hsmc_mt_get((hsmc_mt_state_t *)r->state);
}

static inline unsigned long int gsl_rng_max(const gsl_rng *r) { return r->type->max; }

/* GSL: scale = range/n; k = (get-offset)/scale, redraw while k >= n */
static inline unsigned long int gsl_rng_uniform_int(const gsl_rng *r, unsigned long int n) {
  unsigned long int offset = r->type->min;
  unsigned long int range = r->type->max - offset;
  unsigned long int scale, k;
  if (n > range || n == 0) {
    fprintf(stderr, "gsl shim: invalid n in gsl_rng_uniform_int\n");
    abort();
  }
  scale = range / n;
  do {
    k = (gsl_rng_get(r) - offset) / scale;
  } while (k >= n);
  return k;
}

static inline void gsl_rng_free(gsl_rng *r) {
  if (!r) return;
  free(r->state);
  free(r);
}

static inline int gsl_rng_fwrite(FILE *stream, const gsl_rng *r) {
  return fwrite(r->state, 1, r->type->size, stream) == r->type->size ? 0 : 1;
}

static inline int gsl_rng_fread(FILE *stream, gsl_rng *r) {
  return fread(r->state, 1, r->type->size, stream) == r->type->size ? 0 : 1;
}

#endif
