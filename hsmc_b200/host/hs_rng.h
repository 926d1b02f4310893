/* hs_rng.h -- host random numbers for the decisions that stay on the host (volume-move
 * size and Metropolis draw, optimizer-independent).  MT19937 with the reference's call
 * semantics (rng.c:21-36: u = raw/0xffffffff in [0,1], GSL uniform_int rejection rule)
 * and GSL's on-disk state layout so restart files stay interchangeable. */
#ifndef HS_RNG_H
#define HS_RNG_H
#include <stdio.h>
#include <stdint.h>

typedef struct hs_rng {
  unsigned long mt[624];
  int mti;
} hs_rng;   /* sizeof == 5000: what gsl_rng_fwrite emits for mt19937 */

void hs_rng_seed(hs_rng *r, unsigned long seed);
uint32_t hs_rng_raw(hs_rng *r);
double hs_rng_double(hs_rng *r);
int hs_rng_int(hs_rng *r, int n);
int hs_rng_write(const hs_rng *r, FILE *f);
int hs_rng_read(hs_rng *r, FILE *f);
#endif
