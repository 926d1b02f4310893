/* hs_opt.h -- one step of the acceptance-ratio tuning: the reference's secant update
 * (optimizer.c:45-58 for dr_max, 115-139 for dv_max) with its unguarded division fenced in.
 *
 * The reference computes x2 - (y2 - target) * (x2 - x1) / (y2 - y1) and uses the result as it comes: two samples
 * with the same acceptance ratio (common at high density with few sweeps per sample) give +-inf or NaN, which then
 * poisons every later step and the production run.  Here a step that is not a positive finite number keeps the
 * previous step instead.  Pure functions, so that tests/test_host_cpu.py can exercise them without a GPU. */
#ifndef HS_OPT_H
#define HS_OPT_H

static inline double hs_opt_secant(double x1, double y1, double x2, double y2, double target) {
  return x2 - (y2 - target) * (x2 - x1) / (y2 - y1);
}

static inline int hs_opt_usable(double v) { return v == v && v > 0.0 && v < 1e300; }

/* next dr_max from the samples (x1, y1), (x2, y2): optimizer.c:45-58 (cap at 1.0, sign flip, halving) + guard */
static inline double hs_opt_next_dr(double x1, double y1, double x2, double y2, double target) {
  double v = hs_opt_secant(x1, y1, x2, y2, target);
  if (v > 1.0) v = 1.0;
  else if (v <= 0.0) {
    v = -v;
    if (v > 1.0) v = x2 / 2;
  }
  return hs_opt_usable(v) ? v : x2;
}

/* next dv_max: optimizer.c:115-139 (sign flip, halving above 0.1) + guard */
static inline double hs_opt_next_dv(double x1, double y1, double x2, double y2, double target) {
  double v = hs_opt_secant(x1, y1, x2, y2, target);
  if (v <= 0.0) {
    v = -v;
    if (v > 0.1) v = x2 / 2;
  }
  return hs_opt_usable(v) ? v : x2;
}

#endif
