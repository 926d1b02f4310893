/* oracle/gsl_shim/gsl/gsl_sf_legendre.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Restates gsl_sf_legendre_sphPlm(l, m, x) =
 *   sqrt((2l+1)/(4 pi) * (l-m)!/(l+m)!) * P_l^m(x)   (Condon-Shortley phase),
 * the one GSL special function the reference calls
 * (/root/reference/src/compute_order_parameter.c:202).  Not on the bit-exact
 * contract surface; pinned by q6(fcc) = 0.574524 and scipy's sph_harm_y.
 */
#ifndef HSMC_ORACLE_GSL_SF_LEGENDRE_SHIM_H
#define HSMC_ORACLE_GSL_SF_LEGENDRE_SHIM_H

#include <math.h>

static inline double gsl_sf_legendre_sphPlm(const int l, int m, const double x) {
  /* P_m^m by the closed form, then upward recurrence in l. */
  double pmm = 1.0;
  if (m > 0) {
    double somx2 = sqrt((1.0 - x) * (1.0 + x));
    double fact = 1.0;
    for (int i = 1; i <= m; i++) {
      pmm *= -fact * somx2;
      fact += 2.0;
    }
  }
  double plm;
  if (l == m) {
    plm = pmm;
  } else {
    double pmmp1 = x * (2.0 * m + 1.0) * pmm;
    if (l == m + 1) {
      plm = pmmp1;
    } else {
      double pll = 0.0;
      for (int ll = m + 2; ll <= l; ll++) {
        pll = (x * (2.0 * ll - 1.0) * pmmp1 - (ll + m - 1.0) * pmm) / (double)(ll - m);
        pmm = pmmp1;
        pmmp1 = pll;
      }
      plm = pll;
    }
  }
  /* normalisation: (l-m)!/(l+m)! = 1 / prod_{k=l-m+1}^{l+m} k */
  double ratio = 1.0;
  for (int k = l - m + 1; k <= l + m; k++) ratio /= (double)k;
  return sqrt((2.0 * l + 1.0) / (4.0 * M_PI) * ratio) * plm;
}

#endif
