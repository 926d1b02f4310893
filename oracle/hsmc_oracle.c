/* oracle/hsmc_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * See hsmc_oracle.h.  Each routine cites the reference file:line it follows
 * (paths relative to /root/reference/src).  Double precision, unfused, same
 * operation order as the reference; build with -ffp-contract=off.
 */
#include "hsmc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct orc_sys {
  int N;
  double lx, ly, lz;
  double *c; /* N x {id,x,y,z}: sim_info.h:20 */
  int num_x, num_y, num_z, num_tot, max_part;
  double size_x, size_y, size_z;
  int *pc;    /* occupancy matrix, row = [count, ids...]  cell_list.h:16 */
  int *neigh; /* 27-stencil table                          cell_list.h:17 */
  int64_t pm, apm, rpm, vm, avm, rvm;
  int status;
};

#define X(s, i) ((s)->c[4 * (size_t)(i) + 1])
#define Y(s, i) ((s)->c[4 * (size_t)(i) + 2])
#define Z(s, i) ((s)->c[4 * (size_t)(i) + 3])

/* ---- sim_info.c:32-71 ---- */
void orc_box_from_lattice(int type, int nx, int ny, int nz, double rho, double *box4) {
  int ppc = (type == 1) ? 1 : 4;
  double cell_vol = ppc / rho;
  double cell_size = pow(cell_vol, 1. / 3.);
  box4[0] = nx * cell_size;
  box4[1] = ny * cell_size;
  box4[2] = nz * cell_size;
  box4[3] = nx * ny * nz * cell_vol;
}

int orc_lattice_count(int type, int nx, int ny, int nz) { return nx * ny * nz * ((type == 1) ? 1 : 4); }

/* ---- sim_info.c:125-166 ---- */
void orc_lattice_fill(int type, int nx, int ny, int nz, double rho, double *c) {
  int ppc = (type == 1) ? 1 : 4;
  double aa = pow(ppc / rho, 1. / 3.);
  int id = 0;
#define ADD(xx, yy, zz) do { c[4*(size_t)id] = id; c[4*(size_t)id+1] = (xx); c[4*(size_t)id+2] = (yy); c[4*(size_t)id+3] = (zz); id++; } while (0)
  for (int ii = 0; ii < nx; ii++)
    for (int jj = 0; jj < ny; jj++)
      for (int kk = 0; kk < nz; kk++) {
        ADD(ii * aa, jj * aa, kk * aa);
        if (ppc == 4) {
          ADD((ii + 0.5) * aa, (jj + 0.5) * aa, kk * aa);
          ADD((ii + 0.5) * aa, jj * aa, (kk + 0.5) * aa);
          ADD(ii * aa, (jj + 0.5) * aa, (kk + 0.5) * aa);
        }
      }
#undef ADD
}

/* ---- cell_list.c:237-249 (note the reference's ix*num_x*num_x + iy*num_y + iz:
        a bijection only on cubic grids) ---- */
static int cell_of_xyz(const orc_sys *s, double x, double y, double z) {
  return (int)(x / s->size_x) * s->num_x * s->num_x + (int)(y / s->size_y) * s->num_y +
         (int)(z / s->size_z);
}
int orc_cell_of(const orc_sys *s, int idx) { return cell_of_xyz(s, X(s, idx), Y(s, idx), Z(s, idx)); }

/* ---- cell_list.c:224-233 ---- */
static void cell_check(orc_sys *s, int cell) {
  int n = s->pc[(size_t)cell * s->max_part];
  if (n < 0 || n > s->max_part - 1) s->status = 1;
}

/* ---- cell_list.c:92-131 (init=false branch: sizes only) + :142-175 ---- */
static void cell_list_new(orc_sys *s) {
  s->size_x = s->lx / s->num_x;
  s->size_y = s->ly / s->num_y;
  s->size_z = s->lz / s->num_z;
  if (s->size_x < 1.0 || s->size_y < 1.0 || s->size_z < 1.0) { s->status = 2; return; }
  for (int i = 0; i < s->num_tot; i++) {
    s->pc[(size_t)i * s->max_part] = 0;
    for (int j = 1; j < s->max_part; j++) s->pc[(size_t)i * s->max_part + j] = -1;
  }
  for (int i = 0; i < s->N; i++) {
    int row = orc_cell_of(s, i);
    if (row < 0 || row >= s->num_tot) { s->status = 3; return; }
    int n = ++s->pc[(size_t)row * s->max_part];
    cell_check(s, row);
    if (s->status) return;
    s->pc[(size_t)row * s->max_part + n] = i;
  }
}

/* ---- cell_list.c:253-301 ---- */
static void neigh_init(orc_sys *s) {
  for (int rx = 0; rx < s->num_x; rx++)
    for (int ry = 0; ry < s->num_y; ry++)
      for (int rz = 0; rz < s->num_z; rz++) {
        int ref = rx * s->num_x * s->num_x + ry * s->num_y + rz;
        int cnt = 0;
        for (int ii = -1; ii < 2; ii++)
          for (int jj = -1; jj < 2; jj++)
            for (int kk = -1; kk < 2; kk++) {
              int ix = rx + ii, iy = ry + jj, iz = rz + kk;
              if (ix > s->num_x - 1) ix -= s->num_x; else if (ix < 0) ix += s->num_x;
              if (iy > s->num_y - 1) iy -= s->num_y; else if (iy < 0) iy += s->num_y;
              if (iz > s->num_z - 1) iz -= s->num_z; else if (iz < 0) iz += s->num_z;
              s->neigh[(size_t)ref * 27 + cnt++] = ix * s->num_x * s->num_x + iy * s->num_y + iz;
            }
      }
}

orc_sys *orc_create(int N, double lx, double ly, double lz, const double *conf4, double neigh_dr,
                    int max_part) {
  orc_sys *s = (orc_sys *)calloc(1, sizeof(*s));
  s->N = N; s->lx = lx; s->ly = ly; s->lz = lz;
  s->c = (double *)malloc(sizeof(double) * 4 * (size_t)N);
  memcpy(s->c, conf4, sizeof(double) * 4 * (size_t)N);
  /* cell_list.c:97-117 */
  s->num_x = (int)floor(lx / neigh_dr);
  s->num_y = (int)floor(ly / neigh_dr);
  s->num_z = (int)floor(lz / neigh_dr);
  s->num_tot = s->num_x * s->num_y * s->num_z;
  if (s->num_tot < 27) { s->num_x = s->num_y = s->num_z = 3; s->num_tot = 27; }
  s->max_part = max_part;
  /* the reference's index formula can exceed num_tot on non-cubic grids: size for it */
  size_t rows = (size_t)s->num_tot;
  s->pc = (int *)malloc(sizeof(int) * rows * (size_t)max_part);
  s->neigh = (int *)malloc(sizeof(int) * rows * 27);
  if (s->num_x != s->num_y || s->num_y != s->num_z) s->status = 4; /* parity undefined */
  else { cell_list_new(s); neigh_init(s); }
  return s;
}

void orc_destroy(orc_sys *s) {
  if (!s) return;
  free(s->c); free(s->pc); free(s->neigh); free(s);
}

int orc_N(const orc_sys *s) { return s->N; }
int orc_status(const orc_sys *s) { return s->status; }
void orc_cells(const orc_sys *s, int *num3, double *size3) {
  num3[0] = s->num_x; num3[1] = s->num_y; num3[2] = s->num_z;
  size3[0] = s->size_x; size3[1] = s->size_y; size3[2] = s->size_z;
}
void orc_get_conf(const orc_sys *s, double *c) { memcpy(c, s->c, sizeof(double) * 4 * (size_t)s->N); }
void orc_set_conf(orc_sys *s, const double *c) {
  memcpy(s->c, c, sizeof(double) * 4 * (size_t)s->N);
  cell_list_new(s);
}

/* ---- moves.c:400-431 ---- */
double orc_compute_dist(const orc_sys *s, int i, int j, double sf) {
  double lx = s->lx * sf, ly = s->ly * sf, lz = s->lz * sf;
  double lx_2 = lx / 2.0, ly_2 = ly / 2.0, lz_2 = lz / 2.0;
  double dx = (X(s, i) - X(s, j)) * sf;
  double dy = (Y(s, i) - Y(s, j)) * sf;
  double dz = (Z(s, i) - Z(s, j)) * sf;
  if (dx > lx_2) dx -= lx; else if (dx < -lx_2) dx += lx;
  if (dy > ly_2) dy -= ly; else if (dy < -ly_2) dy += ly;
  if (dz > lz_2) dz -= lz; else if (dz < -lz_2) dz += lz;
  return sqrt(dx * dx + dy * dy + dz * dz);
}

/* ---- moves.c:157-212 ---- */
int orc_check_overlap(const orc_sys *s, int idx, double sf) {
  int cell = orc_cell_of(s, idx);
  for (int ii = 0; ii < 27; ii++) {
    int nb = s->neigh[(size_t)cell * 27 + ii];
    const int *row = &s->pc[(size_t)nb * s->max_part];
    for (int jj = 1; jj <= row[0]; jj++) {
      int p = row[jj];
      double dr = orc_compute_dist(s, idx, p, sf);
      if (dr < 1.0 && p != idx) return 1;
    }
  }
  return 0;
}

int orc_trial_verdict(orc_sys *s, int idx, double x, double y, double z, double sf) {
  double ox = X(s, idx), oy = Y(s, idx), oz = Z(s, idx);
  X(s, idx) = x; Y(s, idx) = y; Z(s, idx) = z;
  int ov = orc_check_overlap(s, idx, sf);
  X(s, idx) = ox; Y(s, idx) = oy; Z(s, idx) = oz;
  return ov;
}

void orc_trial_verdicts(orc_sys *s, int n, const int *idx, const double *xyz, double sf, int *flags) {
  for (int i = 0; i < n; i++)
    flags[i] = orc_trial_verdict(s, idx[i], xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], sf);
}

void orc_overlap_all(const orc_sys *s, double sf, int *flags) {
  for (int i = 0; i < s->N; i++) flags[i] = orc_check_overlap(s, i, sf);
}

int orc_any_overlap(const orc_sys *s, double sf) {
  for (int i = 0; i < s->N; i++)
    if (orc_check_overlap(s, i, sf)) return 1;
  return 0;
}

/* ---- moves.c:215-226 ---- */
static void apply_pbc(orc_sys *s, int i) {
  if (X(s, i) > s->lx) X(s, i) -= s->lx; else if (X(s, i) < 0.0) X(s, i) += s->lx;
  if (Y(s, i) > s->ly) Y(s, i) -= s->ly; else if (Y(s, i) < 0.0) Y(s, i) += s->ly;
  if (Z(s, i) > s->lz) Z(s, i) -= s->lz; else if (Z(s, i) < 0.0) Z(s, i) += s->lz;
}

/* ---- cell_list.c:180-222 ---- */
static void cell_update(orc_sys *s, int cdel, int cadd, int p) {
  int *row = &s->pc[(size_t)cdel * s->max_part];
  int n = row[0], rm = n, shift = 0;
  row[0] -= 1;
  cell_check(s, cdel);
  for (int ii = 1; ii <= n; ii++) {
    if (row[ii] == p) { rm = ii; row[ii] = -1; shift = 1; }
    if (shift && ii > rm) { row[ii - 1] = row[ii]; row[ii] = -1; }
  }
  if (!shift) s->status = 5;
  row = &s->pc[(size_t)cadd * s->max_part];
  n = row[0];
  row[0] += 1;
  cell_check(s, cadd);
  if (!s->status) row[n + 1] = p;
}

/* ---- moves.c:27-80, with the four RNG draws supplied by the caller.
        u = raw/0xffffffff as in rng.c:29-31 ---- */
int orc_part_move_raw(orc_sys *s, int idx, uint32_t rx, uint32_t ry, uint32_t rz, double dr_max) {
  double r_x = (double)rx / (double)0xffffffffUL;
  double r_y = (double)ry / (double)0xffffffffUL;
  double r_z = (double)rz / (double)0xffffffffUL;
  double xo = X(s, idx), yo = Y(s, idx), zo = Z(s, idx);
  int cell_old = orc_cell_of(s, idx);
  X(s, idx) += (r_x - 0.5) * dr_max;
  Y(s, idx) += (r_y - 0.5) * dr_max;
  Z(s, idx) += (r_z - 0.5) * dr_max;
  apply_pbc(s, idx);
  int acc;
  if (orc_check_overlap(s, idx, 1.0)) {
    X(s, idx) = xo; Y(s, idx) = yo; Z(s, idx) = zo;
    s->rpm += 1; acc = 0;
  } else {
    int cell_new = orc_cell_of(s, idx);
    if (cell_new != cell_old) cell_update(s, cell_old, cell_new, idx);
    s->apm += 1; acc = 1;
  }
  s->pm += 1;
  return acc;
}

/* the same, for a whole list of trials (test harness convenience: one call per logged sweep) */
void orc_replay_moves(orc_sys *s, int n, const int *ids, const uint32_t *raw3, double dr_max, int *accepted) {
  for (int k = 0; k < n; k++)
    accepted[k] = orc_part_move_raw(s, ids[k], raw3[3 * k], raw3[3 * k + 1], raw3[3 * k + 2], dr_max);
}

void orc_counters(const orc_sys *s, int64_t *o) {
  o[0] = s->pm; o[1] = s->apm; o[2] = s->rpm; o[3] = s->vm; o[4] = s->avm; o[5] = s->rvm;
}
void orc_reset_counters(orc_sys *s) { s->pm = s->apm = s->rpm = s->vm = s->avm = s->rvm = 0; }

/* ---- moves.c:129-142: coordinates *= sf, PBC, cell_list_new ---- */
void orc_rescale(orc_sys *s, double sf, double lx, double ly, double lz) {
  s->lx = lx; s->ly = ly; s->lz = lz;
  for (int i = 0; i < s->N; i++) {
    X(s, i) *= sf; Y(s, i) *= sf; Z(s, i) *= sf;
    apply_pbc(s, i);
  }
  cell_list_new(s);
}

/* ---- MT19937 as used through GSL by rng.c (published algorithm; seed 0 -> 4357) ---- */
struct orc_mt { uint32_t mt[624]; int mti; };
orc_mt *orc_mt_create(unsigned long seed) {
  orc_mt *m = (orc_mt *)malloc(sizeof(*m));
  if (seed == 0) seed = 4357;
  m->mt[0] = (uint32_t)seed;
  for (int i = 1; i < 624; i++) m->mt[i] = 1812433253u * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
  m->mti = 624;
  return m;
}
void orc_mt_destroy(orc_mt *m) { free(m); }
uint32_t orc_mt_raw(orc_mt *m) {
  if (m->mti >= 624) {
    for (int k = 0; k < 624; k++) {
      uint32_t y = (m->mt[k] & 0x80000000u) | (m->mt[(k + 1) % 624] & 0x7fffffffu);
      m->mt[k] = m->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    m->mti = 0;
  }
  uint32_t k = m->mt[m->mti++];
  k ^= k >> 11; k ^= (k << 7) & 0x9d2c5680u; k ^= (k << 15) & 0xefc60000u; k ^= k >> 18;
  return k;
}
double orc_mt_double(orc_mt *m) { return (double)orc_mt_raw(m) / (double)0xffffffffUL; }
int orc_mt_int(orc_mt *m, int n) {
  unsigned long scale = 0xffffffffUL / (unsigned long)n, k;
  do { k = orc_mt_raw(m) / scale; } while (k >= (unsigned long)n);
  return (int)k;
}

/* ---- nvt.c:201-209 + moves.c:38-43 draw order: index, then x, y, z ---- */
void orc_sweep_nvt(orc_sys *s, orc_mt *m, int n_sweeps, double dr_max) {
  for (int sw = 0; sw < n_sweeps; sw++)
    for (int i = 0; i < s->N; i++) {
      int idx = orc_mt_int(m, s->N);
      uint32_t rx = orc_mt_raw(m), ry = orc_mt_raw(m), rz = orc_mt_raw(m);
      orc_part_move_raw(s, idx, rx, ry, rz, dr_max);
    }
}

/* ---- compute_widom_chem_pot.c:82-160 ---- */
static int widom_overlap(const orc_sys *s, double rx, double ry, double rz) {
  int cell = cell_of_xyz(s, rx, ry, rz);
  double lx_2 = s->lx / 2.0, ly_2 = s->ly / 2.0, lz_2 = s->lz / 2.0;
  for (int ii = 0; ii < 27; ii++) {
    int nb = s->neigh[(size_t)cell * 27 + ii];
    const int *row = &s->pc[(size_t)nb * s->max_part];
    for (int jj = 1; jj <= row[0]; jj++) {
      int p = row[jj];
      double dx = rx - X(s, p), dy = ry - Y(s, p), dz = rz - Z(s, p);
      if (dx > lx_2) dx -= s->lx; else if (dx < -lx_2) dx += s->lx;
      if (dy > ly_2) dy -= s->ly; else if (dy < -ly_2) dy += s->ly;
      if (dz > lz_2) dz -= s->lz; else if (dz < -lz_2) dz += s->lz;
      if (sqrt(dx * dx + dy * dy + dz * dz) < 1.0) return 1;
    }
  }
  return 0;
}

void orc_widom_verdicts(const orc_sys *s, int M, const double *xyz, int *flags) {
  for (int i = 0; i < M; i++) flags[i] = widom_overlap(s, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

/* ---- compute_widom_chem_pot.c:44-80: r = u*L with u = raw/0xffffffff ---- */
int64_t orc_widom_count_raw(const orc_sys *s, int64_t M, const uint32_t *raw3) {
  int64_t wtest = 0;
  for (int64_t i = 0; i < M; i++) {
    double rx = ((double)raw3[3 * i] / (double)0xffffffffUL) * s->lx;
    double ry = ((double)raw3[3 * i + 1] / (double)0xffffffffUL) * s->ly;
    double rz = ((double)raw3[3 * i + 2] / (double)0xffffffffUL) * s->lz;
    if (!widom_overlap(s, rx, ry, rz)) wtest++;
  }
  return wtest;
}

/* ---- compute_rdf.c:76,104,110-128 ---- */
int orc_rdf_nn(double dr, double rmax) { return (int)((rmax - 1.0) / dr); }
void orc_rdf_counts(const orc_sys *s, double dr_bin, double rmax_in, int nn, uint64_t *counts) {
  (void)rmax_in;
  double rmax = dr_bin * nn + 1.0; /* rdf_hist_init re-derives the cutoff */
  for (int i = 0; i < nn; i++) counts[i] = 0;
  for (int ii = 0; ii < s->N; ii++)
    for (int jj = ii + 1; jj < s->N; jj++) {
      double dr = orc_compute_dist(s, ii, jj, 1.0);
      if (dr < rmax) {
        int bin = (int)((dr - 1.0) / dr_bin);
        counts[bin] += 1; /* reference adds 2.0 per pair: hist = 2*count */
      }
    }
}

/* ---- compute_press.c:36,88,116,123-165 ---- */
int orc_pressv_nn(double dr) { return (int)((1.05 - 1.0) / dr); }
void orc_pressv_counts(const orc_sys *s, double dr_bin, int nn, uint64_t *counts) {
  double rmax = dr_bin * nn + 1.0;
  for (int i = 0; i < nn; i++) counts[i] = 0;
  for (int ii = 0; ii < s->N; ii++) {
    int cell = orc_cell_of(s, ii);
    for (int jj = 0; jj < 27; jj++) {
      int nb = s->neigh[(size_t)cell * 27 + jj];
      const int *row = &s->pc[(size_t)nb * s->max_part];
      for (int kk = 1; kk <= row[0]; kk++) {
        int p = row[kk];
        double dr = orc_compute_dist(s, ii, p, 1.0);
        if (dr < rmax && p > ii) {
          int bin = (int)((dr - 1.0) / dr_bin);
          counts[bin] += 1;
        }
      }
    }
  }
}

/* ---- compute_order_parameter.c:84-229 ----
 * gsl_sf_legendre_sphPlm restated in gsl_shim/gsl/gsl_sf_legendre.h (GSL is a third-party
 * dependency absent from the reference tree; not on the bit-exact surface). */
#include "gsl_shim/gsl/gsl_sf_legendre.h"
double orc_order_param(const orc_sys *s, int l, double rmax) {
  const int tlp1 = 2 * l + 1;
  double *qlm2 = (double *)malloc(sizeof(double) * (size_t)tlp1);
  const double lx_2 = s->lx / 2.0, ly_2 = s->ly / 2.0, lz_2 = s->lz / 2.0;   /* :108-110 */
  double ql_ave = 0.0;
  for (int ref = 0; ref < s->N; ref++) {
    int cell = orc_cell_of(s, ref);
    for (int mm = 0; mm < l + 1; mm++) {                                     /* :154-226 */
      int num_bonds = 0;
      double qr = 0., qi = 0., qmr = 0., qmi = 0.;
      for (int ii = 0; ii < 27; ii++) {
        int nb = s->neigh[(size_t)cell * 27 + ii];
        const int *row = &s->pc[(size_t)nb * s->max_part];
        for (int jj = 1; jj <= row[0]; jj++) {
          int p = row[jj];
          double dx = X(s, ref) - X(s, p);
          double dy = Y(s, ref) - Y(s, p);
          double dz = Z(s, ref) - Z(s, p);
          if (dx > lx_2) dx -= s->lx; else if (dx < -lx_2) dx += s->lx;
          if (dy > ly_2) dy -= s->ly; else if (dy < -ly_2) dy += s->ly;
          if (dz > lz_2) dz -= s->lz; else if (dz < -lz_2) dz += s->lz;
          double dr = sqrt(dx * dx + dy * dy + dz * dz);
          if (p != ref && dr <= rmax) {
            double phi = atan2(dy, dx);
            if (phi < 0) phi += 2. * M_PI;
            num_bonds++;
            double plm = gsl_sf_legendre_sphPlm(l, mm, dz / dr);
            qr += plm * cos(mm * phi);
            qi += plm * sin(mm * phi);
            if (mm % 2 != 0) plm *= -1.0;
            qmr += plm * cos(-mm * phi);
            qmi += plm * sin(-mm * phi);
          }
        }
      }
      if (num_bonds != 0) { qr /= num_bonds; qi /= num_bonds; qmr /= num_bonds; qmi /= num_bonds; }
      qlm2[l + mm] = qr * qr + qi * qi;
      qlm2[l - mm] = qmr * qmr + qmi * qmi;
    }
    double q = 0.0;                                                          /* :120-126 */
    for (int jj = 0; jj < tlp1; jj++) q += qlm2[jj];
    q *= 4 * M_PI / tlp1;
    ql_ave += sqrt(q) / s->N;                                                /* :92-96 */
  }
  free(qlm2);
  return ql_ave;
}

/* ---- compute_press.c:211,228-235,250-271 ---- */
int orc_presst_nn(double dxi, double xi_max) { return (int)(xi_max / dxi); }
void orc_presst_flags(const orc_sys *s, double dxi, int nn, int *flags, double *sf_out) {
  for (int ii = 0; ii < nn; ii++) {
    double xi = (ii + 1) * dxi;
    double sf = pow(1 - xi, 1. / 3.);
    if (sf_out) sf_out[ii] = sf;
    flags[ii] = orc_any_overlap(s, sf) ? 0 : 1; /* hist += 1.0 when NO overlap */
  }
}

/* ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11; Random123 constants) ---- */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
