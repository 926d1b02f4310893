/* oracle/ref_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin C entry points over the UNMODIFIED reference sources (compiled where they
 * lie under /root/reference/src by oracle/Makefile into oracle/_ref/libhsmc_ref.so).
 * The reference keeps all state in file-scope globals and void f(void) routines
 * (SURVEY.md 0.2); this file only sets those globals up from explicit arguments,
 * calls the reference's own routines, and copies results out.  No reference
 * arithmetic is re-implemented here.
 *
 * compute_rdf.c / compute_press.c / compute_widom_chem_pot.c / compute_order_parameter.c are compiled with
 * -Dstatic= (oracle/Makefile) so that their histogram arrays are linkable.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdbool.h>
#include <math.h>

#include "read_input.h"
#include "sim_info.h"
#include "cell_list.h"
#include "moves.h"
#include "rng.h"
#include "nvt.h"
#include "npt.h"
#include "compute_rdf.h"
#include "compute_press.h"
#include "compute_widom_chem_pot.h"
#include "compute_order_parameter.h"

/* scripted RNG hooks defined (weak) in the GSL shim, see gsl_shim/gsl/gsl_rng.h */
extern const unsigned int *hsmc_shim_script;
extern size_t hsmc_shim_script_len;
extern size_t hsmc_shim_script_pos;

/* exposed by -Dstatic= */
extern int rdf_hist_nn;
extern double *rdf_rr, *rdf_hist;
extern int pressv_hist_nn;
extern double pressv_rmax;
extern double *pressv_rr, *pressv_hist;
extern int presst_hist_nn;
extern double *presst_xi, *presst_hist;
extern int wtest;
extern double mu;
extern double ql_ave;          /* compute_order_parameter.c:19 */

static int harness_live = 0;

static void input_defaults(void) {
  /* mirrors the defaults block of read_input.c:71-115 only for the fields the
     hot path reads; everything else zero */
  memset(&G_IN, 0, sizeof(G_IN));
  G_IN.rho = 0.5;
  G_IN.nx = G_IN.ny = G_IN.nz = 3;
  G_IN.type = 1;
  G_IN.neigh_dr = 1.0;
  G_IN.neigh_max_part = 10;
  G_IN.dr_max = 0.05;
  G_IN.output_int = 1;
  G_IN.dv_max = 0.001;
  G_IN.pressv_dr = 0.01;
  G_IN.presst_dxi = 0.0001;
  G_IN.presst_xi_max = 0.002;
  G_IN.mu_insertions = 100;
  G_IN.rdf_dr = 0.01;
  G_IN.rdf_rmax = 10;
  G_IN.rdf_samples = 100;
  G_IN.ql_order = 6;
  G_IN.ql_rmax = 1.5;
}

/* Reference lattice start: sim_box_init + part_alloc + part_init + rng_init +
   cell_list_init, i.e. hs_nvt() lines nvt.c:35-55 without the run. */
int ref_setup_lattice(int type, int nx, int ny, int nz, double rho, double neigh_dr,
                      int max_part, unsigned long seed) {
  if (harness_live) return -1;
  input_defaults();
  G_IN.type = type; G_IN.nx = nx; G_IN.ny = ny; G_IN.nz = nz; G_IN.rho = rho;
  G_IN.neigh_dr = neigh_dr; G_IN.neigh_max_part = max_part; G_IN.seed = seed;
  sim_box_init(type, nx, ny, nz, rho);
  part_alloc();
  part_init();
  rng_init();
  cell_list_init(true);
  reset_moves_counters();
  harness_live = 1;
  return 0;
}

/* Arbitrary N and box: the box_info/p_info structs are injected through the
   reference's own restart (de)serialisers (sim_info.c:219-236). */
int ref_setup_box(int N, double lx, double ly, double lz, const double *conf4,
                  double neigh_dr, int max_part, unsigned long seed) {
  if (harness_live) return -1;
  input_defaults();
  box_info b;
  memset(&b, 0, sizeof(b));
  b.vol = lx * ly * lz;
  b.lx = lx; b.ly = ly; b.lz = lz;
  b.min_size = fmin(lx, fmin(ly, lz));
  b.cell_size = 1.0;
  b.cell_x = N; b.cell_y = 1; b.cell_z = 1; b.cell_type = 1;
  char *buf = NULL; size_t len = 0;
  FILE *f = open_memstream(&buf, &len);
  fwrite(&b, sizeof(b), 1, f);
  fclose(f);
  f = fmemopen(buf, len, "rb");
  sim_box_info_read(f);
  fclose(f);
  free(buf);
  part_alloc(); /* N = cell_x*cell_y*cell_z*1 */
  config c = part_config_get();
  memcpy(c, conf4, sizeof(double) * 4 * (size_t)N);
  G_IN.rho = N / b.vol;
  G_IN.type = 1; G_IN.nx = N; G_IN.ny = 1; G_IN.nz = 1;
  G_IN.neigh_dr = neigh_dr; G_IN.neigh_max_part = max_part; G_IN.seed = seed;
  rng_init();
  cell_list_init(true);
  reset_moves_counters();
  harness_live = 1;
  return 0;
}

void ref_teardown(void) {
  if (!harness_live) return;
  hsmc_shim_script = NULL;
  part_free();
  cell_list_free();
  rng_free();
  harness_live = 0;
}

int ref_N(void) { return part_info_get().NN; }

void ref_box(double *out4) {
  box_info b = sim_box_info_get();
  out4[0] = b.lx; out4[1] = b.ly; out4[2] = b.lz; out4[3] = b.vol;
}

void ref_cells(int *num3, double *size3) {
  cl_info nl = get_cell_list_info();
  num3[0] = nl.num_x; num3[1] = nl.num_y; num3[2] = nl.num_z;
  size3[0] = nl.size_x; size3[1] = nl.size_y; size3[2] = nl.size_z;
}

void ref_get_conf(double *out4) {
  memcpy(out4, part_config_get(), sizeof(double) * 4 * (size_t)ref_N());
}

void ref_set_conf(const double *in4) {
  memcpy(part_config_get(), in4, sizeof(double) * 4 * (size_t)ref_N());
  cell_list_new();
}

void ref_set_moves(double dr_max, double dv_max, double press) {
  G_IN.dr_max = dr_max; G_IN.dv_max = dv_max; G_IN.press = press;
}

double ref_get_rho(void) { return G_IN.rho; }

/* ---- scripted RNG ---- */
void ref_script(const unsigned int *raw, size_t n) {
  hsmc_shim_script = raw; hsmc_shim_script_len = n; hsmc_shim_script_pos = 0;
}
size_t ref_script_pos(void) { return hsmc_shim_script_pos; }

/* ---- moves.c ---- */

/* verdict of check_overlap() for particle idx placed at (x,y,z), exactly as
   part_move() evaluates it (moves.c:52-60): coordinates overwritten in place, the
   cell list NOT updated, the particle restored afterwards. */
int ref_trial_verdict(int idx, double x, double y, double z, double sf) {
  config c = part_config_get();
  double ox = c[idx][1], oy = c[idx][2], oz = c[idx][3];
  c[idx][1] = x; c[idx][2] = y; c[idx][3] = z;
  int ov = check_overlap(idx, sf, sf, sf) ? 1 : 0;
  c[idx][1] = ox; c[idx][2] = oy; c[idx][3] = oz;
  return ov;
}

void ref_trial_verdicts(int n, const int *idx, const double *xyz, double sf, int *flags) {
  for (int i = 0; i < n; i++)
    flags[i] = ref_trial_verdict(idx[i], xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], sf);
}

/* check_overlap(ii, sf, sf, sf) for every particle, no early break */
void ref_overlap_all(double sf, int *flags) {
  int N = ref_N();
  for (int i = 0; i < N; i++) flags[i] = check_overlap(i, sf, sf, sf) ? 1 : 0;
}

/* the vol_move()/presst loop (moves.c:106-112): any overlap under scaling sf */
int ref_any_overlap(double sf) {
  int N = ref_N();
  for (int i = 0; i < N; i++)
    if (check_overlap(i, sf, sf, sf)) return 1;
  return 0;
}

double ref_compute_dist(int i, int j, double sf) { return compute_dist(i, j, sf, sf, sf); }

void ref_part_moves(long n) { for (long i = 0; i < n; i++) part_move(); }
void ref_vol_move(void) { vol_move(); }
void ref_sweep_nvt(int n) { for (int i = 0; i < n; i++) sweep_nvt(); }
void ref_sweep_npt(int n) { for (int i = 0; i < n; i++) sweep_npt(); }
void ref_reset_counters(void) { reset_moves_counters(); }
void ref_counters(int *out6) {
  get_moves_counters(&out6[0], &out6[1], &out6[2], &out6[3], &out6[4], &out6[5]);
}
int ref_cell_of(int idx) { return cell_part_idx(idx); }

/* ---- compute_widom_chem_pot.c ---- */
int ref_widom(int M, double *mu_out) {
  G_IN.mu_insertions = M;
  widom_insertion();
  if (mu_out) *mu_out = mu;
  return wtest;
}

/* per-insertion verdicts through the reference's own widom_rand_pos /
   widom_check_overlap (compute_widom_chem_pot.c:73-120) */
void ref_widom_verdicts(int M, int *flags) {
  G_IN.mu_insertions = 0;
  widom_insertion(); /* binds the file-scope cell-list handle, performs 0 insertions */
  for (int i = 0; i < M; i++) {
    widom_rand_pos();
    flags[i] = widom_check_overlap() ? 1 : 0;
  }
}

/* ---- compute_rdf.c: raw (un-normalised) histogram, rdf_hist_compute only ---- */
int ref_rdf_hist(double dr, double rmax, double *hist_out, int cap) {
  G_IN.rdf_dr = dr; G_IN.rdf_rmax = rmax;
  rdf_hist_alloc();
  rdf_hist_init();
  rdf_hist_compute();
  int nn = rdf_hist_nn;
  for (int i = 0; i < nn && i < cap; i++) hist_out[i] = rdf_hist[i];
  rdf_hist_free();
  return nn;
}

/* ---- compute_press.c ---- */
int ref_pressv_hist(double dr, double *hist_out, int cap) {
  G_IN.pressv_dr = dr;
  pressv_rmax = 1.05;
  pressv_hist_alloc();
  pressv_hist_init();
  pressv_compute_hist();
  int nn = pressv_hist_nn;
  for (int i = 0; i < nn && i < cap; i++) hist_out[i] = pressv_hist[i];
  pressv_hist_free();
  return nn;
}

int ref_presst_hist(double dxi, double xi_max, double *hist_out, double *xi_out, int cap) {
  G_IN.presst_dxi = dxi; G_IN.presst_xi_max = xi_max;
  presst_hist_alloc();
  presst_hist_init();
  presst_compute_hist();
  int nn = presst_hist_nn;
  for (int i = 0; i < nn && i < cap; i++) { hist_out[i] = presst_hist[i]; xi_out[i] = presst_xi[i]; }
  presst_hist_free();
  return nn;
}

/* compute_order_parameter.c:84-97: global_ql_compute() = average over particles of the
   Steinhardt q_l (ql_compute / qlm2_compute, :99-229) with cutoff rmax (the caller keeps it
   <= the neighbour-list cell edge, as compute_op(init) does at :29-40) */
double ref_order_param(int l, double rmax) {
  G_IN.ql_order = l;
  G_IN.ql_rmax = rmax;
  ql_alloc();
  global_ql_compute();
  ql_free();
  return ql_ave;
}


/* io_config.c:134-191: the reference's own snapshot writer, into the current directory
   (config_%06d.dat.gz; `samples_per_file` samples are appended to one file).  Pins the
   byte format the parallel writer of the host driver (hs_fastio.c) must reproduce. */
#include "io_config.h"
void ref_write_config(int sweep, int samples_per_file) {
  G_IN.config_write = 1;
  G_IN.config_samples = samples_per_file;
  G_IN.sweep_eq = 1; G_IN.sweep_stat = 1;
  write_config(sweep);
}
