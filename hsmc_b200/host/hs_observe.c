/* hs_observe.c -- observables of the drop-in host driver.
 *
 * The pair loops run on the device through the C ABI and come back as integer counts;
 * this file turns them into the reference's histograms (hist = 2.0 * count, as the
 * reference adds 2.0 per pair), normalises and writes the reference's file formats:
 *   press_virial.dat  compute_press.c:277-307     press_thermo.dat / density.dat  :311-359
 *   chem_pot.dat      compute_widom_chem_pot.c:164-182
 *   rdf_%06d.dat.gz   compute_rdf.c:155-205       order_param.dat  compute_order_parameter.c:232-253
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "hs_sim.h"

static void hist_alloc(hs_hist *h, int nn) {
  h->nn = nn;
  h->x = malloc(sizeof(double) * (size_t)(nn > 0 ? nn : 1));
  h->h = malloc(sizeof(double) * (size_t)(nn > 0 ? nn : 1));
  if (!h->x || !h->h) hs_die("Failed histogram allocation");
  h->live = true;
}

static void hist_free(hs_hist *h) {
  if (!h->live) return;
  free(h->x);
  free(h->h);
  h->live = false;
}

/* shell normalisation shared by the contact histogram and the rdf
   (compute_press.c:170-187, compute_rdf.c:133-150) */
static void shell_normalise(hs_hist *h, double rho, int N) {
  double dr = h->x[1] - h->x[0];
  for (int i = 0; i < h->nn; i++) {
    double r1 = h->x[i] - dr / 2., r2 = h->x[i] + dr / 2.;
    double bin_vol = (4. * M_PI / 3.) * (pow(r2, 3.) - pow(r1, 3.));
    h->h[i] = h->h[i] / (bin_vol * rho * N);
  }
}

static FILE *open_sample_file(const char *name, bool init, const char *what) {
  FILE *f = fopen(name, init ? "w" : "a");
  if (!f) {
    char msg[128];
    snprintf(msg, sizeof(msg), "Error while creating the file for the %s\n", what);
    perror(msg);
    exit(EXIT_FAILURE);
  }
  return f;
}

/* ---- pressure, virial route ---------------------------------------------------------- */
void hs_compute_pressv(hs_sim *s, bool init) {
  const hs_input *in = &s->in;
  if (init) {
    s->pressv_rmax = 1.05;   /* compute_press.c:36 */
    hist_alloc(&s->pressv, (int)((s->pressv_rmax - 1.0) / in->pressv_dr));
  }
  hs_hist *h = &s->pressv;
  for (int i = 0; i < h->nn; i++) h->x[i] = (i + 1. / 2.) * in->pressv_dr + 1.0;
  s->pressv_rmax = in->pressv_dr * h->nn + 1.0;
  uint64_t *cnt = calloc((size_t)(h->nn > 0 ? h->nn : 1), sizeof(uint64_t));
  /* the device refuses, with the reference's message, when its cells are narrower than
     the histogram range (compute_press.c:58-69) */
  hs_gpu_check(hsmc_gpu_contact_counts(s->gpu, in->pressv_dr, h->nn, cnt));
  for (int i = 0; i < h->nn; i++) h->h[i] = 2.0 * (double)cnt[i];
  free(cnt);
  shell_normalise(h, in->rho, s->part.NN);
  if (!HS_ROOT(s)) return;
  FILE *f = open_sample_file("press_virial.dat", init, "virial pressure");
  fprintf(f, "######################################\n");
  fprintf(f, "# Bins, volume, number of particles\n");
  fprintf(f, "######################################\n");
  fprintf(f, "%d %.8e %d\n", h->nn, s->box.vol, s->part.NN);
  fprintf(f, "###############################\n");
  fprintf(f, "# rr, rdf\n");
  fprintf(f, "###############################\n");
  for (int i = 0; i < h->nn; i++) fprintf(f, "%.8e %.8e\n", h->x[i], h->h[i]);
  fclose(f);
}

/* ---- pressure, thermodynamic route ------------------------------------------------------- */
void hs_compute_presst(hs_sim *s, bool init) {
  const hs_input *in = &s->in;
  if (init) hist_alloc(&s->presst, (int)(in->presst_xi_max / in->presst_dxi));
  hs_hist *h = &s->presst;
  double *sf = malloc(sizeof(double) * (size_t)(h->nn > 0 ? h->nn : 1));
  int *free_of_overlap = calloc((size_t)(h->nn > 0 ? h->nn : 1), sizeof(int));
  for (int i = 0; i < h->nn; i++) {
    h->x[i] = (i + 1) * in->presst_dxi;
    sf[i] = pow(1 - h->x[i], 1. / 3.);   /* compute_press.c:253-256 */
  }
  if (h->nn > 0) hs_gpu_check(hsmc_gpu_presst_flags(s->gpu, sf, h->nn, free_of_overlap));
  for (int i = 0; i < h->nn; i++) h->h[i] = free_of_overlap[i] ? 1.0 : 0.0;
  free(sf);
  free(free_of_overlap);
  if (!HS_ROOT(s)) return;
  FILE *f = open_sample_file("press_thermo.dat", init, "thermo pressure");
  fprintf(f, "######################################\n");
  fprintf(f, "# Bins, volume, number of particles\n");
  fprintf(f, "######################################\n");
  fprintf(f, "%d %.8e %d\n", h->nn, s->box.vol, s->part.NN);
  fprintf(f, "######################################\n");
  fprintf(f, "# abs(xi), exp(-beta*U)\n");
  fprintf(f, "######################################\n");
  for (int i = 0; i < h->nn; i++) fprintf(f, "%.8e %.8e\n", h->x[i], h->h[i]);
  fclose(f);
  if (in->press > 0) {
    f = open_sample_file("density.dat", init, "density");
    if (init) {
      fprintf(f, "######################################\n");
      fprintf(f, "# Density (each line is one sample)\n");
      fprintf(f, "######################################\n");
    }
    fprintf(f, "%.8e\n", in->rho);
    fclose(f);
  }
}

/* ---- chemical potential, Widom insertions -------------------------------------------------- */
void hs_compute_mu(hs_sim *s, bool init) {
  const hs_input *in = &s->in;
  int64_t wtest = 0;
  hs_gpu_check(hsmc_gpu_widom(s->gpu, s->mu_samples++, 0, in->mu_insertions, 1, &wtest));
  double mu = (wtest > 0) ? -log((double)wtest / in->mu_insertions) : 0.0;
  if (!HS_ROOT(s)) return;
  FILE *f = open_sample_file("chem_pot.dat", init, "chemical potential");
  if (init) {
    fprintf(f, "##################################################################################\n");
    fprintf(f, "# Chemical potenital (average over %d insertions, Fraction of accepted insertions)\n", in->mu_insertions);
    fprintf(f, "##################################################################################\n");
  }
  fprintf(f, "%.8e %.8e\n", mu, (double)wtest / in->mu_insertions);
  fclose(f);
}

/* ---- radial distribution function -------------------------------------------------------------- */
void hs_compute_rdf(hs_sim *s, bool init, int sweep) {
  hs_input *in = &s->in;
  if (init) {
    /* compute_rdf.c:39-52 (the reference compares against the LARGEST edge) */
    double lmax = s->box.lx;
    if (lmax < s->box.ly) lmax = s->box.ly;
    if (lmax < s->box.lz) lmax = s->box.lz;
    if (lmax < 2.0 * in->rdf_rmax) {
      in->rdf_rmax = lmax / 2.0;
      printf("WARNING: Cutoff for the rdf extraction reduced to %f in order to be consistent with minimum image convention\n",
             in->rdf_rmax);
    }
    hist_alloc(&s->rdf, (int)((in->rdf_rmax - 1.0) / in->rdf_dr));
  }
  hs_hist *h = &s->rdf;
  for (int i = 0; i < h->nn; i++) h->x[i] = (i + 1. / 2.) * in->rdf_dr + 1.0;
  in->rdf_rmax = in->rdf_dr * h->nn + 1.0;
  uint64_t *cnt = calloc((size_t)(h->nn > 0 ? h->nn : 1), sizeof(uint64_t));
  if (h->nn > 0 && s->mp.world == 1) {
    hs_gpu_check(hsmc_gpu_rdf_counts(s->gpu, in->rdf_dr, h->nn, cnt));
  } else if (h->nn > 0) {
    /* slab run: the pair histogram is all-pairs, so it is taken on a REPLICA of the configuration
       (one single-GPU handle per rank), each rank counting its share of the pair-tile triangle
       (SURVEY 8e); the shares meet in the shared scratch block and rank 0 adds them up */
    if (h->nn > HS_MP_SCRATCH) hs_die("rdf: more than %d bins in a multi-GPU run", HS_MP_SCRATCH);
    hs_gpu_pull(s);
    if (!s->gpu_rep) {
      hsmc_gpu_config cfg;
      memset(&cfg, 0, sizeof(cfg));
      const char *dev = getenv("HSMC_DEVICE");
      cfg.device = (dev ? atoi(dev) : 0) + s->mp.rank;
      cfg.world = 1;
      cfg.seed = (uint64_t)in->seed;
      cfg.cell_min = in->neigh_dr;
      double box[3] = {s->box.lx, s->box.ly, s->box.lz};
      hs_gpu_check(hsmc_gpu_create(&s->gpu_rep, &cfg, s->part.NN, box));
    }
    hs_gpu_check(hsmc_gpu_upload(s->gpu_rep, &s->conf[0][0], s->part.NN));
    hs_gpu_check(hsmc_gpu_rdf_counts_part(s->gpu_rep, in->rdf_dr, h->nn, s->mp.rank, s->mp.world, cnt));
    memcpy(hs_mp_scratch(&s->mp, s->mp.rank), cnt, sizeof(uint64_t) * (size_t)h->nn);
    hs_mp_barrier(&s->mp);
    for (int r = 0; r < s->mp.world; r++) {
      if (r == s->mp.rank) continue;
      const uint64_t *o = hs_mp_scratch(&s->mp, r);
      for (int i = 0; i < h->nn; i++) cnt[i] += o[i];
    }
    hs_mp_barrier(&s->mp);
  }
  for (int i = 0; i < h->nn; i++) h->h[i] = 2.0 * (double)cnt[i];
  free(cnt);
  shell_normalise(h, in->rho, s->part.NN);
  if (!HS_ROOT(s)) return;
  if (init && (double)(in->sweep_stat + in->sweep_eq) / (in->rdf_sample_int * in->rdf_samples) > 100000)
    printf("ERROR: Too many (> 100000) rdf files will be produced. Consider increasing number of samples per file\n");
  char name[32];
  snprintf(name, sizeof(name), "rdf_%06d.dat.gz", s->rdf_file_id);
  gzFile f = gzopen(name, s->rdf_samples_in_file == 0 ? "w" : "a");
  if (f == Z_NULL) { perror("Error while creating rdf file"); exit(EXIT_FAILURE); }
  gzprintf(f, "######################################\n");
  gzprintf(f, "# Sweep, Bins, volume, number of particles\n");
  gzprintf(f, "######################################\n");
  gzprintf(f, "%d %d %.8e %d\n", sweep, h->nn, s->box.vol, s->part.NN);
  gzprintf(f, "###############################\n");
  gzprintf(f, "# rr, rdf\n");
  gzprintf(f, "###############################\n");
  for (int i = 0; i < h->nn; i++) gzprintf(f, "%.8e %.8e\n", h->x[i], h->h[i]);
  gzclose(f);
  if (++s->rdf_samples_in_file == in->rdf_samples) {
    s->rdf_samples_in_file = 0;
    s->rdf_file_id++;
  }
}

/* ---- Steinhardt order parameter q_l on the GPU (hsmc_gpu_order_parameter, K8) ---------------
 * q_l(i) = sqrt(4 pi/(2l+1) sum_m |<Y_lm>_bonds|^2), bonds = neighbours within ql_rmax
 * (compute_order_parameter.c:84-229); the host only keeps the reference's cutoff rule, its
 * WARNING text and the output file. */
void hs_compute_op(hs_sim *s, bool init) {
  hs_input *in = &s->in;
  const double L[3] = {s->box.lx, s->box.ly, s->box.lz};
  if (init) {
    /* compute_order_parameter.c:29-40: the reference limits the cutoff to the edge of ITS
       neighbour-list cells, L/floor(L/neigh_dr) (it keeps the largest of the three edges),
       and says so.  The drop-in reproduces value and message so q_l stays comparable. */
    double size[3];
    int num[3], tot = 1;
    for (int a = 0; a < 3; a++) { num[a] = (int)floor(L[a] / in->neigh_dr); tot *= num[a]; }
    if (tot < 27) num[0] = num[1] = num[2] = 3;          /* cell_list.c:108-113 */
    for (int a = 0; a < 3; a++) size[a] = L[a] / num[a];
    double nl_size = size[0];
    if (nl_size < size[1]) nl_size = size[1];
    if (nl_size < size[2]) nl_size = size[2];
    if (nl_size < in->ql_rmax) {
      printf("WARNING: The cutoff for the order parameter was reduced to %f in order to be consistent with the neighbor list size\n", nl_size);
      in->ql_rmax = nl_size;
    }
  }
  /* the GPU cell grid (even cell counts, edge >= neigh_dr) bounds the bond cutoff the same way */
  hsmc_gpu_info gi;
  hs_gpu_check(hsmc_gpu_get_info(s->gpu, &gi));
  double edge = fmin(gi.cell_size[0], fmin(gi.cell_size[1], gi.cell_size[2]));
  if (edge < in->ql_rmax) {
    printf("WARNING: The cutoff for the order parameter was reduced to %f in order to be consistent with the neighbor list size\n", edge);
    in->ql_rmax = edge;
  }
  double ql_ave = 0.0;
  hs_gpu_check(hsmc_gpu_order_parameter(s->gpu, in->ql_order, in->ql_rmax, &ql_ave));
  if (!HS_ROOT(s)) return;
  FILE *f = open_sample_file("order_param.dat", init, "order parameter");
  if (init) {
    fprintf(f, "###############################################################\n");
    fprintf(f, "# Average order parameter of order %d (each line is one sample)\n", in->ql_order);
    fprintf(f, "###############################################################\n");
  }
  fprintf(f, "%.8e\n", ql_ave);
  fclose(f);
}

void hs_observables_free(hs_sim *s) {
  hist_free(&s->pressv);
  hist_free(&s->presst);
  hist_free(&s->rdf);
}
