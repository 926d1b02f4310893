/* hs_run.c -- ensemble drivers and the step-size optimizer of the drop-in host driver.
 *
 * Same control flow, sampling schedule and stdout lines as the reference (nvt.c:30-199,
 * npt.c:25-175, optimizer.c:23-166); every sweep / pair loop goes to the device.
 * Timing is wall-clock (the reference's clock() would only see host CPU time).
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <time.h>

#include "hs_sim.h"
#include "hs_opt.h"

static double now_s(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static bool due(int interval, int sweep) { return interval > 0 && sweep % interval == 0; }

static void setup_state(hs_sim *s) {
  hs_input *in = &s->in;
  if (in->restart_read == 0) {
    hs_box_init(s, in->type, in->nx, in->ny, in->nz, in->rho);
    hs_part_alloc(s);
    hs_part_init(s);
    hs_rng_seed(&s->rng, in->seed);
  } else {
    hs_read_restart(s, in->restart_name);
  }
  hs_print_sim_info(s);
  hs_gpu_open(s);
}

/* ---- optimizer: secant search on the acceptance ratio (optimizer.c:23-166) ----------- */
typedef struct { double dr, dv, acc_part, acc_vol; } opt_sample;

static opt_sample take_sample(hs_sim *s, int sweeps, bool npt) {
  hs_reset_counters(s);
  for (int i = 0; i < sweeps; i++) {
    if (npt) hs_sweep_npt(s); else hs_sweep_nvt(s);
  }
  int64_t c[6];
  hs_counters(s, c);
  opt_sample o = {s->in.dr_max, s->in.dv_max, (double)c[1] / (double)c[0], 0.0};
  if (npt) o.acc_vol = (double)c[4] / (double)c[3];
  return o;
}

static void optimize(hs_sim *s, bool npt) {
  hs_input *in = &s->in;
  int per_sample = in->opt_sweeps / in->opt_samples;
  printf("---------------------------------------------------\n");
  printf("Maximum displacement optimization started ...\n");
  printf("Sweeps for optimization: %d\n", in->opt_sweeps);
  printf("Number of samples: %d\n", in->opt_samples);
  opt_sample a = take_sample(s, per_sample, npt);
  in->dr_max = (a.acc_part > in->opt_part_target) ? in->dr_max * 2 : in->dr_max / 2;
  if (npt) in->dv_max = (a.acc_vol > in->opt_vol_target) ? in->dv_max * 2 : in->dv_max / 2;
  opt_sample b = take_sample(s, per_sample, npt);
  for (int i = 0; i < in->opt_samples; i++) {
    /* the reference's secant step with its clamps; a step that is not a positive finite number (two samples
       with equal acceptance: the reference divides by zero, optimizer.c:45) keeps the previous one (hs_opt.h) */
    in->dr_max = hs_opt_next_dr(a.dr, a.acc_part, b.dr, b.acc_part, in->opt_part_target);
    if (npt) in->dv_max = hs_opt_next_dv(a.dv, a.acc_vol, b.dv, b.acc_vol, in->opt_vol_target);
    a = b;
    b = take_sample(s, per_sample, npt);
  }
  if (!npt) {
    printf("Optimal maximum displacement: %.8f\n", in->dr_max);
    printf("Acceptance ratio: %.8f\n", b.acc_part);
  } else {
    printf("Optimal maximum particle displacement: %.8f\n", in->dr_max);
    printf("Acceptance ratio: %.8f \n", b.acc_part);
    printf("Optimal maximum volume deformation: %.8f\n", in->dv_max);
    printf("Acceptance ratio: %.8f \n", b.acc_vol);
  }
  printf("Maximum displacement optimization completed\n");
}

void hs_opt_nvt(hs_sim *s) { optimize(s, false); }
void hs_opt_npt(hs_sim *s) { optimize(s, true); }

/* ---- sweep loops (nvt.c:105-199, npt.c:107-175) ------------------------------------------ */
typedef struct { bool pressv, presst, ql, mu, rdf; } first_flags;

static void run_phase(hs_sim *s, bool npt, bool production, int offset) {
  hs_input *in = &s->in;
  int n = production ? in->sweep_stat : in->sweep_eq;
  first_flags first = {true, true, true, true, true};
  for (int ii = offset; ii < n + offset; ii++) {
    if (npt) {
      if (ii == 0) printf("Sweep number  Density\n");
      if (ii % in->output_int == 0) { printf("%d  %.8f\n", ii, in->rho); fflush(stdout); }
    } else {
      if (ii == offset) printf("Sweep number\n");
      if (ii % in->output_int == 0) { printf("%d\n", ii); fflush(stdout); }
    }
    if (due(in->restart_write, ii)) hs_write_restart(s, ii);
    if (production) {
      if (due(in->config_write, ii)) hs_write_config(s, ii);
      if (!npt && due(in->pressv_sample_int, ii)) { hs_compute_pressv(s, first.pressv); first.pressv = false; }
      if (due(in->presst_sample_int, ii)) { hs_compute_presst(s, first.presst); first.presst = false; }
      if (due(in->ql_sample_int, ii)) { hs_compute_op(s, first.ql); first.ql = false; }
      if (!npt && due(in->mu_sample_int, ii)) { hs_compute_mu(s, first.mu); first.mu = false; }
      if (!npt && due(in->rdf_sample_int, ii)) { hs_compute_rdf(s, first.rdf, ii); first.rdf = false; }
    }
    if (npt) hs_sweep_npt(s); else hs_sweep_nvt(s);
  }
  hs_observables_free(s);
}

static void simulate(hs_sim *s, bool npt) {
  hs_input *in = &s->in;
  if (in->output_int <= 0) hs_die("the output interval `out` must be a positive number of sweeps");
  setup_state(s);
  if (in->opt_flag == 1) {
    double rho_start = in->rho;
    optimize(s, npt);
    /* the reference restarts from the lattice after tuning (nvt.c:58-62, npt.c:54-61) */
    in->rho = rho_start;
    hs_box_init(s, in->type, in->nx, in->ny, in->nz, rho_start);
    if (npt) {
      double nb[3] = {s->box.lx, s->box.ly, s->box.lz};
      hs_gpu_close(s);
      hs_part_init(s);
      hs_gpu_open(s);
      (void)nb;
    } else {
      hs_part_init(s);
      hs_gpu_push(s);
    }
  }
  double t0 = now_s();
  hs_reset_counters(s);
  printf("---------------------------------------------------\n");
  printf("Equilibration...\n");
  run_phase(s, npt, false, 0);
  printf("Equilibration completed.\n");
  printf("---------------------------------------------------\n");
  printf("Production...\n");
  run_phase(s, npt, true, in->sweep_eq);
  printf("Production completed.\n");
  hs_gpu_check(hsmc_gpu_sync(s->gpu));
  double t1 = now_s();
  int64_t c[6];
  hs_counters(s, c);
  printf("---------------------------------------------------\n");
  printf("-- Particle moves: %.8e\n", (double)c[0]);
  printf("   Acceptance percentage: %f\n", (double)c[1] / (double)c[0]);
  printf("   Rejection percentage: %f\n", (double)c[2] / (double)c[0]);
  if (npt) {
    printf("-- Volume moves: %.8e\n", (double)c[3]);
    printf("   Acceptance percentage: %f\n", (double)c[4] / (double)c[3]);
    printf("   Rejection percentage: %f\n", (double)c[5] / (double)c[3]);
  }
  printf("Elapsed time: %f seconds\n", t1 - t0);
  if (getenv("HSMC_REPORT_LAUNCHES")) {       /* bench.py: kernels this handle launched (stderr: stdout is the reference's format) */
    hsmc_gpu_info gi;
    hs_gpu_check(hsmc_gpu_get_info(s->gpu, &gi));
    fprintf(stderr, "[hsmc_b200] kernel launches: %llu\n", (unsigned long long)gi.kernel_launches);
  }
  hs_gpu_pull(s);
  hs_gpu_close(s);
  if (s->mp.world == 1) free(s->conf);
  s->conf = NULL;
}

void hs_run_nvt_simulation(hs_sim *s) { simulate(s, false); }
void hs_run_npt_simulation(hs_sim *s) { simulate(s, true); }
