/* hs_sim.h -- simulation context of the drop-in host driver.
 *
 * The reference keeps this state in file-scope globals (G_IN, sim_box_info, part_info,
 * part_conf, rng_mt, move counters: SURVEY.md 0.2); here it is one struct that also owns
 * the GPU handle.  The host mirror `conf` keeps the reference layout {id,x,y,z}
 * (sim_info.h:20) and is refreshed from the device before anything reads it.
 */
#ifndef HS_SIM_H
#define HS_SIM_H

#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hsmc_gpu.h"
#include "hs_input.h"
#include "hs_rng.h"
#include "hs_mp.h"

/* on-disk mirror of the reference's box_info / p_info (sim_info.h:6-18): restart files
   are raw struct dumps (io_config.c:53-69) */
typedef struct hs_box {
  double vol;
  double lx, ly, lz;
  double min_size;
  double cell_size;
  int cell_x, cell_y, cell_z;
  int cell_type;
} hs_box;

typedef struct hs_pinfo {
  int NN;
  int Ncell;
} hs_pinfo;

typedef struct hs_hist {
  int nn;
  double *x, *h;
  bool live;
} hs_hist;

typedef struct hs_sim {
  hs_input in;
  hs_box box;
  hs_pinfo part;
  double (*conf)[4];
  bool mirror_current;     /* host mirror == device state */
  bool conf_pinned;        /* the table is page-locked (hs_gpu_open) */
  bool fetch_failed;       /* a slice of a streamed snapshot did not arrive (hs_write_config) */
  hs_rng rng;
  hsmc_gpu *gpu;
  hs_mp mp;                /* process-per-GPU plumbing (`-g K`); world 1 = the plain serial driver */
  hsmc_gpu *gpu_rep;       /* world > 1: single-GPU handle holding a replica, for the sharded RDF */
  uint64_t philox_sweeps;  /* restart: device RNG sweep counter */
  int64_t carry_moves[6];  /* move counters of handles retired by a slab redistribution */
  /* observables */
  hs_hist pressv, presst, rdf;
  double pressv_rmax;
  uint64_t mu_samples;
  int rdf_samples_in_file, rdf_file_id;
  int config_samples_in_file, config_file_id;
  bool restart_checked, config_checked;
} hs_sim;

/* fatal error in the reference's style: "ERROR: ..." + exit(EXIT_FAILURE) */
void hs_die(const char *fmt, ...);
void hs_die_hook(hs_mp *mp);                 /* so that a fatal error on one rank stops the others */
#define HS_ROOT(s) ((s)->mp.rank == 0)       /* rank 0 alone writes output files */
/* particles the run will hold (lattice counts, or the header of the restart file) */
int64_t hs_plan_particles(const hs_input *in);
void hs_gpu_check(int rc);

/* box + lattice (sim_info.c:32-71, 99-166) */
void hs_box_init(hs_sim *s, int type, int nx, int ny, int nz, double rho);
void hs_part_alloc(hs_sim *s);
void hs_part_init(hs_sim *s);
void hs_print_sim_info(const hs_sim *s);

/* device glue */
void hs_gpu_open(hs_sim *s);                 /* where cell_list_init(true) was */
void hs_gpu_close(hs_sim *s);                /* where cell_list_free was */
void hs_gpu_push(hs_sim *s);                 /* host mirror -> device (+ cell list) */
void hs_gpu_pull(hs_sim *s);                 /* device -> host mirror, if stale */

/* restart / configuration files (io_config.c) */
void hs_write_restart(hs_sim *s, int sweep);
void hs_read_restart(hs_sim *s, const char *name);
void hs_write_config(hs_sim *s, int sweep);

/* moves */
void hs_sweep_nvt(hs_sim *s);
void hs_sweep_npt(hs_sim *s);
void hs_vol_move(hs_sim *s);
void hs_counters(hs_sim *s, int64_t out[6]);
void hs_reset_counters(hs_sim *s);

/* observables (compute_press.c, compute_widom_chem_pot.c, compute_rdf.c,
   compute_order_parameter.c) */
void hs_compute_pressv(hs_sim *s, bool init);
void hs_compute_presst(hs_sim *s, bool init);
void hs_compute_mu(hs_sim *s, bool init);
void hs_compute_rdf(hs_sim *s, bool init, int sweep);
void hs_compute_op(hs_sim *s, bool init);
void hs_observables_free(hs_sim *s);

/* optimizer.c */
void hs_opt_nvt(hs_sim *s);
void hs_opt_npt(hs_sim *s);

/* nvt.c / npt.c */
void hs_run_nvt_simulation(hs_sim *s);
void hs_run_npt_simulation(hs_sim *s);

#endif
