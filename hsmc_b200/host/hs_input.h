/* hs_input.h -- input-file keywords of the drop-in host driver.
 *
 * Same keywords, value order and defaults as the reference's parser
 * (read_input.c:71-115 defaults, :138-441 keywords) so that existing input files run
 * unchanged; the implementation is a keyword table, not the reference's if-chain.
 */
#ifndef HS_INPUT_H
#define HS_INPUT_H

typedef struct hs_input {
  double rho;
  int nx, ny, nz, type;
  double neigh_dr;
  int neigh_max_part;
  double dr_max;
  int sweep_eq, sweep_stat;
  int output_int;
  double press, dv_max;
  int opt_flag, opt_sweeps, opt_samples;
  double opt_part_target, opt_vol_target;
  unsigned long seed;
  double cavity_pcav, cavity_maxdr, cavity_mindr, cavity_out_dr;
  int cavity_sample_int;
  int cluster_flag, cluster_moves_sweep, cluster_init_step;
  int restart_read;
  char restart_name[100];
  int restart_write;
  int config_write, config_samples;
  double pressv_dr;
  int pressv_sample_int;
  double presst_dxi, presst_xi_max;
  int presst_sample_int;
  int ql_order;
  double ql_rmax;
  int ql_sample_int;
  int mu_insertions, mu_sample_int;
  double rdf_dr, rdf_rmax;
  int rdf_sample_int, rdf_samples;
} hs_input;

void hs_input_defaults(hs_input *in);
/* prints the reference's messages; exits on unknown key / missing value like it does */
void hs_input_read(hs_input *in, const char *filename);
void hs_input_print_example(void);

#endif
