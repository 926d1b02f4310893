/* hs_input.h -- input-file keywords of the drop-in host driver.
 *
 * Same keywords, value order and defaults as the reference's parser
 * (read_input.c:71-115 defaults, :138-441 keywords) so that existing input files run
 * unchanged; the implementation is a keyword table, not the reference's if-chain.
 */
#ifndef HS_INPUT_H
#define HS_INPUT_H

typedef struct hs_input {
  /* one group per input keyword, in the order of the keyword table of hs_input.c; the member
     names are the reference's G_IN names so that the two drivers read alike */
  double rho;                                       /* rho <density> */
  int nx;                                           /* cells_x <n> */
  int ny;                                           /* cells_y <n> */
  int nz;                                           /* cells_z <n> */
  int type;                                         /* type <1 sc | 2 fcc> */
  double neigh_dr;                                  /* neigh_list <min cell edge> <max particles per cell> */
  int neigh_max_part;
  double dr_max;                                    /* dr_max <max displacement> */
  int sweep_eq;                                     /* sweep_eq <n> */
  int sweep_stat;                                   /* sweep_stat <n> */
  int output_int;                                   /* out <sweeps between progress lines> */
  double press;                                     /* npt <pressure> <max ln V step> */
  double dv_max;
  int opt_flag;                                     /* opt <on> <sweeps> <samples> <target acc.> <target vol. acc.> */
  int opt_sweeps;
  int opt_samples;
  double opt_part_target;
  double opt_vol_target;
  unsigned long seed;                               /* seed <n> */
  double cavity_pcav;                               /* cavity <...>: parsed, refused by the GPU driver */
  double cavity_maxdr;
  double cavity_mindr;
  int cavity_sample_int;
  double cavity_out_dr;
  int cluster_flag;                                 /* cluster <...>: parsed, refused by the GPU driver */
  int cluster_moves_sweep;
  int cluster_init_step;
  int restart_read;                                 /* restart_read <on> <file> */
  char restart_name[100];
  int restart_write;                                /* restart_write <sweeps between files> */
  int config_write;                                 /* config_write <sweeps between samples> <samples per file> */
  int config_samples;
  double pressv_dr;                                 /* press_virial <bin> <sweeps between samples> */
  int pressv_sample_int;
  double presst_dxi;                                /* press_thermo <d xi> <xi max> <sweeps between samples> */
  double presst_xi_max;
  int presst_sample_int;
  int ql_order;                                     /* ql <l> <bond cutoff> <sweeps between samples> */
  double ql_rmax;
  int ql_sample_int;
  int mu_insertions;                                /* widom <insertions> <sweeps between samples> */
  int mu_sample_int;
  double rdf_dr;                                    /* rdf <bin> <r max> <sweeps between samples> <samples per file> */
  double rdf_rmax;
  int rdf_sample_int;
  int rdf_samples;
} hs_input;

void hs_input_defaults(hs_input *in);
/* prints the reference's messages; exits on unknown key / missing value like it does */
void hs_input_read(hs_input *in, const char *filename);
void hs_input_print_example(void);

#endif
