/* hs_input.c -- keyword-table parser for the reference's input format
 * (one keyword per line, '#' comments, blank lines skipped, all values of a keyword
 * mandatory: README.md:19-27, read_input.c:62-500). */
#define _GNU_SOURCE
#include "hs_input.h"

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { T_INT, T_DBL, T_ULONG, T_STR100 };

typedef struct { int type; size_t off; } field;
typedef struct { const char *key; int n; field f[5]; } keyword;

#define F(t, m) { t, offsetof(hs_input, m) }

static const keyword TABLE[] = {
  {"rho", 1, {F(T_DBL, rho)}},
  {"cells_x", 1, {F(T_INT, nx)}},
  {"cells_y", 1, {F(T_INT, ny)}},
  {"cells_z", 1, {F(T_INT, nz)}},
  {"type", 1, {F(T_INT, type)}},
  {"neigh_list", 2, {F(T_DBL, neigh_dr), F(T_INT, neigh_max_part)}},
  {"dr_max", 1, {F(T_DBL, dr_max)}},
  {"sweep_eq", 1, {F(T_INT, sweep_eq)}},
  {"sweep_stat", 1, {F(T_INT, sweep_stat)}},
  {"out", 1, {F(T_INT, output_int)}},
  {"npt", 2, {F(T_DBL, press), F(T_DBL, dv_max)}},
  {"opt", 5, {F(T_INT, opt_flag), F(T_INT, opt_sweeps), F(T_INT, opt_samples), F(T_DBL, opt_part_target),
              F(T_DBL, opt_vol_target)}},
  {"seed", 1, {F(T_ULONG, seed)}},
  {"cavity", 5, {F(T_DBL, cavity_pcav), F(T_DBL, cavity_maxdr), F(T_DBL, cavity_mindr),
                 F(T_INT, cavity_sample_int), F(T_DBL, cavity_out_dr)}},
  {"cluster", 3, {F(T_INT, cluster_flag), F(T_INT, cluster_moves_sweep), F(T_INT, cluster_init_step)}},
  {"restart_read", 2, {F(T_INT, restart_read), F(T_STR100, restart_name)}},
  {"restart_write", 1, {F(T_INT, restart_write)}},
  {"config_write", 2, {F(T_INT, config_write), F(T_INT, config_samples)}},
  {"press_virial", 2, {F(T_DBL, pressv_dr), F(T_INT, pressv_sample_int)}},
  {"press_thermo", 3, {F(T_DBL, presst_dxi), F(T_DBL, presst_xi_max), F(T_INT, presst_sample_int)}},
  {"ql", 3, {F(T_INT, ql_order), F(T_DBL, ql_rmax), F(T_INT, ql_sample_int)}},
  {"widom", 2, {F(T_INT, mu_insertions), F(T_INT, mu_sample_int)}},
  {"rdf", 4, {F(T_DBL, rdf_dr), F(T_DBL, rdf_rmax), F(T_INT, rdf_sample_int), F(T_INT, rdf_samples)}},
};
#define NKEY (sizeof(TABLE) / sizeof(TABLE[0]))

void hs_input_defaults(hs_input *in) {
  memset(in, 0, sizeof(*in));
  in->rho = 0.5;
  in->nx = in->ny = in->nz = 3;
  in->type = 1;
  in->neigh_dr = 1.0;
  in->neigh_max_part = 10;
  in->dr_max = 0.05;
  in->cavity_maxdr = 1.2;
  in->cavity_out_dr = 0.01;
  in->cavity_sample_int = 100;
  in->cluster_moves_sweep = 1;
  in->cluster_init_step = 10000;
  in->config_samples = 128;
  in->pressv_dr = 0.01;
  in->presst_dxi = 0.0001;
  in->presst_xi_max = 0.002;
  in->dv_max = 0.001;
  in->opt_flag = 1;
  in->opt_sweeps = 1000;
  in->opt_samples = 10;
  in->opt_part_target = 0.5;
  in->opt_vol_target = 0.5;
  in->ql_order = 6;
  in->ql_rmax = 1.5;
  in->mu_insertions = 100;
  in->rdf_dr = 0.01;
  in->rdf_rmax = 10;
  in->rdf_samples = 100;
}

/* read_input.c:477-500.  The reference prints its line buffer AFTER strtok(line, " ") has cut it, i.e.
   the first space-delimited token only (with its newline when the line has no space); same here. */
static void input_error(int kind, char *line) {
  if (kind == 1) printf("Missing value to key\n");
  else if (kind == 2) printf("Unknown key\n");
  else printf("Name of restart file is too long, maximum 100 characters\n");
  while (*line == ' ') line++;
  char *sp = strchr(line, ' ');
  if (sp) *sp = '\0';
  printf("Last read line in the input file:\n%s\n", line);
  exit(EXIT_FAILURE);
}

void hs_input_read(hs_input *in, const char *filename) {
  hs_input_defaults(in);
  printf("Reading input data from %s ...\n", filename);
  FILE *fp = fopen(filename, "r");
  if (!fp) {
    printf("Error! Could not open file %s\n", filename);
    exit(EXIT_FAILURE);
  }
  char *line = NULL;
  size_t cap = 0;
  while (getline(&line, &cap, fp) >= 0) {
    if (line[0] == '#' || line[0] == '\n') continue;
    char *copy = strdup(line);
    char *save = NULL;
    char *key = strtok_r(line, " \t\r\n", &save);
    if (!key) { free(copy); continue; }
    const keyword *kw = NULL;
    for (size_t k = 0; k < NKEY; k++)
      if (strcmp(key, TABLE[k].key) == 0) { kw = &TABLE[k]; break; }
    if (!kw) input_error(2, copy);
    for (int v = 0; v < kw->n; v++) {
      char *val = strtok_r(NULL, " \t\r\n", &save);
      if (!val) input_error(1, copy);
      void *dst = (char *)in + kw->f[v].off;
      switch (kw->f[v].type) {
        case T_INT: *(int *)dst = atoi(val); break;
        case T_DBL: *(double *)dst = atof(val); break;
        case T_ULONG: *(unsigned long *)dst = strtoul(val, NULL, 10); break;
        case T_STR100:
          if (strlen(val) >= 100) input_error(3, copy);
          strcpy((char *)dst, val);
          break;
      }
    }
    free(copy);
  }
  free(line);
  fclose(fp);
  printf("Done\n");
  fflush(stdout);
}

void hs_input_print_example(void) {
  fputs(
    "# NVT run at density 0.5/sigma^3 with 1000 particles on a simple-cubic start,\n"
    "# pressure from the contact value of g(r) (virial route).\n"
    "# 1e6 sweeps of equilibration, 1e6 sweeps of statistics.\n\n"
    "rho 0.5\n\n"
    "cells_x 10\ncells_y 10\ncells_z 10\ntype 1\n\n"
    "# cell edge >= 1.05 is required by press_virial\n"
    "neigh_list 1.05 10\n\n"
    "dr_max 0.05\n\n"
    "opt 1 1000 10 0.5 0.5\n\n"
    "press_virial 0.002 10\n\n"
    "seed 124787\n\n"
    "restart_write 100000\n\n"
    "config_write 100000 100\n\n"
    "sweep_eq 1000000\n\n"
    "sweep_stat 1000000\n\n"
    "out 10000\n\n", stdout);
}
