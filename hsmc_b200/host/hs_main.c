/* hs_main.c -- command line of the drop-in host driver: `hsmc_b200 -i IN_FILE [-o OUT_FILE]`,
 * `-e` prints an example input (exec.c:24-146 of the reference; getopt instead of argp). */
#define _GNU_SOURCE
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hs_sim.h"

static void usage(const char *argv0) {
  printf("Usage: %s [-i IN_FILE] [-o OUT_FILE] [-g GPUS] [-e]\n"
         "hsmc_b200 performs Monte Carlo simulations of mono-disperse hard-sphere systems on a B200 GPU.\n"
         "  -i, --input=IN_FILE    Input read from IN_FILE instead of from in.dat\n"
         "  -o, --output=OUT_FILE  Output to OUT_FILE instead of standard output\n"
         "  -g, --gpus=K           Slab-decompose the box over K GPUs of this node (one process per GPU;\n"
         "                         default 1, or the HSMC_GPUS environment variable)\n"
         "  -e, --example          Print example of input file on screen\n", argv0);
}

int main(int argc, char **argv) {
  const char *input = "in.dat", *output = NULL;
  int example = 0;
  int gpus = getenv("HSMC_GPUS") ? atoi(getenv("HSMC_GPUS")) : 1;
  static const struct option longopts[] = {
    {"input", required_argument, 0, 'i'}, {"output", required_argument, 0, 'o'},
    {"gpus", required_argument, 0, 'g'}, {"example", no_argument, 0, 'e'}, {"help", no_argument, 0, '?'}, {0, 0, 0, 0}};
  int c;
  while ((c = getopt_long(argc, argv, "i:o:g:e?", longopts, NULL)) != -1) {
    if (c == 'i') input = optarg;
    else if (c == 'o') output = optarg;
    else if (c == 'g') gpus = atoi(optarg);
    else if (c == 'e') example = 1;
    else { usage(argv[0]); return c == '?' ? 0 : 1; }
  }
  if (example) {
    hs_input_print_example();
    return 0;
  }
  if (output && !freopen(output, "w", stdout)) {
    printf("Failed to pipe output to file %s\n", output);
    exit(EXIT_FAILURE);
  }
  hs_sim *s = calloc(1, sizeof(*s));
  hs_input_read(&s->in, input);
  /* one process per GPU, forked before anything touches CUDA; rank 0 (this process) reports */
  if (gpus > 1 && (s->in.cavity_pcav > 0 || s->in.cluster_flag > 0)) gpus = 1;
  hs_mp_start(&s->mp, gpus, hs_plan_particles(&s->in));
  hs_die_hook(&s->mp);
  if (s->in.press > 0) hs_run_npt_simulation(s);
  else if (s->in.cavity_pcav > 0) hs_die("cavity simulations are not part of the B200 hot path; use the reference CPU build");
  else if (s->in.cluster_flag > 0) hs_die("cluster moves are not part of the B200 hot path; use the reference CPU build");
  else hs_run_nvt_simulation(s);
  if (hs_mp_finish(&s->mp)) hs_die("a GPU rank of this run failed");
  printf("Simulation complete!\n");
  free(s);
  return 0;
}
