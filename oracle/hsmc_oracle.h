/* oracle/hsmc_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of the reference's hard-sphere hot path with explicit state
 * instead of file-scope globals.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may link or call this.  Parity status: PINNED --
 * every routine is checked bit-for-bit against the unmodified reference
 * (oracle/_ref/libhsmc_ref.so) in tests/test_oracle_vs_ref.py and against the
 * committed fixtures under tests/golden/ (generated from the reference by
 * tests/golden/make_golden.py).  The reference itself ships no tests or vectors
 * (SURVEY.md section 4).
 */
#ifndef HSMC_ORACLE_H
#define HSMC_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_sys orc_sys;

/* box + lattice (sim_info.c:32-71, 125-166) */
void orc_box_from_lattice(int type, int nx, int ny, int nz, double rho, double *box4 /*lx,ly,lz,vol*/);
int orc_lattice_count(int type, int nx, int ny, int nz);
void orc_lattice_fill(int type, int nx, int ny, int nz, double rho, double *conf4);

/* system + cell list (cell_list.c:33-58, 92-131, 142-175, 253-301) */
orc_sys *orc_create(int N, double lx, double ly, double lz, const double *conf4,
                    double neigh_dr, int max_part);
void orc_destroy(orc_sys *s);
int orc_N(const orc_sys *s);
void orc_cells(const orc_sys *s, int *num3, double *size3);
void orc_get_conf(const orc_sys *s, double *conf4);
void orc_set_conf(orc_sys *s, const double *conf4);
int orc_cell_of(const orc_sys *s, int idx);
int orc_status(const orc_sys *s); /* nonzero after a condition on which the reference exits */

/* moves.c */
double orc_compute_dist(const orc_sys *s, int i, int j, double sf);              /* :400-431 */
int orc_check_overlap(const orc_sys *s, int idx, double sf);                     /* :157-212 */
int orc_trial_verdict(orc_sys *s, int idx, double x, double y, double z, double sf);
void orc_trial_verdicts(orc_sys *s, int n, const int *idx, const double *xyz, double sf, int *flags);
void orc_overlap_all(const orc_sys *s, double sf, int *flags);
int orc_any_overlap(const orc_sys *s, double sf);                                /* :106-112 */
/* part_move (:27-80) driven by explicit raw 32-bit draws: returns 1 accepted, 0 rejected */
int orc_part_move_raw(orc_sys *s, int idx, uint32_t rx, uint32_t ry, uint32_t rz, double dr_max);
void orc_replay_moves(orc_sys *s, int n, const int *ids, const uint32_t *raw3, double dr_max, int *accepted);
void orc_counters(const orc_sys *s, int64_t *out6);
void orc_reset_counters(orc_sys *s);
/* rescale after an accepted volume move (:129-142); new box passed in */
void orc_rescale(orc_sys *s, double sf, double lx, double ly, double lz);

/* serial sweeps with the reference's own RNG call sequence (nvt.c:201-209) */
typedef struct orc_mt orc_mt;
orc_mt *orc_mt_create(unsigned long seed);
void orc_mt_destroy(orc_mt *m);
uint32_t orc_mt_raw(orc_mt *m);
double orc_mt_double(orc_mt *m);               /* rng.c:29-31 */
int orc_mt_int(orc_mt *m, int n);              /* rng.c:34-36 + gsl_rng_uniform_int */
void orc_sweep_nvt(orc_sys *s, orc_mt *m, int n_sweeps, double dr_max);

/* observables */
void orc_widom_verdicts(const orc_sys *s, int M, const double *xyz, int *flags); /* widom.c:82-160 */
int64_t orc_widom_count_raw(const orc_sys *s, int64_t M, const uint32_t *raw3);  /* :44-80 */
int orc_rdf_nn(double dr, double rmax);                                          /* rdf.c:76 */
void orc_rdf_counts(const orc_sys *s, double dr, double rmax_in, int nn, uint64_t *counts); /* :110-128 */
int orc_pressv_nn(double dr);                                                    /* press.c:88 */
void orc_pressv_counts(const orc_sys *s, double dr, int nn, uint64_t *counts);   /* :123-165 */
int orc_presst_nn(double dxi, double xi_max);                                    /* :211 */
void orc_presst_flags(const orc_sys *s, double dxi, int nn, int *flags, double *sf_out); /* :250-271 */

/* compute_order_parameter.c:84-229: average Steinhardt q_l, bonds = stencil neighbours within rmax */
double orc_order_param(const orc_sys *s, int l, double rmax);

/* Philox4x32-10 (Salmon et al., SC'11) -- the device RNG, restated for the checker */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
