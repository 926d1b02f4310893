/* hs_sim.c -- box, lattice, device glue, restart/config files, trial moves. */
#define _GNU_SOURCE
#include "hs_sim.h"
#include "hs_fastio.h"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static hs_mp *g_mp = NULL;
void hs_die_hook(hs_mp *mp) { g_mp = mp; }

void hs_die(const char *fmt, ...) {
  va_list ap;
  if (g_mp && g_mp->rank != 0) {
    /* the stdout of ranks > 0 is /dev/null: say it on stderr before taking the run down */
    va_start(ap, fmt);
    fprintf(stderr, "ERROR (GPU rank %d): ", g_mp->rank);
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
  }
  if (g_mp) hs_mp_abort(g_mp);
  va_start(ap, fmt);
  printf("ERROR: ");
  vprintf(fmt, ap);
  printf("\n");
  va_end(ap);
  fflush(stdout);
  exit(EXIT_FAILURE);
}

void hs_gpu_check(int rc) {
  if (rc) hs_die("%s", hsmc_gpu_last_error());
}

/* ---- box + lattice ------------------------------------------------------------- */
static int particles_per_cell(int type) {
  if (type == 1) return 1;          /* simple cubic */
  if (type == 2) return 4;          /* face-centred cubic */
  printf("Unknown lattice type, default to fcc\n");
  return 4;
}

/* sim_info.c:32-71: the box follows from the lattice counts and the density */
void hs_box_init(hs_sim *s, int type, int nx, int ny, int nz, double rho) {
  int ppc = particles_per_cell(type);
  double cell_vol = ppc / rho;
  double a = pow(cell_vol, 1. / 3.);
  int nmin = nx < ny ? nx : ny;
  if (nz < nmin) nmin = nz;
  hs_box *b = &s->box;
  b->cell_x = nx; b->cell_y = ny; b->cell_z = nz;
  b->lx = nx * a; b->ly = ny * a; b->lz = nz * a;
  b->min_size = nmin * a;
  b->vol = nx * ny * nz * cell_vol;
  b->cell_size = a;
  b->cell_type = type;
}

int64_t hs_plan_particles(const hs_input *in) {
  if (in->restart_read == 0) return (int64_t)in->nx * in->ny * in->nz * (in->type == 1 ? 1 : 4);
  /* restart_%d.bin: dr_max, dv_max, box_info, p_info, ... (io_config.c:53-61) */
  FILE *f = fopen(in->restart_name, "rb");
  if (!f) { perror("Error while reading restart file"); exit(EXIT_FAILURE); }
  double d[2];
  hs_box b;
  hs_pinfo pi;
  if (fread(d, sizeof(double), 2, f) != 2 || fread(&b, sizeof(b), 1, f) != 1 || fread(&pi, sizeof(pi), 1, f) != 1)
    hs_die("restart file %s is truncated", in->restart_name);
  fclose(f);
  return pi.NN;
}

void hs_part_alloc(hs_sim *s) {
  int ppc = particles_per_cell(s->box.cell_type);
  int n = s->box.cell_x * s->box.cell_y * s->box.cell_z * ppc;
  if (s->mp.world > 1) {
    /* the mirror is the table all ranks share (mapped before the fork) */
    if (n != s->mp.table_rows) hs_die("internal: shared particle table has %ld rows, the run needs %d", (long)s->mp.table_rows, n);
    s->conf = s->mp.table;
  } else {
    s->conf = malloc((size_t)n * sizeof(*s->conf));
  }
  if (!s->conf) hs_die("Failed particle allocation");
  s->part.Ncell = ppc;
  s->part.NN = n;
}

static void lattice_overlap_exit(void) {
  printf("Overlap in the initial configuration. Possible solutions:\n");
  printf("-- If SC lattice was selected, try to change to FCC\n");
  printf("-- If FCC lattice was selected, the selected value of density is unphysical\n");
  exit(EXIT_FAILURE);
}

/* sim_info.c:125-166: lattice sites in (ii,jj,kk) order, fcc basis (0,0,0) (.5,.5,0)
   (.5,0,.5) (0,.5,.5) */
void hs_part_init(hs_sim *s) {
  static const double basis[4][3] = {{0, 0, 0}, {0.5, 0.5, 0}, {0.5, 0, 0.5}, {0, 0.5, 0.5}};
  const hs_box *b = &s->box;
  double a = b->cell_size;
  int nb = (b->cell_type == 1) ? 1 : 4;
  if (nb == 1 ? (a < 1) : (a / sqrt(2.0) < 1)) lattice_overlap_exit();
  int id = 0;
  for (int i = 0; i < b->cell_x; i++)
    for (int j = 0; j < b->cell_y; j++)
      for (int k = 0; k < b->cell_z; k++)
        for (int q = 0; q < nb; q++, id++) {
          s->conf[id][0] = id;
          s->conf[id][1] = (i + basis[q][0]) * a;
          s->conf[id][2] = (j + basis[q][1]) * a;
          s->conf[id][3] = (k + basis[q][2]) * a;
        }
  s->mirror_current = false;
}

void hs_print_sim_info(const hs_sim *s) {
  printf("Simulation box size (x, y, z): %.5f %.5f %.5f\n", s->box.lx, s->box.ly, s->box.lz);
  printf("Number of particles: %d\n", s->part.NN);
  if (s->in.press > 0) printf("Pressure: %.8f\n", s->in.press);
}

/* ---- device glue ----------------------------------------------------------------- */
void hs_gpu_open(hs_sim *s) {
  hsmc_gpu_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  const char *dev = getenv("HSMC_DEVICE");
  cfg.device = (dev ? atoi(dev) : 0) + s->mp.rank;
  cfg.rank = s->mp.rank;
  cfg.world = s->mp.world;
  cfg.seed = (uint64_t)s->in.seed;
  cfg.cell_min = s->in.neigh_dr;
  cfg.regrid_interval = 1;
  unsigned char id[HSMC_GPU_NCCL_ID_BYTES];
  if (s->mp.world > 1) {
    /* rank 0 draws the communicator id, everybody gets it through the shared block */
    if (s->mp.rank == 0) hs_gpu_check(hsmc_gpu_nccl_id(id));
    hs_mp_bcast_id(&s->mp, id, HSMC_GPU_NCCL_ID_BYTES);
    cfg.nccl_id = id;
  } else {
    /* identity checks: a single-GPU run with the block partition of a K-slab run is the K-GPU chain */
    const char *xp = getenv("HSMC_XPART_WORLD");
    if (xp && atoi(xp) > 1) cfg.sweep_impl = atoi(xp) << 8;
  }
  double box[3] = {s->box.lx, s->box.ly, s->box.lz};
  hs_gpu_check(hsmc_gpu_create(&s->gpu, &cfg, s->part.NN, box));
  if (s->mp.world > 1) {
    const char *p2p = getenv("HSMC_P2P");
    if (!p2p || atoi(p2p) != 0) {
      /* NVLink peer-to-peer halo windows: gather the blobs, attach the two neighbours' */
      unsigned char blob[HSMC_GPU_IPC_BYTES];
      const void *left, *right;
      hs_gpu_check(hsmc_gpu_ipc_export(s->gpu, blob));
      hs_mp_exchange_blobs(&s->mp, blob, &left, &right);
      hs_gpu_check(hsmc_gpu_ipc_attach(s->gpu, left, right));
      hs_mp_barrier(&s->mp);
    }
  }
  hs_gpu_check(hsmc_gpu_set_sweep_counter(s->gpu, s->philox_sweeps));
  /* a large table is page-locked so that uploads and (sliced) snapshot downloads run at full PCIe speed and
     asynchronously; failing to lock it only makes them slower */
  if (s->mp.world == 1 && s->part.NN >= (1 << 18) && !s->conf_pinned)
    s->conf_pinned = hsmc_gpu_pin_host(s->conf, (size_t)s->part.NN * sizeof(*s->conf), 1) == 0;
  hs_gpu_push(s);
}

void hs_gpu_close(hs_sim *s) {
  if (s->conf_pinned) { hsmc_gpu_pin_host(s->conf, 0, 0); s->conf_pinned = false; }
  if (s->gpu_rep) hsmc_gpu_destroy(s->gpu_rep);
  s->gpu_rep = NULL;
  if (s->gpu) hsmc_gpu_destroy(s->gpu);
  s->gpu = NULL;
}

void hs_gpu_push(hs_sim *s) {
  /* slab mode: every rank hands over the full table and keeps the rows of its slab */
  hs_gpu_check(hsmc_gpu_upload(s->gpu, &s->conf[0][0], s->part.NN));
  hs_mp_barrier(&s->mp);       /* nobody rewrites the shared table while another rank still reads it */
  s->mirror_current = true;
}

void hs_gpu_pull(hs_sim *s) {
  if (s->mirror_current) return;
  if (s->mp.world == 1) {
    hs_gpu_check(hsmc_gpu_download(s->gpu, &s->conf[0][0]));
  } else {
    /* every rank brings back the rows it owns and files them under their ids in the shared table;
       first make sure nobody (rank 0 writing a snapshot, say) is still reading the previous contents */
    hs_mp_barrier(&s->mp);
    hsmc_gpu_info gi;
    hs_gpu_check(hsmc_gpu_sync(s->gpu));
    hs_gpu_check(hsmc_gpu_get_info(s->gpu, &gi));
    int64_t cap = gi.n_owned + 1024, n = 0;
    double (*rows)[4] = malloc((size_t)cap * sizeof(*rows));
    if (!rows) hs_die("Failed allocation of the download buffer");
    hs_gpu_check(hsmc_gpu_download_owned(s->gpu, &rows[0][0], cap, &n));
    for (int64_t i = 0; i < n; i++) memcpy(s->conf[(int64_t)rows[i][0]], rows[i], sizeof(*rows));
    free(rows);
    hs_mp_barrier(&s->mp);
  }
  s->mirror_current = true;
}

/* ---- restart files: io_config.c:28-130, same byte layout + one trailing field ---- */
void hs_write_restart(hs_sim *s, int sweep) {
  const hs_input *in = &s->in;
  if (!s->restart_checked) {
    if ((double)(in->sweep_stat + in->sweep_eq) / in->restart_write > 100000)
      printf("ERROR: Too many (> 100000) restart files will be produced. Consider writing less often\n");
    s->restart_checked = true;
  }
  int width = (int)ceil(log10(in->sweep_stat + in->sweep_eq));
  char name[64];
  snprintf(name, sizeof(name), "restart_%0*d.bin", width, sweep);
  hs_gpu_pull(s);
  if (!HS_ROOT(s)) return;
  FILE *f = fopen(name, "wb");
  if (!f) { perror("Error while creating restart file\n"); exit(EXIT_FAILURE); }
  fwrite(&in->dr_max, sizeof(double), 1, f);
  fwrite(&in->dv_max, sizeof(double), 1, f);
  fwrite(&s->box, sizeof(hs_box), 1, f);
  fwrite(&s->part, sizeof(hs_pinfo), 1, f);
  fwrite(s->conf, sizeof(double), (size_t)s->part.NN * 4, f);
  hs_rng_write(&s->rng, f);
  /* extension: device Philox sweep counter; the reference ignores trailing bytes */
  hsmc_gpu_info gi;
  hs_gpu_check(hsmc_gpu_get_info(s->gpu, &gi));
  uint64_t tail[2] = {0x48534d4342323030ULL /* "HSMCB200" */, gi.sweeps_done};
  fwrite(tail, sizeof(uint64_t), 2, f);
  fclose(f);
}

void hs_read_restart(hs_sim *s, const char *name) {
  printf("Reading data from restart file %s...\n", name);
  FILE *f = fopen(name, "rb");
  if (!f) { perror("Error while reading restart file"); exit(EXIT_FAILURE); }
  size_t ok = 0;
  ok += fread(&s->in.dr_max, sizeof(double), 1, f);
  ok += fread(&s->in.dv_max, sizeof(double), 1, f);
  ok += fread(&s->box, sizeof(hs_box), 1, f);
  ok += fread(&s->part, sizeof(hs_pinfo), 1, f);
  if (ok != 4) hs_die("restart file %s is truncated", name);
  hs_part_alloc(s);
  if (fread(s->conf, sizeof(double), (size_t)s->part.NN * 4, f) != (size_t)s->part.NN * 4)
    hs_die("restart file %s is truncated", name);
  if (hs_rng_read(&s->rng, f)) hs_die("restart file %s is truncated", name);
  uint64_t tail[2];
  s->philox_sweeps = 0;
  if (fread(tail, sizeof(uint64_t), 2, f) == 2 && tail[0] == 0x48534d4342323030ULL) s->philox_sweeps = tail[1];
  fclose(f);
  s->in.rho = s->part.NN / s->box.vol;
  s->mirror_current = false;
  printf("The following data was initialized via the restart file:\n"
         "- Maximum displacement (override with optimization)\n"
         "- Maximum volume displacement (only for NpT, override with optimization)\n"
         "- Dimensions of the simulation box\n"
         "- Cell list information\n"
         "- Number of particles\n"
         "- Particle's positions\n"
         "- Density\n"
         "- Status of the random number generator\n");
}

/* ---- configuration snapshots: io_config.c:134-191 ----------------------------------- */
/* one slice of the id-ordered table from the device into the host mirror (called from the writer's producer thread) */
static int fetch_slice(void *ctx, long long first, long long n) {
  hs_sim *s = ctx;
  if (hsmc_gpu_fetch_rows(s->gpu, first, n, &s->conf[first][0])) { s->fetch_failed = true; return -1; }
  return 0;
}

void hs_write_config(hs_sim *s, int sweep) {
  const hs_input *in = &s->in;
  if (!s->config_checked) {
    if ((double)(in->sweep_stat + in->sweep_eq) / (in->config_write * in->config_samples) > 100000)
      printf("ERROR: Too many (> 100000) configuration files will be produced. Consider increasing number of samples per file\n");
    s->config_checked = true;
  }
  char name[32];
  snprintf(name, sizeof(name), "config_%06d.dat.gz", s->config_file_id);
  const int append = s->config_samples_in_file != 0;
  /* one GPU, mirror stale: the table comes over in slices while the slices that have arrived are being formatted and
     deflated (hs_fastio.c); otherwise (slabs, serial writer, mirror already current) the whole table first */
  const int stream = s->mp.world == 1 && !s->mirror_current && !getenv("HSMC_IO_SERIAL") && !getenv("HSMC_IO_NO_STREAM");
  if (!stream) hs_gpu_pull(s);
  if (!HS_ROOT(s)) {
    /* keep the file counters in step with rank 0 */
  } else if (getenv("HSMC_IO_SERIAL")) {
    /* the reference's writer, statement for statement (kept for timing comparisons) */
    gzFile f = gzopen(name, append ? "a" : "w");
    if (f == Z_NULL) { perror("Error while creating configuration file"); exit(EXIT_FAILURE); }
    gzprintf(f, "# Sweep number\n%d\n", sweep);
    gzprintf(f, "# Number of particles\n%d\n", s->part.NN);
    gzprintf(f, "# Simulation box size\n%.8f\n%.8f\n%.8f\n", s->box.lx, s->box.ly, s->box.lz);
    gzprintf(f, "# Configuration\n");
    for (int i = 0; i < s->part.NN; i++)
      gzprintf(f, "%d %.8f %.8f %.8f\n", (int)s->conf[i][0], s->conf[i][1], s->conf[i][2], s->conf[i][3]);
    gzclose(f);
  } else {
    /* same bytes after decompression, formatted and deflated chunk-parallel (hs_fastio.c) */
    const double b3[3] = {s->box.lx, s->box.ly, s->box.lz};
    int rc;
    if (stream) {
      hs_gpu_check(hsmc_gpu_pack_table(s->gpu));
      rc = hs_fastio_write_config_stream(name, append, sweep, s->part.NN, b3, (const double (*)[4])s->conf, 0, fetch_slice, s,
                                         1 << 20);
      if (s->fetch_failed) hs_gpu_check(1);
      s->mirror_current = rc == 0;
    } else
      rc = hs_fastio_write_config(name, append, sweep, s->part.NN, b3, (const double (*)[4])s->conf, 0);
    if (rc) {
      perror("Error while creating configuration file");
      exit(EXIT_FAILURE);
    }
  }
  if (++s->config_samples_in_file == in->config_samples) {
    s->config_samples_in_file = 0;
    s->config_file_id++;
  }
}

/* ---- moves ----------------------------------------------------------------------- */
/* nvt.c:201-209: N trial displacements; on the device one checkerboard sweep */
void hs_sweep_nvt(hs_sim *s) {
  hs_gpu_check(hsmc_gpu_sweep_nvt(s->gpu, 1, s->in.dr_max));
  s->mirror_current = false;
}

/* npt.c:177-196 draws, for each of N steps, a volume move with probability 1/(N+1).
   The particle moves run as one device sweep; the number of volume moves of the sweep
   is drawn from the same Binomial(N, 1/(N+1)) law by geometric skipping and they are
   carried out after it (any interleaving leaves the NpT distribution invariant). */
void hs_sweep_npt(hs_sim *s) {
  hs_sweep_nvt(s);
  const int N = s->part.NN;
  const double log_q = log1p(-1.0 / (N + 1.0));
  long pos = 0;
  for (;;) {
    double u = hs_rng_double(&s->rng);
    if (u <= 0.0) break;
    pos += (long)floor(log(u) / log_q) + 1;
    if (pos > N) break;
    hs_vol_move(s);
  }
}

/* moves.c:83-153; the O(N) overlap scan and the rescale run on the device */
void hs_vol_move(hs_sim *s) {
  hs_input *in = &s->in;
  const int N = s->part.NN;
  double r_dv = hs_rng_double(&s->rng);
  double log_vol_new = log(s->box.vol) + (r_dv - 0.5) * in->dv_max;
  double vol_new = exp(log_vol_new);
  double vol_ratio = vol_new / s->box.vol;
  double sf = pow(vol_ratio, 1. / 3.);
  int overlap = 0;
  hs_gpu_check(hsmc_gpu_overlap_scaled(s->gpu, sf, &overlap));
  int accepted = 0;
  if (!overlap) {
    double boltz = exp(in->press * (s->box.vol - vol_new) + (N + 1) * log(vol_ratio));
    double r_acc = hs_rng_double(&s->rng);
    if (r_acc < boltz) {
      accepted = 1;
      in->rho = N / vol_new;
      hs_box_init(s, in->type, in->nx, in->ny, in->nz, in->rho);
      double nb[3] = {s->box.lx, s->box.ly, s->box.lz};
      int rc = hsmc_gpu_rescale(s->gpu, sf, nb);
      if (rc && s->mp.world > 1) {
        /* the move changed the number of cells per axis: slab ownership has to be redrawn.  Same
           verdict on every rank (it only depends on the boxes), and the library has not touched the
           configuration yet: bring the table home, rescale it as the reference does
           (moves.c:135-139), and hand it to fresh handles on the new grid. */
        hsmc_gpu_info gi;
        hs_gpu_check(hsmc_gpu_get_info(s->gpu, &gi));
        if (HS_ROOT(s) && getenv("HSMC_DEBUG_MP"))
          fprintf(stderr, "[hsmc_b200] volume move changes the cell grid (box %.4f -> %.4f): slabs redistributed\n", gi.box[0], nb[0]);
        s->mirror_current = false;
        hs_gpu_pull(s);
        if (HS_ROOT(s)) {
          for (int i = 0; i < N; i++)
            for (int k = 1; k <= 3; k++) {
              double v = s->conf[i][k] * sf;
              if (v > nb[k - 1]) v -= nb[k - 1]; else if (v < 0.0) v += nb[k - 1];
              s->conf[i][k] = v;
            }
        }
        hs_mp_barrier(&s->mp);
        int64_t cnt[6];
        hs_gpu_check(hsmc_gpu_counters(s->gpu, cnt));
        s->philox_sweeps = gi.sweeps_done;
        s->carry_moves[0] += cnt[0]; s->carry_moves[1] += cnt[1]; s->carry_moves[2] += cnt[2];
        s->carry_moves[3] += cnt[3]; s->carry_moves[4] += cnt[4]; s->carry_moves[5] += cnt[5];
        hs_gpu_close(s);
        hs_gpu_open(s);
      } else {
        hs_gpu_check(rc);
      }
      s->mirror_current = false;
    }
  }
  hs_gpu_check(hsmc_gpu_add_vol_move(s->gpu, accepted));
}

/* counters of handles retired by a slab redistribution are carried over */
void hs_counters(hs_sim *s, int64_t out[6]) {
  hs_gpu_check(hsmc_gpu_counters(s->gpu, out));
  for (int k = 0; k < 6; k++) out[k] += s->carry_moves[k];
}
void hs_reset_counters(hs_sim *s) {
  hs_gpu_check(hsmc_gpu_reset_counters(s->gpu));
  memset(s->carry_moves, 0, sizeof(s->carry_moves));
}
