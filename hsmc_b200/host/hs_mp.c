/* hs_mp.c -- process-per-GPU plumbing of the host driver (see hs_mp.h). */
#define _GNU_SOURCE
#include "hs_mp.h"

#include <pthread.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <unistd.h>

struct hs_mp_shared {
  pthread_barrier_t bar;
  pid_t pid[HS_MP_MAX_RANKS];
  unsigned char id[128];
  unsigned char blob[HS_MP_MAX_RANKS][64];
  uint64_t scratch[HS_MP_MAX_RANKS][HS_MP_SCRATCH];
};

/* rank 0 watches its children: a rank that dies (crash, CUDA/NCCL abort, kill) would otherwise leave
   the others waiting in a barrier or an NCCL call for ever */
static struct hs_mp_shared *g_sh = NULL;
static int g_world = 0;
static volatile sig_atomic_t g_child_ok[HS_MP_MAX_RANKS];

static void on_sigchld(int sig) {
  (void)sig;
  int saved = errno, st;
  pid_t p;
  while ((p = waitpid(-1, &st, WNOHANG)) > 0) {
    int r = -1;
    for (int k = 1; k < g_world; k++)
      if (g_sh->pid[k] == p) r = k;
    if (r < 0) continue;
    if (WIFEXITED(st) && WEXITSTATUS(st) == 0) { g_child_ok[r] = 1; continue; }
    static const char msg[] = "ERROR: a GPU rank of this run died; stopping the others\n";
    if (write(STDERR_FILENO, msg, sizeof(msg) - 1) < 0) { /* nothing left to do about it */ }
    for (int k = 1; k < g_world; k++)
      if (k != r && !g_child_ok[k] && g_sh->pid[k] > 0) kill(g_sh->pid[k], SIGTERM);
    _exit(EXIT_FAILURE);
  }
  errno = saved;
}

int hs_mp_start(hs_mp *mp, int world, int64_t n_rows) {
  memset(mp, 0, sizeof(*mp));
  mp->world = world < 1 ? 1 : world;
  if (mp->world == 1) return 0;
  /* stdout is the reference's report: NCCL's own messages (version banner under NCCL_DEBUG=VERSION, ...)
     go to stderr unless the user chose a file for them */
  setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
  if (mp->world > HS_MP_MAX_RANKS) { printf("ERROR: at most %d GPUs\n", HS_MP_MAX_RANKS); exit(EXIT_FAILURE); }
  size_t bytes = sizeof(struct hs_mp_shared) + (size_t)n_rows * 4 * sizeof(double);
  void *m = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (m == MAP_FAILED) { perror("Failed allocation of the shared particle table"); exit(EXIT_FAILURE); }
  mp->sh = m;
  mp->table = (double (*)[4])((char *)m + sizeof(struct hs_mp_shared));
  mp->table_rows = n_rows;
  pthread_barrierattr_t a;
  pthread_barrierattr_init(&a);
  pthread_barrierattr_setpshared(&a, PTHREAD_PROCESS_SHARED);
  pthread_barrier_init(&mp->sh->bar, &a, (unsigned)mp->world);
  pthread_barrierattr_destroy(&a);
  fflush(NULL);
  mp->sh->pid[0] = getpid();
  g_sh = mp->sh;
  g_world = mp->world;
  struct sigaction sa;
  memset(&sa, 0, sizeof(sa));
  sa.sa_handler = on_sigchld;
  sa.sa_flags = SA_RESTART | SA_NOCLDSTOP;
  sigaction(SIGCHLD, &sa, NULL);
  for (int r = 1; r < mp->world; r++) {
    pid_t p = fork();
    if (p < 0) { perror("fork"); hs_mp_abort(mp); exit(EXIT_FAILURE); }
    if (p == 0) {
      signal(SIGCHLD, SIG_DFL);
      prctl(PR_SET_PDEATHSIG, SIGTERM);      /* no orphans holding a GPU if rank 0 goes away */
      if (getppid() != mp->sh->pid[0]) _exit(EXIT_FAILURE);
      mp->rank = r;
      mp->sh->pid[r] = getpid();
      /* only rank 0 reports; the others run the same program silently */
      if (!freopen("/dev/null", "w", stdout)) _exit(EXIT_FAILURE);
      return r;
    }
    mp->sh->pid[r] = p;
  }
  return 0;
}

void hs_mp_barrier(hs_mp *mp) {
  if (mp->world > 1) pthread_barrier_wait(&mp->sh->bar);
}

void hs_mp_bcast_id(hs_mp *mp, void *buf, int bytes) {
  if (mp->world == 1) return;
  if (mp->rank == 0) memcpy(mp->sh->id, buf, (size_t)bytes);
  hs_mp_barrier(mp);
  if (mp->rank != 0) memcpy(buf, mp->sh->id, (size_t)bytes);
  hs_mp_barrier(mp);
}

void hs_mp_exchange_blobs(hs_mp *mp, const void *mine, const void **left, const void **right) {
  memcpy(mp->sh->blob[mp->rank], mine, 64);
  hs_mp_barrier(mp);
  *left = mp->sh->blob[(mp->rank + mp->world - 1) % mp->world];
  *right = mp->sh->blob[(mp->rank + 1) % mp->world];
}

uint64_t *hs_mp_scratch(hs_mp *mp, int rank) { return mp->sh->scratch[rank]; }

void hs_mp_abort(hs_mp *mp) {
  if (mp->world == 1 || !mp->sh) return;
  for (int r = 0; r < mp->world; r++)
    if (r != mp->rank && mp->sh->pid[r] > 0) kill(mp->sh->pid[r], SIGTERM);
}

int hs_mp_finish(hs_mp *mp) {
  if (mp->world == 1) return 0;
  fflush(NULL);
  if (mp->rank != 0) _exit(EXIT_SUCCESS);
  int bad = 0;
  for (int r = 1; r < mp->world; r++) {
    int st = 0;
    while (!g_child_ok[r]) {
      pid_t p = waitpid(mp->sh->pid[r], &st, 0);
      if (p == mp->sh->pid[r]) { if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1; break; }
      if (p < 0 && errno == ECHILD) { if (!g_child_ok[r]) bad = 1; break; }   /* reaped by the handler */
      if (p < 0 && errno != EINTR) { bad = 1; break; }
    }
  }
  return bad;
}
