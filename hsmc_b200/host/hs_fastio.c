/* hs_fastio.c -- parallel, byte-compatible writer of configuration snapshots (see hs_fastio.h).
 *
 * File format: io_config.c:165-178 of the reference.  Nothing here is on the bit-exact
 * physics surface, but the decompressed file must equal the reference's byte for byte, so
 * the "%.8f" formatter below is exact rather than fast-and-approximate: it rounds the exact
 * binary value of the double to 8 decimals, ties to even, which is what glibc's printf does
 * in the default rounding mode (tests/test_fastio_cpu.py checks it against printf, against
 * ties such as k/512, carries such as 0.999999995, and against the reference's own
 * write_config()).
 */
#define _GNU_SOURCE
#include "hs_fastio.h"

#include <errno.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

/* ---- exact "%.8f" ---------------------------------------------------------------------- */
static inline char *put_u64(char *p, uint64_t v) {
  char tmp[20];
  int k = 0;
  do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (k) *p++ = tmp[--k];
  return p;
}

/* round(f * 1e8) for 0 <= f < 1, on the exact value of f, ties to even */
static inline uint64_t frac_1e8(double f) {
  if (f == 0.0) return 0;
  int e;
  const double m = frexp(f, &e);                      /* f = m * 2^e, 0.5 <= m < 1 */
  const uint64_t M = (uint64_t)ldexp(m, 53);          /* exact: 53-bit integer mantissa */
  const int s = 53 - e;                               /* f = M * 2^-s, s >= 53 */
  if (s > 120) return 0;                              /* f < 2^-67: f*1e8 < 1e-12, rounds to 0 */
  const unsigned __int128 P = (unsigned __int128)M * 100000000u;   /* < 2^80 */
  const unsigned __int128 one = 1;
  uint64_t q = (uint64_t)(P >> s);
  const unsigned __int128 rem = P & ((one << s) - 1), half = one << (s - 1);
  if (rem > half || (rem == half && (q & 1))) q++;
  return q;
}

int hs_fmt_f8(char *out, double x) {
  char *p = out;
  if (!(fabs(x) < 9.0e15)) return sprintf(out, "%.8f", x);   /* huge / inf / nan: not a coordinate */
  if (signbit(x)) { *p++ = '-'; x = -x; }
  const double ip = floor(x);
  uint64_t I = (uint64_t)ip;
  uint64_t q = frac_1e8(x - ip);                        /* x - floor(x) is exact */
  if (q == 100000000u) { q = 0; I++; }
  p = put_u64(p, I);
  *p++ = '.';
  for (int k = 7; k >= 0; k--) { p[k] = (char)('0' + q % 10); q /= 10; }
  return (int)(p + 8 - out);
}

/* ---- chunked format + deflate ---------------------------------------------------------------- */
#define ROWS_PER_CHUNK 32768
#define LINE_MAX_BYTES 1400      /* "%d" + 3 x "%.8f" of any double (<= 1 + 309 + 1 + 8 each) */

/* Output ring: chunk c lands in slot c % window; a worker may only take chunk c once
   c < written + window, i.e. once the writer has drained the slot's previous tenant.  Slots and
   the workers' text buffers are allocated once per call and reused, so a 16.8M-particle
   snapshot touches a few tens of MB of scratch, not the 0.7 GB of text it produces. */
typedef struct {
  unsigned char *gz;      /* finished gzip member */
  size_t cap, len;
  int chunk;              /* which chunk the slot holds */
  int done;               /* 1 ok, -1 failed */
} chunk_out;

typedef struct {
  const double (*conf)[4];
  int n, nchunks, window, level, strategy;
  const char *header;
  size_t header_len;
  chunk_out *out;
  int avail;              /* rows [0, avail) of conf are valid (streaming source; n when the table is complete) */
  int next;               /* next chunk to hand out */
  int written;            /* chunks already written by the main thread */
  int failed;
  pthread_mutex_t mu;
  pthread_cond_t cv_done, cv_room;
} job;

/* text of chunk c into the worker's own buffer (*text, *cap; grown if ever needed) */
static int format_chunk(const job *J, int c, char **text, size_t *cap_io, size_t *len) {
  const int r0 = c * ROWS_PER_CHUNK;
  const int r1 = (r0 + ROWS_PER_CHUNK < J->n) ? r0 + ROWS_PER_CHUNK : J->n;
  size_t cap = *cap_io;
  char *buf = *text;
  char *p = buf;
  if (c == 0) { memcpy(p, J->header, J->header_len); p += J->header_len; }
  for (int i = r0; i < r1; i++) {
    if ((size_t)(p - buf) + LINE_MAX_BYTES > cap) {
      const size_t used = (size_t)(p - buf);
      cap = cap * 2 + LINE_MAX_BYTES;
      char *nb = realloc(buf, cap);
      if (!nb) return -1;
      buf = nb; p = buf + used;
      *text = buf; *cap_io = cap;
    }
    const double *row = J->conf[i];
    int id = (int)row[0];                               /* "%d", (int)part_conf[ii][0] */
    if (id < 0) { *p++ = '-'; p = put_u64(p, (uint64_t)(-(int64_t)id)); }
    else p = put_u64(p, (uint64_t)id);
    *p++ = ' '; p += hs_fmt_f8(p, row[1]);
    *p++ = ' '; p += hs_fmt_f8(p, row[2]);
    *p++ = ' '; p += hs_fmt_f8(p, row[3]);
    *p++ = '\n';
  }
  *len = (size_t)(p - buf);
  return 0;
}

/* one gzip member into the ring slot (*gz, *cap; grown if ever needed); zs is the worker's own
   stream, initialised once and reset per member */
static int deflate_member(z_stream *zs, const char *text, size_t len, unsigned char **gz, size_t *cap, size_t *gzlen) {
  if (deflateReset(zs) != Z_OK) return -1;
  const size_t bound = deflateBound(zs, (uLong)len) + 64;
  if (bound > *cap) {
    unsigned char *nb = realloc(*gz, bound);
    if (!nb) return -1;
    *gz = nb; *cap = bound;
  }
  zs->next_in = (Bytef *)text;
  zs->avail_in = (uInt)len;
  zs->next_out = *gz;
  zs->avail_out = (uInt)bound;
  const int rc = deflate(zs, Z_FINISH);
  if (rc != Z_STREAM_END) return -1;
  *gzlen = bound - zs->avail_out;
  return 0;
}

static void *worker(void *arg) {
  job *J = arg;
  size_t cap = (size_t)ROWS_PER_CHUNK * 48 + LINE_MAX_BYTES + J->header_len;
  char *text = malloc(cap);
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  const int zok = deflateInit2(&zs, J->level, Z_DEFLATED, 15 + 16 /* gzip wrapper */, 8, J->strategy) == Z_OK;
  for (;;) {
    pthread_mutex_lock(&J->mu);
    while (!J->failed && J->next < J->nchunks && J->next >= J->written + J->window)
      pthread_cond_wait(&J->cv_room, &J->mu);            /* do not run far ahead of the writer */
    if (J->failed || J->next >= J->nchunks) { pthread_mutex_unlock(&J->mu); break; }
    const int c = J->next++;
    {                                                      /* streaming source: wait for the rows of this chunk */
      const int r1 = ((c + 1) * ROWS_PER_CHUNK < J->n) ? (c + 1) * ROWS_PER_CHUNK : J->n;
      while (!J->failed && J->avail < r1) pthread_cond_wait(&J->cv_done, &J->mu);
    }
    pthread_mutex_unlock(&J->mu);

    chunk_out *o = &J->out[c % J->window];
    size_t len = 0, gzlen = 0;
    int rc = (text && zok) ? format_chunk(J, c, &text, &cap, &len) : -1;
    if (rc == 0) rc = deflate_member(&zs, text, len, &o->gz, &o->cap, &gzlen);

    pthread_mutex_lock(&J->mu);
    o->len = gzlen;
    o->chunk = c;
    o->done = rc == 0 ? 1 : -1;
    if (rc) J->failed = 1;
    pthread_cond_broadcast(&J->cv_done);
    pthread_cond_broadcast(&J->cv_room);
    pthread_mutex_unlock(&J->mu);
  }
  if (zok) deflateEnd(&zs);
  free(text);
  return NULL;
}

/* producer of a streamed snapshot: fetches consecutive slices of the table and publishes how far it got */
typedef struct { job *J; hs_fastio_fetch_fn fetch; void *ctx; int slice; } feeder;

static void *feed(void *arg) {
  feeder *F = arg;
  job *J = F->J;
  for (int r0 = 0; r0 < J->n; r0 += F->slice) {
    const int m = (r0 + F->slice < J->n) ? F->slice : J->n - r0;
    const int rc = F->fetch(F->ctx, r0, m);
    pthread_mutex_lock(&J->mu);
    if (rc) J->failed = 1; else J->avail = r0 + m;
    pthread_cond_broadcast(&J->cv_done);
    pthread_cond_broadcast(&J->cv_room);
    const int stop = J->failed;
    pthread_mutex_unlock(&J->mu);
    if (stop) break;
  }
  return NULL;
}

int hs_fastio_write_config(const char *name, int append, int sweep, int n, const double box[3],
                           const double (*conf)[4], int threads) {
  return hs_fastio_write_config_stream(name, append, sweep, n, box, conf, threads, NULL, NULL, 0);
}

int hs_fastio_write_config_stream(const char *name, int append, int sweep, int n, const double box[3],
                                  const double (*conf)[4], int threads, hs_fastio_fetch_fn fetch, void *ctx,
                                  int slice_rows) {
  if (n < 0) { errno = EINVAL; return -1; }
  const char *env = getenv("HSMC_IO_THREADS");
  if (env && atoi(env) > 0) threads = atoi(env);
  if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (threads < 1) threads = 1;
  /* The text is ~85 % decimal digits of coordinates: LZ77 finds next to no matches in them, so
     entropy coding alone (Z_HUFFMAN_ONLY) gives a file slightly SMALLER than gzopen(name,"w")'s
     level-6 search at ~10x its speed (scripts/io_bench.py).  HSMC_IO_LEVEL=1..9 selects zlib's
     ordinary match search at that level instead. */
  int level = Z_DEFAULT_COMPRESSION, strategy = Z_HUFFMAN_ONLY;
  env = getenv("HSMC_IO_LEVEL");
  if (env && atoi(env) >= 1 && atoi(env) <= 9) { level = atoi(env); strategy = Z_DEFAULT_STRATEGY; }

  /* header, io_config.c:165-173 */
  char header[256];
  int hl = snprintf(header, sizeof(header), "# Sweep number\n%d\n# Number of particles\n%d\n# Simulation box size\n", sweep, n);
  for (int k = 0; k < 3; k++) { hl += hs_fmt_f8(header + hl, box[k]); header[hl++] = '\n'; }
  hl += snprintf(header + hl, sizeof(header) - (size_t)hl, "# Configuration\n");

  job J;
  memset(&J, 0, sizeof(J));
  J.conf = conf; J.n = n; J.level = level; J.strategy = strategy;
  J.avail = fetch ? 0 : n;
  J.header = header; J.header_len = (size_t)hl;
  J.nchunks = n == 0 ? 1 : (n + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK;
  if (threads > J.nchunks) threads = J.nchunks;
  J.window = 4 * threads;
  J.out = calloc((size_t)J.window, sizeof(chunk_out));
  if (!J.out) return -1;
  FILE *f = fopen(name, append ? "ab" : "wb");
  if (!f) { free(J.out); return -1; }
  pthread_mutex_init(&J.mu, NULL);
  pthread_cond_init(&J.cv_done, NULL);
  pthread_cond_init(&J.cv_room, NULL);

  pthread_t *tid = calloc((size_t)threads, sizeof(pthread_t));
  int started = 0;
  if (tid)
    for (; started < threads; started++)
      if (pthread_create(&tid[started], NULL, worker, &J)) break;
  int rc = 0;
  if (started == 0) { rc = -1; J.failed = 1; }
  feeder F = {&J, fetch, ctx, slice_rows > 0 ? slice_rows : (1 << 20)};
  pthread_t feed_tid;
  int feeding = 0;
  if (fetch && rc == 0) {
    if (pthread_create(&feed_tid, NULL, feed, &F)) {     /* no thread: fetch everything here, then go on as usual */
      if (fetch(ctx, 0, n)) { rc = -1; J.failed = 1; }
      pthread_mutex_lock(&J.mu);
      J.avail = n;
      pthread_cond_broadcast(&J.cv_done);
      pthread_mutex_unlock(&J.mu);
    } else feeding = 1;
  }

  /* ordered writer */
  for (int c = 0; c < J.nchunks && rc == 0; c++) {
    chunk_out *o = &J.out[c % J.window];
    pthread_mutex_lock(&J.mu);
    while (!(o->done && o->chunk == c) && !J.failed) pthread_cond_wait(&J.cv_done, &J.mu);
    const int st = (o->chunk == c) ? o->done : 0;
    pthread_mutex_unlock(&J.mu);
    if (st != 1) { rc = -1; break; }
    if (fwrite(o->gz, 1, o->len, f) != o->len) rc = -1;
    pthread_mutex_lock(&J.mu);
    o->done = 0;
    J.written = c + 1;
    if (rc) J.failed = 1;
    pthread_cond_broadcast(&J.cv_room);
    pthread_mutex_unlock(&J.mu);
  }
  pthread_mutex_lock(&J.mu);
  if (rc) J.failed = 1;
  pthread_cond_broadcast(&J.cv_room);
  pthread_mutex_unlock(&J.mu);
  pthread_mutex_lock(&J.mu);
  pthread_cond_broadcast(&J.cv_done);                      /* (workers waiting for rows see `failed`) */
  pthread_mutex_unlock(&J.mu);
  for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
  if (feeding) pthread_join(feed_tid, NULL);
  for (int c = 0; c < J.window; c++) free(J.out[c].gz);
  free(tid);
  free(J.out);
  pthread_mutex_destroy(&J.mu);
  pthread_cond_destroy(&J.cv_done);
  pthread_cond_destroy(&J.cv_room);
  if (fclose(f)) rc = -1;
  return rc;
}
