/* hs_fastio.h -- parallel writer of the reference's configuration snapshots.
 *
 * The reference writes `config_%06d.dat.gz` with one gzprintf per particle
 * (io_config.c:165-178): at N = 16.8M that is 16.8M printf + a single-threaded deflate of
 * ~0.7 GB of text per snapshot.  Here the particle table is cut into chunks; worker threads
 * format each chunk with an exact "%.8f" formatter and deflate it into its own gzip member,
 * and the members are written in order.  A sequence of gzip members is a valid gzip file
 * (the reference itself produces one whenever it appends a sample with gzopen(...,"a")),
 * so every reader of the reference's files reads these; the decompressed bytes are
 * identical to the reference's.
 */
#ifndef HS_FASTIO_H
#define HS_FASTIO_H

#include <stddef.h>

/* "%.8f" of x into out (no terminator); returns the number of characters.  Bit-for-bit the
   glibc printf result in round-to-nearest mode (exact binary value, ties to even). */
int hs_fmt_f8(char *out, double x);

/* One snapshot = header (sweep, N, box; io_config.c:165-173) + N lines "%d %.8f %.8f %.8f".
   append = 0 truncates the file, 1 appends a sample.  threads <= 0: one per online core
   (HSMC_IO_THREADS overrides).  Returns 0, or -1 with errno set. */
int hs_fastio_write_config(const char *name, int append, int sweep, int n, const double box[3],
                           const double (*conf)[4], int threads);

/* The same writer fed piecewise: `fetch(ctx, first_row, n_rows)` must fill conf[first_row .. first_row + n_rows)
   (rows in id order) and return 0; it is called from a producer thread for consecutive slices of `slice_rows` rows
   while the workers format and deflate the rows that have already arrived -- the device-to-host copy of a large
   table overlaps its compression.  After the call conf holds the whole table. */
typedef int (*hs_fastio_fetch_fn)(void *ctx, long long first_row, long long n_rows);
int hs_fastio_write_config_stream(const char *name, int append, int sweep, int n, const double box[3],
                                  const double (*conf)[4], int threads, hs_fastio_fetch_fn fetch, void *ctx,
                                  int slice_rows);

#endif
