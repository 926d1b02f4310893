/* hs_mp.h -- one process per GPU for the drop-in host driver (`hsmc_b200 -g K`).
 *
 * The reference is a single serial process.  For slab-decomposed runs (x-slabs, one
 * rank per GPU) the driver forks K-1 copies of itself BEFORE any CUDA call; every copy runs the
 * same deterministic host program (same input, same MT19937 stream, all device verdicts and
 * counters are all-reduced by the library), rank 0 alone talks to stdout and to the output
 * files.  What the ranks share lives in one anonymous MAP_SHARED mapping made before the fork:
 * the NCCL id and the NVLink window blobs (hsmc_gpu_nccl_id / hsmc_gpu_ipc_export), a process-
 * shared barrier, scratch for sharded observables, and the host mirror of the particle table
 * ({id,x,y,z} rows, sim_info.h:20), which each rank refreshes with the rows it owns.
 */
#ifndef HS_MP_H
#define HS_MP_H

#include <stdint.h>
#include <sys/types.h>

#define HS_MP_MAX_RANKS 16
#define HS_MP_SCRATCH 8192          /* uint64 words of scratch per rank (sharded histograms) */

typedef struct hs_mp {
  int rank, world;
  struct hs_mp_shared *sh;          /* NULL when world == 1 */
  double (*table)[4];               /* shared host mirror (world > 1) */
  int64_t table_rows;
} hs_mp;

/* world <= 1: no-op (rank 0 of 1).  Otherwise: map the shared block + an n_rows table, fork
   world-1 children (ranks 1..), silence their stdout.  Returns this process's rank. */
int hs_mp_start(hs_mp *mp, int world, int64_t n_rows);
void hs_mp_barrier(hs_mp *mp);
/* rank `root` fills buf[0..bytes) before the call; everyone has it after (bytes <= 128) */
void hs_mp_bcast_id(hs_mp *mp, void *buf, int bytes);
/* every rank contributes a 64-byte blob; returns pointers to the left/right neighbours' blobs */
void hs_mp_exchange_blobs(hs_mp *mp, const void *mine, const void **left, const void **right);
/* per-rank scratch area (HS_MP_SCRATCH words) */
uint64_t *hs_mp_scratch(hs_mp *mp, int rank);
/* a rank that hits a fatal error takes the others down instead of leaving them in a barrier */
void hs_mp_abort(hs_mp *mp);
/* children: _exit(0); rank 0: wait for the children, nonzero if any failed */
int hs_mp_finish(hs_mp *mp);

#endif
