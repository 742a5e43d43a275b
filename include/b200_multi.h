/* include/b200_multi.h -- one handle over several GPUs of a node, plain C ABI.
 *
 * The reference's GPU scorer drives every visible device from one process (Marv::getDeviceIds, lib/libmarv/src/marv.cu:103-111;
 * per-GPU DB partitions + result merge, lib/libmarv/src/cudasw4.cuh:2097-2230); its CPU modules split work by query range
 * (DBReader::decomposeDomainByAminoAcid, src/commons/DBReader.cpp:1108).  b200_multi offers both splits behind one handle of
 * per-device b200_ctx (include/b200_align.h), one host thread per device inside every call, no data-path collective:
 *
 *   query-sharded  (shard_targets = 0)  DB replicated on every device; a batch of queries is cut into contiguous ranges balanced by
 *                                       the sum of query lengths; every (query, target) pair is independent, results land in the
 *                                       caller's arrays at the query's own position.  The throughput layout (BASELINE configs 3-5).
 *   target-sharded (shard_targets = 1)  the DB is cut into contiguous id ranges balanced by residues; every device scans every query
 *                                       against its slice and the per-device top lists are merged with the comparator
 *                                       (score desc, global id asc) -- hit_t::compareHitsByScoreAndId -- so the result does not
 *                                       depend on the number of devices.  The latency layout (one query at a time, Marv::scan) and
 *                                       the one for a DB that does not fit one GPU.
 *
 * Multi-process layouts (one rank per GPU, torch.distributed / MPI) use one b200_ctx per rank instead and gather hit lists with
 * NCCL (bench.py --gpus N, mmseqs2_b200/sharding.py).
 */
#ifndef B200_MULTI_H
#define B200_MULTI_H

#include <stddef.h>
#include <stdint.h>

#include "b200_align.h"
#include "b200_alignment.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_multi b200_multi;

/* devices == NULL: every visible device (n_devices ignored).  The same device may be listed more than once (two contexts on one
 * GPU): useful to exercise the sharding logic on a single-GPU box. */
int b200_multi_create(const int *devices, int n_devices, b200_multi **out);
void b200_multi_destroy(b200_multi *m);
int b200_multi_size(const b200_multi *m);
b200_ctx *b200_multi_ctx(b200_multi *m, int i);
const char *b200_multi_last_error(const b200_multi *m);

/* residues / offsets as b200_db_load.  shard_targets selects the split (see above). */
int b200_multi_db_load(b200_multi *m, const uint8_t *residues, const uint64_t *offsets, uint64_t n_seq, int alphabet, int shard_targets);
/* the padded GPU DB of makepaddedseqdb as Marv::loadDb takes it (b200_db_load_padded) */
int b200_multi_db_load_padded(b200_multi *m, const uint8_t *data, const size_t *offsets, const int32_t *lengths, uint64_t n_seq,
                              int alphabet, int shard_targets);
/* replicated, mask bit stripped (b200_db_load_padded_unmasked): the target DB of the `align` module */
int b200_multi_db_load_padded_unmasked(b200_multi *m, const uint8_t *data, const size_t *offsets, const int32_t *lengths, uint64_t n_seq,
                                       int alphabet);

/* b200_ungapped_scan over all devices; hits [n_queries][max_hits], n_hits [n_queries], global target ids. */
int b200_multi_ungapped_scan(b200_multi *m, const b200_query *queries, int n_queries, int min_score_excl, uint32_t max_hits,
                             b200_hit *hits, uint32_t *n_hits);

/* b200_align_batch over all devices (query-sharded DBs only): arguments and outputs exactly as b200_align_batch. */
int b200_multi_align_batch(b200_multi *m, const int16_t *sub_matrix, const double *p_back, int alphabet,
                           const uint8_t *query_residues, const uint64_t *query_offsets, const uint32_t *query_keys,
                           uint32_t n_queries, const uint64_t *hit_offsets, const uint32_t *hit_targets,
                           const uint32_t *target_keys, const b200_align_params *params, const b200_evalue_params *evalue,
                           b200_result *results, uint32_t *n_results, char *bt_pool, uint64_t bt_cap, uint64_t *n_alignments);

/* ---- the sharding arithmetic on its own (host only, no device) ------------------------------------------------------------------------
 * bounds[parts + 1]: contiguous ranges of [0, n) balanced by weight -- the query ranges (weights = query lengths) and target slices
 * (weights = sequence length + 1) of the calls above. */
void b200h_balanced_ranges(const uint64_t *weights, uint64_t n, int parts, uint64_t *bounds);
/* One query's per-device top lists (local ids, lists[d][0..counts[d])) -> the global list: ids + first_id[d], (score desc, global id
 * asc), cut to max_hits.  Returns the number of hits written to out[max_hits]. */
uint32_t b200h_merge_top_hits(const b200_hit *const *lists, const uint32_t *counts, const uint64_t *first_id, int n_lists, uint32_t max_hits,
                              b200_hit *out);

#ifdef __cplusplus
}
#endif
#endif
