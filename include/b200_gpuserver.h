/* include/b200_gpuserver.h -- a persistent B200 scan server behind the reference's `gpuserver` shared-memory protocol
 * (SURVEY.md 8f row 1): GPUSharedMemory, src/commons/GpuUtil.h:9-52 (layout), GpuUtil.cpp:40-63 (alloc), the server loop of
 * src/util/gpuserver.cpp:77-91 and the client side of src/prefiltering/ungappedprefilter.cpp:209-250.
 *
 * A client (the reference's `ungappedprefilter --gpu-server 1`, or anything following the same state machine) maps the segment,
 * claims it IDLE -> RESERVED, copies the numeric query and its [21][L] int8 profile, sets READY, waits for DONE, reads resultLen
 * Marv::Result records {id, score, qEndPos, dbEndPos}, and sets IDLE again.  The server answers every READY request with the
 * hits of b200_ungapped_scan on the DB loaded in ctx: score > min_score_excl, ordered (score desc, id asc), at most
 * maxResListLen; qEndPos / dbEndPos are 0 (AlignmentType::GAPLESS).  The SSW profile bias the saturating scorer needs is
 * recovered exactly from the request: profile[a][j] - subMatrix[a][q[j]] is the rounded composition bias of position j.
 */
#ifndef B200_GPUSERVER_H
#define B200_GPUSERVER_H

#include <stdint.h>

#include "b200_align.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_server b200_server;

/* Creates (shm_open O_CREAT + ftruncate + mmap) and initialises the segment as GPUSharedMemory::alloc does.
 * shm_name: POSIX shared memory name, e.g. "/123456" (the reference uses the decimal hash of the DB path, GpuUtil.cpp:18-36). */
int b200_gpuserver_create(b200_ctx *ctx, const char *shm_name, unsigned max_seq_len, unsigned max_res_list_len,
                          const int16_t *sub_matrix, int alphabet, int min_score_excl, b200_server **out);
/* Serves requests until b200_gpuserver_stop() is called (max_requests == 0) or max_requests have been answered.
 * Returns B200_OK, or the error of a failed scan (the segment is left in DONE with resultLen 0 in that case). */
int b200_gpuserver_serve(b200_server *srv, uint64_t max_requests);
void b200_gpuserver_stop(b200_server *srv);          /* async-signal-safe: only sets a flag */
uint64_t b200_gpuserver_served(const b200_server *srv);
/* serverExit = true, munmap, shm_unlink (gpuserver.cpp:94-97) */
void b200_gpuserver_destroy(b200_server *srv);

#ifdef __cplusplus
}
#endif
#endif
