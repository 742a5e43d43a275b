// mmseqs2_b200/csrc/b200_gpuserver.cpp -- the `gpuserver` shared-memory protocol served by the B200 scan (include/b200_gpuserver.h).
#include "b200_gpuserver.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "b200_host.h"
#include "b200_internal.h"

namespace {

// the segment header exactly as GPUSharedMemory lays it out (GpuUtil.h:9-25); the payload follows at the offsets it records
struct ShmHeader {
    enum State { IDLE, RESERVED, READY, DONE };
    unsigned int maxSeqLen;
    unsigned int maxResListLen;
    std::atomic<int> state;
    std::atomic<bool> serverExit;
    unsigned int queryOffset;
    unsigned int queryLen;
    unsigned int resultsOffset;
    unsigned int resultLen;
    unsigned int profileOffset;
};
static_assert(sizeof(ShmHeader) == 36, "GPUSharedMemory header is 36 bytes with 4-byte std::atomic<int> and 1-byte std::atomic<bool>");
struct ShmResult { unsigned int id; int score; int qEndPos; int dbEndPos; };   // Marv::Result, lib/libmarv/src/marv.h:37-45

size_t shm_bytes(unsigned max_seq_len, unsigned max_res) {      // GPUSharedMemory::calculateSize
    return sizeof(ShmHeader) + (size_t) max_seq_len + sizeof(ShmResult) * (size_t) max_res + (size_t) 21 * max_seq_len;
}

}  // namespace

struct b200_server {
    b200_ctx *ctx = nullptr;
    std::string name;
    ShmHeader *hdr = nullptr;
    size_t bytes = 0;
    std::vector<int16_t> mat;
    int A = 0, min_score = 0;
    std::atomic<bool> stop{false};
    std::atomic<uint64_t> served{0};
};

extern "C" {

int b200_gpuserver_create(b200_ctx *ctx, const char *shm_name, unsigned max_seq_len, unsigned max_res_list_len,
                          const int16_t *sub_matrix, int alphabet, int min_score_excl, b200_server **out) {
    if (ctx == nullptr || out == nullptr) return B200_ERR_ARG;
    *out = nullptr;
    if (shm_name == nullptr || sub_matrix == nullptr || max_seq_len == 0 || max_res_list_len == 0 || alphabet <= 0 || alphabet > 21)
        return b200_set_err(ctx, B200_ERR_ARG, "b200_gpuserver_create: bad argument");
    if (b200_db_num_seqs(ctx) == 0) return b200_set_err(ctx, B200_ERR_NODB, "no target DB loaded");
    const size_t bytes = shm_bytes(max_seq_len, max_res_list_len);
    const int fd = shm_open(shm_name, O_CREAT | O_RDWR, 0666);
    if (fd == -1) return b200_set_err(ctx, B200_ERR_ARG, "b200_gpuserver_create: shm_open failed");
    if (ftruncate(fd, (off_t) bytes) == -1) { close(fd); shm_unlink(shm_name); return b200_set_err(ctx, B200_ERR_NOMEM, "b200_gpuserver_create: ftruncate failed"); }
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { shm_unlink(shm_name); return b200_set_err(ctx, B200_ERR_NOMEM, "b200_gpuserver_create: mmap failed"); }
    ShmHeader *h = new (p) ShmHeader;
    h->maxSeqLen = max_seq_len; h->maxResListLen = max_res_list_len;
    h->state.store(ShmHeader::IDLE); h->serverExit.store(false);
    h->queryOffset = (unsigned) sizeof(ShmHeader);
    h->queryLen = 0;
    h->resultsOffset = h->queryOffset + max_seq_len;
    h->resultLen = 0;
    h->profileOffset = h->resultsOffset + (unsigned) (sizeof(ShmResult) * max_res_list_len);
    b200_server *s = new b200_server();
    s->ctx = ctx; s->name = shm_name; s->hdr = h; s->bytes = bytes; s->A = alphabet; s->min_score = min_score_excl;
    s->mat.assign(sub_matrix, sub_matrix + (size_t) alphabet * alphabet);
    *out = s;
    return B200_OK;
}

int b200_gpuserver_serve(b200_server *s, uint64_t max_requests) {
    if (s == nullptr) return B200_ERR_ARG;
    ShmHeader *h = s->hdr;
    char *base = reinterpret_cast<char *>(h);
    std::vector<b200_hit> hits(h->maxResListLen);
    std::vector<int8_t> cb;
    uint64_t answered = 0;
    int rc_all = B200_OK;
    while (!s->stop.load(std::memory_order_acquire) && (max_requests == 0 || answered < max_requests)) {
        if (h->state.load(std::memory_order_acquire) != ShmHeader::READY) { std::this_thread::yield(); continue; }
        std::atomic_thread_fence(std::memory_order_acquire);
        const int L = (int) h->queryLen;
        const uint8_t *q = reinterpret_cast<const uint8_t *>(base + h->queryOffset);
        const int8_t *prof = reinterpret_cast<const int8_t *>(base + h->profileOffset);
        ShmResult *res = reinterpret_cast<ShmResult *>(base + h->resultsOffset);
        uint32_t n = 0;
        int rc = B200_OK;
        if (L <= 0 || (unsigned) L > h->maxSeqLen) rc = B200_ERR_ARG;
        if (rc == B200_OK) {
            // SSW bias from the request itself: matrix column + one composition bias per position for sequence queries, the profile
            // rule otherwise (the reference client also sends profile_for_alignment of HMM queries) -- b200h_ssw_bias_from_profile
            for (int j = 0; j < L; j++)
                if (q[j] >= s->A) { rc = B200_ERR_ARG; break; }
            if (rc == B200_OK) {
                b200_query bq; bq.profile = prof; bq.qlen = L; bq.bias = b200h_ssw_bias_from_profile(s->mat.data(), s->A, q, L, prof);
                rc = b200_ungapped_scan(s->ctx, &bq, 1, s->min_score, h->maxResListLen, hits.data(), &n, nullptr);
            }
        }
        if (rc != B200_OK) { n = 0; rc_all = rc; }
        for (uint32_t k = 0; k < n; k++) { res[k].id = hits[k].id; res[k].score = hits[k].score; res[k].qEndPos = 0; res[k].dbEndPos = 0; }
        h->resultLen = n;
        std::atomic_thread_fence(std::memory_order_release);
        h->state.store(ShmHeader::DONE, std::memory_order_release);
        answered++;
        s->served.fetch_add(1, std::memory_order_relaxed);
    }
    return rc_all;
}

void b200_gpuserver_stop(b200_server *s) { if (s != nullptr) s->stop.store(true, std::memory_order_release); }
uint64_t b200_gpuserver_served(const b200_server *s) { return s ? s->served.load() : 0; }

void b200_gpuserver_destroy(b200_server *s) {
    if (s == nullptr) return;
    s->hdr->serverExit.store(true, std::memory_order_release);
    std::atomic_thread_fence(std::memory_order_release);
    munmap(s->hdr, s->bytes);
    shm_unlink(s->name.c_str());
    delete s;
}

}  // extern "C"
