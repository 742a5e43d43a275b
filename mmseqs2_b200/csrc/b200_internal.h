// mmseqs2_b200/csrc/b200_internal.h -- shared between the translation units of libb200align.so (not installed).
#ifndef B200_INTERNAL_H
#define B200_INTERNAL_H
#include "b200_align.h"

#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#define CU_TRY(ctx, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                     \
            return B200_ERR_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

static inline uint64_t round_up(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }

// Device scratch.  cudaMalloc/cudaFree cost hundreds of microseconds each and synchronise the device, so released
// buffers go to a small per-process free list (best fit, bounded) instead of back to the driver; one-shot calls such as
// b200_sw_score then run without touching the allocator after warm-up.
struct DevPool {
    struct Slot { void *p; size_t cap; int dev; };
    std::mutex mu;
    std::vector<Slot> free_list;
    size_t held = 0;
    bool dirty = false;   // a buffer was returned since the last device-wide synchronisation
    static DevPool &get() { static DevPool pool; return pool; }
    // Buffers come back without a stream synchronisation (cudaFree would have synchronised implicitly), so work queued on the
    // releasing context's stream may still be reading them.  Before the first buffer is handed out again after any release, the
    // device is synchronised once: a few microseconds on an idle device, and it makes reuse across contexts / streams safe.
    void *take(size_t n, int dev, size_t *cap_out) {
        std::lock_guard<std::mutex> lk(mu);
        int best = -1;
        for (size_t i = 0; i < free_list.size(); i++)
            if (free_list[i].dev == dev && free_list[i].cap >= n && free_list[i].cap <= 4 * n + (1u << 20) &&
                (best < 0 || free_list[i].cap < free_list[best].cap)) best = (int) i;
        if (best < 0) return nullptr;
        if (dirty) { cudaDeviceSynchronize(); dirty = false; }
        void *p = free_list[best].p;
        *cap_out = free_list[best].cap;
        held -= free_list[best].cap;
        free_list.erase(free_list.begin() + best);
        return p;
    }
    void give(void *p, size_t cap, int dev) {
        std::lock_guard<std::mutex> lk(mu);
        if (free_list.size() >= 256 || held + cap > ((size_t) 32 << 30)) { cudaFree(p); return; }   // of 180 GB HBM
        Slot s = {p, cap, dev};
        dirty = true;
        free_list.push_back(s);
        held += cap;
    }
};

struct DevBuf {  // grow-only device scratch
    void *p = nullptr;
    size_t cap = 0;
    int dev = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        release();
        cudaGetDevice(&dev);
        const size_t want = n + n / 4 + 256;
        p = DevPool::get().take(want, dev, &cap);
        if (p) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release() { if (p) DevPool::get().give(p, cap, dev); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // result downloads of resident jobs: waits on the job's own event, not on later launches
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};   // the capacity classes of one scan batch run on stream + side[] concurrently
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[16];
    std::mutex mu;
    std::string err;
    uint64_t launches = 0;
    float last_kernel_ms = 0.f;  // device time of the kernels of the last one-shot call that measures it (b200_last_kernel_ms)
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    uint64_t hbm = 0;
    int max_smem_optin = 0;
    // resident target DB
    uint8_t *d_res = nullptr;     // residues, every sequence 16-byte aligned and padded with code `alphabet`
    uint64_t *d_off = nullptr;    // [n_seq] byte offset of sequence i in d_res
    int32_t *d_len = nullptr;     // [n_seq]
    uint32_t *d_order = nullptr;  // [n_seq] ids sorted by length descending (scan schedule)
    std::vector<int32_t> h_len;
    uint64_t n_seq = 0, n_res = 0;
    int alphabet = 0;
    int max_len = 0;
    // resident ASCII copy of a sequence DB for rescorediagonal (b200_db_load_ascii): bytes as stored, concatenated
    uint8_t *d_ares = nullptr;
    uint64_t *d_aoff = nullptr;   // [n_aseq + 1]
    uint64_t n_aseq = 0;
    std::vector<uint64_t> h_aoff;
    // scratch
    DevBuf raw, pad, qdesc, dense, hits, nhits, pairs, items, out4, bnd, ids, diags, counts, rawout, counter;
};

static inline int b200_set_err(b200_ctx *ctx, int code, const char *msg) { ctx->err = msg; return code; }

// b200_sw_backtrace with the ops left in one dense pool (b200_backtrace.cu); the public entry point scatters them into the caller's slots
int b200_sw_backtrace_impl(b200_ctx *ctx, const b200_query *queries, const uint8_t *const *query_seqs, int nq, const b200_pair *pairs,
                           uint64_t n, int gap_open, int gap_extend, const b200_sw_aln *alns, b200_sw_bt *out, uint32_t *cigars,
                           const uint64_t *cigar_offsets, std::vector<uint32_t> *pool_out, std::vector<uint64_t> *base_out);

// common body of b200_db_load / b200_db_load_padded / b200_db_load_ascii (b200_align.cu); the caller holds ctx->mu
int b200_db_load_impl(b200_ctx *ctx, const uint8_t *base, const uint64_t *starts, const int32_t *lens, uint64_t n_seq, int alphabet,
                      int mask_from, uint64_t n_res, bool strip_mask);

void b200_ascii_db_free(b200_ctx *ctx);   // b200_rescore.cu

#endif  // B200_INTERNAL_H
