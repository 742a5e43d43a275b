// mmseqs2_b200/csrc/b200_rescore.cu -- SURVEY 8(f) row 3: the scorer of `rescorediagonal`, DistanceCalculator::computeUngappedAlignment
// (src/alignment/DistanceCalculator.h:93-174, per-mode scorers :15-37,177-271) on the ASCII sequences of a resident DB, for whole hit
// lists of many queries per call.  The same "gather one diagonal per hit" shape as the per-diagonal prefilter scorer (A1), with the
// five reductions of Parameters::RESCORE_MODE_* (rescorediagonal.cpp:231-236):
//   0 HAMMING        number of identical characters on the diagonal
//   1 SUBSTITUTION   best 0-reset running sum of matrix scores
//   2 ALIGNMENT      the same + the segment that attains it (first maximum; start = position after the last reset)
//   3 END_TO_END     the whole diagonal ('*' at either end skipped), floored at 0
//   4 WINDOW_QUALITY longest stretch with <= 5 mismatches in every window of 20, scored
// One warp per hit.  The 32 lanes fetch 32 diagonal cells at a time (coalesced bytes of query and target, matrix lookups, match
// flags) and park them in shared memory; modes 0/1/3 then reduce in parallel (sum / maximum-subarray, both associative), modes 2 and
// 4 carry order-dependent tie rules and an error window, so lane 0 replays the reference's scalar recurrence over the staged
// values -- a few instructions per cell on data that is already on chip.  Latency/HBM-gather bound like A1: diagonal length bytes
// of each sequence + 6 B hit record in, 28 B out per hit.
#include "b200_internal.h"

#include <cstring>
#include <vector>

namespace {

struct RescoreQuery { uint64_t seq_off; uint64_t hit_begin; int32_t qlen; int32_t pad_; };

constexpr int RS_WARPS = 8;

// one real diagonal of the (query, target) rectangle; out = {score, startPos, endPos, diagonalLen, distToDiagonal, diagonal, idCnt}
template <int MODE>
__device__ __forceinline__ void rescore_one(const uint8_t *__restrict__ q, unsigned qL, const uint8_t *__restrict__ t, unsigned tL, int diagonal,
                                            const int8_t *__restrict__ m, int lane, int *sc_s, uint8_t *eq_s, long long out[7]) {
    const unsigned dist = (unsigned) (diagonal < 0 ? -diagonal : diagonal);
    out[0] = 0; out[1] = -1; out[2] = -1; out[3] = 0; out[4] = dist; out[5] = diagonal; out[6] = 0;
    const uint8_t *a, *b;
    unsigned len;
    if (diagonal >= 0 && dist < qL) { len = tL < qL - dist ? tL : qL - dist; a = q + dist; b = t; }
    else if (diagonal < 0 && dist < tL) { len = tL - dist < qL ? tL - dist : qL; a = q; b = t + dist; }
    else return;
    out[3] = len;
    if (len == 0) return;
    unsigned first = 0, last = len - 1;
    if (MODE >= 3) {
        first = (a[0] == '*' || b[0] == '*') ? 1 : 0;
        if (last > 0 && (a[len - 1] == '*' || b[len - 1] == '*')) last--;
    }
    // ---- parallel part: per-lane partials for the associative modes
    int p_sum = 0, p_pre = 0, p_suf = 0, p_best = 0;       // maximum-subarray tuple of this lane's cells (mode 1), plain sums otherwise
    // ---- sequential state (lane 0 only) for modes 2 and 4
    int s2 = 0, best2 = 0, minPos = -1, st2 = 0, en2 = 0;
    unsigned long long window = 0; unsigned errs = 0, maxLen = 0, curLen = 0, maxEnd = 0, maxStart = 0, start = first;
    int carry_run = 0;       // mode 1: running 0-reset sum entering the current block (all lanes hold it)
    int blk_best = 0;
    for (unsigned i0 = first; i0 <= last; i0 += 32) {
        const unsigned i = i0 + lane;
        const bool on = i <= last;
        int sc = 0; int eq = 0;
        if (on) {
            const unsigned char ca = a[i], cb = b[i];
            eq = ca == cb;
            if (MODE != 0) sc = (int) m[(size_t) ca * 123 + cb];
        }
        if (MODE == 0) { p_sum += eq; continue; }
        if (MODE == 3) { p_sum += sc; continue; }
        if (MODE == 1) {
            // block-level maximum subarray by an ordered shuffle tree over (sum, best prefix, best suffix, best); then stitched to the
            // running sum that enters the block: best = max(best, carry + prefix), carry' = max(block suffix, carry + sum, 0)
            int sum = sc, pre = sc > 0 ? sc : 0, suf = pre, bst = pre;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int rs = __shfl_down_sync(0xffffffffu, sum, o), rp = __shfl_down_sync(0xffffffffu, pre, o);
                const int rf = __shfl_down_sync(0xffffffffu, suf, o), rb = __shfl_down_sync(0xffffffffu, bst, o);
                if ((lane & (2 * o - 1)) == 0) {
                    const int nb = max(max(bst, rb), suf + rp);
                    const int np = max(pre, sum + rp), nf = max(rf, rs + suf);
                    sum += rs; pre = np; suf = nf; bst = nb;
                }
            }
            sum = __shfl_sync(0xffffffffu, sum, 0); pre = __shfl_sync(0xffffffffu, pre, 0);
            suf = __shfl_sync(0xffffffffu, suf, 0); bst = __shfl_sync(0xffffffffu, bst, 0);
            blk_best = max(blk_best, max(bst, carry_run + pre));
            carry_run = max(max(suf, carry_run + sum), 0);
            continue;
        }
        // modes 2 and 4: stage the block, lane 0 replays the scalar recurrence
        sc_s[lane] = sc; eq_s[lane] = (uint8_t) eq;
        __syncwarp();
        if (lane == 0) {
            const unsigned nb = (last - i0 + 1) < 32u ? (last - i0 + 1) : 32u;
            if (MODE == 2) {
                for (unsigned k = 0; k < nb; k++) {
                    s2 += sc_s[k];
                    if (s2 <= 0) { s2 = 0; minPos = (int) (i0 + k); }
                    if (s2 > best2) { best2 = s2; en2 = (int) (i0 + k); st2 = minPos + 1; }
                }
            } else {
                const unsigned W = 20, E = 5;
                const unsigned long long mask = 1ull << (W - 1);
                for (unsigned k = 0; k < nb; k++) {
                    const unsigned ii = i0 + k;
                    if (window & mask) errs -= 1;
                    window <<= 1;
                    if (!eq_s[k]) { window |= 1; errs += 1; }
                    curLen += 1;
                    if (ii >= W - 1 && errs > E) { start = ii - W + 2; curLen = W - 1; }
                    if (curLen > maxLen) { maxStart = start; maxEnd = ii; maxLen = curLen; }
                }
            }
        }
        __syncwarp();
    }
    if (MODE == 0 || MODE == 3) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p_sum += __shfl_xor_sync(0xffffffffu, p_sum, o);
        if (MODE == 0) out[0] = p_sum;
        else { out[0] = p_sum > 0 ? p_sum : 0; out[1] = first; out[2] = last; }
    } else if (MODE == 1) {
        out[0] = blk_best;
    } else if (MODE == 2) {
        out[0] = __shfl_sync(0xffffffffu, best2, 0); out[1] = __shfl_sync(0xffffffffu, st2, 0); out[2] = __shfl_sync(0xffffffffu, en2, 0);
    } else {
        const unsigned ms = __shfl_sync(0xffffffffu, maxStart, 0), me = __shfl_sync(0xffffffffu, maxEnd, 0);
        int s = 0;
        for (unsigned i = ms + lane; i < me; i += 32) s += (int) m[(size_t) a[i] * 123 + b[i]];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        out[0] = (unsigned) s; out[1] = ms; out[2] = me;
    }
    (void) p_pre; (void) p_suf; (void) p_best;
    // identical residues of the reported segment, case-insensitive, as rescorediagonal.cpp:296-301 counts them for the modes that
    // report positions (the caller applies its E-value condition)
    if (MODE >= 2 && out[1] >= 0 && out[2] >= out[1]) {
        int c = 0;
        for (long long i = out[1] + lane; i <= out[2]; i += 32) c += (a[i] & 0xDF) == (b[i] & 0xDF);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        out[6] = c;
    }
}

template <int MODE>
__global__ void __launch_bounds__(RS_WARPS * 32)
rescore_diagonal_kernel(const uint8_t *__restrict__ qseq, const RescoreQuery *__restrict__ rq, int nq, const uint8_t *__restrict__ db,
                        const uint64_t *__restrict__ off, const uint32_t *__restrict__ ids, const uint16_t *__restrict__ diags, uint64_t n,
                        const int8_t *__restrict__ mat, int32_t *__restrict__ out) {
    __shared__ int sc_s[RS_WARPS][32];
    __shared__ uint8_t eq_s[RS_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t warps_total = (uint64_t) gridDim.x * RS_WARPS;
    for (uint64_t h = (uint64_t) blockIdx.x * RS_WARPS + warp; h < n; h += warps_total) {
        int lo = 0, hi = nq - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (rq[mid].hit_begin <= h) lo = mid; else hi = mid - 1; }
        const uint8_t *q = qseq + rq[lo].seq_off;
        const unsigned qL = (unsigned) rq[lo].qlen;
        const uint32_t id = ids[h];
        const uint8_t *t = db + off[id];
        const unsigned tL = (unsigned) (off[id + 1] - off[id]);
        const unsigned diagonal = diags[h];
        // the unsigned short diagonal stands for every real diagonal congruent to it mod 65536 that meets the rectangle
        // (DistanceCalculator.h:98-112); the best score wins, the first of equal ones is kept
        long long best[7] = {0, -1, -1, 0, 0, 0, 0}, cur[7];
        for (unsigned d = 1; d <= 1 + tL / 32768; d++) {
            rescore_one<MODE>(q, qL, t, tL, -(int) (d * 65536) + (int) diagonal, mat, lane, sc_s[warp], eq_s[warp], cur);
            if ((uint32_t) cur[0] > (uint32_t) best[0]) for (int k = 0; k < 7; k++) best[k] = cur[k];
        }
        for (unsigned d = 0; d <= qL / 65536; d++) {
            rescore_one<MODE>(q, qL, t, tL, (int) (d * 65536) + (int) diagonal, mat, lane, sc_s[warp], eq_s[warp], cur);
            if ((uint32_t) cur[0] > (uint32_t) best[0]) for (int k = 0; k < 7; k++) best[k] = cur[k];
        }
        if (lane < 7) out[h * 7 + lane] = (int32_t) best[lane];
    }
}

void adb_free(b200_ctx *ctx) {
    if (ctx->d_ares) cudaFree(ctx->d_ares);
    if (ctx->d_aoff) cudaFree(ctx->d_aoff);
    ctx->d_ares = nullptr; ctx->d_aoff = nullptr; ctx->n_aseq = 0; ctx->h_aoff.clear();
}

}  // namespace

void b200_ascii_db_free(b200_ctx *ctx) { adb_free(ctx); }

extern "C" {

int b200_db_load_ascii(b200_ctx *ctx, const char *data, const uint64_t *offsets, uint64_t n_seq) {
    if (ctx == nullptr) return B200_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (data == nullptr || offsets == nullptr || n_seq == 0) return b200_set_err(ctx, B200_ERR_ARG, "b200_db_load_ascii: bad arguments");
    if (n_seq >= 0xffffffffull) return b200_set_err(ctx, B200_ERR_RANGE, "b200_db_load_ascii: more than 2^32-1 sequences");
    for (uint64_t i = 0; i < n_seq; i++)
        if (offsets[i + 1] < offsets[i]) return b200_set_err(ctx, B200_ERR_ARG, "b200_db_load_ascii: offsets not monotone");
    const uint64_t total = offsets[n_seq] - offsets[0];
    for (uint64_t i = 0; i < total; i++)
        if ((unsigned char) data[offsets[0] + i] >= 123) return b200_set_err(ctx, B200_ERR_ARG, "b200_db_load_ascii: character outside the 123-entry ASCII matrix");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    adb_free(ctx);
    ctx->h_aoff.resize(n_seq + 1);
    for (uint64_t i = 0; i <= n_seq; i++) ctx->h_aoff[i] = offsets[i] - offsets[0];
    cudaError_t e = cudaMalloc(&ctx->d_ares, total + 16);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_aoff, (n_seq + 1) * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_ares, data + offsets[0], total, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_aoff, ctx->h_aoff.data(), (n_seq + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        ctx->err = std::string("b200_db_load_ascii: ") + cudaGetErrorString(e);
        adb_free(ctx);
        return e == cudaErrorMemoryAllocation ? B200_ERR_NOMEM : B200_ERR_CUDA;
    }
    ctx->n_aseq = n_seq;
    return B200_OK;
}

int b200_rescore_diagonal(b200_ctx *ctx, const char *query_data, const uint64_t *query_offsets, int nq, const uint64_t *hit_offsets,
                          const uint32_t *ids, const uint16_t *diagonals, const int8_t *ascii_matrix, int mode, b200_rescore *out) {
    if (ctx == nullptr) return B200_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->n_aseq == 0) return b200_set_err(ctx, B200_ERR_NODB, "no ASCII target DB loaded (b200_db_load_ascii)");
    if (query_data == nullptr || query_offsets == nullptr || nq <= 0 || hit_offsets == nullptr || ascii_matrix == nullptr || mode < 0 || mode > 4)
        return b200_set_err(ctx, B200_ERR_ARG, "b200_rescore_diagonal: bad arguments");
    const uint64_t n = hit_offsets[nq] - hit_offsets[0];
    if (n == 0) return B200_OK;
    if (ids == nullptr || diagonals == nullptr || out == nullptr) return b200_set_err(ctx, B200_ERR_ARG, "b200_rescore_diagonal: bad arguments");
    std::vector<RescoreQuery> h_rq(nq);
    for (int i = 0; i < nq; i++) {
        if (query_offsets[i + 1] < query_offsets[i] || hit_offsets[i + 1] < hit_offsets[i]) return b200_set_err(ctx, B200_ERR_ARG, "b200_rescore_diagonal: offsets not monotone");
        if (query_offsets[i + 1] - query_offsets[i] > 0x7fffffffull) return b200_set_err(ctx, B200_ERR_RANGE, "b200_rescore_diagonal: query too long");
        h_rq[i].seq_off = query_offsets[i] - query_offsets[0]; h_rq[i].hit_begin = hit_offsets[i] - hit_offsets[0];
        h_rq[i].qlen = (int32_t) (query_offsets[i + 1] - query_offsets[i]); h_rq[i].pad_ = 0;
    }
    const uint64_t qbytes = query_offsets[nq] - query_offsets[0];
    for (uint64_t i = 0; i < qbytes; i++)
        if ((unsigned char) query_data[query_offsets[0] + i] >= 123) return b200_set_err(ctx, B200_ERR_ARG, "b200_rescore_diagonal: character outside the 123-entry ASCII matrix");
    const uint32_t *idp = ids + hit_offsets[0];
    const uint16_t *dgp = diagonals + hit_offsets[0];
    for (uint64_t i = 0; i < n; i++)
        if (idp[i] >= ctx->n_aseq) return b200_set_err(ctx, B200_ERR_ARG, "b200_rescore_diagonal: target id out of range");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf d_q, d_rq, d_mat, d_out;
    cudaError_t e = d_q.reserve(qbytes + 16);
    if (e == cudaSuccess) e = d_rq.reserve(sizeof(RescoreQuery) * nq);
    if (e == cudaSuccess) e = d_mat.reserve(123 * 123);
    if (e == cudaSuccess) e = d_out.reserve(n * 7 * sizeof(int32_t));
    if (e == cudaSuccess) e = ctx->ids.reserve(n * sizeof(uint32_t));
    if (e == cudaSuccess) e = ctx->diags.reserve(n * sizeof(uint16_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_q.p, query_data + query_offsets[0], qbytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rq.p, h_rq.data(), sizeof(RescoreQuery) * nq, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_mat.p, ascii_matrix, 123 * 123, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->ids.p, idp, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->diags.p, dgp, n * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream);
    const unsigned ctas = (unsigned) std::min<uint64_t>((n + RS_WARPS - 1) / RS_WARPS, (uint64_t) ctx->sm_count * 8);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[12], ctx->stream);
    if (e == cudaSuccess) {
#define RS_LAUNCH(M) rescore_diagonal_kernel<M><<<ctas, RS_WARPS * 32, 0, ctx->stream>>>(d_q.as<uint8_t>(), d_rq.as<RescoreQuery>(), nq, ctx->d_ares, \
        ctx->d_aoff, ctx->ids.as<uint32_t>(), ctx->diags.as<uint16_t>(), n, d_mat.as<int8_t>(), d_out.as<int32_t>())
        switch (mode) {
            case 0: RS_LAUNCH(0); break;
            case 1: RS_LAUNCH(1); break;
            case 2: RS_LAUNCH(2); break;
            case 3: RS_LAUNCH(3); break;
            default: RS_LAUNCH(4); break;
        }
#undef RS_LAUNCH
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[13], ctx->stream);
    static_assert(sizeof(b200_rescore) == 7 * sizeof(int32_t), "b200_rescore is seven int32");
    if (e == cudaSuccess) e = cudaMemcpyAsync(out + hit_offsets[0], d_out.p, n * 7 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev[12], ctx->ev[13]);
    d_q.release(); d_rq.release(); d_mat.release(); d_out.release();
    if (e != cudaSuccess) { ctx->err = std::string("b200_rescore_diagonal: ") + cudaGetErrorString(e); return e == cudaErrorMemoryAllocation ? B200_ERR_NOMEM : B200_ERR_CUDA; }
    return B200_OK;
}

}  // extern "C"
