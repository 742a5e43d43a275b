// mmseqs2_b200/csrc/b200_backtrace.cu -- A6: CIGAR of protein alignments whose score and end points are already known.
//
// Reference semantics (restated and pinned in oracle/oracle.c, orc_sw_backtrace):
//   SmithWaterman::banded_sw          src/alignment/StripedSmithWaterman.cpp:1478-1693
//   SmithWaterman::computerBacktrace  src/alignment/StripedSmithWaterman.cpp:1280-1308
// banded_sw is a scalar int32 DP over the sub-rectangle [qStart..qEnd] x [dbStart..dbEnd]: band |dbLen-qLen|+1, doubled
// until the banded maximum reaches the known score, three direction bytes per cell, trace back from the bottom-right
// corner.  Rows and the F chain are sequential, so the parallelism is across alignments: one thread per alignment.
// Only hits that survived the E-value / coverage gates get here (a few hundred per query), so this kernel is about
// completeness of alignment mode 3 on the device, not about throughput.
//
// Two passes over the same code: pass 1 finds the final band of every alignment (no direction bytes), the host sizes the
// direction buffers exactly, pass 2 replays every band iteration into the alignment's buffer (earlier iterations leave
// their bytes behind exactly as the reference's realloc'ed buffer does) and walks back.
#include "b200_internal.h"

#include <algorithm>
#include <cstring>

namespace {

struct BtTask {
    uint32_t target;
    int32_t prof_off_lo, prof_off_hi;  // byte offset of the query's [A][qlen] profile in d_prof (64-bit split)
    int32_t qlen;
    int32_t qstart, qend, dbstart, dbend, score;
    uint32_t seq_off_lo, seq_off_hi;   // query residues
    uint32_t pad_;
};

__device__ __forceinline__ int band_u(int band, int i, int j) { int x = i - band; x = x > 0 ? x : 0; return j - x + 1; }
__device__ __forceinline__ long long band_d(int band, int i, int j, int p) { int x = i - band; x = x > 0 ? x : 0; return (long long) (j - x) * 3 + p; }

// PASS 1: out_band[task] = final band.  PASS 2: direction bytes + trace back -> cigar ops, identities.
template <int PASS>
__global__ void __launch_bounds__(128)
sw_backtrace_kernel(const BtTask *__restrict__ tasks, unsigned n_tasks, const int8_t *__restrict__ prof, const uint8_t *__restrict__ qseq,
                    const uint8_t *__restrict__ db, const uint64_t *__restrict__ off, int go, int ge, int32_t *__restrict__ rows,
                    size_t rows_stride, unsigned *__restrict__ counter, int32_t *__restrict__ out_band, int8_t *__restrict__ dirs,
                    const uint64_t *__restrict__ dir_off, uint32_t *__restrict__ cigars, const uint64_t *__restrict__ cigar_off,
                    int32_t *__restrict__ out /* [n][4]: n_ops, identical, bt_len, ok */) {
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    int32_t *hb = rows + tid * rows_stride * 3, *eb = hb + rows_stride, *hc = eb + rows_stride;
    while (true) {
        const unsigned ti = atomicAdd(counter, 1u);
        if (ti >= n_tasks) break;
        const BtTask tk = tasks[ti];
        const int8_t *pr = prof + (((uint64_t) (uint32_t) tk.prof_off_hi << 32) | (uint32_t) tk.prof_off_lo);
        const uint8_t *q = qseq + (((uint64_t) tk.seq_off_hi << 32) | tk.seq_off_lo);
        const uint8_t *t = db + off[tk.target];
        const int q_len = tk.qend - tk.qstart + 1, db_len = tk.dbend - tk.dbstart + 1;
        int band = abs(db_len - q_len) + 1;
        const int final_band = PASS == 2 ? out_band[ti] : 0;
        int8_t *direction = PASS == 2 ? dirs + dir_off[ti] : nullptr;
        int8_t *dl = direction;
        long long width = 0, width_d = 0;
        int maxv = 0;
        // rows persist across band doublings (the reference reallocs, contents kept); start from zero like a fresh object
        {
            const long long w_final = PASS == 2 ? (long long) final_band * 2 + 3 : (long long) rows_stride;
            for (long long k = 0; k < w_final && k < (long long) rows_stride; k++) { hb[k] = 0; eb[k] = 0; hc[k] = 0; }
        }
        do {
            width = (long long) band * 2 + 3; width_d = (long long) band * 2 + 1;
            for (long long j = 1; j < width - 1; j++) hb[j] = 0;
            for (int i = 0; i < q_len; i++) {
                int beg = i - band; if (beg < 0) beg = 0;
                int end = i + band; if (end > db_len - 1) end = db_len - 1;
                const long long edge = end + 1 < width - 1 ? end + 1 : width - 1;
                int f = 0, u = 0;
                hb[0] = 0; eb[0] = 0; hb[edge] = 0; eb[edge] = 0; hc[0] = 0;
                if (PASS == 2) dl = direction + width_d * i * 3;
                const int qi = tk.qstart + i;
                for (int j = beg; j <= end; j++) {
                    u = band_u(band, i, j);
                    const int e_ = band_u(band, i - 1, j), b = band_u(band, i, j - 1), d = band_u(band, i - 1, j - 1);
                    int t1 = (i == 0) ? -go : hb[e_] - go;
                    int t2 = (i == 0) ? -ge : eb[e_] - ge;
                    const int ev = t1 > t2 ? t1 : t2;
                    eb[u] = ev;
                    const int8_t dde = t1 > t2 ? 3 : 2;
                    t1 = hc[b] - go; t2 = f - ge;
                    f = t1 > t2 ? t1 : t2;
                    const int8_t ddf = t1 > t2 ? 5 : 4;
                    const int f1 = f > 0 ? f : 0, e1 = ev > 0 ? ev : 0;
                    t1 = e1 > f1 ? e1 : f1;
                    t2 = hb[d] + (int) pr[(size_t) t[tk.dbstart + j] * tk.qlen + qi];
                    const int h = t1 > t2 ? t1 : t2;
                    hc[u] = h;
                    if (h > maxv) maxv = h;
                    if (PASS == 2) {
                        const long long de = band_d(band, i, j, 0);
                        dl[de] = dde; dl[de + 1] = ddf;
                        dl[de + 2] = (t1 <= t2) ? (int8_t) 1 : (e1 > f1 ? dde : ddf);
                    }
                }
                for (int j = 1; j <= u; j++) hb[j] = hc[j];
            }
            band *= 2;
        } while (maxv < tk.score && band <= (1 << 28));
        band /= 2;
        if (PASS == 1) { out_band[ti] = band; continue; }
        // ---- trace back (bottom-right corner to the origin), ops emitted end -> start, then reversed
        uint32_t *c = cigars + cigar_off[ti];
        int i = q_len - 1, j = db_len - 1, e = 0, n = 0, state = 2;
        bool ok = true;
        char op = 'M', prev_op = 'M';
        while (i > 0 || j > 0) {
            const long long idx = band_d(band, i, j, state);
            switch (dl[idx]) {
                case 1: --i; --j; state = 2; dl -= width_d * 3; op = 'M'; break;
                case 2: --i; state = 0; dl -= width_d * 3; op = 'I'; break;
                case 3: --i; state = 2; dl -= width_d * 3; op = 'I'; break;
                case 4: --j; state = 1; op = 'D'; break;
                case 5: --j; state = 2; op = 'D'; break;
                default: ok = false; break;
            }
            if (!ok) break;
            if (op == prev_op) ++e;
            else { c[n++] = (uint32_t) e << 4 | (prev_op == 'M' ? 0u : prev_op == 'I' ? 1u : 2u); prev_op = op; e = 1; }
        }
        int ids = 0, bt_len = 0;
        if (ok) {
            if (op == 'M') c[n++] = (uint32_t) (e + 1) << 4;
            else { c[n++] = (uint32_t) e << 4 | (op == 'I' ? 1u : 2u); c[n++] = 1u << 4; }
            for (int k = 0; k < n >> 1; k++) { const uint32_t x = c[k]; c[k] = c[n - 1 - k]; c[n - 1 - k] = x; }
            int tp = tk.dbstart, qp = tk.qstart;
            for (int k = 0; k < n; k++) {
                const int L = (int) (c[k] >> 4), o = (int) (c[k] & 0xfu);
                if (o == 0) { for (int r = 0; r < L; r++) ids += t[tp + r] == q[qp + r]; tp += L; qp += L; }
                else if (o == 1) qp += L;
                else tp += L;
                bt_len += L;
            }
        } else n = 0;
        out[(size_t) ti * 4 + 0] = n; out[(size_t) ti * 4 + 1] = ids; out[(size_t) ti * 4 + 2] = bt_len; out[(size_t) ti * 4 + 3] = ok ? 1 : 0;
    }
}

}  // namespace

int b200_sw_backtrace(b200_ctx *ctx, const b200_query *queries, const uint8_t *const *query_seqs, int nq, const b200_pair *pairs,
                      uint64_t n, int gap_open, int gap_extend, const b200_sw_aln *alns, b200_sw_bt *out, uint32_t *cigars,
                      const uint64_t *cigar_offsets) {
    if (ctx == nullptr) return B200_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->n_seq == 0) return b200_set_err(ctx, B200_ERR_NODB, "no target DB loaded");
    if (n == 0) return B200_OK;
    if (queries == nullptr || query_seqs == nullptr || pairs == nullptr || alns == nullptr || out == nullptr || cigars == nullptr ||
        cigar_offsets == nullptr || nq <= 0)
        return b200_set_err(ctx, B200_ERR_ARG, "b200_sw_backtrace: NULL argument");
    if (n >= 0xffffffffull) return b200_set_err(ctx, B200_ERR_RANGE, "b200_sw_backtrace: too many alignments");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const int A = ctx->alphabet;
    // stage profiles + query residues
    std::vector<uint64_t> poff(nq), soff(nq);
    uint64_t pbytes = 0, sbytes = 0;
    for (int i = 0; i < nq; i++) {
        if (queries[i].profile == nullptr || queries[i].qlen <= 0 || query_seqs[i] == nullptr) return b200_set_err(ctx, B200_ERR_ARG, "b200_sw_backtrace: query without profile or residues");
        poff[i] = pbytes; pbytes += (uint64_t) A * queries[i].qlen;
        soff[i] = sbytes; sbytes += (uint64_t) queries[i].qlen;
    }
    std::vector<int8_t> h_prof(pbytes);
    std::vector<uint8_t> h_seq(sbytes);
    for (int i = 0; i < nq; i++) {
        memcpy(h_prof.data() + poff[i], queries[i].profile, (size_t) A * queries[i].qlen);
        memcpy(h_seq.data() + soff[i], query_seqs[i], (size_t) queries[i].qlen);
    }
    std::vector<BtTask> h_tasks;
    std::vector<uint64_t> idx;  // caller index of every task
    int max_span = 1;
    for (uint64_t i = 0; i < n; i++) {
        out[i].n_cigar = 0; out[i].identical = 0; out[i].bt_len = 0; out[i].ok = 0;
        if (pairs[i].query >= (uint32_t) nq || pairs[i].target >= ctx->n_seq) return b200_set_err(ctx, B200_ERR_ARG, "b200_sw_backtrace: pair index out of range");
        const b200_sw_aln &a = alns[i];
        if (a.dbend < 0 || a.qstart < 0 || a.dbstart < 0) continue;   // gated out upstream: nothing to trace
        const int ql = queries[pairs[i].query].qlen, tl = ctx->h_len[pairs[i].target];
        if (a.qend >= ql || a.dbend >= tl || a.qstart > a.qend || a.dbstart > a.dbend || a.score <= 0)
            return b200_set_err(ctx, B200_ERR_ARG, "b200_sw_backtrace: inconsistent alignment coordinates");
        const uint64_t need = (uint64_t) (a.qend - a.qstart + 1) + (uint64_t) (a.dbend - a.dbstart + 1) + 2;
        if (cigar_offsets[i + 1] - cigar_offsets[i] < need) return b200_set_err(ctx, B200_ERR_ARG, "b200_sw_backtrace: cigar slot smaller than qAlnLen + dbAlnLen + 2");
        BtTask t;
        t.target = pairs[i].target;
        t.prof_off_lo = (int32_t) (uint32_t) (poff[pairs[i].query] & 0xffffffffu); t.prof_off_hi = (int32_t) (uint32_t) (poff[pairs[i].query] >> 32);
        t.seq_off_lo = (uint32_t) (soff[pairs[i].query] & 0xffffffffu); t.seq_off_hi = (uint32_t) (soff[pairs[i].query] >> 32);
        t.qlen = ql; t.qstart = a.qstart; t.qend = a.qend; t.dbstart = a.dbstart; t.dbend = a.dbend; t.score = a.score; t.pad_ = 0;
        h_tasks.push_back(t);
        idx.push_back(i);
        max_span = std::max(max_span, std::max(a.qend - a.qstart + 1, a.dbend - a.dbstart + 1));
    }
    const size_t m = h_tasks.size();
    if (m == 0) return B200_OK;
    // the band stops doubling once it covers the rectangle: final band < 2 * max_span, rows need 2*band+3 ints
    const size_t rows_stride = (size_t) 4 * max_span + 16;
    size_t threads = std::min<size_t>(round_up(m, 128), (size_t) ctx->sm_count * 1024);
    while (threads > 128 && threads * rows_stride * 3 * sizeof(int32_t) > ((size_t) 4 << 30)) threads /= 2;
    threads = round_up(threads, 128);
    const unsigned grid = (unsigned) (threads / 128);
    DevBuf d_tasks, d_prof, d_seq, d_rows, d_band, d_dirs, d_diroff, d_cig, d_cigoff, d_out;
    std::vector<uint64_t> h_cigoff(m + 1, 0);
    for (size_t k = 0; k < m; k++) h_cigoff[k + 1] = h_cigoff[k] + (cigar_offsets[idx[k] + 1] - cigar_offsets[idx[k]]);
    cudaError_t e = d_tasks.reserve(sizeof(BtTask) * m);
    if (e == cudaSuccess) e = d_prof.reserve(pbytes + 16);
    if (e == cudaSuccess) e = d_seq.reserve(sbytes + 16);
    if (e == cudaSuccess) e = d_rows.reserve(threads * rows_stride * 3 * sizeof(int32_t));
    if (e == cudaSuccess) e = d_band.reserve(sizeof(int32_t) * m);
    if (e == cudaSuccess) e = d_cig.reserve(sizeof(uint32_t) * h_cigoff[m] + 16);
    if (e == cudaSuccess) e = d_cigoff.reserve(sizeof(uint64_t) * (m + 1));
    if (e == cudaSuccess) e = d_diroff.reserve(sizeof(uint64_t) * (m + 1));
    if (e == cudaSuccess) e = d_out.reserve(sizeof(int32_t) * 4 * m);
    if (e == cudaSuccess) e = ctx->counter.reserve(sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tasks.p, h_tasks.data(), sizeof(BtTask) * m, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_prof.p, h_prof.data(), pbytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_seq.p, h_seq.data(), sbytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cigoff.p, h_cigoff.data(), sizeof(uint64_t) * (m + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->counter.p, 0, sizeof(unsigned), ctx->stream);
    if (e == cudaSuccess) {
        sw_backtrace_kernel<1><<<grid, 128, 0, ctx->stream>>>(d_tasks.as<BtTask>(), (unsigned) m, d_prof.as<int8_t>(), d_seq.as<uint8_t>(), ctx->d_res,
                                                             ctx->d_off, gap_open, gap_extend, d_rows.as<int32_t>(), rows_stride,
                                                             ctx->counter.as<unsigned>(), d_band.as<int32_t>(), nullptr, nullptr, nullptr,
                                                             nullptr, nullptr);
        ctx->launches++;
        e = cudaGetLastError();
    }
    std::vector<int32_t> h_band(m);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_band.data(), d_band.p, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    std::vector<uint64_t> h_diroff(m + 1, 0);
    if (e == cudaSuccess) {
        for (size_t k = 0; k < m; k++) {
            const uint64_t q_len = (uint64_t) (h_tasks[k].qend - h_tasks[k].qstart + 1);
            h_diroff[k + 1] = h_diroff[k] + round_up(((uint64_t) h_band[k] * 2 + 1) * q_len * 3 + 16, 16);
        }
        e = d_dirs.reserve(h_diroff[m] + 16);
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(d_dirs.p, 0, h_diroff[m] + 16, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_diroff.p, h_diroff.data(), sizeof(uint64_t) * (m + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->counter.p, 0, sizeof(unsigned), ctx->stream);
    if (e == cudaSuccess) {
        sw_backtrace_kernel<2><<<grid, 128, 0, ctx->stream>>>(d_tasks.as<BtTask>(), (unsigned) m, d_prof.as<int8_t>(), d_seq.as<uint8_t>(), ctx->d_res,
                                                             ctx->d_off, gap_open, gap_extend, d_rows.as<int32_t>(), rows_stride,
                                                             ctx->counter.as<unsigned>(), d_band.as<int32_t>(), d_dirs.as<int8_t>(),
                                                             d_diroff.as<uint64_t>(), d_cig.as<uint32_t>(), d_cigoff.as<uint64_t>(),
                                                             d_out.as<int32_t>());
        ctx->launches++;
        e = cudaGetLastError();
    }
    std::vector<int32_t> h_out(4 * m);
    std::vector<uint32_t> h_cig(h_cigoff[m] + 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out.data(), d_out.p, sizeof(int32_t) * 4 * m, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_cig.data(), d_cig.p, sizeof(uint32_t) * h_cigoff[m], cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&d_tasks, &d_prof, &d_seq, &d_rows, &d_band, &d_dirs, &d_diroff, &d_cig, &d_cigoff, &d_out};
    for (DevBuf *b : bufs) b->release();
    if (e != cudaSuccess) { ctx->err = std::string("b200_sw_backtrace: ") + cudaGetErrorString(e); return e == cudaErrorMemoryAllocation ? B200_ERR_NOMEM : B200_ERR_CUDA; }
    for (size_t k = 0; k < m; k++) {
        const uint64_t i = idx[k];
        out[i].n_cigar = h_out[4 * k]; out[i].identical = h_out[4 * k + 1]; out[i].bt_len = h_out[4 * k + 2]; out[i].ok = h_out[4 * k + 3];
        memcpy(cigars + cigar_offsets[i], h_cig.data() + h_cigoff[k], sizeof(uint32_t) * (size_t) std::max(0, h_out[4 * k]));
    }
    return B200_OK;
}
