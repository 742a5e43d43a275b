// mmseqs2_b200/csrc/b200_backtrace.cu -- A6: CIGAR of protein alignments whose score and end points are already known.
//
// Reference semantics (restated and pinned in oracle/oracle.c, orc_sw_backtrace):
//   SmithWaterman::banded_sw          src/alignment/StripedSmithWaterman.cpp:1478-1693
//   SmithWaterman::computerBacktrace  src/alignment/StripedSmithWaterman.cpp:1280-1308
// banded_sw is a scalar int32 DP over the sub-rectangle [qStart..qEnd] x [dbStart..dbEnd]: band |dbLen-qLen|+1, doubled
// until the banded maximum reaches the known score, three direction bytes per cell, trace back from the bottom-right
// corner.  One warp per alignment: rows are sequential, the cells of a row go 32 at a time (the F chain along the row is
// a warp prefix maximum), lane 0 walks the direction bytes back.
//
// Two passes over the same code: pass 1 finds the final band of every alignment (no direction bytes), the host sizes the
// direction buffers exactly, pass 2 replays every band iteration into the alignment's buffer (earlier iterations leave
// their bytes behind exactly as the reference's realloc'ed buffer does) and walks back.
#include "b200_internal.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
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

constexpr int BT_WARPS = 8;          // warps per CTA, one alignment per warp at a time
constexpr int BT_SMEM_W = 259;       // shared-memory row entries per warp: 2 * band + 3 for band <= 128

// One warp per alignment.  banded_sw keeps three int rows indexed by band position (h of the previous row, e, h of the
// current row) and walks the cells of a row left to right; the only left-to-right dependency is the F chain
//     f(j) = max(h(j-1) - go, f(j-1) - ge),   h(j) = max(max(e1, max(f,0)), diag + s)        (e1 = max(e,0) >= 0)
// With g = max(e1, diag + s) >= 0 this is h = max(g, f) and f(j) = max(g(j-1) - go, f(j-1) - min(go, ge)), i.e.
// f(j) + j*m is a running maximum of g(k) - go + (k+1)*m  -- a warp prefix-max per 32 cells plus a carry.  Every value,
// and therefore every direction byte (the reference's strict/non-strict comparisons are evaluated on the same numbers),
// equals the sequential evaluation.  The rows, their persistence across band doublings and the edge resets are kept as
// in the reference (hb[0], eb[0], hb[edge], eb[edge], hc[0] = 0 at the start of a row; hb[1..u] = hc[1..u] at its end).
// PASS 1: out_band[task] = final band.  PASS 2: direction bytes + trace back -> cigar ops, identities.
template <int PASS>
__global__ void __launch_bounds__(BT_WARPS * 32)
sw_backtrace_kernel(const BtTask *__restrict__ tasks, unsigned n_tasks, const int8_t *__restrict__ prof, const uint8_t *__restrict__ qseq,
                    const uint8_t *__restrict__ db, const uint64_t *__restrict__ off, int go, int ge, int32_t *__restrict__ rows,
                    size_t rows_stride, unsigned *__restrict__ counter, int32_t *__restrict__ out_band, int8_t *__restrict__ dirs,
                    const uint64_t *__restrict__ dir_off, uint32_t *__restrict__ cigars, const uint64_t *__restrict__ cigar_off,
                    int32_t *__restrict__ out /* [n][4]: n_ops, identical, bt_len, ok */, uint32_t *__restrict__ pool,
                    unsigned long long *__restrict__ pool_used, unsigned long long *__restrict__ pool_base) {
    __shared__ int32_t srows[BT_WARPS][3][BT_SMEM_W];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t gwarp = (size_t) blockIdx.x * BT_WARPS + warp;
    int32_t *ghb = rows + gwarp * rows_stride * 3, *geb = ghb + rows_stride, *ghc = geb + rows_stride;
    const int m = go < ge ? go : ge;
    while (true) {
        unsigned ti = 0;
        if (lane == 0) ti = atomicAdd(counter, 1u);
        ti = __shfl_sync(0xffffffffu, ti, 0);
        if (ti >= n_tasks) break;
        const BtTask tk = tasks[ti];
        const int8_t *pr = prof + (((uint64_t) (uint32_t) tk.prof_off_hi << 32) | (uint32_t) tk.prof_off_lo);
        const uint8_t *q = qseq + (((uint64_t) tk.seq_off_hi << 32) | tk.seq_off_lo);
        const uint8_t *t = db + off[tk.target];
        const int q_len = tk.qend - tk.qstart + 1, db_len = tk.dbend - tk.dbstart + 1;
        int band = abs(db_len - q_len) + 1;
        int8_t *direction = PASS == 2 ? dirs + dir_off[ti] : nullptr;
        int8_t *dl = direction;
        long long width = 0, width_d = 0, zeroed = 0;
        int maxv = 0;
        bool covered = false;
        int32_t *hb = srows[warp][0], *eb = srows[warp][1], *hc = srows[warp][2];
        bool in_smem = true;
        do {
            width = (long long) band * 2 + 3; width_d = (long long) band * 2 + 1;
            if (in_smem && width > BT_SMEM_W) {          // wider than the shared rows: move to this warp's global rows
                for (long long k = lane; k < zeroed; k += 32) { ghb[k] = hb[k]; geb[k] = eb[k]; ghc[k] = hc[k]; }
                hb = ghb; eb = geb; hc = ghc; in_smem = false;
            }
            // rows start from zero like a fresh object; entries beyond the previous width have never been touched
            for (long long k = zeroed + lane; k < width; k += 32) { hb[k] = 0; eb[k] = 0; hc[k] = 0; }
            zeroed = width;
            __syncwarp();
            for (long long k = 1 + lane; k < width - 1; k += 32) hb[k] = 0;
            __syncwarp();
            int lmax = 0;
            for (int i = 0; i < q_len; i++) {
                int beg = i - band; if (beg < 0) beg = 0;
                int end = i + band; if (end > db_len - 1) end = db_len - 1;
                const long long edge = end + 1 < width - 1 ? end + 1 : width - 1;
                if (lane == 0) { hb[0] = 0; eb[0] = 0; hb[edge] = 0; eb[edge] = 0; hc[0] = 0; }
                __syncwarp();
                if (PASS == 2) dl = direction + width_d * i * 3;
                const int8_t *prow = pr + (tk.qstart + i);
                const int x = (i - band) > 0 ? (i - band) : 0, xp = (i - 1 - band) > 0 ? (i - 1 - band) : 0;
                // carry of the F chain into the next cell: h and f of the cell to the left (hc[0] = 0 and f = 0 before the first)
                int h_left = 0, f_left = 0;
                for (int j0 = beg; j0 <= end; j0 += 32) {
                    const int j = j0 + lane;
                    const bool on = j <= end;
                    const int u = j - x + 1, e_ = j - xp + 1;
                    int ev = 0, g = 0, diag = 0, e1 = 0;
                    int8_t dde = 2;
                    if (on) {
                        const int t1 = (i == 0) ? -go : hb[e_] - go;
                        const int t2 = (i == 0) ? -ge : eb[e_] - ge;
                        ev = t1 > t2 ? t1 : t2;
                        dde = t1 > t2 ? 3 : 2;
                        e1 = ev > 0 ? ev : 0;
                        diag = hb[e_ - 1] + (int) prow[(size_t) t[tk.dbstart + j] * tk.qlen];
                        g = e1 > diag ? e1 : diag;
                    }
                    __syncwarp();                        // every read of eb[] is done before the row's new e values land
                    if (on) eb[u] = ev;
                    // f of lane l = max(F0 - l*m, max over k < l of g(k) - go - (l-1-k)*m): inclusive prefix max of
                    // w(k) = g(k) - go + (k+1)*m, shifted by one lane, minus l*m; F0 = f of the chunk's first cell, exact from the carry
                    int w = on ? g - go + (lane + 1) * m : INT_MIN / 2;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, w, o);
                        if (lane >= o) w = w > v ? w : v;
                    }
                    const int f0 = (h_left - go) > (f_left - ge) ? (h_left - go) : (f_left - ge);
                    int pre = __shfl_up_sync(0xffffffffu, w, 1);
                    if (lane == 0) pre = INT_MIN / 2;
                    int f = f0 - lane * m;
                    { const int alt = pre - lane * m; f = f > alt ? f : alt; }
                    const int h = g > f ? g : f;
                    // left neighbour's (h, f) for the direction byte of the F move
                    int hl = __shfl_up_sync(0xffffffffu, h, 1), fl = __shfl_up_sync(0xffffffffu, f, 1);
                    if (lane == 0) { hl = h_left; fl = f_left; }
                    if (on) {
                        hc[u] = h;
                        lmax = h > lmax ? h : lmax;
                        if (PASS == 2) {
                            const int t1 = hl - go, t2 = fl - ge;
                            const int8_t ddf = t1 > t2 ? 5 : 4;
                            const int f1 = f > 0 ? f : 0;
                            const int tt1 = e1 > f1 ? e1 : f1;
                            const long long de = (long long) (j - x) * 3;
                            dl[de] = dde; dl[de + 1] = ddf;
                            dl[de + 2] = (tt1 <= diag) ? (int8_t) 1 : (e1 > f1 ? dde : ddf);
                        }
                    }
                    // carry out: the last active cell of this chunk
                    const int last = (end - j0) < 31 ? (end - j0) : 31;
                    h_left = __shfl_sync(0xffffffffu, h, last);
                    f_left = __shfl_sync(0xffffffffu, f, last);
                }
                __syncwarp();
                const int u_last = end - x + 1;
                for (int k = 1 + lane; k <= u_last; k += 32) hb[k] = hc[k];
                __syncwarp();
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const int v = __shfl_xor_sync(0xffffffffu, lmax, o); lmax = lmax > v ? lmax : v; }
            maxv = maxv > lmax ? maxv : lmax;
            // a band that already spans the whole rectangle cannot find more: stop there even if the claimed score was not reached
            // (inconsistent caller input, or a score capped at 32767) -- the rows and the direction buffer are sized for exactly this
            // bound (final band < 2 * max_span); the reference would keep doubling and reallocating (StripedSmithWaterman.cpp:1614-1640)
            covered = band >= (q_len > db_len ? q_len : db_len);
            band *= 2;
        } while (maxv < tk.score && !covered && band <= (1 << 28));
        band /= 2;
        const bool reached = maxv >= tk.score;
        if (PASS == 1) { if (lane == 0) out_band[ti] = band; continue; }
        __syncwarp();                                    // the direction bytes of every lane are visible to lane 0
        if (lane != 0) continue;                         // lane 0 walks back; the other lanes wait at the next task fetch
        // ---- trace back (bottom-right corner to the origin), ops emitted end -> start, then reversed
        uint32_t *c = cigars + cigar_off[ti];
        int i = q_len - 1, j = db_len - 1, e = 0, n = 0, state = 2;
        bool ok = reached;      // score not reachable inside the rectangle: no CIGAR, ok = 0
        char op = 'M', prev_op = 'M';
        while (ok && (i > 0 || j > 0)) {
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
            for (int k = 0; k < n >> 1; k++) { const uint32_t xx = c[k]; c[k] = c[n - 1 - k]; c[n - 1 - k] = xx; }
            int tp = tk.dbstart, qp = tk.qstart;
            for (int k = 0; k < n; k++) {
                const int L = (int) (c[k] >> 4), o = (int) (c[k] & 0xfu);
                if (o == 0) { for (int r = 0; r < L; r++) ids += t[tp + r] == q[qp + r]; tp += L; qp += L; }
                else if (o == 1) qp += L;
                else tp += L;
                bt_len += L;
            }
        } else n = 0;
        // the ops go to a dense pool (the per-task slots are worst-case sized: only the pool travels back to the host)
        const unsigned long long base = atomicAdd(pool_used, (unsigned long long) n);
        for (int k = 0; k < n; k++) pool[base + k] = c[k];
        pool_base[ti] = base;
        out[(size_t) ti * 4 + 0] = n; out[(size_t) ti * 4 + 1] = ids; out[(size_t) ti * 4 + 2] = bt_len; out[(size_t) ti * 4 + 3] = ok ? 1 : 0;
    }
}

}  // namespace

int b200_sw_backtrace(b200_ctx *ctx, const b200_query *queries, const uint8_t *const *query_seqs, int nq, const b200_pair *pairs,
                      uint64_t n, int gap_open, int gap_extend, const b200_sw_aln *alns, b200_sw_bt *out, uint32_t *cigars,
                      const uint64_t *cigar_offsets) {
    if (ctx == nullptr) return B200_ERR_ARG;
    if (n > 0 && (cigars == nullptr || cigar_offsets == nullptr)) return b200_set_err(ctx, B200_ERR_ARG, "b200_sw_backtrace: NULL argument");
    return b200_sw_backtrace_impl(ctx, queries, query_seqs, nq, pairs, n, gap_open, gap_extend, alns, out, cigars, cigar_offsets, nullptr, nullptr);
}

// cigars/cigar_offsets == NULL: the ops stay in one dense pool, *pool_out, with alignment i at (*base_out)[i] (what
// b200_align_batch uses: scattering a few ops into worst-case sized slots only costs page faults)
int b200_sw_backtrace_impl(b200_ctx *ctx, const b200_query *queries, const uint8_t *const *query_seqs, int nq, const b200_pair *pairs,
                           uint64_t n, int gap_open, int gap_extend, const b200_sw_aln *alns, b200_sw_bt *out, uint32_t *cigars,
                           const uint64_t *cigar_offsets, std::vector<uint32_t> *pool_out, std::vector<uint64_t> *base_out) {
    if (ctx == nullptr) return B200_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->n_seq == 0) return b200_set_err(ctx, B200_ERR_NODB, "no target DB loaded");
    if (base_out != nullptr) base_out->assign(n, 0);
    if (pool_out != nullptr) pool_out->clear();
    if (n == 0) return B200_OK;
    if (queries == nullptr || query_seqs == nullptr || pairs == nullptr || alns == nullptr || out == nullptr || nq <= 0 ||
        (cigar_offsets == nullptr && (pool_out == nullptr || base_out == nullptr)))
        return b200_set_err(ctx, B200_ERR_ARG, "b200_sw_backtrace: NULL argument");
    if (n >= 0xffffffffull) return b200_set_err(ctx, B200_ERR_RANGE, "b200_sw_backtrace: too many alignments");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const int A = ctx->alphabet;
    typedef std::chrono::steady_clock Clock;
    const Clock::time_point tc0 = Clock::now();
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
        if (cigar_offsets != nullptr && cigar_offsets[i + 1] - cigar_offsets[i] < need)
            return b200_set_err(ctx, B200_ERR_ARG, "b200_sw_backtrace: cigar slot smaller than qAlnLen + dbAlnLen + 2");
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
    {   // longest alignments first: threads pull tasks from a counter, so the long ones must not start last.
        // Stable counting sort on the query span (<= 65535): linear, where a comparison sort of 5e5 tasks cost ~60 ms per call.
        std::vector<uint32_t> cnt(65537, 0);
        for (size_t k = 0; k < m; k++) cnt[(size_t) std::min(65535, h_tasks[k].qend - h_tasks[k].qstart)]++;
        uint32_t run = 0;
        for (int v = 65535; v >= 0; v--) { const uint32_t c = cnt[v]; cnt[v] = run; run += c; }   // descending spans
        std::vector<BtTask> st(m);
        std::vector<uint64_t> si(m);
        for (size_t k = 0; k < m; k++) {
            const uint32_t pos = cnt[(size_t) std::min(65535, h_tasks[k].qend - h_tasks[k].qstart)]++;
            st[pos] = h_tasks[k]; si[pos] = idx[k];
        }
        h_tasks.swap(st); idx.swap(si);
    }

    // the band stops doubling once it covers the rectangle: final band < 2 * max_span, rows need 2*band+3 ints
    const size_t rows_stride = (size_t) 4 * max_span + 16;
    // resident warps only: every warp pulls alignments from the counter until none are left
    size_t warps = std::min<size_t>(round_up(m, BT_WARPS), (size_t) ctx->sm_count * 4 * BT_WARPS);
    while (warps > BT_WARPS && warps * rows_stride * 3 * sizeof(int32_t) > ((size_t) 4 << 30)) warps /= 2;
    warps = round_up(warps, BT_WARPS);
    const unsigned grid = (unsigned) (warps / BT_WARPS);
    const size_t threads = warps;   // scratch rows are per warp
    DevBuf d_tasks, d_prof, d_seq, d_rows, d_band, d_dirs, d_diroff, d_cig, d_cigoff, d_out, d_pool, d_pbase;
    std::vector<uint64_t> h_cigoff(m + 1, 0);
    for (size_t k = 0; k < m; k++) {
        const BtTask &bk = h_tasks[k];
        const uint64_t need = (uint64_t) (bk.qend - bk.qstart + 1) + (uint64_t) (bk.dbend - bk.dbstart + 1) + 2;
        h_cigoff[k + 1] = h_cigoff[k] + (cigar_offsets != nullptr ? cigar_offsets[idx[k] + 1] - cigar_offsets[idx[k]] : need);
    }
    cudaError_t e = d_tasks.reserve(sizeof(BtTask) * m);
    if (e == cudaSuccess) e = d_prof.reserve(pbytes + 16);
    if (e == cudaSuccess) e = d_seq.reserve(sbytes + 16);
    if (e == cudaSuccess) e = d_rows.reserve(threads * rows_stride * 3 * sizeof(int32_t));
    if (e == cudaSuccess) e = d_band.reserve(sizeof(int32_t) * m);
    if (e == cudaSuccess) e = d_cig.reserve(sizeof(uint32_t) * h_cigoff[m] + 16);
    if (e == cudaSuccess) e = d_cigoff.reserve(sizeof(uint64_t) * (m + 1));
    if (e == cudaSuccess) e = d_diroff.reserve(sizeof(uint64_t) * (m + 1));
    if (e == cudaSuccess) e = d_out.reserve(sizeof(int32_t) * 4 * m);
    if (e == cudaSuccess) e = d_pool.reserve(sizeof(uint32_t) * h_cigoff[m] + 16);
    if (e == cudaSuccess) e = d_pbase.reserve(sizeof(unsigned long long) * (m + 1));
    if (e == cudaSuccess) e = ctx->counter.reserve(sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tasks.p, h_tasks.data(), sizeof(BtTask) * m, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_prof.p, h_prof.data(), pbytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_seq.p, h_seq.data(), sbytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cigoff.p, h_cigoff.data(), sizeof(uint64_t) * (m + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->counter.p, 0, sizeof(unsigned), ctx->stream);
    const Clock::time_point tc1 = Clock::now();
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[12], ctx->stream);
    if (e == cudaSuccess) {
        sw_backtrace_kernel<1><<<grid, BT_WARPS * 32, 0, ctx->stream>>>(d_tasks.as<BtTask>(), (unsigned) m, d_prof.as<int8_t>(), d_seq.as<uint8_t>(), ctx->d_res,
                                                             ctx->d_off, gap_open, gap_extend, d_rows.as<int32_t>(), rows_stride,
                                                             ctx->counter.as<unsigned>(), d_band.as<int32_t>(), nullptr, nullptr, nullptr,
                                                             nullptr, nullptr, nullptr, nullptr, nullptr);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[13], ctx->stream);
    std::vector<int32_t> h_band(m);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_band.data(), d_band.p, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    const Clock::time_point tc2 = Clock::now();
    if (e == cudaSuccess && getenv("B200_TRACE") != nullptr) {   // development aid: where the band search ended, and its cost
        uint64_t cells = 0, worst = 0;
        int hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (size_t k = 0; k < m; k++) {
            const uint64_t ql = (uint64_t) (h_tasks[k].qend - h_tasks[k].qstart + 1);
            const int b0 = abs((h_tasks[k].dbend - h_tasks[k].dbstart) - (h_tasks[k].qend - h_tasks[k].qstart)) + 1;
            uint64_t c = 0;
            for (int64_t b = b0; b <= h_band[k]; b *= 2) c += ql * (uint64_t) (2 * b + 1);
            cells += c; worst = std::max(worst, c);
            int bin = 0;
            while (bin < 7 && (8 << bin) < h_band[k]) bin++;
            hist[bin]++;
        }
        fprintf(stderr, "[b200 trace] backtrace: %zu alignments, %.3g cells per pass, largest task %.3g cells; final band <=8:%d <=16:%d <=32:%d "
                        "<=64:%d <=128:%d <=256:%d <=512:%d more:%d\n", m, (double) cells, (double) worst, hist[0], hist[1], hist[2], hist[3],
                hist[4], hist[5], hist[6], hist[7]);
    }
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
    if (e == cudaSuccess) e = cudaMemsetAsync(d_pbase.as<unsigned long long>() + m, 0, sizeof(unsigned long long), ctx->stream);   // pool_used
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[14], ctx->stream);
    if (e == cudaSuccess) {
        sw_backtrace_kernel<2><<<grid, BT_WARPS * 32, 0, ctx->stream>>>(d_tasks.as<BtTask>(), (unsigned) m, d_prof.as<int8_t>(), d_seq.as<uint8_t>(), ctx->d_res,
                                                             ctx->d_off, gap_open, gap_extend, d_rows.as<int32_t>(), rows_stride,
                                                             ctx->counter.as<unsigned>(), d_band.as<int32_t>(), d_dirs.as<int8_t>(),
                                                             d_diroff.as<uint64_t>(), d_cig.as<uint32_t>(), d_cigoff.as<uint64_t>(),
                                                             d_out.as<int32_t>(), d_pool.as<uint32_t>(), d_pbase.as<unsigned long long>() + m,
                                                             d_pbase.as<unsigned long long>());
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[15], ctx->stream);
    const Clock::time_point tc3 = Clock::now();
    std::vector<int32_t> h_out(4 * m);
    std::vector<unsigned long long> h_pbase(m + 1, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out.data(), d_out.p, sizeof(int32_t) * 4 * m, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_pbase.data(), d_pbase.p, sizeof(unsigned long long) * (m + 1), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    std::vector<uint32_t> h_pool((size_t) h_pbase[m] + 1);
    if (e == cudaSuccess && h_pbase[m] > 0) e = cudaMemcpyAsync(h_pool.data(), d_pool.p, sizeof(uint32_t) * (size_t) h_pbase[m], cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) {
        float p1 = 0.f, p2 = 0.f;
        cudaEventElapsedTime(&p1, ctx->ev[12], ctx->ev[13]);
        cudaEventElapsedTime(&p2, ctx->ev[14], ctx->ev[15]);
        ctx->last_kernel_ms = p1 + p2;
        if (getenv("B200_TRACE") != nullptr) {
            auto ms = [](Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "[b200 trace] backtrace kernels: band search %.2f ms, directions + trace back %.2f ms, direction bytes %.3g; host: "
                            "stage %.1f ms, pass 1 + bands back %.1f ms, dir buffer + pass 2 launch %.1f ms, wait + results back %.1f ms\n",
                    p1, p2, (double) h_diroff[m], ms(tc0, tc1), ms(tc1, tc2), ms(tc2, tc3), ms(tc3, Clock::now()));
        }
    }
    DevBuf *bufs[] = {&d_tasks, &d_prof, &d_seq, &d_rows, &d_band, &d_dirs, &d_diroff, &d_cig, &d_cigoff, &d_out, &d_pool, &d_pbase};
    for (DevBuf *b : bufs) b->release();
    if (e != cudaSuccess) { ctx->err = std::string("b200_sw_backtrace: ") + cudaGetErrorString(e); return e == cudaErrorMemoryAllocation ? B200_ERR_NOMEM : B200_ERR_CUDA; }
    for (size_t k = 0; k < m; k++) {
        const uint64_t i = idx[k];
        out[i].n_cigar = h_out[4 * k]; out[i].identical = h_out[4 * k + 1]; out[i].bt_len = h_out[4 * k + 2]; out[i].ok = h_out[4 * k + 3];
        if (cigar_offsets != nullptr) memcpy(cigars + cigar_offsets[i], h_pool.data() + h_pbase[k], sizeof(uint32_t) * (size_t) std::max(0, h_out[4 * k]));
        else (*base_out)[i] = h_pbase[k];
    }
    if (pool_out != nullptr) pool_out->swap(h_pool);
    return B200_OK;
}
