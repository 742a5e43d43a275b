// mmseqs2_b200/csrc/b200_nucl.cu -- A7: the nucleotide gapped aligner on sm_100a, one warp per (query, target, diagonal).
//
// Reference semantics (restated and pinned in oracle/oracle_ksw.c):
//   BandedNucleotideAligner::align      src/alignment/BandedNucleotideAligner.cpp:73-263
//   DistanceCalculator ungapped seed    src/alignment/DistanceCalculator.h:93-200
//   ksw_extz2_sse                       lib/ksw2/ksw2_extz2_sse.cpp:44-285   (band 64, z-drop, exact 32-bit maximum)
//   ksw_backtrack / ksw_apply_zdrop     lib/ksw2/ksw2.h:145-202
//
// ksw2's observable results include artefacts of its 16-lane SSE blocks (band rounded outwards to multiples of 16, stale
// lanes outside the band, 16-byte score stores past the band end, a four-stream running maximum).  To stay bit-exact the
// kernel keeps ksw2's byte arrays (u v x y s | target copy | reversed query) with the same relative layout and lets the 32
// lanes of a warp play the SSE lanes: one anti-diagonal r per iteration, lane = target index t.  The arrays live in a
// per-warp scratch slice (L2-resident: ~3 KB per 150-bp read); the direction matrix for the CIGAR pass goes to a second
// per-warp slice.  Integer only; the E-value / coverage arithmetic stays with the caller.
#include "b200_internal.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int KSW_NEG_INF = -0x40000000;
constexpr int EZ_SCORE_ONLY = 0x01, EZ_EXTZ_ONLY = 0x40;
constexpr int NUCL_WARPS = 8;
constexpr int MAX_CHUNKS = 3;  // band 64: at most 65 cells per anti-diagonal, rounded outwards to 16 -> at most 96 lanes

struct NuclTask { uint32_t query; uint32_t target; uint32_t diagonal; uint32_t cigar_off; };

struct SeqView {  // element i of a ksw operand: forward slice or the reference's shifted reverse (rev[k] = seq[L-k], rev[0] = X)
    const uint8_t *base;
    int L, start, rev;
    __device__ __forceinline__ uint8_t get(int i) const {
        const int idx = start + i;
        if (!rev) return base[idx];
        return idx == 0 ? (uint8_t) 4 : base[L - idx];
    }
};

struct Ez { int max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar; };

// One ksw_extz2 call, executed by a full warp.  mem: zero-initialisable scratch (layout below), H: int32[Ls],
// pmat/poff: direction bytes and per-row (st,en) for the CIGAR pass (unused when score-only).
template <bool SMEM>
__device__ __forceinline__ void ksw_extz2_warp(int qlen, const SeqView &qv, int tlen, const SeqView &tv, int8_t sc_mch, int8_t sc_mis, int q, int e,
                               int w, int zdrop, int flag, uint8_t *gmem, size_t mem_bytes_cap, uint8_t *pmat, int2 *poff, Ez &ez,
                               uint32_t *cigar, int cigar_cap) {
    extern __shared__ __align__(16) uint8_t nucl_smem[];
    // the ksw byte arrays + H: shared memory slice of this warp when they fit (SMEM), else the warp's global scratch slice
    uint8_t *mem = SMEM ? nucl_smem + (size_t) (threadIdx.x >> 5) * ((mem_bytes_cap + 15) / 16 * 16) : gmem;
    const int lane = threadIdx.x & 31;
    const bool with_cigar = !(flag & EZ_SCORE_ONLY);
    const int qe = q + e;
    ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
    ez.max = 0; ez.score = ez.mqe = ez.mte = KSW_NEG_INF; ez.n_cigar = 0; ez.zdropped = 0;
    if (qlen <= 0 || tlen <= 0) return;
    if (-(int) sc_mis > 2 * qe) return;
    const int tlen_ = (tlen + 15) / 16, qlen_ = (qlen + 15) / 16;
    int n_col_ = qlen < tlen ? qlen : tlen;
    n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
    const int n_col = n_col_ * 16;
    const int L = tlen_ * 16;
    // only target indices below qlen + w + 16 are ever touched; when the real stride L is larger than that, the arrays
    // cannot run into each other and a compact stride is equivalent
    const int Ls = min(L, ((qlen + w + 48 + 15) / 16) * 16);
    int8_t *u = (int8_t *) mem, *v = u + Ls, *x = v + Ls, *y = x + Ls, *s = y + Ls;
    uint8_t *sf = (uint8_t *) (s + Ls), *qr = sf + Ls;
    const int mem_bytes = 6 * Ls + qlen_ * 16 + 32;
    int32_t *H = reinterpret_cast<int32_t *>(mem + (mem_bytes + 15) / 16 * 16);
    for (int i = lane; i < mem_bytes; i += 32) mem[i] = 0;
    for (int i = lane; i < Ls; i += 32) H[i] = KSW_NEG_INF;
    __syncwarp();
    for (int t = lane; t < qlen; t += 32) qr[t] = qv.get(qlen - 1 - t);
    for (int t = lane; t < min(tlen, Ls); t += 32) sf[t] = tv.get(t);
    __syncwarp();
    const int8_t qe2 = (int8_t) (qe * 2);
    const uint8_t max_sc = (uint8_t) (int8_t) (sc_mch + qe * 2);

    int last_st = -1, last_en = -1;
    const int n_rows = qlen + tlen - 1;
    for (int r = 0; r < n_rows; ++r) {
        int st = 0, en = tlen - 1;
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
        if (en > ((r + w) >> 1)) en = (r + w) >> 1;
        if (st > en) { ez.zdropped = 1; break; }
        const int st0 = st, en0 = en;
        st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
        int8_t x1, v1;
        if (st > 0) {
            if (st - 1 >= last_st && st - 1 <= last_en) { x1 = x[st - 1]; v1 = v[st - 1]; }
            else x1 = v1 = 0;
        } else { x1 = 0; v1 = r ? (int8_t) q : (int8_t) 0; }
        if (lane == 0 && en >= r) { y[r] = 0; u[r] = r ? (int8_t) q : (int8_t) 0; }
        // scores, 16 bytes per block starting at st0 (runs past en0 exactly as the vector stores do)
        {
            const uint8_t *qrr = qr + (qlen - 1 - r);
            const int n_sc = ((en0 - st0) / 16 + 1) * 16;
            int8_t val[MAX_CHUNKS];
#pragma unroll
            for (int k = 0; k < MAX_CHUNKS; k++) {
                const int i = lane + 32 * k;
                val[k] = 0;
                if (i < n_sc) {
                    const uint8_t sq = sf[st0 + i], sv = qrr[st0 + i];
                    val[k] = (sq == sv) ? sc_mch : sc_mis;
                    if (sq == 4 || sv == 4) val[k] = 0;
                }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < MAX_CHUNKS; k++) {
                const int i = lane + 32 * k;
                if (i < n_sc) s[st0 + i] = val[k];
            }
        }
        __syncwarp();
        // gather every input of the row (old values) before anything is overwritten
        int8_t in_s[MAX_CHUNKS], in_x[MAX_CHUNKS], in_v[MAX_CHUNKS], in_u[MAX_CHUNKS], in_y[MAX_CHUNKS];
#pragma unroll
        for (int k = 0; k < MAX_CHUNKS; k++) {
            const int t = st + lane + 32 * k;
            if (t <= en) {
                in_s[k] = s[t]; in_u[k] = u[t]; in_y[k] = y[t];
                in_x[k] = (t == st) ? x1 : x[t - 1];
                in_v[k] = (t == st) ? v1 : v[t - 1];
            }
        }
        __syncwarp();
        uint8_t *pr = with_cigar ? pmat + (size_t) r * n_col - st : nullptr;
        if (with_cigar && lane == 0) poff[r] = make_int2(st, en);
#pragma unroll
        for (int k = 0; k < MAX_CHUNKS; k++) {
            const int t = st + lane + 32 * k;
            if (t <= en) {
                int8_t z = (int8_t) (in_s[k] + qe2);
                const int8_t vt1 = in_v[k], ut = in_u[k];
                int8_t a = (int8_t) (in_x[k] + vt1);
                int8_t b = (int8_t) (in_y[k] + ut);
                uint8_t d = 0;
                if (with_cigar) {           // gap left-alignment (KSW_EZ_RIGHT is never set by the aligner)
                    d = (a > z) ? 1 : 0;
                    z = z > a ? z : a;
                    if (b > z) d = 2;
                } else {
                    z = z > a ? z : a;
                }
                uint8_t zu = (uint8_t) z;
                const uint8_t bu = (uint8_t) b;
                zu = zu > bu ? zu : bu;
                zu = zu < max_sc ? zu : max_sc;
                z = (int8_t) zu;
                u[t] = (int8_t) (z - vt1);
                v[t] = (int8_t) (z - ut);
                z = (int8_t) (z - (int8_t) q);
                a = (int8_t) (a - z);
                b = (int8_t) (b - z);
                x[t] = a > 0 ? a : (int8_t) 0;
                y[t] = b > 0 ? b : (int8_t) 0;
                if (with_cigar) {
                    if (a > 0) d |= 0x08;
                    if (b > 0) d |= 0x10;
                    pr[t] = d;
                }
            }
        }
        __syncwarp();
        // exact maximum over H (32-bit), with the reference's four-stream tie rule
        int max_H, max_t;
        if (r > 0) {
            const int en1 = st0 + (en0 - st0) / 4 * 4;
            int h_en0 = 0;
            if (lane == 0) h_en0 = en0 > 0 ? H[en0 - 1] + (int) (uint8_t) u[en0] - qe : H[en0] + (int) (uint8_t) v[en0] - qe;
            __syncwarp();
            int sv = KSW_NEG_INF - 1, stt = 0x7fffffff;  // this lane's stream ( (t - st0) & 3 == lane & 3 ): best value, first t
#pragma unroll
            for (int k = 0; k < MAX_CHUNKS; k++) {
                const int t = st0 + lane + 32 * k;
                if (t < en0) {
                    const int h = H[t] + (int) (uint8_t) v[t] - qe;
                    H[t] = h;
                    if (t < en1 && h > sv) { sv = h; stt = t; }
                }
            }
            if (lane == 0) H[en0] = h_en0;
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {  // combine lanes of the same stream: larger value, then smaller t
                const int ov = __shfl_xor_sync(0xffffffffu, sv, o), ot = __shfl_xor_sync(0xffffffffu, stt, o);
                if (ov > sv || (ov == sv && ot < stt)) { sv = ov; stt = ot; }
            }
            max_H = __shfl_sync(0xffffffffu, h_en0, 0);
            max_t = en0;
#pragma unroll
            for (int i = 0; i < 4; i++) {       // streams in order, strictly greater only
                const int hv = __shfl_sync(0xffffffffu, sv, i), ht = __shfl_sync(0xffffffffu, stt, i);
                if (hv > max_H) { max_H = hv; max_t = ht; }
            }
            __syncwarp();
            for (int t = en1; t < en0; ++t) {   // <= 3 tail elements, in order
                const int h = H[t];
                if (h > max_H) { max_H = h; max_t = t; }
            }
        } else {
            if (lane == 0) H[0] = (int) (uint8_t) v[0] - qe - qe;
            __syncwarp();
            max_H = H[0]; max_t = 0;
        }
        __syncwarp();
        const int h_en0_now = H[en0], h_st0_now = H[st0];
        if (en0 == tlen - 1 && h_en0_now > ez.mte) { ez.mte = h_en0_now; ez.mte_q = r - en; }
        if (r - st0 == qlen - 1 && h_st0_now > ez.mqe) { ez.mqe = h_st0_now; ez.mqe_t = st0; }
        {   // ksw_apply_zdrop (every lane evaluates the same scalars)
            bool stop = false;
            if (max_H > ez.max) { ez.max = max_H; ez.max_t = max_t; ez.max_q = r - max_t; }
            else if (max_t >= ez.max_t && r - max_t >= ez.max_q) {
                const int tl = max_t - ez.max_t, ql = (r - max_t) - ez.max_q, l = tl > ql ? tl - ql : ql - tl;
                if (zdrop >= 0 && ez.max - max_H > zdrop + l * e) { ez.zdropped = 1; stop = true; }
            }
            if (stop) break;
        }
        if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H[tlen - 1];
        last_st = st; last_en = en;
    }
    __syncwarp();
    if (with_cigar) {
        int i0 = -1, j0 = -1;
        bool do_bt = true;
        if (!ez.zdropped && !(flag & EZ_EXTZ_ONLY)) { i0 = tlen - 1; j0 = qlen - 1; }
        else if (ez.max_t >= 0 && ez.max_q >= 0) { i0 = ez.max_t; j0 = ez.max_q; }
        else do_bt = false;
        int n = 0;
        if (do_bt && lane == 0) {               // ksw_backtrack (is_rot, no N): sequential walk, emitted end -> start
            int i = i0, j = j0, state = 0;
            while (i >= 0 && j >= 0) {
                int force_state = -1;
                const int r = i + j;
                const int2 oe = poff[r];
                if (i < oe.x) force_state = 2;
                if (i > oe.y) force_state = 1;
                const unsigned tmp = force_state < 0 ? pmat[(size_t) r * n_col + i - oe.x] : 0u;
                if (state == 0) state = tmp & 7;
                else if (!(tmp >> (state + 2) & 1)) state = 0;
                if (state == 0) state = tmp & 7;
                if (force_state >= 0) state = force_state;
                unsigned op;
                if (state == 0) { op = 0; --i; --j; }
                else if (state == 1 || state == 3) { op = 2; --i; }
                else { op = 1; --j; }
                if (n == 0 || op != (cigar[n - 1] & 0xfu)) { if (n < cigar_cap) cigar[n] = 1u << 4 | op; n++; }
                else cigar[n - 1] += 1u << 4;
            }
            if (i >= 0) { if (n == 0 || 2u != (cigar[n - 1] & 0xfu)) { if (n < cigar_cap) cigar[n] = (unsigned) (i + 1) << 4 | 2u; n++; } else cigar[n - 1] += (unsigned) (i + 1) << 4; }
            if (j >= 0) { if (n == 0 || 1u != (cigar[n - 1] & 0xfu)) { if (n < cigar_cap) cigar[n] = (unsigned) (j + 1) << 4 | 1u; n++; } else cigar[n - 1] += (unsigned) (j + 1) << 4; }
            const int m = min(n, cigar_cap);
            for (int k = 0; k < m >> 1; ++k) { const uint32_t t2 = cigar[k]; cigar[k] = cigar[m - 1 - k]; cigar[m - 1 - k] = t2; }
        }
        ez.n_cigar = __shfl_sync(0xffffffffu, n, 0);
        __syncwarp();
    }
}

// DistanceCalculator::computeSubstitutionStartEndDistance on one diagonal (sequential rule; a few hundred steps)
__device__ void seed_segment(const int8_t *smat, const uint8_t *s1, const uint8_t *s2, int length, int &start, int &end, int &score) {
    int maxScore = 0, maxEnd = 0, maxStart = 0, minPos = -1, sc = 0;
    for (int pos = 0; pos < length; pos++) {
        sc += smat[s1[pos] * 5 + s2[pos]];
        if (sc <= 0) { sc = 0; minPos = pos; }
        if (sc > maxScore) { maxEnd = pos; maxStart = minPos + 1; maxScore = sc; }
    }
    start = maxStart; end = maxEnd; score = maxScore;
}

// MINB = resident CTAs per SM the kernel is compiled for (register cap 128 / 80 / 64): the kernel is latency bound (one short
// dependent chain per anti-diagonal), so more resident warps can pay for a few spilled registers; B200_NUCL_MINB picks at run time (default 4: 2.47 M alignments/s against 2.31 M at 2 CTAs/SM).
template <bool SMEM, int MINB>
__global__ void __launch_bounds__(NUCL_WARPS * 32, MINB)
nucl_align_kernel(const NuclTask *__restrict__ tasks, unsigned n_tasks, const uint8_t *__restrict__ qres,
                  const uint64_t *__restrict__ qoff, const uint8_t *__restrict__ db, const uint64_t *__restrict__ off,
                  const int32_t *__restrict__ len, int gapo, int gape, int zdrop, int w, uint8_t *__restrict__ scratch,
                  size_t scratch_stride, size_t mem_bytes, size_t h_bytes, size_t p_bytes, size_t pm_bytes, unsigned *__restrict__ counter,
                  int32_t *__restrict__ out, uint32_t *__restrict__ pool, unsigned long long *__restrict__ pool_used,
                  int cigar_cap_slack) {
    __shared__ int8_t smat[25];
    if (threadIdx.x < 25) {  // nucleotide.out at bit factor 1: +2 / -3, X column/row -3 (NucleotideMatrix), ksw treats X as 0
        const int a = threadIdx.x / 5, b = threadIdx.x % 5;
        smat[threadIdx.x] = (a == b && a < 4) ? 2 : -3;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * NUCL_WARPS + (threadIdx.x >> 5);
    uint8_t *mem = scratch + (size_t) warp_global * scratch_stride;
    const size_t mh_cap = mem_bytes + h_bytes;   // per-warp capacity of (byte arrays + H)
    int2 *poff = reinterpret_cast<int2 *>(mem + mem_bytes + h_bytes);
    uint8_t *pmat = mem + mem_bytes + h_bytes + p_bytes;  // p_bytes = size of the poff area, pm_bytes = direction matrix
    uint32_t *cigar = reinterpret_cast<uint32_t *>(pmat + pm_bytes);   // per-warp CIGAR staging; finished ops go to the pool
    const int8_t sc_mch = 2, sc_mis = -3;

    while (true) {
        unsigned ti = 0;
        if (lane == 0) ti = atomicAdd(counter, 1u);
        ti = __shfl_sync(0xffffffffu, ti, 0);
        if (ti >= n_tasks) break;
        const NuclTask task = tasks[ti];
        const uint8_t *q = qres + qoff[task.query];
        const int qL = (int) (qoff[task.query + 1] - qoff[task.query]);
        const uint8_t *t = db + off[task.target];
        const int tL = len[task.target];
        const int cigar_cap = 2 * qL + cigar_cap_slack;
        int32_t *o = out + (size_t) ti * 8;
        // ---- ungapped seed on the prefilter diagonal (two wrap candidates), lane 0
        int best_score = 0, best_start = -1, best_end = -1, best_diag = 0, best_dist = 0;
        if (lane == 0) {
            const int cand[2] = {-65536 + (int) task.diagonal, (int) task.diagonal};
            for (int c = 0; c < 2; c++) {
                const int d = cand[c], dist = d < 0 ? -d : d;
                int st = 0, en = 0, sc = 0;
                if (d >= 0 && dist < qL) seed_segment(smat, q + dist, t, min(tL, qL - dist), st, en, sc);
                else if (d < 0 && dist < tL) seed_segment(smat, q, t + dist, min(tL - dist, qL), st, en, sc);
                if (sc > best_score) { best_score = sc; best_start = st; best_end = en; best_diag = d; best_dist = dist; }
            }
        }
        best_score = __shfl_sync(0xffffffffu, best_score, 0); best_start = __shfl_sync(0xffffffffu, best_start, 0);
        best_end = __shfl_sync(0xffffffffu, best_end, 0); best_diag = __shfl_sync(0xffffffffu, best_diag, 0);
        best_dist = __shfl_sync(0xffffffffu, best_dist, 0);
        int qUs, qUe, dUs, dUe;
        if (best_diag >= 0) { qUs = best_start + best_dist; qUe = best_end + best_dist; dUs = best_start; dUe = best_end; }
        else { qUs = best_start; qUe = best_end; dUs = best_start + best_dist; dUe = best_end + best_dist; }
        if (qUe - qUs == qL - 1 && dUs == 0 && dUe == tL - 1) {  // the seed already spans both sequences (:132-159)
            int ids = 0;
            for (int i = qUs + lane; i <= qUe; i += 32) ids += q[i] == t[dUs + (i - qUs)];
#pragma unroll
            for (int s2 = 16; s2 > 0; s2 >>= 1) ids += __shfl_xor_sync(0xffffffffu, ids, s2);
            if (lane == 0) {
                const unsigned long long po = atomicAdd(pool_used, 1ull);
                pool[po] = (uint32_t) qL << 4;
                o[0] = best_score; o[1] = qUs; o[2] = qUe; o[3] = dUs; o[4] = dUe; o[5] = ids; o[6] = 1; o[7] = (int32_t) (uint32_t) po;
            }
            continue;
        }
        const int qStartRev = (qL - qUe) - 1, tStartRev = (tL - dUe) - 1;
        SeqView qrv = {q, qL, qStartRev, 1}, trv = {t, tL, tStartRev, 1};
        Ez ez, ezA;
        ksw_extz2_warp<SMEM>(qL - qStartRev, qrv, tL - tStartRev, trv, sc_mch, sc_mis, gapo, gape, w, zdrop, EZ_SCORE_ONLY | EZ_EXTZ_ONLY, mem, mh_cap,
                       pmat, poff, ez, cigar, cigar_cap);
        const int qStartPos = qL - (qStartRev + ez.max_q) - 1, tStartPos = tL - (tStartRev + ez.max_t) - 1;
        SeqView qfv = {q, qL, qStartPos, 0}, tfv = {t, tL, tStartPos, 0};
        ksw_extz2_warp<SMEM>(qL - qStartPos, qfv, tL - tStartPos, tfv, sc_mch, sc_mis, gapo, gape, w, zdrop, EZ_EXTZ_ONLY, mem, mh_cap, pmat, poff,
                       ezA, cigar, cigar_cap);
        bool reversed = false;
        if (ez.max_q > ezA.max_q && ez.max_t > ezA.max_t) {
            ksw_extz2_warp<SMEM>(qL - qStartRev, qrv, tL - tStartRev, trv, sc_mch, sc_mis, gapo, gape, w, zdrop, EZ_EXTZ_ONLY, mem, mh_cap, pmat,
                           poff, ezA, cigar, cigar_cap);
            reversed = true;
        }
        const int n = min(ezA.n_cigar, cigar_cap);
        if (lane == 0) {
            if (reversed) for (int k = 0; k < n >> 1; ++k) { const uint32_t t2 = cigar[k]; cigar[k] = cigar[n - 1 - k]; cigar[n - 1 - k] = t2; }
            int ids = 0, tp = tStartPos, qp = qStartPos;
            for (int c = 0; c < n; c++) {
                const uint32_t cv = cigar[c];
                const int op = (int) (cv & 0xfu);
                const int ln = (int) (cv >> 4);
                if (op == 0) { for (int i = 0; i < ln; i++) ids += t[tp + i] == q[qp + i]; tp += ln; qp += ln; }
                else if (op == 1) qp += ln;
                else tp += ln;
            }
            const unsigned long long po = atomicAdd(pool_used, (unsigned long long) n);
            for (int c = 0; c < n; c++) pool[po + c] = cigar[c];
            o[0] = ezA.max; o[1] = qStartPos; o[2] = qStartPos + ezA.max_q; o[3] = tStartPos; o[4] = tStartPos + ezA.max_t;
            o[5] = ids; o[6] = n; o[7] = (int32_t) (uint32_t) po;
        }
        __syncwarp();
    }
}

}  // namespace

// ---- host ---------------------------------------------------------------------------------------------------------
int b200_nucl_align(b200_ctx *ctx, const uint8_t *query_residues, const uint64_t *query_offsets, uint32_t n_queries,
                    const b200_nucl_task *tasks, uint64_t n, int gap_open, int gap_extend, int zdrop, b200_nucl_aln *out,
                    uint32_t *cigars, const uint64_t *cigar_offsets) {
    if (ctx == nullptr) return B200_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->n_seq == 0) return b200_set_err(ctx, B200_ERR_NODB, "no target DB loaded");
    if (ctx->alphabet != 5) return b200_set_err(ctx, B200_ERR_ARG, "b200_nucl_align: the loaded DB is not a nucleotide DB (alphabet 5)");
    if (n == 0) return B200_OK;
    if (query_residues == nullptr || query_offsets == nullptr || tasks == nullptr || out == nullptr || cigars == nullptr ||
        cigar_offsets == nullptr || n_queries == 0)
        return b200_set_err(ctx, B200_ERR_ARG, "b200_nucl_align: NULL argument");
    if (n >= 0xffffffffull) return b200_set_err(ctx, B200_ERR_RANGE, "b200_nucl_align: too many tasks");
    if (gap_open < 0 || gap_extend < 0 || gap_open + gap_extend > 60) return b200_set_err(ctx, B200_ERR_ARG, "b200_nucl_align: gap penalties out of the int8 range of ksw2");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const int w = 64;
    const uint64_t q_total = query_offsets[n_queries];
    int max_q = 1, max_t = 1;
    for (uint64_t i = 0; i < q_total; i++) if (query_residues[i] > 4) return b200_set_err(ctx, B200_ERR_ARG, "b200_nucl_align: query residue code > 4");
    std::vector<NuclTask> h_tasks(n);
    for (uint64_t i = 0; i < n; i++) {
        if (tasks[i].query >= n_queries || tasks[i].target >= ctx->n_seq) return b200_set_err(ctx, B200_ERR_ARG, "b200_nucl_align: task index out of range");
        const int qL = (int) (query_offsets[tasks[i].query + 1] - query_offsets[tasks[i].query]);
        const int tL = ctx->h_len[tasks[i].target];
        if (qL <= 0 || qL >= 32768 || tL >= 32768) return b200_set_err(ctx, B200_ERR_RANGE, "b200_nucl_align: sequences must be 1..32767 long (longer ones are split upstream, blastn.sh:27-52)");
        if (cigar_offsets[i + 1] - cigar_offsets[i] < (uint64_t) (2 * qL + w + 8)) return b200_set_err(ctx, B200_ERR_ARG, "b200_nucl_align: cigar slot smaller than 2*qlen + 72");
        if (cigar_offsets[i] >= 0xffffffffull) return b200_set_err(ctx, B200_ERR_RANGE, "b200_nucl_align: cigar buffer too large");
        max_q = std::max(max_q, qL); max_t = std::max(max_t, tL);
        h_tasks[i].query = tasks[i].query; h_tasks[i].target = tasks[i].target; h_tasks[i].diagonal = tasks[i].diagonal;
        h_tasks[i].cigar_off = (uint32_t) cigar_offsets[i];
    }
    // per-warp scratch: ksw byte arrays, H, per-row (st,en), direction matrix
    const size_t Ls = std::min<size_t>(round_up((size_t) max_t, 16), round_up((size_t) max_q + w + 48, 16));
    const size_t mem_bytes = round_up(6 * Ls + round_up((size_t) max_q, 16) + 32, 16);
    const size_t h_bytes = Ls * sizeof(int32_t);
    const size_t rows = std::min<size_t>((size_t) max_q + max_t, 2 * (size_t) std::min(max_q, max_t) + w + 16);
    const size_t poff_bytes = round_up(rows * sizeof(int2), 16);
    const size_t n_col = ((std::min<size_t>(std::min(max_q, max_t), w + 1) + 15) / 16 + 1) * 16;
    const size_t pm_bytes = round_up(rows * n_col + 64, 16);
    const size_t cig_stage_bytes = round_up(((size_t) 2 * max_q + w + 8) * sizeof(uint32_t), 16);
    const size_t stride = mem_bytes + h_bytes + poff_bytes + pm_bytes + cig_stage_bytes;
    const size_t smem_need = round_up(mem_bytes + h_bytes, 16) * NUCL_WARPS;
    const bool use_smem = smem_need <= 96 * 1024;
    int per_sm = 0;
    static const int minb = [] { const char *e = getenv("B200_NUCL_MINB"); const int v = e ? atoi(e) : 4; return v >= 4 ? 4 : (v == 3 ? 3 : 2); }();
#define NUCL_DISPATCH(EXPR_TRUE, EXPR_FALSE) \
    do { if (use_smem) { if (minb == 4) { EXPR_TRUE(4); } else if (minb == 3) { EXPR_TRUE(3); } else { EXPR_TRUE(2); } } \
         else { if (minb == 4) { EXPR_FALSE(4); } else if (minb == 3) { EXPR_FALSE(3); } else { EXPR_FALSE(2); } } } while (0)
#define NUCL_ATTR_T(M) do { cudaFuncSetAttribute(nucl_align_kernel<true, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_need); \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nucl_align_kernel<true, M>, NUCL_WARPS * 32, smem_need) != cudaSuccess) per_sm = 1; } while (0)
#define NUCL_ATTR_F(M) do { if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nucl_align_kernel<false, M>, NUCL_WARPS * 32, 0) != cudaSuccess) per_sm = 1; } while (0)
    NUCL_DISPATCH(NUCL_ATTR_T, NUCL_ATTR_F);
    per_sm = std::max(1, per_sm);
    const unsigned grid = (unsigned) std::max<uint64_t>(1, std::min<uint64_t>((uint64_t) ctx->sm_count * per_sm, (n + NUCL_WARPS - 1) / NUCL_WARPS));
    uint64_t pool_cap = 0;   // worst case: every alignment fills its slot; what comes back over PCIe is only what was used
    for (uint64_t i = 0; i < n; i++) pool_cap += 2 * (query_offsets[tasks[i].query + 1] - query_offsets[tasks[i].query]) + w + 8;
    if (pool_cap >= 0xffffffffull) return b200_set_err(ctx, B200_ERR_RANGE, "b200_nucl_align: CIGAR pool too large, split the batch");
    DevBuf d_tasks, d_q, d_qoff, d_scratch, d_out, d_cig, d_used;
    cudaError_t e = d_tasks.reserve(sizeof(NuclTask) * n);
    if (e == cudaSuccess) e = d_q.reserve(q_total + 16);
    if (e == cudaSuccess) e = d_qoff.reserve(sizeof(uint64_t) * ((size_t) n_queries + 1));
    if (e == cudaSuccess) e = d_scratch.reserve(stride * (size_t) grid * NUCL_WARPS);
    if (e == cudaSuccess) e = d_out.reserve(sizeof(int32_t) * 8 * n);
    if (e == cudaSuccess) e = d_cig.reserve(sizeof(uint32_t) * pool_cap + 16);
    if (e == cudaSuccess) e = d_used.reserve(sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_used.p, 0, sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess) e = ctx->counter.reserve(sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->counter.p, 0, sizeof(unsigned), ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tasks.p, h_tasks.data(), sizeof(NuclTask) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_q.p, query_residues, q_total, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_qoff.p, query_offsets, sizeof(uint64_t) * ((size_t) n_queries + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[14], ctx->stream);
    if (e == cudaSuccess) {
#define NUCL_ARGS d_tasks.as<NuclTask>(), (unsigned) n, d_q.as<uint8_t>(), d_qoff.as<uint64_t>(), ctx->d_res, ctx->d_off, ctx->d_len, gap_open, \
                gap_extend, zdrop, w, d_scratch.as<uint8_t>(), stride, mem_bytes, h_bytes, poff_bytes, pm_bytes, ctx->counter.as<unsigned>(), \
                d_out.as<int32_t>(), d_cig.as<uint32_t>(), d_used.as<unsigned long long>(), w + 8
#define NUCL_LAUNCH_T(M) nucl_align_kernel<true, M><<<grid, NUCL_WARPS * 32, smem_need, ctx->stream>>>(NUCL_ARGS)
#define NUCL_LAUNCH_F(M) nucl_align_kernel<false, M><<<grid, NUCL_WARPS * 32, 0, ctx->stream>>>(NUCL_ARGS)
        NUCL_DISPATCH(NUCL_LAUNCH_T, NUCL_LAUNCH_F);
        ctx->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[15], ctx->stream);
    }
    std::vector<int32_t> h_out(8 * n);
    unsigned long long used = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out.data(), d_out.p, sizeof(int32_t) * 8 * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&used, d_used.p, sizeof(used), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    std::vector<uint32_t> h_pool((size_t) used + 1);
    if (e == cudaSuccess && used > 0) e = cudaMemcpyAsync(h_pool.data(), d_cig.p, sizeof(uint32_t) * used, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev[14], ctx->ev[15]);
    d_tasks.release(); d_q.release(); d_qoff.release(); d_scratch.release(); d_out.release(); d_cig.release(); d_used.release();
    if (e != cudaSuccess) { ctx->err = std::string("b200_nucl_align: ") + cudaGetErrorString(e); return B200_ERR_CUDA; }
    for (uint64_t i = 0; i < n; i++) {
        const int32_t *o = h_out.data() + 8 * i;
        out[i].score = o[0]; out[i].qstart = o[1]; out[i].qend = o[2]; out[i].dbstart = o[3]; out[i].dbend = o[4];
        out[i].identical = o[5]; out[i].n_cigar = o[6];
        if (o[6] > 0) memcpy(cigars + cigar_offsets[i], h_pool.data() + (uint32_t) o[7], sizeof(uint32_t) * (size_t) o[6]);
    }
    return B200_OK;
}
