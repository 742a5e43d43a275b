/* oracle/oracle_ksw.c -- TEST INFRASTRUCTURE (see oracle/oracle.c header for the rules).
 *
 * A7: the nucleotide gapped aligner.  Restates
 *   ksw_extz2_sse                  lib/ksw2/ksw2_extz2_sse.cpp:44-285   (Suzuki-Kasahara difference DP, band w, z-drop)
 *   ksw_apply_zdrop / ksw_backtrack lib/ksw2/ksw2.h:186-202, :145-177
 *   DistanceCalculator::computeSubstitutionStartEndDistance / ungappedAlignmentByDiagonal
 *                                  src/alignment/DistanceCalculator.h:178-200, :115-174
 *   BandedNucleotideAligner::align src/alignment/BandedNucleotideAligner.cpp:73-263
 *
 * ksw2's results depend on artefacts of its 16-lane SSE blocks (the band is rounded outwards to multiples of 16, lanes
 * outside the band carry whatever the previous rows left there, score bytes are written 16 at a time past the band end,
 * the running maximum is tracked in four interleaved lane streams).  They are part of the reference's observable
 * behaviour, so this restatement keeps the same byte arrays and applies every vector operation lane by lane in scalar C.
 *
 * Parity status: PINNED against oracle/_ref (ref_ksw_extz2, ref_banded_nucl_align) by tests/test_oracle_vs_ref.py and
 * against tests/golden/nucl_v1.npz.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KSW_NEG_INF (-0x40000000)
#define KSW_EZ_SCORE_ONLY 0x01
#define KSW_EZ_RIGHT 0x02
#define KSW_EZ_GENERIC_SC 0x04
#define KSW_EZ_APPROX_MAX 0x08
#define KSW_EZ_APPROX_DROP 0x10
#define KSW_EZ_EXTZ_ONLY 0x40
#define KSW_EZ_REV_CIGAR 0x80

typedef struct {
    int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar;
} orc_ez;

static int ez_zdrop(orc_ez *ez, int32_t H, int r, int t, int zdrop, int e) { /* ksw2.h:186-202, is_rot = 1 */
    if (H > ez->max) {
        ez->max = H; ez->max_t = t; ez->max_q = r - t;
    } else if (t >= ez->max_t && r - t >= ez->max_q) {
        int tl = t - ez->max_t, ql = (r - t) - ez->max_q, l = tl > ql ? tl - ql : ql - tl;
        if (zdrop >= 0 && ez->max - H > zdrop + l * e) { ez->zdropped = 1; return 1; }
    }
    return 0;
}

/* cigar: up to cap u32 (len<<4|op, op 0=M 1=I 2=D), written in final order; returns n_cigar via ez */
void orc_ksw_extz2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat, int q, int e,
                   int w, int zdrop, int flag, orc_ez *ez, uint32_t *cigar, int cap) {
    const int with_cigar = !(flag & KSW_EZ_SCORE_ONLY);
    const int qe = q + e;
    ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
    ez->max = 0; ez->score = ez->mqe = ez->mte = KSW_NEG_INF; ez->n_cigar = 0; ez->zdropped = 0;
    if (m <= 0 || qlen <= 0 || tlen <= 0) return;
    if (w < 0) w = tlen > qlen ? tlen : qlen;
    const int wl = w, wr = w;
    const int tlen_ = (tlen + 15) / 16, qlen_ = (qlen + 15) / 16;
    int n_col_ = qlen < tlen ? qlen : tlen;
    n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
    int min_sc = mat[1];
    for (int t = 1; t < m * m; ++t) min_sc = min_sc < mat[t] ? min_sc : mat[t];
    if (-min_sc > 2 * qe) return;
    const int8_t sc_mch = mat[0], sc_mis = mat[1], qe2 = (int8_t) (qe * 2);
    const uint8_t max_sc = (uint8_t) (int8_t) (mat[0] + qe * 2);
    /* one zeroed block with the reference's layout: u v x y s (tlen_*16 each), sf (tlen_*16), qr (qlen_*16 + 16) */
    const size_t L = (size_t) tlen_ * 16;
    uint8_t *mem = (uint8_t *) calloc(L * 6 + (size_t) qlen_ * 16 + 32, 1);
    int8_t *u = (int8_t *) mem, *v = u + L, *x = v + L, *y = x + L, *s = y + L;
    uint8_t *sf = (uint8_t *) (s + L), *qr = sf + L;
    int32_t *H = (int32_t *) malloc(L * sizeof(int32_t));
    for (size_t t = 0; t < L; ++t) H[t] = KSW_NEG_INF;
    uint8_t *p = NULL;
    int *off = NULL, *off_end = NULL;
    const int n_col = n_col_ * 16;
    if (with_cigar) {
        p = (uint8_t *) calloc((size_t) (qlen + tlen - 1) * n_col + 32, 1);
        off = (int *) malloc((size_t) (qlen + tlen - 1) * sizeof(int) * 2);
        off_end = off + qlen + tlen - 1;
    }
    for (int t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];
    memcpy(sf, target, (size_t) tlen);

    int last_st = -1, last_en = -1;
    for (int r = 0; r < qlen + tlen - 1; ++r) {
        int st = 0, en = tlen - 1;
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < ((r - wr + 1) >> 1)) st = (r - wr + 1) >> 1;
        if (en > ((r + wl) >> 1)) en = (r + wl) >> 1;
        if (st > en) { ez->zdropped = 1; break; }
        const int st0 = st, en0 = en;
        st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
        int8_t x1, v1;
        if (st > 0) {
            if (st - 1 >= last_st && st - 1 <= last_en) { x1 = x[st - 1]; v1 = v[st - 1]; }
            else x1 = v1 = 0;
        } else { x1 = 0; v1 = r ? (int8_t) q : 0; }
        if (en >= r) { y[r] = 0; u[r] = r ? (int8_t) q : 0; }
        const uint8_t *qrr = qr + (qlen - 1 - r);   /* may point below qr for r >= qlen: stays inside `mem` (sf region) */
        for (int t = st0; t <= en0; t += 16) {      /* 16 score bytes per store, past en0 as the reference does */
            for (int k = 0; k < 16; k++) {
                const uint8_t sq = sf[t + k], sv = qrr[t + k];
                int8_t val = (sq == sv) ? sc_mch : sc_mis;
                if (sq == (uint8_t) (m - 1) || sv == (uint8_t) (m - 1)) val = 0;
                s[t + k] = val;
            }
        }
        uint8_t *pr = with_cigar ? p + (size_t) r * n_col - st : NULL;
        if (with_cigar) { off[r] = st; off_end[r] = en; }
        int8_t xprev = x1, vprev = v1;              /* old x[t-1], v[t-1] */
        for (int t = st; t <= en; ++t) {
            int8_t z = (int8_t) (s[t] + qe2);
            const int8_t xt1 = xprev, vt1 = vprev;
            xprev = x[t]; vprev = v[t];
            int8_t a = (int8_t) (xt1 + vt1);
            const int8_t ut = u[t];
            int8_t b = (int8_t) (y[t] + ut);
            uint8_t d = 0;
            if (with_cigar) {
                if (!(flag & KSW_EZ_RIGHT)) {
                    d = (a > z) ? 1 : 0;
                    z = z > a ? z : a;
                    if (b > z) d = 2;
                } else {
                    d = (z > a) ? 0 : 1;
                    z = z > a ? z : a;
                    if (!(z > b)) d = 2;
                }
            } else {
                z = z > a ? z : a;
            }
            uint8_t zu = (uint8_t) z, bu = (uint8_t) b;
            zu = zu > bu ? zu : bu;                 /* max_epu8 */
            zu = zu < max_sc ? zu : max_sc;         /* min_epu8 */
            z = (int8_t) zu;
            u[t] = (int8_t) (z - vt1);
            v[t] = (int8_t) (z - ut);
            z = (int8_t) (z - (int8_t) q);
            a = (int8_t) (a - z);
            b = (int8_t) (b - z);
            if (!with_cigar) {
                x[t] = a > 0 ? a : 0;
                y[t] = b > 0 ? b : 0;
            } else if (!(flag & KSW_EZ_RIGHT)) {
                x[t] = a > 0 ? a : 0; if (a > 0) d |= 0x08;
                y[t] = b > 0 ? b : 0; if (b > 0) d |= 0x10;
                pr[t] = d;
            } else {
                x[t] = (0 > a) ? 0 : a; if (!(0 > a)) d |= 0x08;
                y[t] = (0 > b) ? 0 : b; if (!(0 > b)) d |= 0x10;
                pr[t] = d;
            }
        }
        /* exact maximum with the 32-bit H array (approx_max is never set by BandedNucleotideAligner) */
        int32_t max_H, max_t;
        if (r > 0) {
            const int en1 = st0 + (en0 - st0) / 4 * 4;
            max_H = H[en0] = en0 > 0 ? H[en0 - 1] + (uint8_t) u[en0] - qe : H[en0] + (uint8_t) v[en0] - qe;
            max_t = en0;
            int32_t HH[4], tt[4];
            for (int i = 0; i < 4; i++) { HH[i] = max_H; tt[i] = max_t; }
            int t;
            for (t = st0; t < en1; t += 4) {
                for (int i = 0; i < 4; i++) {
                    H[t + i] += (int32_t) (uint8_t) v[t + i] - qe;
                    if (H[t + i] > HH[i]) { HH[i] = H[t + i]; tt[i] = t; }
                }
            }
            for (int i = 0; i < 4; i++) if (max_H < HH[i]) { max_H = HH[i]; max_t = tt[i] + i; }
            for (; t < en0; ++t) {
                H[t] += (int32_t) (uint8_t) v[t] - qe;
                if (H[t] > max_H) { max_H = H[t]; max_t = t; }
            }
        } else { H[0] = (uint8_t) v[0] - qe - qe; max_H = H[0]; max_t = 0; }
        if (en0 == tlen - 1 && H[en0] > ez->mte) { ez->mte = H[en0]; ez->mte_q = r - en; }
        if (r - st0 == qlen - 1 && H[st0] > ez->mqe) { ez->mqe = H[st0]; ez->mqe_t = st0; }
        if (ez_zdrop(ez, max_H, r, max_t, zdrop, e)) break;
        if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
        last_st = st; last_en = en;
    }
    if (with_cigar) {
        int i0, j0, do_bt = 1;
        if (!ez->zdropped && !(flag & KSW_EZ_EXTZ_ONLY)) { i0 = tlen - 1; j0 = qlen - 1; }
        else if (ez->max_t >= 0 && ez->max_q >= 0) { i0 = ez->max_t; j0 = ez->max_q; }
        else { do_bt = 0; i0 = j0 = -1; }
        if (do_bt) {                                /* ksw_backtrack, is_rot = 1, with_N = 0 */
            int n = 0, i = i0, j = j0, state = 0;
            uint32_t *cg = (uint32_t *) malloc((size_t) (qlen + tlen + 2) * sizeof(uint32_t));
            while (i >= 0 && j >= 0) {
                int force_state = -1;
                const int r = i + j;
                if (i < off[r]) force_state = 2;
                if (i > off_end[r]) force_state = 1;
                const uint32_t tmp = force_state < 0 ? p[(size_t) r * n_col + i - off[r]] : 0;
                if (state == 0) state = tmp & 7;
                else if (!(tmp >> (state + 2) & 1)) state = 0;
                if (state == 0) state = tmp & 7;
                if (force_state >= 0) state = force_state;
                uint32_t op;
                if (state == 0) { op = 0; --i; --j; }
                else if (state == 1 || state == 3) { op = 2; --i; }
                else { op = 1; --j; }
                if (n == 0 || op != (cg[n - 1] & 0xf)) cg[n++] = 1u << 4 | op; else cg[n - 1] += 1u << 4;
            }
            if (i >= 0) { if (n == 0 || 2 != (cg[n - 1] & 0xf)) cg[n++] = (uint32_t) (i + 1) << 4 | 2; else cg[n - 1] += (uint32_t) (i + 1) << 4; }
            if (j >= 0) { if (n == 0 || 1 != (cg[n - 1] & 0xf)) cg[n++] = (uint32_t) (j + 1) << 4 | 1; else cg[n - 1] += (uint32_t) (j + 1) << 4; }
            if (!(flag & KSW_EZ_REV_CIGAR))
                for (int k = 0; k < n >> 1; ++k) { uint32_t t2 = cg[k]; cg[k] = cg[n - 1 - k]; cg[n - 1 - k] = t2; }
            ez->n_cigar = n;
            for (int k = 0; k < n && k < cap; k++) cigar[k] = cg[k];
            free(cg);
        }
        free(p); free(off);
    }
    free(mem); free(H);
}

/* DistanceCalculator::computeSubstitutionStartEndDistance (:178-200) on one diagonal segment */
static void seg_best(const int8_t *mat5, const uint8_t *s1, const uint8_t *s2, int length, int *start, int *end, int *score) {
    int maxScore = 0, maxEnd = 0, maxStart = 0, minPos = -1, sc = 0;
    for (int pos = 0; pos < length; pos++) {
        sc += mat5[s1[pos] * 5 + s2[pos]];
        const int isMin = sc <= 0;
        if (isMin) { sc = 0; minPos = pos; }
        if (sc > maxScore) { maxEnd = pos; maxStart = minPos + 1; maxScore = sc; }
    }
    *start = maxStart; *end = maxEnd; *score = maxScore;
}

/* BandedNucleotideAligner::align (:73-263), non-wrapped scoring, forward strand, sequences < 32768 (T6).
 * q/t numeric (A,C,T,G,X = 0..4); smat = the 5x5 NucleotideMatrix scores used for the ungapped seed (fastMatrix),
 * kmat = the 5x5 ksw matrix the aligner builds (mat[0]/mat[1] match/mismatch).  diagonal as passed by the prefilter (u16).
 * out: score, qStart, qEnd, dbStart, dbEnd, identical, n_cigar; cigar in alignment order; bt = M/I/D string (cap bytes) */
void orc_banded_nucl_align(const uint8_t *q, int qL, const uint8_t *t, int tL, uint16_t diagonal16, const int8_t *smat,
                           const int8_t *kmat, int gapo, int gape, int zdrop, int32_t *out, uint32_t *cigar, int cap, char *bt) {
    /* ungapped seed: computeUngappedAlignment with one wrap candidate each way (:93-112) */
    int best_score = 0, best_start = -1, best_end = -1, best_diag = 0, best_dist = 0;   /* LocalAlignment() defaults */
    int cand[2];
    cand[0] = -65536 + (int) diagonal16;
    cand[1] = (int) diagonal16;
    for (int c = 0; c < 2; c++) {
        const int d = cand[c];
        const int dist = d < 0 ? -d : d;
        int st = 0, en = 0, sc = 0, ok = 0;
        if (d >= 0 && dist < qL) { int n = tL < qL - dist ? tL : qL - dist; seg_best(smat, q + dist, t, n, &st, &en, &sc); ok = 1; }
        else if (d < 0 && dist < tL) { int n = tL - dist < qL ? tL - dist : qL; seg_best(smat, q, t + dist, n, &st, &en, &sc); ok = 1; }
        (void) ok;
        if (sc > best_score) { best_score = sc; best_start = st; best_end = en; best_diag = d; best_dist = dist; }
    }
    int qUs, qUe, dUs, dUe;
    if (best_diag >= 0) { qUs = best_start + best_dist; qUe = best_end + best_dist; dUs = best_start; dUe = best_end; }
    else { qUs = best_start; qUe = best_end; dUs = best_start + best_dist; dUe = best_end + best_dist; }
    int n_bt = 0;
    if (qUe - qUs == qL - 1 && dUs == 0 && dUe == tL - 1) {
        out[0] = best_score; out[1] = qUs; out[2] = qUe; out[3] = dUs; out[4] = dUe;
        int ids = 0;
        for (int i = qUs; i <= qUe; i++) ids += q[i] == t[dUs + (i - qUs)];
        out[5] = ids; out[6] = 1;
        if (cap > 0) cigar[0] = (uint32_t) qL << 4;
        for (int i = 0; i < qL && i < cap - 1; i++) bt[n_bt++] = 'M';
        bt[n_bt] = 0;
        return;
    }
    /* The reference reverses with seq_reverse(rev, seq, L) -- L, not L-1 (BandedNucleotideAligner.cpp:60,88) -- so
     * rev[k] = seq[L-k], k = 0..L: shifted by one, and rev[0] is the byte PAST the sequence, which in the reference is
     * whatever the Sequence buffer held before.  That byte is defined here (and forced in the pinning harness) as X. */
    uint8_t *qrev = (uint8_t *) malloc((size_t) qL + 2), *trev = (uint8_t *) malloc((size_t) tL + 2);
    qrev[0] = 4; trev[0] = 4;
    for (int i = 1; i <= qL; i++) qrev[i] = q[qL - i];
    for (int i = 1; i <= tL; i++) trev[i] = t[tL - i];
    const int qStartRev = (qL - qUe) - 1, tStartRev = (tL - dUe) - 1;
    orc_ez ez, ezA;
    orc_ksw_extz2(qL - qStartRev, qrev + qStartRev, tL - tStartRev, trev + tStartRev, 5, kmat, gapo, gape, 64, zdrop,
                  KSW_EZ_SCORE_ONLY | KSW_EZ_EXTZ_ONLY, &ez, NULL, 0);
    const int qStartPos = qL - (qStartRev + ez.max_q) - 1, tStartPos = tL - (tStartRev + ez.max_t) - 1;
    uint32_t *cg = (uint32_t *) malloc((size_t) (qL + tL + 2) * sizeof(uint32_t));
    orc_ksw_extz2(qL - qStartPos, q + qStartPos, tL - tStartPos, t + tStartPos, 5, kmat, gapo, gape, 64, zdrop,
                  KSW_EZ_EXTZ_ONLY, &ezA, cg, qL + tL + 2);
    int reversed = 0;
    if (ez.max_q > ezA.max_q && ez.max_t > ezA.max_t) {
        orc_ksw_extz2(qL - qStartRev, qrev + qStartRev, tL - tStartRev, trev + tStartRev, 5, kmat, gapo, gape, 64, zdrop,
                      KSW_EZ_EXTZ_ONLY, &ezA, cg, qL + tL + 2);
        reversed = 1;
    }
    const int n = ezA.n_cigar;
    for (int i = 0; i < n && i < cap; i++) cigar[i] = reversed ? cg[n - 1 - i] : cg[i];
    out[0] = ezA.max; out[1] = qStartPos; out[2] = qStartPos + ezA.max_q; out[3] = tStartPos; out[4] = tStartPos + ezA.max_t;
    int ids = 0, tp = tStartPos, qp = qStartPos;
    for (int c = 0; c < n; c++) {
        const uint32_t cv = reversed ? cg[n - 1 - c] : cg[c];
        const int op = (int) (cv & 0xf);
        const uint32_t len = cv >> 4;
        for (uint32_t i = 0; i < len; i++) {
            char letter = "MID"[op];
            if (op == 0) { ids += t[tp] == q[qp]; ++qp; ++tp; }
            else if (op == 1) ++qp;
            else ++tp;
            if (n_bt < cap - 1) bt[n_bt++] = letter;
        }
    }
    bt[n_bt] = 0;
    out[5] = ids; out[6] = n;
    free(cg); free(qrev); free(trev);
}
