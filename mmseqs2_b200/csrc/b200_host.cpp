// mmseqs2_b200/csrc/b200_host.cpp -- host-side profile construction (see include/b200_host.h for the reference cites).
// Compiled with -ffp-contract=off: the float/double mixing below must not be fused.
#include "b200_host.h"

#include <algorithm>
#include <cstdlib>

extern "C" {

void b200h_comp_bias(const int16_t *mat, const double *pback, int A, const uint8_t *seq, int L, float scale, float *out) {
    const int windowSize = 40;
    for (int i = 0; i < L; i++) {
        const int lo = std::max(0, i - windowSize / 2);
        const int hi = std::min(L, i + windowSize / 2);
        const int16_t *row = mat + (size_t) seq[i] * A;
        int sum = 0;
        for (int j = lo; j < hi; j++) sum += row[seq[j]];
        sum -= row[seq[i]];
        float delta = (float) sum;
        delta = (float) ((double) delta / (-1.0 * (double) (float) (hi - lo)));               // float /= double
        for (int a = 0; a < A; a++) delta = (float) ((double) delta + pback[a] * (double) (float) row[a]);  // float += double*float
        out[i] = scale * delta;
    }
}

void b200h_round_bias_ssw(const float *in, int L, int8_t *out) {
    for (int i = 0; i < L; i++) {
        const double v = (in[i] < 0.0) ? (in[i] - 0.5) : (in[i] + 0.5);
        out[i] = (int8_t) v;
    }
}

void b200h_round_bias_diag(const float *in, int L, int8_t *out) {
    for (int i = 0; i < L; i++) {
        float v = in[i];
        v = (v < 0.0) ? (float) (v / 4 - 0.5) : (float) (v / 4 + 0.5);
        out[i] = (int8_t) static_cast<char>(v);
    }
}

int b200h_ssw_bias(const int16_t *mat, int A, const int8_t *cb, int L, int cb_enabled) {
    int bias = 0;
    for (int i = 0; i < A * A; i++) bias = std::min(bias, (int) (int8_t) mat[i]);
    int comp = 0;
    if (cb_enabled) {
        for (int i = 0; i < L; i++) comp = std::min(comp, (int) cb[i]);
    }
    return std::abs(bias) + std::abs(comp);
}

int b200h_build_profile(const int16_t *mat, int A, const uint8_t *q, int L, const int8_t *cb, int target_major, int8_t *out) {
    int rc = 0;
    for (int a = 0; a < A; a++)
        for (int j = 0; j < L; j++) {
            const int m = target_major ? mat[(size_t) a * A + q[j]] : mat[(size_t) q[j] * A + a];
            const int v = m + (cb ? cb[j] : 0);
            if (v < -127 || v > 127) rc = -4;
            out[(size_t) a * L + j] = (int8_t) v;
        }
    return rc;
}

// HMM_PROFILE branch of ssw_init (StripedSmithWaterman.cpp:1388-1406): profile->mat = the alignment profile with the X row
// cleared; bias = |smallest entry of the caller's [rows][L] table| (abs(compositionBias) is 0 here).
int b200h_build_profile_pssm(const int8_t *pssm, int rows, int L, int A, int8_t *out) {
    if (pssm == nullptr || out == nullptr || rows <= 0 || rows > A || L <= 0) return -1;
    int lowest = 0;
    for (size_t i = 0; i < (size_t) rows * L; i++) {
        out[i] = pssm[i];
        if (pssm[i] < lowest) lowest = pssm[i];
    }
    for (size_t i = (size_t) rows * L; i < (size_t) A * L; i++) out[i] = 0;
    return lowest < 0 ? -lowest : lowest;
}

// The SSW profile bias from what Marv::scan / the gpuserver protocol carry: the encoded query, its [A][L] profile and (optionally) the
// substitution matrix.  A sequence query's profile is mat[a][q[j]] + cb[j] (ungappedprefilter.cpp:195-203), so profile - matrix column
// is the same rounded composition bias in every residue row, and the bias is ssw_init's |min(mat)| + |min(0, min cb)|
// (StripedSmithWaterman.cpp:1375-1406).  Anything else is a profile (HMM) query, whose bias is |min(0, lowest entry)| with the X row
// counted as 0 (the isProfile branch, :1388-1406; the caller's table is the alignment profile itself).
int b200h_ssw_bias_from_profile(const int16_t *mat, int A, const uint8_t *q, int L, const int8_t *profile) {
    if (profile == nullptr || L <= 0 || A <= 0) return 0;
    bool is_seq = mat != nullptr && q != nullptr;
    int comp = 0;
    for (int j = 0; j < L && is_seq; j++) {
        if (q[j] >= A) { is_seq = false; break; }
        const int d = (int) profile[j] - (int) mat[q[j]];
        for (int a = 1; a < A; a++)
            if ((int) profile[(size_t) a * L + j] - (int) mat[(size_t) a * A + q[j]] != d) { is_seq = false; break; }
        comp = std::min(comp, d);
    }
    if (is_seq) {
        int bias = 0;
        for (int i = 0; i < A * A; i++) bias = std::min(bias, (int) (int8_t) mat[i]);
        return std::abs(bias) + std::abs(comp);
    }
    int lowest = 0;
    for (size_t i = 0; i < (size_t) (A - 1) * L; i++) lowest = std::min(lowest, (int) profile[i]);
    return -lowest;
}

}  // extern "C"
