/* include/b200_host.h -- host-side (CPU, no CUDA) pieces of the drop-in: the float/rounding logic that turns a query
 * into the int8 profiles the device entry points of b200_align.h consume.  The reference keeps exactly this on the host
 * (SURVEY.md fact 6); bit-exactness of the rounding is part of the contract, so it lives here once, in C.
 *
 *   b200h_comp_bias        SubstitutionMatrix::calcLocalAaBiasCorrection   src/commons/SubstitutionMatrix.cpp:79-109
 *   b200h_round_bias_ssw   ssw_init rounding                               src/alignment/StripedSmithWaterman.cpp:1379
 *   b200h_round_bias_diag  createProfile rounding (bias/4)                 src/prefiltering/UngappedAlignment.cpp:395-400
 *   b200h_ssw_bias         profile bias constant                           src/alignment/StripedSmithWaterman.cpp:1375-1406
 *   b200h_build_profile    profile_word_linear / createProfile contents    StripedSmithWaterman.cpp:1434-1439,
 *                                                                          UngappedAlignment.cpp:412-420
 *   b200h_build_profile_pssm  the HMM_PROFILE branches of both             StripedSmithWaterman.cpp:1388-1406,
 *                                                                          UngappedAlignment.cpp:405-411
 */
#ifndef B200_HOST_H
#define B200_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* mat: A*A int16 row-major (BaseMatrix::subMatrix), pback: A doubles (BaseMatrix::pBack) */
void b200h_comp_bias(const int16_t *mat, const double *pback, int A, const uint8_t *seq, int L, float scale, float *out);
void b200h_round_bias_ssw(const float *in, int L, int8_t *out);
void b200h_round_bias_diag(const float *in, int L, int8_t *out);
int b200h_ssw_bias(const int16_t *mat, int A, const int8_t *cb, int L, int cb_enabled);
/* out[a*L + j] = (target_major ? mat[a][q[j]] : mat[q[j]][a]) + cb[j]; returns 0, or -4 (B200_ERR_RANGE) if a value
 * leaves int8.  target_major=1 is the gapped / scan profile, 0 the per-diagonal scorer's. */
int b200h_build_profile(const int16_t *mat, int A, const uint8_t *q, int L, const int8_t *cb, int target_major, int8_t *out);
/* Profile (PSSM) query: pssm = Sequence::getAlignmentProfile(), [rows][L] int8 with rows = Sequence::PROFILE_AA_SIZE (20).
 * out[A][L]: the first `rows` residue rows copied, the remaining ones (X: "neutral state", score 0) zero -- the layout every
 * device entry point takes (per-diagonal scorer included: its [pos][A+... ] table holds the same values).  Returns the SSW
 * profile bias |min(0, min pssm)| (no composition bias in the profile branch), or -1 (B200_ERR_ARG) on bad sizes. */
int b200h_build_profile_pssm(const int8_t *pssm, int rows, int L, int A, int8_t *out);

/* SSW profile bias recovered from (encoded query, [A][L] profile, matrix) -- what Marv::scan (lib/libmarv/src/marv.h:47) and the
 * gpuserver protocol (src/commons/GpuUtil.h) carry.  mat may be NULL (then, or when profile - matrix is not one composition bias per
 * column, the profile-query rule applies: |min(0, lowest entry of the first A-1 rows)|). */
int b200h_ssw_bias_from_profile(const int16_t *mat, int A, const uint8_t *q, int L, const int8_t *profile);

#ifdef __cplusplus
}
#endif
#endif
