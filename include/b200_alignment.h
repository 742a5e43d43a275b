/* include/b200_alignment.h -- the caller of the gapped hot path, batched: what Alignment::run does per query
 * (src/alignment/Alignment.cpp:312-420) for a whole batch of queries in one call, and the record formats either side of it.
 * SURVEY.md section 8(f) rows 1 and 2.  Plain C ABI, no exceptions, caller-allocated outputs.
 *
 *   prefilter record   key \t score \t diagonal \n              QueryMatcher::prefilterHitToBuffer / parsePrefilterHit
 *                                                               (src/prefiltering/QueryMatcher.h:78-99,120-132)
 *   alignment record   key \t bitscore \t seqId \t evalue \t qStart \t qEnd \t qLen \t dbStart \t dbEnd \t dbLen [\t cigar] \n
 *                                                               Matcher::resultToBuffer (src/alignment/Matcher.cpp:282-329)
 *   E-values           EvalueComputation (src/alignment/EvalueComputation.h:22-44) over the ALP finite-size-corrected
 *                      area (lib/alp/sls_pvalues.cpp:366-531, lib/alp/sls_alignment_evaluer.hpp:154-162)
 *
 * Scope: amino-acid sequence queries (SEQ_SEQ), Matcher::SCORE_ONLY / SCORE_COV / SCORE_COV_SEQID, no realign, no
 * alternative alignments, no wrapped scoring, correlationScoreWeight 0 -- the defaults of `mmseqs align` / easy-search.
 */
#ifndef B200_ALIGNMENT_H
#define B200_ALIGNMENT_H

#include <stddef.h>
#include <stdint.h>

#include "b200_align.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- E-value statistics ------------------------------------------------------------------------------------------ */
/* Gumbel parameters in ALP's naming (ALP_set_of_parameters); db_residues = EvalueComputation::dbResCount. */
typedef struct b200_evalue_params {
    double lambda, K;
    double a_J, b_J, a_I, b_I;
    double alpha_J, beta_J, alpha_I, beta_I;
    double sigma, tau;
    uint64_t db_residues;
} b200_evalue_params;

/* The hard-coded parameter sets of EvalueComputation.h:56-81.  matrix: "blosum62.out" or "nucleotide.out".
 * Returns B200_OK, or B200_ERR_ARG when the reference would estimate parameters with ALP at start-up (not done here:
 * the caller passes evaluer.parameters() in that case). */
int b200h_evalue_defaults(const char *matrix, int gap_open, int gap_extend, int gapped, uint64_t db_residues,
                          b200_evalue_params *out);
/* EvalueComputation::computeEvalue / computeBitScore (EvalueComputation.h:22-44) */
double b200h_evalue(const b200_evalue_params *p, double score, double query_len);
double b200h_bit_score(const b200_evalue_params *p, double score);

/* ---- records ----------------------------------------------------------------------------------------------------- */
typedef struct b200_pref_hit {   /* hit_t, QueryMatcher.h:33-49 */
    uint32_t seq_id;             /* database KEY as written in the record */
    int32_t pref_score;
    uint16_t diagonal;
    uint16_t pad_;
} b200_pref_hit;

/* parsePrefilterHits: every line of a NUL-terminated entry; returns the number of records (stores at most cap). */
size_t b200h_parse_prefilter_hits(const char *data, b200_pref_hit *out, size_t cap);
/* prefilterHitToBuffer: writes "id\tscore\tdiag\n\0", returns the length without the NUL; buf needs >= 40 bytes. */
size_t b200h_prefilter_hit_to_buffer(char *buf, const b200_pref_hit *h);

typedef struct b200_result {     /* Matcher::result_t, Matcher.h:32-95, without the ORF fields */
    uint32_t db_key;
    int32_t score;               /* bit score, rounded */
    float qcov, dbcov, seq_id;
    double eval;
    uint32_t aln_length;
    int32_t q_start, q_end, q_len;
    int32_t db_start, db_end, db_len;
    uint64_t bt_off;             /* offset of the M/I/D backtrace string in the batch's backtrace pool */
    uint32_t bt_len;             /* 0 when no backtrace was computed */
    uint32_t pad_;
} b200_result;

/* resultToBuffer(buffer, result, addBacktrace, compress, addOrfPosition=false); backtrace may be NULL when
 * add_backtrace == 0.  buf needs >= 256 + 2 * bt_len bytes.  Returns the record length without the NUL. */
size_t b200h_result_to_buffer(char *buf, const b200_result *r, const char *backtrace, int add_backtrace, int compress);
/* Matcher::compressAlignment (Matcher.cpp:168-187): "MMMIID" -> "3M2I1D".  Returns the length written (no NUL). */
size_t b200h_compress_alignment(const char *bt, size_t bt_len, char *out);

/* ---- Alignment::run for a batch of queries ----------------------------------------------------------------------- */
typedef struct b200_align_params {
    int gap_open, gap_extend;      /* par.gapOpen / par.gapExtend (aa: 11 / 1) */
    int sw_mode;                   /* Matcher::SCORE_ONLY 0, SCORE_COV 1, SCORE_COV_SEQID 2 (Alignment::initSWMode) */
    double eval_thr;               /* -e */
    float cov_thr;                 /* -c */
    int cov_mode;                  /* --cov-mode, Parameters::COV_MODE_* */
    float seq_id_thr;              /* --min-seq-id */
    int aln_len_thr;               /* --min-aln-len */
    int seq_id_mode;               /* --seq-id-mode, Parameters::SEQ_ID_* */
    uint32_t max_accept;           /* --max-accept (UINT32_MAX: none) */
    uint32_t max_rejected;         /* --max-rejected (UINT32_MAX: none) */
    int comp_bias;                 /* --comp-bias-corr */
    float comp_bias_scale;         /* --comp-bias-corr-scale */
    int include_identity;          /* isIdentity rule of Alignment.cpp:379: query key == target key && (includeIdentity || sameQTDB) */
} b200_align_params;

/* One call = the per-query loop of Alignment::run (Alignment.cpp:312-420) for n_queries queries:
 *   canBeCovered pre-check, ssw_align of every remaining (query, prefilter hit) -- score / end / E-value + coverage gate /
 *   start / backtrace on the device --, getSWResult's result assembly (Matcher.cpp:62-145), checkCriteria,
 *   --max-accept / --max-rejected in list order, compareHits ordering.
 * Inputs
 *   sub_matrix   A*A int16 (BaseMatrix::subMatrix), p_back A doubles, alphabet A (must equal the loaded DB's)
 *   query_residues / query_offsets[n_queries+1]   numeric query sequences, concatenated
 *   query_keys[n_queries]                         database keys of the queries (identity rule only; may be NULL)
 *   hit_offsets[n_queries+1], hit_targets[n_hits] prefilter lists in list order; targets are DB-local ids of b200_db_load
 *   target_keys[n_db]                             database key of every DB-local id (NULL: key == id)
 * Outputs
 *   results[n_hits]      the accepted results of query i at results[hit_offsets[i] .. hit_offsets[i] + n_results[i]), sorted
 *   n_results[n_queries]
 *   bt_pool / bt_cap     backtrace strings (only sw_mode 2); B200_ERR_RANGE if bt_cap is too small
 *   n_alignments (may be NULL)   number of ssw_align calls the reference would have made
 */
int b200_align_batch(b200_ctx *ctx, const int16_t *sub_matrix, const double *p_back, int alphabet,
                     const uint8_t *query_residues, const uint64_t *query_offsets, const uint32_t *query_keys,
                     uint32_t n_queries, const uint64_t *hit_offsets, const uint32_t *hit_targets,
                     const uint32_t *target_keys, const b200_align_params *params, const b200_evalue_params *evalue,
                     b200_result *results, uint32_t *n_results, char *bt_pool, uint64_t bt_cap, uint64_t *n_alignments);

/* The same loop for nucleotide searches: Matcher::getSWResult takes the BandedNucleotideAligner branch (Matcher.cpp:72-78),
 * alignmentMode is forced to SCORE_COV_SEQID, and every hit carries its prefilter diagonal and strand.
 *   hit_diagonals[n_hits]   (short) hit_t::diagonal of the prefilter record
 *   hit_reverse[n_hits]     may be NULL; non-zero = reverse-strand hit (reversePrefilterResult && prefScore < 0, Alignment.cpp:360):
 *                           the read is reverse-complemented (BandedNucleotideAligner::initQuery, :60-66) and the result carries
 *                           dbStartPos/dbEndPos swapped as Matcher.cpp:131 does
 * The DB must be loaded with alphabet 5; params->gap_open/gap_extend are 5/2 by default in the reference, params->sw_mode and
 * comp_bias are ignored; zdrop = par.zdrop (40).  Not covered: wrapped scoring. */
int b200_align_batch_nucl(b200_ctx *ctx, const uint8_t *query_residues, const uint64_t *query_offsets, const uint32_t *query_keys,
                          uint32_t n_queries, const uint64_t *hit_offsets, const uint32_t *hit_targets,
                          const int16_t *hit_diagonals, const uint8_t *hit_reverse, const uint32_t *target_keys,
                          const b200_align_params *params, int zdrop, const b200_evalue_params *evalue, b200_result *results,
                          uint32_t *n_results, char *bt_pool, uint64_t bt_cap, uint64_t *n_alignments);

#ifdef __cplusplus
}
#endif
#endif
