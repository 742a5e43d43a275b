/* include/b200_db.h -- the on-disk formats either side of the hot path, host side only:
 *
 *   DB triple     <name> (entries, each followed by one NUL byte), <name>.index (text lines "key \t offset \t length \n", length
 *                 counts the NUL; kept sorted by key), <name>.dbtype (int32, Parameters::DBTYPE_*)
 *                 -- DBReader / DBWriter, src/commons/DBReader.cpp, DBWriter.cpp:401-420,483-494,654-667
 *   sequence DB   entry = residues as ASCII + '\n' (+ NUL): sequence length = index length - 2 (DBReader::getSeqLen)
 *   letters       ASCII -> numeric residue codes as Sequence::mapSequence applies them through BaseMatrix::aa2num
 *                 (SubstitutionMatrix::setupLetterMapping, SubstitutionMatrix.cpp:257-298; NucleotideMatrix.cpp:17-62)
 *
 * and b200_align_db: the `align` module over DB files -- query DB + target DB + prefilter result DB -> alignment result DB --
 * built on b200_align_batch (include/b200_alignment.h).  Single-part or ".0 .1 ..." split data files, plain or zstd-compressed
 * (dbtype bit 31, DBReader::getDataCompressed, src/commons/DBReader.cpp:575-607: inflated at open through a run-time dlopen of
 * libzstd.so.1, b200h_db_type reports the plain type); no lookup / source / header files (not needed by the path).  Plain C ABI, status codes, no exceptions.
 */
#ifndef B200_DB_H
#define B200_DB_H

#include <stddef.h>
#include <stdint.h>

#include "b200_alignment.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Parameters::DBTYPE_* (src/commons/Parameters.h) */
enum { B200_DBTYPE_AMINO_ACIDS = 0, B200_DBTYPE_NUCLEOTIDES = 1, B200_DBTYPE_HMM_PROFILE = 2, B200_DBTYPE_ALIGNMENT_RES = 5,
       B200_DBTYPE_PREFILTER_RES = 7 };

typedef struct b200h_db b200h_db;
/* Opens <data_path>, <data_path>.index and (if present) <data_path>.dbtype.  Entries are addressed by id = rank of the key
 * (DBReader's order).  Returns B200_OK or B200_ERR_ARG (missing / malformed files; b200h_db_last_error() has the text). */
int b200h_db_open(const char *data_path, b200h_db **out);
void b200h_db_close(b200h_db *db);
uint64_t b200h_db_size(const b200h_db *db);
int b200h_db_type(const b200h_db *db);                          /* -1 when there is no .dbtype file */
uint32_t b200h_db_key(const b200h_db *db, uint64_t id);
const char *b200h_db_data(const b200h_db *db, uint64_t id);    /* the entry, NUL-terminated */
uint64_t b200h_db_entry_len(const b200h_db *db, uint64_t id);  /* index length: payload + NUL */
int64_t b200h_db_id(const b200h_db *db, uint32_t key);         /* DBReader::getId; -1 if the key is absent */
const char *b200h_db_last_error(void);

typedef struct b200h_dbw b200h_dbw;
int b200h_dbw_open(const char *data_path, int dbtype, b200h_dbw **out);
int b200h_dbw_write(b200h_dbw *w, uint32_t key, const char *data, uint64_t len);   /* payload; the NUL is appended */
int b200h_dbw_close(b200h_dbw *w);                              /* index sorted by key + dbtype file; frees w */

/* BaseMatrix::aa2num after setupLetterMapping.  num2aa: the matrix alphabet in numeric order (A letters, 'X' last). */
void b200h_aa2num_table(const char *num2aa, int A, int nucleotide, uint8_t table[256]);

/* `mmseqs align` on DB files (amino-acid sequence DBs): loads target_db into the context, walks prefilter_db in buckets of
 * bucket_queries queries, writes alignment_db.  add_backtrace / compress as --add-backtrace / -a in the reference.
 * n_alignments / n_records may be NULL. */
int b200_align_db(b200_ctx *ctx, const char *query_db, const char *target_db, const char *prefilter_db, const char *alignment_db,
                  const int16_t *sub_matrix, const double *p_back, const char *num2aa, int alphabet, const b200_align_params *params,
                  const b200_evalue_params *evalue /* NULL: blosum62 11/1 defaults with the target DB's residue count */,
                  int add_backtrace, uint32_t bucket_queries, uint64_t *n_alignments, uint64_t *n_records);

/* `mmseqs ungappedprefilter` on DB files (amino-acid sequence queries, PREF_MODE_UNGAPPED: runFilterOnCpu / runFilterOnGpu,
 * src/prefiltering/ungappedprefilter.cpp:346-482, :41-343): every query against every target with the saturating ungapped scorer,
 * hits with score > min_diag_score, ordered by (score desc, target key asc), truncated to max_res_list_len, written as prefilter
 * records "key \t score \t 0".  Soft-masked (lowercase) target residues score as X, as in runFilterOnCpu (:401-404).  Not applied:
 * identity inclusion (same query/target DB keeps its self hit only through its score) and the canBeCovered pre-check (:409-411), which is
 * a no-op at the default -c 0.  n_hits may be NULL. */
int b200_prefilter_db(b200_ctx *ctx, const char *query_db, const char *target_db, const char *prefilter_db, const int16_t *sub_matrix,
                      const double *p_back, const char *num2aa, int alphabet, int comp_bias, float comp_bias_scale, int min_diag_score,
                      uint32_t max_res_list_len, uint32_t bucket_queries, uint64_t *n_hits);

/* `mmseqs rescorediagonal` on DB files (amino-acid DBs; src/alignment/rescorediagonal.cpp:45-396): every prefilter hit (target key,
 * diagonal) of every query is rescored on its diagonal by DistanceCalculator::computeUngappedAlignment -- on the device,
 * b200_rescore_diagonal -- and turned into the reference's records: rescore_mode 0 (HAMMING) "key \t 100*seqId \t diagonal", 1
 * (SUBSTITUTION) "key \t bitScore \t diagonal", 2-4 (ALIGNMENT / END_TO_END / WINDOW_QUALITY) alignment records with "<alnLen>M" as
 * backtrace; hits kept when identity, or alnLen, coverage, seqId and E-value pass (:318-327).  Not taken over: --wrapped-scoring and
 * reverse-strand prefilter DBs (nucleotide), --filter-hits (its precision library).  n_hits / n_records may be NULL. */
typedef struct b200_rescore_params {
    int rescore_mode;              /* --rescore-mode, Parameters::RESCORE_MODE_* 0..4 */
    double eval_thr;               /* -e */
    float cov_thr;                 /* -c */
    int cov_mode;                  /* --cov-mode */
    float seq_id_thr;              /* --min-seq-id */
    int aln_len_thr;               /* --min-aln-len */
    int seq_id_mode;               /* --seq-id-mode */
    int include_identity;          /* --add-self-matches */
    int add_backtrace;             /* -a */
    int sort_results;              /* --sort-results */
} b200_rescore_params;
int b200_rescorediagonal_db(b200_ctx *ctx, const char *query_db, const char *target_db, const char *prefilter_db, const char *out_db,
                            const int16_t *sub_matrix, const char *num2aa, int alphabet, const b200_rescore_params *params,
                            const b200_evalue_params *evalue /* NULL: the ungapped blosum62 set with the target DB's residue count */,
                            uint32_t bucket_queries, uint64_t *n_hits, uint64_t *n_records);
/* The host half of the module with the per-hit scorer abstracted (arguments of b200_rescore_diagonal plus the target sequences as the
 * module staged them: target i = target_data[target_offsets[i] .. target_offsets[i+1])).  Called once with mode = -1 and no hits
 * ("these are the targets"), then once per bucket of prefilter entries.  b200_rescorediagonal_db passes the device scorer; a test can
 * pass a checker to exercise the host half alone. */
typedef int (*b200_rescore_fn)(void *user, const char *query_data, const uint64_t *query_offsets, int n_queries, const uint64_t *hit_offsets,
                               const uint32_t *ids, const uint16_t *diagonals, const char *target_data, const uint64_t *target_offsets,
                               uint64_t n_targets, const int8_t *ascii_matrix, int mode, b200_rescore *out);
int b200h_rescorediagonal_db_with(b200_rescore_fn scorer, void *user, const char *query_db, const char *target_db, const char *prefilter_db,
                                  const char *out_db, const int16_t *sub_matrix, const char *num2aa, int alphabet, const b200_rescore_params *params,
                                  const b200_evalue_params *evalue, uint32_t bucket_queries, uint64_t *n_hits, uint64_t *n_records);
/* SubstitutionMatrix::createAsciiSubMat (SubstitutionMatrix.h:55-72): out [123][123] int8, indexed by the bytes 0..'z' */
void b200h_ascii_matrix(const int16_t *sub_matrix, const char *num2aa, int alphabet, int nucleotide, int8_t *out);
const char *b200h_rescore_module_last_error(void);

/* ---- padded GPU sequence DB: the writer (`mmseqs makepaddedseqdb`, src/util/makepaddedseqdb.cpp:14-153) -------------------------------
 * Reads the amino-acid DB <src_db> (+ <src_db>_h, plain or compressed) and writes <dst_db>, .index, .dbtype (extended flag
 * DBTYPE_EXTENDED_GPU), <dst_db>_h (+ .index, .dbtype) and, with write_lookup, <dst_db>.lookup and a copy of <src_db>.source: entries
 * ordered by ascending length and renumbered 0..n-1, residues as numeric codes (aa2num, Sequence::mapSequence), +32 on the codes the
 * masker would replace by X (mask_mode 1: tantan repeats with posterior >= mask_prob, runs longer than mask_n_repeats, lower-case
 * letters if mask_lower_case) or, with mask_mode 0, on lower-case letters; every entry padded with code 20 to a multiple of 4 bytes.
 * This is the DB b200_db_load_padded / Marv::loadDb read.  likelihood_ratio [alphabet * alphabet] = Qxy / (Px Py), the reference's
 * ProbabilityMatrix (src/commons/BaseMatrix.h:83-96); may be NULL with mask_mode 0.  The reference's defaults: mask_mode 1,
 * mask_prob 0.9, mask_lower_case 0, mask_n_repeats 0, write_lookup 1.
 * The masker's doubles are accumulated in the order of the reference's AVX2 build (what the released x86-64 binaries are); its SSE /
 * NEON / scalar builds sum in another order and may differ from it -- and from this -- in the last bit of a posterior. */
int b200h_make_padded_db(const char *src_db, const char *dst_db, const uint8_t aa2num[256], int alphabet, const double *likelihood_ratio,
                         int mask_mode, double mask_prob, int mask_lower_case, int mask_n_repeats, int write_lookup, int threads);
/* tantan::getProbabilities as Masker.cpp:22-32 calls it (maxRepeatOffset 50, repeatProb 0.005, repeatEndProb 0.05, decay 0.9, no
 * indel states): posterior probability of "inside a repeat" per residue.  seq: numeric codes < alphabet. */
int b200h_tantan_probabilities(const uint8_t *seq, int L, int alphabet, const double *likelihood_ratio, float *probs);
/* Masker::maskSequence on numeric codes (text = the entry's letters, needed for mask_lower_case; may be NULL otherwise): masked
 * residues become code alphabet-1 (X).  Returns the number of masked residues, -1 on bad arguments. */
int b200h_mask_sequence(uint8_t *seq, const char *text, int L, int alphabet, const double *likelihood_ratio, int mask_tantan, double mask_prob,
                        int mask_lower_case, int mask_n_repeats);
const char *b200h_paddeddb_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
