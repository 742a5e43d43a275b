// mmseqs2_b200/csrc/b200_alignment.cpp -- the batched caller of the gapped hot path and the records either side of it
// (include/b200_alignment.h).  Host code: double/float statistics, gates, ordering and text
// formatting stay on the CPU exactly as in the reference; every DP runs on the device through the C ABI of b200_align.h.
#include "b200_alignment.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "b200_host.h"
#include "b200_internal.h"

// =====================================================================================================================
// E-value statistics
// =====================================================================================================================
namespace {

struct NamedSet { const char *matrix; int go, ge; bool gapped; double v[12]; };
// EvalueComputation.h:56-81 -- order of v[]: lambda, K, a1, b1, a2, b2, alpha1, beta1, alpha2, beta2, sigma, tau
// (Sls::AlignmentEvaluerParameters, lib/alp/sls_basic.hpp:70-84; index 1 = J, index 2 = I in initParameters,
// sls_alignment_evaluer.cpp:669-724)
const NamedSet kSets[] = {
    {"nucleotide.out", 7, 1, true, {1.0960171987681839, 0.33538787507026158, 2.0290734315292083, -0.46514786408422282,
                                     2.0290734315292083, -0.46514786408422282, 5.0543294182155085, 15.130999712620039,
                                     5.0543294182155085, 15.130999712620039, 5.0543962679167036, 15.129930117400917}},
    {"nucleotide.out", 5, 2, true, {0.62092274139392822363, 0.35177597988201619872, 0.74528059208662511548,
                                     -0.71027220445456995535, 0.74528059208662511548, -0.71027220445456995535,
                                     1.0135243407674570104, -2.5226486486783059604, 1.0135243407674570104,
                                     -2.5226486486783059604, 1.0031949332622873694, -2.3780369436059309862}},
    {"blosum62.out", 11, 1, true, {0.27359865037097330642, 0.044620920658722244834, 1.5938724404943873658,
                                    -19.959867650284412122, 1.5938724404943873658, -19.959867650284412122,
                                    30.455610143099914211, -622.28684628915891608, 30.455610143099914211,
                                    -622.28684628915891608, 29.602444874818868215, -601.81087985041381216}},
    {"blosum62.out", 0, 0, false, {0.3207378152604042354, 0.13904657125294345166, 0.76221128839920349041, 0,
                                    0.76221128839920349041, 0, 4.5269915477182944841, 0, 4.5269915477182944841, 0,
                                    4.5269915477182944841, 0}},
};

// standard normal CDF as ALP evaluates it (sls_basic.hpp:195-198)
inline double normal_cdf(double x) { return 0.5 * erfc(-sqrt(0.5) * x); }

// One side of the finite-size correction (sls_pvalues.cpp:423-456 for the database side "I", :459-490 for the query
// side "J"): expected usable length L - (a*y + b) under a normal with variance max(thr, alpha*y + beta).
struct Side { double p, cdf; };
inline Side fsc_side(double len, double y, double a, double b, double alpha, double beta, double var_floor) {
    const double pi = 3.1415926535897932384626433832795;
    const double inv_sqrt_2pi = 1 / sqrt(2.0 * pi);
    const double usable = len - (a * y + b);
    const double var = std::max(var_floor, alpha * y + beta);
    const double sd = sqrt(var);
    const double z = (sd == 0.0) ? 1e100 : usable / sd;
    Side s;
    s.cdf = normal_cdf(z);
    const double dens = -inv_sqrt_2pi * exp(-0.5 * z * z);
    s.p = usable * s.cdf - sd * dens;
    return s;
}

// AlignmentEvaluer::area(score, seqlen1 = query, seqlen2 = database) (sls_alignment_evaluer.cpp:989-1028) with
// compute_only_area and blast == false: p_I * p_J + c(y) * Phi_I * Phi_J  (sls_pvalues.cpp:495-505)
double fsc_area(const b200_evalue_params &p, double y, double query_len) {
    const double cut = 2.0;  // nat_cut_off_in_max, sls_pvalues.cpp:46,352-354
    const double floor_i = std::max(cut * p.alpha_I / p.lambda, 0.0);
    const double floor_j = std::max(cut * p.alpha_J / p.lambda, 0.0);
    const double floor_c = std::max(cut * p.sigma / p.lambda, 0.0);
    const Side si = fsc_side((double) p.db_residues, y, p.a_I, p.b_I, p.alpha_I, p.beta_I, floor_i);
    const Side sj = fsc_side(query_len, y, p.a_J, p.b_J, p.alpha_J, p.beta_J, floor_j);
    const double cov = std::max(floor_c, p.sigma * y + p.tau);
    const double both = si.cdf * sj.cdf;
    const double cov_term = cov * both;
    const double prod = si.p * sj.p;
    return prod + cov_term;
}

}  // namespace

int b200h_evalue_defaults(const char *matrix, int gap_open, int gap_extend, int gapped, uint64_t db_residues,
                          b200_evalue_params *out) {
    if (matrix == nullptr || out == nullptr) return B200_ERR_ARG;
    for (const NamedSet &s : kSets) {
        if (strcmp(s.matrix, matrix) != 0 || s.gapped != (gapped != 0)) continue;
        if (s.go != gap_open || s.ge != gap_extend) continue;
        out->lambda = s.v[0]; out->K = s.v[1];
        out->a_J = s.v[2]; out->b_J = s.v[3]; out->a_I = s.v[4]; out->b_I = s.v[5];
        out->alpha_J = s.v[6]; out->beta_J = s.v[7]; out->alpha_I = s.v[8]; out->beta_I = s.v[9];
        out->sigma = s.v[10]; out->tau = s.v[11];
        out->db_residues = db_residues;
        return B200_OK;
    }
    return B200_ERR_ARG;
}

double b200h_evalue(const b200_evalue_params *p, double score, double query_len) {
    const double per_area = p->K * exp(-p->lambda * score);   // evaluePerArea, sls_alignment_evaluer.hpp:154-157
    const double a = fsc_area(*p, score, query_len);
    return per_area * a;                                      // EvalueComputation::computeEvalue, :35-39
}

double b200h_bit_score(const b200_evalue_params *p, double score) {
    const double log_k = log(p->K);                           // EvalueComputation::logK, :157
    return (p->lambda * score - log_k) / log(2.0);            // bitScore(score, logK), sls_alignment_evaluer.hpp:159-162
}

// =====================================================================================================================
// records
// =====================================================================================================================
namespace {

inline const char *skip_ws(const char *d) { while (*d == ' ' || *d == '\t') d++; return d; }       // Util::skipWhitespace
inline const char *skip_word(const char *d) { while (*d != ' ' && *d != '\t' && *d != '\n' && *d != '\0') d++; return d; }
inline const char *skip_line(const char *d) { while (*d != '\n' && *d != '\0') d++; return *d == '\n' ? d + 1 : d; }

template <typename T> T parse_int(const char *s) {   // Util::fast_atoi (Util.h:131-146): no overflow check, stops at non-digit
    T val = 0;
    int sign = 1;
    if (std::numeric_limits<T>::is_signed && *s == '-') { sign = -1; s++; }
    while (*s >= '0' && *s <= '9') val = (T) (val * 10 + (*s++ - '0'));
    return (T) (sign * val);
}

// decimal text of v at p, returns the position after the last digit
char *put_u64(char *p, uint64_t v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char) ('0' + v % 10); v /= 10; } while (v != 0);
    while (n > 0) *p++ = tmp[--n];
    return p;
}
char *put_i32(char *p, int32_t v) {
    if (v < 0) { *p++ = '-'; return put_u64(p, (uint64_t) (-(int64_t) v)); }
    return put_u64(p, (uint64_t) v);
}

// Util::fastSeqIdToBuffer as it ends up in a record (Util.cpp:251-280 + the caller's "*(tmpBuff-1) = '\t'" in
// Matcher.cpp:288-289): the 1.0 branch returns a pointer AT its terminator instead of behind it, so the separator
// lands on the last digit and "1.000" reaches the record as "1.00".
char *put_seq_id(char *p, float seq_id) {
    if (seq_id == 1.0) { memcpy(p, "1.00", 4); return p + 4; }
    *p++ = '0'; *p++ = '.';
    if (seq_id < 0.10) *p++ = '0';
    if (seq_id < 0.01) *p++ = '0';
    return put_i32(p, (int) (seq_id * 1000));
}

}  // namespace

size_t b200h_parse_prefilter_hits(const char *data, b200_pref_hit *out, size_t cap) {
    size_t n = 0;
    if (data == nullptr) return 0;
    while (*data != '\0') {
        // three whitespace-separated words (getWordsOfLine); anything else is a format error in the reference
        const char *w0 = skip_ws(data);
        const char *w1 = skip_ws(skip_word(w0));
        const char *w2 = skip_ws(skip_word(w1));
        if (n < cap) {
            b200_pref_hit h;
            h.seq_id = parse_int<uint32_t>(w0);
            h.pref_score = parse_int<int>(w1);
            h.diagonal = (uint16_t) parse_int<short>(w2);
            h.pad_ = 0;
            out[n] = h;
        }
        n++;
        data = skip_line(data);
    }
    return n;
}

size_t b200h_prefilter_hit_to_buffer(char *buf, const b200_pref_hit *h) {
    char *p = put_u64(buf, h->seq_id);
    *p++ = '\t';
    p = put_i32(p, h->pref_score);
    *p++ = '\t';
    p = put_i32(p, (int32_t) (int16_t) h->diagonal);
    *p++ = '\n';
    *p = '\0';
    return (size_t) (p - buf);
}

size_t b200h_compress_alignment(const char *bt, size_t bt_len, char *out) {
    char *p = out;
    char state = 'M';
    size_t run = 0;
    for (size_t i = 0; i < bt_len; i++) {
        if (bt[i] != state) { p = put_u64(p, run); *p++ = state; state = bt[i]; run = 1; }
        else run++;
    }
    p = put_u64(p, run);
    *p++ = state;
    return (size_t) (p - out);
}

size_t b200h_result_to_buffer(char *buf, const b200_result *r, const char *backtrace, int add_backtrace, int compress) {
    char *p = put_u64(buf, r->db_key);
    *p++ = '\t';
    p = put_i32(p, r->score);
    *p++ = '\t';
    p = put_seq_id(p, r->seq_id);
    *p++ = '\t';
    p += snprintf(p, 32, "%.3E", r->eval);
    *p++ = '\t';
    p = put_i32(p, r->q_start); *p++ = '\t';
    p = put_i32(p, r->q_end); *p++ = '\t';
    p = put_i32(p, r->q_len); *p++ = '\t';
    p = put_i32(p, r->db_start); *p++ = '\t';
    p = put_i32(p, r->db_end); *p++ = '\t';
    p = put_i32(p, r->db_len);
    if (add_backtrace) {
        *p++ = '\t';
        if (compress) p += b200h_compress_alignment(backtrace, r->bt_len, p);
        else { memcpy(p, backtrace, r->bt_len); p += r->bt_len; }
    }
    *p++ = '\n';
    *p = '\0';
    return (size_t) (p - buf);
}

// =====================================================================================================================
// Alignment::run, batched
// =====================================================================================================================
namespace {

// SmithWaterman::computeCov (StripedSmithWaterman.cpp:1762-1764)
inline float compute_cov(unsigned int start, unsigned int end, unsigned int len) {
    return (std::min(len, std::max(start, end)) - std::min(start, end) + 1) / (float) len;
}

// Util::canBeCovered / hasCoverage (Util.cpp:542-576); modes Parameters::COV_MODE_* (Parameters.h:284-289)
bool can_be_covered(float thr, int mode, float ql, float tl) {
    switch (mode) {
        case 0: return (ql / tl >= thr) && (tl / ql >= thr);
        case 2: return (tl / ql) >= thr;
        case 1: return (ql / tl) >= thr;
        case 3: return ((tl / ql) >= thr) && (tl / ql) <= 1.0;
        case 4: return ((ql / tl) >= thr) && (ql / tl) <= 1.0;
        case 5: return (std::min(tl, ql) / std::max(tl, ql)) >= thr;
        default: return true;
    }
}
bool has_coverage(float thr, int mode, float qcov, float tcov) {
    switch (mode) {
        case 0: return (qcov >= thr) && (tcov >= thr);
        case 2: return qcov >= thr;
        case 1: return tcov >= thr;
        default: return true;
    }
}

// Matcher::estimateSeqIdByScorePerCol (Matcher.cpp:162-166): float quotient, double affine map, back to float
float estimate_seq_id(uint16_t score, unsigned int q_aln, unsigned int t_aln) {
    float est = (score / static_cast<float>(std::max(q_aln, t_aln))) * 0.1656 + 0.1141;
    est = std::min(est, 1.0f);
    return std::max(0.0f, est);
}

// Util::computeSeqId (Util.cpp:597-607)
float compute_seq_id(int mode, int ids, int ql, int tl, int aln_len) {
    switch (mode) {
        case 1: return static_cast<float>(ids) / static_cast<float>(std::min(ql, tl));
        case 2: return static_cast<float>(ids) / static_cast<float>(std::max(ql, tl));
        case 0: return static_cast<float>(ids) / static_cast<float>(aln_len);
    }
    return 0.0;
}

// body(first, last) over [0, n) on up to 16 host threads (the reference runs this part of Alignment::run under OpenMP too)
template <typename F> void parallel_ranges(size_t n, size_t min_per_thread, F body) {
    size_t nt = std::min<size_t>(std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency())), n / std::max<size_t>(1, min_per_thread));
    if (nt <= 1) { body((size_t) 0, n); return; }
    std::vector<std::thread> th;
    const size_t step = (n + nt - 1) / nt;
    for (size_t a = 0; a < n; a += step) th.emplace_back(body, a, std::min(n, a + step));
    for (std::thread &t : th) t.join();
}

struct HitState {        // s_align of one (query, hit) as ssw_align leaves it
    uint32_t score1 = 0;
    int32_t q_start = -1, q_end = 0, db_start = -1, db_end = -1;
    float qcov = 0, tcov = 0;
    double evalue = 0;
    uint32_t identical = 0;
    bool aligned = false;        // getSWResult was called (passed canBeCovered)
    bool defined = false;        // the reference's s_align fields are all initialised (dbEnd != -1)
    bool identity = false;
    std::string backtrace;
};

}  // namespace

int b200_align_batch(b200_ctx *ctx, const int16_t *sub_matrix, const double *p_back, int alphabet,
                     const uint8_t *query_residues, const uint64_t *query_offsets, const uint32_t *query_keys,
                     uint32_t n_queries, const uint64_t *hit_offsets, const uint32_t *hit_targets,
                     const uint32_t *target_keys, const b200_align_params *params, const b200_evalue_params *evalue,
                     b200_result *results, uint32_t *n_results, char *bt_pool, uint64_t bt_cap, uint64_t *n_alignments) {
    if (ctx == nullptr) return B200_ERR_ARG;
    if (sub_matrix == nullptr || p_back == nullptr || query_residues == nullptr || query_offsets == nullptr ||
        hit_offsets == nullptr || params == nullptr || evalue == nullptr || n_results == nullptr)
        return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch: NULL argument");
    const uint64_t n_hits = hit_offsets[n_queries];
    if (n_hits > 0 && (hit_targets == nullptr || results == nullptr))
        return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch: NULL hit list / result array");
    if (params->sw_mode < 0 || params->sw_mode > 2) return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch: sw_mode must be 0, 1 or 2");
    const uint64_t n_db = b200_db_num_seqs(ctx);
    if (n_db == 0) return b200_set_err(ctx, B200_ERR_NODB, "no target DB loaded");
    if (alphabet != ctx->alphabet) return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch: alphabet differs from the loaded DB's");
    for (uint64_t k = 0; k < n_hits; k++)
        if (hit_targets[k] >= n_db) return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch: target id out of range");
    const int A = alphabet;
    const int go = params->gap_open, ge = params->gap_extend;
    const int mode = params->sw_mode;
    static const bool trace = getenv("B200_TRACE") != nullptr;   // stderr phase times (development aid)
    typedef std::chrono::steady_clock Clock;
    const Clock::time_point t0 = Clock::now();
    Clock::time_point t_prof = t0, t_end = t0, t_start = t0, t_bt = t0;
    float bt_kernel_ms = 0;
    const int32_t *db_len = ctx->h_len.data();   // read-only after b200_db_load

    // ---- Matcher::initQuery -> ssw_init for every query (host float/int8 logic, b200_host.h) -------------------------
    std::vector<std::vector<int8_t>> profiles(n_queries);
    std::vector<b200_query> queries(n_queries);
    for (uint32_t qi = 0; qi < n_queries; qi++) {
        if (query_offsets[qi + 1] < query_offsets[qi] || query_offsets[qi + 1] - query_offsets[qi] > 65535)
            return b200_set_err(ctx, B200_ERR_RANGE, "b200_align_batch: query length outside [0, 65535]");
        if (query_offsets[qi + 1] == query_offsets[qi] && hit_offsets[qi + 1] != hit_offsets[qi])
            return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch: empty query with hits");
    }
    std::vector<int> prof_rc(n_queries, 0);
    parallel_ranges(n_queries, 16, [&](size_t qa, size_t qb) {
        std::vector<float> fbias;
        std::vector<int8_t> cb;
        for (size_t qi = qa; qi < qb; qi++) {
            const int L = (int) (query_offsets[qi + 1] - query_offsets[qi]);
            const uint8_t *seq = query_residues + query_offsets[qi];
            queries[qi].profile = nullptr; queries[qi].qlen = L; queries[qi].bias = 0;
            if (L == 0) continue;
            for (int j = 0; j < L; j++)
                if (seq[j] >= A) prof_rc[qi] = B200_ERR_ARG;
            if (prof_rc[qi] != 0) continue;
            cb.assign((size_t) L, 0);
            if (params->comp_bias) {
                fbias.resize((size_t) L);
                b200h_comp_bias(sub_matrix, p_back, A, seq, L, params->comp_bias_scale, fbias.data());
                b200h_round_bias_ssw(fbias.data(), L, cb.data());
            }
            profiles[qi].resize((size_t) A * L);
            if (b200h_build_profile(sub_matrix, A, seq, L, cb.data(), 1, profiles[qi].data()) != 0) { prof_rc[qi] = B200_ERR_RANGE; continue; }
            queries[qi].profile = profiles[qi].data();
            queries[qi].bias = b200h_ssw_bias(sub_matrix, A, cb.data(), L, params->comp_bias ? 1 : 0);
        }
    });
    for (uint32_t qi = 0; qi < n_queries; qi++) {
        if (prof_rc[qi] == B200_ERR_ARG) return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch: query residue code >= alphabet");
        if (prof_rc[qi] == B200_ERR_RANGE) return b200_set_err(ctx, B200_ERR_RANGE, "b200_align_batch: profile value outside int8");
    }
    // a query without hits never reaches initQuery; give the ABI a harmless one-residue stand-in
    static const int8_t kStub[64] = {0};
    for (uint32_t qi = 0; qi < n_queries; qi++)
        if (queries[qi].profile == nullptr) { queries[qi].profile = kStub; queries[qi].qlen = 1; queries[qi].bias = 0; }

    t_prof = Clock::now();
    // ---- the pairs the reference would align (Alignment.cpp:346-381) -------------------------------------------------
    std::vector<HitState> st(n_hits);
    std::vector<b200_pair> pairs;
    std::vector<uint64_t> pair_hit;
    pairs.reserve(n_hits); pair_hit.reserve(n_hits);
    for (uint32_t qi = 0; qi < n_queries; qi++) {
        const int L = (int) (query_offsets[qi + 1] - query_offsets[qi]);
        for (uint64_t k = hit_offsets[qi]; k < hit_offsets[qi + 1]; k++) {
            const uint32_t t = hit_targets[k];
            if (!can_be_covered(params->cov_thr, params->cov_mode, static_cast<float>(L), static_cast<float>(db_len[t]))) continue;
            st[k].aligned = true;
            const uint32_t tkey = target_keys ? target_keys[t] : t;
            if (params->include_identity && query_keys != nullptr && query_keys[qi] == tkey) { st[k].identity = true; continue; }
            b200_pair pr; pr.query = qi; pr.target = t;
            pairs.push_back(pr); pair_hit.push_back(k);
        }
    }

    // ---- alignScoreEndPos, split at the E-value gate: the score of every pair (packed score launch), then end positions only
    //      for the pairs whose E-value passes -- a hit that fails it is rejected by checkCriteria whatever its positions are
    //      (Alignment.cpp:549-566), and in a search most prefilter hits fail it
    const uint64_t np = pairs.size();
    std::vector<int32_t> scores(np, 0);
    if (np > 0) {
        int rc = b200_sw_score(ctx, queries.data(), (int) n_queries, pairs.data(), np, go, ge, scores.data());
        if (rc != B200_OK) return rc;
    }
    std::vector<b200_pair> epairs;
    std::vector<int32_t> escores;
    std::vector<uint64_t> ehit;
    // the E-value of every scored pair (ALP's finite-size-corrected area: a few exp/log per call) on all host threads, then the ordered
    // compaction of the pairs that pass
    std::vector<uint8_t> pass(np, 0);
    parallel_ranges(np, 4096, [&](size_t a, size_t b) {
        for (size_t i = a; i < b; i++) {
            HitState &h = st[pair_hit[i]];
            const int L = queries[pairs[i].query].qlen;
            h.score1 = (uint32_t) scores[i]; h.q_end = L - 1; h.db_end = -1;
            if (scores[i] <= 0) continue;          // "no residue could be aligned": the reference returns uninitialised fields
            h.defined = true;
            h.evalue = b200h_evalue(evalue, (double) h.score1, (double) L);
            pass[i] = h.evalue > params->eval_thr ? 0 : 1;
        }
    });
    for (uint64_t i = 0; i < np; i++)
        if (pass[i]) { epairs.push_back(pairs[i]); escores.push_back(scores[i]); ehit.push_back(pair_hit[i]); }
    std::vector<b200_sw_end> ends(epairs.size());
    if (!epairs.empty()) {
        int rc = b200_sw_endpos(ctx, queries.data(), (int) n_queries, epairs.data(), epairs.size(), go, ge, escores.data(), ends.data());
        if (rc != B200_OK) return rc;
    }
    t_end = Clock::now();
    // ---- the coverage half of ssw_align_private's gate (StripedSmithWaterman.cpp:846-863): host float math -----------------
    std::vector<b200_pair> sub;
    std::vector<b200_sw_end> sub_ends;
    std::vector<uint64_t> sub_hit;
    for (uint64_t i = 0; i < epairs.size(); i++) {
        HitState &h = st[ehit[i]];
        const int L = queries[epairs[i].query].qlen;
        h.q_end = ends[i].qend; h.db_end = ends[i].dbend;
        h.qcov = compute_cov(0, (unsigned) h.q_end, (unsigned) L);
        h.tcov = compute_cov(0, (unsigned) h.db_end, (unsigned) db_len[epairs[i].target]);
        const bool low_cov = !has_coverage(params->cov_thr, params->cov_mode, h.qcov, h.tcov);
        if (mode == 0 || low_cov) continue;
        sub.push_back(epairs[i]); sub_ends.push_back(ends[i]); sub_hit.push_back(ehit[i]);
    }
    // ---- alignStartPosBacktrace: reverse pass, then (mode 2) banded_sw + computerBacktrace for what still has coverage -----
    if (!sub.empty()) {
        std::vector<b200_sw_aln> aln(sub.size());
        int rc = b200_sw_startpos(ctx, queries.data(), (int) n_queries, sub.data(), sub.size(), go, ge, sub_ends.data(), aln.data());
        if (rc != B200_OK) return rc;
        t_start = Clock::now();
        std::vector<b200_pair> bt_pairs;
        std::vector<b200_sw_aln> bt_aln;
        std::vector<uint64_t> bt_hit;
        std::vector<uint64_t> coff(1, 0);
        for (size_t i = 0; i < sub.size(); i++) {
            HitState &h = st[sub_hit[i]];
            const int L = queries[sub[i].query].qlen;
            h.q_start = aln[i].qstart; h.db_start = aln[i].dbstart;
            h.qcov = compute_cov((unsigned) h.q_start, (unsigned) h.q_end, (unsigned) L);
            h.tcov = compute_cov((unsigned) h.db_start, (unsigned) h.db_end, (unsigned) db_len[sub[i].target]);
            const bool low_cov = !has_coverage(params->cov_thr, params->cov_mode, h.qcov, h.tcov);
            if (mode == 1 || low_cov) continue;
            bt_pairs.push_back(sub[i]); bt_aln.push_back(aln[i]); bt_hit.push_back(sub_hit[i]);
            coff.push_back(coff.back() + (uint64_t) (aln[i].qend - aln[i].qstart + 1) + (uint64_t) (aln[i].dbend - aln[i].dbstart + 1) + 2);
        }
        if (!bt_pairs.empty()) {
            std::vector<const uint8_t *> qseq(n_queries);
            for (uint32_t qi = 0; qi < n_queries; qi++) qseq[qi] = query_residues + query_offsets[qi];
            std::vector<b200_sw_bt> bt(bt_pairs.size());
            std::vector<uint32_t> cig_pool;
            std::vector<uint64_t> cig_base;
            rc = b200_sw_backtrace_impl(ctx, queries.data(), qseq.data(), (int) n_queries, bt_pairs.data(), bt_pairs.size(), go, ge,
                                        bt_aln.data(), bt.data(), nullptr, nullptr, &cig_pool, &cig_base);
            if (rc != B200_OK) return rc;
            bt_kernel_ms = b200_last_kernel_ms(ctx);
            for (size_t i = 0; i < bt_pairs.size(); i++)
                if (!bt[i].ok) return b200_set_err(ctx, B200_ERR_CUDA, "b200_align_batch: backtrace did not reach the alignment score");
            const uint32_t *cigp = cig_pool.data();
            parallel_ranges(bt_pairs.size(), 512, [&](size_t a, size_t b) {
                for (size_t i = a; i < b; i++) {
                    HitState &h = st[bt_hit[i]];
                    h.identical = (uint32_t) bt[i].identical;
                    h.backtrace.reserve((size_t) bt[i].bt_len);
                    for (int c = 0; c < bt[i].n_cigar; c++) {
                        const uint32_t op = cigp[cig_base[i] + c];
                        h.backtrace.append((size_t) (op >> 4), "MID"[op & 0xfu]);
                    }
                }
            });
        }
    }

    t_bt = Clock::now();
    // ---- getSWResult's assembly, checkCriteria, accept / reject walk, ordering (Matcher.cpp:84-141, Alignment.cpp:381-403) ----
    // Queries are independent: every host thread assembles, filters and orders the results of a range of queries; the
    // backtrace strings are laid into the caller's pool in query order afterwards (the reference runs this loop under OpenMP too).
    struct QueryOut { std::vector<b200_result> acc; std::vector<std::string> acc_bt; uint64_t n_aln = 0; int err = 0; };
    std::vector<QueryOut> qout(n_queries);
    parallel_ranges(n_queries, 4, [&](size_t q_first, size_t q_last) {
    for (uint32_t qi = (uint32_t) q_first; qi < (uint32_t) q_last; qi++) {
        const int L = (int) (query_offsets[qi + 1] - query_offsets[qi]);
        std::vector<b200_result> &acc = qout[qi].acc;
        std::vector<std::string> &acc_bt = qout[qi].acc_bt;
        uint64_t &n_aln = qout[qi].n_aln;
        uint32_t passed = 0, rejected = 0;
        for (uint64_t k = hit_offsets[qi]; k < hit_offsets[qi + 1] && passed < params->max_accept && rejected < params->max_rejected; k++) {
            HitState &h = st[k];
            if (!h.aligned) { rejected++; continue; }
            n_aln++;
            const uint32_t t = hit_targets[k];
            const int tl = db_len[t];
            if (h.identity) {          // SmithWaterman::scoreIdentical (StripedSmithWaterman.cpp:1770-1805)
                if (tl != L) { qout[qi].err = 1; break; }
                short s = 0;
                const uint8_t *ts = query_residues + query_offsets[qi];   // same key => same sequence as the query
                for (int pos = 0; pos < L; pos++) s = (short) (s + profiles[qi][(size_t) ts[pos] * L + pos]);
                h.score1 = (uint32_t) (int) s;
                h.q_start = mode == 0 ? -1 : 0; h.db_start = h.q_start;
                h.q_end = L - 1; h.db_end = L - 1;
                h.qcov = 1.0f; h.tcov = 1.0f;
                h.evalue = b200h_evalue(evalue, (double) h.score1, (double) L);
                h.identical = (uint32_t) L;
                h.backtrace.assign((size_t) L, 'M');
                h.defined = true;
            }
            if (!h.defined) { rejected++; continue; }   // reference: undefined E-value / coverage; a zero-score hit is never kept
            const unsigned int q_start = (unsigned int) h.q_start, db_start = (unsigned int) h.db_start;
            const unsigned int q_end = (unsigned int) h.q_end, db_end = (unsigned int) h.db_end;
            float qcov = 0.0, dbcov = 0.0, seq_id = 0.0;
            if (mode == 1 || mode == 2) { qcov = h.qcov; dbcov = h.tcov; }
            unsigned int aln_len = (unsigned int) (std::max(abs((int) q_end - (int) q_start), abs((int) db_end - (int) db_start)) + 1);
            if (mode == 2) {
                if (h.backtrace.size() > 0) aln_len = (unsigned int) h.backtrace.size();
                seq_id = compute_seq_id(params->seq_id_mode, (int) h.identical, L, tl, (int) aln_len);
            } else if (mode == 1) {
                const unsigned int qa = std::max(q_end - q_start, 1u), ta = std::max(db_end - db_start, 1u);
                seq_id = estimate_seq_id((uint16_t) h.score1, qa, ta);
            } else {
                const unsigned int qa = std::max(q_end, 1u), ta = std::max(db_end, 1u);
                seq_id = estimate_seq_id((uint16_t) h.score1, qa, ta);
            }
            b200_result r;
            memset(&r, 0, sizeof(r));
            r.db_key = target_keys ? target_keys[t] : t;
            r.score = static_cast<int>(b200h_bit_score(evalue, (double) h.score1) + 0.5);
            r.qcov = qcov; r.dbcov = dbcov; r.seq_id = seq_id; r.eval = h.evalue; r.aln_length = aln_len;
            r.q_start = (int) q_start; r.q_end = (int) q_end; r.q_len = L;
            r.db_start = (int) db_start; r.db_end = (int) db_end; r.db_len = tl;
            if (h.identity) { r.qcov = 1.0f; r.dbcov = 1.0f; r.seq_id = 1.0f; }
            // Alignment::checkCriteria (Alignment.cpp:548-567)
            const bool ok = h.identity || ((r.eval <= params->eval_thr) && (r.seq_id >= (double) params->seq_id_thr) &&
                                           has_coverage(params->cov_thr, params->cov_mode, r.qcov, r.dbcov) &&
                                           ((int) r.aln_length >= params->aln_len_thr));
            if (ok) {
                r.bt_len = (uint32_t) h.backtrace.size();
                r.bt_off = acc_bt.size();     // index for now; turned into a pool offset after sorting
                acc.push_back(r); acc_bt.push_back(std::move(h.backtrace));
                passed++; rejected = 0;
            } else {
                rejected++;
            }
        }
        // Matcher::compareHits (Matcher.h:161-172)
        if (acc.size() > 1)
            std::sort(acc.begin(), acc.end(), [](const b200_result &a, const b200_result &b) {
                if (a.eval != b.eval) return a.eval < b.eval;
                if (a.score != b.score) return a.score > b.score;
                if (a.db_len != b.db_len) return a.db_len < b.db_len;
                return a.db_key < b.db_key;
            });
    }
    });
    uint64_t n_aln = 0, bt_used = 0;
    std::vector<uint64_t> bt_base(n_queries + 1, 0);
    for (uint32_t qi = 0; qi < n_queries; qi++) {
        if (qout[qi].err) return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch: identity hit with a different length");
        n_aln += qout[qi].n_aln;
        uint64_t sz = 0;
        for (const b200_result &r : qout[qi].acc) sz += r.bt_len;
        bt_base[qi + 1] = bt_base[qi] + sz;
    }
    bt_used = bt_base[n_queries];
    if (bt_used > bt_cap || (bt_used > 0 && bt_pool == nullptr)) return b200_set_err(ctx, B200_ERR_RANGE, "b200_align_batch: backtrace pool too small");
    parallel_ranges(n_queries, 4, [&](size_t q_first, size_t q_last) {
        for (size_t qi = q_first; qi < q_last; qi++) {
            std::vector<b200_result> &acc = qout[qi].acc;
            uint64_t at = bt_base[qi];
            for (size_t i = 0; i < acc.size(); i++) {
                const std::string &bt = qout[qi].acc_bt[(size_t) acc[i].bt_off];
                if (!bt.empty()) memcpy(bt_pool + at, bt.data(), bt.size());
                acc[i].bt_off = at;
                at += bt.size();
                results[hit_offsets[qi] + i] = acc[i];
            }
            n_results[qi] = (uint32_t) acc.size();
        }
    });
    if (n_alignments != nullptr) *n_alignments = n_aln;
    if (trace) {
        auto ms = [](Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[b200 trace] align_batch: %llu hits, profiles %.1f ms, score+end %.1f ms, gate+start %.1f ms, backtrace %.1f ms "
                        "(kernels %.1f ms), assembly %.1f ms\n", (unsigned long long) n_hits, ms(t0, t_prof), ms(t_prof, t_end),
                ms(t_end, t_start), ms(t_start, t_bt), bt_kernel_ms, ms(t_bt, Clock::now()));
    }
    return B200_OK;
}

// ---- nucleotide searches: the BandedNucleotideAligner branch of getSWResult ---------------------------------------------------
int b200_align_batch_nucl(b200_ctx *ctx, const uint8_t *query_residues, const uint64_t *query_offsets, const uint32_t *query_keys,
                          uint32_t n_queries, const uint64_t *hit_offsets, const uint32_t *hit_targets,
                          const int16_t *hit_diagonals, const uint8_t *hit_reverse, const uint32_t *target_keys,
                          const b200_align_params *params, int zdrop, const b200_evalue_params *evalue, b200_result *results,
                          uint32_t *n_results, char *bt_pool, uint64_t bt_cap, uint64_t *n_alignments) {
    if (ctx == nullptr) return B200_ERR_ARG;
    if (query_residues == nullptr || query_offsets == nullptr || hit_offsets == nullptr || params == nullptr || evalue == nullptr ||
        n_results == nullptr)
        return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch_nucl: NULL argument");
    const uint64_t n_hits = hit_offsets[n_queries];
    if (n_hits > 0 && (hit_targets == nullptr || hit_diagonals == nullptr || results == nullptr))
        return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch_nucl: NULL hit list / result array");
    const uint64_t n_db = b200_db_num_seqs(ctx);
    if (n_db == 0) return b200_set_err(ctx, B200_ERR_NODB, "no target DB loaded");
    if (ctx->alphabet != 5) return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch_nucl: the DB must be loaded with alphabet 5");
    for (uint64_t k = 0; k < n_hits; k++)
        if (hit_targets[k] >= n_db) return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch_nucl: target id out of range");
    const int32_t *db_len = ctx->h_len.data();

    // reads as given, followed by the reverse complements of the reads that have reverse-strand hits
    // (NucleotideMatrix::reverseResidue, NucleotideMatrix.cpp:9-13: A<->T, C<->G, X stays; codes A,C,T,G,X = 0..4)
    static const uint8_t kComplement[5] = {2, 3, 0, 1, 4};
    std::vector<uint8_t> reads(query_residues, query_residues + query_offsets[n_queries]);
    std::vector<uint64_t> roff(query_offsets, query_offsets + n_queries + 1);
    std::vector<uint32_t> rc_index(n_queries, UINT32_MAX);
    for (uint32_t qi = 0; qi < n_queries; qi++) {
        const uint64_t L = query_offsets[qi + 1] - query_offsets[qi];
        if (L > 32767) return b200_set_err(ctx, B200_ERR_RANGE, "b200_align_batch_nucl: read longer than 32767");
        for (uint64_t j = 0; j < L; j++)
            if (query_residues[query_offsets[qi] + j] > 4) return b200_set_err(ctx, B200_ERR_ARG, "b200_align_batch_nucl: residue code > 4");
        bool any_rev = false;
        if (hit_reverse != nullptr)
            for (uint64_t k = hit_offsets[qi]; k < hit_offsets[qi + 1] && !any_rev; k++) any_rev = hit_reverse[k] != 0;
        if (!any_rev) continue;
        rc_index[qi] = (uint32_t) (roff.size() - 1);
        const uint8_t *src = query_residues + query_offsets[qi];
        for (uint64_t j = 0; j < L; j++) reads.push_back(kComplement[src[L - 1 - j]]);
        roff.push_back(reads.size());
    }

    // tasks of every hit that can be covered (Alignment.cpp:370-373)
    std::vector<b200_nucl_task> tasks;
    std::vector<uint64_t> task_hit;
    std::vector<uint64_t> coff(1, 0);
    std::vector<int64_t> hit_task(n_hits, -1);
    for (uint32_t qi = 0; qi < n_queries; qi++) {
        const int L = (int) (query_offsets[qi + 1] - query_offsets[qi]);
        for (uint64_t k = hit_offsets[qi]; k < hit_offsets[qi + 1]; k++) {
            const uint32_t t = hit_targets[k];
            if (!can_be_covered(params->cov_thr, params->cov_mode, static_cast<float>(L), static_cast<float>(db_len[t]))) continue;
            const bool rev = hit_reverse != nullptr && hit_reverse[k] != 0;
            b200_nucl_task tk;
            tk.query = rev ? rc_index[qi] : qi; tk.target = t; tk.diagonal = (uint16_t) hit_diagonals[k]; tk.reserved = 0;
            hit_task[k] = (int64_t) tasks.size();
            tasks.push_back(tk); task_hit.push_back(k);
            coff.push_back(coff.back() + 2 * (uint64_t) L + 72);
        }
    }
    std::vector<b200_nucl_aln> aln(tasks.size());
    std::vector<uint32_t> cig(coff.back() + 1);
    if (!tasks.empty()) {
        int rc = b200_nucl_align(ctx, reads.data(), roff.data(), (uint32_t) (roff.size() - 1), tasks.data(), tasks.size(), params->gap_open,
                                 params->gap_extend, zdrop, aln.data(), cig.data(), coff.data());
        if (rc != B200_OK) return rc;
    }

    // result assembly (BandedNucleotideAligner.cpp:214-260 + Matcher.cpp:84-141 with alignmentMode SCORE_COV_SEQID), criteria, order
    uint64_t n_aln = 0, bt_used = 0;
    std::vector<b200_result> acc;
    std::vector<std::string> acc_bt;
    std::string bt;
    for (uint32_t qi = 0; qi < n_queries; qi++) {
        const int L = (int) (query_offsets[qi + 1] - query_offsets[qi]);
        acc.clear(); acc_bt.clear();
        uint32_t passed = 0, rejected = 0;
        for (uint64_t k = hit_offsets[qi]; k < hit_offsets[qi + 1] && passed < params->max_accept && rejected < params->max_rejected; k++) {
            if (hit_task[k] < 0) { rejected++; continue; }
            n_aln++;
            const b200_nucl_aln &a = aln[(size_t) hit_task[k]];
            const uint32_t t = hit_targets[k];
            const int tl = db_len[t];
            const bool rev = hit_reverse != nullptr && hit_reverse[k] != 0;
            bt.clear();
            for (int c = 0; c < a.n_cigar; c++) {
                const uint32_t op = cig[coff[(size_t) hit_task[k]] + c];
                bt.append((size_t) (op >> 4), "MID"[op & 0xfu]);
            }
            const uint32_t score1 = (uint32_t) a.score;
            const float qcov = compute_cov((unsigned) a.qstart, (unsigned) a.qend, (unsigned) L);
            const float tcov = compute_cov((unsigned) a.dbstart, (unsigned) a.dbend, (unsigned) tl);
            unsigned int aln_len = (unsigned int) (std::max(abs(a.qend - a.qstart), abs(a.dbend - a.dbstart)) + 1);
            if (bt.size() > 0) aln_len = (unsigned int) bt.size();
            b200_result r;
            memset(&r, 0, sizeof(r));
            r.db_key = target_keys ? target_keys[t] : t;
            r.score = static_cast<int>(b200h_bit_score(evalue, (double) score1) + 0.5);
            r.qcov = qcov; r.dbcov = tcov;
            r.seq_id = compute_seq_id(params->seq_id_mode, a.identical, L, tl, (int) aln_len);
            r.eval = b200h_evalue(evalue, (double) score1, (double) L);
            r.aln_length = aln_len;
            r.q_start = a.qstart; r.q_end = a.qend; r.q_len = L;
            r.db_start = rev ? a.dbend : a.dbstart; r.db_end = rev ? a.dbstart : a.dbend; r.db_len = tl;
            const uint32_t tkey = r.db_key;
            const bool identity = params->include_identity && query_keys != nullptr && query_keys[qi] == tkey;
            if (identity) { r.qcov = 1.0f; r.dbcov = 1.0f; r.seq_id = 1.0f; }
            const bool ok = identity || ((r.eval <= params->eval_thr) && (r.seq_id >= (double) params->seq_id_thr) &&
                                         has_coverage(params->cov_thr, params->cov_mode, r.qcov, r.dbcov) &&
                                         ((int) r.aln_length >= params->aln_len_thr));
            if (ok) {
                r.bt_len = (uint32_t) bt.size();
                r.bt_off = acc_bt.size();
                acc.push_back(r); acc_bt.push_back(bt);
                passed++; rejected = 0;
            } else {
                rejected++;
            }
        }
        if (acc.size() > 1)
            std::sort(acc.begin(), acc.end(), [](const b200_result &x, const b200_result &y) {
                if (x.eval != y.eval) return x.eval < y.eval;
                if (x.score != y.score) return x.score > y.score;
                if (x.db_len != y.db_len) return x.db_len < y.db_len;
                return x.db_key < y.db_key;
            });
        for (size_t i = 0; i < acc.size(); i++) {
            const std::string &b = acc_bt[(size_t) acc[i].bt_off];
            if (bt_used + b.size() > bt_cap || (b.size() > 0 && bt_pool == nullptr))
                return b200_set_err(ctx, B200_ERR_RANGE, "b200_align_batch_nucl: backtrace pool too small");
            if (!b.empty()) memcpy(bt_pool + bt_used, b.data(), b.size());
            acc[i].bt_off = bt_used;
            bt_used += b.size();
            results[hit_offsets[qi] + i] = acc[i];
        }
        n_results[qi] = (uint32_t) acc.size();
    }
    if (n_alignments != nullptr) *n_alignments = n_aln;
    return B200_OK;
}

