// mmseqs2_b200/csrc/b200_rescore_module.cpp -- `mmseqs rescorediagonal` over DB files (src/alignment/rescorediagonal.cpp:45-396): query DB +
// target DB + prefilter DB -> rescored result DB.  The per-hit scorer (DistanceCalculator::computeUngappedAlignment, :231-236) runs on the
// device (b200_rescore_diagonal, b200_rescore.cu); what the reference does with its seven integers per hit -- sequence identity, E-value,
// bit score, coverage, the acceptance rule, record formats, ordering (:237-345) -- is host float math restated here.  Host code only.
#include "b200_db.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "b200_internal.h"

namespace {

// SmithWaterman::computeCov (StripedSmithWaterman.cpp:1762-1764)
inline float cov_of(unsigned int start, unsigned int end, unsigned int len) {
    return (std::min(len, std::max(start, end)) - std::min(start, end) + 1) / (float) len;
}
// Util::canBeCovered / hasCoverage (Util.cpp:542-576); modes Parameters::COV_MODE_* (Parameters.h:284-289)
bool coverable(float thr, int mode, float ql, float tl) {
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
bool covered(float thr, int mode, float qcov, float tcov) {
    switch (mode) {
        case 0: return (qcov >= thr) && (tcov >= thr);
        case 2: return qcov >= thr;
        case 1: return tcov >= thr;
        default: return true;
    }
}
// Util::computeSeqId (Util.cpp:597-607)
float seq_id_of(int mode, int ids, int ql, int tl, int aln_len) {
    switch (mode) {
        case 1: return static_cast<float>(ids) / static_cast<float>(std::min(ql, tl));
        case 2: return static_cast<float>(ids) / static_cast<float>(std::max(ql, tl));
        case 0: return static_cast<float>(ids) / static_cast<float>(aln_len);
    }
    return 0.0;
}

thread_local std::string g_rs_err;
int rs_fail(int code, const std::string &msg) { g_rs_err = msg; return code; }

struct DeviceScorer { b200_ctx *ctx; };
int device_scorer(void *user, const char *query_data, const uint64_t *query_offsets, int n_queries, const uint64_t *hit_offsets, const uint32_t *ids,
                  const uint16_t *diagonals, const char *, const uint64_t *, uint64_t, const int8_t *ascii_matrix, int mode, b200_rescore *out) {
    return b200_rescore_diagonal(static_cast<DeviceScorer *>(user)->ctx, query_data, query_offsets, n_queries, hit_offsets, ids, diagonals, ascii_matrix, mode, out);
}

struct Entry {                 // one prefilter entry = one query
    uint32_t key;
    int64_t qid;               // id in the query DB, -1 when the entry is empty
    int qlen;
    uint64_t hit_begin;        // into the bucket's kept-hit arrays
};

}  // namespace

extern "C" {

const char *b200h_rescore_module_last_error(void) { return g_rs_err.c_str(); }

void b200h_ascii_matrix(const int16_t *sub_matrix, const char *num2aa, int alphabet, int nucleotide, int8_t *out) {
    uint8_t a2n[256];
    b200h_aa2num_table(num2aa, alphabet, nucleotide, a2n);
    for (int i = 0; i < 123; i++)                                   // SubstitutionMatrix::createAsciiSubMat (SubstitutionMatrix.h:55-72): bytes 0..'z'
        for (int j = 0; j < 123; j++) out[i * 123 + j] = (int8_t) sub_matrix[(size_t) a2n[i] * alphabet + a2n[j]];
}

int b200h_rescorediagonal_db_with(b200_rescore_fn scorer, void *user, const char *query_db, const char *target_db, const char *prefilter_db,
                                  const char *out_db, const int16_t *sub_matrix, const char *num2aa, int alphabet, const b200_rescore_params *par,
                                  const b200_evalue_params *evalue, uint32_t bucket_queries, uint64_t *n_hits, uint64_t *n_records) {
    if (scorer == nullptr || query_db == nullptr || target_db == nullptr || prefilter_db == nullptr || out_db == nullptr || sub_matrix == nullptr ||
        num2aa == nullptr || par == nullptr)
        return rs_fail(B200_ERR_ARG, "b200_rescorediagonal_db: NULL argument");
    if (par->rescore_mode < 0 || par->rescore_mode > 4) return rs_fail(B200_ERR_ARG, "b200_rescorediagonal_db: rescore_mode outside 0..4");
    if (bucket_queries == 0) bucket_queries = 4096;
    g_rs_err.clear();
    const int mode = par->rescore_mode;
    const bool aln_mode = mode >= 2;                                 // ALIGNMENT, END_TO_END_ALIGNMENT, WINDOW_QUALITY_ALIGNMENT
    b200h_db *tdb = nullptr, *qdb = nullptr, *pdb = nullptr;
    b200h_dbw *out = nullptr;
    const bool same = std::string(query_db) == std::string(target_db);            // sameQTDB, rescorediagonal.cpp:59
    int rc = b200h_db_open(target_db, &tdb);
    if (rc == B200_OK) rc = same ? B200_OK : b200h_db_open(query_db, &qdb);
    if (rc == B200_OK && same) qdb = tdb;
    if (rc == B200_OK) rc = b200h_db_open(prefilter_db, &pdb);
    auto amino = [](const b200h_db *d) { const int t = b200h_db_type(d); return t == -1 || (t & 0xffff) == B200_DBTYPE_AMINO_ACIDS; };   // -1: no .dbtype file
    if (rc == B200_OK && (!amino(tdb) || !amino(qdb)))
        rc = rs_fail(B200_ERR_ARG, "b200_rescorediagonal_db: amino-acid sequence DBs expected (nucleotide / wrapped scoring keep the reference path)");
    else if (rc != B200_OK) rs_fail(rc, b200h_db_last_error());
    if (rc == B200_OK && b200h_dbw_open(out_db, aln_mode ? B200_DBTYPE_ALIGNMENT_RES : B200_DBTYPE_PREFILTER_RES, &out) != B200_OK)
        rc = rs_fail(B200_ERR_ARG, b200h_db_last_error());
    uint64_t total_hits = 0, total_rec = 0;
    if (rc == B200_OK) {
        int8_t ascii[123 * 123];
        b200h_ascii_matrix(sub_matrix, num2aa, alphabet, 0, ascii);
        // ---- target sequences as they lie in the DB, without "\n\0" (DBReader::getSeqLen = index length - 2) ----------------------
        const uint64_t nt = b200h_db_size(tdb);
        std::vector<uint64_t> toff(nt + 1, 0);
        for (uint64_t i = 0; i < nt; i++) {
            const uint64_t l = b200h_db_entry_len(tdb, i);
            toff[i + 1] = toff[i] + (l >= 2 ? l - 2 : 0);
        }
        std::vector<char> tdata(toff[nt] + 1);
        for (uint64_t i = 0; i < nt; i++) memcpy(tdata.data() + toff[i], b200h_db_data(tdb, i), toff[i + 1] - toff[i]);
        b200_evalue_params ev_default;
        if (evalue == nullptr) {           // EvalueComputation(tdbr->getAminoAcidDBSize(), subMat): the ungapped parameter set (:107)
            if (b200h_evalue_defaults("blosum62.out", 0, 0, 0, toff[nt], &ev_default) != B200_OK) rc = rs_fail(B200_ERR_ARG, "b200_rescorediagonal_db: no built-in E-value parameters");
            evalue = &ev_default;
        }
        if (rc == B200_OK) {               // mode -1: "these are the targets" (the device scorer uploads them once)
            rc = scorer(user, nullptr, nullptr, 0, nullptr, nullptr, nullptr, tdata.data(), toff.data(), nt, ascii, -1, nullptr);
            if (rc != B200_OK) rs_fail(rc, "b200_rescorediagonal_db: the scorer failed");
        }
        const uint64_t np = b200h_db_size(pdb);
        // the reference walks the result DB in the order of its data file (DBReader::LINEAR_ACCCESS, :402) and writes in that order
        std::vector<uint64_t> walk(np);
        for (uint64_t i = 0; i < np; i++) walk[i] = i;
        std::sort(walk.begin(), walk.end(), [&](uint64_t x, uint64_t y) { return b200h_db_data(pdb, x) < b200h_db_data(pdb, y); });
        std::vector<Entry> entries;
        std::vector<char> qdata;
        std::vector<uint64_t> qoff, hoff;
        std::vector<uint32_t> ids, keys;
        std::vector<uint16_t> diags;
        std::vector<b200_pref_hit> parsed;
        std::vector<b200_rescore> scored;
        std::vector<b200_result> alns;
        std::vector<std::string> alns_bt;
        std::vector<b200_pref_hit> shorts;
        std::string entry;
        char line[1024];
        for (uint64_t b0 = 0; b0 < np && rc == B200_OK; b0 += bucket_queries) {
            const uint64_t b1 = std::min<uint64_t>(np, b0 + bucket_queries);
            entries.clear(); qdata.clear(); qoff.assign(1, 0); hoff.assign(1, 0); ids.clear(); keys.clear(); diags.clear();
            for (uint64_t w = b0; w < b1 && rc == B200_OK; w++) {
                const uint64_t i = walk[w];
                Entry e;
                e.key = b200h_db_key(pdb, i); e.qid = -1; e.qlen = -1; e.hit_begin = ids.size();
                const char *pe = b200h_db_data(pdb, i);
                parsed.resize(b200h_db_entry_len(pdb, i) / 4 + 2);
                const size_t nh = b200h_parse_prefilter_hits(pe, parsed.data(), parsed.size());
                if (*pe != '\0') {          // the query is looked up only for non-empty entries (:157-164)
                    e.qid = b200h_db_id(qdb, e.key);
                    if (e.qid < 0) { rc = rs_fail(B200_ERR_ARG, "b200_rescorediagonal_db: query key " + std::to_string(e.key) + " of the result DB missing from the query DB"); break; }
                    const uint64_t l = b200h_db_entry_len(qdb, (uint64_t) e.qid);
                    e.qlen = (int) (l >= 2 ? l - 2 : 0);
                    const char *s = b200h_db_data(qdb, (uint64_t) e.qid);
                    qdata.insert(qdata.end(), s, s + e.qlen);
                }
                qoff.push_back(qdata.size());
                for (size_t k = 0; k < nh; k++) {
                    const int64_t tid = b200h_db_id(tdb, parsed[k].seq_id);
                    if (tid < 0) { rc = rs_fail(B200_ERR_ARG, "b200_rescorediagonal_db: target key " + std::to_string(parsed[k].seq_id) + " missing from the target DB"); break; }
                    const int db_len = (int) (toff[tid + 1] - toff[tid]);
                    if (!coverable(par->cov_thr, par->cov_mode, static_cast<float>(e.qlen), static_cast<float>(db_len))) continue;   // :218-220
                    ids.push_back((uint32_t) tid); keys.push_back(parsed[k].seq_id); diags.push_back(parsed[k].diagonal);
                }
                hoff.push_back(ids.size());
                entries.push_back(e);
            }
            if (rc != B200_OK) break;
            scored.resize(ids.size() + 1);
            if (!ids.empty()) {
                qdata.push_back('\0');
                rc = scorer(user, qdata.data(), qoff.data(), (int) entries.size(), hoff.data(), ids.data(), diags.data(), tdata.data(), toff.data(), nt, ascii, mode, scored.data());
                if (rc != B200_OK) { rs_fail(rc, "b200_rescorediagonal_db: the scorer failed"); break; }
            }
            total_hits += ids.size();
            for (size_t ei = 0; ei < entries.size() && rc == B200_OK; ei++) {
                const Entry &e = entries[ei];
                alns.clear(); alns_bt.clear(); shorts.clear();
                for (uint64_t h = hoff[ei]; h < hoff[ei + 1]; h++) {
                    const b200_rescore &a = scored[h];
                    const uint32_t tid = ids[h];
                    const int db_len = (int) (toff[tid + 1] - toff[tid]);
                    const int q_len = e.qlen;
                    const bool is_identity = ((uint64_t) e.qid == (uint64_t) tid && (par->include_identity || same));   // :213 (ids of the two readers)
                    const unsigned int dist_to_diag = (unsigned int) a.dist_to_diagonal;
                    const int diagonal_len = a.diagonal_len, distance = a.score, diagonal = a.diagonal;
                    double seq_id = 0, eval = 0.0;
                    int bit_score = 0, aln_len = 0;
                    float tcov = static_cast<float>(diagonal_len) / static_cast<float>(db_len);
                    float qcov = static_cast<float>(diagonal_len) / static_cast<float>(q_len);
                    b200_result res;
                    std::string bt;
                    memset(&res, 0, sizeof(res));
                    if (mode == 0) {
                        const int id_cnt = (int) (static_cast<float>(distance));
                        seq_id = seq_id_of(par->seq_id_mode, id_cnt, q_len, db_len, diagonal_len);
                        aln_len = diagonal_len;
                    } else {
                        eval = b200h_evalue(evalue, distance, q_len);
                        bit_score = static_cast<int>(b200h_bit_score(evalue, distance) + 0.5);
                        if (aln_mode) {
                            aln_len = (a.end_pos - a.start_pos) + 1;
                            int qs, qe, ds, de;
                            if (diagonal >= 0) { qs = a.start_pos + (int) dist_to_diag; qe = a.end_pos + (int) dist_to_diag; ds = a.start_pos; de = a.end_pos; }
                            else { qs = a.start_pos; qe = a.end_pos; ds = a.start_pos + (int) dist_to_diag; de = a.end_pos + (int) dist_to_diag; }
                            if (eval <= par->eval_thr || is_identity)      // the identity count only matters for hits that can still pass (:293-303)
                                seq_id = seq_id_of(par->seq_id_mode, a.identical, q_len, db_len, aln_len);
                            if (par->add_backtrace) { bt = std::to_string(aln_len); bt.push_back('M'); }
                            qcov = cov_of((unsigned) qs, (unsigned) qe, (unsigned) q_len);
                            tcov = cov_of((unsigned) ds, (unsigned) de, (unsigned) db_len);
                            res.db_key = keys[h]; res.score = bit_score; res.qcov = qcov; res.dbcov = tcov; res.seq_id = (float) seq_id; res.eval = eval;
                            res.aln_length = (uint32_t) aln_len; res.q_start = qs; res.q_end = qe; res.q_len = q_len; res.db_start = ds; res.db_end = de;
                            res.db_len = db_len; res.bt_len = (uint32_t) bt.size();
                        }
                    }
                    const bool has_cov = covered(par->cov_thr, par->cov_mode, qcov, tcov);
                    const bool has_seq_id = seq_id >= (par->seq_id_thr - std::numeric_limits<float>::epsilon());
                    const bool has_eval = (eval <= par->eval_thr);
                    const bool has_aln_len = (aln_len >= par->aln_len_thr);
                    if (!(is_identity || (has_aln_len && has_cov && has_seq_id && has_eval))) continue;
                    if (aln_mode) {
                        alns.push_back(res); alns_bt.push_back(bt);
                        alns.back().bt_off = alns_bt.size() - 1;
                    } else {
                        b200_pref_hit hit;
                        hit.seq_id = keys[h]; hit.pad_ = 0; hit.diagonal = (uint16_t) diagonal;
                        if (mode == 1) hit.pref_score = bit_score;
                        else { hit.pref_score = 100 * seq_id; }                                  // double product truncated to int (:335)
                        shorts.push_back(hit);
                    }
                }
                if (par->sort_results > 0 && alns.size() > 1)
                    std::sort(alns.begin(), alns.end(), [](const b200_result &x, const b200_result &y) {      // Matcher::compareHits
                        if (x.eval != y.eval) return x.eval < y.eval;
                        if (x.score != y.score) return x.score > y.score;
                        if (x.db_len != y.db_len) return x.db_len < y.db_len;
                        return x.db_key < y.db_key;
                    });
                if (par->sort_results > 0 && shorts.size() > 1)
                    std::sort(shorts.begin(), shorts.end(), [](const b200_pref_hit &x, const b200_pref_hit &y) {    // hit_t::compareHitsByScoreAndId
                        const int ax = abs(x.pref_score), ay = abs(y.pref_score);
                        if (ax != ay) return ax > ay;
                        return x.seq_id < y.seq_id;
                    });
                entry.clear();
                for (const b200_result &r : alns) {
                    const std::string &bt = alns_bt[r.bt_off];
                    const size_t len = b200h_result_to_buffer(line, &r, bt.c_str(), par->add_backtrace, 0);
                    entry.append(line, len);
                }
                for (const b200_pref_hit &hh : shorts) entry.append(line, b200h_prefilter_hit_to_buffer(line, &hh));
                total_rec += alns.size() + shorts.size();
                if (b200h_dbw_write(out, e.key, entry.data(), entry.size()) != B200_OK) rc = rs_fail(B200_ERR_ARG, b200h_db_last_error());
            }
        }
    }
    if (out != nullptr && b200h_dbw_close(out) != B200_OK && rc == B200_OK) rc = rs_fail(B200_ERR_ARG, b200h_db_last_error());
    if (pdb != nullptr) b200h_db_close(pdb);
    if (qdb != nullptr && qdb != tdb) b200h_db_close(qdb);
    if (tdb != nullptr) b200h_db_close(tdb);
    if (n_hits != nullptr) *n_hits = total_hits;
    if (n_records != nullptr) *n_records = total_rec;
    return rc;
}

int b200_rescorediagonal_db(b200_ctx *ctx, const char *query_db, const char *target_db, const char *prefilter_db, const char *out_db,
                            const int16_t *sub_matrix, const char *num2aa, int alphabet, const b200_rescore_params *par,
                            const b200_evalue_params *evalue, uint32_t bucket_queries, uint64_t *n_hits, uint64_t *n_records) {
    if (ctx == nullptr) return B200_ERR_ARG;
    DeviceScorer dev = {ctx};
    // the module's "targets are ready" call uploads the ASCII target DB; every later call scores one bucket of hit lists
    struct Adapter {
        static int call(void *user, const char *qd, const uint64_t *qo, int nq, const uint64_t *ho, const uint32_t *ids, const uint16_t *dg,
                        const char *td, const uint64_t *to, uint64_t nt, const int8_t *m, int mode, b200_rescore *out) {
            if (mode < 0) return b200_db_load_ascii(static_cast<DeviceScorer *>(user)->ctx, td, to, nt);
            return device_scorer(user, qd, qo, nq, ho, ids, dg, td, to, nt, m, mode, out);
        }
    };
    const int rc = b200h_rescorediagonal_db_with(&Adapter::call, &dev, query_db, target_db, prefilter_db, out_db, sub_matrix, num2aa, alphabet, par, evalue,
                                                 bucket_queries, n_hits, n_records);
    // a failure of the device scorer has left its own text in the context; everything else is reported from here
    if (rc != B200_OK && g_rs_err.find("the scorer failed") == std::string::npos) b200_set_err(ctx, rc, g_rs_err.c_str());
    return rc;
}

}  // extern "C"
