// include/b200_mmseqs.hpp -- host-side C++ mirror of the reference's operator surface, header-only, over the C ABI
// (b200_align.h + b200_host.h).  C++11, no exceptions (MMseqs2 is built -fno-exceptions), no CUDA headers.
//
// The classes keep the reference's names, argument meaning and calling pattern so that the call sites in
// ungappedprefilter.cpp / QueryMatcher.cpp / Alignment.cpp change by a type name and a batch boundary, not by logic:
//
//   b200::Marv               class Marv                      lib/libmarv/src/marv.h:6-58  (scan called at ungappedprefilter.cpp:207)
//   b200::UngappedAlignment  class UngappedAlignment         src/prefiltering/UngappedAlignment.h:16-34  (QueryMatcher.cpp:119,131)
//   b200::DiagSubmitQueue    (no counterpart: QueryMatcher owns one UngappedAlignment per OpenMP thread, QueryMatcher.cpp:73 --
//                            here those threads' align() calls are combined into one device call)
//   b200::SmithWaterman      class SmithWaterman             src/alignment/StripedSmithWaterman.h:84-222 (Matcher.cpp:51-144)
//
// What differs, and why: the device wants many (query,target) pairs per call.  SmithWaterman therefore collects the
// targets of one query (addTarget) and resolves them in one flush(); the E-value / coverage gate between "score",
// "end position" and "start position" stays with the caller's EvalueComputation (host double math), expressed as a
// callback, exactly where ssw_align_private applies it (StripedSmithWaterman.cpp:854-863).
#ifndef B200_MMSEQS_HPP
#define B200_MMSEQS_HPP

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

#include "b200_align.h"
#include "b200_host.h"

namespace b200 {

// One process-wide device context + resident DB (the role Marv's cudasw handle plays).
class Device {
public:
    explicit Device(int deviceId = 0) : ctx_(NULL), status_(b200_create(deviceId, &ctx_)) {}
    ~Device() { if (ctx_) b200_destroy(ctx_); }
    bool ok() const { return status_ == B200_OK; }
    int status() const { return status_; }
    const char *error() const { return ctx_ ? b200_last_error(ctx_) : "b200_create failed"; }
    b200_ctx *ctx() { return ctx_; }
    // SequenceLookup layout (src/prefiltering/SequenceLookup.cpp:42-46): data + offsets[n+1]
    int loadLookup(const unsigned char *data, const size_t *offsets, size_t nSeq, int alphabetSize) {
        std::vector<uint64_t> off(nSeq + 1);
        for (size_t i = 0; i <= nSeq; i++) off[i] = offsets[i];
        return b200_db_load(ctx_, data, off.data(), nSeq, alphabetSize);
    }
private:
    Device(const Device &);
    Device &operator=(const Device &);
    b200_ctx *ctx_;
    int status_;
};

// ---------------------------------------------------------------------------------------------------------------
// Marv: all-diagonals ungapped scan of one query profile against the resident DB.
// ---------------------------------------------------------------------------------------------------------------
class Marv {
public:
    enum AlignmentType { GAPLESS, SMITH_WATERMAN, GAPLESS_SMITH_WATERMAN };
    struct Stats { size_t results; int numOverflows; double seconds; double gcups; };
    struct Result {
        unsigned int id; int score; int qEndPos; int dbEndPos;
        Result() : id(0), score(0), qEndPos(0), dbEndPos(0) {}
        Result(unsigned int id, int score, int qEndPos, int dbEndPos) : id(id), score(score), qEndPos(qEndPos), dbEndPos(dbEndPos) {}
    };

    Marv(Device *dev, size_t dbEntries, int alphabetSize, int maxSeqLength, size_t maxSeqs, AlignmentType type = GAPLESS)
        : dev_(dev), dbEntries_(dbEntries), alphabetSize_(alphabetSize), maxSeqs_(maxSeqs), type_(type), minScoreExcl_(0) {
        (void) maxSeqLength;
    }

    // The padded GPU DB of makepaddedseqdb (src/util/makepaddedseqdb.cpp:66-98): numeric codes, +32 = soft-masked,
    // every entry padded to a multiple of 4 with code 20.  offset[i] / length[i] as ungappedprefilter.cpp:132-136 builds
    // them.  Masked residues are mapped to X (alphabetSize-1) as runFilterOnCpu does (ungappedprefilter.cpp:402-405).
    void *loadDb(char *data, size_t *offset, int32_t *length, size_t dbByteSize) {
        (void) dbByteSize;
        lastStatus_ = b200_db_load_padded(dev_->ctx(), reinterpret_cast<const uint8_t *>(data), offset, length, dbEntries_, alphabetSize_);
        return dev_;
    }
    void setDb(void *) {}
    void prefetch() {}
    void setMinScore(int minScoreExcl) { minScoreExcl_ = minScoreExcl; }  // minDiagScoreThr, applied on the device
    int lastStatus() const { return lastStatus_; }

    // pssm: [alphabetSize][sequenceLength] int8 exactly as ungappedprefilter.cpp:195-203 fills it; `bias` is the SSW
    // profile bias (b200h_ssw_bias) that makes the u8 saturation of the CPU scorer reproducible.
    Stats scan(const char *sequence, size_t sequenceLength, int8_t *pssm, Result *results, int bias) {
        (void) sequence;
        Stats st; st.results = 0; st.numOverflows = 0; st.seconds = 0; st.gcups = 0;
        b200_query q; q.profile = pssm; q.qlen = (int32_t) sequenceLength; q.bias = bias;
        hits_.resize(maxSeqs_);
        uint32_t n = 0;
        lastStatus_ = b200_ungapped_scan(dev_->ctx(), &q, 1, minScoreExcl_, (uint32_t) maxSeqs_, hits_.data(), &n, NULL);
        if (lastStatus_ != B200_OK) return st;
        if (type_ == GAPLESS) {
            for (uint32_t i = 0; i < n; i++) results[i] = Result(hits_[i].id, hits_[i].score, -1, -1);
        } else {
            // GAPLESS_SMITH_WATERMAN (what `search --gpu 1 --alignment-mode 1` selects, ungappedprefilter.cpp:151-152; libmarv:
            // cudasw4.cuh:1003-1026): the best maxSeqs targets of the ungapped scan are rescored with the gapped kernel; results carry the
            // gapped score and its end positions, ordered (gapped score desc, id asc).  ungappedprefilter.cpp:282-325 turns them
            // into alignment records directly.
            std::vector<b200_pair> pairs(n);
            std::vector<b200_sw_end> ends(n);
            for (uint32_t i = 0; i < n; i++) { pairs[i].query = 0; pairs[i].target = hits_[i].id; }
            if (n > 0) {
                lastStatus_ = b200_sw_score_endpos(dev_->ctx(), &q, 1, pairs.data(), n, gapOpen_, gapExtend_, ends.data());
                if (lastStatus_ != B200_OK) return st;
            }
            std::vector<uint32_t> order(n);
            for (uint32_t i = 0; i < n; i++) order[i] = i;
            const b200_sw_end *e = ends.data();
            const b200_hit *h = hits_.data();
            std::sort(order.begin(), order.end(), [e, h](uint32_t a, uint32_t b) {
                return e[a].score != e[b].score ? e[a].score > e[b].score : h[a].id < h[b].id; });
            for (uint32_t i = 0; i < n; i++) results[i] = Result(hits_[order[i]].id, ends[order[i]].score, ends[order[i]].qend, ends[order[i]].dbend);
        }
        st.results = n;
        return st;
    }
    void setGapCosts(int open, int extend) { gapOpen_ = open; gapExtend_ = extend; }

private:
    Device *dev_;
    size_t dbEntries_;
    int alphabetSize_;
    size_t maxSeqs_;
    AlignmentType type_;
    int minScoreExcl_;
    int gapOpen_ = 11, gapExtend_ = 1;
    int lastStatus_ = B200_OK;
    std::vector<b200_hit> hits_;
};

// ---------------------------------------------------------------------------------------------------------------
// UngappedAlignment: per-(target, diagonal) scorer used inside the k-mer prefilter.
// ---------------------------------------------------------------------------------------------------------------
struct __attribute__((__packed__)) CounterResult {  // src/prefiltering/CacheFriendlyOperations.h:46-51
    unsigned int id;
    unsigned short diagonal;
    unsigned char count;
};

// Many host threads -> one device call (seam B2).  QueryMatcher runs one UngappedAlignment per OpenMP thread and each of
// them scores its own query's hits; sending those one by one costs a round trip (uploads, launch, download, synchronisation) per
// query under the context mutex.  The queue combines them: a thread that arrives while no call is in flight becomes the leader of a
// round -- it takes every request queued so far (its own among them), stages their hit arrays back to back, issues ONE
// b200_diag_score_batch, hands the results back and wakes the owners; threads that arrive during the call queue up and one of them
// leads the next round.  No helper thread, no spinning, no added latency for a lone caller (it leads a round of one).
// Backend: int operator()(const b200_query*, int nq, const uint64_t* hitOffsets, const uint32_t* ids, const uint16_t* diagonals,
//                         uint8_t* counts, int32_t* raw) -- b200_diag_score_batch on a context, or a stand-in for tests.
template <typename Backend>
class DiagSubmitQueueT {
public:
    explicit DiagSubmitQueueT(Backend backend, size_t maxQueriesPerCall = 256) : backend_(backend), maxQueries_(std::max<size_t>(1, maxQueriesPerCall)), leader_(false), rounds_(0), requests_(0) {}

    // blocks until the hits of this query are scored; counts in/out as b200_diag_score, raw may be NULL
    int submit(const b200_query &query, const uint32_t *ids, const uint16_t *diagonals, size_t n, uint8_t *counts, int32_t *raw) {
        Request req;
        req.query = query; req.ids = ids; req.diagonals = diagonals; req.counts = counts; req.raw = raw; req.n = n; req.rc = B200_OK; req.done = false;
        std::unique_lock<std::mutex> lk(mu_);
        pending_.push_back(&req);
        requests_++;
        while (!req.done) {
            if (leader_) { cv_.wait(lk); continue; }
            leader_ = true;                                          // this thread leads a round; its own request is still pending
            std::vector<Request *> batch;
            const size_t take = std::min(pending_.size(), maxQueries_);
            batch.assign(pending_.begin(), pending_.begin() + take);
            pending_.erase(pending_.begin(), pending_.begin() + take);
            lk.unlock();
            const int rc = run(batch);
            lk.lock();
            for (size_t i = 0; i < batch.size(); i++) { batch[i]->rc = rc; batch[i]->done = true; }
            rounds_++;
            leader_ = false;
            cv_.notify_all();
        }
        return req.rc;
    }

    size_t rounds() const { std::lock_guard<std::mutex> lk(mu_); return rounds_; }        // device calls issued
    size_t requests() const { std::lock_guard<std::mutex> lk(mu_); return requests_; }    // submit() calls served or queued

private:
    struct Request {
        b200_query query;
        const uint32_t *ids;
        const uint16_t *diagonals;
        uint8_t *counts;
        int32_t *raw;
        size_t n;
        int rc;
        bool done;
    };

    // only the leader of the current round runs this, so the staging vectors need no lock
    int run(const std::vector<Request *> &batch) {
        const size_t nq = batch.size();
        queries_.resize(nq); offsets_.assign(nq + 1, 0);
        bool wantRaw = false;
        for (size_t i = 0; i < nq; i++) {
            queries_[i] = batch[i]->query;
            offsets_[i + 1] = offsets_[i] + batch[i]->n;
            wantRaw = wantRaw || batch[i]->raw != NULL;
        }
        const size_t total = (size_t) offsets_[nq];
        if (total == 0) return B200_OK;
        ids_.resize(total); diags_.resize(total); counts_.resize(total);
        if (wantRaw) raw_.resize(total);
        for (size_t i = 0; i < nq; i++) {
            const Request &r = *batch[i];
            if (r.n == 0) continue;
            memcpy(&ids_[offsets_[i]], r.ids, r.n * sizeof(uint32_t));
            memcpy(&diags_[offsets_[i]], r.diagonals, r.n * sizeof(uint16_t));
            memcpy(&counts_[offsets_[i]], r.counts, r.n);
        }
        const int rc = backend_(queries_.data(), (int) nq, offsets_.data(), ids_.data(), diags_.data(), counts_.data(), wantRaw ? raw_.data() : NULL);
        if (rc != B200_OK) return rc;
        for (size_t i = 0; i < nq; i++) {
            const Request &r = *batch[i];
            if (r.n == 0) continue;
            memcpy(r.counts, &counts_[offsets_[i]], r.n);
            if (r.raw != NULL) memcpy(r.raw, &raw_[offsets_[i]], r.n * sizeof(int32_t));
        }
        return B200_OK;
    }

    DiagSubmitQueueT(const DiagSubmitQueueT &);
    DiagSubmitQueueT &operator=(const DiagSubmitQueueT &);
    Backend backend_;
    size_t maxQueries_;
    mutable std::mutex mu_;
    std::condition_variable cv_;
    std::vector<Request *> pending_;
    bool leader_;
    size_t rounds_, requests_;
    std::vector<b200_query> queries_;
    std::vector<uint64_t> offsets_;
    std::vector<uint32_t> ids_;
    std::vector<uint16_t> diags_;
    std::vector<uint8_t> counts_;
    std::vector<int32_t> raw_;
};

struct DeviceDiagBackend {       // the product backend: one batched call on the shared context
    b200_ctx *ctx;
    int operator()(const b200_query *q, int nq, const uint64_t *off, const uint32_t *ids, const uint16_t *dg, uint8_t *counts, int32_t *raw) const {
        return b200_diag_score_batch(ctx, q, nq, off, ids, dg, counts, raw);
    }
};

class DiagSubmitQueue : public DiagSubmitQueueT<DeviceDiagBackend> {
public:
    explicit DiagSubmitQueue(Device *dev, size_t maxQueriesPerCall = 256) : DiagSubmitQueueT<DeviceDiagBackend>(make(dev), maxQueriesPerCall) {}
private:
    static DeviceDiagBackend make(Device *dev) { DeviceDiagBackend b; b.ctx = dev->ctx(); return b; }
};

class UngappedAlignment {
public:
    // subMatrix: A*A int16 row-major copy of BaseMatrix::subMatrix; the lookup must already be loaded into `dev`.
    // queue (optional): the DiagSubmitQueue shared by the UngappedAlignment objects of all OpenMP threads
    UngappedAlignment(Device *dev, const int16_t *subMatrix, int alphabetSize, DiagSubmitQueue *queue = NULL)
        : dev_(dev), queue_(queue), mat_(subMatrix, subMatrix + (size_t) alphabetSize * alphabetSize), A_(alphabetSize), qlen_(0) {}

    // createProfile(Sequence*, float* biasCorrection)   UngappedAlignment.cpp:388-421
    int createProfile(const unsigned char *numSequence, int L, const float *biasCorrection) {
        qlen_ = L;
        cb_.assign(L, 0);
        if (biasCorrection != NULL) b200h_round_bias_diag(biasCorrection, L, cb_.data());
        profile_.resize((size_t) A_ * L);
        return b200h_build_profile(mat_.data(), A_, numSequence, L, cb_.data(), /*target_major=*/0, profile_.data());
    }

    // createProfile for a Sequence of type HMM_PROFILE (UngappedAlignment.cpp:405-411): alignmentProfile =
    // seq->getAlignmentProfile(), [PROFILE_AA_SIZE][L] int8; no bias correction in this branch
    int createProfile(const int8_t *alignmentProfile, int profileRows, int L) {
        qlen_ = L;
        profile_.resize((size_t) A_ * L);
        return b200h_build_profile_pssm(alignmentProfile, profileRows, L, A_, profile_.data()) < 0 ? B200_ERR_ARG : B200_OK;
    }

    // align(CounterResult*, n)   UngappedAlignment.cpp:36-42 -- counts updated in place
    int align(CounterResult *results, size_t n) { return run(results, n, NULL); }

    // scoreSingelSequenceByCounterResult for a batch (QueryMatcher::getResult re-scores clamped hits, :448-450)
    int rescore(CounterResult *results, size_t n, int32_t *raw) { return run(results, n, raw); }

private:
    int run(CounterResult *results, size_t n, int32_t *raw) {
        ids_.resize(n); diags_.resize(n); counts_.resize(n);
        for (size_t i = 0; i < n; i++) { ids_[i] = results[i].id; diags_[i] = results[i].diagonal; counts_[i] = results[i].count; }
        b200_query q; q.profile = profile_.data(); q.qlen = qlen_; q.bias = 0;
        const int rc = queue_ != NULL ? queue_->submit(q, ids_.data(), diags_.data(), n, counts_.data(), raw)
                                      : b200_diag_score(dev_->ctx(), &q, ids_.data(), diags_.data(), n, counts_.data(), raw);
        if (rc == B200_OK) for (size_t i = 0; i < n; i++) results[i].count = counts_[i];
        return rc;
    }
    Device *dev_;
    DiagSubmitQueue *queue_;
    std::vector<int16_t> mat_;
    int A_, qlen_;
    std::vector<int8_t> cb_, profile_;
    std::vector<uint32_t> ids_;
    std::vector<uint16_t> diags_;
    std::vector<uint8_t> counts_;
};

// ---------------------------------------------------------------------------------------------------------------
// SmithWaterman: ssw_init once per query, then the query's whole prefilter list in one flush.
// ---------------------------------------------------------------------------------------------------------------
struct s_align {  // the fields of StripedSmithWaterman.h:59-74 this path produces
    uint32_t score1;
    int32_t dbStartPos1, dbEndPos1, qStartPos1, qEndPos1;
    float qCov, tCov;
    int word;
    uint32_t identicalAACnt;   // alignment modes >= 2
    std::string backtrace;     // M/I/D string of computerBacktrace, alignment modes >= 2
};

class SmithWaterman {
public:
    SmithWaterman(Device *dev, const int16_t *subMatrix, const double *pBack, int alphabetSize, bool aaBiasCorrection,
                  float aaBiasCorrectionScale)
        : dev_(dev), mat_(subMatrix, subMatrix + (size_t) alphabetSize * alphabetSize), pback_(pBack, pBack + alphabetSize),
          A_(alphabetSize), biasCorr_(aaBiasCorrection), scale_(aaBiasCorrectionScale), qlen_(0), bias_(0) {}

    // ssw_init(const Sequence* q, const int8_t* mat, const BaseMatrix* m)   StripedSmithWaterman.cpp:1364-1476
    int ssw_init(const unsigned char *numSequence, int L) {
        qlen_ = L;
        seq_.assign(numSequence, numSequence + L);
        cb_.assign(L, 0);
        if (biasCorr_) {
            tmp_.resize(L);
            b200h_comp_bias(mat_.data(), pback_.data(), A_, numSequence, L, scale_, tmp_.data());
            b200h_round_bias_ssw(tmp_.data(), L, cb_.data());
        }
        bias_ = b200h_ssw_bias(mat_.data(), A_, cb_.data(), L, biasCorr_ ? 1 : 0);
        profile_.resize((size_t) A_ * L);
        targets_.clear();
        return b200h_build_profile(mat_.data(), A_, numSequence, L, cb_.data(), /*target_major=*/1, profile_.data());
    }

    // ssw_init for a profile query (the isProfile branch, StripedSmithWaterman.cpp:1388-1425): consensus = q->numSequence
    // (identity counting in the backtrace), alignmentProfile = q->getAlignmentProfile(), [PROFILE_AA_SIZE][L] int8
    int ssw_init(const unsigned char *consensus, const int8_t *alignmentProfile, int profileRows, int L) {
        qlen_ = L;
        seq_.assign(consensus, consensus + L);
        cb_.assign(L, 0);
        profile_.resize((size_t) A_ * L);
        targets_.clear();
        bias_ = b200h_build_profile_pssm(alignmentProfile, profileRows, L, A_, profile_.data());
        return bias_ < 0 ? B200_ERR_ARG : B200_OK;
    }

    void addTarget(uint32_t dbId) { targets_.push_back(dbId); }
    size_t pending() const { return targets_.size(); }

    static float computeCov(unsigned int startPos, unsigned int endPos, unsigned int len) {  // StripedSmithWaterman.cpp:1762
        return (std::min(len, std::max(startPos, endPos)) - std::min(startPos, endPos) + 1) / (float) len;
    }

    // Resolves every pending target the way ssw_align_private does (StripedSmithWaterman.cpp:831-890), alignment modes
    // 0/1/2 up to start positions:
    //   1. score of all targets (packed kernel)
    //   2. gateScore(score1, userData) -> false: the hit is dropped before positions are computed (E-value gate;
    //      such hits fail Alignment::checkCriteria anyway, Alignment.cpp:389)
    //   3. end positions of the survivors; gateCov(qCov, tCov, userData) as hasLowerCoverage does with start 0
    //   4. start positions of the remaining ones (alignmentMode >= 1)
    //   5. CIGAR / identities of those (alignmentMode >= 2): banded_sw + computerBacktrace on the device
    // dbLen[i] is needed for the coverage values only.  out[i].dbEndPos1 == -1 marks "no residue aligned" or "gated out".
    typedef bool (*GateScore)(uint32_t score1, void *userData);
    typedef bool (*GateCov)(float qCov, float tCov, void *userData);
    int flush(const int32_t *dbLen, uint8_t gapOpen, uint8_t gapExtend, uint8_t alignmentMode, GateScore gateScore,
              GateCov gateCov, void *userData, std::vector<s_align> &out) {
        const size_t n = targets_.size();
        out.resize(n);
        if (n == 0) return B200_OK;
        b200_query q; q.profile = profile_.data(); q.qlen = qlen_; q.bias = bias_;
        std::vector<b200_pair> pairs(n);
        for (size_t i = 0; i < n; i++) { pairs[i].query = 0; pairs[i].target = targets_[i]; }
        std::vector<int32_t> score(n);
        int rc = b200_sw_score(dev_->ctx(), &q, 1, pairs.data(), n, gapOpen, gapExtend, score.data());
        if (rc != B200_OK) return rc;
        std::vector<b200_pair> sub;
        std::vector<int32_t> subScore;
        std::vector<size_t> idx;
        for (size_t i = 0; i < n; i++) {
            s_align &r = out[i];
            r.score1 = (uint32_t) score[i]; r.dbStartPos1 = -1; r.qStartPos1 = -1; r.dbEndPos1 = -1; r.qEndPos1 = qlen_ - 1;
            r.qCov = 0; r.tCov = 0; r.word = (score[i] + bias_ >= 255) ? 1 : 0; r.identicalAACnt = 0; r.backtrace.clear();
            if (score[i] > 0 && (gateScore == NULL || gateScore(r.score1, userData))) { sub.push_back(pairs[i]); subScore.push_back(score[i]); idx.push_back(i); }
        }
        if (sub.empty()) { targets_.clear(); return B200_OK; }
        std::vector<b200_sw_end> ends(sub.size());
        rc = b200_sw_endpos(dev_->ctx(), &q, 1, sub.data(), sub.size(), gapOpen, gapExtend, subScore.data(), ends.data());   // known scores
        if (rc != B200_OK) return rc;
        std::vector<b200_pair> sub2;
        std::vector<b200_sw_end> ends2;
        std::vector<size_t> idx2;
        for (size_t k = 0; k < sub.size(); k++) {
            s_align &r = out[idx[k]];
            r.qEndPos1 = ends[k].qend; r.dbEndPos1 = ends[k].dbend; r.word = ends[k].word;
            if (ends[k].dbend == -1) continue;
            r.qCov = computeCov(0, r.qEndPos1, qlen_);
            r.tCov = computeCov(0, r.dbEndPos1, dbLen[idx[k]]);
            const bool covOk = gateCov == NULL || gateCov(r.qCov, r.tCov, userData);
            if (alignmentMode == 0 || !covOk) continue;
            sub2.push_back(sub[k]); ends2.push_back(ends[k]); idx2.push_back(idx[k]);
        }
        if (!sub2.empty()) {
            std::vector<b200_sw_aln> aln(sub2.size());
            rc = b200_sw_startpos(dev_->ctx(), &q, 1, sub2.data(), sub2.size(), gapOpen, gapExtend, ends2.data(), aln.data());
            if (rc != B200_OK) return rc;
            for (size_t k = 0; k < sub2.size(); k++) {
                s_align &r = out[idx2[k]];
                r.qStartPos1 = aln[k].qstart; r.dbStartPos1 = aln[k].dbstart;
                r.qCov = computeCov(r.qStartPos1, r.qEndPos1, qlen_);
                r.tCov = computeCov(r.dbStartPos1, r.dbEndPos1, dbLen[idx2[k]]);
            }
            if (alignmentMode >= 2) {   // alignStartPosBacktrace generates the CIGAR unless the start-based coverage fails (:1213-1218)
                std::vector<b200_pair> sub3; std::vector<b200_sw_aln> aln3; std::vector<size_t> idx3;
                std::vector<uint64_t> coff(1, 0);
                for (size_t k = 0; k < sub2.size(); k++) {
                    const s_align &r = out[idx2[k]];
                    if (gateCov != NULL && !gateCov(r.qCov, r.tCov, userData)) continue;
                    sub3.push_back(sub2[k]); aln3.push_back(aln[k]); idx3.push_back(idx2[k]);
                    coff.push_back(coff.back() + (uint64_t) (aln[k].qend - aln[k].qstart + 1) + (uint64_t) (aln[k].dbend - aln[k].dbstart + 1) + 2);
                }
                if (!sub3.empty()) {
                    std::vector<b200_sw_bt> bt(sub3.size());
                    std::vector<uint32_t> cig(coff.back() + 1);
                    const uint8_t *seqp = seq_.data();
                    rc = b200_sw_backtrace(dev_->ctx(), &q, &seqp, 1, sub3.data(), sub3.size(), gapOpen, gapExtend, aln3.data(), bt.data(),
                                           cig.data(), coff.data());
                    if (rc != B200_OK) return rc;
                    for (size_t k = 0; k < sub3.size(); k++) {
                        s_align &r = out[idx3[k]];
                        r.identicalAACnt = (uint32_t) bt[k].identical;
                        r.backtrace.reserve((size_t) bt[k].bt_len);
                        for (int c = 0; c < bt[k].n_cigar; c++) r.backtrace.append((size_t) (cig[coff[k] + c] >> 4), "MID"[cig[coff[k] + c] & 0xfu]);
                    }
                }
            }
        }
        targets_.clear();
        return B200_OK;
    }

    int bias() const { return bias_; }
    const int8_t *compositionBias() const { return cb_.data(); }

private:
    Device *dev_;
    std::vector<int16_t> mat_;
    std::vector<double> pback_;
    int A_;
    bool biasCorr_;
    float scale_;
    int qlen_, bias_;
    std::vector<float> tmp_;
    std::vector<int8_t> cb_, profile_;
    std::vector<uint8_t> seq_;
    std::vector<uint32_t> targets_;
};

}  // namespace b200
#endif  // B200_MMSEQS_HPP
