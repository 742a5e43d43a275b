// integration/shim/b200_align_module.h -- the `align` module's per-query loop (src/alignment/Alignment.cpp:283-520) handed to
// libb200align.so in buckets of queries.  Included by the patched Alignment.cpp under HAVE_B200 (integration/mmseqs_b200.patch);
// written against the reference's own classes (DBReader, Sequence, Matcher::result_t, DBWriter), so everything either side of the
// device call -- reading the prefilter DB, mapping sequences, writing records -- is the reference's code, and the part in between
// is b200_align_batch (include/b200_alignment.h): canBeCovered, ssw_align of every (query, hit), getSWResult's assembly,
// checkCriteria, --max-accept/--max-rejected, compareHits.
//
// Eligibility (everything else keeps the reference loop -- an input-domain guard, not a fallback of the kernels):
// amino-acid sequence queries and targets, BLOSUM62 with gap 11/1 (the hard-coded E-value parameter set of
// EvalueComputation.h:56-81), no realign / alternative alignments / LCA / wrapped scoring / correlation score / score bias,
// alignment output mode 0.  MMSEQS_B200_ALIGN=0 in the environment disables the device path.
#ifndef B200_ALIGN_MODULE_H
#define B200_ALIGN_MODULE_H

#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "b200_align.h"
#include "b200_alignment.h"
#include "b200_multi.h"

#include "DBReader.h"
#include "DBWriter.h"
#include "Debug.h"
#include "Matcher.h"
#include "Parameters.h"
#include "QueryMatcher.h"
#include "Sequence.h"
#include "Util.h"

#ifdef OPENMP
#include <omp.h>
#endif

namespace b200shim {

struct AlignSettings {
    int gapOpen, gapExtend;
    unsigned int swMode;
    double evalThr;
    double covThr, canCovThr;
    int covMode;
    double seqIdThr;
    int alnLenThr;
    int seqIdMode;
    unsigned int maxAccept, maxReject;
    bool compBiasCorrection;
    float compBiasCorrectionScale;
    bool includeIdentity, sameQTDB;
    bool addBacktrace;
    size_t maxSeqLen;
    unsigned int threads;
};

static inline bool alignEnabled() {
    const char *e = getenv("MMSEQS_B200_ALIGN");
    return !(e != NULL && e[0] == '0');
}

// Returns false (nothing written) when the device path does not apply; true after the whole [dbFrom, dbFrom+dbSize) range of the
// prefilter DB has been aligned and written to dbw.
static inline bool alignOnDevice(DBReader<DBKeyType> *qdbr, DBReader<DBKeyType> *tdbr, DBReader<DBKeyType> *prefdbr, DBWriter &dbw,
                                 BaseMatrix *m, int querySeqType, int targetSeqType, const AlignSettings &s, size_t dbFrom, size_t dbSize,
                                 size_t &alignmentsNum, size_t &totalPassedNum) {
    if (!alignEnabled()) return false;
    if (!Parameters::isEqualDbtype(querySeqType, Parameters::DBTYPE_AMINO_ACIDS) ||
        !Parameters::isEqualDbtype(targetSeqType, Parameters::DBTYPE_AMINO_ACIDS)) return false;
    if (s.canCovThr != s.covThr) return false;
    b200_evalue_params ev;
    if (b200h_evalue_defaults(m->getMatrixName().c_str(), s.gapOpen, s.gapExtend, 1, tdbr->getAminoAcidDBSize(), &ev) != B200_OK) return false;

    const int A = m->alphabetSize;
    const unsigned int threads = s.threads > 0 ? s.threads : 1;
    typedef std::chrono::steady_clock Clock;
    const Clock::time_point tStart = Clock::now();
    auto since = [](Clock::time_point a) { return std::chrono::duration<double>(Clock::now() - a).count(); };
    // every visible device (B200_DEVICE=<id> pins one): target DB replicated, each bucket's queries split across the devices
    b200_multi *multi = NULL;
    const char *dev = getenv("B200_DEVICE");
    const int oneDev = dev != NULL ? atoi(dev) : 0;
    if ((dev != NULL ? b200_multi_create(&oneDev, 1, &multi) : b200_multi_create(NULL, 0, &multi)) != B200_OK) {
        Debug(Debug::ERROR) << "libb200align: b200_multi_create failed\n";
        EXIT(EXIT_FAILURE);
    }

    const double tCreate = since(tStart);
    // ---- target DB -> numeric residues in HBM (DB-local id == DBReader id) ------------------------------------------------
    const size_t nT = tdbr->getSize();
    std::vector<uint64_t> tOff(nT + 1, 0);
    std::vector<uint32_t> tKeys(nT);
    for (size_t i = 0; i < nT; i++) {
        tOff[i + 1] = tOff[i] + tdbr->getSeqLen(i);
        tKeys[i] = (uint32_t) tdbr->getDbKey(i);
    }
    int rc;
    double tMap = 0;
    const bool paddedDb = (DBReader<DBKeyType>::getExtendedDbtype(tdbr->getDbtype()) & Parameters::DBTYPE_EXTENDED_GPU) != 0;
    if (paddedDb && tdbr->getDataFileCnt() == 1) {
        // the padded GPU DB already holds numeric codes (makepaddedseqdb.cpp:77-86): hand it over as it lies, mask bit stripped --
        // the residues Sequence::mapSequence would produce from DBReader::getUnpadded, without touching every sequence on the host
        std::vector<size_t> pOff(nT);
        std::vector<int32_t> pLen(nT);
        for (size_t i = 0; i < nT; i++) {
            pOff[i] = tdbr->getOffset(i);
            pLen[i] = (int32_t) tdbr->getSeqLen(i);
        }
        rc = b200_multi_db_load_padded_unmasked(multi, reinterpret_cast<const uint8_t *>(tdbr->getDataForFile(0)), pOff.data(), pLen.data(), nT, A);
    } else {
        std::vector<uint8_t> tRes(tOff[nT] + 1);
#pragma omp parallel num_threads(threads)
        {
            unsigned int thread_idx = 0;
#ifdef OPENMP
            thread_idx = static_cast<unsigned int>(omp_get_thread_num());
#endif
            Sequence dbSeq(s.maxSeqLen, targetSeqType, m, 0, false, s.compBiasCorrection);
#pragma omp for schedule(dynamic, 1000)
            for (size_t i = 0; i < nT; i++) {
                dbSeq.mapSequence(i, tKeys[i], tdbr->getData(i, thread_idx), tdbr->getSeqLen(i));
                memcpy(tRes.data() + tOff[i], dbSeq.numSequence, (size_t) dbSeq.L);
            }
        }
        tMap = since(tStart) - tCreate;
        rc = b200_multi_db_load(multi, tRes.data(), tOff.data(), nT, A, /*shard_targets=*/0);
    }
    if (rc != B200_OK) {
        Debug(Debug::ERROR) << "libb200align: b200_db_load failed: " << b200_multi_last_error(multi) << "\n";
        EXIT(EXIT_FAILURE);
    }
    const double tLoad = since(tStart) - tCreate - tMap;
    double tParse = 0, tDevice = 0, tWrite = 0;

    std::vector<int16_t> mat((size_t) A * A);
    for (int i = 0; i < A; i++) {
        for (int j = 0; j < A; j++) {
            mat[(size_t) i * A + j] = m->subMatrix[i][j];
        }
    }

    b200_align_params p;
    memset(&p, 0, sizeof(p));
    p.gap_open = s.gapOpen; p.gap_extend = s.gapExtend; p.sw_mode = (int) s.swMode; p.eval_thr = s.evalThr;
    p.cov_thr = (float) s.covThr; p.cov_mode = s.covMode; p.seq_id_thr = (float) s.seqIdThr; p.aln_len_thr = s.alnLenThr;
    p.seq_id_mode = s.seqIdMode; p.max_accept = s.maxAccept; p.max_rejected = s.maxReject;
    p.comp_bias = s.compBiasCorrection ? 1 : 0; p.comp_bias_scale = s.compBiasCorrectionScale;
    p.include_identity = (s.includeIdentity || s.sameQTDB) ? 1 : 0;

    // ---- buckets of queries ---------------------------------------------------------------------------------------------
    const size_t maxBucketHits = 4u << 20;          // ~4 M (query, hit) pairs per device call
    size_t id = dbFrom;
    const size_t idEnd = dbFrom + dbSize;
    Debug::Progress progress(dbSize);
    std::vector<uint64_t> qOff, hOff;
    std::vector<uint32_t> qKeys, hTargets, nResults;
    std::vector<uint8_t> qRes;
    std::vector<b200_result> results;
    std::vector<char> btPool;
    std::vector<std::vector<uint32_t> > perQueryHits;
    std::vector<std::vector<uint8_t> > perQuerySeq;
    while (id < idEnd) {
        // bucket boundary: by the prefilter entries' byte size (about 12 bytes per record)
        size_t end = id, bytes = 0;
        while (end < idEnd && (end == id || bytes + prefdbr->getEntryLen(end) < maxBucketHits * 12)) {
            bytes += prefdbr->getEntryLen(end);
            end++;
        }
        const size_t nQ = end - id;
        const Clock::time_point tb0 = Clock::now();
        perQueryHits.assign(nQ, std::vector<uint32_t>());
        perQuerySeq.assign(nQ, std::vector<uint8_t>());
        qKeys.resize(nQ);
#pragma omp parallel num_threads(threads)
        {
            unsigned int thread_idx = 0;
#ifdef OPENMP
            thread_idx = static_cast<unsigned int>(omp_get_thread_num());
#endif
            Sequence qSeq(s.maxSeqLen, querySeqType, m, 0, false, s.compBiasCorrection);
            char buffer[1024];
#pragma omp for schedule(dynamic, 16)
            for (size_t k = 0; k < nQ; k++) {
                char *data = prefdbr->getData(id + k, thread_idx);
                const DBKeyType queryDbKey = prefdbr->getDbKey(id + k);
                qKeys[k] = (uint32_t) queryDbKey;
                if (*data == '\0') continue;
                const size_t qId = qdbr->getId(queryDbKey);
                char *querySeqData = qdbr->getData(qId, thread_idx);
                if (querySeqData == NULL) {
                    Debug(Debug::ERROR) << "Query sequence " << queryDbKey
                                        << " is required in the prefiltering, but is not contained in the query sequence database.\nPlease check your database.\n";
                    EXIT(EXIT_FAILURE);
                }
                qSeq.mapSequence(qId, queryDbKey, querySeqData, qdbr->getSeqLen(qId));
                perQuerySeq[k].assign(qSeq.numSequence, qSeq.numSequence + qSeq.L);
                while (*data != '\0') {
                    Util::parseKey(data, buffer);
                    const DBKeyType dbKey = Util::fast_atoi<DBKeyType>(buffer);
                    data = Util::skipLine(data);
                    const size_t dbId = tdbr->getId(dbKey);
                    if (dbId == DB_ENTRY_NOT_FOUND) {
                        Debug(Debug::ERROR) << "Sequence " << dbKey << " is required in the prefiltering, but is not contained in the target sequence database!\nPlease check your database.\n";
                        EXIT(EXIT_FAILURE);
                    }
                    perQueryHits[k].push_back((uint32_t) dbId);
                }
            }
        }
        qOff.assign(nQ + 1, 0);
        hOff.assign(nQ + 1, 0);
        for (size_t k = 0; k < nQ; k++) {
            qOff[k + 1] = qOff[k] + perQuerySeq[k].size();
            hOff[k + 1] = hOff[k] + perQueryHits[k].size();
        }
        qRes.resize(qOff[nQ] + 1);
        hTargets.resize(hOff[nQ] + 1);
        for (size_t k = 0; k < nQ; k++) {
            if (!perQuerySeq[k].empty()) memcpy(qRes.data() + qOff[k], perQuerySeq[k].data(), perQuerySeq[k].size());
            if (!perQueryHits[k].empty()) memcpy(hTargets.data() + hOff[k], perQueryHits[k].data(), perQueryHits[k].size() * sizeof(uint32_t));
        }
        const size_t nHits = hOff[nQ];
        results.resize(nHits + 1);
        nResults.assign(nQ, 0);
        uint64_t nAln = 0;
        tParse += since(tb0);
        const Clock::time_point tb1 = Clock::now();
        if (nHits > 0) {
            uint64_t btCap = s.swMode == Matcher::SCORE_COV_SEQID ? (uint64_t) 64 << 20 : 16;
            while (true) {
                btPool.resize(btCap);
                rc = b200_multi_align_batch(multi, mat.data(), m->pBack, A, qRes.data(), qOff.data(), qKeys.data(), (uint32_t) nQ, hOff.data(),
                                      hTargets.data(), tKeys.data(), &p, &ev, results.data(), nResults.data(), btPool.data(), btCap, &nAln);
                if (rc == B200_ERR_RANGE && btCap < ((uint64_t) 1 << 36)) { btCap *= 4; continue; }   // backtrace pool too small
                break;
            }
            if (rc != B200_OK) {
                Debug(Debug::ERROR) << "libb200align: b200_align_batch failed: " << b200_multi_last_error(multi) << "\n";
                EXIT(EXIT_FAILURE);
            }
        }
        alignmentsNum += nAln;
        tDevice += since(tb1);
        const Clock::time_point tb2 = Clock::now();
        // ---- records: Matcher::resultToBuffer on the reference's own result_t ---------------------------------------------
#pragma omp parallel num_threads(threads)
        {
            unsigned int thread_idx = 0;
#ifdef OPENMP
            thread_idx = static_cast<unsigned int>(omp_get_thread_num());
#endif
            std::string out;
            out.reserve(1024 * 1024);
            char buffer[1024 + 32768 * 4];
            size_t passed = 0;
#pragma omp for schedule(dynamic, 16)
            for (size_t k = 0; k < nQ; k++) {
                for (uint32_t r = 0; r < nResults[k]; r++) {
                    const b200_result &b = results[hOff[k] + r];
                    std::string bt;
                    if (b.bt_len > 0) bt.assign(btPool.data() + b.bt_off, b.bt_len);
                    Matcher::result_t res(b.db_key, b.score, b.qcov, b.dbcov, b.seq_id, b.eval, b.aln_length, b.q_start, b.q_end, b.q_len,
                                          b.db_start, b.db_end, b.db_len, bt);
                    const size_t len = Matcher::resultToBuffer(buffer, res, s.addBacktrace);
                    out.append(buffer, len);
                    passed++;
                }
                dbw.writeData(out.c_str(), out.length(), qKeys[k], thread_idx);
                out.clear();
                progress.updateProgress();
            }
#pragma omp atomic
            totalPassedNum += passed;
        }
        tWrite += since(tb2);
        id = end;
    }
    Debug(Debug::INFO) << "libb200align align module: context " << tCreate << " s, target DB mapping " << tMap << " s, upload " << tLoad
                       << " s, prefilter lists + queries " << tParse << " s, b200_align_batch " << tDevice << " s, records " << tWrite << " s\n";
    b200_multi_destroy(multi);
    return true;
}

}  // namespace b200shim
#endif
