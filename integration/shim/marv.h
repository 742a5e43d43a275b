// integration/shim/marv.h -- drop-in for lib/libmarv/src/marv.h (the reference's GPU scorer interface, class Marv :6-58),
// served by libb200align.so.  With -DENABLE_B200=1 (integration/mmseqs_b200.patch) this directory replaces lib/libmarv/src on the
// include path, so src/prefiltering/ungappedprefilter.cpp, src/util/gpuserver.cpp and src/commons/GpuUtil.h compile against it
// UNCHANGED apart from one added call, setScoringMatrix(), after each construction.
//
// What differs from libmarv, on purpose: scan() returns the scores of the reference's CPU scorer (SmithWaterman::ungapped_alignment,
// saturating u8 with the SSW bias) ordered (score desc, id asc), i.e. the hit lists of runFilterOnCpu
// (ungappedprefilter.cpp:346-482), not libmarv's unsaturated half2/short2 scores.  To do that it needs the SSW profile bias
// |min(mat)| + |min(0, min compBias)| (StripedSmithWaterman.cpp:1375-1406), which it derives from the query, its profile and the
// substitution matrix given to setScoringMatrix().
//
// Header-only, C++11, no exceptions (the host is built -fno-exceptions), no CUDA headers.
#ifndef MARV_H
#define MARV_H

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "b200_align.h"
#include "b200_host.h"
#include "b200_multi.h"

class Marv {
public:
    enum AlignmentType {
        GAPLESS,
        SMITH_WATERMAN,
        GAPLESS_SMITH_WATERMAN
    };

    struct Stats {
        size_t results;
        int numOverflows;
        double seconds;
        double gcups;
    };

    struct Result {
        unsigned int id;
        int score;
        int qEndPos;
        int dbEndPos;

        Result(unsigned int id, int score, int qEndPos, int dbEndPos) :
            id(id), score(score), qEndPos(qEndPos), dbEndPos(dbEndPos) {};
    };

    Marv(size_t dbEntries, int alphabetSize, int maxSeqLength, size_t maxSeqs, AlignmentType alignmentType = AlignmentType::GAPLESS)
        : dbEntries(dbEntries), alphabetSize(alphabetSize), maxSeqs(maxSeqs), alignmentType(alignmentType), multi(NULL), ctx(NULL),
          gapOpen(11), gapExtend(1) {
        (void) maxSeqLength;
        t0 = std::chrono::steady_clock::now();
        // Like libmarv, every visible device is used (CUDA_VISIBLE_DEVICES selects them): the DB is cut into per-device slices and
        // each query is scanned against all slices at once (target-sharded, include/b200_multi.h).  B200_DEVICE=<id> pins one device;
        // the gapped rescoring mode needs the whole DB next to the scan and also stays on one device.
        const char *dev = getenv("B200_DEVICE");
        int rc;
        if (dev != NULL || alignmentType != GAPLESS) {
            const int id = dev != NULL ? atoi(dev) : 0;
            rc = b200_multi_create(&id, 1, &multi);
        } else {
            rc = b200_multi_create(NULL, 0, &multi);
        }
        if (rc != B200_OK) {
            fprintf(stderr, "libb200align: b200_multi_create failed (%d)\n", rc);
            exit(EXIT_FAILURE);
        }
        ctx = b200_multi_ctx(multi, 0);
        tCreate = elapsed();
    }

    ~Marv() {
        if (getenv("B200_TRACE") != NULL) {
            fprintf(stderr, "libb200align Marv: context %.3f s, loadDb %.3f s, %zu scans in %.3f s (%.3f ms each), lifetime %.3f s\n", tCreate,
                    tLoad, nScans, tScan, nScans ? 1e3 * tScan / nScans : 0.0, elapsed());
        }
        if (multi != NULL) {
            b200_multi_destroy(multi);
        }
    }

    static std::vector<int> getDeviceIds() {
        std::vector<int> ids;
        const int n = b200_device_count();
        for (int i = 0; i < n; i++) {
            ids.push_back(i);
        }
        return ids;
    }

    // data/offset/length: the padded GPU DB exactly as ungappedprefilter.cpp:127-139 / gpuserver.cpp:41-49 hand it over
    // (makepaddedseqdb layout: numeric codes, +32 = masked, entries padded to x4 with code 20, sorted by length).
    // Masked residues become X, as the CPU scorer treats them (ungappedprefilter.cpp:401-404).
    void* loadDb(char* data, size_t* offset, int32_t* length, size_t dbByteSize) {
        (void) dbByteSize;
        const double a = elapsed();
        const int rc = b200_multi_db_load_padded(multi, reinterpret_cast<const uint8_t*>(data), offset, length, dbEntries, alphabetSize,
                                                 /*shard_targets=*/1);
        checkMulti(rc, "loadDb");
        tLoad = elapsed() - a;
        return this;
    }

    // The remaining members exist for source compatibility with marv.h; the B200 context owns one resident copy of the DB.
    void* loadDb(char* data, size_t dbByteSize, void* otherdb) { (void) data; (void) dbByteSize; return otherdb; }
    void setDb(void* dbhandle) { (void) dbhandle; }
    void setDbWithAllocation(void* dbhandle, const std::string& allocationinfo) { (void) dbhandle; (void) allocationinfo; }
    std::string getDbMemoryHandle() { return std::string(); }
    void printInfo() {}
    void prefetch() {}
    void startTimer() {}
    void stopTimer() {}

    // B200 addition: BaseMatrix::subMatrix of the matrix the profiles are built from (ungappedprefilter.cpp:195-203)
    void setScoringMatrix(short** subMatrix, int alphabet) {
        mat.resize((size_t) alphabet * alphabet);
        for (int i = 0; i < alphabet; i++) {
            for (int j = 0; j < alphabet; j++) {
                mat[(size_t) i * alphabet + j] = subMatrix[i][j];
            }
        }
    }
    // gap costs of the GAPLESS_SMITH_WATERMAN rescoring (libmarv hard-codes BLOSUM62 11/1)
    void setGapCosts(int open, int extend) { gapOpen = open; gapExtend = extend; }

    //sequence must be encoded
    Stats scan(const char* sequence, size_t sequenceLength, int8_t* pssm, Result* results) {
        Stats st;
        st.results = 0; st.numOverflows = 0; st.seconds = 0; st.gcups = 0;
        const double a = elapsed();
        b200_query q;
        q.profile = pssm;
        q.qlen = (int32_t) sequenceLength;
        q.bias = b200h_ssw_bias_from_profile(mat.empty() ? NULL : mat.data(), alphabetSize,
                                             reinterpret_cast<const uint8_t*>(sequence), (int) sequenceLength, pssm);
        hits.resize(maxSeqs);
        uint32_t n = 0;
        // libmarv returns the maxSeqs best targets whatever their score; the caller filters by --min-ungapped-score
        int rc = b200_multi_ungapped_scan(multi, &q, 1, /*min_score_excl=*/-1, (uint32_t) maxSeqs, hits.data(), &n);
        checkMulti(rc, "scan");
        if (alignmentType == GAPLESS) {
            for (uint32_t i = 0; i < n; i++) {
                results[i] = Result(hits[i].id, hits[i].score, -1, -1);
            }
        } else {
            // GAPLESS_SMITH_WATERMAN (cudasw4.cuh:1003-1026): gapped score + end positions of the best ungapped hits
            pairs.resize(n);
            ends.resize(n);
            for (uint32_t i = 0; i < n; i++) {
                pairs[i].query = 0;
                pairs[i].target = hits[i].id;
            }
            if (n > 0) {
                rc = b200_sw_score_endpos(ctx, &q, 1, pairs.data(), n, gapOpen, gapExtend, ends.data());
                check(rc, "scan (gapped rescoring)");
            }
            order.resize(n);
            for (uint32_t i = 0; i < n; i++) {
                order[i] = i;
            }
            const b200_sw_end *e = ends.data();
            const b200_hit *h = hits.data();
            std::sort(order.begin(), order.end(), [e, h](uint32_t a, uint32_t b) {
                return e[a].score != e[b].score ? e[a].score > e[b].score : h[a].id < h[b].id;
            });
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t k = order[i];
                results[i] = Result(hits[k].id, ends[k].score, ends[k].qend, ends[k].dbend);
            }
        }
        st.results = n;
        st.seconds = elapsed() - a;
        tScan += st.seconds;
        nScans++;
        return st;
    }

private:
    double elapsed() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
    std::chrono::steady_clock::time_point t0;
    double tCreate = 0, tLoad = 0, tScan = 0;
    size_t nScans = 0;

    void checkMulti(int rc, const char *what) {
        if (rc != B200_OK) {
            fprintf(stderr, "libb200align: %s failed (%d): %s\n", what, rc, b200_multi_last_error(multi));
            exit(EXIT_FAILURE);
        }
    }

    void check(int rc, const char *what) {
        if (rc != B200_OK) {
            fprintf(stderr, "libb200align: %s failed (%d): %s\n", what, rc, b200_last_error(ctx));
            exit(EXIT_FAILURE);
        }
    }

    size_t dbEntries;
    int alphabetSize;
    size_t maxSeqs;
    AlignmentType alignmentType;
    b200_multi* multi;
    b200_ctx* ctx;            // device 0 of the handle (the gapped rescoring runs there)
    int gapOpen, gapExtend;
    std::vector<int16_t> mat;
    std::vector<b200_hit> hits;
    std::vector<b200_pair> pairs;
    std::vector<b200_sw_end> ends;
    std::vector<uint32_t> order;
};

#endif
