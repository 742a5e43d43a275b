// baseline/marv_harness.cpp -- BASELINE / MEASUREMENT INFRASTRUCTURE, not product code.
//
// A C-ABI shim around the reference's own GPU scorer `class Marv` (lib/libmarv/src/marv.h:6-58), so that bench.py can
// drive Marv::scan through ctypes on the same padded target DB and the same queries it gives libb200align.so.
// The calls below are the ones src/prefiltering/ungappedprefilter.cpp:150-158,207 makes: Marv(dbEntries, alphabetSize,
// maxTargetLength, maxResListLen, type) -> loadDb(data, offsets, lengths, bytes) -> setDb -> scan(query, L, profile,
// results) once per query.  Nothing of libmarv is modified; it is compiled in place by baseline/Makefile.
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>
// cudasw4.cuh defines non-inline kernels, so it can live in one translation unit only: this file therefore INCLUDES the
// reference's lib/libmarv/src/marv.cu where it lies (it is not compiled separately, see baseline/Makefile) and adds the
// C entry points below.  The harness needs cudasw4::CudaSW4's public setCustomKernelConfig_Gapless (cudasw4.cuh:3669) to
// time the reference's DPX (short2) tables on a cc 10.0 device: the shipped selection (gapless_kernel_config.cuh:526-559)
// has no cc 10.0 entry and falls back to the sm_89 half2 table.  `class Marv` keeps its CudaSW4 pointer private.
#include <algorithm>
#include <memory>
#include <sstream>
#include <fstream>
#include <iostream>
#include <map>
#include <optional>
#include <numeric>
#include <thread>
#include <mutex>
#include <future>
#define private public
#include "marv.h"
#undef private
#include "marv.cu"

struct MarvHarness {
    Marv* marv = nullptr;
    void* db = nullptr;
    std::vector<Marv::Result> results;
    std::vector<cudasw4::GaplessKernelConfig> table;   // empty = whatever Marv selects on this device (as shipped)
};

extern "C" {

// type: 0 GAPLESS, 1 SMITH_WATERMAN, 2 GAPLESS_SMITH_WATERMAN (marv.h:8-12)
void* marvh_create(size_t db_entries, int alphabet, int max_target_len, size_t max_seqs, int type) {
    MarvHarness* h = new MarvHarness();
    h->marv = new Marv(db_entries, alphabet, max_target_len, max_seqs, static_cast<Marv::AlignmentType>(type));
    h->table.clear();
    h->results.reserve(max_seqs);
    for (size_t i = 0; i < max_seqs; i++) h->results.emplace_back(0u, 0, 0, 0);
    return h;
}

// Kernel table override: 0 = as shipped, 90 = the reference's H100 table (short2/DPX), 103 = its B300 table (short2/DPX),
// 89 = its sm_89 table (half2; what "as shipped" resolves to on cc 10.0).  The tables are the reference's own
// (gapless_kernel_config.cuh:156-355), obtained by calling its functions; per query the entry with the smallest
// tilesize >= qlen is installed through setCustomKernelConfig_Gapless, as getSingleTileGroupRegConfigForPSSM_Gapless would pick it.
int marvh_set_table(void* handle, int which) {
    MarvHarness* h = static_cast<MarvHarness*>(handle);
    switch (which) {
        case 0: h->table.clear(); break;
        case 89: h->table = cudasw4::getOptimalKernelConfigs_gapless_sm89(); break;
        case 90: h->table = cudasw4::getOptimalKernelConfigs_gapless_sm90(); break;
        case 103: h->table = cudasw4::getOptimalKernelConfigs_gapless_sm103(); break;
        default: return 1;
    }
    cudasw4::CudaSW4* sw = static_cast<cudasw4::CudaSW4*>(h->marv->cudasw);
    sw->useCustomKernelConfig_Gapless = false;
    return 0;
}

// data: the padded GPU DB bytes (makepaddedseqdb layout), offsets[n+1], lengths[n]; the arrays must outlive the handle.
int marvh_load_db(void* handle, char* data, size_t* offsets, int32_t* lengths, size_t db_bytes) {
    MarvHarness* h = static_cast<MarvHarness*>(handle);
    h->db = h->marv->loadDb(data, offsets, lengths, db_bytes);
    h->marv->setDb(h->db);
    h->marv->prefetch();
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}

// One Marv::scan.  out_ids/out_scores receive up to max_seqs results; stats[0] = seconds, stats[1] = gcups (libmarv's own
// Stats, cudasw4.cuh), stats[2] = overflows.  Returns the number of results, or -1.
long marvh_scan(void* handle, const char* query, size_t qlen, int8_t* pssm, uint32_t* out_ids, int32_t* out_scores,
                int32_t* out_qend, int32_t* out_dbend, double* stats) {
    MarvHarness* h = static_cast<MarvHarness*>(handle);
    if (!h->table.empty()) {
        cudasw4::CudaSW4* sw = static_cast<cudasw4::CudaSW4*>(h->marv->cudasw);
        const cudasw4::GaplessKernelConfig* pick = nullptr;
        for (const auto& c : h->table) if (c.tilesize >= (int)qlen && (!pick || c.tilesize < pick->tilesize)) pick = &c;
        if (pick) sw->setCustomKernelConfig_Gapless(*pick); else sw->useCustomKernelConfig_Gapless = false;
    }
    Marv::Stats st = h->marv->scan(query, qlen, pssm, h->results.data());
    size_t n = st.results < h->results.size() ? st.results : h->results.size();
    for (size_t i = 0; i < n; i++) {
        if (out_ids) out_ids[i] = h->results[i].id;
        if (out_scores) out_scores[i] = h->results[i].score;
        if (out_qend) out_qend[i] = h->results[i].qEndPos;
        if (out_dbend) out_dbend[i] = h->results[i].dbEndPos;
    }
    if (stats) { stats[0] = st.seconds; stats[1] = st.gcups; stats[2] = st.numOverflows; }
    return (long)n;
}

void marvh_destroy(void* handle) {
    MarvHarness* h = static_cast<MarvHarness*>(handle);
    if (!h) return;
    delete h->marv;
    delete h;
}

}  // extern "C"
