// tests/cpp/operators_smoke.cpp -- drives include/b200_mmseqs.hpp (the C++ operator mirror) the way the reference's call
// sites would, on inputs written by tests/test_gpu_cpp_operators.py; results go back as a flat int32 file.
//   g++ -std=c++11 -fno-exceptions -Iinclude tests/cpp/operators_smoke.cpp -Lmmseqs2_b200 -lb200align -o operators_smoke
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "b200_mmseqs.hpp"

template <typename T>
static std::vector<T> rd(FILE *f, size_t n) {
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}

static bool gate_score(uint32_t score, void *ud) { return score >= *(uint32_t *) ud; }

int main(int argc, char **argv) {
    if (argc < 3) return 1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    const std::vector<int32_t> hdr = rd<int32_t>(f, 5);  // A, nT, qlen, nHits, minScore
    const int A = hdr[0], nT = hdr[1], qlen = hdr[2], nHits = hdr[3];
    uint32_t minScore = (uint32_t) hdr[4];
    std::vector<int16_t> mat = rd<int16_t>(f, (size_t) A * A);
    std::vector<double> pback = rd<double>(f, A);
    std::vector<uint64_t> off64 = rd<uint64_t>(f, (size_t) nT + 1);
    std::vector<uint8_t> res = rd<uint8_t>(f, (size_t) off64[nT]);
    std::vector<uint8_t> q = rd<uint8_t>(f, qlen);
    std::vector<uint32_t> hitIds = rd<uint32_t>(f, nHits);
    std::vector<uint16_t> hitDiags = rd<uint16_t>(f, nHits);
    fclose(f);

    b200::Device dev(0);
    if (!dev.ok()) { fprintf(stderr, "no device: %s\n", dev.error()); return 3; }
    std::vector<size_t> off(off64.begin(), off64.end());
    if (dev.loadLookup(res.data(), off.data(), nT, A) != B200_OK) { fprintf(stderr, "%s\n", dev.error()); return 4; }
    std::vector<int32_t> dbLen(nT);
    for (int i = 0; i < nT; i++) dbLen[i] = (int32_t) (off[i + 1] - off[i]);

    std::vector<int32_t> out;
    // --- Matcher-style: ssw_init + whole hit list
    b200::SmithWaterman sw(&dev, mat.data(), pback.data(), A, true, 1.0f);
    if (sw.ssw_init(q.data(), qlen) != 0) return 5;
    for (int i = 0; i < nT; i++) sw.addTarget((uint32_t) i);
    std::vector<b200::s_align> aln;
    if (sw.flush(dbLen.data(), 11, 1, 2, gate_score, NULL, &minScore, aln) != B200_OK) { fprintf(stderr, "%s\n", dev.error()); return 6; }
    for (int i = 0; i < nT; i++) {
        out.push_back((int32_t) aln[i].score1); out.push_back(aln[i].qStartPos1); out.push_back(aln[i].qEndPos1);
        out.push_back(aln[i].dbStartPos1); out.push_back(aln[i].dbEndPos1); out.push_back(aln[i].word);
        out.push_back((int32_t) aln[i].identicalAACnt); out.push_back((int32_t) aln[i].backtrace.size());
    }
    // --- Marv-style scan with the profile ungappedprefilter.cpp:195-203 builds
    std::vector<int8_t> pssm((size_t) A * qlen);
    b200h_build_profile(mat.data(), A, q.data(), qlen, sw.compositionBias(), 1, pssm.data());
    b200::Marv marv(&dev, nT, A, 0, 50);
    marv.setMinScore(15);
    std::vector<b200::Marv::Result> results(50);
    b200::Marv::Stats st = marv.scan((const char *) q.data(), qlen, pssm.data(), results.data(), sw.bias());
    out.push_back((int32_t) st.results);
    for (size_t i = 0; i < 50; i++) { out.push_back(i < st.results ? (int32_t) results[i].id : -1); out.push_back(i < st.results ? results[i].score : -1); }
    // --- Marv in GAPLESS_SMITH_WATERMAN mode (search --gpu 1 --alignment-mode 1): gapped score + end positions of the best ungapped hits
    b200::Marv marv2(&dev, nT, A, 0, 50, b200::Marv::GAPLESS_SMITH_WATERMAN);
    marv2.setMinScore(15);
    std::vector<b200::Marv::Result> results2(50);
    b200::Marv::Stats st2 = marv2.scan((const char *) q.data(), qlen, pssm.data(), results2.data(), sw.bias());
    out.push_back((int32_t) st2.results);
    for (size_t i = 0; i < 50; i++) {
        const bool on = i < st2.results;
        out.push_back(on ? (int32_t) results2[i].id : -1); out.push_back(on ? results2[i].score : -1);
        out.push_back(on ? results2[i].qEndPos : -1); out.push_back(on ? results2[i].dbEndPos : -1);
    }
    // --- UngappedAlignment-style
    b200::UngappedAlignment ua(&dev, mat.data(), A);
    std::vector<float> fb(qlen);
    b200h_comp_bias(mat.data(), pback.data(), A, q.data(), qlen, 1.0f, fb.data());
    ua.createProfile(q.data(), qlen, fb.data());
    std::vector<b200::CounterResult> hits(nHits);
    for (int i = 0; i < nHits; i++) { hits[i].id = hitIds[i]; hits[i].diagonal = hitDiags[i]; hits[i].count = 0; }
    if (ua.align(hits.data(), nHits) != B200_OK) return 7;
    for (int i = 0; i < nHits; i++) out.push_back(hits[i].count);

    FILE *o = fopen(argv[2], "wb");
    fwrite(out.data(), sizeof(int32_t), out.size(), o);
    fclose(o);
    return 0;
}
