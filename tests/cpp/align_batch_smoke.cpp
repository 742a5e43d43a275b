// tests/cpp/align_batch_smoke.cpp -- the `align` step from a C++ host (what Alignment::run would call):
// b200_db_load + b200_align_batch + b200h_result_to_buffer, built with -fno-exceptions like the reference.
// in:  header int32[5] = {A, nTargets, nQueries, nHits, swMode}; int16 mat[A*A]; double pback[A]; uint64 toff[nT+1]; uint8 tres[];
//      uint64 qoff[nQ+1]; uint8 qres[]; uint64 hoff[nQ+1]; uint32 hitTargets[nHits]; uint32 targetKeys[nT]; uint32 queryKeys[nQ]
// out: for every query its alignment-DB entry (records as Matcher::resultToBuffer writes them) followed by one NUL byte
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "b200_alignment.h"

template <typename T> static std::vector<T> rd(FILE *f, size_t n) {
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}

int main(int argc, char **argv) {
    if (argc < 3) return 1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    const std::vector<int32_t> hdr = rd<int32_t>(f, 5);
    const int A = hdr[0], nT = hdr[1], nQ = hdr[2], nHits = hdr[3], swMode = hdr[4];
    std::vector<int16_t> mat = rd<int16_t>(f, (size_t) A * A);
    std::vector<double> pback = rd<double>(f, A);
    std::vector<uint64_t> toff = rd<uint64_t>(f, (size_t) nT + 1);
    std::vector<uint8_t> tres = rd<uint8_t>(f, (size_t) toff[nT]);
    std::vector<uint64_t> qoff = rd<uint64_t>(f, (size_t) nQ + 1);
    std::vector<uint8_t> qres = rd<uint8_t>(f, (size_t) qoff[nQ]);
    std::vector<uint64_t> hoff = rd<uint64_t>(f, (size_t) nQ + 1);
    std::vector<uint32_t> hits = rd<uint32_t>(f, nHits);
    std::vector<uint32_t> tkeys = rd<uint32_t>(f, nT);
    std::vector<uint32_t> qkeys = rd<uint32_t>(f, nQ);
    fclose(f);

    b200_ctx *ctx = NULL;
    if (b200_create(0, &ctx) != B200_OK) { fprintf(stderr, "no device\n"); return 3; }
    if (b200_db_load(ctx, tres.data(), toff.data(), (uint64_t) nT, A) != B200_OK) { fprintf(stderr, "%s\n", b200_last_error(ctx)); return 4; }

    b200_align_params p;
    p.gap_open = 11; p.gap_extend = 1; p.sw_mode = swMode; p.eval_thr = 1e-3; p.cov_thr = 0.0f; p.cov_mode = 0; p.seq_id_thr = 0.0f;
    p.aln_len_thr = 0; p.seq_id_mode = 0; p.max_accept = 0x7fffffff; p.max_rejected = 0x7fffffff; p.comp_bias = 1; p.comp_bias_scale = 1.0f;
    p.include_identity = 0;
    b200_evalue_params ev;
    if (b200h_evalue_defaults("blosum62.out", 11, 1, 1, toff[nT], &ev) != B200_OK) return 5;

    std::vector<b200_result> results((size_t) nHits + 1);
    std::vector<uint32_t> nres((size_t) nQ + 1);
    uint64_t btCap = 16, nAln = 0;
    for (int i = 0; i < nQ; i++)
        for (uint64_t k = hoff[i]; k < hoff[i + 1]; k++) btCap += (qoff[i + 1] - qoff[i]) + (toff[hits[k] + 1] - toff[hits[k]]);
    std::vector<char> pool(btCap);
    if (b200_align_batch(ctx, mat.data(), pback.data(), A, qres.data(), qoff.data(), qkeys.data(), (uint32_t) nQ, hoff.data(), hits.data(),
                         tkeys.data(), &p, &ev, results.data(), nres.data(), pool.data(), btCap, &nAln) != B200_OK) {
        fprintf(stderr, "%s\n", b200_last_error(ctx));
        return 6;
    }
    FILE *o = fopen(argv[2], "wb");
    if (!o) return 7;
    std::vector<char> buf;
    for (int i = 0; i < nQ; i++) {
        for (uint32_t k = 0; k < nres[i]; k++) {
            const b200_result &r = results[hoff[i] + k];
            buf.resize(512 + 2 * (size_t) r.bt_len);
            const size_t len = b200h_result_to_buffer(buf.data(), &r, pool.data() + r.bt_off, swMode == 2, 1);
            fwrite(buf.data(), 1, len, o);
        }
        fputc('\0', o);
    }
    fclose(o);
    fprintf(stderr, "%llu alignments\n", (unsigned long long) nAln);
    b200_destroy(ctx);
    return 0;
}
