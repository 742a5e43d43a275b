// mmseqs2_b200/csrc/b200_paddeddb.cpp -- writer of the padded GPU sequence DB (`mmseqs makepaddedseqdb`, src/util/makepaddedseqdb.cpp:14-153)
// and the repeat masker it runs on every target (Masker::maskSequence, src/commons/Masker.cpp:16-58, over lib/tantan).  Host code only;
// declared in include/b200_db.h.  The layout written here is what b200_db_load_padded (b200_align.cu) and the reference's own GPU path read.
#include "b200_db.h"

#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---- repeat masker ------------------------------------------------------------------------------------------------------------------
// tantan's HMM (Frith 2011) as Masker.cpp:22-32 configures it: one background state, W = 50 repeat states (period 1..W), no indel
// states (firstGapProb = otherGapProb = 0).  Forward pass, backward pass, posterior of "not in a repeat" per letter.  The sums over
// the repeat states are accumulated the way the reference's AVX2 build does (four interleaved partial sums, then (s0+s2)+(s1+s3),
// then the scalar tail: lib/tantan/tantan.cpp:306-347, mcf_simd.h:175-179) so that the doubles -- and with them every >= threshold
// decision -- come out the same; compiled with -ffp-contract=off like the rest of the library.
struct RepeatHmm {
    static constexpr int W = 50;              // maxRepeatOffset
    static constexpr int kScaleStep = 16;     // probabilities are renormalised every 16 letters
    double b2b, f2b, f2f, b2f[W];

    RepeatHmm() {
        const double repeatProb = 0.005, repeatEndProb = 0.05, decay = 0.9;
        b2b = 1 - repeatProb;
        f2b = repeatEndProb;
        f2f = 1 - repeatEndProb;
        double p = repeatProb * ((1 - decay) / (1 - std::pow(decay, W)));
        for (int i = 0; i < W; i++) { b2f[i] = p; p *= decay; }
    }

    // probs[p] = posterior probability that letter p lies in a repeat (float, as the reference stores it)
    void posteriors(const uint8_t *seq, int L, int A, const double *lr, float *probs, std::vector<double> &scale) const {
        double bg = 1.0, fg[W];
        for (int i = 0; i < W; i++) fg[i] = 0.0;
        scale.assign((size_t) L / kScaleStep, 0.0);
        for (int p = 0; p < L; p++) {
            const double *row = lr + (size_t) seq[p] * A;
            const int m = p < W ? p : W;
            const double b = bg;
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            int i = 0;
            for (; i <= m - 4; i += 4) {
                const double f0 = fg[i], f1 = fg[i + 1], f2 = fg[i + 2], f3 = fg[i + 3];
                s0 += f0; s1 += f1; s2 += f2; s3 += f3;
                fg[i] = (b * b2f[i] + f0 * f2f) * row[seq[p - 1 - i]];
                fg[i + 1] = (b * b2f[i + 1] + f1 * f2f) * row[seq[p - 2 - i]];
                fg[i + 2] = (b * b2f[i + 2] + f2 * f2f) * row[seq[p - 3 - i]];
                fg[i + 3] = (b * b2f[i + 3] + f3 * f2f) * row[seq[p - 4 - i]];
            }
            double from = (s0 + s2) + (s1 + s3);
            for (; i < m; i++) {
                const double f = fg[i];
                from += f;
                fg[i] = (b * b2f[i] + f * f2f) * row[seq[p - 1 - i]];
            }
            bg = b * b2b + from * f2b;
            if (p % kScaleStep == kScaleStep - 1) {
                const double sc = 1 / bg;
                scale[(size_t) p / kScaleStep] = sc;
                bg *= sc;
                for (int k = 0; k < W; k++) fg[k] *= sc;
            }
            probs[p] = (float) bg;
        }
        double tot = 0.0;
        for (int i = 0; i < W; i++) tot += fg[i];
        const double z = bg * b2b + tot * f2b;

        bg = b2b;
        for (int i = 0; i < W; i++) fg[i] = f2b;
        for (int p = L - 1; p >= 0; p--) {
            const double non_repeat = probs[p] * bg / z;
            probs[p] = 1 - (float) non_repeat;
            if (p % kScaleStep == kScaleStep - 1) {
                const double sc = scale[(size_t) p / kScaleStep];
                bg *= sc;
                for (int k = 0; k < W; k++) fg[k] *= sc;
            }
            const double *row = lr + (size_t) seq[p] * A;
            const int m = p < W ? p : W;
            const double to_bg = f2b * bg;
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            int i = 0;
            for (; i <= m - 4; i += 4) {
                const double f0 = fg[i] * row[seq[p - 1 - i]], f1 = fg[i + 1] * row[seq[p - 2 - i]];
                const double f2 = fg[i + 2] * row[seq[p - 3 - i]], f3 = fg[i + 3] * row[seq[p - 4 - i]];
                s0 += b2f[i] * f0; s1 += b2f[i + 1] * f1; s2 += b2f[i + 2] * f2; s3 += b2f[i + 3] * f3;
                fg[i] = to_bg + f2f * f0; fg[i + 1] = to_bg + f2f * f1; fg[i + 2] = to_bg + f2f * f2; fg[i + 3] = to_bg + f2f * f3;
            }
            double to = (s0 + s2) + (s1 + s3);
            for (; i < m; i++) {
                const double f = fg[i] * row[seq[p - 1 - i]];
                to += b2f[i] * f;
                fg[i] = to_bg + f2f * f;
            }
            bg = b2b * bg + to;
        }
    }
};

// Masker::maskRepeats (Masker.cpp:83-118): runs of more than n identical letters become the mask letter
void mask_runs(uint8_t *seq, int L, int n, uint8_t mask) {
    int run = 0, start = -1, prev = 0;      // the reference starts with previousChar = '\0' (code 0) and a run of 0
    for (int p = 0; p < L; p++) {
        if (seq[p] == prev) { run++; continue; }
        if (run > n && start >= 0) for (int k = start; k < p; k++) seq[k] = mask;
        run = 1; start = p; prev = seq[p];
    }
    // a sequence that opens with code-0 letters extends the initial "run" without a start position: the reference then loops from
    // index -1 as an unsigned (4294967295 < pos is false: nothing is written); start stays -1 here and the same run is skipped
    if (run > n && start >= 0) for (int k = start; k < L; k++) seq[k] = mask;
}

thread_local std::string g_pad_err;
int pad_fail(const std::string &msg) { g_pad_err = msg; return B200_ERR_ARG; }

bool write_file(const std::string &path, const void *data, size_t n) {
    FILE *f = fopen(path.c_str(), "wb");
    if (f == nullptr) return false;
    const bool ok = (n == 0 || fwrite(data, 1, n, f) == n);
    return (fclose(f) == 0) && ok;
}

bool starts_with(const std::string &s, size_t off, const char *prefix) {
    const size_t n = strlen(prefix);
    return s.size() >= n && s.compare(off, n, prefix) == 0;       // Util::startWith: the length test ignores the offset
}

// Util::parseFastaHeader (src/commons/Util.cpp:147-229): the accession inside the first word of a header
std::string accession_of(const char *header) {
    size_t len = 0;
    while (!(header[len] == ' ' || header[len] == '\t' || header[len] == '\n' || header[len] == '\0')) len++;
    const std::string h(header, len);
    if (h.empty()) return "";
    size_t off = starts_with(h, 0, "consensus_") ? 10 : 0;
    static const struct { const char *prefix; int bars; } kDatabases[] = {
        {"cl|", 1}, {"sp|", 1}, {"tr|", 1}, {"gb|", 1}, {"ref|", 1}, {"pdb|", 1}, {"bbs|", 1}, {"lcl|", 1}, {"pir||", 1}, {"prf||", 1},
        {"gnl|", 2}, {"pat|", 2}, {"gi|", 3}};
    for (const auto &d : kDatabases) {
        if (!starts_with(h, off, d.prefix)) continue;
        size_t start = off + strlen(d.prefix);
        for (int j = 0; j + 1 < d.bars; j++) {
            const size_t bar = h.find('|', start);
            if (bar == std::string::npos) return "";
            start = bar + 1;
        }
        size_t end = h.find('|', start);
        if (end == std::string::npos) end = h.find_first_of(" \n", start);
        if (end == std::string::npos) end = h.size();
        return h.substr(start, end - start);
    }
    size_t end = h.find_first_of(" \n", off);
    if (end == std::string::npos) end = h.size();
    return h.substr(off, end - off);
}

}  // namespace

extern "C" {

const char *b200h_paddeddb_last_error(void) { return g_pad_err.c_str(); }

int b200h_tantan_probabilities(const uint8_t *seq, int L, int alphabet, const double *likelihood_ratio, float *probs) {
    if (L < 0 || alphabet <= 0 || likelihood_ratio == nullptr || (L > 0 && (seq == nullptr || probs == nullptr))) return pad_fail("b200h_tantan_probabilities: bad arguments");
    for (int i = 0; i < L; i++)
        if (seq[i] >= alphabet) return pad_fail("b200h_tantan_probabilities: residue code outside the alphabet");
    static const RepeatHmm hmm;
    std::vector<double> scale;
    hmm.posteriors(seq, L, alphabet, likelihood_ratio, probs, scale);
    return B200_OK;
}

int b200h_mask_sequence(uint8_t *seq, const char *text, int L, int alphabet, const double *likelihood_ratio, int mask_tantan, double mask_prob,
                        int mask_lower_case, int mask_n_repeats) {
    if (L < 0 || alphabet <= 1 || (L > 0 && seq == nullptr)) return -1;
    const uint8_t x = (uint8_t) (alphabet - 1);                 // subMat.aa2num['X'] = the last code
    int masked = 0;
    if (mask_tantan) {
        if (likelihood_ratio == nullptr) return -1;
        std::vector<float> probs((size_t) L);
        if (b200h_tantan_probabilities(seq, L, alphabet, likelihood_ratio, probs.data()) != B200_OK) return -1;
        for (int i = 0; i < L; i++)
            if (probs[i] >= mask_prob) { seq[i] = x; masked++; }
    }
    if (mask_n_repeats > 0) {
        const int before = (int) std::count(seq, seq + L, x);
        mask_runs(seq, L, mask_n_repeats, x);
        masked += (int) std::count(seq, seq + L, x) - before;
    }
    if (mask_lower_case && text != nullptr)
        for (int i = 0; i < L; i++)
            if (std::islower((unsigned char) text[i])) { seq[i] = x; masked++; }
    return masked;
}

int b200h_make_padded_db(const char *src_db, const char *dst_db, const uint8_t aa2num[256], int alphabet, const double *likelihood_ratio,
                         int mask_mode, double mask_prob, int mask_lower_case, int mask_n_repeats, int write_lookup, int threads) {
    if (src_db == nullptr || dst_db == nullptr || aa2num == nullptr || alphabet != 21) return pad_fail("b200h_make_padded_db: bad arguments (amino-acid alphabet of 21 letters expected)");
    if (mask_mode && likelihood_ratio == nullptr) return pad_fail("b200h_make_padded_db: mask_mode needs the likelihood-ratio matrix");
    b200h_db *seqs = nullptr, *hdrs = nullptr;
    if (b200h_db_open(src_db, &seqs) != B200_OK) return pad_fail(b200h_db_last_error());
    if (b200h_db_open((std::string(src_db) + "_h").c_str(), &hdrs) != B200_OK) { b200h_db_close(seqs); return pad_fail(b200h_db_last_error()); }
    struct Closer { b200h_db *a, *b; ~Closer() { b200h_db_close(a); b200h_db_close(b); } } closer{seqs, hdrs};
    const int src_type = b200h_db_type(seqs);
    if ((src_type & 0xffff) != B200_DBTYPE_AMINO_ACIDS) return pad_fail("b200h_make_padded_db: amino-acid sequence DB expected");
    const uint64_t n = b200h_db_size(seqs);

    // DBReader::SORT_BY_LENGTH (DBReaderSortIndex.cpp:110-126): (index length desc, id asc); makepaddedseqdb walks it backwards
    std::vector<uint64_t> order(n);
    for (uint64_t i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
        const uint32_t la = (uint32_t) b200h_db_entry_len(seqs, a), lb = (uint32_t) b200h_db_entry_len(seqs, b);
        return la != lb ? la > lb : a < b;
    });
    std::reverse(order.begin(), order.end());

    // per-entry output: codes (+32 where masked / lower case), padded with code 20 to a multiple of 4 (makepaddedseqdb.cpp:62-88)
    std::vector<uint64_t> off(n + 1, 0);
    std::vector<int64_t> hdr_id(n);
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t len = b200h_db_entry_len(seqs, order[i]);
        const uint64_t L = (uint32_t) len >= 2 ? (uint32_t) len - 2 : 0;                // DBReader::getSeqLen: without newline and NUL
        off[i + 1] = off[i] + ((L + 3) & ~(uint64_t) 3);
        hdr_id[i] = b200h_db_id(hdrs, b200h_db_key(seqs, order[i]));
        if (hdr_id[i] < 0) return pad_fail("b200h_make_padded_db: no header for key " + std::to_string(b200h_db_key(seqs, order[i])));
    }
    std::vector<uint8_t> out(off[n]);
    const int nt = std::max(1, std::min(threads, 256));
    std::atomic<uint64_t> next(0);
    std::atomic<int> bad(0);
    auto work = [&]() {
        std::vector<uint8_t> codes, masked;
        for (;;) {
            const uint64_t lo = next.fetch_add(256);
            if (lo >= n) break;
            for (uint64_t i = lo; i < std::min(n, lo + 256); i++) {
                const char *text = b200h_db_data(seqs, order[i]);
                const uint64_t len = b200h_db_entry_len(seqs, order[i]);
                const int L = (uint32_t) len >= 2 ? (int) ((uint32_t) len - 2) : 0;
                uint8_t *dst = out.data() + off[i];
                codes.resize((size_t) L);
                for (int k = 0; k < L; k++) codes[k] = aa2num[(unsigned char) text[k]];
                if (mask_mode) {
                    masked = codes;
                    if (b200h_mask_sequence(masked.data(), text, L, alphabet, likelihood_ratio, 1, mask_prob, mask_lower_case, mask_n_repeats) < 0) { bad = 1; return; }
                    for (int k = 0; k < L; k++) dst[k] = (uint8_t) (masked[k] == alphabet - 1 ? codes[k] + 32 : codes[k]);
                } else {
                    for (int k = 0; k < L; k++) dst[k] = (uint8_t) (std::islower((unsigned char) text[k]) ? codes[k] + 32 : codes[k]);
                }
                for (uint64_t k = (uint64_t) L; k < off[i + 1] - off[i]; k++) dst[k] = 20;
            }
        }
    };
    if (nt == 1) work();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    if (bad) return pad_fail("b200h_make_padded_db: residue code outside the alphabet (aa2num table and alphabet disagree)");

    const std::string dst(dst_db);
    std::string index, hindex, hdata, lookup;
    char line[128];
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t len = b200h_db_entry_len(seqs, order[i]);
        const uint64_t L = (uint32_t) len >= 2 ? (uint32_t) len - 2 : 0;
        snprintf(line, sizeof(line), "%llu\t%llu\t%llu\n", (unsigned long long) i, (unsigned long long) off[i], (unsigned long long) (L + 2));
        index += line;
        const uint64_t hl = b200h_db_entry_len(hdrs, (uint64_t) hdr_id[i]);              // text + NUL, copied as it lies
        snprintf(line, sizeof(line), "%llu\t%llu\t%llu\n", (unsigned long long) i, (unsigned long long) hdata.size(), (unsigned long long) hl);
        hindex += line;
        const char *h = b200h_db_data(hdrs, (uint64_t) hdr_id[i]);
        hdata.append(h, hl);
        if (write_lookup) {
            // id, accession, and -- in the file-number column -- the key the sequence had in the source DB (makepaddedseqdb.cpp:121-127)
            lookup += std::to_string(i); lookup += '\t'; lookup += accession_of(h); lookup += '\t';
            lookup += std::to_string(b200h_db_key(seqs, order[i])); lookup += '\n';
        }
    }
    const int32_t seq_type = (int32_t) (((uint32_t) src_type | (8u << 16)) & 0x7FFFFFFFu);   // DBReader::setExtendedDbtype(.., DBTYPE_EXTENDED_GPU)
    const int32_t hdr_type = 12;                                                          // Parameters::DBTYPE_GENERIC_DB
    bool ok = write_file(dst, out.data(), out.size()) && write_file(dst + ".index", index.data(), index.size()) &&
              write_file(dst + ".dbtype", &seq_type, 4) && write_file(dst + "_h", hdata.data(), hdata.size()) &&
              write_file(dst + "_h.index", hindex.data(), hindex.size()) && write_file(dst + "_h.dbtype", &hdr_type, 4);
    if (ok && write_lookup) {
        ok = write_file(dst + ".lookup", lookup.data(), lookup.size());
        FILE *s = fopen((std::string(src_db) + ".source").c_str(), "rb");                // carried over when present (:147-150)
        if (ok && s != nullptr) {
            std::string src;
            char buf[4096];
            size_t r;
            while ((r = fread(buf, 1, sizeof(buf), s)) > 0) src.append(buf, r);
            ok = write_file(dst + ".source", src.data(), src.size());
        }
        if (s != nullptr) fclose(s);
    }
    if (!ok) return pad_fail("b200h_make_padded_db: cannot write " + dst);
    return B200_OK;
}

}  // extern "C"
