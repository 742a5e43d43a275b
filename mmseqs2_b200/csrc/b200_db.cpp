// mmseqs2_b200/csrc/b200_db.cpp -- DB triple reader / writer, letter mapping and the `align` module over DB files
// (include/b200_db.h).  Host code only.
#include "b200_db.h"

#include <algorithm>
#include <cctype>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "b200_host.h"
#include "b200_internal.h"

namespace {

thread_local std::string g_db_err;

int db_fail(const std::string &msg) { g_db_err = msg; return B200_ERR_ARG; }

bool read_file(const std::string &path, std::vector<char> &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (f == nullptr) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t) n : 0);
    const bool ok = out.empty() || fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

struct IndexEntry { uint32_t key; uint64_t offset; uint64_t length; };

// zstd streaming decompression, bound at run time: compressed DBs (createdb --compressed 1, `mmseqs compress`) are the only users,
// so the library carries no load-time dependency on libzstd.  The two buffer structs are zstd's stable public ABI (zstd.h,
// "Streaming decompression"): { pointer, size, pos }.
struct ZInBuf { const void *src; size_t size; size_t pos; };
struct ZOutBuf { void *dst; size_t size; size_t pos; };
struct Zstd {
    void *(*createDStream)() = nullptr;
    size_t (*initDStream)(void *) = nullptr;
    size_t (*decompressStream)(void *, ZOutBuf *, ZInBuf *) = nullptr;
    size_t (*freeDStream)(void *) = nullptr;
    unsigned (*isError)(size_t) = nullptr;
    bool ok = false;
};

const Zstd &zstd_api() {
    static Zstd z;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
        if (h == nullptr) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
        if (h == nullptr) return;
        z.createDStream = (void *(*)()) dlsym(h, "ZSTD_createDStream");
        z.initDStream = (size_t (*)(void *)) dlsym(h, "ZSTD_initDStream");
        z.decompressStream = (size_t (*)(void *, ZOutBuf *, ZInBuf *)) dlsym(h, "ZSTD_decompressStream");
        z.freeDStream = (size_t (*)(void *)) dlsym(h, "ZSTD_freeDStream");
        z.isError = (unsigned (*)(size_t)) dlsym(h, "ZSTD_isError");
        z.ok = z.createDStream && z.initDStream && z.decompressStream && z.freeDStream && z.isError;
    });
    return z;
}

// A compressed DB (dbtype bit 31) stores every entry as  [uint32 n][n payload bytes][marker]  with the index keeping the entry's
// offset and its UNCOMPRESSED length + 1 (DBWriter::writeEnd, src/commons/DBWriter.cpp:372-413): marker 0 = the payload is one zstd
// frame, anything else (0xFF) = the payload is the text itself, stored raw because it was shorter than 60 bytes.  The reader tells
// the two apart by that marker byte (DBReader::getDataCompressed, src/commons/DBReader.cpp:575-607).  Rewrites `data` / `index`
// into the plain layout (text + NUL back to back) the rest of this file works on.
bool inflate_db(std::vector<char> &data, std::vector<IndexEntry> &index, std::string &err) {
    const Zstd &z = zstd_api();
    if (!z.ok) { err = "compressed DB needs libzstd.so.1 at run time (dlopen failed)"; return false; }
    const uint64_t data_size = data.size() - 1;
    uint64_t total = 0;
    for (const IndexEntry &e : index) total += e.length;
    std::vector<char> plain;
    void *ds = z.createDStream();
    if (ds == nullptr) { err = "ZSTD_createDStream failed"; return false; }
    bool ok = true;
    try {
    plain.reserve(total + 1);
    for (IndexEntry &e : index) {
        if (e.offset + 5 > data_size) { err = "compressed entry beyond the data file"; ok = false; break; }
        uint32_t n;
        memcpy(&n, data.data() + e.offset, 4);
        const char *payload = data.data() + e.offset + 4;
        if ((uint64_t) n + 1 > data_size - e.offset - 4) { err = "compressed entry beyond the data file"; ok = false; break; }
        const uint64_t at = plain.size();
        const uint64_t want = e.length - 1;
        if (payload[n] == 0) {
            plain.resize(at + want + 1);
            z.initDStream(ds);
            ZInBuf in = {payload, n, 0};
            ZOutBuf out = {plain.data() + at, (size_t) want, 0};
            while (in.pos < in.size) {
                const size_t before_in = in.pos, before_out = out.pos;
                const size_t rc = z.decompressStream(ds, &out, &in);
                if (z.isError(rc)) { err = "zstd: corrupt entry"; ok = false; break; }
                if (rc == 0) break;                                              // frame complete
                if (in.pos == before_in && out.pos == before_out) { err = "zstd: entry longer than its index length"; ok = false; break; }
            }
            if (!ok) break;
            if (out.pos != want) { err = "zstd: entry shorter than its index length"; ok = false; break; }
            plain[at + want] = '\0';
        } else {
            if (n != want) { err = "raw entry of a compressed DB disagrees with its index length"; ok = false; break; }
            plain.insert(plain.end(), payload, payload + n);
            plain.push_back('\0');
        }
        e.offset = at;
    }
    } catch (const std::bad_alloc &) { err = "compressed DB: index lengths exceed the memory available"; ok = false; }
    z.freeDStream(ds);
    if (!ok) return false;
    plain.push_back('\0');
    data.swap(plain);
    return true;
}

}  // namespace

struct b200h_db {
    std::vector<char> data;            // all data parts back to back (+ one guard NUL)
    std::vector<IndexEntry> index;     // sorted by key
    int dbtype = -1;
};

struct b200h_dbw {
    std::string path;
    FILE *data = nullptr;
    int dbtype = 0;
    uint64_t offset = 0;
    std::vector<IndexEntry> index;
};

extern "C" {

const char *b200h_db_last_error(void) { return g_db_err.c_str(); }

int b200h_db_open(const char *data_path, b200h_db **out) {
    if (data_path == nullptr || out == nullptr) return db_fail("b200h_db_open: NULL argument");
    *out = nullptr;
    b200h_db *db = new b200h_db();
    const std::string base(data_path);
    // data: <name>, or the split parts <name>.0, <name>.1, ... in order (FileUtil::findDatafiles)
    if (!read_file(base, db->data)) {
        std::vector<char> part;
        int parts = 0;
        while (read_file(base + "." + std::to_string(parts), part)) { db->data.insert(db->data.end(), part.begin(), part.end()); parts++; }
        if (parts == 0) { delete db; return db_fail("cannot open data file " + base); }
    }
    db->data.push_back('\0');
    bool compressed = false;
    std::vector<char> ty;
    if (read_file(base + ".dbtype", ty) && ty.size() >= 4) {
        int32_t v; memcpy(&v, ty.data(), 4);
        // extended flag 8 = GPU-padded residues (Parameters.h:94): that layout is read by b200_db_load_padded from the caller's
        // DBReader, not as text through this reader -- refuse instead of misreading it
        const uint32_t ext = ((uint32_t) v >> 16) & 0x7FFEu;                       // DBReader::getExtendedDbtype
        if ((ext & 8u) != 0) { delete db; return db_fail("GPU-padded DB: not supported by b200h_db_open: " + base); }
        compressed = ((uint32_t) v & (1u << 31)) != 0;                             // DBReader::isCompressed
        db->dbtype = (int32_t) ((uint32_t) v & ~(1u << 31));                       // what callers see is the plain DB
    }
    std::vector<char> idx;
    if (!read_file(base + ".index", idx)) { delete db; return db_fail("cannot open index file " + base + ".index"); }
    idx.push_back('\0');
    const char *p = idx.data();
    const uint64_t data_size = db->data.size() - 1;   // without the sentinel NUL appended above
    while (*p != '\0') {
        // key \t offset \t length \n   (DBReader::readIndex) -- all three fields on this line, nothing borrowed from the next one
        const char *eol = p;
        while (*eol != '\n' && *eol != '\0') eol++;
        uint64_t f[3];
        const char *c = p;
        bool ok = true;
        for (int k = 0; k < 3 && ok; k++) {
            while (c < eol && (*c == ' ' || *c == '\t')) c++;
            if (c >= eol || *c < '0' || *c > '9') { ok = false; break; }
            uint64_t v = 0;
            while (c < eol && *c >= '0' && *c <= '9') { v = v * 10 + (uint64_t) (*c - '0'); c++; }
            f[k] = v;
        }
        if (!ok) { delete db; return db_fail("malformed index line in " + base + ".index"); }
        IndexEntry e;
        e.key = (uint32_t) f[0]; e.offset = f[1]; e.length = f[2];
        // every entry carries at least its NUL (DBWriter::writeEnd); offset + length must not wrap
        // (a compressed DB's index holds the UNCOMPRESSED length: inflate_db bounds-checks those entries against the frame sizes)
        if (e.length == 0 || e.offset > data_size || (!compressed && (e.length > data_size || e.offset > data_size - e.length))) { delete db; return db_fail("index entry beyond the data file (or empty) in " + base); }
        db->index.push_back(e);
        p = (*eol == '\n') ? eol + 1 : eol;
    }
    if (compressed) {
        std::string err;
        if (!inflate_db(db->data, db->index, err)) { delete db; return db_fail(err + ": " + base); }
    }
    std::stable_sort(db->index.begin(), db->index.end(), [](const IndexEntry &a, const IndexEntry &b) { return a.key < b.key; });
    *out = db;
    return B200_OK;
}

void b200h_db_close(b200h_db *db) { delete db; }
uint64_t b200h_db_size(const b200h_db *db) { return db ? db->index.size() : 0; }
int b200h_db_type(const b200h_db *db) { return db ? db->dbtype : -1; }
uint32_t b200h_db_key(const b200h_db *db, uint64_t id) { return db->index[id].key; }
const char *b200h_db_data(const b200h_db *db, uint64_t id) { return db->data.data() + db->index[id].offset; }
uint64_t b200h_db_entry_len(const b200h_db *db, uint64_t id) { return db->index[id].length; }

int64_t b200h_db_id(const b200h_db *db, uint32_t key) {
    auto it = std::lower_bound(db->index.begin(), db->index.end(), key, [](const IndexEntry &e, uint32_t k) { return e.key < k; });
    if (it == db->index.end() || it->key != key) return -1;
    return (int64_t) (it - db->index.begin());
}

int b200h_dbw_open(const char *data_path, int dbtype, b200h_dbw **out) {
    if (data_path == nullptr || out == nullptr) return db_fail("b200h_dbw_open: NULL argument");
    *out = nullptr;
    FILE *f = fopen(data_path, "wb");
    if (f == nullptr) return db_fail(std::string("cannot create ") + data_path);
    b200h_dbw *w = new b200h_dbw();
    w->path = data_path; w->data = f; w->dbtype = dbtype;
    *out = w;
    return B200_OK;
}

int b200h_dbw_write(b200h_dbw *w, uint32_t key, const char *data, uint64_t len) {
    if (w == nullptr || (data == nullptr && len > 0)) return db_fail("b200h_dbw_write: NULL argument");
    if (len > 0 && fwrite(data, 1, len, w->data) != len) return db_fail("short write to " + w->path);
    if (fputc('\0', w->data) == EOF) return db_fail("short write to " + w->path);     // DBWriter::writeEnd
    IndexEntry e; e.key = key; e.offset = w->offset; e.length = len + 1;
    w->index.push_back(e);
    w->offset += len + 1;
    return B200_OK;
}

int b200h_dbw_close(b200h_dbw *w) {
    if (w == nullptr) return db_fail("b200h_dbw_close: NULL argument");
    int rc = B200_OK;
    if (fclose(w->data) != 0) rc = db_fail("cannot close " + w->path);
    std::stable_sort(w->index.begin(), w->index.end(), [](const IndexEntry &a, const IndexEntry &b) { return a.key < b.key; });
    FILE *fi = fopen((w->path + ".index").c_str(), "wb");
    if (fi == nullptr) rc = db_fail("cannot create " + w->path + ".index");
    else {
        for (const IndexEntry &e : w->index) fprintf(fi, "%u\t%llu\t%llu\n", e.key, (unsigned long long) e.offset, (unsigned long long) e.length);
        if (fclose(fi) != 0) rc = db_fail("cannot close " + w->path + ".index");
    }
    FILE *ft = fopen((w->path + ".dbtype").c_str(), "wb");
    if (ft == nullptr) rc = db_fail("cannot create " + w->path + ".dbtype");
    else {
        const int32_t v = w->dbtype;
        if (fwrite(&v, 4, 1, ft) != 1) rc = db_fail("short write to " + w->path + ".dbtype");
        fclose(ft);
    }
    delete w;
    return rc;
}

void b200h_aa2num_table(const char *num2aa, int A, int nucleotide, uint8_t table[256]) {
    // BaseMatrix starts every letter as "unknown"; the alphabet letters get their index; every other byte falls to X, the
    // last code, with the reference's folding of ambiguity letters and lower case
    int code[256];
    for (int i = 0; i < 256; i++) code[i] = -1;
    for (int i = 0; i < A; i++) code[(unsigned char) num2aa[i]] = i;
    const int x = A - 1;
    for (int letter = 0; letter < 256; letter++) {
        const int up = toupper(letter);
        int target;
        if (nucleotide) {
            switch (up) {
                case 'A': case 'T': case 'G': case 'C': target = up; break;
                case 'U': case 'W': target = 'T'; break;
                case 'K': case 'B': case 'D': case 'V': case 'R': case 'S': target = 'G'; break;
                case 'M': case 'Y': case 'H': target = 'C'; break;
                default: target = 'X'; break;
            }
        } else {
            switch (up) {
                case 'A': case 'T': case 'G': case 'C': case 'D': case 'E': case 'F': case 'H': case 'I': case 'K': case 'L': case 'M':
                case 'N': case 'P': case 'Q': case 'R': case 'S': case 'V': case 'W': case 'Y': case 'X': target = up; break;
                case 'J': target = 'L'; break;
                case 'U': case 'O': target = 'X'; break;
                case 'Z': target = 'E'; break;
                case 'B': target = 'D'; break;
                default: target = 'X'; break;
            }
        }
        const int c = code[target];
        table[letter] = (uint8_t) (c >= 0 ? c : x);
    }
}

int b200_align_db(b200_ctx *ctx, const char *query_db, const char *target_db, const char *prefilter_db, const char *alignment_db,
                  const int16_t *sub_matrix, const double *p_back, const char *num2aa, int alphabet, const b200_align_params *params,
                  const b200_evalue_params *evalue, int add_backtrace, uint32_t bucket_queries, uint64_t *n_alignments,
                  uint64_t *n_records) {
    if (ctx == nullptr) return B200_ERR_ARG;
    if (query_db == nullptr || target_db == nullptr || prefilter_db == nullptr || alignment_db == nullptr || sub_matrix == nullptr ||
        p_back == nullptr || num2aa == nullptr || params == nullptr)
        return b200_set_err(ctx, B200_ERR_ARG, "b200_align_db: NULL argument");
    if (bucket_queries == 0) bucket_queries = 4096;
    b200h_db *tdb = nullptr, *qdb = nullptr, *pdb = nullptr;
    b200h_dbw *out = nullptr;
    int rc = b200h_db_open(target_db, &tdb);
    const bool same = std::string(query_db) == std::string(target_db);           // sameQTDB, Alignment.cpp:65
    if (rc == B200_OK) rc = same ? B200_OK : b200h_db_open(query_db, &qdb);
    if (rc == B200_OK && same) qdb = tdb;
    if (rc == B200_OK) rc = b200h_db_open(prefilter_db, &pdb);
    if (rc == B200_OK) rc = b200h_dbw_open(alignment_db, B200_DBTYPE_ALIGNMENT_RES, &out);
    uint64_t total_aln = 0, total_rec = 0;
    if (rc != B200_OK) b200_set_err(ctx, rc, g_db_err.c_str());
    if (rc == B200_OK) {
        uint8_t a2n[256];
        b200h_aa2num_table(num2aa, alphabet, 0, a2n);
        // ---- target DB -> numeric residues in HBM (Sequence::mapSequence per entry; length = index length - 2) -----------
        const uint64_t nt = b200h_db_size(tdb);
        std::vector<uint64_t> toff(nt + 1, 0);
        for (uint64_t i = 0; i < nt; i++) {
            const uint64_t l = b200h_db_entry_len(tdb, i);
            toff[i + 1] = toff[i] + (l >= 2 ? l - 2 : 0);
        }
        std::vector<uint8_t> tres(toff[nt] + 1);
        std::vector<uint32_t> tkeys(nt);
        for (uint64_t i = 0; i < nt; i++) {
            const char *s = b200h_db_data(tdb, i);
            uint8_t *d = tres.data() + toff[i];
            for (uint64_t j = 0; j < toff[i + 1] - toff[i]; j++) d[j] = a2n[(unsigned char) s[j]];
            tkeys[i] = b200h_db_key(tdb, i);
        }
        rc = b200_db_load(ctx, tres.data(), toff.data(), nt, alphabet);
        b200_evalue_params ev_default;
        if (rc == B200_OK && evalue == nullptr) {
            if (b200h_evalue_defaults("blosum62.out", params->gap_open, params->gap_extend, 1, toff[nt], &ev_default) != B200_OK)
                rc = b200_set_err(ctx, B200_ERR_ARG, "b200_align_db: no built-in E-value parameters for these gap costs");
            evalue = &ev_default;
        }
        // ---- prefilter entries in buckets of queries (Alignment.cpp:262-312) -----------------------------------------------
        const uint64_t np = b200h_db_size(pdb);
        b200_align_params par = *params;
        if (same) par.include_identity = 1;
        std::vector<uint8_t> qres;
        std::vector<uint64_t> qoff, hoff;
        std::vector<uint32_t> qkeys, hits, nres;
        std::vector<b200_pref_hit> parsed;
        std::vector<b200_result> results;
        std::vector<char> pool, line, entry;
        for (uint64_t b0 = 0; b0 < np && rc == B200_OK; b0 += bucket_queries) {
            const uint64_t b1 = std::min<uint64_t>(np, b0 + bucket_queries);
            qres.clear(); qoff.assign(1, 0); hoff.assign(1, 0); qkeys.clear(); hits.clear();
            uint64_t bt_cap = 16;
            for (uint64_t i = b0; i < b1 && rc == B200_OK; i++) {
                const uint32_t qkey = b200h_db_key(pdb, i);
                const char *pe = b200h_db_data(pdb, i);
                parsed.resize(b200h_db_entry_len(pdb, i) / 4 + 2);
                const size_t nh = b200h_parse_prefilter_hits(pe, parsed.data(), parsed.size());
                uint64_t qlen = 0;
                if (nh > 0) {      // "only load query data if data != \\0" (Alignment.cpp:323)
                    const int64_t qid = b200h_db_id(qdb, qkey);
                    if (qid < 0) { rc = b200_set_err(ctx, B200_ERR_ARG, "b200_align_db: prefilter query key missing from the query DB"); break; }
                    const uint64_t l = b200h_db_entry_len(qdb, (uint64_t) qid);
                    qlen = l >= 2 ? l - 2 : 0;
                    const char *s = b200h_db_data(qdb, (uint64_t) qid);
                    for (uint64_t j = 0; j < qlen; j++) qres.push_back(a2n[(unsigned char) s[j]]);
                }
                qoff.push_back(qres.size());
                qkeys.push_back(qkey);
                for (size_t k = 0; k < nh; k++) {
                    const int64_t tid = b200h_db_id(tdb, parsed[k].seq_id);
                    if (tid < 0) { rc = b200_set_err(ctx, B200_ERR_ARG, "b200_align_db: prefilter target key missing from the target DB"); break; }
                    hits.push_back((uint32_t) tid);
                    bt_cap += qlen + (toff[tid + 1] - toff[tid]);
                }
                hoff.push_back(hits.size());
            }
            if (rc != B200_OK) break;
            const uint32_t nq = (uint32_t) (b1 - b0);
            results.resize(hits.size() + 1); nres.assign(nq + 1, 0); pool.resize(bt_cap);
            uint64_t n_aln = 0;
            rc = b200_align_batch(ctx, sub_matrix, p_back, alphabet, qres.data(), qoff.data(), qkeys.data(), nq, hoff.data(), hits.data(),
                                  tkeys.data(), &par, evalue, results.data(), nres.data(), pool.data(), bt_cap, &n_aln);
            if (rc != B200_OK) break;
            total_aln += n_aln;
            for (uint32_t i = 0; i < nq && rc == B200_OK; i++) {
                entry.clear();
                for (uint32_t k = 0; k < nres[i]; k++) {
                    const b200_result &r = results[hoff[i] + k];
                    line.resize(512 + 2 * (size_t) r.bt_len);
                    const size_t len = b200h_result_to_buffer(line.data(), &r, pool.data() + r.bt_off, add_backtrace, 1);
                    entry.insert(entry.end(), line.data(), line.data() + len);
                }
                total_rec += nres[i];
                if (b200h_dbw_write(out, qkeys[i], entry.data(), entry.size()) != B200_OK) rc = b200_set_err(ctx, B200_ERR_ARG, g_db_err.c_str());
            }
        }
    }
    if (out != nullptr && b200h_dbw_close(out) != B200_OK && rc == B200_OK) rc = b200_set_err(ctx, B200_ERR_ARG, g_db_err.c_str());
    if (pdb != nullptr) b200h_db_close(pdb);
    if (qdb != nullptr && qdb != tdb) b200h_db_close(qdb);
    if (tdb != nullptr) b200h_db_close(tdb);
    if (n_alignments != nullptr) *n_alignments = total_aln;
    if (n_records != nullptr) *n_records = total_rec;
    return rc;
}

namespace {
// numeric residues of every entry of a sequence DB (Sequence::mapSequence: length = index length - 2)
// lower_to_x: soft-masked (lowercase) residues become X, as runFilterOnCpu does with the targets before scoring (ungappedprefilter.cpp:401-404)
void load_sequences(const b200h_db *db, const uint8_t a2n[256], std::vector<uint8_t> &res, std::vector<uint64_t> &off, std::vector<uint32_t> &keys,
                    bool lower_to_x = false) {
    const uint64_t n = b200h_db_size(db);
    off.assign(n + 1, 0);
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t l = b200h_db_entry_len(db, i);
        off[i + 1] = off[i] + (l >= 2 ? l - 2 : 0);
    }
    res.resize(off[n] + 1);
    keys.resize(n);
    for (uint64_t i = 0; i < n; i++) {
        const char *s = b200h_db_data(db, i);
        uint8_t *d = res.data() + off[i];
        const uint8_t x = a2n[(unsigned char) 'X'];
        for (uint64_t j = 0; j < off[i + 1] - off[i]; j++) d[j] = (lower_to_x && (unsigned char) s[j] >= 97) ? x : a2n[(unsigned char) s[j]];
        keys[i] = b200h_db_key(db, i);
    }
}
}  // namespace

int b200_prefilter_db(b200_ctx *ctx, const char *query_db, const char *target_db, const char *prefilter_db, const int16_t *sub_matrix,
                      const double *p_back, const char *num2aa, int alphabet, int comp_bias, float comp_bias_scale, int min_diag_score,
                      uint32_t max_res_list_len, uint32_t bucket_queries, uint64_t *n_hits) {
    if (ctx == nullptr) return B200_ERR_ARG;
    if (query_db == nullptr || target_db == nullptr || prefilter_db == nullptr || sub_matrix == nullptr || p_back == nullptr ||
        num2aa == nullptr || max_res_list_len == 0)
        return b200_set_err(ctx, B200_ERR_ARG, "b200_prefilter_db: bad argument");
    if (bucket_queries == 0) bucket_queries = 64;
    b200h_db *tdb = nullptr, *qdb = nullptr;
    b200h_dbw *out = nullptr;
    int rc = b200h_db_open(target_db, &tdb);
    if (rc == B200_OK) rc = b200h_db_open(query_db, &qdb);
    if (rc == B200_OK) rc = b200h_dbw_open(prefilter_db, B200_DBTYPE_PREFILTER_RES, &out);
    if (rc != B200_OK) b200_set_err(ctx, rc, g_db_err.c_str());
    uint64_t total = 0;
    if (rc == B200_OK) {
        uint8_t a2n[256];
        b200h_aa2num_table(num2aa, alphabet, 0, a2n);
        std::vector<uint8_t> tres, qres;
        std::vector<uint64_t> toff, qoff;
        std::vector<uint32_t> tkeys, qkeys;
        load_sequences(tdb, a2n, tres, toff, tkeys, /*lower_to_x=*/true);
        load_sequences(qdb, a2n, qres, qoff, qkeys);
        rc = b200_db_load(ctx, tres.data(), toff.data(), tkeys.size(), alphabet);
        const uint64_t nq = qkeys.size();
        const int A = alphabet;
        std::vector<std::vector<int8_t>> profiles;
        std::vector<b200_query> queries;
        std::vector<b200_hit> hits;
        std::vector<uint32_t> nh;
        std::vector<float> fbias;
        std::vector<int8_t> cb;
        std::vector<char> entry;
        char line[64];
        for (uint64_t b0 = 0; b0 < nq && rc == B200_OK; b0 += bucket_queries) {
            const uint64_t b1 = std::min<uint64_t>(nq, b0 + bucket_queries);
            const int nb = (int) (b1 - b0);
            profiles.assign(nb, std::vector<int8_t>());
            queries.assign(nb, b200_query());
            std::vector<int> live;              // zero-length queries get an empty entry without touching the device
            for (int i = 0; i < nb; i++) {
                const uint64_t qi = b0 + i;
                const int L = (int) (qoff[qi + 1] - qoff[qi]);
                if (L == 0) continue;
                const uint8_t *seq = qres.data() + qoff[qi];
                // the profile ungappedprefilter.cpp:186-203 hands to the scorer: matrix column + rounded composition bias
                cb.assign((size_t) L, 0);
                if (comp_bias) {
                    fbias.resize((size_t) L);
                    b200h_comp_bias(sub_matrix, p_back, A, seq, L, comp_bias_scale, fbias.data());
                    b200h_round_bias_ssw(fbias.data(), L, cb.data());
                }
                profiles[i].resize((size_t) A * L);
                if (b200h_build_profile(sub_matrix, A, seq, L, cb.data(), 1, profiles[i].data()) != 0) { rc = b200_set_err(ctx, B200_ERR_RANGE, "b200_prefilter_db: profile value outside int8"); break; }
                b200_query q; q.profile = profiles[i].data(); q.qlen = L; q.bias = b200h_ssw_bias(sub_matrix, A, cb.data(), L, comp_bias ? 1 : 0);
                queries[live.size()] = q;
                live.push_back(i);
            }
            if (rc != B200_OK) break;
            hits.resize((size_t) live.size() * max_res_list_len + 1);
            nh.assign(live.size() + 1, 0);
            if (!live.empty())
                rc = b200_ungapped_scan(ctx, queries.data(), (int) live.size(), min_diag_score, max_res_list_len, hits.data(), nh.data(), nullptr);
            if (rc != B200_OK) break;
            size_t lk = 0;
            for (int i = 0; i < nb; i++) {
                entry.clear();
                if (lk < live.size() && live[lk] == i) {
                    // device order is (score desc, DB-local id asc); ids are key ranks, so this is compareHitsByScoreAndId on keys
                    for (uint32_t k = 0; k < nh[lk]; k++) {
                        const b200_hit &h = hits[lk * max_res_list_len + k];
                        b200_pref_hit ph; ph.seq_id = tkeys[h.id]; ph.pref_score = h.score; ph.diagonal = 0; ph.pad_ = 0;
                        const size_t len = b200h_prefilter_hit_to_buffer(line, &ph);
                        entry.insert(entry.end(), line, line + len);
                    }
                    total += nh[lk];
                    lk++;
                }
                if (b200h_dbw_write(out, qkeys[b0 + i], entry.data(), entry.size()) != B200_OK) { rc = b200_set_err(ctx, B200_ERR_ARG, g_db_err.c_str()); break; }
            }
        }
    }
    if (out != nullptr && b200h_dbw_close(out) != B200_OK && rc == B200_OK) rc = b200_set_err(ctx, B200_ERR_ARG, g_db_err.c_str());
    if (qdb != nullptr) b200h_db_close(qdb);
    if (tdb != nullptr) b200h_db_close(tdb);
    if (n_hits != nullptr) *n_hits = total;
    return rc;
}

}  // extern "C"
