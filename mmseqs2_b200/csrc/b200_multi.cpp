// mmseqs2_b200/csrc/b200_multi.cpp -- several GPUs behind one handle (include/b200_multi.h): host-side sharding over per-device
// b200_ctx objects, one std::thread per device inside every call.  No CUDA here: everything goes through the C ABI of b200_align.h.
#include "b200_multi.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct b200_multi {
    std::vector<b200_ctx *> ctx;
    std::vector<int> device;
    std::string err;
    bool shard_targets = false;
    uint64_t n_seq = 0;
    std::vector<uint64_t> tgt_begin;   // target-sharded: device d owns global ids [tgt_begin[d], tgt_begin[d+1])
};

namespace {

int fail(b200_multi *m, int code, const std::string &msg) { m->err = msg; return code; }

// contiguous ranges balanced by weight (DBReader::decomposeDomainByAminoAcid's idea): bounds[parts + 1]
std::vector<uint64_t> balanced_ranges(const std::vector<uint64_t> &weight, int parts) {
    const uint64_t n = weight.size();
    uint64_t total = 0;
    for (uint64_t w : weight) total += w;
    std::vector<uint64_t> bounds(parts + 1, n);
    bounds[0] = 0;
    uint64_t acc = 0, i = 0;
    for (int p = 1; p < parts; p++) {
        const double target = (double) total * p / parts;
        while (i < n && (double) (acc + weight[i]) <= target) { acc += weight[i]; i++; }
        if (i < n && (double) acc < target && target - (double) acc > (double) (acc + weight[i]) - target) { acc += weight[i]; i++; }
        bounds[p] = i;
    }
    return bounds;
}

// per-device top lists of one query (local ids) -> the global list: ids shifted by the slice's first id, ordered by the deterministic
// comparator (score desc, global id asc) -- hit_t::compareHitsByScoreAndId -- and cut to max_hits; returns the number written
uint32_t merge_top_hits(const b200_hit *const *lists, const uint32_t *counts, const uint64_t *first_id, int n_lists, uint32_t max_hits, b200_hit *out) {
    std::vector<b200_hit> all;
    for (int d = 0; d < n_lists; d++)
        for (uint32_t k = 0; k < counts[d]; k++) {
            b200_hit h = lists[d][k];
            h.id += (uint32_t) first_id[d];
            all.push_back(h);
        }
    std::sort(all.begin(), all.end(), [](const b200_hit &a, const b200_hit &b) { return a.score != b.score ? a.score > b.score : a.id < b.id; });
    const uint32_t n = (uint32_t) std::min<size_t>(all.size(), max_hits);
    if (n > 0) memcpy(out, all.data(), sizeof(b200_hit) * n);
    return n;
}

template <typename F>
int on_every_device(b200_multi *m, F f) {
    const int n = (int) m->ctx.size();
    std::vector<int> rc(n, B200_OK);
    if (n == 1) { rc[0] = f(0); }
    else {
        std::vector<std::thread> th;
        for (int d = 0; d < n; d++) th.emplace_back([&, d]() { rc[d] = f(d); });
        for (auto &t : th) t.join();
    }
    for (int d = 0; d < n; d++)
        if (rc[d] != B200_OK) return fail(m, rc[d], std::string("device ") + std::to_string(m->device[d]) + ": " + b200_last_error(m->ctx[d]));
    return B200_OK;
}

}  // namespace

extern "C" {

int b200_multi_create(const int *devices, int n_devices, b200_multi **out) {
    if (out == nullptr) return B200_ERR_ARG;
    *out = nullptr;
    std::vector<int> ids;
    if (devices == nullptr) {
        const int n = b200_device_count();
        for (int i = 0; i < n; i++) ids.push_back(i);
    } else {
        ids.assign(devices, devices + std::max(0, n_devices));
    }
    if (ids.empty()) return B200_ERR_ARG;
    b200_multi *m = new b200_multi();
    for (int id : ids) {
        b200_ctx *c = nullptr;
        const int rc = b200_create(id, &c);
        if (rc != B200_OK) { b200_multi_destroy(m); return rc; }
        m->ctx.push_back(c);
        m->device.push_back(id);
    }
    *out = m;
    return B200_OK;
}

void b200_multi_destroy(b200_multi *m) {
    if (m == nullptr) return;
    for (b200_ctx *c : m->ctx) b200_destroy(c);
    delete m;
}

int b200_multi_size(const b200_multi *m) { return m ? (int) m->ctx.size() : 0; }
b200_ctx *b200_multi_ctx(b200_multi *m, int i) { return (m && i >= 0 && i < (int) m->ctx.size()) ? m->ctx[i] : nullptr; }
const char *b200_multi_last_error(const b200_multi *m) { return m ? m->err.c_str() : "null handle"; }

int b200_multi_db_load(b200_multi *m, const uint8_t *residues, const uint64_t *offsets, uint64_t n_seq, int alphabet, int shard_targets) {
    if (m == nullptr) return B200_ERR_ARG;
    if (residues == nullptr || offsets == nullptr || n_seq == 0) return fail(m, B200_ERR_ARG, "b200_multi_db_load: bad arguments");
    const int nd = (int) m->ctx.size();
    m->shard_targets = shard_targets != 0 && nd > 1;
    m->n_seq = n_seq;
    m->tgt_begin.assign(nd + 1, 0);
    if (!m->shard_targets) {
        m->tgt_begin[nd] = n_seq;
        return on_every_device(m, [&](int d) { return b200_db_load(m->ctx[d], residues, offsets, n_seq, alphabet); });
    }
    std::vector<uint64_t> w(n_seq);
    for (uint64_t i = 0; i < n_seq; i++) w[i] = offsets[i + 1] - offsets[i] + 1;
    m->tgt_begin = balanced_ranges(w, nd);
    for (int d = 0; d < nd; d++)
        if (m->tgt_begin[d + 1] == m->tgt_begin[d]) return fail(m, B200_ERR_ARG, "b200_multi_db_load: fewer sequences than devices");
    return on_every_device(m, [&](int d) {
        const uint64_t a = m->tgt_begin[d], b = m->tgt_begin[d + 1];
        return b200_db_load(m->ctx[d], residues, offsets + a, b - a, alphabet);     // offsets are absolute into residues: a sub-range works as is
    });
}

int b200_multi_db_load_padded(b200_multi *m, const uint8_t *data, const size_t *offsets, const int32_t *lengths, uint64_t n_seq,
                              int alphabet, int shard_targets) {
    if (m == nullptr) return B200_ERR_ARG;
    if (data == nullptr || offsets == nullptr || lengths == nullptr || n_seq == 0) return fail(m, B200_ERR_ARG, "b200_multi_db_load_padded: bad arguments");
    const int nd = (int) m->ctx.size();
    m->shard_targets = shard_targets != 0 && nd > 1;
    m->n_seq = n_seq;
    m->tgt_begin.assign(nd + 1, 0);
    if (!m->shard_targets) {
        m->tgt_begin[nd] = n_seq;
        return on_every_device(m, [&](int d) { return b200_db_load_padded(m->ctx[d], data, offsets, lengths, n_seq, alphabet); });
    }
    // the padded DB is sorted by length: contiguous slices would give one device all the long sequences; slices still have to be
    // contiguous id ranges for the (score, id) merge to be a plain offset, so balance them by residues
    std::vector<uint64_t> w(n_seq);
    for (uint64_t i = 0; i < n_seq; i++) w[i] = (uint64_t) lengths[i] + 1;
    m->tgt_begin = balanced_ranges(w, nd);
    for (int d = 0; d < nd; d++)
        if (m->tgt_begin[d + 1] == m->tgt_begin[d]) return fail(m, B200_ERR_ARG, "b200_multi_db_load_padded: fewer sequences than devices");
    return on_every_device(m, [&](int d) {
        const uint64_t a = m->tgt_begin[d], b = m->tgt_begin[d + 1];
        return b200_db_load_padded(m->ctx[d], data, offsets + a, lengths + a, b - a, alphabet);
    });
}

int b200_multi_db_load_padded_unmasked(b200_multi *m, const uint8_t *data, const size_t *offsets, const int32_t *lengths, uint64_t n_seq, int alphabet) {
    if (m == nullptr) return B200_ERR_ARG;
    if (data == nullptr || offsets == nullptr || lengths == nullptr || n_seq == 0) return fail(m, B200_ERR_ARG, "b200_multi_db_load_padded_unmasked: bad arguments");
    const int nd = (int) m->ctx.size();
    m->shard_targets = false;
    m->n_seq = n_seq;
    m->tgt_begin.assign(nd + 1, 0);
    m->tgt_begin[nd] = n_seq;
    return on_every_device(m, [&](int d) { return b200_db_load_padded_unmasked(m->ctx[d], data, offsets, lengths, n_seq, alphabet); });
}

int b200_multi_ungapped_scan(b200_multi *m, const b200_query *queries, int nq, int min_score_excl, uint32_t max_hits, b200_hit *hits,
                             uint32_t *n_hits) {
    if (m == nullptr) return B200_ERR_ARG;
    if (queries == nullptr || nq <= 0 || max_hits == 0 || hits == nullptr || n_hits == nullptr) return fail(m, B200_ERR_ARG, "b200_multi_ungapped_scan: bad arguments");
    if (m->n_seq == 0) return fail(m, B200_ERR_NODB, "no target DB loaded");
    const int nd = (int) m->ctx.size();
    if (!m->shard_targets) {
        std::vector<uint64_t> w(nq);
        for (int i = 0; i < nq; i++) w[i] = (uint64_t) std::max(1, queries[i].qlen);
        const std::vector<uint64_t> qb = balanced_ranges(w, nd);
        return on_every_device(m, [&](int d) {
            const uint64_t a = qb[d], b = qb[d + 1];
            if (a == b) return (int) B200_OK;
            return b200_ungapped_scan(m->ctx[d], queries + a, (int) (b - a), min_score_excl, max_hits, hits + a * max_hits, n_hits + a, nullptr);
        });
    }
    // target-sharded: local top lists, then the k-way merge with the deterministic comparator (score desc, global id asc)
    std::vector<std::vector<b200_hit>> lh(nd, std::vector<b200_hit>((size_t) nq * max_hits));
    std::vector<std::vector<uint32_t>> ln(nd, std::vector<uint32_t>(nq, 0));
    const int rc = on_every_device(m, [&](int d) {
        return b200_ungapped_scan(m->ctx[d], queries, nq, min_score_excl, max_hits, lh[d].data(), ln[d].data(), nullptr);
    });
    if (rc != B200_OK) return rc;
    std::vector<const b200_hit *> lists(nd);
    std::vector<uint32_t> counts(nd);
    for (int q = 0; q < nq; q++) {
        for (int d = 0; d < nd; d++) { lists[d] = lh[d].data() + (size_t) q * max_hits; counts[d] = ln[d][q]; }
        n_hits[q] = merge_top_hits(lists.data(), counts.data(), m->tgt_begin.data(), nd, max_hits, hits + (size_t) q * max_hits);
    }
    return B200_OK;
}

int b200_multi_align_batch(b200_multi *m, const int16_t *sub_matrix, const double *p_back, int alphabet, const uint8_t *query_residues,
                           const uint64_t *query_offsets, const uint32_t *query_keys, uint32_t n_queries, const uint64_t *hit_offsets,
                           const uint32_t *hit_targets, const uint32_t *target_keys, const b200_align_params *params,
                           const b200_evalue_params *evalue, b200_result *results, uint32_t *n_results, char *bt_pool, uint64_t bt_cap,
                           uint64_t *n_alignments) {
    if (m == nullptr) return B200_ERR_ARG;
    if (m->shard_targets) return fail(m, B200_ERR_ARG, "b200_multi_align_batch needs a replicated (query-sharded) DB");
    if (query_offsets == nullptr || hit_offsets == nullptr || n_results == nullptr) return fail(m, B200_ERR_ARG, "b200_multi_align_batch: bad arguments");
    const int nd = (int) m->ctx.size();
    if (nd == 1 || n_queries <= 1)
        return on_every_device(m, [&](int d) {
            if (d != 0) return (int) B200_OK;
            return b200_align_batch(m->ctx[0], sub_matrix, p_back, alphabet, query_residues, query_offsets, query_keys, n_queries, hit_offsets,
                                    hit_targets, target_keys, params, evalue, results, n_results, bt_pool, bt_cap, n_alignments);
        });
    // queries balanced by the gapped work they carry: sum over their hits of the query length (target lengths are not known here)
    std::vector<uint64_t> w(n_queries);
    for (uint32_t i = 0; i < n_queries; i++) w[i] = 1 + (hit_offsets[i + 1] - hit_offsets[i]) * (query_offsets[i + 1] - query_offsets[i]);
    const std::vector<uint64_t> qb = balanced_ranges(w, nd);
    std::vector<std::vector<char>> pools(nd);
    std::vector<uint64_t> naln(nd, 0);
    std::vector<std::vector<uint64_t>> qoff(nd), hoff(nd);
    const bool want_bt = params != nullptr && params->sw_mode == 2 && bt_pool != nullptr;
    int rc = on_every_device(m, [&](int d) {
        const uint64_t a = qb[d], b = qb[d + 1];
        if (a == b) return (int) B200_OK;
        // rebased copies of the offset arrays of this query range
        qoff[d].resize(b - a + 1); hoff[d].resize(b - a + 1);
        for (uint64_t i = a; i <= b; i++) { qoff[d][i - a] = query_offsets[i] - query_offsets[a]; hoff[d][i - a] = hit_offsets[i] - hit_offsets[a]; }
        uint64_t cap = want_bt ? std::max<uint64_t>(1 << 20, bt_cap / nd) : 16;
        int r;
        while (true) {
            pools[d].resize(cap);
            r = b200_align_batch(m->ctx[d], sub_matrix, p_back, alphabet, query_residues + query_offsets[a], qoff[d].data(),
                                 query_keys ? query_keys + a : nullptr, (uint32_t) (b - a), hoff[d].data(), hit_targets + hit_offsets[a],
                                 target_keys, params, evalue, results + hit_offsets[a], n_results + a, pools[d].data(), cap, &naln[d]);
            if (r == B200_ERR_RANGE && want_bt && cap < bt_cap) { cap = std::min(bt_cap, cap * 4); continue; }
            break;
        }
        return r;
    });
    if (rc != B200_OK) return rc;
    // backtrace strings: per-device pools appended to the caller's pool, offsets rebased
    uint64_t used = 0, total_aln = 0;
    for (int d = 0; d < nd; d++) {
        total_aln += naln[d];
        const uint64_t a = qb[d], b = qb[d + 1];
        if (!want_bt || a == b) continue;
        uint64_t dev_used = 0;
        for (uint64_t q = a; q < b; q++)
            for (uint32_t k = 0; k < n_results[q]; k++) {
                b200_result &r = results[hit_offsets[q] + k];
                if (r.bt_len == 0) continue;
                dev_used = std::max(dev_used, r.bt_off + r.bt_len);
                r.bt_off += used;
            }
        if (used + dev_used > bt_cap) return fail(m, B200_ERR_RANGE, "b200_multi_align_batch: bt_cap too small");
        if (dev_used > 0) memcpy(bt_pool + used, pools[d].data(), dev_used);
        used += dev_used;
    }
    if (n_alignments) *n_alignments = total_aln;
    return B200_OK;
}

// ---- the sharding arithmetic on its own (no device): what the multi-device calls cut and merge with; used by the CPU tests ------------
void b200h_balanced_ranges(const uint64_t *weights, uint64_t n, int parts, uint64_t *bounds) {
    const std::vector<uint64_t> b = balanced_ranges(std::vector<uint64_t>(weights, weights + n), parts);
    for (int p = 0; p <= parts; p++) bounds[p] = b[p];
}

uint32_t b200h_merge_top_hits(const b200_hit *const *lists, const uint32_t *counts, const uint64_t *first_id, int n_lists, uint32_t max_hits,
                              b200_hit *out) {
    return merge_top_hits(lists, counts, first_id, n_lists, max_hits, out);
}

}  // extern "C"
