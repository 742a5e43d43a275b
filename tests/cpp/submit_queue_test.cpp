// tests/cpp/submit_queue_test.cpp -- b200::DiagSubmitQueueT (include/b200_mmseqs.hpp) with a stand-in backend: many threads submit
// the hit lists of their own query; every thread must get exactly the results a direct call would have given it, and concurrent
// submissions must have been combined into fewer backend calls.  No GPU, no library: the queue is host logic.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "b200_mmseqs.hpp"

static std::atomic<long> g_calls(0), g_queries(0), g_maxBatch(0);

static uint8_t expectCount(uint32_t id, uint16_t dg, int qlen, int8_t tag) { return (uint8_t) ((id * 31u + dg + (unsigned) qlen + (unsigned) (uint8_t) tag) & 0xffu); }
static int32_t expectRaw(uint32_t id, uint16_t dg, int qlen) { return (int32_t) (id ^ ((uint32_t) dg << 8)) + qlen; }

struct MockBackend {
    int failEvery;
    int operator()(const b200_query *q, int nq, const uint64_t *off, const uint32_t *ids, const uint16_t *dg, uint8_t *counts, int32_t *raw) const {
        const long call = ++g_calls;
        g_queries += nq;
        long seen = g_maxBatch.load();
        while (nq > seen && !g_maxBatch.compare_exchange_weak(seen, nq)) {}
        std::this_thread::sleep_for(std::chrono::microseconds(300));        // a device round trip: lets other threads queue up
        if (failEvery > 0 && call % failEvery == 0) return B200_ERR_CUDA;
        for (int i = 0; i < nq; i++)
            for (uint64_t h = off[i]; h < off[i + 1]; h++) {
                if (counts[h] == 0) counts[h] = expectCount(ids[h], dg[h], q[i].qlen, q[i].profile[0]);   // non-zero counts are skipped
                if (raw != NULL) raw[h] = expectRaw(ids[h], dg[h], q[i].qlen);
            }
        return B200_OK;
    }
};

int main(int argc, char **argv) {
    const int nThreads = argc > 1 ? atoi(argv[1]) : 16, perThread = argc > 2 ? atoi(argv[2]) : 60, failEvery = argc > 3 ? atoi(argv[3]) : 0;
    MockBackend be; be.failEvery = failEvery;
    b200::DiagSubmitQueueT<MockBackend> queue(be, 8);                           // at most 8 queries per backend call
    std::atomic<long> wrong(0), failed(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nThreads; t++)
        th.emplace_back([&, t]() {
            unsigned s = 1234u + 77u * (unsigned) t;
            auto rnd = [&s]() { s = s * 1664525u + 1013904223u; return s >> 8; };
            for (int it = 0; it < perThread; it++) {
                const size_t n = it % 7 == 0 ? 0 : rnd() % 500;
                std::vector<uint32_t> ids(n); std::vector<uint16_t> dg(n); std::vector<uint8_t> counts(n), before(n); std::vector<int32_t> raw(n, -1);
                for (size_t i = 0; i < n; i++) { ids[i] = rnd(); dg[i] = (uint16_t) rnd(); before[i] = counts[i] = (rnd() % 5 == 0) ? (uint8_t) (1 + rnd() % 200) : 0; }
                int8_t tag[4] = {(int8_t) (t * 3 + it), 0, 0, 0};
                b200_query q; q.profile = tag; q.qlen = 10 + (int) (rnd() % 900); q.bias = 0;
                const bool wantRaw = it % 3 != 0;
                const int rc = queue.submit(q, ids.data(), dg.data(), n, counts.data(), wantRaw ? raw.data() : NULL);
                if (rc != B200_OK) { failed++; continue; }
                for (size_t i = 0; i < n; i++) {
                    const uint8_t e = before[i] != 0 ? before[i] : expectCount(ids[i], dg[i], q.qlen, tag[0]);
                    if (counts[i] != e) wrong++;
                    if (wantRaw && raw[i] != expectRaw(ids[i], dg[i], q.qlen)) wrong++;
                }
            }
        });
    for (size_t i = 0; i < th.size(); i++) th[i].join();
    const long total = (long) nThreads * perThread;
    printf("requests %ld rounds %ld backend_calls %ld queries_sent %ld max_batch %ld wrong %ld failed %ld\n", (long) queue.requests(), (long) queue.rounds(),
           g_calls.load(), g_queries.load(), g_maxBatch.load(), wrong.load(), failed.load());
    if (wrong != 0 || (long) queue.requests() != total || g_queries.load() > total) return 1;      // no request reaches the backend twice (rounds without hits skip it)
    if (failEvery == 0 && failed != 0) return 1;
    if (g_maxBatch.load() > 8) return 1;                                                              // the per-call cap holds
    if (nThreads >= 8 && g_calls.load() * 2 > total) return 1;                                        // combining happened
    return 0;
}
