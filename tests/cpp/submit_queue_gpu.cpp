// tests/cpp/submit_queue_gpu.cpp -- b200::DiagSubmitQueue on a real context: the UngappedAlignment objects of several host threads
// (QueryMatcher's shape: one per OpenMP thread, QueryMatcher.cpp:73) share one queue; every thread must read the counts and raw
// scores a direct b200_diag_score call gives for the same hits, and the device must have been called fewer times than align() was.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "b200_mmseqs.hpp"

int main() {
    const int A = 21, nSeq = 3000, nThreads = 8, iters = 12;
    unsigned s = 99u;
    auto rnd = [&s]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    std::vector<int16_t> mat((size_t) A * A);
    for (int i = 0; i < A; i++)
        for (int j = 0; j < A; j++) mat[i * A + j] = (int16_t) (i == j ? 4 + (i % 3) : ((i + j) % 5 == 0 ? 1 : -2));
    std::vector<size_t> off(nSeq + 1, 0);
    for (int i = 0; i < nSeq; i++) off[i + 1] = off[i] + 30 + rnd() % 400;
    std::vector<unsigned char> data(off[nSeq]);
    for (size_t i = 0; i < data.size(); i++) data[i] = (unsigned char) (rnd() % 20);
    b200::Device dev(0);
    if (!dev.ok()) { fprintf(stderr, "no device: %s\n", dev.error()); return 2; }
    if (dev.loadLookup(data.data(), off.data(), nSeq, A) != B200_OK) { fprintf(stderr, "loadLookup: %s\n", dev.error()); return 2; }
    b200::DiagSubmitQueue queue(&dev, 64);
    std::atomic<long> wrong(0), errors(0), calls(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nThreads; t++)
        th.emplace_back([&, t]() {
            unsigned r = 4321u + 17u * (unsigned) t;
            auto rn = [&r]() { r = r * 1664525u + 1013904223u; return r >> 8; };
            b200::UngappedAlignment queued(&dev, mat.data(), A, &queue), direct(&dev, mat.data(), A);
            for (int it = 0; it < iters; it++) {
                const int L = 50 + (int) (rn() % 350);
                std::vector<unsigned char> q(L);
                for (int i = 0; i < L; i++) q[i] = (unsigned char) (rn() % 20);
                std::vector<float> bias(L);
                for (int i = 0; i < L; i++) bias[i] = ((int) (rn() % 200) - 100) / 25.0f;
                if (queued.createProfile(q.data(), L, bias.data()) != B200_OK || direct.createProfile(q.data(), L, bias.data()) != B200_OK) { errors++; continue; }
                const size_t n = it == 3 ? 0 : 200 + rn() % 4000;
                std::vector<b200::CounterResult> a(n), b;
                for (size_t i = 0; i < n; i++) {
                    a[i].id = rn() % nSeq;
                    a[i].diagonal = (unsigned short) (short) ((int) (rn() % 600) - 300);
                    a[i].count = (rn() % 9 == 0) ? 7 : 0;                   // non-zero counts are left alone (UngappedAlignment.cpp:327-329)
                }
                b = a;
                std::vector<int32_t> rawA(n, -7), rawB(n, -7);
                const bool rescore = it % 2 == 1;
                const int rcA = rescore ? queued.rescore(a.data(), n, rawA.data()) : queued.align(a.data(), n);
                const int rcB = rescore ? direct.rescore(b.data(), n, rawB.data()) : direct.align(b.data(), n);
                calls++;
                if (rcA != B200_OK || rcB != B200_OK) { errors++; continue; }
                for (size_t i = 0; i < n; i++)
                    if (a[i].count != b[i].count || (rescore && rawA[i] != rawB[i])) wrong++;
            }
        });
    for (size_t i = 0; i < th.size(); i++) th[i].join();
    printf("align calls %ld queue requests %ld device rounds %ld wrong %ld errors %ld\n", calls.load(), (long) queue.requests(), (long) queue.rounds(), wrong.load(), errors.load());
    return (wrong == 0 && errors == 0 && (long) queue.requests() == calls.load() && (long) queue.rounds() <= calls.load()) ? 0 : 1;
}
