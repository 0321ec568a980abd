// The n-way merge order of csrc/host/nway_merge.hpp (libstdc++'s heap on a fixed array, key carried in the heap element) against
// the formulation of the host decoder and of the reference (BamMerger.cpp: a std::priority_queue whose elements are streams, the
// comparator looks at the streams' current records): the same order on keys full of ties between streams, for 1..9 streams,
// sorted and unsorted, with empty streams. Then the parallel form (parts cut at position boundaries, one run per valid layout of
// the heap at the cut, chained afterwards) with tiny parts, so that hundreds of cuts and every kind of tie at a cut are met.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <queue>
#include <random>
#include <vector>
#include "../../breakdancer_b200/csrc/host/nway_merge.hpp"

struct Head { int bam; uint64_t i; };

int main(int argc, char** argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 400;
    std::mt19937_64 rng(99);
    long checked = 0, parallel_runs = 0;
    for (int r = 0; r < rounds; ++r) {
        const int n = 1 + (int)(rng() % 9);
        std::vector<std::vector<uint64_t>> keys(n);
        const uint64_t span = 1 + rng() % (r % 3 == 0 ? 8 : 300);          // few distinct keys: ties everywhere
        uint64_t total = 0;
        for (int b = 0; b < n; ++b) {
            const uint64_t cnt = rng() % 7 == 0 ? 0 : rng() % 400;
            uint64_t k = rng() % span;
            for (uint64_t i = 0; i < cnt; ++i) {
                if (r % 5 == 4) k = rng() % span;                          // unsorted streams: the queue still defines an order
                else k += rng() % 3 == 0 ? rng() % 4 : 0;
                keys[b].push_back(k);
            }
            total += cnt;
        }
        // the reference's form
        auto greater = [&](const Head& x, const Head& y) { return keys[x.bam][x.i] > keys[y.bam][y.i]; };
        std::priority_queue<Head, std::vector<Head>, decltype(greater)> pq(greater);
        for (int b = 0; b < n; ++b) if (!keys[b].empty()) pq.push(Head{b, 0});
        std::vector<uint32_t> want;
        while (!pq.empty()) {
            Head h = pq.top(); pq.pop();
            want.push_back((uint32_t)h.i | ((uint32_t)h.bam << 28));
            if (h.i + 1 < keys[h.bam].size()) pq.push(Head{h.bam, h.i + 1});
        }
        std::vector<const uint64_t*> kp(n);
        std::vector<uint64_t> counts(n);
        for (int b = 0; b < n; ++b) { kp[b] = keys[b].data(); counts[b] = keys[b].size(); }
        std::vector<uint32_t> got(total + 1, 0xdeadbeefu);
        bdh::nway_merge_order(kp.data(), counts.data(), n, 28, got.data());
        if (got[total] != 0xdeadbeefu) { fprintf(stderr, "wrote past the end\n"); return 2; }
        for (uint64_t o = 0; o < total; ++o)
            if (got[o] != want[o]) { fprintf(stderr, "MISMATCH round %d at %llu of %llu (n = %d)\n", r, (unsigned long long)o, (unsigned long long)total, n); return 1; }
        checked += (long)total;
        // the parallel form: sorted streams (by key >> 1; the low bit -- the strand -- goes either way), 2..5 of them
        if (r % 5 != 4 && n >= 2 && n <= 5) {
            for (int b = 0; b < n; ++b) {
                std::vector<uint64_t>& k = keys[b];
                std::sort(k.begin(), k.end(), [](uint64_t x, uint64_t y) { return (x >> 1) < (y >> 1); });
            }
            // reference order again on the re-sorted streams
            std::priority_queue<Head, std::vector<Head>, decltype(greater)> pq2(greater);
            for (int b = 0; b < n; ++b) if (!keys[b].empty()) pq2.push(Head{b, 0});
            want.clear();
            while (!pq2.empty()) {
                Head h = pq2.top(); pq2.pop();
                want.push_back((uint32_t)h.i | ((uint32_t)h.bam << 28));
                if (h.i + 1 < keys[h.bam].size()) pq2.push(Head{h.bam, h.i + 1});
            }
            for (int b = 0; b < n; ++b) kp[b] = keys[b].data();
            std::vector<uint32_t> par(total + 1, 0xdeadbeefu);
            const uint64_t part = 8 + rng() % 60;
            if (bdh::nway_merge_order_parallel(kp.data(), counts.data(), n, 28, par.data(), 1 + (int)(rng() % 4) + 1, part)) {
                ++parallel_runs;
                if (par[total] != 0xdeadbeefu) { fprintf(stderr, "parallel form wrote past the end\n"); return 2; }
                for (uint64_t o = 0; o < total; ++o)
                    if (par[o] != want[o]) { fprintf(stderr, "PARALLEL MISMATCH round %d at %llu of %llu (n = %d, part %llu)\n", r, (unsigned long long)o, (unsigned long long)total, n, (unsigned long long)part); return 1; }
            }
        }
    }
    if (parallel_runs < rounds / 8) { fprintf(stderr, "the parallel form ran only %ld times\n", parallel_runs); return 3; }
    printf("ok rounds=%d records=%ld parallel_runs=%ld\n", rounds, checked, parallel_runs);
    return 0;
}
