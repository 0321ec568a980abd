// nway_merge.hpp -- the order in which the reference's BamMerger delivers the records of n position-sorted bams
// (src/lib/io/BamMerger.cpp:40-126: a std::priority_queue over the streams' heads ordered by (tid, pos, strand); pop the top,
// emit it, push the stream's next record), computed from the streams' packed keys alone. Used by bdk_push_bams for three or more
// bams decoded on the device (bdk_bam.inl): the keys come back from the GPU, the order goes there, the ten columns are gathered
// through it on the device.
//
// For three or more streams the order among equal keys depends on the heap's LAYOUT, i.e. on its whole history (the two-stream
// case has a closed form: csrc/bam_merge.cuh), so the queue itself has to run. Two things make that fast and keep it exact:
//   * SmallHeap is libstdc++'s binary heap (std::push_heap / std::pop_heap as priority_queue::push / pop call them: bits/stl_heap.h
//     __push_heap, __adjust_heap) on a fixed array with the key carried in the element -- the comparator returns what the
//     reference's returns, so the heap makes the same moves (tests/hostsim/nway_merge_check.cpp compares with a
//     std::priority_queue over (stream, index) on tie-ridden keys);
//   * the merge is cut at (tid, pos) boundaries: everything in front of a boundary leaves the queue before anything behind it,
//     so the streams' cursors at a boundary are known (a binary search each) -- what is NOT known is the layout the heap has
//     there. But a heap of m heads has only a few valid layouts (2 for three distinct keys, 3 for four, 8 for five), so every
//     part is merged once per valid layout, all parts and layouts in parallel, and afterwards the parts are chained: part j's
//     true initial layout is the final layout of part j - 1's chosen run. Exact by construction, a few times the work, all cores.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

namespace bdh {

struct MergeHead { uint64_t key; uint32_t i; uint32_t bam; };

// libstdc++'s heap with comp(a, b) = a.key > b.key (a min-heap on the key), on at most CAP elements
struct SmallHeap {
    static constexpr int CAP = 64;
    MergeHead e[CAP];
    int len = 0;
    static bool comp(const MergeHead& a, const MergeHead& b) { return a.key > b.key; }
    void push_heap_at(int hole, int top, const MergeHead& value) {            // __push_heap
        int parent = (hole - 1) / 2;
        while (hole > top && comp(e[parent], value)) {
            e[hole] = e[parent];
            hole = parent;
            parent = (hole - 1) / 2;
        }
        e[hole] = value;
    }
    void push(const MergeHead& x) { push_heap_at(len, 0, x); ++len; }         // c.push_back(x); std::push_heap
    void pop() {                                                               // std::pop_heap; c.pop_back()
        if (len > 1) {
            const int n = len - 1;                                             // __pop_heap: value = last, hole at the root, heap of n
            const MergeHead value = e[n];
            int hole = 0, second = 0;
            while (second < (n - 1) / 2) {                                     // __adjust_heap
                second = 2 * (second + 1);
                if (comp(e[second], e[second - 1])) --second;
                e[hole] = e[second];
                hole = second;
            }
            if ((n & 1) == 0 && second == (n - 2) / 2) {
                second = 2 * (second + 1);
                e[hole] = e[second - 1];
                hole = second - 1;
            }
            push_heap_at(hole, 0, value);
        }
        --len;
    }
};

// One run of the queue: from the heap as it stands, until it is empty or its top reaches `stop` ((tid, pos) part of the key,
// i.e. key >> 1; ~0 = no stop). counts[b] = records of stream b. Returns the number of records emitted.
inline uint64_t nway_run(SmallHeap& h, const uint64_t* const* keys, const uint64_t* counts, uint64_t stop_pos, int bam_shift, uint32_t* order) {
    uint64_t o = 0;
    while (h.len) {
        const MergeHead t = h.e[0];
        if ((t.key >> 1) >= stop_pos) break;
        h.pop();
        order[o++] = t.i | (t.bam << bam_shift);
        if ((uint64_t)t.i + 1 < counts[t.bam]) h.push(MergeHead{keys[t.bam][t.i + 1], t.i + 1, t.bam});
    }
    return o;
}

// order[o] = index | bam << bam_shift for o < sum(counts). keys[b][i] = packed (tid, pos, strand) of record i of bam b. One thread.
inline void nway_merge_order(const uint64_t* const* keys, const uint64_t* counts, int n, int bam_shift, uint32_t* order) {
    SmallHeap h;
    for (int b = 0; b < n; ++b)
        if (counts[b]) h.push(MergeHead{keys[b][0], 0u, (uint32_t)b});
    nway_run(h, keys, counts, ~0ull, bam_shift, order);
}

// The same order with the work spread over `threads` threads (see the header). Needs every stream ordered by (tid, pos)
// (key >> 1 non-decreasing; the strand bit may go either way) and n <= 5; the caller falls back to nway_merge_order otherwise.
// `part_records`: about how many records a part should hold (tests use tiny parts).
inline bool nway_merge_order_parallel(const uint64_t* const* keys, const uint64_t* counts, int n, int bam_shift, uint32_t* order,
                                      int threads, uint64_t part_records = 1u << 18) {
    constexpr int MAXV = 8;                                  // valid layouts of a part's first heap we are willing to run
    if (n < 2 || n > 5 || threads < 2) return false;
    uint64_t total = 0;
    int big = 0;
    for (int b = 0; b < n; ++b) { total += counts[b]; if (counts[b] > counts[big]) big = b; }
    if (total < 4 * part_records) return false;
    struct Variant { int m; uint32_t lay[5]; uint32_t fin[5]; int fin_m; std::vector<uint32_t> out; };
    struct Part { uint64_t pos; uint64_t cur[5]; uint64_t first_out, n_out; std::vector<Variant> v; };
    auto cursor = [&](int b, uint64_t pos) {
        return (uint64_t)(std::lower_bound(keys[b], keys[b] + counts[b], pos, [](uint64_t k, uint64_t p) { return (k >> 1) < p; }) - keys[b]);
    };
    // valid heap layouts of the heads at a boundary: arrangements of the non-exhausted streams with no parent greater than its child
    auto layouts = [&](const uint64_t* cur, std::vector<Variant>& out) {
        uint32_t ids[5]; int m = 0;
        for (int b = 0; b < n; ++b) if (cur[b] < counts[b]) ids[m++] = (uint32_t)b;      // ascending: the first permutation
        out.clear();
        do {
            bool ok = true;
            for (int c = 1; c < m && ok; ++c) ok = !(keys[ids[(c - 1) / 2]][cur[ids[(c - 1) / 2]]] > keys[ids[c]][cur[ids[c]]]);
            if (ok) {
                if ((int)out.size() == MAXV) return false;
                Variant v; v.m = m; v.fin_m = 0;
                for (int c = 0; c < m; ++c) v.lay[c] = ids[c];
                out.push_back(std::move(v));
            }
        } while (std::next_permutation(ids, ids + m));
        return true;
    };
    std::vector<Part> parts;
    {   // part 0 starts from the queue as BamMerger builds it: the streams pushed in order
        Part p0; p0.pos = 0; p0.first_out = 0;
        for (int b = 0; b < n; ++b) p0.cur[b] = 0;
        SmallHeap h;
        for (int b = 0; b < n; ++b) if (counts[b]) h.push(MergeHead{keys[b][0], 0u, (uint32_t)b});
        Variant v; v.m = h.len; v.fin_m = 0;
        for (int c = 0; c < h.len; ++c) v.lay[c] = h.e[c].bam;
        p0.v.push_back(std::move(v));
        parts.push_back(std::move(p0));
    }
    const uint64_t want_parts = std::max<uint64_t>(2, total / part_records);
    for (uint64_t q = 1; q < want_parts; ++q) {
        uint64_t i = counts[big] * q / want_parts;
        for (int tries = 0; tries < 64 && i < counts[big]; ++tries, ++i) {
            if (i == 0 || (keys[big][i] >> 1) == (keys[big][i - 1] >> 1)) continue;         // a boundary starts a (tid, pos)
            const uint64_t pos = keys[big][i] >> 1;
            if (pos <= parts.back().pos) continue;
            Part p; p.pos = pos; p.first_out = 0;
            for (int b = 0; b < n; ++b) { p.cur[b] = cursor(b, pos); p.first_out += p.cur[b]; }
            if (p.first_out <= parts.back().first_out) continue;
            if (!layouts(p.cur, p.v)) continue;               // too many ties between the heads here: try the next position
            parts.push_back(std::move(p));
            break;
        }
    }
    const size_t np = parts.size();
    if (np < 2) return false;
    for (size_t j = 0; j < np; ++j) parts[j].n_out = (j + 1 < np ? parts[j + 1].first_out : total) - parts[j].first_out;
    // every (part, layout) run, largest first
    std::vector<std::pair<uint32_t, uint32_t>> tasks;
    for (size_t j = 0; j < np; ++j) for (size_t v = 0; v < parts[j].v.size(); ++v) tasks.push_back({(uint32_t)j, (uint32_t)v});
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        for (;;) {
            const size_t t = next.fetch_add(1);
            if (t >= tasks.size()) return;
            Part& P = parts[tasks[t].first];
            Variant& V = P.v[tasks[t].second];
            SmallHeap h;
            h.len = V.m;
            for (int c = 0; c < V.m; ++c) { const uint32_t b = V.lay[c]; h.e[c] = MergeHead{keys[b][P.cur[b]], (uint32_t)P.cur[b], b}; }
            V.out.resize(P.n_out);
            const uint64_t stop = tasks[t].first + 1 < np ? parts[tasks[t].first + 1].pos : ~0ull;
            const uint64_t got = nway_run(h, keys, counts, stop, bam_shift, V.out.data());
            V.fin_m = got == P.n_out ? h.len : -1;            // (-1: cannot happen for sorted streams; the chain below then gives up)
            for (int c = 0; c < h.len && c < 5; ++c) V.fin[c] = h.e[c].bam;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < threads; ++t) th.emplace_back(worker);
        worker();
        for (auto& x : th) x.join();
    }
    // chain: the layout a part really starts from is the one its predecessor's chosen run ended in
    std::vector<int> chosen(np, -1);
    chosen[0] = 0;
    for (size_t j = 0; j + 1 < np; ++j) {
        const Variant& V = parts[j].v[chosen[j]];
        if (V.fin_m < 0) return false;
        for (size_t w = 0; w < parts[j + 1].v.size(); ++w) {
            const Variant& W = parts[j + 1].v[w];
            if (W.m == V.fin_m && std::equal(W.lay, W.lay + W.m, V.fin)) { chosen[j + 1] = (int)w; break; }
        }
        if (chosen[j + 1] < 0) return false;
    }
    if (parts[np - 1].v[chosen[np - 1]].fin_m < 0) return false;
    std::atomic<size_t> nextp(0);
    auto copier = [&]() {
        for (;;) {
            const size_t j = nextp.fetch_add(1);
            if (j >= np) return;
            const std::vector<uint32_t>& src = parts[j].v[chosen[j]].out;
            std::copy(src.begin(), src.end(), order + parts[j].first_out);
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < threads; ++t) th.emplace_back(copier);
        copier();
        for (auto& x : th) x.join();
    }
    return true;
}

}  // namespace bdh
