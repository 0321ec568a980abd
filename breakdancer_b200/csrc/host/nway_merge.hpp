// nway_merge.hpp -- the order in which the reference's BamMerger delivers the records of n position-sorted bams
// (src/lib/io/BamMerger.cpp:40-126: a std::priority_queue over the streams' heads ordered by (tid, pos, strand); pop the top,
// emit it, push the stream's next record), computed from the streams' packed keys alone. For three or more streams the tie order
// depends on the heap's layout, i.e. on its whole history, so there is no cut where a parallel merge could start afresh (the
// two-stream case has a closed form: csrc/bam_merge.cuh): this IS the priority queue, with the key carried in the heap element
// instead of being fetched through (bam, index) at every comparison -- the comparator returns the same answers, so libstdc++'s
// heap makes the same moves as the reference's (tests/hostsim/nway_merge_check.cpp compares the two forms on tie-ridden keys).
// Used by bdk_push_bams for three or more bams decoded on the device (bdk_bam.inl): the keys come back from the GPU, the order goes
// there, the ten columns are gathered through it on the device.
#pragma once
#include <stdint.h>
#include <queue>
#include <vector>

namespace bdh {

struct MergeHead { uint64_t key; uint32_t i; uint32_t bam; };

// order[o] = index | bam << bam_shift for o < sum(counts). keys[b][i] = packed (tid, pos, strand) of record i of bam b.
inline void nway_merge_order(const uint64_t* const* keys, const uint64_t* counts, int n, int bam_shift, uint32_t* order) {
    auto greater = [](const MergeHead& x, const MergeHead& y) { return x.key > y.key; };
    std::priority_queue<MergeHead, std::vector<MergeHead>, decltype(greater)> pq(greater);
    for (int b = 0; b < n; ++b)
        if (counts[b]) pq.push(MergeHead{keys[b][0], 0u, (uint32_t)b});
    uint64_t o = 0;
    while (!pq.empty()) {
        const MergeHead h = pq.top();
        pq.pop();
        order[o++] = h.i | (h.bam << bam_shift);
        if ((uint64_t)h.i + 1 < counts[h.bam]) pq.push(MergeHead{keys[h.bam][h.i + 1], h.i + 1, h.bam});
    }
}

}  // namespace bdh
