// bam_merge.cuh -- two decoded bams merged on the device in the reference's order (BamMerger: a std::priority_queue over the
// streams' heads ordered by (tid, pos, strand), src/lib/io/BamMerger.cpp:40-126), for bdk_push_bams (bdk_bam.inl).
//
// For two streams the heap's behaviour has a closed form (the host merge in csrc/host/bam_io.cpp uses the same one and is tested
// against the priority queue): after a pop from stream A the heap holds B's head alone; A's next record is pushed below it and
// sifts up only if it is STRICTLY smaller, so on a tie the stream that did not emit last goes first (at the very start: bam 0,
// which was pushed first). That rule needs no history across a (tid, pos) that occurs in one bam only: everything before it leaves
// both streams first and the heads differ after it. So: find such positions in the larger bam (one binary search per run start),
// keep about one per 64 records as cuts, let one thread merge each part sequentially with the rule (keys need not be monotone
// inside a (tid, pos) group: the strand bit is not part of a bam's order, the procedure is what defines the result), then gather
// the ten columns through the merge order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bam_decode.cuh"

namespace bammerge {

using bamdev::Columns;

constexpr uint32_t CUT_EVERY = 64;
constexpr uint32_t NO_CUT = 0xffffffffu;

// (tid, pos, strand) lexicographic in one word, as the host merge packs it
__global__ void __launch_bounds__(256) keys_kernel(const int32_t* __restrict__ tid, const int32_t* __restrict__ pos, const uint16_t* __restrict__ flag,
                                                   uint32_t n, unsigned long long* __restrict__ key) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        key[i] = (unsigned long long)(uint32_t)tid[i] << 33 | (unsigned long long)((uint32_t)pos[i] ^ 0x80000000u) << 1 | (unsigned long long)((flag[i] & 0x10) != 0);
}

__device__ __forceinline__ uint32_t lower_bound_pos(const unsigned long long* __restrict__ k, uint32_t n, unsigned long long want /* key >> 1 */) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if ((k[mid] >> 1) < want) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One thread per block of CUT_EVERY records of the cutting bam X (the larger one): the first record of the block that starts a
// (tid, pos) run which does not occur in the other bam Y. cut_x[b] = its index (NO_CUT: none), cut_y[b] = records of Y before it.
__global__ void __launch_bounds__(128) find_cuts_kernel(const unsigned long long* __restrict__ kx, uint32_t nx, const unsigned long long* __restrict__ ky, uint32_t ny,
                                                        uint32_t nblocks, uint32_t* __restrict__ cut_x, uint32_t* __restrict__ cut_y) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    uint32_t found = NO_CUT, fy = 0;
    const uint32_t lo = b * CUT_EVERY, hi = min(nx, lo + CUT_EVERY);
    for (uint32_t i = max(lo, 1u); i < hi; ++i) {                    // (index 0 is the start of the first part anyway)
        const unsigned long long p = kx[i] >> 1;
        if ((kx[i - 1] >> 1) == p) continue;
        const uint32_t j = lower_bound_pos(ky, ny, p);
        if (j < ny && (ky[j] >> 1) == p) continue;
        found = i; fy = j;
        break;
    }
    cut_x[b] = found; cut_y[b] = fy;
}

struct CutFlag {
    const uint32_t* cut_x;
    __device__ uint32_t operator()(uint32_t b, uint32_t) const { return cut_x[b] != NO_CUT ? 1u : 0u; }
};
struct CutOut {       // part q + 1 starts at cut q (part 0 starts at (0, 0))
    const uint32_t* cut_x; const uint32_t* cut_y; uint32_t* part_x; uint32_t* part_y;
    __device__ void operator()(uint32_t b, uint32_t inc, uint32_t v, uint32_t) const {
        if (v) { part_x[inc] = cut_x[b]; part_y[inc] = cut_y[b]; }
    }
};

// One thread per part: the sequential two-way merge with the reference's tie rule. k0 / k1 are ALWAYS bam 0 / bam 1 (the rule
// is not symmetric); part_a / part_b are the part starts in bam 0 / bam 1. order[o] = index | bam << 31.
__global__ void __launch_bounds__(128) merge_parts_kernel(const unsigned long long* __restrict__ k0, uint32_t n0, const unsigned long long* __restrict__ k1, uint32_t n1,
                                                          const uint32_t* __restrict__ part_a, const uint32_t* __restrict__ part_b, const uint32_t* __restrict__ nparts_ptr,
                                                          uint32_t* __restrict__ order, uint32_t* __restrict__ longest) {
    const uint32_t nparts = *nparts_ptr + 1;
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nparts) return;
    uint32_t i = q ? part_a[q] : 0, j = q ? part_b[q] : 0;
    const uint32_t e0 = q + 1 < nparts ? part_a[q + 1] : n0, e1 = q + 1 < nparts ? part_b[q + 1] : n1;
    atomicMax(longest, (e0 - i) + (e1 - j));
    uint32_t o = i + j;
    int last = 1;
    while (i < e0 && j < e1) {
        const unsigned long long a = k0[i], b = k1[j];
        const bool take0 = last == 0 ? a < b : a <= b;
        if (take0) { order[o++] = i++; last = 0; } else { order[o++] = (j++) | 0x80000000u; last = 1; }
    }
    while (i < e0) order[o++] = i++;
    while (j < e1) order[o++] = (j++) | 0x80000000u;
}

__global__ void __launch_bounds__(256) gather_kernel(const uint32_t* __restrict__ order, uint32_t n, Columns a, Columns b, Columns out) {
    for (uint32_t o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
        const uint32_t s = order[o], i = s & 0x7fffffffu;
        const Columns& c = (s >> 31) ? b : a;
        out.pos[o] = c.pos[i]; out.mpos[o] = c.mpos[i]; out.tid[o] = c.tid[i]; out.mtid[o] = c.mtid[i]; out.isize[o] = c.isize[i];
        out.qlen[o] = c.qlen[i]; out.flag[o] = c.flag[i]; out.rgid[o] = c.rgid[i]; out.mapq[o] = c.mapq[i]; out.qid[o] = c.qid[i];
    }
}

// Three or more bams: the order comes from the host (csrc/host/nway_merge.hpp: the priority queue itself; its tie order depends on
// the heap's history, so there is nothing to cut) as index | bam << BAM_SHIFT; the columns are gathered through it here.
constexpr int MAX_BAMS = 16, BAM_SHIFT = 28;
struct ColumnsN { Columns c[MAX_BAMS]; };
__global__ void __launch_bounds__(256) gather_n_kernel(const uint32_t* __restrict__ order, uint32_t n, ColumnsN src, Columns out) {
    for (uint32_t o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
        const uint32_t s = order[o], i = s & ((1u << BAM_SHIFT) - 1u);
        const Columns& c = src.c[s >> BAM_SHIFT];
        out.pos[o] = c.pos[i]; out.mpos[o] = c.mpos[i]; out.tid[o] = c.tid[i]; out.mtid[o] = c.mtid[i]; out.isize[o] = c.isize[i];
        out.qlen[o] = c.qlen[i]; out.flag[o] = c.flag[i]; out.rgid[o] = c.rgid[i]; out.mapq[o] = c.mapq[i]; out.qid[o] = c.qid[i];
    }
}

}  // namespace bammerge
