// scan_sort.cuh -- device-wide prefix scan and stable LSD radix sort with DEVICE-SIDE element
// counts, so the region/link/graph stages can be chained on one stream without host round trips.
// Grids are fixed (multiples of the 148 SMs); each block owns a contiguous chunk of the input.
#pragma once
#include <utility>
#include "common.cuh"

namespace bdk {

constexpr int SS_THREADS = 256;
constexpr int SS_WARPS = SS_THREADS / 32;
constexpr int SS_GRID = kNumSMs * 2;          // 296 blocks

__device__ __forceinline__ uint32_t ss_chunk(uint32_t n) {   // per-block chunk, multiple of the tile size
    uint32_t c = div_up<uint32_t>(n, SS_GRID);
    return div_up<uint32_t>(c, SS_THREADS) * SS_THREADS;
}

// inclusive scan of one value per thread across the block; returns inclusive value, *total = block sum
__device__ __forceinline__ uint32_t ss_block_scan(uint32_t v, uint32_t* s_warp /*[SS_WARPS + 1]*/, uint32_t* total) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SS_WARPS ? s_warp[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < SS_WARPS; d <<= 1) { uint32_t t = __shfl_up_sync(FULL, wi, d); if (lane >= d) wi += t; }
        if (lane < SS_WARPS) s_warp[lane] = wi - w;
        if (lane == SS_WARPS - 1) s_warp[SS_WARPS] = wi;
    }
    __syncthreads();
    inc += s_warp[warp];
    *total = s_warp[SS_WARPS];
    __syncthreads();
    return inc;
}

// ---- scan: out(i, inclusive_prefix, value) for i < *n, value = f(i) ------------------------------
template <class F>
__global__ void __launch_bounds__(SS_THREADS) scan_reduce_kernel(F f, const uint32_t* __restrict__ n_ptr, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t s_warp[SS_WARPS + 1];
    const uint32_t n = *n_ptr, chunk = ss_chunk(n);
    const uint32_t lo = min(n, blockIdx.x * chunk), hi = min(n, lo + chunk);
    uint32_t s = 0;
    for (uint32_t i = lo + threadIdx.x; i < hi; i += SS_THREADS) s += f(i, n);
    uint32_t total;
    ss_block_scan(s, s_warp, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// exclusive scan of the SS_GRID block sums in place; total -> *total_out (+ add)
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t* __restrict__ block_sums, uint32_t* __restrict__ total_out, uint32_t add) {
    __shared__ uint32_t s[1024];
    const int t = threadIdx.x;
    uint32_t v = t < SS_GRID ? block_sums[t] : 0;
    s[t] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        uint32_t x = t >= d ? s[t - d] : 0;
        __syncthreads();
        s[t] += x;
        __syncthreads();
    }
    if (t < SS_GRID) block_sums[t] = s[t] - v;
    if (t == 1023 && total_out) *total_out = s[t] + add;
}

template <class F, class O>
__global__ void __launch_bounds__(SS_THREADS) scan_apply_kernel(F f, O out, const uint32_t* __restrict__ n_ptr, const uint32_t* __restrict__ block_offs) {
    __shared__ uint32_t s_warp[SS_WARPS + 1];
    const uint32_t n = *n_ptr, chunk = ss_chunk(n);
    const uint32_t lo = min(n, blockIdx.x * chunk), hi = min(n, lo + chunk);
    uint32_t carry = block_offs[blockIdx.x];
    for (uint32_t base = lo; base < hi; base += SS_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < hi ? f(i, n) : 0;
        uint32_t total;
        const uint32_t inc = ss_block_scan(v, s_warp, &total) + carry;
        if (i < hi) out(i, inc, v, n);
        carry += total;
    }
}

struct ScanScratch { uint32_t* block_sums; };   // [SS_GRID]

template <class F, class O>
inline void device_scan(cudaStream_t st, F f, O out, const uint32_t* n_ptr, uint32_t* total_out, uint32_t total_add, ScanScratch sc) {
    scan_reduce_kernel<<<SS_GRID, SS_THREADS, 0, st>>>(f, n_ptr, sc.block_sums);
    scan_sums_kernel<<<1, 1024, 0, st>>>(sc.block_sums, total_out, total_add);
    scan_apply_kernel<<<SS_GRID, SS_THREADS, 0, st>>>(f, out, n_ptr, sc.block_sums);
}

// inclusive scan of one value per thread across a block of up to 1024 threads; s_warp: [33]
__device__ __forceinline__ uint32_t ss_block_scan_any(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nw ? s_warp[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULL, wi, d); if (lane >= d) wi += t; }
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    inc += s_warp[warp];
    *total = s_warp[32];
    __syncthreads();
    return inc;
}

// ---- stable LSD radix sort of u64 items, 8 bits per pass, the digit of pass p given by a functor D(item, p) -------
template <class D>
__global__ void __launch_bounds__(SS_THREADS) sort_hist_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ n_ptr,
                                                               D digit, int pass, uint32_t* __restrict__ hist /*[256][SS_GRID]*/) {
    __shared__ uint32_t s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n = *n_ptr, chunk = ss_chunk(n);
    const uint32_t lo = min(n, blockIdx.x * chunk), hi = min(n, lo + chunk);
    for (uint32_t i = lo + threadIdx.x; i < hi; i += SS_THREADS) atomicAdd(&s_h[digit(keys[i], pass)], 1u);
    __syncthreads();
    hist[threadIdx.x * SS_GRID + blockIdx.x] = s_h[threadIdx.x];
}

// exclusive scan over the digit-major [256][SS_GRID] table, in place (one block of 32 warps).
// Each warp owns 8 digit rows: coalesced row sums, a 256-entry scan of the row totals, then a
// warp-scan of every row seeded with its base.
__global__ void __launch_bounds__(1024) sort_scan_kernel(uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_row[256];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int CH = (SS_GRID + 31) / 32;
    for (int d = warp * 8; d < warp * 8 + 8; ++d) {
        uint32_t s = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) { const int i = c * 32 + lane; if (i < SS_GRID) s += hist[d * SS_GRID + i]; }
        s = __reduce_add_sync(FULL, s);
        if (lane == 0) s_row[d] = s;
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the 256 row totals: 8 per lane
        uint32_t v[8], t = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = s_row[lane * 8 + k]; t += v[k]; }
        uint32_t inc = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t x = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += x; }
        uint32_t run = inc - t;
#pragma unroll
        for (int k = 0; k < 8; ++k) { s_row[lane * 8 + k] = run; run += v[k]; }
    }
    __syncthreads();
    for (int d = warp * 8; d < warp * 8 + 8; ++d) {
        uint32_t v[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) { const int i = c * 32 + lane; v[c] = i < SS_GRID ? hist[d * SS_GRID + i] : 0; }
        uint32_t base = s_row[d];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            uint32_t inc = v[c];
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) { uint32_t x = __shfl_up_sync(FULL, inc, k); if (lane >= k) inc += x; }
            const int i = c * 32 + lane;
            if (i < SS_GRID) hist[d * SS_GRID + i] = base + inc - v[c];
            base += __shfl_sync(FULL, inc, 31);
        }
    }
}

template <class D>
__global__ void __launch_bounds__(SS_THREADS) sort_scatter_kernel(const unsigned long long* __restrict__ keys_in, unsigned long long* __restrict__ keys_out,
        const uint32_t* __restrict__ n_ptr, D digit, int pass, const uint32_t* __restrict__ offs /*[256][SS_GRID] exclusive*/) {
    __shared__ uint32_t s_base[256];
    __shared__ uint32_t s_wcnt[SS_WARPS][256];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s_base[threadIdx.x] = offs[threadIdx.x * SS_GRID + blockIdx.x];
    const uint32_t n = *n_ptr, chunk = ss_chunk(n);
    const uint32_t lo = min(n, blockIdx.x * chunk), hi = min(n, lo + chunk);
    for (uint32_t base = lo; base < hi; base += SS_THREADS) {
        for (int w = 0; w < SS_WARPS; ++w) s_wcnt[w][threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = base + threadIdx.x;
        const bool v = i < hi;
        unsigned long long k = 0; uint32_t d = 0;
        if (v) { k = keys_in[i]; d = digit(k, pass); }
        const unsigned vm = __ballot_sync(FULL, v);
        uint32_t rank = 0;
        if (v) {
            const unsigned peers = __match_any_sync(vm, d);
            rank = __popc(peers & lanemask_lt());
            if (lane == __ffs(peers) - 1) s_wcnt[warp][d] = __popc(peers);
        }
        __syncthreads();
        {   // per digit: exclusive prefix over the warps; threadIdx.x is the digit
            uint32_t run = 0;
            for (int w = 0; w < SS_WARPS; ++w) { uint32_t c = s_wcnt[w][threadIdx.x]; s_wcnt[w][threadIdx.x] = run; run += c; }
            __syncthreads();
            if (v) keys_out[s_base[d] + s_wcnt[warp][d] + rank] = k;
            __syncthreads();
            s_base[threadIdx.x] += run;
        }
        __syncthreads();
    }
}

struct SortScratch { uint32_t* hist; unsigned long long* keys_tmp; };  // hist [256][SS_GRID]

// `npasses` passes, ping-pong between *keys and sc.keys_tmp; returns the buffer that holds the result.
template <class D>
inline unsigned long long* device_radix_sort(cudaStream_t st, unsigned long long* keys, const uint32_t* n_ptr, D digit, int npasses, SortScratch& sc) {
    unsigned long long* kin = keys; unsigned long long* kout = sc.keys_tmp;
    for (int pass = 0; pass < npasses; ++pass) {
        sort_hist_kernel<<<SS_GRID, SS_THREADS, 0, st>>>(kin, n_ptr, digit, pass, sc.hist);
        sort_scan_kernel<<<1, 1024, 0, st>>>(sc.hist);
        sort_scatter_kernel<<<SS_GRID, SS_THREADS, 0, st>>>(kin, kout, n_ptr, digit, pass, sc.hist);
        std::swap(kin, kout);
    }
    return kin;
}

// The same sort by ONE CTA of 1024 threads, all passes inside one launch (no per-pass launch cost: a few microseconds per pass
// for tens of thousands of items), followed by an epilogue over the sorted array. Each warp owns a contiguous segment of
// the input and keeps per-digit cursors in shared memory; ranks inside a group of 32 items come from __match_any_sync.
// The result is in `a` when npasses is even, else in `b`.
constexpr int SORT1_THREADS = 1024;
template <class D, class Epi>
__global__ void __launch_bounds__(SORT1_THREADS) sort_single_cta_kernel(unsigned long long* __restrict__ a, unsigned long long* __restrict__ b,
        const uint32_t* __restrict__ n_ptr, D digit, int npasses, Epi epi) {
    __shared__ uint32_t s_cnt[256][33];          // [digit][warp], padded
    __shared__ uint32_t s_scratch[33];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n = *n_ptr;
    const uint32_t seg = div_up<uint32_t>(div_up<uint32_t>(n, 32u), 32u) * 32u;
    const uint32_t lo = min(n, warp * seg), hi = min(n, lo + seg);
    unsigned long long* in = a; unsigned long long* out = b;
    for (int pass = 0; pass < npasses; ++pass) {
        for (int i = threadIdx.x; i < 256 * 33; i += SORT1_THREADS) (&s_cnt[0][0])[i] = 0;
        __syncthreads();
        for (uint32_t base = lo; base < hi; base += 32) {
            const uint32_t i = base + lane;
            const bool v = i < hi;
            const uint32_t d = v ? digit(in[i], pass) : 0;
            const unsigned vm = __ballot_sync(FULL, v);
            if (v) {
                const unsigned peers = __match_any_sync(vm, d);
                if (lane == __ffs(peers) - 1) s_cnt[d][warp] += __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        {   // exclusive scan over (digit, warp) in digit-major order: 8 consecutive entries per thread
            const int e0 = threadIdx.x * 8;
            uint32_t v[8], t = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { const int e = e0 + k; v[k] = s_cnt[e >> 5][e & 31]; t += v[k]; }
            uint32_t total;
            uint32_t run = ss_block_scan_any(t, s_scratch, &total) - t;
#pragma unroll
            for (int k = 0; k < 8; ++k) { const int e = e0 + k; s_cnt[e >> 5][e & 31] = run; run += v[k]; }
        }
        __syncthreads();
        for (uint32_t base = lo; base < hi; base += 32) {
            const uint32_t i = base + lane;
            const bool v = i < hi;
            unsigned long long k = 0; uint32_t d = 0;
            if (v) { k = in[i]; d = digit(k, pass); }
            const unsigned vm = __ballot_sync(FULL, v);
            if (v) {
                const unsigned peers = __match_any_sync(vm, d);
                const uint32_t at = s_cnt[d][warp] + __popc(peers & lanemask_lt());
                out[at] = k;
                __syncwarp(peers);
                if (lane == __ffs(peers) - 1) s_cnt[d][warp] += __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        unsigned long long* t = in; in = out; out = t;
    }
    __threadfence_block();
    epi(in, n, s_scratch);
}

}  // namespace bdk
