// scan_sort.cuh -- device-wide prefix scan with DEVICE-SIDE element counts, so the region/link/graph
// stages can be chained on one stream without host round trips.
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

}  // namespace bdk
