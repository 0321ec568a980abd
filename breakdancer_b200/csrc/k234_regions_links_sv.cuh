// k234_regions_links_sv.cuh -- the stages that run on the compacted anomalous-read stream
// (about 1-3 % of the records):
//   K2  break flags -> candidate regions (segmented reductions) -> accepted regions
//       (BreakDancer::push_read:209-241, process_breakpoint:244-264, ReadRegionData::add_region)
//   K3  mate join by read-name key (hash table), region-region mate links aggregated into weighted
//       edges (ReadRegionData.cpp:109-113, Graph.hpp:41-46), followed edges sorted per flush window
//   K4  table of region deletion windows (fixed point), calls per window, pairs per call, Poisson
//       score (build_connection, process_sv, SvBuilder, is_region_final, ComputeProbScore)
// All element counts stay on the device (d_cnt[]); kernels are grid-stride over fixed grids.
#pragma once
#include "common.cuh"
#include "scan_sort.cuh"
#include "bdk_finalize.h"

namespace bdk {

enum { CNT_A = 0, CNT_NCAND, CNT_NREG, CNT_NSE, CNT_NROW, CNT_ERR, CNT_NEMIT, CNT_K4_CHANGED, CNT_K4_NBIGWIN, CNT_NDUP, CNT_N };
constexpr int GS_THREADS = 256;
constexpr int GS_GRID = kNumSMs * 4;

// ---- pass-1 statistics -> BamSummary numbers, densities, window (one CTA) ---------------------
__global__ void __launch_bounds__(256) finalize_kernel(FinalizeIn in, uint64_t n_records, const uint32_t* __restrict__ d_cnt,
                                                       bdk_summary_t* __restrict__ S, float* __restrict__ density) {
    __shared__ unsigned long long s_ref[BDK_MAX_BAMS];
    for (int b = threadIdx.x; b < BDK_MAX_BAMS; b += blockDim.x) s_ref[b] = 0;
    __syncthreads();
    const int total = in.nbam * in.ntid;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        long long t = ref_len_term(in.first[i], in.last[i]);
        if (t) atomicAdd(&s_ref[i / in.ntid], (unsigned long long)t);
    }
    __syncthreads();
    if (threadIdx.x == 0) finalize_rest(in, s_ref, n_records, d_cnt[CNT_A], S, density);
}

// ---- K2 ------------------------------------------------------------------------------------------
struct BreakFlag {   // scan input: 1 where a read starts a new candidate region
    const bdk_aread* ar; const bdk_summary_t* S;
    __device__ uint32_t operator()(uint32_t j, uint32_t) const {
        if (j == 0) return 1u;
        const bdk_aread a = ar[j - 1], b = ar[j];
        return k2_is_break(a.tid, a.pos, b.tid, b.pos, S->window) ? 1u : 0u;
    }
};
struct BreakOut {
    int32_t* read_cand; uint32_t* cand_first;
    __device__ void operator()(uint32_t j, uint32_t inc, uint32_t v, uint32_t) const {
        read_cand[j] = (int32_t)inc - 1;
        if (v) cand_first[inc - 1] = j;
    }
};

struct CandInfo { int32_t fwd, rev, nonctx, accept; };

__global__ void __launch_bounds__(GS_THREADS) k2_candidates_kernel(const bdk_aread* __restrict__ ar, const uint32_t* __restrict__ cand_first,
        const uint32_t* __restrict__ d_cnt, int32_t min_len, int32_t cov_lim, int32_t* __restrict__ cand_maxlen, CandInfo* __restrict__ cand_info) {
    const uint32_t A = d_cnt[CNT_A], ncand = d_cnt[CNT_NCAND];
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < ncand; c += gridDim.x * blockDim.x) {
        const int64_t s = cand_first[c], e = (int64_t)(c + 1 < ncand ? cand_first[c + 1] : A) - 1;
        const CandAgg g = k2_cand_aggregate(ar, s, e, (int64_t)A);
        cand_maxlen[c] = g.maxlen;
        CandInfo ci; ci.fwd = g.fwd; ci.rev = g.rev; ci.nonctx = g.nonctx;
        ci.accept = k2_accept(ar[s].pos, ar[e].pos, g, min_len, cov_lim) ? 1 : 0;
        cand_info[c] = ci;
    }
}

struct AcceptFlag {
    const CandInfo* ci;
    __device__ uint32_t operator()(uint32_t c, uint32_t) const { return (uint32_t)ci[c].accept; }
};
struct RegionOut {   // writes the region table and the read -> region map
    const bdk_aread* ar; const uint32_t* cand_first; const CandInfo* ci; const uint32_t* d_cnt;
    RegionRec* reg; int32_t* read_region; int32_t* cand_regs; int32_t dummy, chr_restricted, min_read_pair;
    __device__ void operator()(uint32_t c, uint32_t inc, uint32_t v, uint32_t ncand) const {
        const uint32_t A = d_cnt[CNT_A];
        cand_regs[c] = (int32_t)inc + dummy;          // regions registered up to and including candidate c = index of the first region behind it
        const uint32_t s = cand_first[c], e = (c + 1 < ncand ? cand_first[c + 1] : A) - 1;
        int32_t r = -1;
        if (v) {
            r = (int32_t)(inc - 1) + dummy;
            const CandInfo k = ci[c];
            RegionRec R;
            R.tid = ar[s].tid; R.start = ar[s].pos; R.end = ar[e].pos; R.fwd = k.fwd; R.rev = k.rev;
            R.first_read = (int32_t)s; R.n_reads = (int32_t)(e - s + 1);
            const int valid = chr_restricted ? k.nonctx : R.n_reads;
            R.stored = valid >= min_read_pair ? 1 : 0; R.cand = (int32_t)c;
            reg[r] = R;
        }
        for (uint32_t j = s; j <= e; ++j) read_region[j] = r;
        if (c == 0 && dummy) {   // region 0 of a run with -s < 0: registered from empty state
            RegionRec D; D.tid = -1; D.start = -1; D.end = -1; D.fwd = 0; D.rev = 0; D.first_read = 0; D.n_reads = 0; D.stored = 0; D.cand = -1;
            reg[0] = D;
        }
    }
};

// ---- K3 ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (uint32_t)x;
}

// Open-addressing table of read indices keyed by the read-name key; the second read of a name
// finds the first one and both learn their mate. A name with MORE than two reads among the anomalous ones (bams with
// overlapping read names, a 64-bit key collision) is flagged on its first read (the table entry every later read of the
// name meets); k3_links_kernel then leaves all reads of such a name without a mate: they are never paired and keep their
// regions from being cleared, the job goes on (the reference pairs whichever two of them a process_sv call meets first).
__global__ void __launch_bounds__(GS_THREADS) k3_mate_join_kernel(const bdk_aread* __restrict__ ar, uint32_t A, uint32_t* __restrict__ table,
        uint32_t mask, int32_t* __restrict__ mate, uint8_t* __restrict__ dup, uint32_t* __restrict__ d_cnt) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < A; j += gridDim.x * blockDim.x) {
        const unsigned long long q = ar[j].qid;
        uint32_t h = hash64(q) & mask;
        for (;;) {
            const uint32_t prev = atomicCAS(table + h, 0xffffffffu, j);
            if (prev == 0xffffffffu) break;
            if (ar[prev].qid == q) {
                const int32_t old = atomicExch(mate + prev, (int32_t)j);
                if (old != -1) { dup[prev] = 1; atomicAdd(d_cnt + CNT_NDUP, 1u); }
                mate[j] = (int32_t)prev;
                break;
            }
            h = (h + 1) & mask;
        }
    }
}

// one link per pair whose two reads were both registered: key = (earlier region << 32 | later region).
// Links are aggregated straight into weighted edges in an open-addressing table (key -> count); the edge list is
// never sorted -- everything downstream is order-independent until each component's edges are ranked.
constexpr unsigned long long EDGE_EMPTY = ~0ull;
__global__ void __launch_bounds__(GS_THREADS) k3_links_kernel(int32_t* __restrict__ mate, const uint8_t* __restrict__ dup, const int32_t* __restrict__ read_region, uint32_t A,
        unsigned long long* __restrict__ tkeys, uint32_t* __restrict__ tcnt, uint32_t mask) {
    for (uint32_t y = blockIdx.x * blockDim.x + threadIdx.x; y < A; y += gridDim.x * blockDim.x) {
        int32_t x = mate[y];
        // every read of a name with more than two reads points at the name's first read (or is it): all lose their mate. Only
        // this thread reads or writes mate[y] in this kernel.
        if (x >= 0 && (dup[y] || dup[x])) { mate[y] = -1; x = -1; }
        if (x < 0 || (uint32_t)x >= y) continue;
        const int32_t rx = read_region[x], ry = read_region[y];
        if (rx < 0 || ry < 0) continue;
        const unsigned long long key = ((unsigned long long)(uint32_t)rx << 32) | (uint32_t)ry;
        uint32_t h = hash64(key) & mask;
        for (;;) {
            const unsigned long long prev = atomicCAS(tkeys + h, EDGE_EMPTY, key);
            if (prev == EDGE_EMPTY || prev == key) { atomicAdd(tcnt + h, 1u); break; }
            h = (h + 1) & mask;
        }
    }
}

// the weight of edge (r0 <= r1) from the link table (0: no such edge)
__device__ __forceinline__ uint32_t k3_edge_weight(const unsigned long long* __restrict__ tkeys, const uint32_t* __restrict__ tcnt, uint32_t mask, int r0, int r1) {
    const unsigned long long key = ((unsigned long long)(uint32_t)r0 << 32) | (uint32_t)r1;
    uint32_t h = hash64(key) & mask;
    for (;;) {
        const unsigned long long k = tkeys[h];
        if (k == key) return tcnt[h];
        if (k == EDGE_EMPTY) return 0;
        h = (h + 1) & mask;
    }
}

// per-read static information for K4 (bdk_logic.h: ReadInfo2), one thread per anomalous read
__global__ void __launch_bounds__(GS_THREADS) k3_read_info_kernel(const bdk_aread* __restrict__ ar, const int32_t* __restrict__ mate,
        const int32_t* __restrict__ read_region, const int32_t* __restrict__ read_cand, const RegionRec* __restrict__ reg, const int32_t* __restrict__ cand_regs,
        const uint32_t* __restrict__ d_cnt, int period, int min_read_pair, const unsigned long long* __restrict__ tkeys, const uint32_t* __restrict__ tcnt, uint32_t mask, uint32_t A,
        ReadInfo2* __restrict__ ri, int32_t* __restrict__ sv_of_read) {
    (void)d_cnt;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < A; j += gridDim.x * blockDim.x) {
        const int m = mate[j];
        bool strong = false;
        if (m >= 0) {
            const int rj = read_region[j], rm = read_region[m];
            if (rj >= 0 && rm >= 0 && rj != rm) strong = (int)k3_edge_weight(tkeys, tcnt, mask, min(rj, rm), max(rj, rm)) >= min_read_pair;
        }
        ri[j] = k4n_make_read_info(ar, mate, read_region, read_cand, reg, cand_regs, period, (int)j, strong);
        sv_of_read[j] = -1;
    }
}

// ---- the followed edges (weight >= -r), bucketed by flush window ------------------------------------------------------
// An edge (r0 <= r1) belongs to the window in which r1 is registered. Pass 1 counts per window the directed copies and
// the edges themselves (= call slots); one CTA scans the windows; pass 2 puts the directed copies (src << 32 | dst) into
// their window's range, in arrival order -- each window's warp sorts its own few edges (k4n_windows_kernel).
__global__ void __launch_bounds__(GS_THREADS) k3_strong_count_kernel(const unsigned long long* __restrict__ tkeys, const uint32_t* __restrict__ tcnt, uint32_t tsize,
        int32_t min_read_pair, int period, uint32_t* __restrict__ wdir, uint32_t* __restrict__ wund) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < tsize; e += gridDim.x * blockDim.x) {
        const unsigned long long k = tkeys[e];
        if (k == EDGE_EMPTY || (int32_t)tcnt[e] < min_read_pair) continue;
        const uint32_t r0 = (uint32_t)(k >> 32), r1 = (uint32_t)k, w = r1 / (uint32_t)period;
        atomicAdd(wdir + w, r0 != r1 ? 2u : 1u);
        atomicAdd(wund + w, 1u);
    }
}
constexpr int K3S_THREADS = 1024;
__global__ void __launch_bounds__(K3S_THREADS) k3_window_scan_kernel(const uint32_t* __restrict__ wdir, const uint32_t* __restrict__ wund, int period,
        uint32_t* __restrict__ wstart, uint32_t* __restrict__ slot_base, uint32_t* __restrict__ wfill, uint32_t* __restrict__ d_cnt) {
    __shared__ uint32_t s_a[33], s_b[33];
    const uint32_t nwin = d_cnt[CNT_NREG] / (uint32_t)period + 1;
    uint32_t run_a = 0, run_b = 0;
    for (uint32_t base = 0; base < nwin; base += K3S_THREADS) {
        const uint32_t w = base + threadIdx.x;
        const uint32_t a = w < nwin ? wdir[w] : 0, b = w < nwin ? wund[w] : 0;
        uint32_t ta, tb;
        const uint32_t ia = ss_block_scan_any(a, s_a, &ta), ib = ss_block_scan_any(b, s_b, &tb);
        if (w < nwin) { wstart[w] = run_a + ia - a; slot_base[w] = run_b + ib - b; wfill[w] = 0; }
        run_a += ta; run_b += tb;
    }
    if (threadIdx.x == 0) { d_cnt[CNT_NSE] = run_a; d_cnt[CNT_NROW] = run_b; }
}
__global__ void __launch_bounds__(GS_THREADS) k3_strong_scatter_kernel(const unsigned long long* __restrict__ tkeys, const uint32_t* __restrict__ tcnt, uint32_t tsize,
        int32_t min_read_pair, int period, const uint32_t* __restrict__ wstart, uint32_t* __restrict__ wfill, unsigned long long* __restrict__ se) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < tsize; e += gridDim.x * blockDim.x) {
        const unsigned long long k = tkeys[e];
        if (k == EDGE_EMPTY || (int32_t)tcnt[e] < min_read_pair) continue;
        const uint32_t r0 = (uint32_t)(k >> 32), r1 = (uint32_t)k, w = r1 / (uint32_t)period;
        const uint32_t n = r0 != r1 ? 2u : 1u;
        const uint32_t at = wstart[w] + atomicAdd(wfill + w, n);
        se[at] = k;
        if (n == 2) se[at + 1] = ((unsigned long long)r1 << 32) | r0;
    }
}

// ---- K4: the connection walk in closed form (bdk_logic.h, "second formulation") ------------------------------------
//   sweeps   one persistent cooperative kernel: the table of deletion windows starts empty; sweep 0 evaluates every region,
//            a later sweep the regions stamped by a neighbour whose entry changed; the table is rewritten in place; grid-wide
//            barrier between sweeps, done when a sweep changes nothing
//   calls    first call window per region; one warp per flush window sorts the window's followed edges in shared memory and
//            orders its calls (build_connection); one warp per call counts its pairs (process_sv); one thread per call
//            scores it (k4_score_kernel)
// Regions are taken 32 at a time: a tile of eight lanes evaluates a region (four at a time per warp), the whole warp a large one.
constexpr int K4_THREADS = 256;
constexpr int K4N_TILE_MAX = 256;         // reads up to which eight lanes evaluate a region (the whole warp beyond)
constexpr int K4_TRACE_SWEEPS = 64;
struct K4Trace { unsigned long long t[1 + K4_TRACE_SWEEPS]; uint32_t nchanged[K4_TRACE_SWEEPS]; };
struct K4Tab {
    int32_t* del;                    // [nreg] flush window in which the region is cleared (K4_NEVER: never)
    uint32_t* stamp;                 // [nreg] sweep in which the region is evaluated again
    int32_t* c1;                     // [nreg] first window with a call involving the region
    const uint32_t* d_cnt;
    int32_t count_changes;           // trace: count the changed regions (else: a flag)
};

template <class Fn>
__device__ __forceinline__ void k4n_for_block(const K4N& S, uint32_t base, uint32_t v_end, bool want, Fn fn) {
    const unsigned FULL = 0xffffffffu;
    const uint32_t v = base + lane_id();
    const bool mine = v < v_end && want;
    const int nr = mine ? S.reg[v].n_reads : 0;
    const unsigned small = __ballot_sync(FULL, mine && nr <= K4N_TILE_MAX), big = __ballot_sync(FULL, mine && nr > K4N_TILE_MAX);
    unsigned m = (small >> (lane_id() & 24u)) & 0xffu;         // the eight regions of this lane's tile
    while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        fn(Tile8Team(), (int)(base + (lane_id() & 24u) + l));
    }
    __syncwarp();
    unsigned bg = big;
    while (bg) {
        const int l = __ffs(bg) - 1;
        bg &= bg - 1;
        fn(WarpTeam(), (int)(base + l));
    }
}

// grid-wide barrier of a cooperative launch (all CTAs resident): monotone arrival counter. The spin uses relaxed loads
// (an acquire load per iteration would invalidate L1 every time); the fences on both sides order the data.
__device__ __forceinline__ void k4_grid_barrier(uint32_t* counter, uint32_t& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++epoch;
        __threadfence();
        atomicAdd(counter, 1u);
        const uint32_t target = epoch * gridDim.x;
        uint32_t v;
        for (;;) {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v >= target) break;
            __nanosleep(32);
        }
        __threadfence();
    }
    __syncthreads();
}

// one sweep: the regions stamped `sweep` (sweep 0: all) against the table, which is rewritten in place; a region whose
// entry changes stamps the regions its reads' mates sit in for the next sweep
__device__ __forceinline__ void k4n_sweep(const K4N& S, const K4Tab& Tb, uint32_t sweep, uint32_t* changed) {
    const uint32_t nreg = (uint32_t)S.nreg;
    const uint32_t nwarp = gridDim.x * (blockDim.x >> 5), warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint32_t nchanged = 0;
    for (uint32_t base = warp * 32u; base < nreg; base += nwarp * 32u) {
        const uint32_t v0 = base + lane_id();
        const bool want = v0 < nreg && (sweep == 0 || __ldcg(Tb.stamp + v0) == sweep);
        k4n_for_block(S, base, nreg, want, [&](auto T, int v) {
            const int d = k4n_region_deletion(T, S, Tb.del, v);
            const int old = __ldcg(Tb.del + v);
            if (d == old) return;
            const RegionRec R = S.reg[v];
            if (T.lane() == 0) { __stcg(Tb.del + v, d); ++nchanged; }
            for (int j = R.first_read + T.lane(); j < R.first_read + R.n_reads; j += T.width()) {
                const int rm = S.ri[j].mate_region;
                if (rm >= 0 && rm != v) __stcg(Tb.stamp + rm, sweep + 1);
            }
        });
    }
    nchanged = __reduce_add_sync(0xffffffffu, nchanged);
    if (nchanged && lane_id() == 0) { if (Tb.count_changes) atomicAdd(changed, nchanged); else __stcg(changed, 1u); }
}

// sync[0]: barrier counter; sync[1 + s % 3]: regions changed in sweep s; sync[7]: number of sweeps (result)
__global__ void __launch_bounds__(K4_THREADS) k4n_sweeps_kernel(K4N S, K4Tab Tb, uint32_t* __restrict__ sync, K4Trace* __restrict__ trace) {
    S.nreg = (int32_t)Tb.d_cnt[CNT_NREG];
    uint32_t epoch = 0;
    const bool tr = trace && blockIdx.x == 0 && threadIdx.x == 0;
    if (tr) trace->t[0] = globaltimer_ns();
    for (uint32_t sweep = 0;; ++sweep) {
        if (blockIdx.x == 0 && threadIdx.x == 0) sync[1 + (sweep + 1) % 3] = 0;      // last read two barriers ago
        k4n_sweep(S, Tb, sweep, sync + 1 + sweep % 3);
        k4_grid_barrier(sync, epoch);
        const uint32_t nchanged = ld_acquire_u32(sync + 1 + sweep % 3);
        if (tr && sweep < K4_TRACE_SWEEPS) { trace->t[1 + sweep] = globaltimer_ns(); trace->nchanged[sweep] = nchanged; }
        if (!nchanged) { if (blockIdx.x == 0 && threadIdx.x == 0) sync[7] = sweep + 1; break; }
    }
}

// the same, one launch per sweep (when the cooperative launch cannot be resident)
__global__ void __launch_bounds__(K4_THREADS) k4n_sweep_kernel(K4N S, K4Tab Tb, uint32_t sweep, uint32_t* __restrict__ changed) {
    S.nreg = (int32_t)Tb.d_cnt[CNT_NREG];
    k4n_sweep(S, Tb, sweep, changed);
}

__global__ void __launch_bounds__(K4_THREADS) k4n_first_call_kernel(K4N S, K4Tab Tb) {
    S.nreg = (int32_t)Tb.d_cnt[CNT_NREG];
    const uint32_t nreg = (uint32_t)S.nreg;
    for (uint32_t base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u; base < nreg; base += gridDim.x * (blockDim.x >> 5) * 32u)
        k4n_for_block(S, base, nreg, true, [&](auto T, int v) {
            const int c = k4n_first_call(T, S, Tb.del, v);
            if (T.lane() == 0) Tb.c1[v] = c;
        });
}

// ---- the calls of a flush window -------------------------------------------------------------------------------------
// One warp per window: its followed edges (a few to a few hundred directed copies) are sorted by (src, dst) in shared
// memory (bitonic), every lane resolves which of them are dead already (a region cleared before this window) and which
// regions were touched by an earlier window, then lane 0 runs build_connection's walk over the window (bdk_logic.h:
// k4n_window_calls) entirely out of shared memory. A window with more than K4W_CAP directed edges goes to the list of
// big windows (k4n_big_windows_kernel: a CTA each, sorted in global memory).
constexpr int K4W_CAP = 512, K4W_WARPS = 4;
struct K4Windows {
    const unsigned long long* se; const uint32_t* wstart; const uint32_t* wdir; const uint32_t* slot_base;
    unsigned long long* scratch;      // [2 * (number of directed edges)] big windows: padded copy to sort
    uint8_t* fl; int32_t* queue;      // big windows: flags [directed edges], queue [directed edges + windows]
    uint32_t* big_list; uint32_t* big_count;
    bdk_sv* rows; uint8_t* row_emit; int32_t period;
    uint32_t cap;                     // <= K4W_CAP (tests lower it to send windows to the big-window kernel)
};
template <class KeyPtr>
__device__ __forceinline__ void k4w_bitonic(KeyPtr key, uint32_t n2, uint32_t tid, uint32_t nthreads, bool cta) {
    for (uint32_t k = 2; k <= n2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < n2; i += nthreads) {
                const uint32_t x = i ^ j;
                if (x > i) {
                    const unsigned long long a = key[i], b = key[x];
                    if ((a > b) == ((i & k) == 0)) { key[i] = b; key[x] = a; }
                }
            }
            if (cta) __syncthreads(); else __syncwarp();
        }
}
__global__ void __launch_bounds__(K4W_WARPS * 32) k4n_windows_kernel(K4Tab Tb, K4Windows W) {
    __shared__ unsigned long long s_key[K4W_WARPS][K4W_CAP];
    __shared__ int32_t s_queue[K4W_WARPS][K4W_CAP + 1];
    __shared__ uint8_t s_fl[K4W_WARPS][K4W_CAP];
    const uint32_t nwin = Tb.d_cnt[CNT_NREG] / (uint32_t)W.period + 1;
    const uint32_t wib = threadIdx.x >> 5, lane = lane_id();
    for (uint32_t w = blockIdx.x * K4W_WARPS + wib; w < nwin; w += gridDim.x * K4W_WARPS) {
        const uint32_t n = W.wdir[w];
        if (!n) continue;
        if (n > W.cap) { if (lane == 0) W.big_list[atomicAdd(W.big_count, 1u)] = w; continue; }
        const uint32_t s = W.wstart[w];
        uint32_t n2 = 32; while (n2 < n) n2 <<= 1;
        unsigned long long* key = s_key[wib];
        for (uint32_t i = lane; i < n2; i += 32) key[i] = i < n ? W.se[s + i] : ~0ull;
        __syncwarp();
        k4w_bitonic(key, n2, lane, 32, false);
        const SEdge* e = reinterpret_cast<const SEdge*>(key);
        for (uint32_t k = lane; k < n; k += 32) k4n_window_prepare(Tb.del, Tb.c1, e, (int)n, s_fl[wib], (int)w, (int)k);
        __syncwarp();
        if (lane == 0) k4n_window_calls(e, (int)n, s_fl[wib], s_queue[wib], (int)w, (int)W.slot_base[w], W.rows, W.row_emit);
        __syncwarp();
    }
}
constexpr int K4WB_THREADS = 512;
__global__ void __launch_bounds__(K4WB_THREADS) k4n_big_windows_kernel(K4Tab Tb, K4Windows W) {
    const uint32_t nbig = *W.big_count;
    for (uint32_t b = blockIdx.x; b < nbig; b += gridDim.x) {
        const uint32_t w = W.big_list[b], n = W.wdir[w], s = W.wstart[w];
        uint32_t n2 = 32; while (n2 < n) n2 <<= 1;
        unsigned long long* key = W.scratch + 2 * (size_t)s;          // n2 < 2 n: the padded copies of different windows do not overlap
        for (uint32_t i = threadIdx.x; i < n2; i += K4WB_THREADS) key[i] = i < n ? W.se[s + i] : ~0ull;
        __syncthreads();
        k4w_bitonic(key, n2, threadIdx.x, K4WB_THREADS, true);
        const SEdge* e = reinterpret_cast<const SEdge*>(key);
        uint8_t* fl = W.fl + s;
        for (uint32_t k = threadIdx.x; k < n; k += K4WB_THREADS) k4n_window_prepare(Tb.del, Tb.c1, e, (int)n, fl, (int)w, (int)k);
        __syncthreads();
        if (threadIdx.x == 0) k4n_window_calls(e, (int)n, fl, W.queue + s + w, (int)w, (int)W.slot_base[w], W.rows, W.row_emit);
        __syncthreads();
    }
}

// eight lanes per call slot: the pairs the call consumes (bdk_logic.h: k4n_call)
__global__ void __launch_bounds__(K4_THREADS) k4n_calls_kernel(K4N S, K4NOut M, const uint32_t* __restrict__ d_cnt) {
    S.nreg = (int32_t)d_cnt[CNT_NREG];
    const uint32_t nrow = d_cnt[CNT_NROW];
    const Tile8Team T;
    for (uint32_t r = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); r < nrow; r += gridDim.x * (blockDim.x >> 3))
        if (M.row_emit[r] & K4_ROW_CALL) k4n_call(T, S, M, (int)r);
}

// second half of process_sv for every call slot left pending, one thread per slot
__global__ void __launch_bounds__(GS_THREADS) k4_score_kernel(K4Static S, K4Mut M, const bdk_summary_t* __restrict__ summary, const uint32_t* __restrict__ d_cnt) {
    S.nreg = (int32_t)d_cnt[CNT_NREG]; S.ncand = (int32_t)d_cnt[CNT_NCAND];
    S.covered_ref_len = summary->covered_ref_len;
    const uint32_t nrow = d_cnt[CNT_NROW];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrow; r += gridDim.x * blockDim.x)
        if (M.row_emit[r] == K4_ROW_PENDING) k4_score_row(S, M, (int)r);
}

// ---- output order --------------------------------------------------------------------------------------
// Call slots are numbered window by window and, inside a window, in the order build_connection makes the calls: slot
// order IS the reference's output order (BreakDancer.cpp:280-338). The emitted rows are compacted by one scan.
struct RowPack {           // per-row arrays, by slot (K4 output) or by output position (what the host receives)
    bdk_sv* rows; int32_t* lib_count; uint32_t* cn_count; float* cn;
};
struct EmitFlag {
    const uint8_t* row_emit;
    __device__ uint32_t operator()(uint32_t s, uint32_t) const { return row_emit[s] == K4_ROW_EMIT ? 1u : 0u; }
};
struct GatherOut {
    RowPack in, out; int32_t* slot_order; int nlib, nkey;
    __device__ void operator()(uint32_t s, uint32_t inc, uint32_t v, uint32_t) const {
        if (!v) { slot_order[s] = -1; return; }
        const uint32_t i = inc - 1;
        bdk_sv r = in.rows[s];
        r.order = (int32_t)i;
        out.rows[i] = r;
        slot_order[s] = (int32_t)i;
        for (int l = 0; l < nlib; ++l) out.lib_count[(size_t)i * nlib + l] = in.lib_count[(size_t)s * nlib + l];
        for (int k = 0; k < nkey; ++k) { out.cn_count[(size_t)i * nkey + k] = in.cn_count[(size_t)s * nkey + k]; out.cn[(size_t)i * nkey + k] = in.cn[(size_t)s * nkey + k]; }
    }
};

// Poisson tail known-answer entry point
__global__ void poisson_logsf_kernel(const double* __restrict__ lambda, const int32_t* __restrict__ k, double* __restrict__ out, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = poisson_log_sf(lambda[i], k[i]);
}

}  // namespace bdk
