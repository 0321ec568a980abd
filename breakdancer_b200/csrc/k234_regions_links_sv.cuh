// k234_regions_links_sv.cuh -- the stages that run on the compacted anomalous-read stream
// (about 1-3 % of the records):
//   K2  break flags -> candidate regions (segmented reductions) -> accepted regions
//       (BreakDancer::push_read:209-241, process_breakpoint:244-264, ReadRegionData::add_region)
//   K3  mate join by read-name key (hash table), region-region mate links, radix sort +
//       run-length -> weighted edges (ReadRegionData.cpp:109-113, Graph.hpp:41-46)
//   K4  connected components (lock-free union-find), per-component connection walk, SV
//       evaluation and Poisson score (build_connection, process_sv, SvBuilder, ComputeProbScore)
// All element counts stay on the device (d_cnt[]); kernels are grid-stride over fixed grids.
#pragma once
#include "common.cuh"
#include "scan_sort.cuh"
#include "bdk_finalize.h"

namespace bdk {

enum { CNT_A = 0, CNT_NCAND, CNT_NREG, CNT_NLINK, CNT_NEDGE, CNT_NDE, CNT_NROW, CNT_ERR, CNT_NREG_REAL, CNT_K4_TICKET, CNT_NEMIT, CNT_NDIRTY, CNT_N };
constexpr uint32_t K3_ERR_DUPNAME = 1u;
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
    RegionRec* reg; int32_t* read_region; uint8_t* alive; int32_t dummy, chr_restricted, min_read_pair;
    __device__ void operator()(uint32_t c, uint32_t inc, uint32_t v, uint32_t ncand) const {
        const uint32_t A = d_cnt[CNT_A];
        const uint32_t s = cand_first[c], e = (c + 1 < ncand ? cand_first[c + 1] : A) - 1;
        int32_t r = -1; uint8_t st = 0;
        if (v) {
            r = (int32_t)(inc - 1) + dummy;
            const CandInfo k = ci[c];
            RegionRec R;
            R.tid = ar[s].tid; R.start = ar[s].pos; R.end = ar[e].pos; R.fwd = k.fwd; R.rev = k.rev;
            R.first_read = (int32_t)s; R.n_reads = (int32_t)(e - s + 1);
            const int valid = chr_restricted ? k.nonctx : R.n_reads;
            R.stored = valid >= min_read_pair ? 1 : 0; R.cand = (int32_t)c;
            reg[r] = R;
            st = (uint8_t)R.stored;
        }
        for (uint32_t j = s; j <= e; ++j) { read_region[j] = r; alive[j] = st; }
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
// finds the first one and both learn their mate.
__global__ void __launch_bounds__(GS_THREADS) k3_mate_join_kernel(const bdk_aread* __restrict__ ar, uint32_t A, uint32_t* __restrict__ table,
        uint32_t mask, int32_t* __restrict__ mate, uint32_t* __restrict__ d_cnt) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < A; j += gridDim.x * blockDim.x) {
        const unsigned long long q = ar[j].qid;
        uint32_t h = hash64(q) & mask;
        for (;;) {
            const uint32_t prev = atomicCAS(table + h, 0xffffffffu, j);
            if (prev == 0xffffffffu) break;
            if (ar[prev].qid == q) {
                const int32_t old = atomicExch(mate + prev, (int32_t)j);
                if (old != -1) atomicOr(d_cnt + CNT_ERR, K3_ERR_DUPNAME);
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
__global__ void __launch_bounds__(GS_THREADS) k3_links_kernel(const int32_t* __restrict__ mate, const int32_t* __restrict__ read_region, uint32_t A,
        unsigned long long* __restrict__ tkeys, uint32_t* __restrict__ tcnt, uint32_t mask) {
    for (uint32_t y = blockIdx.x * blockDim.x + threadIdx.x; y < A; y += gridDim.x * blockDim.x) {
        const int32_t x = mate[y];
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

__device__ __forceinline__ int uf_find(int32_t* parent, int x) {
    for (;;) {
        int p = parent[x];
        if (p == x) return x;
        int gp = parent[p];
        if (gp != p) parent[x] = gp;   // path halving (benign race: only ever points further up)
        x = p;
    }
}
__device__ __forceinline__ void uf_union(int32_t* parent, int a, int b) {
    for (;;) {
        a = uf_find(parent, a); b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }          // hook the larger root under the smaller
        if (atomicCAS(parent + a, a, b) == a) return;
    }
}

__global__ void __launch_bounds__(GS_THREADS) k3_init_regions_kernel(int32_t* __restrict__ parent, uint32_t* __restrict__ comp_ne,
        uint32_t* __restrict__ comp_strong, uint32_t* __restrict__ comp_fill, uint8_t* __restrict__ deleted, int2* __restrict__ win_range,
        const uint32_t* __restrict__ d_cnt) {
    const uint32_t nreg = d_cnt[CNT_NREG];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nreg; r += gridDim.x * blockDim.x) {
        parent[r] = (int32_t)r; comp_ne[r] = 0; comp_strong[r] = 0; comp_fill[r] = 0; deleted[r] = 0;
        win_range[r] = make_int2(0x7fffffff, -1);      // first / last flush window in which the region has an edge
    }
}

// Components = connected components over the edges the connection walk can follow (weight >= -r): only those
// couple two regions through shared reads. A weaker edge still makes both its ends "active" in its flush window
// (is_region_final is asked for them) and is kept, as a directed copy, in the edge list of each end's component.
__global__ void __launch_bounds__(GS_THREADS) k3_union_kernel(const unsigned long long* __restrict__ tkeys, const uint32_t* __restrict__ tcnt, uint32_t tsize,
                                                              int32_t* __restrict__ parent, int32_t min_read_pair) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < tsize; e += gridDim.x * blockDim.x) {
        const unsigned long long k = tkeys[e];
        if (k == EDGE_EMPTY || (int32_t)tcnt[e] < min_read_pair) continue;
        const int r0 = (int)(k >> 32), r1 = (int)(k & 0xffffffffu);
        if (r0 != r1) uf_union(parent, r0, r1);
    }
}

// parent[] -> root table (no unions after this)
__global__ void __launch_bounds__(GS_THREADS) k3_flatten_kernel(int32_t* __restrict__ parent, const uint32_t* __restrict__ d_cnt) {
    const uint32_t nreg = d_cnt[CNT_NREG];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nreg; r += gridDim.x * blockDim.x) {
        int x = (int)r;
        for (int p = parent[x]; p != x; p = parent[x]) x = p;      // roots never change any more; concurrent writers store roots
        parent[r] = x;
    }
}

__global__ void __launch_bounds__(GS_THREADS) k3_comp_count_kernel(const unsigned long long* __restrict__ tkeys, const uint32_t* __restrict__ tcnt, uint32_t tsize,
        const int32_t* __restrict__ root_of, uint32_t* __restrict__ comp_ne, uint32_t* __restrict__ comp_strong, int32_t min_read_pair) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < tsize; e += gridDim.x * blockDim.x) {
        const unsigned long long k = tkeys[e];
        if (k == EDGE_EMPTY) continue;
        const int r0 = (int)(k >> 32), r1 = (int)(k & 0xffffffffu);
        const int root0 = root_of[r0];
        atomicAdd(comp_ne + root0, 1u);
        if (r0 != r1) atomicAdd(comp_ne + root_of[r1], 1u);
        if ((int32_t)tcnt[e] >= min_read_pair) atomicAdd(comp_strong + root0, 1u);   // a followed edge: both ends in one component
    }
}

struct LoadU32 { const uint32_t* p; __device__ uint32_t operator()(uint32_t i, uint32_t) const { return p[i]; } };
struct ExclOut { uint32_t* o; __device__ void operator()(uint32_t i, uint32_t inc, uint32_t v, uint32_t) const { o[i] = inc - v; } };

__global__ void __launch_bounds__(GS_THREADS) k3_scatter_edges_kernel(const unsigned long long* __restrict__ tkeys, const uint32_t* __restrict__ tcnt, uint32_t tsize,
        const int32_t* __restrict__ root_of, const uint32_t* __restrict__ de_off, uint32_t* __restrict__ comp_fill, DEdge* __restrict__ de,
        int32_t* __restrict__ de_root, int32_t period, int2* __restrict__ win_range) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < tsize; e += gridDim.x * blockDim.x) {
        const unsigned long long k = tkeys[e];
        if (k == EDGE_EMPTY) continue;
        const int r0 = (int)(k >> 32), r1 = (int)(k & 0xffffffffu);
        const int win = r1 / period;          // r0 <= r1: the pair is counted when r1 is registered
        DEdge d; d.win = win; d.src = r0; d.dst = r1; d.w = (int)tcnt[e]; d.flags = 0;
        atomicMin(&win_range[r0].x, win); atomicMax(&win_range[r0].y, win);
        if (r0 != r1) { atomicMin(&win_range[r1].x, win); atomicMax(&win_range[r1].y, win); }
        const int root0 = root_of[r0];
        const uint32_t s0 = de_off[root0] + atomicAdd(comp_fill + root0, 1u);
        de[s0] = d; de_root[s0] = root0;
        if (r0 != r1) {                       // the copy seen from r1 goes to r1's component (the same one iff the edge is followed)
            const int root1 = root_of[r1];
            const uint32_t s1 = de_off[root1] + atomicAdd(comp_fill + root1, 1u);
            d.src = r1; d.dst = r0; de[s1] = d; de_root[s1] = root1;
        }
    }
}

// Rank sort of every component's directed edges by (win, src, dst), one thread per edge: the keys are unique, so
// the number of smaller edges in the component is the edge's final position. Components with more than
// DE_RANK_SORT_MAX edges are left to the walk (in-place heap sort).
__global__ void __launch_bounds__(GS_THREADS) k3_rank_edges_kernel(const DEdge* __restrict__ de, const int32_t* __restrict__ de_root,
        const uint32_t* __restrict__ de_off, const uint32_t* __restrict__ comp_ne, DEdge* __restrict__ de_sorted, const uint32_t* __restrict__ d_cnt) {
    const uint32_t nde = d_cnt[CNT_NDE];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nde; t += gridDim.x * blockDim.x) {
        const int root = de_root[t];
        const uint32_t lo = de_off[root], n = comp_ne[root];
        if (n > (uint32_t)DE_RANK_SORT_MAX) continue;
        const DEdge x = de[t];
        uint32_t r = 0;
        for (uint32_t j = 0; j < n; ++j) r += de_less(de[lo + j], x) ? 1u : 0u;
        de_sorted[lo + r] = x;
    }
}

// ---- K4: one warp walks one connected component ------------------------------------------------------
// Components are found 32 regions at a time (a region with edges is the root of its component); the warp
// then walks them one after the other, its lanes sharing the loops over the reads of the regions involved.
// The walks are repeated in sweeps (bdk_logic.h, K4Static) until the table of deletion times is stable:
//   walk phase   sweep 0: every component; sweep s > 0: the components stamped s
//   mark phase   one thread per directed edge: a component that looks across a never-followed edge at a region
//                whose deletion time differs from the table it used is stamped s + 1
//   next phase   del_prev <- del_cur; the regions of the stamped components start again from "never cleared"
// Single GPU: one persistent cooperative kernel runs all sweeps with grid-wide barriers in between
// (k4_sweeps_kernel). Multi-GPU: one launch per phase, the deletion times are all-reduced between walk and mark.
constexpr int K4_THREADS = 128;
struct K4Graph {
    const uint32_t* comp_ne; const uint32_t* comp_strong; const uint32_t* de_off; const uint32_t* row_off;
    DEdge* de; DEdge* de_sorted; const int32_t* de_root; int32_t* queue;
    uint32_t* stamp;                 // [nreg] by root: sweep in which the component is walked again
    int32_t* del_prev;               // = S.del_prev, writable for the next phase
    const int2* win_range;           // [nreg] first / last flush window in which the region is active
    uint8_t* never_final;            // = S.never_final, written by k4_guess_kernel
    const bdk_summary_t* summary; uint32_t* d_cnt;
    uint32_t v_lo, v_hi;             // this GPU walks the components whose root region is in [v_lo, v_hi); single GPU: [0, ~0)
};

// A component with more than K4_BIG directed edges is walked by its warp for milliseconds to seconds. It waits (keeps its
// stamp for the next sweep) while smaller components are still changing, so that it is walked as few times as possible.
constexpr uint32_t K4_BIG = 4096;
__device__ __forceinline__ void k4_walk_phase(K4Static& S, K4Mut& M, const K4Graph& G, uint32_t sweep, uint32_t* ticket,
                                              bool defer_big = false, uint32_t* n_dirty = nullptr) {
    const unsigned FULL = 0xffffffffu;
    const WarpTeam T;
    const uint32_t lane = lane_id();
    S.rerun = sweep ? 1 : 0;
    const uint32_t v_end = min(G.v_hi, (uint32_t)S.nreg);
    for (;;) {                                  // 32 regions at a time, handed out dynamically (components differ a lot in size)
        uint32_t base = 0;
        if (lane == 0) base = G.v_lo + atomicAdd(ticket, 1u) * 32u;
        base = __shfl_sync(FULL, base, 0);
        if (base >= v_end) break;
        const uint32_t r = base + lane;
        uint32_t ne = (r < v_end && (!sweep || G.stamp[r] == sweep)) ? G.comp_ne[r] : 0;
        if (defer_big && ne > K4_BIG) { G.stamp[r] = sweep + 1; atomicAdd(n_dirty, 1u); ne = 0; }
        unsigned m = __ballot_sync(FULL, ne != 0);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t rr = base + src;
            const int n = (int)__shfl_sync(FULL, ne, src);
            DEdge* e = n <= DE_RANK_SORT_MAX ? G.de_sorted + G.de_off[rr] : de_sort_team(T, G.de + G.de_off[rr], n, (DEdge*)nullptr);
            k4_component(T, S, M, e, n, G.queue + G.de_off[rr] + 2 * (size_t)rr, (int)G.row_off[rr], (int)G.comp_strong[rr]);
        }
    }
}

__device__ __forceinline__ void k4_mark_phase(const K4Static& S, const K4Mut& M, const K4Graph& G, uint32_t sweep, uint32_t* n_dirty,
                                              uint32_t* n_small = nullptr) {
    const uint32_t nde = G.d_cnt[CNT_NDE];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nde; t += gridDim.x * blockDim.x) {
        const int root = G.de_root[t];
        const DEdge x = G.de[t];                               // de[]: the component's edges as a set (sorted or not)
        if (S.root_of[x.dst] == root) continue;
        const int a = S.del_prev[x.dst], b = M.del_cur[x.dst];
        if (a == b || G.stamp[root] == sweep + 1) continue;
        if (S.never_final[x.src]) continue;                    // x.src is never checked against other regions' state
        const int2 wr = G.win_range[x.src];
        if (k4_change_matters(x.src, x.dst, wr.x, min(wr.y, M.del_cur[x.src]), a, b)) {
            G.stamp[root] = sweep + 1; atomicAdd(n_dirty, 1u);
            if (n_small && G.comp_ne[root] <= K4_BIG) atomicAdd(n_small, 1u);
        }
    }
}

// reset_stamped: multi-GPU only -- every rank forgets the deletion times of the components that will be walked again (their
// owner rewrites them, the min all-reduce then takes the owner's values); on a single GPU the walk resets its own regions.
__device__ __forceinline__ void k4_next_phase(const K4Static& S, const K4Mut& M, const K4Graph& G, uint32_t sweep, bool reset_stamped) {
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < (uint32_t)S.nreg; v += gridDim.x * blockDim.x) {
        G.del_prev[v] = M.del_cur[v];
        if (reset_stamped && G.stamp[S.root_of[v]] == sweep + 1) M.del_cur[v] = K4_NEVER;
    }
}

__device__ __forceinline__ void k4_load_counts(K4Static& S, const K4Graph& G) {
    S.nreg = (int32_t)G.d_cnt[CNT_NREG]; S.ncand = (int32_t)G.d_cnt[CNT_NCAND];
    S.covered_ref_len = G.summary->covered_ref_len;
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
            __nanosleep(64);
        }
        __threadfence();
    }
    __syncthreads();
}

// single GPU: all sweeps in one persistent kernel. sync[0]: barrier counter, sync[1..2]: walk tickets (alternating),
// sync[3..4]: stamped-component counts (alternating), sync[5]: number of sweeps done (result), sync[6..7]: stamped small
// components (alternating), sync[8]: there are big components
constexpr int K4_TRACE_SWEEPS = 32;
struct K4Trace { unsigned long long t[1 + 3 * K4_TRACE_SWEEPS]; uint32_t ndirty[K4_TRACE_SWEEPS]; };   // BDK_K4_TRACE=1: phase time stamps (ns)
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void __launch_bounds__(K4_THREADS) k4_sweeps_kernel(K4Static S, K4Mut M, K4Graph G, uint32_t* __restrict__ sync, K4Trace* __restrict__ trace) {
    k4_load_counts(S, G);
    uint32_t epoch = 0;
    const bool tr = trace && blockIdx.x == 0 && threadIdx.x == 0;
    if (tr) trace->t[0] = globaltimer_ns();
    uint32_t nsmall_prev = 1;                                      // sweep 0: the small components go first
    for (uint32_t sweep = 0;; ++sweep) {
        // big components wait while small ones are still being corrected (they are walked at the latest when nothing else is left)
        k4_walk_phase(S, M, G, sweep, sync + 1 + (sweep & 1), nsmall_prev != 0, sync + 3 + (sweep & 1));
        k4_grid_barrier(sync, epoch);
        if (tr && sweep < K4_TRACE_SWEEPS) trace->t[1 + 3 * sweep] = globaltimer_ns();
        if (blockIdx.x == 0 && threadIdx.x == 0) { sync[1 + ((sweep + 1) & 1)] = 0; sync[3 + ((sweep + 1) & 1)] = 0; sync[6 + ((sweep + 1) & 1)] = 0; }   // last used before the barrier two phases back
        k4_mark_phase(S, M, G, sweep, sync + 3 + (sweep & 1), sync + 6 + (sweep & 1));
        k4_grid_barrier(sync, epoch);
        if (tr && sweep < K4_TRACE_SWEEPS) trace->t[2 + 3 * sweep] = globaltimer_ns();
        const uint32_t ndirty = ld_acquire_u32(sync + 3 + (sweep & 1));
        nsmall_prev = ld_acquire_u32(sync + 6 + (sweep & 1));
        k4_next_phase(S, M, G, sweep, false);
        k4_grid_barrier(sync, epoch);
        if (tr && sweep < K4_TRACE_SWEEPS) { trace->t[3 + 3 * sweep] = globaltimer_ns(); trace->ndirty[sweep] = ndirty; }
        if (!ndirty) { if (blockIdx.x == 0 && threadIdx.x == 0) sync[5] = sweep + 1; break; }
    }
}

// starting table of deletion times (bdk_logic.h: k4_guess_deletion), one thread per region
__global__ void __launch_bounds__(GS_THREADS) k4_guess_kernel(K4Static S, K4Mut M, K4Graph G) {
    k4_load_counts(S, G);
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < (uint32_t)S.nreg; v += gridDim.x * blockDim.x) {
        G.del_prev[v] = k4_guess_deletion(S, M.alive, (int)v, G.win_range[v].y);
        G.never_final[v] = k4_never_final(S, M.alive, (int)v) ? 1 : 0;
        M.del_cur[v] = K4_NEVER;
        G.stamp[v] = 0;
    }
}

// multi-GPU: one launch per phase
__global__ void __launch_bounds__(K4_THREADS) k4_components_kernel(K4Static S, K4Mut M, K4Graph G, uint32_t sweep, uint32_t* __restrict__ ticket) {
    k4_load_counts(S, G);
    k4_walk_phase(S, M, G, sweep, ticket);
}
__global__ void __launch_bounds__(GS_THREADS) k4_mark_dirty_kernel(K4Static S, K4Mut M, K4Graph G, uint32_t sweep, uint32_t* __restrict__ n_dirty) {
    k4_load_counts(S, G);
    k4_mark_phase(S, M, G, sweep, n_dirty);
}
__global__ void __launch_bounds__(GS_THREADS) k4_next_sweep_kernel(K4Static S, K4Mut M, K4Graph G, uint32_t sweep) {
    k4_load_counts(S, G);
    k4_next_phase(S, M, G, sweep, true);
}

// ---- output order --------------------------------------------------------------------------------------
// The reference prints window by window, BFS by BFS, and the calls of one BFS in the order they were made:
// sort the emitted rows by (key = window << 32 | BFS start vertex, row slot). One CTA, bitonic sort in
// shared memory (an SV table has thousands of rows, not millions).
constexpr int K5_THREADS = 1024;
constexpr int K5_SMEM_ROWS = 16384;                       // 12 bytes per row -> 192 KB
__global__ void __launch_bounds__(K5_THREADS, 1) k5_order_smem_kernel(const uint64_t* __restrict__ emit_key, const uint32_t* __restrict__ emit_slot,
        const uint32_t* __restrict__ d_cnt, uint32_t* __restrict__ order_slot) {
    extern __shared__ __align__(16) unsigned char s_k5[];
    const uint32_t n = d_cnt[CNT_NEMIT];
    uint32_t m = 1024; while (m < n) m <<= 1;            // n <= K5_SMEM_ROWS (checked by the host against the row-slot count)
    uint64_t* key = reinterpret_cast<uint64_t*>(s_k5);
    uint32_t* slot = reinterpret_cast<uint32_t*>(key + m);
    for (uint32_t i = threadIdx.x; i < m; i += K5_THREADS) {
        key[i] = i < n ? emit_key[i] : ~0ull;
        slot[i] = i < n ? emit_slot[i] : 0xffffffffu;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < (m >> 1); t += K5_THREADS) {
                const uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;   // lo has bit j clear
                const bool up = (lo & k) == 0;
                const uint64_t ka = key[lo], kb = key[hi];
                const uint32_t sa = slot[lo], sb = slot[hi];
                const bool gt = ka > kb || (ka == kb && sa > sb);
                if (gt == up) { key[lo] = kb; key[hi] = ka; slot[lo] = sb; slot[hi] = sa; }
            }
            __syncthreads();
        }
    for (uint32_t i = threadIdx.x; i < n; i += K5_THREADS) order_slot[i] = slot[i];
}

// large tables: order_slot comes from the device radix sort; this just seeds its value array
__global__ void __launch_bounds__(GS_THREADS) k5_copy_u32_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, const uint32_t* __restrict__ n_ptr) {
    const uint32_t n = *n_ptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// (key = slot, value = slot) pairs of the emitted rows, and the replacement of those keys by the rows' sort keys
__global__ void __launch_bounds__(GS_THREADS) k5_slot_keys_kernel(const uint64_t* __restrict__ emit_key, const uint32_t* __restrict__ emit_slot,
        const uint32_t* __restrict__ n_ptr, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t n = *n_ptr;
    (void)emit_key;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { keys[i] = emit_slot[i]; vals[i] = emit_slot[i]; }
}
__global__ void __launch_bounds__(GS_THREADS) k5_row_keys_kernel(const uint64_t* __restrict__ row_key, const uint32_t* __restrict__ n_ptr,
        unsigned long long* __restrict__ keys) {
    const uint32_t n = *n_ptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) keys[i] = row_key[keys[i]];
}

struct RowPack {           // per-row arrays, by slot (K4 output) or by output position (what the host receives)
    bdk_sv* rows; int32_t* lib_count; uint32_t* cn_count; float* cn;
};
__global__ void __launch_bounds__(GS_THREADS) k5_gather_kernel(RowPack in, RowPack out, const uint32_t* __restrict__ order_slot, int32_t* __restrict__ slot_order,
        int nlib, int nkey, const uint32_t* __restrict__ d_cnt, uint32_t* __restrict__ n_out) {
    const uint32_t n = d_cnt[CNT_NEMIT];
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = order_slot[i];
        bdk_sv r = in.rows[s];
        r.order = (int32_t)i;
        out.rows[i] = r;
        slot_order[s] = (int32_t)i;
        for (int l = 0; l < nlib; ++l) out.lib_count[(size_t)i * nlib + l] = in.lib_count[(size_t)s * nlib + l];
        for (int k = 0; k < nkey; ++k) { out.cn_count[(size_t)i * nkey + k] = in.cn_count[(size_t)s * nkey + k]; out.cn[(size_t)i * nkey + k] = in.cn[(size_t)s * nkey + k]; }
    }
}

// Poisson tail known-answer entry point
__global__ void poisson_logsf_kernel(const double* __restrict__ lambda, const int32_t* __restrict__ k, double* __restrict__ out, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = poisson_log_sf(lambda[i], k[i]);
}

}  // namespace bdk
