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

enum { CNT_A = 0, CNT_NCAND, CNT_NREG, CNT_NLINK, CNT_NEDGE, CNT_NDE, CNT_NROW, CNT_ERR, CNT_NREG_REAL, CNT_K4_TICKET, CNT_NEMIT, CNT_NDIRTY, CNT_NBIG, CNT_K4_NBIGLIST, CNT_K4_BIGCUR, CNT_N };
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
        int32_t* __restrict__ de_root, int32_t period, int2* __restrict__ win_range, const uint32_t* __restrict__ comp_ne, uint32_t* __restrict__ d_cnt) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < tsize; e += gridDim.x * blockDim.x) {
        const unsigned long long k = tkeys[e];
        if (k == EDGE_EMPTY) continue;
        const int r0 = (int)(k >> 32), r1 = (int)(k & 0xffffffffu);
        const int win = r1 / period;          // r0 <= r1: the pair is counted when r1 is registered
        DEdge d; d.win = win; d.src = r0; d.dst = r1; d.w = (int)tcnt[e]; d.flags = 0;
        atomicMin(&win_range[r0].x, win); atomicMax(&win_range[r0].y, win);
        if (r0 != r1) { atomicMin(&win_range[r1].x, win); atomicMax(&win_range[r1].y, win); }
        const int root0 = root_of[r0];
        const uint32_t f0 = atomicAdd(comp_fill + root0, 1u), s0 = de_off[root0] + f0;
        if (f0 == 0 && comp_ne[root0] > (uint32_t)DE_RANK_SORT_MAX) atomicAdd(d_cnt + CNT_NBIG, 1u);   // components too large for the rank sort
        de[s0] = d; de_root[s0] = root0;
        if (r0 != r1) {                       // the copy seen from r1 goes to r1's component (the same one iff the edge is followed)
            const int root1 = root_of[r1];
            const uint32_t f1 = atomicAdd(comp_fill + root1, 1u), s1 = de_off[root1] + f1;
            if (f1 == 0 && comp_ne[root1] > (uint32_t)DE_RANK_SORT_MAX) atomicAdd(d_cnt + CNT_NBIG, 1u);
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

// When some component is too large for the rank sort, ALL directed edges are sorted at once with the device radix sort:
// by (win, src, dst) packed into one key, then stably by the offset of the edge's component, which leaves every
// component's segment in place and sorted.
__global__ void __launch_bounds__(GS_THREADS) k3_edge_keys_kernel(const DEdge* __restrict__ de, const uint32_t* __restrict__ d_cnt, int vbits,
                                                                  unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t nde = d_cnt[CNT_NDE];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nde; t += gridDim.x * blockDim.x) {
        const DEdge x = de[t];
        keys[t] = ((unsigned long long)(uint32_t)x.win << (2 * vbits)) | ((unsigned long long)(uint32_t)x.src << vbits) | (uint32_t)x.dst;
        vals[t] = t;
    }
}
__global__ void __launch_bounds__(GS_THREADS) k3_edge_segment_keys_kernel(const uint32_t* __restrict__ vals, const int32_t* __restrict__ de_root,
        const uint32_t* __restrict__ de_off, const uint32_t* __restrict__ d_cnt, unsigned long long* __restrict__ keys) {
    const uint32_t nde = d_cnt[CNT_NDE];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nde; i += gridDim.x * blockDim.x) keys[i] = de_off[de_root[vals[i]]];
}
__global__ void __launch_bounds__(GS_THREADS) k3_edge_gather_kernel(const DEdge* __restrict__ de, const uint32_t* __restrict__ vals,
                                                                    const uint32_t* __restrict__ d_cnt, DEdge* __restrict__ de_sorted) {
    const uint32_t nde = d_cnt[CNT_NDE];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nde; i += gridDim.x * blockDim.x) de_sorted[i] = de[vals[i]];
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
constexpr int K4_THREADS = 256;
constexpr int K4_TRACE_SWEEPS = 32;
struct K4Trace {   // BDK_K4_TRACE=1: phase time stamps of the sweeps (ns) and, for CTA-walked components, time per stage (ns, summed)
    unsigned long long t[1 + 3 * K4_TRACE_SWEEPS]; uint32_t ndirty[K4_TRACE_SWEEPS];
    unsigned long long cta[8]; unsigned long long cta_windows, cta_pieces, cta_cands, cta_rounds, cta_chunks, cta_survivors, cta_maxreads;
};
struct K4Graph {
    const uint32_t* comp_ne; const uint32_t* comp_strong; const uint32_t* de_off; const uint32_t* row_off;
    DEdge* de; DEdge* de_sorted; const int32_t* de_root; int32_t* queue;
    uint32_t* stamp;                 // [nreg] by root: sweep in which the component is walked again
    int32_t* del_prev;               // = S.del_prev, writable for the next phase
    const int2* win_range;           // [nreg] first / last flush window in which the region is active
    uint8_t* never_final;            // = S.never_final, written by k4_guess_kernel
    const bdk_summary_t* summary; uint32_t* d_cnt;
    uint32_t v_lo, v_hi;             // this GPU walks the components whose root region is in [v_lo, v_hi); single GPU: [0, ~0)
    int32_t all_sorted;              // de_sorted holds every component's sorted edges (radix path), not only the rank-sorted ones
    uint32_t* big_list;              // roots of the components with more than K4_CTA_MIN directed edges (k4_guess_kernel), any order
    uint32_t* big_count;
    uint32_t cta_min, big_min;       // K4_CTA_MIN / K4_BIG (tests lower them to force those paths on small inputs)
    int32_t maxr;                    // <= K4C_MAXR (tests lower it to force the sequential-window fallback)
    int32_t defer_first;             // big components also sit out the first sweep (else they are walked once with everybody first)
    K4Trace* trace;                  // or null
};

// ---- a whole CTA walks one large component ---------------------------------------------------------------------------
// The sequential walk costs about 3 us per directed edge (a chain of dependent loads). For a component with more than
// K4_CTA_MIN directed edges the windows are still taken one after the other, but inside a window the independent pieces
// (bdk_logic.h: "the same window, split into independent pieces") are walked by the warps of the CTA concurrently, and the
// is_region_final pass is evaluated for all active nodes at once and then resolved. tests/hostsim runs the same
// decomposition on the host (component_by_pieces) against the oracle.
constexpr uint32_t K4_CTA_MIN = 512;      // directed edges from which a component gets a CTA instead of a warp
constexpr uint32_t K4_BIG = 4096;         // ... and from which it waits for the smaller components to settle before it is walked
constexpr int K4C_MAXR = 1024;            // distinct regions of one component in one window handled in shared memory (else: sequential window)
constexpr int K4C_MAXE = 1024;            // directed edges of one window staged in shared memory (the walk is a chain of dependent edge reads)
struct K4CtaSmem {
    DEdge edges[K4C_MAXE];
    int32_t vtx[K4C_MAXR], rs[K4C_MAXR + 1], label[K4C_MAXR], prow[K4C_MAXR], pq[K4C_MAXR], piece[K4C_MAXR], prowoff[K4C_MAXR], pqoff[K4C_MAXR],
            cand[K4C_MAXR];
    uint8_t state[K4C_MAXR];
    int32_t warp_tot[33];
    int32_t next_piece, changed, pending, cur, row_base;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// exclusive prefix of v over the CTA (all threads call it); total = sum over the CTA
__device__ __forceinline__ int k4_block_excl_scan(int v, int32_t* warp_tot, int& total) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int t = lane < nw ? warp_tot[lane] : 0;
        int ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(FULL, ti, d); if (lane >= d) ti += u; }
        warp_tot[lane] = ti - t;
        if (lane == 31) warp_tot[32] = ti;
    }
    __syncthreads();
    const int res = warp_tot[warp] + inc - v;
    total = warp_tot[32];
    __syncthreads();
    return res;
}

// ordered compaction of the indices r in [0, n) with pred(r) into out[]; returns their number (uniform)
template <class Pred>
__device__ __forceinline__ int k4_block_compact(int n, int32_t* out, int cap, int32_t* warp_tot, Pred pred) {
    int running = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int r = base + threadIdx.x;
        const int p = (r < n && pred(r)) ? 1 : 0;
        int tot;
        const int pos = running + k4_block_excl_scan(p, warp_tot, tot);
        if (p && pos < cap) out[pos] = r;
        running += tot;
    }
    __syncthreads();          // out[] is complete for every thread
    return running;
}

__device__ __forceinline__ int k4_run_of_edge(const K4CtaSmem& sm, int R, int t) {      // run whose edge range holds t
    int a = 0, b = R;
    while (a < b) { const int m = (a + b) >> 1; if (sm.rs[m + 1] <= t) a = m + 1; else b = m; }
    return a;
}
__device__ __forceinline__ int k4_run_of_vtx(const K4CtaSmem& sm, int R, int v) {       // run of region v (it has one: the reverse copy of the edge)
    int a = 0, b = R;
    while (a < b) { const int m = (a + b) >> 1; if (sm.vtx[m] < v) a = m + 1; else b = m; }
    return a < R && sm.vtx[a] == v ? a : -1;
}

__device__ void k4_component_cta(const K4Static& S, K4Mut& M, DEdge* e, int ne, int32_t* queue /* 2 * ne + 2 */, int row0, int nrows, K4CtaSmem& sm, int maxr, K4Trace* trace) {
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, NW = NT >> 5;
    const WarpTeam T;
    const bool tr = trace != nullptr && tid == 0;
    unsigned long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_last = tr ? globaltimer_ns() : 0, n_win = 0, n_pc = 0, n_cd = 0, n_rd = 0, n_ch = 0, n_sv = 0, n_mx = 0;
    auto lap = [&](int k) { if (tr) { const unsigned long long now = globaltimer_ns(); t_acc[k] += now - t_last; t_last = now; } };
    if (S.rerun) {
        for (int q = tid; q < ne; q += NT) k4_reset_slot(S, M, e, q);
        for (int r = tid; r < nrows; r += NT) M.row_emit[row0 + r] = 0;
    }
    __syncthreads();
    int i = 0, row_base = row0;
    while (i < ne) {
        const int w = e[i].win;
        int lo = i + 1, hi = ne;                                     // e[] is sorted by window: end of this one
        while (lo < hi) { const int m = (lo + hi) >> 1; if (e[m].win == w) lo = m + 1; else hi = m; }
        const int j = lo;
        const WindowInfo wi = k4_window_info(S, w);
        // The window's edges in shared memory: every step of the walk reads edges (binary searches for a region's run, scans
        // of a run, the flags). The erased marks only matter inside the window, so the copy is never written back.
        const int nE = j - i;
        DEdge* ew = e + i;                                            // the window's edges, indexed 0 .. nE
        if (nE <= K4C_MAXE) {
            const int* src = reinterpret_cast<const int*>(e + i);
            int* dst = reinterpret_cast<int*>(sm.edges);
            for (int t = tid; t < 5 * nE; t += NT) dst[t] = src[t];
            ew = sm.edges;
            __syncthreads();
        }
        lap(0);      // stage 0: window bounds + edge staging
        // ---- runs of equal source = the active nodes of the window, ascending ----------------------------------
        const int R = k4_block_compact(nE, sm.rs, K4C_MAXR, sm.warp_tot, [&](int t) { return t == 0 || ew[t].src != ew[t - 1].src; });
        if (R > maxr) {                                               // too many for shared memory: this window sequentially, by one warp
            if (warp == 0) { const int row = k4_window_seq(T, S, M, ew, 0, nE, w, wi, queue, row_base); if (lane == 0) sm.row_base = row; }
            __syncthreads();
            row_base = sm.row_base;
            __syncthreads();
            i = j;
            continue;
        }
        for (int r = tid; r < R; r += NT) { sm.vtx[r] = ew[sm.rs[r]].src; sm.label[r] = r; sm.prow[r] = 0; sm.pq[r] = 0; }
        if (tid == 0) sm.rs[R] = nE;
        __syncthreads();
        auto followable = [&](const DEdge& x) { return x.w >= S.min_read_pair && !M.deleted[x.dst] && !M.deleted[x.src]; };
        lap(1);      // stage 1: runs
        // ---- pieces: label propagation over the edges the walk would follow ------------------------------------
        for (;;) {
            if (tid == 0) sm.changed = 0;
            __syncthreads();
            for (int t = tid; t < nE; t += NT) {
                const DEdge x = ew[t];
                if (x.src == x.dst || !followable(x)) continue;
                const int r = k4_run_of_edge(sm, R, t), r2 = k4_run_of_vtx(sm, R, x.dst);
                if (r2 < 0) continue;
                const int la = ((volatile int32_t*)sm.label)[r], lb = ((volatile int32_t*)sm.label)[r2];
                if (la < lb) { atomicMin(&sm.label[r2], la); sm.changed = 1; }
                else if (lb < la) { atomicMin(&sm.label[r], lb); sm.changed = 1; }
            }
            __syncthreads();
            for (int r = tid; r < R; r += NT) {                        // pointer jumping (labels only ever decrease, within the piece)
                int l = ((volatile int32_t*)sm.label)[r];
                while (((volatile int32_t*)sm.label)[l] < l) l = ((volatile int32_t*)sm.label)[l];
                sm.label[r] = l;
            }
            __syncthreads();
            const int ch = sm.changed;
            __syncthreads();
            if (!ch) break;
        }
        lap(2);      // stage 2: piece labels
        for (int t = tid; t < nE; t += NT) {                           // per piece: edges it will follow (= row slots), queue entries
            const DEdge x = ew[t];
            if (!followable(x)) continue;
            const int root = sm.label[k4_run_of_edge(sm, R, t)];
            atomicAdd(&sm.pq[root], 1);
            if (x.src <= x.dst) atomicAdd(&sm.prow[root], 1);
        }
        __syncthreads();
        const int np = k4_block_compact(R, sm.piece, K4C_MAXR, sm.warp_tot, [&](int r) { return sm.label[r] == r && sm.prow[r] > 0; });
        int rows_total = 0, q_total = 0;
        for (int base = 0; base < np; base += NT) {
            const int p = base + tid;
            const int v = p < np ? sm.prow[sm.piece[p]] : 0;
            int tot, tot2;
            const int o1 = k4_block_excl_scan(v, sm.warp_tot, tot);
            const int o2 = k4_block_excl_scan(p < np ? v + 1 : 0, sm.warp_tot, tot2);
            if (p < np) { sm.prowoff[p] = rows_total + o1; sm.pqoff[p] = q_total + o2; }
            rows_total += tot; q_total += tot2;
        }
        if (tid == 0) sm.next_piece = 0;
        __syncthreads();
        lap(3);      // stage 3: piece list + offsets
        n_win += 1; n_pc += np;
        // ---- the pieces, a warp each ----------------------------------------------------------------------------
        for (;;) {
            int p = 0;
            if (lane == 0) p = atomicAdd(&sm.next_piece, 1);
            p = __shfl_sync(FULL, p, 0);
            if (p >= np) break;
            const int root = sm.piece[p];
            int row = row_base + sm.prowoff[p];
            int32_t* q = queue + sm.pqoff[p];
            for (int rb = root; rb < R; rb += 32) {                    // members of the piece, ascending (the root is its smallest)
                const int r = rb + lane;
                unsigned mask = __ballot_sync(FULL, r < R && sm.label[r] == root);
                while (mask) {
                    const int rr = rb + __ffs(mask) - 1;
                    mask &= mask - 1;
                    row = k4_bfs_from(T, S, M, ew, 0, nE, w, wi, sm.rs[rr], sm.rs[rr + 1], q, row);
                }
            }
        }
        __syncthreads();
        lap(4);      // stage 4: the pieces' walks
        row_base += rows_total;
        // ---- is_region_final over the active nodes: evaluate against the state before the pass, then resolve ----
        const int nc = k4_block_compact(R, sm.cand, K4C_MAXR, sm.warp_tot, [&](int r) {
            const int v = sm.vtx[r];
            return !S.never_final[v] && !M.deleted[v] && v != wi.last_region; });
        for (int c = tid; c < nc; c += NT) sm.cand[c] = sm.vtx[sm.cand[c]];          // run index -> region (still ascending)
        __syncthreads();
        // first the last 32 reads of every candidate, a warp per candidate (the reads whose mates lie ahead are a region's last
        // ones: most candidates are refused here); then all (candidate, 32 reads) chunks of the survivors spread over the warps
        for (int c = warp; c < nc; c += NW) {
            const RegionRec& Rg = S.reg[sm.cand[c]];
            const int jr = Rg.first_read + Rg.n_reads - 1 - lane;
            const bool bad = jr >= Rg.first_read && k4_read_blocks_final(S, M, jr, sm.cand[c], wi);
            const bool refused = __any_sync(FULL, bad) != 0;
            if (lane == 0) sm.state[c] = refused ? K4_FIN_NOT : K4_FIN_UNDECIDED;
        }
        __syncthreads();
        int chunks_total = 0;
        for (int base = 0; base < nc; base += NT) {
            const int c = base + tid;
            const int nr = c < nc && sm.state[c] == K4_FIN_UNDECIDED ? S.reg[sm.cand[c]].n_reads - 32 : 0;      // the last 32 are done
            const int v = nr > 0 ? (nr + 31) >> 5 : 0;
            int tot;
            const int o = k4_block_excl_scan(v, sm.warp_tot, tot);
            if (c < nc) sm.prowoff[c] = chunks_total + o;
            chunks_total += tot;
        }
        __syncthreads();
        if (tr) {
            n_ch += chunks_total;
            for (int c = 0; c < nc; ++c) if (sm.state[c] == K4_FIN_UNDECIDED) { n_sv += 1; if ((unsigned long long)S.reg[sm.cand[c]].n_reads > n_mx) n_mx = S.reg[sm.cand[c]].n_reads; }
        }
        for (int item = warp; item < chunks_total; item += NW) {
            int a = 0, b = nc;                                         // last candidate whose first chunk is <= item
            while (b - a > 1) { const int m = (a + b) >> 1; if (sm.prowoff[m] <= item) a = m; else b = m; }
            const RegionRec& Rg = S.reg[sm.cand[a]];
            const int jr = Rg.first_read + ((item - sm.prowoff[a]) << 5) + lane;
            const bool bad = jr < Rg.first_read + Rg.n_reads - 32 && k4_read_blocks_final(S, M, jr, sm.cand[a], wi);
            if (__any_sync(FULL, bad) && lane == 0) sm.state[a] = K4_FIN_NOT;
        }
        __syncthreads();
        lap(5);      // stage 5: is_region_final of the candidates
        n_cd += nc;
        for (;;) {
            if (tid == 0) sm.pending = 0;
            __syncthreads();
            for (int c = warp; c < nc; c += NW) {
                if (((volatile uint8_t*)sm.state)[c] != K4_FIN_UNDECIDED) continue;
                const int d = k4_final_deps(T, S, M, sm.cand[c], sm.cand, (const uint8_t*)sm.state, nc);
                if (lane == 0) {
                    if (d & 1) sm.state[c] = K4_FIN_NOT;
                    else if (!(d & 2)) sm.state[c] = K4_FIN_CLEARED;
                    else sm.pending = 1;
                }
            }
            __syncthreads();
            const int pend = sm.pending;
            __syncthreads();
            n_rd += 1;
            if (!pend) break;
        }
        for (int c = tid; c < nc; c += NT)
            if (sm.state[c] == K4_FIN_CLEARED) { M.deleted[sm.cand[c]] = 1; M.del_cur[sm.cand[c]] = w; }
        __syncthreads();
        lap(6);      // stage 6: resolution rounds + commit
        i = j;
    }
    if (tr) {
        for (int k = 0; k < 8; ++k) atomicAdd(&trace->cta[k], t_acc[k]);
        atomicAdd(&trace->cta_windows, n_win); atomicAdd(&trace->cta_pieces, n_pc); atomicAdd(&trace->cta_cands, n_cd); atomicAdd(&trace->cta_rounds, n_rd);
        atomicAdd(&trace->cta_chunks, n_ch); atomicAdd(&trace->cta_survivors, n_sv); atomicMax(&trace->cta_maxreads, n_mx);
    }
}

__device__ __forceinline__ void k4_walk_phase(K4Static& S, K4Mut& M, const K4Graph& G, uint32_t sweep, uint32_t* ticket, uint32_t* big_cursor,
                                              K4CtaSmem& sm, bool defer_big = false, uint32_t* n_dirty = nullptr) {
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
        if (ne > G.cta_min) ne = 0;             // large components: below, a CTA each
        unsigned m = __ballot_sync(FULL, ne != 0);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t rr = base + src;
            const int n = (int)__shfl_sync(FULL, ne, src);
            DEdge* e = G.de_sorted + G.de_off[rr];
            k4_component(T, S, M, e, n, G.queue + 2 * (size_t)G.de_off[rr] + 2 * (size_t)rr, (int)G.row_off[rr], (int)G.comp_strong[rr]);
        }
    }
    const uint32_t nbig = *G.big_count;
    if (!nbig) return;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sm.cur = (int32_t)atomicAdd(big_cursor, 1u);
        __syncthreads();
        const uint32_t idx = (uint32_t)sm.cur;
        if (idx >= nbig) break;
        const uint32_t rr = G.big_list[idx];
        if (rr < G.v_lo || rr >= v_end || (sweep && G.stamp[rr] != sweep)) continue;
        const uint32_t n = G.comp_ne[rr];
        if (defer_big && n > G.big_min) {          // waits for the small components to settle: stays stamped for the next sweep
            if (threadIdx.x == 0) { G.stamp[rr] = sweep + 1; atomicAdd(n_dirty, 1u); }
            continue;
        }
        DEdge* e = (n <= (uint32_t)DE_RANK_SORT_MAX || G.all_sorted) ? G.de_sorted + G.de_off[rr] : G.de + G.de_off[rr];
        if (!(n <= (uint32_t)DE_RANK_SORT_MAX || G.all_sorted)) {          // (only when the radix keys did not fit 64 bits)
            if (threadIdx.x < 32) de_sort_team(T, e, (int)n, (DEdge*)nullptr);
            __syncthreads();
        }
        k4_component_cta(S, M, e, (int)n, G.queue + 2 * (size_t)G.de_off[rr] + 2 * (size_t)rr, (int)G.row_off[rr], (int)G.comp_strong[rr], sm, G.maxr, n > G.big_min ? G.trace : nullptr);
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
            if (n_small && G.comp_ne[root] <= G.big_min) atomicAdd(n_small, 1u);
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
// components (alternating), sync[9..10]: cursors into the list of large components (alternating)

__global__ void __launch_bounds__(K4_THREADS, 3) k4_sweeps_kernel(K4Static S, K4Mut M, K4Graph G, uint32_t* __restrict__ sync, K4Trace* __restrict__ trace) {
    extern __shared__ __align__(16) unsigned char k4_smem_raw[];
    K4CtaSmem& sm = *reinterpret_cast<K4CtaSmem*>(k4_smem_raw);
    k4_load_counts(S, G);
    uint32_t epoch = 0;
    const bool tr = trace && blockIdx.x == 0 && threadIdx.x == 0;
    if (tr) trace->t[0] = globaltimer_ns();
    uint32_t nsmall_prev = G.defer_first ? 1 : 0;                  // sweep 0: the small components go first (or everybody at once)
    for (uint32_t sweep = 0;; ++sweep) {
        // big components wait while small ones are still being corrected (they are walked at the latest when nothing else is left)
        k4_walk_phase(S, M, G, sweep, sync + 1 + (sweep & 1), sync + 9 + (sweep & 1), sm, nsmall_prev != 0, sync + 3 + (sweep & 1));
        k4_grid_barrier(sync, epoch);
        if (tr && sweep < K4_TRACE_SWEEPS) trace->t[1 + 3 * sweep] = globaltimer_ns();
        if (blockIdx.x == 0 && threadIdx.x == 0) {     // the other parity's counters were last used before the barrier two phases back
            sync[1 + ((sweep + 1) & 1)] = 0; sync[3 + ((sweep + 1) & 1)] = 0; sync[6 + ((sweep + 1) & 1)] = 0; sync[9 + ((sweep + 1) & 1)] = 0;
        }
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

// per-read static information packed for the walk (bdk_logic.h: ReadInfo), one thread per anomalous read
__global__ void __launch_bounds__(GS_THREADS) k4_read_info_kernel(const bdk_aread* __restrict__ ar, const int32_t* __restrict__ mate,
        const int32_t* __restrict__ read_region, const int32_t* __restrict__ read_cand, uint32_t A, ReadInfo* __restrict__ ri) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < A; j += gridDim.x * blockDim.x) ri[j] = make_read_info(ar, mate, read_region, read_cand, (int)j);
}

// starting table of deletion times (bdk_logic.h: k4_guess_deletion), one thread per region
__global__ void __launch_bounds__(GS_THREADS) k4_guess_kernel(K4Static S, K4Mut M, K4Graph G) {
    k4_load_counts(S, G);
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < (uint32_t)S.nreg; v += gridDim.x * blockDim.x) {
        G.del_prev[v] = k4_guess_deletion(S, M.alive, (int)v, G.win_range[v].y);
        G.never_final[v] = k4_never_final(S, M.alive, (int)v) ? 1 : 0;
        if (G.comp_ne[v] > G.cta_min) G.big_list[atomicAdd(G.big_count, 1u)] = v;
        M.del_cur[v] = K4_NEVER;
        G.stamp[v] = 0;
    }
}

// after the sweeps: second half of process_sv for every row slot the walk left pending, one thread per slot
__global__ void __launch_bounds__(GS_THREADS) k4_score_kernel(K4Static S, K4Mut M, K4Graph G) {
    k4_load_counts(S, G);
    const uint32_t nrow = G.d_cnt[CNT_NROW];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrow; r += gridDim.x * blockDim.x)
        if (M.row_emit[r] == K4_ROW_PENDING) k4_score_row(S, M, (int)r);
}

// multi-GPU: one launch per phase
__global__ void __launch_bounds__(K4_THREADS, 3) k4_components_kernel(K4Static S, K4Mut M, K4Graph G, uint32_t sweep, uint32_t* __restrict__ ticket,
                                                                   uint32_t* __restrict__ big_cursor) {
    extern __shared__ __align__(16) unsigned char k4_smem_raw[];
    K4CtaSmem& sm = *reinterpret_cast<K4CtaSmem*>(k4_smem_raw);
    k4_load_counts(S, G);
    k4_walk_phase(S, M, G, sweep, ticket, big_cursor, sm);
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
constexpr int K5_THREADS = 256;
constexpr int K5_SMEM_ROWS = 16384;                       // tables up to this many row slots are ordered by the rank sort below
// Rank sort over the whole grid: (key, slot) pairs are unique, so the number of smaller pairs is a row's final position.
// CTA (bx, by) compares the 256 rows of block bx with the 256 rows of tile by (staged in shared memory) and adds its partial
// counts to rank[]; a second kernel scatters. n^2 comparisons (16 K rows = 2.7e8) spread over up to 4096 CTAs: a few
// microseconds, where a single-CTA bitonic sort of the same table took over 100 us.
__global__ void __launch_bounds__(K5_THREADS) k5_rank_partial_kernel(const uint64_t* __restrict__ emit_key, const uint32_t* __restrict__ emit_slot,
        const uint32_t* __restrict__ d_cnt, uint32_t* __restrict__ rank) {
    __shared__ uint64_t s_key[K5_THREADS];
    __shared__ uint32_t s_slot[K5_THREADS];
    const uint32_t n = d_cnt[CNT_NEMIT];
    const uint32_t i = blockIdx.x * K5_THREADS + threadIdx.x, t0 = blockIdx.y * K5_THREADS;
    if (blockIdx.x * K5_THREADS >= n || t0 >= n) return;           // uniform per CTA
    const uint32_t j = t0 + threadIdx.x;
    s_key[threadIdx.x] = j < n ? emit_key[j] : ~0ull;
    s_slot[threadIdx.x] = j < n ? emit_slot[j] : 0xffffffffu;
    __syncthreads();
    if (i >= n) return;
    const uint64_t ki = emit_key[i];
    const uint32_t si = emit_slot[i];
    const uint32_t m = min((uint32_t)K5_THREADS, n - t0);
    uint32_t r = 0;
    for (uint32_t q = 0; q < m; ++q) {
        const uint64_t kq = s_key[q];
        r += (kq < ki || (kq == ki && s_slot[q] < si)) ? 1u : 0u;
    }
    if (r) atomicAdd(rank + i, r);
}
__global__ void __launch_bounds__(GS_THREADS) k5_rank_scatter_kernel(const uint32_t* __restrict__ emit_slot, const uint32_t* __restrict__ rank,
        const uint32_t* __restrict__ d_cnt, uint32_t* __restrict__ order_slot) {
    const uint32_t n = d_cnt[CNT_NEMIT];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) order_slot[rank[i]] = emit_slot[i];
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
