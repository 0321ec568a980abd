// k1_classify.cuh -- K1: the one pass over the 25-byte-per-record hot columns.
//
// Per record (reference: IlluminaPEReadClassifier::classify, BamSummary::_analyze_bam,
// BreakDancer::push_read up to the point a read is found anomalous):
//   * classify against the library's cut-offs                          -> bdk::classify_record
//   * pass-1 statistics: proper-pair counts per (library, bam), flag histogram per library
//   * pass-2 filter; kept proper pairs feed the per-key running counts (nread_ROI / nread_FR),
//     anomalous reads are compacted IN STREAM ORDER together with their inclusive per-key counts.
//
// Layout of the work: a tile is 4096 consecutive records; each of the 8 warps of a CTA owns a
// contiguous 512-record span and walks it in 4 iterations of 128 records, every lane loading 4
// consecutive records with 16-byte (int32 columns), 8-byte (u16) and 4-byte (u8) streaming loads
// -> 100 bytes in flight per lane and iteration, fully coalesced. The per-read-group constants
// (cut-offs, mapping-quality threshold, library / bam / key ids) are one 16-byte shared-memory
// load; the pass-1 proper-pair counters are thread-private shared-memory columns (no conflicts,
// no warp votes). Phase 1 keeps only two 16-bit masks per thread (anomalous, kept-proper).
// Tiles are handed out by a ticket counter to a persistent grid (a multiple of the 148 SMs) and
// chained by a decoupled look-back over (anomalous count, kept-proper count per key): every
// anomalous read gets its final position in the stream-ordered output and its global inclusive
// proper-pair counts in the same pass -- no staging, no second pass over the records.
// The covered-reference-length statistic (first / last record of every (bam, chromosome)) needs
// only the run boundaries of the sorted tid column: k1_span_kernel finds them by search.
#pragma once
#include "common.cuh"

namespace bdk {

constexpr int K1_THREADS = 256;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_IPT = 4;                                   // consecutive records per lane per iteration
constexpr int K1_ITERS = 4;
constexpr int K1_UNIT = 32 * K1_IPT * K1_ITERS;             // 512 records per warp and tile
constexpr int K1_TILE = K1_UNIT * K1_WARPS;                 // 4096
constexpr int K1_MAXK = 64;                                 // copy-number keys (bams, or libraries with -a)
constexpr int K1_MAXB = BDK_MAX_BAMS;
constexpr int K1_MAXCOMP = K1_MAXK + 1;                     // look-back vector: anomalous count + one per key
constexpr int K1_PRIV_CNT = 32;                             // (library, bam) pairs counted in private columns
constexpr int K1_RG_SMEM = 1023;                            // read groups whose constants live in shared memory
constexpr uint32_t K1_ERR_RG = 1u, K1_ERR_OVERFLOW = 2u;

// Per-read-group constants: the library's cut-offs and the ids the record maps to.
struct alignas(16) RgDev {
    float upper, lower;
    int32_t min_mapq;      // effective: library override or -q
    uint32_t info;         // RGI_* fields
};
constexpr uint32_t RGI_LIB_MASK = 0xffu;                    // bits 0-7   library index
constexpr int RGI_KEY_SHIFT = 8;                            // bits 8-13  copy-number key
constexpr int RGI_BAM_SHIFT = 14;                           // bits 14-19 source bam
constexpr int RGI_CNT_SHIFT = 20;                           // bits 20-24 private counter column (31: none)
constexpr uint32_t RGI_CNT_NONE = 31u;
constexpr uint32_t RGI_INVALID = 0x80000000u;               // read group without a library

struct K1Args {
    bdk_soa c;                 // device columns of this push (16-byte aligned); qlen / qid may be mapped host memory
    uint64_t n;                // records in this push
    uint32_t base_index;       // stream index of record 0 of this push
    const RgDev* rgtab;        // [nrg + 1]: entry nrg = invalid read group
    int32_t nrg, nlib, nbam, nkey;
    int32_t pad_rg;            // a read group with a library, used by padding records
    int32_t ncnt;              // private counter columns in use (0: count with warp votes + global atomics)
    const int32_t* cnt_rg;     // [ncnt] read group that receives the column's total
    ClassifyOpts co;
    bdk_aread* ar;             // [cap] anomalous reads in stream order
    uint32_t* P;               // [cap][nkey] inclusive kept-proper-pair counts per key at each anomalous read
    uint32_t cap;
    uint32_t* carry;           // [1 + nkey] anomalous reads / kept proper pairs per key before this push (updated)
    uint32_t* ticket;          // tile ticket counter (zero at launch)
    uint32_t* tile_status;     // [tiles] epoch << 2 | state
    uint32_t* tile_agg;        // [tiles][1 + nkey]
    uint32_t* tile_inc;        // [tiles][1 + nkey]
    unsigned long long* tile_bams;   // [tiles] source bams present in the tile (nbam > 1 only)
    uint32_t epoch;
    unsigned long long* rg_sproper;   // [nrg]
    uint32_t* hist;                   // [nlib][BDK_NUM_FLAGS]
    uint32_t* err;
};

struct K1Rec4 {
    int32_t pos[4], mpos[4], tid[4], mtid[4], isz[4];
    uint32_t flag[4], mapq[4], rg[4];
};

__device__ __forceinline__ void k1_load4(const bdk_soa& c, uint64_t g, int nv, uint32_t pad_rg, K1Rec4& r) {
    if (nv == 4) {
        int4 a = ld_stream_v4(c.pos + g);   r.pos[0] = a.x; r.pos[1] = a.y; r.pos[2] = a.z; r.pos[3] = a.w;
        int4 b = ld_stream_v4(c.mpos + g);  r.mpos[0] = b.x; r.mpos[1] = b.y; r.mpos[2] = b.z; r.mpos[3] = b.w;
        int4 d = ld_stream_v4(c.tid + g);   r.tid[0] = d.x; r.tid[1] = d.y; r.tid[2] = d.z; r.tid[3] = d.w;
        int4 e = ld_stream_v4(c.mtid + g);  r.mtid[0] = e.x; r.mtid[1] = e.y; r.mtid[2] = e.z; r.mtid[3] = e.w;
        int4 f = ld_stream_v4(c.isize + g); r.isz[0] = f.x; r.isz[1] = f.y; r.isz[2] = f.z; r.isz[3] = f.w;
        uint2 fl = ld_stream_v2(c.flag + g);
        r.flag[0] = fl.x & 0xffffu; r.flag[1] = fl.x >> 16; r.flag[2] = fl.y & 0xffffu; r.flag[3] = fl.y >> 16;
        uint32_t mq = ld_stream_u32(c.mapq + g);
        r.mapq[0] = mq & 0xffu; r.mapq[1] = (mq >> 8) & 0xffu; r.mapq[2] = (mq >> 16) & 0xffu; r.mapq[3] = mq >> 24;
        uint2 rg = ld_stream_v2(c.rgid + g);
        r.rg[0] = rg.x & 0xffffu; r.rg[1] = rg.x >> 16; r.rg[2] = rg.y & 0xffffu; r.rg[3] = rg.y >> 16;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bool v = j < nv;    // padding records: flag 0 (unpaired -> dropped by every filter), a valid read group
            r.pos[j] = v ? c.pos[g + j] : 0; r.mpos[j] = v ? c.mpos[g + j] : 0;
            r.tid[j] = v ? c.tid[g + j] : 0; r.mtid[j] = v ? c.mtid[g + j] : 0;
            r.isz[j] = v ? c.isize[g + j] : 0; r.flag[j] = v ? c.flag[g + j] : 0;
            r.mapq[j] = v ? c.mapq[g + j] : 0; r.rg[j] = v ? c.rgid[g + j] : pad_rg;
        }
    }
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const unsigned FULL = 0xffffffffu;
    const int lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULL, v, d); if (lane >= d) v += t; }
    return v;
}
// population count of each nibble of a 16-bit mask, one count per byte of the result
__device__ __forceinline__ uint32_t nibble_counts(uint32_t m) {
    m = m - ((m >> 1) & 0x5555u);
    m = (m & 0x3333u) + ((m >> 2) & 0x3333u);
    return (m & 0xFu) | ((m & 0xF0u) << 4) | ((m & 0xF00u) << 8) | ((m & 0xF000u) << 12);
}
__device__ __forceinline__ uint32_t byte_sum(uint32_t v) { return __dp4a(v, 0x01010101u, 0u); }
// sum of the bytes below byte `it`
__device__ __forceinline__ uint32_t bytes_before(uint32_t v, int it) { return byte_sum(v & ((1u << (8 * it)) - 1u)); }
__device__ __forceinline__ uint32_t byte_of(uint32_t v, int it) { return (v >> (8 * it)) & 0xffu; }

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
    uint32_t v; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}

enum : uint32_t { TS_AGG = 1u, TS_INC = 2u };

// SINGLE_KEY: one copy-number key (the common single-bam run): its counts live in registers.
// RG_SMEM: the read-group table fits in shared memory (nrg <= K1_RG_SMEM).
template <bool SINGLE_KEY, bool RG_SMEM>
__global__ void __launch_bounds__(K1_THREADS, 4) k1_classify_kernel(const K1Args a) {
    extern __shared__ int4 s_dyn4[];
    const int ncomp = 1 + a.nkey;
    const int nhist = a.nlib * BDK_NUM_FLAGS;
    const int ncol = a.ncnt > 1 ? a.ncnt : 0;                                      // one column: a register does it
    RgDev* s_rg = reinterpret_cast<RgDev*>(s_dyn4);                                // RG_SMEM: [nrg + 1]
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_rg + (RG_SMEM ? a.nrg + 1 : 0));   // [nlib * 11]
    uint32_t* s_cnt = s_hist + nhist;                                              // [ncol][K1_THREADS]
    uint32_t* s_px = s_cnt + ncol * K1_THREADS;                                    // !SINGLE_KEY: [nkey][K1_THREADS]
    uint32_t* s_pt = s_px + (SINGLE_KEY ? 0 : a.nkey * K1_THREADS);                // !SINGLE_KEY: [nkey][K1_WARPS]
    uint32_t* s_woff = s_pt + (SINGLE_KEY ? 0 : a.nkey * K1_WARPS);                // [ncomp][K1_WARPS]
    uint32_t* s_base = s_woff + ncomp * K1_WARPS;                                  // [ncomp]
    __shared__ uint32_t s_wa[K1_WARPS], s_wp[K1_WARPS];
    __shared__ unsigned long long s_wb[K1_WARPS];
    __shared__ uint32_t s_tile[2];

    const unsigned FULL = 0xffffffffu;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < nhist + ncol * K1_THREADS; i += K1_THREADS) s_hist[i] = 0;
    if (RG_SMEM) for (int i = threadIdx.x; i < a.nrg + 1; i += K1_THREADS) s_rg[i] = a.rgtab[i];
    uint32_t* my_cnt = s_cnt + threadIdx.x;                                        // this thread's private counter column
    uint32_t spcnt = 0;                                                            // ncnt == 1: pass-1 proper pairs seen by this thread
    uint32_t bad = 0;                                                              // OR of the info words (bit 31: invalid read group)
    if (threadIdx.x == 0) s_tile[0] = atomicAdd(a.ticket, 1u);
    __syncthreads();

    const uint32_t ntiles = (uint32_t)div_up<uint64_t>(a.n, K1_TILE);
    int par = 0;
    for (;;) {
        const uint32_t tile = s_tile[par];
        if (tile >= ntiles) break;
        if (threadIdx.x == 0) s_tile[par ^ 1] = atomicAdd(a.ticket, 1u);   // read after the next barrier
        par ^= 1;
        const uint64_t span = (uint64_t)tile * K1_TILE + (uint64_t)warp * K1_UNIT;   // first record of this warp's span
        // ---------------- phase 1: four decisions per record, kept as bit masks ------------------------
        uint32_t amask = 0, pmask = 0, hmask = 0, keys[K1_ITERS] = {0, 0, 0, 0};
        unsigned long long bams = 0;
#pragma unroll (SINGLE_KEY ? 1 : K1_ITERS)
        for (int it = 0; it < K1_ITERS; ++it) {
            const uint64_t g = span + (uint64_t)it * 128 + (uint64_t)lane * 4;
            const int nv = g + 4 <= a.n ? 4 : (g < a.n ? (int)(a.n - g) : 0);
            K1Rec4 r;
            k1_load4(a.c, g, nv, (uint32_t)a.pad_rg, r);
            uint32_t kk = 0, a4 = 0, p4 = 0, h4 = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t ri = min(r.rg[j], (uint32_t)a.nrg);
                RgDev L;
                if (RG_SMEM) L = s_rg[ri];
                else { const int4 q = __ldg(reinterpret_cast<const int4*>(a.rgtab) + ri); L.upper = __int_as_float(q.x); L.lower = __int_as_float(q.y); L.min_mapq = q.z; L.info = (uint32_t)q.w; }
                const uint32_t ch = classify_hot(r.pos[j], r.mpos[j], r.tid[j], r.mtid[j], r.isz[j], r.flag[j], r.mapq[j],
                                                 L.upper, L.lower, L.min_mapq, a.co);
                bad |= L.info;
                if (ch & CH_ANOM) a4 |= 1u << j;
                if (ch & CH_MPROPER) p4 |= 1u << j;
                if (ch & CH_HIST) h4 |= 1u << j;
                if (!SINGLE_KEY) kk |= ((L.info >> RGI_KEY_SHIFT) & 0x3fu) << (8 * j);
                if (a.nbam > 1) bams |= 1ull << ((L.info >> RGI_BAM_SHIFT) & 0x3fu);
                // pass-1 proper-pair count per (library, bam)
                const uint32_t sp = (ch >> 3) & 1u;          // CH_SPROPER
                if (a.ncnt == 1) spcnt += sp;
                else if (a.ncnt) atomicAdd(my_cnt + ((L.info >> RGI_CNT_SHIFT) & 31u) * K1_THREADS, sp);   // private column: no conflicts
                else {                                       // many (library, bam) pairs: one atomic per distinct read group and warp
                    const unsigned spm = __ballot_sync(FULL, sp && r.rg[j] < (uint32_t)a.nrg);
                    if ((spm >> lane) & 1u) {
                        const unsigned peers = __match_any_sync(spm, r.rg[j]);
                        if (lane == __ffs(peers) - 1) atomicAdd(&a.rg_sproper[r.rg[j]], (unsigned long long)__popc(peers));
                    }
                }
            }
            amask |= a4 << (4 * it); pmask |= p4 << (4 * it); hmask |= h4 << (4 * it);
            if (!SINGLE_KEY) keys[it] = kk;
        }
        // ---------------- warp level: ranks inside the 512-record span -----------------------------
        const uint32_t ca = nibble_counts(amask);
        const uint32_t ainc = warp_incl_scan(ca);
        const uint32_t aex = ainc - ca;                                   // per iteration: anomalous reads in lower lanes
        const uint32_t atot = __shfl_sync(FULL, ainc, 31);                // per iteration: anomalous reads of the warp
        uint32_t pex = 0, ptot = 0;
        if (SINGLE_KEY) {
            const uint32_t cp = nibble_counts(pmask);
            const uint32_t pinc = warp_incl_scan(cp);
            pex = pinc - cp;
            ptot = __shfl_sync(FULL, pinc, 31);
            if (lane == 0) { s_wa[warp] = byte_sum(atot); s_wp[warp] = byte_sum(ptot); }
        } else {
            if (lane == 0) s_wa[warp] = byte_sum(atot);
            for (int k = lane; k < a.nkey; k += 32) s_pt[k * K1_WARPS + warp] = 0;
            unsigned long long present = 0;
#pragma unroll
            for (int it = 0; it < K1_ITERS; ++it)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((pmask >> (it * 4 + j)) & 1u) present |= 1ull << ((keys[it] >> (8 * j)) & 0x3fu);
            present = (unsigned long long)__reduce_or_sync(FULL, (uint32_t)present) |
                      ((unsigned long long)__reduce_or_sync(FULL, (uint32_t)(present >> 32)) << 32);
            __syncwarp();
            while (present) {                                             // warp-uniform loop over the keys present
                const int k = __ffsll((long long)present) - 1;
                present &= present - 1;
                uint32_t c = 0;
#pragma unroll
                for (int it = 0; it < K1_ITERS; ++it)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (((pmask >> (it * 4 + j)) & 1u) && ((keys[it] >> (8 * j)) & 0x3fu) == (uint32_t)k) c += 1u << (8 * it);
                const uint32_t inc = warp_incl_scan(c);
                s_px[k * K1_THREADS + threadIdx.x] = inc - c;
                if (lane == 31) s_pt[k * K1_WARPS + warp] = inc;
            }
        }
        if (a.nbam > 1) {
            bams = (unsigned long long)__reduce_or_sync(FULL, (uint32_t)bams) | ((unsigned long long)__reduce_or_sync(FULL, (uint32_t)(bams >> 32)) << 32);
            if (lane == 0) s_wb[warp] = bams;
        }
        __syncthreads();                                                     // S1
        // ---------------- warp 0: tile totals, look-back --------------------------------------------
        if (warp == 0) {
            if (a.nbam > 1 && lane == 0) {
                unsigned long long b = 0;
                for (int w = 0; w < K1_WARPS; ++w) b |= s_wb[w];
                a.tile_bams[tile] = b;
            }
            // lane owns components lane, lane + 32, lane + 64 (component 0: anomalous, 1 + k: key k)
            uint32_t agg[3] = {0, 0, 0}, exc[3] = {0, 0, 0};
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int comp = lane + 32 * q;
                if (comp < ncomp) {
                    uint32_t run = 0;
                    for (int w = 0; w < K1_WARPS; ++w) {
                        uint32_t v;
                        if (comp == 0) v = s_wa[w];
                        else if (SINGLE_KEY) v = s_wp[w];
                        else v = byte_sum(s_pt[(comp - 1) * K1_WARPS + w]);
                        s_woff[comp * K1_WARPS + w] = run;
                        run += v;
                    }
                    agg[q] = run;
                }
                if (SINGLE_KEY && q == 0) break;
            }
            if (tile == 0) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int comp = lane + 32 * q;
                    if (comp < ncomp) { exc[q] = a.carry[comp]; a.tile_inc[(size_t)tile * ncomp + comp] = exc[q] + agg[q]; }
                    if (SINGLE_KEY && q == 0) break;
                }
                __threadfence();
                __syncwarp();
                if (lane == 0) st_release_u32(a.tile_status + tile, (a.epoch << 2) | TS_INC);
            } else {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int comp = lane + 32 * q;
                    if (comp < ncomp) a.tile_agg[(size_t)tile * ncomp + comp] = agg[q];
                    if (SINGLE_KEY && q == 0) break;
                }
                __threadfence();
                __syncwarp();
                if (lane == 0) st_release_u32(a.tile_status + tile, (a.epoch << 2) | TS_AGG);
                // look back over windows of 32 predecessor tiles (lane l inspects tile p - l)
                int p = (int)tile - 1;
                for (;;) {
                    const int q_tile = p - lane;
                    uint32_t st = TS_AGG;                                    // tiles before 0 contribute nothing
                    if (q_tile >= 0) {
                        do { st = ld_acquire_u32(a.tile_status + q_tile); } while ((st >> 2) != a.epoch);
                        st &= 3u;
                    }
                    const unsigned incm = __ballot_sync(FULL, st == TS_INC);
                    const int cut = incm ? __ffs(incm) - 1 : 31;            // nearest tile with an inclusive prefix
                    const bool use = lane <= cut && q_tile >= 0;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        for (int cl = 0; cl < 32 && 32 * q + cl < ncomp; ++cl) {
                            uint32_t v = 0;
                            if (use) v = ld_cg_u32((st == TS_INC ? a.tile_inc : a.tile_agg) + (size_t)q_tile * ncomp + 32 * q + cl);
                            v = __reduce_add_sync(FULL, v);
                            if (cl == lane) exc[q] += v;
                        }
                        if (SINGLE_KEY && q == 0) break;
                    }
                    if (incm) break;
                    p -= 32;
                }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int comp = lane + 32 * q;
                    if (comp < ncomp) a.tile_inc[(size_t)tile * ncomp + comp] = exc[q] + agg[q];
                    if (SINGLE_KEY && q == 0) break;
                }
                __threadfence();
                __syncwarp();
                if (lane == 0) st_release_u32(a.tile_status + tile, (a.epoch << 2) | TS_INC);
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int comp = lane + 32 * q;
                if (comp < ncomp) {
                    s_base[comp] = exc[q];
                    if (tile == ntiles - 1) a.carry[comp] = exc[q] + agg[q];   // prefix for the next push
                }
                if (SINGLE_KEY && q == 0) break;
            }
            if (lane == 0 && exc[0] + agg[0] > a.cap) atomicOr(a.err, K1_ERR_OVERFLOW);
        }
        __syncthreads();                                                     // S2
        // ---------------- phase 2: full classification of the flagged records (~1-3 %) ----------------
        // hmask (pass-1 histogram) is a superset of amask (anomalous reads to write out)
        if (hmask) {
            const uint32_t out0 = s_base[0] + s_woff[warp];                  // rank of the span's first anomalous read
            uint32_t m = hmask;
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                const int it = bit >> 2, j = bit & 3;
                const uint64_t i = span + (uint64_t)it * 128 + (uint64_t)lane * 4 + j;
                const uint32_t ri = min((uint32_t)a.c.rgid[i], (uint32_t)a.nrg);
                const RgDev L = RG_SMEM ? s_rg[ri] : a.rgtab[ri];
                const uint32_t mq = a.c.mapq[i];
                const int32_t pos = a.c.pos[i], tid = a.c.tid[i], isz = a.c.isize[i];
                const uint32_t cr = classify_record(pos, a.c.mpos[i], tid, a.c.mtid[i], isz, a.c.flag[i], mq, L.upper, L.lower, L.min_mapq, a.co);
                const uint32_t hf = (cr >> CR_HIST_SHIFT) & 0xFu;
                if (hf) atomicAdd(&s_hist[(L.info & RGI_LIB_MASK) * BDK_NUM_FLAGS + hf], 1u);
                if (!((amask >> bit) & 1u)) continue;
                const uint32_t a4 = (amask >> (4 * it)) & 0xFu, p4 = (pmask >> (4 * it)) & 0xFu;
                const uint32_t below = (2u << j) - 1u;                       // items 0..j of the iteration
                const uint32_t o = out0 + bytes_before(atot, it) + byte_of(aex, it) + __popc(a4 & (below >> 1));
                if (o >= a.cap) continue;
                bdk_aread rec;
                rec.pos = pos; rec.tid = tid; rec.qlen = a.c.qlen[i];
                rec.abs_isize = isz < 0 ? -isz : isz;
                rec.meta = make_meta(cr, (int)(L.info & RGI_LIB_MASK), mq);
                rec.record = a.base_index + (uint32_t)i;
                rec.qid = a.c.qid[i];
                int4* dst = reinterpret_cast<int4*>(a.ar + o);
                const int4* src = reinterpret_cast<const int4*>(&rec);
                dst[0] = src[0]; dst[1] = src[1];
                if (SINGLE_KEY) {
                    a.P[o] = s_base[1] + s_woff[K1_WARPS + warp] + bytes_before(ptot, it) + byte_of(pex, it) + __popc(p4 & below);
                } else {
                    for (int k = 0; k < a.nkey; ++k) {
                        uint32_t v = s_base[1 + k] + s_woff[(1 + k) * K1_WARPS + warp];
                        const uint32_t pt = s_pt[k * K1_WARPS + warp];
                        if (pt) {                                            // key k occurs in this warp's span
                            uint32_t mine = 0;
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj)
                                if (((p4 & below) >> jj) & 1u) mine += ((keys[it] >> (8 * jj)) & 0x3fu) == (uint32_t)k;
                            v += bytes_before(pt, it) + byte_of(s_px[k * K1_THREADS + threadIdx.x], it) + mine;
                        }
                        a.P[(size_t)o * a.nkey + k] = v;
                    }
                }
            }
        }
        // no barrier here: s_woff / s_base are rewritten by warp 0 only after S1 of the next tile, which every
        // warp reaches after its phase 2; s_pt / s_px rows are rewritten by the warp that alone reads them in phase 2.
    }
    // ---------------- CTA epilogue: flush the accumulators ---------------------------------------------
    if (__any_sync(FULL, (bad & RGI_INVALID) != 0) && lane == 0) atomicOr(a.err, K1_ERR_RG);
    if (a.ncnt == 1) {
        spcnt = __reduce_add_sync(FULL, spcnt);
        if (lane == 0 && spcnt) atomicAdd(a.rg_sproper + a.cnt_rg[0], (unsigned long long)spcnt);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nhist; i += K1_THREADS) if (s_hist[i]) atomicAdd(a.hist + i, s_hist[i]);
    for (int col = warp; col < ncol; col += K1_WARPS) {
        uint32_t s = 0;
        for (int t = lane; t < K1_THREADS; t += 32) s += s_cnt[col * K1_THREADS + t];
        s = __reduce_add_sync(FULL, s);
        if (lane == 0 && s) atomicAdd(a.rg_sproper + a.cnt_rg[col], (unsigned long long)s);
    }
}

// ---- first / last record of every (bam, chromosome) of one push -----------------------------------
// BamSummary::_analyze_bam's ref_len (BamSummary.cpp:70-74) telescopes to last - first position per
// (bam, tid). The stream is sorted by tid, so chromosome t occupies the run [lower_bound(t),
// lower_bound(t + 1)); one warp per (bam, t) finds the run by 32-ary search and then the first and
// the last record of the run that came from that bam (immediately, unless the bam is sparse there:
// then it skips 4096-record tiles using the per-tile bam masks K1 wrote).
__device__ __forceinline__ uint64_t warp_lower_bound_tid(const int32_t* __restrict__ tid, uint64_t n, int32_t t, int lane) {
    uint64_t lo = 0, hi = n;   // answer in [lo, hi]
    while (hi - lo > 0) {
        const uint64_t len = hi - lo;
        const uint64_t step = div_up<uint64_t>(len, 33);
        const uint64_t probe = lo + step * (lane + 1) - 1;          // 32 probes splitting the range into 33 parts
        const bool lt = probe < hi ? tid[probe] < t : false;
        const unsigned m = __ballot_sync(0xffffffffu, lt);          // monotone: a prefix of lanes
        const int c = __popc(m);
        const uint64_t nlo = c ? lo + step * c : lo;                // probes below are < t -> answer after the last of them
        const uint64_t nhi = c < 32 ? min(hi, lo + step * (c + 1) - 1) : hi;
        lo = nlo; hi = nhi;
    }
    return lo;
}

__global__ void __launch_bounds__(128) k1_span_kernel(const int32_t* __restrict__ tid, const int32_t* __restrict__ pos, const uint16_t* __restrict__ rgid,
        uint64_t n, uint32_t base_index, const RgDev* __restrict__ rgtab, int32_t nrg, int32_t nbam, int32_t ntid,
        const unsigned long long* __restrict__ tile_bams, unsigned long long* __restrict__ first, unsigned long long* __restrict__ last) {
    const unsigned FULL = 0xffffffffu;
    const int lane = lane_id();
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    if (n == 0) return;
    const int32_t t_lo = max(tid[0], 0), t_hi = min(tid[n - 1], ntid - 1);
    const int64_t nitems = (int64_t)(t_hi - t_lo + 1) * nbam;
    for (int64_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < nitems; item += nwarps) {
        const int32_t t = t_lo + (int32_t)(item / nbam);
        const uint32_t b = (uint32_t)(item % nbam);
        const uint64_t s = warp_lower_bound_tid(tid, n, t, lane), e = warp_lower_bound_tid(tid, n, t + 1, lane);
        if (s >= e) continue;
        auto from_bam = [&](uint64_t i) -> bool {
            const uint32_t rg = rgid[i];
            return rg < (uint32_t)nrg && ((rgtab[rg].info >> RGI_BAM_SHIFT) & 0x3fu) == b;
        };
        // forward
        uint64_t i0 = s, f = 0; bool found = false;
        while (i0 < e) {
            if (nbam > 1 && !((tile_bams[i0 / K1_TILE] >> b) & 1ull)) { i0 = (i0 / K1_TILE + 1) * K1_TILE; continue; }
            const uint64_t i = i0 + lane;
            const unsigned m = __ballot_sync(FULL, i < e && from_bam(i));
            if (m) { f = i0 + (__ffs(m) - 1); found = true; break; }
            i0 += 32;
        }
        if (!found) continue;
        // backward (stops at f at the latest)
        uint64_t l = f, i1 = e;
        while (i1 > f + 1) {
            if (nbam > 1 && !((tile_bams[(i1 - 1) / K1_TILE] >> b) & 1ull)) { i1 = (i1 - 1) / K1_TILE * K1_TILE; continue; }
            const int64_t i = (int64_t)i1 - 1 - lane;
            const unsigned m = __ballot_sync(FULL, i > (int64_t)f && from_bam((uint64_t)i));
            if (m) { l = i1 - 1 - (__ffs(m) - 1); break; }
            i1 = i1 > 32 ? i1 - 32 : 0;
        }
        if (lane == 0) {
            const unsigned long long kf = ((unsigned long long)(base_index + (uint32_t)f) << 32) | (uint32_t)pos[f];
            const unsigned long long kl = ((unsigned long long)(base_index + (uint32_t)l) << 32) | (uint32_t)pos[l];
            atomicMin(first + (size_t)b * ntid + t, kf);
            atomicMax(last + (size_t)b * ntid + t, kl);
        }
    }
}

}  // namespace bdk
