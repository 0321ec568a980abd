// k1_classify.cuh -- K1: the one pass over the 25-byte-per-record hot columns.
//
// Per record (reference: IlluminaPEReadClassifier::classify, BamSummary::_analyze_bam,
// BreakDancer::push_read up to the point a read is found anomalous):
//   * classify against the library's cut-offs                          -> bdk::classify_record
//   * pass-1 statistics: proper-pair counts per read group, flag histogram per library,
//     first/last record of every (bam, tid) for the covered reference length
//   * pass-2 filter; kept proper pairs feed the per-key running counts (nread_ROI / nread_FR),
//     anomalous reads are compacted, in stream order, with their inclusive per-key counts.
//
// Layout of the work: a tile is 4096 consecutive records; each of the 8 warps of a CTA owns a
// contiguous 512-record span ("unit") and walks it in 4 iterations of 128 records, every lane
// loading 4 consecutive records with 16-byte (int32 columns), 8-byte (u16) and 4-byte (u8)
// streaming loads -> 100 bytes in flight per lane and iteration, fully coalesced.
// Phase 1 classifies and keeps only bit masks in registers; one block-wide exchange reserves
// the tile's segment in the staging array (one global atomic per tile); phase 2 turns the masks
// into ranks with warp ballots and writes the anomalous reads. Tiles are handed out round-robin
// to a persistent grid (a multiple of the 148 SMs).
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
constexpr uint32_t K1_ERR_RG = 1u, K1_ERR_OVERFLOW = 2u;
constexpr uint32_t RG_INVALID = 0x80000000u;                // rg_info: lib | srcbam << 8 | invalid << 31

struct K1Args {
    bdk_soa c;                 // device columns of this push (16-byte aligned)
    uint64_t n;                // records in this push
    uint32_t base_index;       // stream index of record 0 of this push
    const LibDev* libs;
    const uint32_t* rg_info;
    int32_t nrg, nlib, nbam, nkey, ntid;
    int32_t nrg_smem;          // read groups counted in shared memory (0: global atomics)
    ClassifyOpts co;
    bdk_aread* stage;          // anomalous reads, tile segments in arrival order
    uint32_t* stage_p;         // [stage_cap][nkey] unit-relative inclusive proper-pair counts
    uint32_t stage_cap;
    uint32_t* cursor;          // staging cursor
    uint32_t* unit_cnt;        // [units] anomalous reads per unit           (this push: + unit_base)
    uint32_t* unit_p;          // [units][nkey] kept proper pairs per unit
    uint32_t* tile_seg;        // [tiles] staging offset of the tile's segment (this push: + tile_base)
    uint64_t unit_base, tile_base;
    unsigned long long* rg_sproper;   // [nrg]
    uint32_t* hist;                   // [nlib][BDK_NUM_FLAGS]
    unsigned long long* first;        // [nbam][ntid]
    unsigned long long* last;
    uint32_t* err;
};

struct K1Rec4 {
    int32_t pos[4], mpos[4], tid[4], mtid[4], isz[4];
    uint32_t flag[4], mapq[4], rg[4];
};

__device__ __forceinline__ void k1_load4(const bdk_soa& c, uint64_t g, int nv, K1Rec4& r) {
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
            bool v = j < nv;
            r.pos[j] = v ? c.pos[g + j] : 0; r.mpos[j] = v ? c.mpos[g + j] : 0;
            r.tid[j] = v ? c.tid[g + j] : 0; r.mtid[j] = v ? c.mtid[g + j] : 0;
            r.isz[j] = v ? c.isize[g + j] : 0; r.flag[j] = v ? c.flag[g + j] : 0;
            r.mapq[j] = v ? c.mapq[g + j] : 0; r.rg[j] = v ? c.rgid[g + j] : 0;
        }
    }
}

// SINGLE_KEY: one copy-number key (the common single-bam run): the running count lives in a register.
template <bool SINGLE_KEY>
__global__ void __launch_bounds__(K1_THREADS, 4) k1_classify_kernel(const K1Args a) {
    extern __shared__ uint32_t s_dyn[];              // [nlib * 11] flag histogram, [nrg_smem] proper counts
    uint32_t* s_hist = s_dyn;
    uint32_t* s_rg = s_dyn + a.nlib * BDK_NUM_FLAGS;
    __shared__ unsigned long long s_wfirst[K1_WARPS][K1_MAXB], s_wlast[K1_WARPS][K1_MAXB];
    __shared__ unsigned long long s_whas[K1_WARPS];
    __shared__ unsigned long long s_rfirst[K1_MAXB], s_rlast[K1_MAXB];   // CTA-running first/last of s_cur_tid
    __shared__ unsigned long long s_rhas;
    __shared__ int s_cur_tid;
    __shared__ uint32_t s_wcnt[K1_WARPS], s_woff[K1_WARPS], s_seg;
    __shared__ uint32_t s_run[SINGLE_KEY ? 1 : K1_WARPS][SINGLE_KEY ? 1 : K1_MAXK];

    const unsigned FULL = 0xffffffffu;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int nhist = a.nlib * BDK_NUM_FLAGS;
    for (int i = threadIdx.x; i < nhist + a.nrg_smem; i += K1_THREADS) s_dyn[i] = 0;
    if (threadIdx.x == 0) { s_rhas = 0; s_cur_tid = -1; }
    __syncthreads();

    const uint64_t ntiles = div_up<uint64_t>(a.n, K1_TILE);
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t span = tile * K1_TILE + (uint64_t)warp * K1_UNIT;   // first record of this warp's unit
        const int32_t t0 = a.c.tid[tile * K1_TILE];                        // tid of the tile's first record
        // ---------------- phase 1: classify, statistics, masks -----------------------------------
        uint32_t amask = 0, pmask = 0, keys[K1_ITERS] = {0, 0, 0, 0};
        unsigned long long flags4 = 0;      // 16 x 4-bit final ReadFlag
        unsigned long long whas = 0;        // bams for which this warp recorded a first/last key
#pragma unroll
        for (int it = 0; it < K1_ITERS; ++it) {
            const uint64_t g = span + (uint64_t)it * 128 + (uint64_t)lane * 4;
            const int nv = g + 4 <= a.n ? 4 : (g < a.n ? (int)(a.n - g) : 0);
            K1Rec4 r;
            k1_load4(a.c, g, nv, r);
            uint32_t bam_item[4];
            uint32_t samebits = 0;          // items with tid == t0 (candidates for the tile-level first/last)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool v = j < nv;
                uint32_t info = 0;
                if (v) {
                    if (r.rg[j] < (uint32_t)a.nrg) info = __ldg(a.rg_info + r.rg[j]); else info = RG_INVALID;
                    if (info & RG_INVALID) { atomicOr(a.err, K1_ERR_RG); info = 0; }
                }
                const int lib = info & 0xffu;
                bam_item[j] = (info >> 8) & 0xffu;
                uint32_t cr = 0;
                LibDev L;
                if (v) {
                    L = a.libs[lib];
                    cr = classify_record(r.pos[j], r.mpos[j], r.tid[j], r.mtid[j], r.isz[j], r.flag[j], r.mapq[j], L, a.co);
                }
                const int bit = it * 4 + j;
                if (cr & CR_ANOM) { amask |= 1u << bit; flags4 |= (unsigned long long)(cr & CR_FLAG_MASK) << (4 * bit); }
                if (cr & CR_MPROPER) { pmask |= 1u << bit; if (!SINGLE_KEY) keys[it] |= (uint32_t)L.key << (8 * j); }
                // pass-1 proper-pair count per read group: one atomic per distinct read group and warp
                const bool sp = (cr & CR_SPROPER) != 0;
                const unsigned spm = __ballot_sync(FULL, sp);
                if (sp) {
                    const unsigned peers = __match_any_sync(spm, r.rg[j]);
                    if (lane == __ffs(peers) - 1) {
                        if (a.nrg_smem) atomicAdd(&s_rg[r.rg[j]], (uint32_t)__popc(peers));
                        else atomicAdd(&a.rg_sproper[r.rg[j]], (unsigned long long)__popc(peers));
                    }
                }
                const int hf = (cr >> CR_HIST_SHIFT) & 0xF;
                if (hf) atomicAdd(&s_hist[lib * BDK_NUM_FLAGS + hf], 1u);
                if (v) {
                    if (r.tid[j] == t0) samebits |= 1u << j;
                    else {   // tile straddles a chromosome boundary: rare, go straight to memory
                        const unsigned long long key = ((unsigned long long)(a.base_index + (uint32_t)(g + j)) << 32) | (uint32_t)r.pos[j];
                        const size_t bt = (size_t)bam_item[j] * a.ntid + r.tid[j];
                        if ((uint32_t)r.tid[j] < (uint32_t)a.ntid) { atomicMin(a.first + bt, key); atomicMax(a.last + bt, key); }
                    }
                }
            }
            // first / last record per source bam inside this warp's unit (index-major keys, so the
            // first hit of the lowest lane in the earliest iteration is the minimum)
            for (int b = 0; b < a.nbam; ++b) {
                uint32_t mb = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) if (((samebits >> j) & 1u) && bam_item[j] == (uint32_t)b) mb |= 1u << j;
                const unsigned has = __ballot_sync(FULL, mb != 0);
                if (!has) continue;
                const int lo = __ffs(has) - 1, hi = 31 - __clz(has);
                if (!((whas >> b) & 1ull) && lane == lo) {
                    const int j = __ffs(mb) - 1;
                    s_wfirst[warp][b] = ((unsigned long long)(a.base_index + (uint32_t)(g + j)) << 32) | (uint32_t)r.pos[j];
                }
                if (lane == hi) {
                    const int j = 31 - __clz(mb);
                    s_wlast[warp][b] = ((unsigned long long)(a.base_index + (uint32_t)(g + j)) << 32) | (uint32_t)r.pos[j];
                }
                whas |= 1ull << b;
            }
        }
        const uint32_t wcnt = __reduce_add_sync(FULL, (uint32_t)__popc(amask));
        if (lane == 0) { s_wcnt[warp] = wcnt; s_whas[warp] = whas; }
        __syncthreads();                                                     // S1
        // ---------------- block exchange: reserve the tile's staging segment ----------------------
        if (warp == 0) {
            uint32_t c = lane < K1_WARPS ? s_wcnt[lane] : 0, inc = c;
#pragma unroll
            for (int d = 1; d < K1_WARPS; d <<= 1) { uint32_t t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
            if (lane < K1_WARPS) s_woff[lane] = inc - c;
            if (lane == K1_WARPS - 1) {
                uint32_t seg = inc ? atomicAdd(a.cursor, inc) : 0;
                s_seg = seg;
                a.tile_seg[a.tile_base + tile] = seg;
                if (inc && seg + inc > a.stage_cap) atomicOr(a.err, K1_ERR_OVERFLOW);
            }
        } else if (threadIdx.x == 32) {
            // merge this tile's per-warp first/last into the CTA-running values of chromosome t0
            if (t0 != s_cur_tid) {
                unsigned long long h = s_rhas;
                if ((uint32_t)s_cur_tid < (uint32_t)a.ntid)
                    for (int b = 0; b < a.nbam; ++b)
                        if ((h >> b) & 1ull) {
                            atomicMin(a.first + (size_t)b * a.ntid + s_cur_tid, s_rfirst[b]);
                            atomicMax(a.last + (size_t)b * a.ntid + s_cur_tid, s_rlast[b]);
                        }
                s_rhas = 0; s_cur_tid = t0;
            }
            unsigned long long h = s_rhas;
            for (int w = 0; w < K1_WARPS; ++w) {
                unsigned long long wh = s_whas[w];
                for (int b = 0; b < a.nbam; ++b) {
                    if (!((wh >> b) & 1ull)) continue;
                    unsigned long long f = s_wfirst[w][b], l = s_wlast[w][b];
                    if (!((h >> b) & 1ull)) { s_rfirst[b] = f; s_rlast[b] = l; h |= 1ull << b; }
                    else { if (f < s_rfirst[b]) s_rfirst[b] = f; if (l > s_rlast[b]) s_rlast[b] = l; }
                }
            }
            s_rhas = h;
        }
        __syncthreads();                                                     // S2
        // ---------------- phase 2: ranks from ballots, write the anomalous reads --------------------
        {
            const uint32_t out0 = s_seg + s_woff[warp];
            uint32_t arun = 0;            // anomalous reads of this unit before the current iteration
            uint32_t prun = 0;            // SINGLE_KEY: kept proper pairs before the current iteration
            if (!SINGLE_KEY) { for (int k = lane; k < a.nkey; k += 32) s_run[warp][k] = 0; __syncwarp(); }
            const unsigned lt = lanemask_lt();
#pragma unroll
            for (int it = 0; it < K1_ITERS; ++it) {
                const uint32_t a4 = (amask >> (4 * it)) & 0xFu, p4 = (pmask >> (4 * it)) & 0xFu;
                unsigned ab[4], pb[4];
                uint32_t abefore = 0, atotal = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ab[j] = __ballot_sync(FULL, (a4 >> j) & 1u);
                    abefore += __popc(ab[j] & lt);
                    atotal += __popc(ab[j]);
                }
                if (SINGLE_KEY) {
                    uint32_t pbefore = 0, ptotal = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        pb[j] = __ballot_sync(FULL, (p4 >> j) & 1u);
                        pbefore += __popc(pb[j] & lt);
                        ptotal += __popc(pb[j]);
                    }
                    if (a4) {
                        uint32_t rank = abefore, pin = prun + pbefore;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            pin += (p4 >> j) & 1u;                       // inclusive of the read itself
                            if ((a4 >> j) & 1u) {
                                const uint32_t o = out0 + arun + rank++;
                                if (o < a.stage_cap) {
                                    const uint64_t i = span + (uint64_t)it * 128 + (uint64_t)lane * 4 + j;
                                    const uint32_t rg = a.c.rgid[i];
                                    const uint32_t info = __ldg(a.rg_info + rg);
                                    const uint32_t fl = a.c.flag[i];
                                    const int32_t isz = a.c.isize[i];
                                    bdk_aread rec;
                                    rec.pos = a.c.pos[i]; rec.tid = a.c.tid[i]; rec.qlen = a.c.qlen[i];
                                    rec.abs_isize = isz < 0 ? -isz : isz;
                                    const uint32_t fnib = (uint32_t)(flags4 >> (4 * (it * 4 + j))) & 0xFu;
                                    rec.meta = fnib | ((fl & 0x10u) ? 16u : 0u) | ((info & 0xffu) << 8) | ((uint32_t)a.c.mapq[i] << 16);
                                    rec.record = a.base_index + (uint32_t)i;
                                    rec.qid = a.c.qid[i];
                                    int4* dst = reinterpret_cast<int4*>(a.stage + o);
                                    const int4* src = reinterpret_cast<const int4*>(&rec);
                                    dst[0] = src[0]; dst[1] = src[1];
                                    a.stage_p[o] = pin;
                                }
                            }
                        }
                    }
                    prun += ptotal;
                } else {
                    // general case: one round of ballots per key present in this iteration
                    unsigned long long present = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) if ((p4 >> j) & 1u) present |= 1ull << ((keys[it] >> (8 * j)) & 0xffu);
                    present |= __shfl_xor_sync(FULL, present, 16); present |= __shfl_xor_sync(FULL, present, 8);
                    present |= __shfl_xor_sync(FULL, present, 4); present |= __shfl_xor_sync(FULL, present, 2);
                    present |= __shfl_xor_sync(FULL, present, 1);
                    const uint32_t obase = out0 + arun + abefore;
                    for (int k = 0; k < a.nkey; ++k) {
                        uint32_t pbefore = 0, ptotal = 0;
                        uint32_t mine = 0;
                        if ((present >> k) & 1ull) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const bool m = ((p4 >> j) & 1u) && ((keys[it] >> (8 * j)) & 0xffu) == (uint32_t)k;
                                const unsigned b = __ballot_sync(FULL, m);
                                pbefore += __popc(b & lt); ptotal += __popc(b);
                                mine |= (uint32_t)m << j;
                            }
                        }
                        const uint32_t base = s_run[warp][k];
                        if (a4) {
                            uint32_t rank = 0, pin = base + pbefore;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                pin += (mine >> j) & 1u;
                                if ((a4 >> j) & 1u) {
                                    const uint32_t o = obase + rank++;
                                    if (o < a.stage_cap) a.stage_p[(size_t)o * a.nkey + k] = pin;
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0 && ptotal) s_run[warp][k] = base + ptotal;
                        __syncwarp();
                    }
                    if (a4) {
                        uint32_t rank = abefore;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if ((a4 >> j) & 1u) {
                                const uint32_t o = out0 + arun + rank++;
                                if (o < a.stage_cap) {
                                    const uint64_t i = span + (uint64_t)it * 128 + (uint64_t)lane * 4 + j;
                                    const uint32_t rg = a.c.rgid[i];
                                    const uint32_t info = __ldg(a.rg_info + rg);
                                    const uint32_t fl = a.c.flag[i];
                                    const int32_t isz = a.c.isize[i];
                                    bdk_aread rec;
                                    rec.pos = a.c.pos[i]; rec.tid = a.c.tid[i]; rec.qlen = a.c.qlen[i];
                                    rec.abs_isize = isz < 0 ? -isz : isz;
                                    const uint32_t fnib = (uint32_t)(flags4 >> (4 * (it * 4 + j))) & 0xFu;
                                    rec.meta = fnib | ((fl & 0x10u) ? 16u : 0u) | ((info & 0xffu) << 8) | ((uint32_t)a.c.mapq[i] << 16);
                                    rec.record = a.base_index + (uint32_t)i;
                                    rec.qid = a.c.qid[i];
                                    int4* dst = reinterpret_cast<int4*>(a.stage + o);
                                    const int4* src = reinterpret_cast<const int4*>(&rec);
                                    dst[0] = src[0]; dst[1] = src[1];
                                }
                            }
                        }
                    }
                }
                arun += atotal;
            }
            // unit table: anomalous and per-key proper-pair totals of this warp's 512 records
            const uint64_t unit = a.unit_base + tile * K1_WARPS + warp;
            if (lane == 0) a.unit_cnt[unit] = arun;
            if (SINGLE_KEY) { if (lane == 0) a.unit_p[unit] = prun; }
            else { __syncwarp(); for (int k = lane; k < a.nkey; k += 32) a.unit_p[unit * a.nkey + k] = s_run[warp][k]; }
        }
    }
    // ---------------- CTA epilogue: flush the shared accumulators -----------------------------------
    __syncthreads();
    for (int i = threadIdx.x; i < nhist; i += K1_THREADS) if (s_hist[i]) atomicAdd(a.hist + i, s_hist[i]);
    for (int i = threadIdx.x; i < a.nrg_smem; i += K1_THREADS) if (s_rg[i]) atomicAdd(a.rg_sproper + i, (unsigned long long)s_rg[i]);
    if (threadIdx.x == 0 && (uint32_t)s_cur_tid < (uint32_t)a.ntid) {
        unsigned long long h = s_rhas;
        for (int b = 0; b < a.nbam; ++b)
            if ((h >> b) & 1ull) {
                atomicMin(a.first + (size_t)b * a.ntid + s_cur_tid, s_rfirst[b]);
                atomicMax(a.last + (size_t)b * a.ntid + s_cur_tid, s_rlast[b]);
            }
    }
}

// ---- exclusive scan of the unit table (one CTA; the table is a few hundred KB) -----------------
// cnt_off[u] = sum of unit_cnt[0..u), p_off[u][k] = sum of unit_p[0..u)[k]; totals[0] = A.
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS, 1) k1_scan_units_kernel(const uint32_t* __restrict__ unit_cnt,
        const uint32_t* __restrict__ unit_p, uint64_t nunits, int nkey, uint32_t* __restrict__ cnt_off,
        uint32_t* __restrict__ p_off, uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_part[SCAN_THREADS];
    const int t = threadIdx.x;
    const uint64_t per = div_up<uint64_t>(nunits, SCAN_THREADS);
    const uint64_t lo = min(nunits, per * t), hi = min(nunits, lo + per);
    for (int q = 0; q <= nkey; ++q) {            // q == 0: anomalous counts, q >= 1: key q - 1
        const uint32_t* src = q == 0 ? unit_cnt : unit_p + (q - 1);
        const int stride = q == 0 ? 1 : nkey;
        uint32_t* dst = q == 0 ? cnt_off : p_off + (q - 1);
        uint32_t s = 0;
        for (uint64_t u = lo; u < hi; ++u) s += src[u * stride];
        s_part[t] = s;
        __syncthreads();
        for (int d = 1; d < SCAN_THREADS; d <<= 1) {   // Hillis-Steele inclusive scan of the partials
            uint32_t v = t >= d ? s_part[t - d] : 0;
            __syncthreads();
            s_part[t] += v;
            __syncthreads();
        }
        uint32_t run = s_part[t] - s;
        for (uint64_t u = lo; u < hi; ++u) { uint32_t v = src[u * stride]; dst[u * stride] = run; run += v; }
        if (q == 0 && t == SCAN_THREADS - 1) totals[0] = s_part[t];
        __syncthreads();
    }
}

// ---- bring the staged reads into stream order and make the proper-pair counts global -----------
// Final position d of a staged read: cnt_off[unit] + rank inside the unit; the units of a tile are
// contiguous in the tile's staging segment.
__global__ void __launch_bounds__(256) k1_reorder_kernel(const bdk_aread* __restrict__ stage, const uint32_t* __restrict__ stage_p,
        const uint32_t* __restrict__ cnt_off, const uint32_t* __restrict__ p_off, const uint32_t* __restrict__ tile_seg,
        uint64_t nunits, const uint32_t* __restrict__ totals, int nkey, bdk_aread* __restrict__ ar, uint32_t* __restrict__ P) {
    const uint32_t A = totals[0];
    for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < A; d += gridDim.x * blockDim.x) {
        uint64_t lo = 0, hi = nunits;            // last unit with cnt_off[u] <= d
        while (hi - lo > 1) { uint64_t m = (lo + hi) >> 1; if (cnt_off[m] <= d) lo = m; else hi = m; }
        const uint64_t u = lo, tile = u / K1_WARPS;
        const uint32_t src = tile_seg[tile] + (d - cnt_off[tile * K1_WARPS]);
        const int4* s = reinterpret_cast<const int4*>(stage + src);
        int4* o = reinterpret_cast<int4*>(ar + d);
        o[0] = s[0]; o[1] = s[1];
        for (int k = 0; k < nkey; ++k) P[(size_t)k * A + d] = stage_p[(size_t)src * nkey + k] + p_off[u * nkey + k];
    }
}

}  // namespace bdk
