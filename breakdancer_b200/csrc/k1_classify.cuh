// k1_classify.cuh -- K1: the one pass over the 25-byte-per-record hot columns.
//
// Per record (reference: IlluminaPEReadClassifier::classify, BamSummary::_analyze_bam,
// BreakDancer::push_read up to the point a read is found anomalous):
//   * classify against the library's cut-offs                          -> bdk::classify_hot / classify_record
//   * pass-1 statistics: proper-pair counts per (library, bam), flag histogram per library
//   * pass-2 filter; kept proper pairs feed the per-key running counts (nread_ROI / nread_FR),
//     anomalous reads are compacted IN STREAM ORDER together with their inclusive per-key counts.
//
// Shape of the kernel (persistent, two CTAs per SM, warp-specialised). Every CTA owns one contiguous
// range of 8192-record tiles, so its running counts are a local carry and no CTA ever waits for another:
//   * producer warp: streams the CTA's tiles through a 3-stage shared-memory ring, 1024 records
//     (25 600 B, eight column slices) per stage, with TMA bulk copies (cp.async.bulk ...
//     mbarrier::complete_tx) -- no registers, ~150 KB in flight per SM;
//   * 8 consumer warps: wait on a stage's "full" mbarrier, read 4 consecutive records per lane with
//     conflict-free 16/8/4-byte shared loads, make the four decisions of classify_hot() per record
//     and keep them as bit masks (32 records per thread and tile); the per-read-group constants are
//     one 16-byte shared load, the pass-1 proper-pair counters a register (one (library, bam) pair)
//     or thread-private shared-memory columns; the 1-3 % flagged records are classified in full on
//     the spot (histogram) and, if anomalous, parked in a thread-private stash slot;
//   * scan warp: turns the per-(stage, warp) totals of a tile into exclusive offsets and advances the
//     CTA's running counts. The consumers do not wait for it: they classify the next tile first and
//     only then write the previous tile's anomalous reads, in stream order, into the CTA's output
//     segment together with their CTA-relative inclusive proper-pair counts.
// k1_compact_kernel then moves the segments to their final place (prefix over the per-CTA totals)
// and makes the counts global. Every record is read from HBM exactly once; only the anomalous
// 1-3 % are written.
// The covered-reference-length statistic (first / last record of every (bam, chromosome)) needs
// only the run boundaries of the sorted tid column: k1_span_kernel finds them by search.
#pragma once
#include "common.cuh"

namespace bdk {

constexpr int K1_CWARPS = 8;                                // consumer warps
constexpr int K1_CTHREADS = K1_CWARPS * 32;                 // 256
constexpr int K1_THREADS = K1_CTHREADS + 64;                // + producer warp + scan warp
constexpr int K1_SUB = 1024;                                // records per ring stage (128 per consumer warp)
constexpr int K1_SUBS = 8;                                  // stages per look-back tile
constexpr int K1_TILE = K1_SUB * K1_SUBS;                   // 8192
#ifndef K1_STAGES_N
#define K1_STAGES_N 3
#endif
constexpr int K1_STAGES = K1_STAGES_N;                                // ring depth
constexpr int K1_STAGE_BYTES = K1_SUB * 25;                 // 25 600
constexpr int K1_STASH = 32;                                // stash slots per thread and tile (= records per thread)
constexpr int K1_MAXK = 64;                                 // copy-number keys (bams, or libraries with -a)
constexpr int K1_MAXB = BDK_MAX_BAMS;
constexpr int K1_PRIV_CNT = 32;                             // (library, bam) pairs counted in private columns
constexpr int K1_RG_SMEM = 255;                             // read groups whose constants live in shared memory
constexpr uint32_t K1_ERR_RG = 1u, K1_ERR_OVERFLOW = 2u;
// byte offsets of the column slices inside a ring stage
constexpr int K1_OFF_POS = 0, K1_OFF_MPOS = 4 * K1_SUB, K1_OFF_TID = 8 * K1_SUB, K1_OFF_MTID = 12 * K1_SUB, K1_OFF_ISZ = 16 * K1_SUB,
              K1_OFF_FLAG = 20 * K1_SUB, K1_OFF_RG = 22 * K1_SUB, K1_OFF_MAPQ = 24 * K1_SUB;
// named barriers (0 is __syncthreads)
constexpr int K1_BAR_A = 1, K1_BAR_B = 3, K1_BAR_C = 5;     // A[2]: totals ready, B[2]: offsets ready, C: consumers only

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

struct alignas(16) K1Stash { int32_t pos, tid, abs_isize; uint32_t meta; };

struct K1Args {
    bdk_soa c;                 // device columns of this push (16-byte aligned); qlen / qid may be mapped host memory
    uint64_t n;                // records in this push
    uint32_t base_index;       // stream index of record 0 of this push
    uint32_t tiles_per_cta;    // CTA b owns tiles [b * tiles_per_cta, (b + 1) * tiles_per_cta)
    const RgDev* rgtab;        // [nrg + 1]: entry nrg = invalid read group
    int32_t nrg, nlib, nbam, nkey;
    int32_t pad_rg;            // a read group with a library, used by padding records
    int32_t ncnt;              // private counter columns in use (0: count with warp votes + global atomics)
    int32_t cn_lib;            // -a: keys are libraries (else the source bams)
    const int32_t* cnt_rg;     // [ncnt] read group that receives the column's total
    ClassifyOpts co;
    bdk_aread* seg_ar;         // [grid][seg_cap] anomalous reads of each CTA's range, in stream order
    uint32_t* seg_P;           // [grid][seg_cap][nkey] CTA-relative inclusive kept-proper-pair counts per key
    uint32_t seg_cap;
    uint32_t* seg_cnt;         // [grid][1 + nkey] totals of each CTA: anomalous reads, kept proper pairs per key
    K1Stash* stash;            // [grid][2][K1_CTHREADS][K1_STASH]
    unsigned long long* tile_bams;   // [tiles] source bams present in the tile (nbam > 1 only)
    unsigned long long* rg_sproper;   // [nrg]
    uint32_t* hist;                   // [nlib][BDK_NUM_FLAGS]
    uint32_t* err;                    // err[0]: K1_ERR_* bits, err[1]: largest per-CTA anomalous count seen
};

struct K1Rec4 {
    int32_t pos[4], mpos[4], tid[4], mtid[4], isz[4];
    uint32_t flag[4], mapq[4], rg[4];
};

__device__ __forceinline__ void k1_unpack(const int4 a, const int4 b, const int4 d, const int4 e, const int4 f, const uint2 fl, const uint32_t mq,
                                          const uint2 rg, K1Rec4& r) {
    r.pos[0] = a.x; r.pos[1] = a.y; r.pos[2] = a.z; r.pos[3] = a.w;
    r.mpos[0] = b.x; r.mpos[1] = b.y; r.mpos[2] = b.z; r.mpos[3] = b.w;
    r.tid[0] = d.x; r.tid[1] = d.y; r.tid[2] = d.z; r.tid[3] = d.w;
    r.mtid[0] = e.x; r.mtid[1] = e.y; r.mtid[2] = e.z; r.mtid[3] = e.w;
    r.isz[0] = f.x; r.isz[1] = f.y; r.isz[2] = f.z; r.isz[3] = f.w;
    r.flag[0] = fl.x & 0xffffu; r.flag[1] = fl.x >> 16; r.flag[2] = fl.y & 0xffffu; r.flag[3] = fl.y >> 16;
    r.mapq[0] = mq & 0xffu; r.mapq[1] = (mq >> 8) & 0xffu; r.mapq[2] = (mq >> 16) & 0xffu; r.mapq[3] = mq >> 24;
    r.rg[0] = rg.x & 0xffffu; r.rg[1] = rg.x >> 16; r.rg[2] = rg.y & 0xffffu; r.rg[3] = rg.y >> 16;
}

// 4 consecutive records of this lane straight from global memory, with a ragged end (last, partial stage of a push; whole stages
// come through the ring: k1_load4_stage)
__device__ __forceinline__ void k1_load4_global(const bdk_soa& c, uint64_t g, int nv, uint32_t pad_rg, K1Rec4& r) {
    if (nv == 4) {
        k1_unpack(ld_stream_v4(c.pos + g), ld_stream_v4(c.mpos + g), ld_stream_v4(c.tid + g), ld_stream_v4(c.mtid + g), ld_stream_v4(c.isize + g),
                  ld_stream_v2(c.flag + g), ld_stream_u32(c.mapq + g), ld_stream_v2(c.rgid + g), r);
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
// byte `it` (0..7) of the pair (lo, hi)
__device__ __forceinline__ uint32_t byte_of2(uint32_t lo, uint32_t hi, int it) { return ((it < 4 ? lo : hi) >> (8 * (it & 3))) & 0xffu; }


// ---- mbarrier / TMA bulk copy / named barrier wrappers ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint32_t keep_u32(uint32_t v) { uint32_t r; asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(addr) : "memory");
}
// shared-memory loads through 32-bit addresses computed once per kernel: from generic pointers the compiler rebuilds the shared
// window address (S2R + LEA) at every use inside the stage loop
__device__ __forceinline__ int4 lds128(uint32_t addr) {
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 4 consecutive records of this lane from a ring stage, the stage given by its shared-memory address
__device__ __forceinline__ void k1_load4_stage(uint32_t st, int idx, K1Rec4& r) {
    const int4 a = lds128(st + K1_OFF_POS + idx * 4);
    const int4 b = lds128(st + K1_OFF_MPOS + idx * 4);
    const int4 d = lds128(st + K1_OFF_TID + idx * 4);
    const int4 e = lds128(st + K1_OFF_MTID + idx * 4);
    const int4 f = lds128(st + K1_OFF_ISZ + idx * 4);
    const uint2 fl = lds64(st + K1_OFF_FLAG + idx * 2);
    const uint2 rg = lds64(st + K1_OFF_RG + idx * 2);
    const uint32_t mq = lds32(st + K1_OFF_MAPQ + idx);
    k1_unpack(a, b, d, e, f, fl, mq, rg, r);
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory"); }

// dynamic shared memory of k1_classify_kernel (bytes), host and device agree through this one function
__host__ __device__ inline size_t k1_smem_bytes(int nrg, int nlib, int ncnt, int nkey, bool single_key, bool rg_smem) {
    const size_t ncomp = 1 + (size_t)nkey, ncol = ncnt > 1 ? ncnt : 0;
    size_t b = (size_t)K1_STAGES * K1_STAGE_BYTES;
    b += rg_smem ? (size_t)(nrg + 1) * sizeof(RgDev) : 0;
    b += ((((size_t)nlib * BDK_NUM_FLAGS + 3) & ~size_t(3)) + ncol * K1_CTHREADS) * 4;   // histogram padded to 16 bytes
    b += 2 * ncomp * K1_CWARPS * 8;                       // s_tot [2][ncomp][warps] packed per-stage totals (2 words)
    b += 2 * ncomp * 64 * 2;                              // s_off [2][ncomp][64] u16
    b += 2 * ncomp * 4;                                   // s_base [2][ncomp]
    (void)single_key;
    return b + 128;
}

// MODE K1_FAST: one copy-number key, one source bam, one (library, bam) pair -- the common single-bam run: the
// running counts live in registers and the hot loop has no per-record branches.
// MODE K1_KEYS4: up to four keys, four (library, bam) pairs and four bams (tumor / normal pairs, -a with a few libraries):
// the key of a record is kept as two bit planes next to the decision masks, so the per-key counts are the single-key
// computation repeated on masked decision bits -- thread-local arithmetic and one warp scan per key and tile, no votes and no
// shared-memory atomics in the hot loop; the pass-1 proper-pair counts are four byte counters in one register.
// MODE K1_GENERAL: anything else (up to 64 keys).
// RG_SMEM: the read-group table fits in shared memory.
// PLAIN: a run without -t and without -l (the defaults): the two options are compile-time zeros in the classifier (eight of 66
// instructions per record; the kernel sits on the integer pipe: 0.554 -> 0.516 ms per 100 M records, profiles/k1_classify_r22a.md).
enum { K1_GENERAL = 0, K1_FAST = 1, K1_KEYS4 = 2 };
template <int MODE, bool RG_SMEM, bool PLAIN>
__global__ void __launch_bounds__(K1_THREADS, 2) k1_classify_kernel(const K1Args a) {
    constexpr bool FAST = MODE == K1_FAST, K4 = MODE == K1_KEYS4, SINGLE_KEY = FAST, GEN = MODE == K1_GENERAL;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    const int ncomp = 1 + a.nkey;
    const int nhist = a.nlib * BDK_NUM_FLAGS;
    const int ncol = (a.ncnt > 1 && !K4) ? a.ncnt : 0;                             // one column (or K1_KEYS4): registers do it
    unsigned char* s_ring = s_dyn;                                                 // [K1_STAGES][K1_STAGE_BYTES]
    RgDev* s_rg = reinterpret_cast<RgDev*>(s_ring + K1_STAGES * K1_STAGE_BYTES);   // RG_SMEM: [nrg + 1]
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_rg + (RG_SMEM ? a.nrg + 1 : 0));   // [nlib * 11]
    uint32_t* s_cnt = s_hist + ((nhist + 3) & ~3);                                 // [ncol][K1_CTHREADS], 16-byte aligned
    uint2* s_tot = reinterpret_cast<uint2*>(s_cnt + ncol * K1_CTHREADS);           // [2][ncomp][K1_CWARPS] per-stage totals, a byte each
    uint16_t* s_off = reinterpret_cast<uint16_t*>(s_tot + 2 * ncomp * K1_CWARPS);  // [2][ncomp][64] exclusive offset of (stage, warp) in the tile
    uint32_t* s_base = reinterpret_cast<uint32_t*>(s_off + 2 * ncomp * 64);        // [2][ncomp] CTA-relative exclusive prefix of the tile
    __shared__ __align__(8) uint64_t s_full[K1_STAGES], s_empty[K1_STAGES];
    __shared__ unsigned long long s_wb[2][K1_CWARPS];

    const unsigned FULL = 0xffffffffu;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < ((nhist + 3) & ~3) + ncol * K1_CTHREADS; i += K1_THREADS) s_hist[i] = 0;
    if (RG_SMEM) for (int i = threadIdx.x; i < a.nrg + 1; i += K1_THREADS) s_rg[i] = a.rgtab[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < K1_STAGES; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], K1_CWARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t ntiles = (uint32_t)div_up<uint64_t>(a.n, K1_TILE);
    const uint32_t tile0 = min(ntiles, blockIdx.x * a.tiles_per_cta), tile1 = min(ntiles, tile0 + a.tiles_per_cta);

    // =============================== producer warp ===================================================
    if (warp == K1_CWARPS) {
        if (lane != 0) return;
        uint32_t g = 0;                                       // ring uses so far
        for (uint32_t tile = tile0; tile < tile1; ++tile) {
            for (int s = 0; s < K1_SUBS; ++s) {
                const uint64_t rec0 = (uint64_t)tile * K1_TILE + (uint64_t)s * K1_SUB;
                if (rec0 + K1_SUB > a.n) break;               // ragged stage: the consumers read it straight from global memory
                const int stage = g % K1_STAGES;
                const uint32_t use = g / K1_STAGES;
                if (use) mbar_wait(&s_empty[stage], (use - 1) & 1);
                unsigned char* st = s_ring + stage * K1_STAGE_BYTES;
                mbar_arrive_expect_tx(&s_full[stage], K1_STAGE_BYTES);
                tma_load_1d(st + K1_OFF_POS, a.c.pos + rec0, 4 * K1_SUB, &s_full[stage]);
                tma_load_1d(st + K1_OFF_MPOS, a.c.mpos + rec0, 4 * K1_SUB, &s_full[stage]);
                tma_load_1d(st + K1_OFF_TID, a.c.tid + rec0, 4 * K1_SUB, &s_full[stage]);
                tma_load_1d(st + K1_OFF_MTID, a.c.mtid + rec0, 4 * K1_SUB, &s_full[stage]);
                tma_load_1d(st + K1_OFF_ISZ, a.c.isize + rec0, 4 * K1_SUB, &s_full[stage]);
                tma_load_1d(st + K1_OFF_FLAG, a.c.flag + rec0, 2 * K1_SUB, &s_full[stage]);
                tma_load_1d(st + K1_OFF_RG, a.c.rgid + rec0, 2 * K1_SUB, &s_full[stage]);
                tma_load_1d(st + K1_OFF_MAPQ, a.c.mapq + rec0, K1_SUB, &s_full[stage]);
                ++g;
            }
        }
        return;
    }

    // =============================== scan warp =======================================================
    if (warp == K1_CWARPS + 1) {
        uint32_t run[3] = {0, 0, 0};                                         // lane owns components lane, lane + 32, lane + 64
        for (uint32_t tile = tile0; tile < tile1; ++tile) {
            const int tb = (tile - tile0) & 1;
            named_bar_sync(K1_BAR_A + tb, K1_CTHREADS + 32);                 // the consumers' totals of this tile are in s_tot
            if (a.nbam > 1 && lane == 0) {
                unsigned long long b = 0;
                for (int w = 0; w < K1_CWARPS; ++w) b |= s_wb[tb][w];
                a.tile_bams[tile] = b;
            }
            // per component: exclusive offsets of the 64 (stage, warp) cells in stream order, and the tile total
            for (int comp = 0; comp < ncomp; ++comp) {
                const uint2* tot = s_tot + ((size_t)tb * ncomp + comp) * K1_CWARPS;
                // cell q = stage * 8 + warp; this lane scans cells 2 * lane and 2 * lane + 1
                const int q0 = 2 * lane, st0 = q0 >> 3, w0 = q0 & 7;
                const uint2 t0 = tot[w0], t1 = tot[w0 + 1];
                const uint32_t v0 = byte_of2(t0.x, t0.y, st0), v1 = byte_of2(t1.x, t1.y, st0);
                const uint32_t inc = warp_incl_scan(v0 + v1);
                uint16_t* off = s_off + ((size_t)tb * ncomp + comp) * 64;
                off[q0] = (uint16_t)(inc - v0 - v1); off[q0 + 1] = (uint16_t)(inc - v1);
                const uint32_t total = __shfl_sync(FULL, inc, 31);
                if ((comp & 31) == lane) {
                    uint32_t& r = comp < 32 ? run[0] : (comp < 64 ? run[1] : run[2]);
                    s_base[tb * ncomp + comp] = r;
                    r += total;
                }
            }
            __syncwarp();
            named_bar_arrive(K1_BAR_B + tb, K1_CTHREADS + 32);               // s_off / s_base of this tile are ready
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int comp = lane + 32 * q;
            if (comp < ncomp) a.seg_cnt[(size_t)blockIdx.x * ncomp + comp] = run[q];
        }
        if (lane == 0) {
            atomicMax(a.err + 1, run[0]);
            if (run[0] > a.seg_cap) atomicOr(a.err, K1_ERR_OVERFLOW);
        }
        return;
    }

    // =============================== consumer warps ==================================================
    ClassifyOpts co = a.co;
    if (PLAIN) { co.transchr = 0; co.long_insert = 0; }
    uint32_t* my_cnt = s_cnt + threadIdx.x;                                        // this thread's private counter column
    uint32_t spcnt = 0;                                                            // ncnt == 1: pass-1 proper pairs seen by this thread
    uint32_t bad = 0;                                                              // OR of the info words (bit 31: invalid read group)
    K1Stash* my_stash = a.stash + ((size_t)blockIdx.x * 2 * K1_CTHREADS + threadIdx.x) * K1_STASH;   // + tb * K1_CTHREADS * K1_STASH
    uint32_t cstage = 0, cphase = 0;                                               // ring position and its parity (same sequence as the producer)
    // (through an opaque move: left to itself the compiler rebuilds these from S2R SR_CgaCtaId + LEA at every use)
    const uint32_t ring_addr = keep_u32(smem_u32(s_ring)), full_addr = keep_u32(smem_u32(s_full)), empty_addr = keep_u32(smem_u32(s_empty));
    const uint32_t rg_addr = RG_SMEM ? keep_u32(smem_u32(s_rg)) : 0u;
    // state of the previous tile, whose anomalous reads are written out after the current tile is classified
    bool prev_have = false;
    bdk_aread* seg_ar = a.seg_ar + (size_t)blockIdx.x * a.seg_cap;
    uint32_t* seg_P = a.seg_P + (size_t)blockIdx.x * a.seg_cap * a.nkey;
    uint32_t prev_tile = 0, prev_amask = 0, prev_pmask = 0, prev_aex0 = 0, prev_aex1 = 0, prev_pex0 = 0, prev_pex1 = 0;
    uint32_t prev_keys[GEN ? K1_SUBS : 1];
#pragma unroll
    for (int s = 0; s < (GEN ? K1_SUBS : 1); ++s) prev_keys[s] = 0;
    // K1_KEYS4: bit planes of the key per record, exclusive per-stage prefixes per key (a byte per stage), pass-1 counts per column
    uint32_t prev_kb0 = 0, prev_kb1 = 0, prev_kpex[K4 ? 4 : 1][2], spc[K4 ? 4 : 1];
#pragma unroll
    for (int k = 0; k < (K4 ? 4 : 1); ++k) { prev_kpex[k][0] = 0; prev_kpex[k][1] = 0; spc[k] = 0; }

    for (uint32_t tile = tile0;; ++tile) {
        const int tb = (tile - tile0) & 1;
        const bool have = tile < tile1;
        uint32_t amask = 0, pmask = 0, aex0 = 0, aex1 = 0, pex0 = 0, pex1 = 0, keys[GEN ? K1_SUBS : 1];
#pragma unroll
        for (int s = 0; s < (GEN ? K1_SUBS : 1); ++s) keys[s] = 0;
        uint32_t kb0 = 0, kb1 = 0, kpex[K4 ? 4 : 1][2], spw = 0, bams32 = 0;
        if (have) {
            // ---------------- phase 1: four decisions per record, kept as bit masks ----------------------
            unsigned long long bams = 0;
            uint32_t nst = 0;                                                  // stash slots used by this thread
            K1Stash* stash = my_stash + (size_t)tb * K1_CTHREADS * K1_STASH;
#pragma unroll 1      // (unrolling the stage loop of the multi-key variant made 11 K instructions: instruction-cache misses were its top stall)
            // stages of this tile that are whole (they come through the ring) and stages that hold records at all: one 64-bit
            // look at the end of the records per tile instead of two per stage
            const uint64_t tile_rec0 = (uint64_t)tile * K1_TILE;
            const uint64_t tile_left = a.n > tile_rec0 ? a.n - tile_rec0 : 0;
            const int s_whole = (int)(tile_left / K1_SUB < (uint64_t)K1_SUBS ? tile_left / K1_SUB : (uint64_t)K1_SUBS);
            const int s_any = (int)((tile_left + K1_SUB - 1) / K1_SUB < (uint64_t)K1_SUBS ? (tile_left + K1_SUB - 1) / K1_SUB : (uint64_t)K1_SUBS);
            const int idx = (warp << 7) + (lane << 2);                         // first of this lane's 4 records inside a stage
            for (int s = 0; s < s_any; ++s) {
                K1Rec4 r;
                const bool staged = s < s_whole;
                const uint32_t stage = cstage;
                if (staged) {
                    mbar_wait_addr(full_addr + stage * 8, cphase);
                    k1_load4_stage(ring_addr + stage * K1_STAGE_BYTES, idx, r);
                } else {
                    const uint64_t rec0 = tile_rec0 + (uint64_t)s * K1_SUB;
                    const uint64_t gi = rec0 + idx;
                    const int nv = gi + 4 <= a.n ? 4 : (gi < a.n ? (int)(a.n - gi) : 0);
                    k1_load4_global(a.c, gi, nv, (uint32_t)a.pad_rg, r);
                }
                uint32_t kk = 0, a4 = 0, p4 = 0, k04 = 0, k14 = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t ri = min(r.rg[j], (uint32_t)a.nrg);
                    RgDev L;
                    if (RG_SMEM) { const int4 q = lds128(rg_addr + ri * 16); L.upper = __int_as_float(q.x); L.lower = __int_as_float(q.y); L.min_mapq = q.z; L.info = (uint32_t)q.w; }
                    else { const int4 q = __ldg(reinterpret_cast<const int4*>(a.rgtab) + ri); L.upper = __int_as_float(q.x); L.lower = __int_as_float(q.y); L.min_mapq = q.z; L.info = (uint32_t)q.w; }
                    const uint32_t ch = classify_hot(r.pos[j], r.mpos[j], r.tid[j], r.mtid[j], r.isz[j], r.flag[j], r.mapq[j],
                                                     L.upper, L.lower, L.min_mapq, co);
                    bad |= L.info;
                    if (ch & CH_ANOM) a4 |= 1u << j;
                    if (ch & CH_MPROPER) p4 |= 1u << j;
                    if (GEN) kk |= ((L.info >> RGI_KEY_SHIFT) & 0x3fu) << (8 * j);
                    if (GEN && a.nbam > 1) bams |= 1ull << ((L.info >> RGI_BAM_SHIFT) & 0x3fu);
                    // pass-1 proper-pair count per (library, bam)
                    const uint32_t sp = (ch >> 3) & 1u;          // CH_SPROPER
                    if (K4) {
                        k04 |= ((L.info >> RGI_KEY_SHIFT) & 1u) << j; k14 |= ((L.info >> (RGI_KEY_SHIFT + 1)) & 1u) << j;
                        if (a.cn_lib) bams32 |= 1u << ((L.info >> RGI_BAM_SHIFT) & 3u);      // without -a the key IS the bam: taken from the key planes below
                        spw += sp << (((L.info >> RGI_CNT_SHIFT) & 3u) * 8);          // at most 32 records per thread and tile: a byte per column
                    } else if (FAST || a.ncnt == 1) spcnt += sp;
                    else if (a.ncnt) atomicAdd(my_cnt + ((L.info >> RGI_CNT_SHIFT) & 31u) * K1_CTHREADS, sp);   // private column: no conflicts
                    else {                                       // many (library, bam) pairs: one atomic per distinct read group and warp
                        const unsigned spm = __ballot_sync(FULL, sp && r.rg[j] < (uint32_t)a.nrg);
                        if ((spm >> lane) & 1u) {
                            const unsigned peers = __match_any_sync(spm, r.rg[j]);
                            if (lane == __ffs(peers) - 1) atomicAdd(&a.rg_sproper[r.rg[j]], (unsigned long long)__popc(peers));
                        }
                    }
                    if (ch & CH_HIST) {                          // ~1-3 % of the records: full classification, histogram, stash
                        const uint32_t cr = classify_record(r.pos[j], r.mpos[j], r.tid[j], r.mtid[j], r.isz[j], r.flag[j], r.mapq[j],
                                                            L.upper, L.lower, L.min_mapq, co);
                        const uint32_t hf = (cr >> CR_HIST_SHIFT) & 0xFu;
                        if (hf) atomicAdd(&s_hist[(L.info & RGI_LIB_MASK) * BDK_NUM_FLAGS + hf], 1u);
                        if (cr & CR_ANOM) {
                            K1Stash e;
                            e.pos = r.pos[j]; e.tid = r.tid[j]; e.abs_isize = r.isz[j] < 0 ? -r.isz[j] : r.isz[j];
                            e.meta = make_meta(cr, (int)(L.info & RGI_LIB_MASK), r.mapq[j]);
                            *reinterpret_cast<int4*>(stash + nst) = *reinterpret_cast<const int4*>(&e);
                            ++nst;
                        }
                    }
                }
                if (staged) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_addr(empty_addr + stage * 8);
                    if (++cstage == K1_STAGES) { cstage = 0; cphase ^= 1u; }
                }
                amask |= a4 << (4 * s); pmask |= p4 << (4 * s);
                if (GEN) keys[s] = kk;
                if (K4) { kb0 |= k04 << (4 * s); kb1 |= k14 << (4 * s); }
            }
            // ---------------- warp level: ranks inside the warp's 8 x 128 records ------------------------
            {
                const uint32_t c0 = nibble_counts(amask & 0xffffu), c1 = nibble_counts(amask >> 16);
                const uint32_t i0 = warp_incl_scan(c0), i1 = warp_incl_scan(c1);
                aex0 = i0 - c0; aex1 = i1 - c1;
                if (lane == 31) s_tot[((size_t)tb * ncomp + 0) * K1_CWARPS + warp] = make_uint2(i0, i1);
            }
            if (SINGLE_KEY) {
                const uint32_t c0 = nibble_counts(pmask & 0xffffu), c1 = nibble_counts(pmask >> 16);
                const uint32_t i0 = warp_incl_scan(c0), i1 = warp_incl_scan(c1);
                pex0 = i0 - c0; pex1 = i1 - c1;
                if (lane == 31) s_tot[((size_t)tb * ncomp + 1) * K1_CWARPS + warp] = make_uint2(i0, i1);
            } else if (K4) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k >= a.nkey) break;
                    const uint32_t pm = pmask & ((k & 1) ? kb0 : ~kb0) & ((k & 2) ? kb1 : ~kb1);      // kept proper pairs of key k
                    const uint32_t c0 = nibble_counts(pm & 0xffffu), c1 = nibble_counts(pm >> 16);
                    const uint32_t i0 = warp_incl_scan(c0), i1 = warp_incl_scan(c1);
                    kpex[k][0] = i0 - c0; kpex[k][1] = i1 - c1;
                    if (lane == 31) s_tot[((size_t)tb * ncomp + 1 + k) * K1_CWARPS + warp] = make_uint2(i0, i1);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) spc[c] += (spw >> (8 * c)) & 0xffu;
                if (a.nbam > 1) {
                    if (!a.cn_lib) {
                        // records beyond the end of the push are padding with key bits of pad_rg: its bam is present anyway or the
                        // span search skips nothing it should not (a set bit only costs a look)
#pragma unroll
                        for (int k = 0; k < 4; ++k) if ((((k & 1) ? kb0 : ~kb0) & ((k & 2) ? kb1 : ~kb1)) != 0u) bams32 |= 1u << k;
                    }
                    bams32 = __reduce_or_sync(FULL, bams32);
                    if (lane == 0) s_wb[tb][warp] = bams32;
                }
            } else {
                unsigned long long present = 0;
#pragma unroll
                for (int s = 0; s < K1_SUBS; ++s)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if ((pmask >> (s * 4 + j)) & 1u) present |= 1ull << ((keys[s] >> (8 * j)) & 0x3fu);
                present = (unsigned long long)__reduce_or_sync(FULL, (uint32_t)present) | ((unsigned long long)__reduce_or_sync(FULL, (uint32_t)(present >> 32)) << 32);
                for (int k = 0; k < a.nkey; ++k) {                            // per key: per-stage totals of the warp
                    uint32_t c0 = 0, c1 = 0;
                    if (!((present >> k) & 1ull)) {
                        if (lane == 0) s_tot[((size_t)tb * ncomp + 1 + k) * K1_CWARPS + warp] = make_uint2(0u, 0u);
                        continue;
                    }
#pragma unroll
                    for (int s = 0; s < K1_SUBS; ++s)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (((pmask >> (s * 4 + j)) & 1u) && ((keys[s] >> (8 * j)) & 0x3fu) == (uint32_t)k) { if (s < 4) c0 += 1u << (8 * s); else c1 += 1u << (8 * (s - 4)); }
                    c0 = __reduce_add_sync(FULL, c0); c1 = __reduce_add_sync(FULL, c1);
                    if (lane == 0) s_tot[((size_t)tb * ncomp + 1 + k) * K1_CWARPS + warp] = make_uint2(c0, c1);
                }
            }
            if (GEN && a.nbam > 1) {
                bams = (unsigned long long)__reduce_or_sync(FULL, (uint32_t)bams) | ((unsigned long long)__reduce_or_sync(FULL, (uint32_t)(bams >> 32)) << 32);
                if (lane == 0) s_wb[tb][warp] = bams;
            }
            __syncwarp();
            named_bar_arrive(K1_BAR_A + tb, K1_CTHREADS + 32);               // hand the totals to the scan warp, do not wait
        }
        // ---------------- phase 2 of the PREVIOUS tile: write its anomalous reads -------------------------
        if (prev_have) {
            const int pb = tb ^ 1;
            named_bar_sync(K1_BAR_B + pb, K1_CTHREADS + 32);                 // its look-back is done (it had a whole tile of time)
            if (prev_amask) {
                const uint32_t base_a = s_base[pb * ncomp];
                const uint16_t* off_a = s_off + ((size_t)pb * ncomp) * 64;
                const K1Stash* stash = my_stash + (size_t)pb * K1_CTHREADS * K1_STASH;
                uint32_t m = prev_amask, k = 0;
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    const int s = bit >> 2, j = bit & 3;
                    const uint32_t a4 = (prev_amask >> (4 * s)) & 0xFu, p4 = (prev_pmask >> (4 * s)) & 0xFu;
                    const uint32_t below = (2u << j) - 1u;                   // items 0..j of the stage
                    const uint32_t o = base_a + off_a[s * 8 + warp] + byte_of2(prev_aex0, prev_aex1, s) + __popc(a4 & (below >> 1));
                    const K1Stash e = stash[k++];
                    if (o >= a.seg_cap) continue;
                    const uint64_t idx = (uint64_t)prev_tile * K1_TILE + (uint64_t)s * K1_SUB + (warp << 7) + (lane << 2) + j;
                    bdk_aread rec;
                    rec.pos = e.pos; rec.tid = e.tid; rec.qlen = a.c.qlen[idx];
                    rec.abs_isize = e.abs_isize; rec.meta = e.meta;
                    rec.record = a.base_index + (uint32_t)idx;
                    rec.qid = a.c.qid[idx];
                    int4* dst = reinterpret_cast<int4*>(seg_ar + o);
                    const int4* src = reinterpret_cast<const int4*>(&rec);
                    dst[0] = src[0]; dst[1] = src[1];
                    if (SINGLE_KEY) {
                        seg_P[o] = s_base[pb * ncomp + 1] + s_off[((size_t)pb * ncomp + 1) * 64 + s * 8 + warp] + byte_of2(prev_pex0, prev_pex1, s) + __popc(p4 & below);
                    }
                    if (K4) {
                        const uint32_t k04 = (prev_kb0 >> (4 * s)) & 0xFu, k14 = (prev_kb1 >> (4 * s)) & 0xFu;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k >= a.nkey) break;
                            const uint32_t pk = p4 & ((k & 1) ? k04 : ~k04) & ((k & 2) ? k14 : ~k14);
                            seg_P[(size_t)o * a.nkey + k] = s_base[pb * ncomp + 1 + k] + s_off[((size_t)pb * ncomp + 1 + k) * 64 + s * 8 + warp]
                                                            + byte_of2(prev_kpex[k][0], prev_kpex[k][1], s) + __popc(pk & below);
                        }
                    }
                }
            }
            if (GEN) {
                // per key: proper pairs of lower lanes in the same stage come from warp votes (all lanes take part)
                const uint32_t base_a = s_base[pb * ncomp];
                const uint16_t* off_a = s_off + ((size_t)pb * ncomp) * 64;
                const unsigned lt = lanemask_lt();
#pragma unroll 1
                for (int s = 0; s < K1_SUBS; ++s) {
                    const uint32_t a4 = (prev_amask >> (4 * s)) & 0xFu, p4 = (prev_pmask >> (4 * s)) & 0xFu;
                    if (!__any_sync(FULL, a4 != 0)) continue;
                    unsigned long long present = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) if ((p4 >> j) & 1u) present |= 1ull << ((prev_keys[s] >> (8 * j)) & 0x3fu);
                    present = (unsigned long long)__reduce_or_sync(FULL, (uint32_t)present) | ((unsigned long long)__reduce_or_sync(FULL, (uint32_t)(present >> 32)) << 32);
                    const uint32_t o0 = base_a + off_a[s * 8 + warp] + byte_of2(prev_aex0, prev_aex1, s);
                    for (int k = 0; k < a.nkey; ++k) {
                        uint32_t before = 0, mine = 0;
                        if ((present >> k) & 1ull) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const bool mk = ((p4 >> j) & 1u) && ((prev_keys[s] >> (8 * j)) & 0x3fu) == (uint32_t)k;
                                before += __popc(__ballot_sync(FULL, mk) & lt);
                                mine |= (uint32_t)mk << j;
                            }
                        }
                        if (a4) {
                            const uint32_t v0 = s_base[pb * ncomp + 1 + k] + s_off[((size_t)pb * ncomp + 1 + k) * 64 + s * 8 + warp] + before;
                            uint32_t rank = 0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if ((a4 >> j) & 1u) {
                                    const uint32_t o = o0 + rank++;
                                    if (o < a.seg_cap) seg_P[(size_t)o * a.nkey + k] = v0 + __popc(mine & ((2u << j) - 1u));
                                }
                            }
                        }
                    }
                }
            }
        }
        if (!have) break;
        prev_have = true; prev_tile = tile; prev_amask = amask; prev_pmask = pmask;
        prev_aex0 = aex0; prev_aex1 = aex1; prev_pex0 = pex0; prev_pex1 = pex1;
        if (GEN) {
#pragma unroll
            for (int s = 0; s < K1_SUBS; ++s) prev_keys[s] = keys[s];
        }
        if (K4) {
            prev_kb0 = kb0; prev_kb1 = kb1;
#pragma unroll
            for (int k = 0; k < 4; ++k) { prev_kpex[k][0] = kpex[k][0]; prev_kpex[k][1] = kpex[k][1]; }
        }
    }
    // ---------------- epilogue of the consumers: flush the accumulators --------------------------------------
    if (__any_sync(FULL, (bad & RGI_INVALID) != 0) && lane == 0) atomicOr(a.err, K1_ERR_RG);
    if (K4) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c >= a.ncnt) break;
            const uint32_t t = __reduce_add_sync(FULL, spc[c]);
            if (lane == 0 && t) atomicAdd(a.rg_sproper + a.cnt_rg[c], (unsigned long long)t);
        }
    } else if (FAST || a.ncnt == 1) {
        spcnt = __reduce_add_sync(FULL, spcnt);
        if (lane == 0 && spcnt) atomicAdd(a.rg_sproper + a.cnt_rg[0], (unsigned long long)spcnt);
    }
    named_bar_sync(K1_BAR_C, K1_CTHREADS);
    for (int i = threadIdx.x; i < nhist; i += K1_CTHREADS) if (s_hist[i]) atomicAdd(a.hist + i, s_hist[i]);
    for (int col = warp; col < ncol; col += K1_CWARPS) {
        uint32_t sum = 0;
        for (int t = lane; t < K1_CTHREADS; t += 32) sum += s_cnt[col * K1_CTHREADS + t];
        sum = __reduce_add_sync(FULL, sum);
        if (lane == 0 && sum) atomicAdd(a.rg_sproper + a.cnt_rg[col], (unsigned long long)sum);
    }
}

// ---- wire format -> hot columns -------------------------------------------------------------------------
// A chunk of a packed run (include/bdk.h: bdk_packed, 12 bytes per record) back into the eight 25-byte columns K1 reads;
// the records that did not fit the packed form are then overwritten from the exception arrays. HBM traffic 12 + 25 bytes per
// record, hidden behind the host-to-device copy of the next chunk.
__global__ void __launch_bounds__(256) k1_expand_kernel(const int32_t* __restrict__ ppos, const uint32_t* __restrict__ meta, const uint32_t* __restrict__ rel,
        uint64_t n, int32_t tid, uint16_t pad_rg, int32_t* __restrict__ pos, int32_t* __restrict__ mpos, int32_t* __restrict__ otid, int32_t* __restrict__ mtid,
        int32_t* __restrict__ isize, uint16_t* __restrict__ flag, uint8_t* __restrict__ mapq, uint16_t* __restrict__ rgid) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int32_t p = ppos[i];
        const uint32_t m = meta[i], r = rel[i];
        const uint32_t rg = m >> 20;
        pos[i] = p; mpos[i] = p + (int32_t)(int16_t)(r >> 16); otid[i] = tid; mtid[i] = tid;
        isize[i] = (int32_t)(int16_t)(r & 0xffffu);
        flag[i] = (uint16_t)(m & 0xfffu); mapq[i] = (uint8_t)((m >> 12) & 0xffu);
        rgid[i] = rg == BDK_PACKED_EXCEPT ? pad_rg : (uint16_t)rg;        // (overwritten by k1_expand_exceptions_kernel)
    }
}
__global__ void __launch_bounds__(256) k1_expand_exceptions_kernel(const uint32_t* __restrict__ x_index, const int32_t* __restrict__ x_mpos,
        const int32_t* __restrict__ x_mtid, const int32_t* __restrict__ x_isize, const uint16_t* __restrict__ x_flag, const uint16_t* __restrict__ x_rgid,
        uint64_t nx, uint32_t chunk_first, int32_t* __restrict__ mpos, int32_t* __restrict__ mtid, int32_t* __restrict__ isize, uint16_t* __restrict__ flag,
        uint16_t* __restrict__ rgid) {
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < nx; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t i = x_index[k] - chunk_first;
        mpos[i] = x_mpos[k]; mtid[i] = x_mtid[k]; isize[i] = x_isize[k]; flag[i] = x_flag[k]; rgid[i] = x_rgid[k];
    }
}

// ---- move the per-CTA segments to their final place ---------------------------------------------------
// Block b copies segment b to ar[carry[0] + sum of the anomalous counts of segments < b ...) and adds the
// matching kept-proper-pair prefixes to its P rows. Launched with the same grid as K1. The last block leaves
// the totals after this push in carry_out.
__global__ void __launch_bounds__(256) k1_compact_kernel(const bdk_aread* __restrict__ seg_ar, const uint32_t* __restrict__ seg_P, uint32_t seg_cap,
        const uint32_t* __restrict__ seg_cnt, int nkey, const uint32_t* __restrict__ carry, uint32_t* __restrict__ carry_out,
        bdk_aread* __restrict__ ar, uint32_t* __restrict__ P, uint32_t cap, uint32_t* __restrict__ err) {
    __shared__ uint32_t s_base[K1_MAXK + 1];
    __shared__ uint32_t s_red[8];
    const int ncomp = 1 + nkey, b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int comp = 0; comp < ncomp; ++comp) {
        uint32_t s = 0;
        for (int i = threadIdx.x; i < b; i += 256) s += seg_cnt[(size_t)i * ncomp + comp];
        s = __reduce_add_sync(0xffffffffu, s);
        if (lane == 0) s_red[warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = carry[comp];
            for (int w = 0; w < 8; ++w) t += s_red[w];
            s_base[comp] = t;
            if (b == (int)gridDim.x - 1) carry_out[comp] = t + seg_cnt[(size_t)b * ncomp + comp];
        }
        __syncthreads();
    }
    const uint32_t cnt = min(seg_cnt[(size_t)b * ncomp], seg_cap), base = s_base[0];
    if (threadIdx.x == 0 && base + seg_cnt[(size_t)b * ncomp] > cap) atomicOr(err, K1_ERR_OVERFLOW);
    const int4* src = reinterpret_cast<const int4*>(seg_ar + (size_t)b * seg_cap);
    int4* dst = reinterpret_cast<int4*>(ar + base);
    for (uint32_t i = threadIdx.x; i < 2 * cnt; i += 256) if (base + (i >> 1) < cap) dst[i] = src[i];
    const uint32_t* sp = seg_P + (size_t)b * seg_cap * nkey;
    for (uint32_t i = threadIdx.x; i < cnt * (uint32_t)nkey; i += 256) {
        const uint32_t d = i / nkey, k = i - d * nkey;
        if (base + d < cap) P[(size_t)(base + d) * nkey + k] = sp[i] + s_base[1 + k];
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
