// bdk_logic.h -- the per-record and per-connection logic of the hot path, written once as
// __host__ __device__ inline functions.  The CUDA kernels (k1..k4 in this directory) call these
// from device code; tests/hostsim compiles the very same functions for the host so the logic
// (not the parallel mechanics) can be checked against the oracle without a GPU.
//
// Reference code restated here (paths relative to the reference root, commit 4e44b43):
//   classify_record    IlluminaPEReadClassifier.cpp:13-101, Alignment.hpp:72-159,
//                      BamSummary.cpp:70-113 (pass 1), BreakDancer.cpp:150-207 (pass 2)
//   poisson_log_sf     boost poisson complement cdf as used by ComputeProbScore, BreakDancer.cpp:62-68
//   k4n_* / k4_*       build_connection / process_sv / SvBuilder / is_region_final / clear_region,
//                      BreakDancer.cpp:266-512, SvBuilder.cpp:18-118, ReadRegionData.cpp:70-175
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/bdk.h"

#if defined(__CUDACC__)
#define BDK_HD __host__ __device__ __forceinline__
#else
#define BDK_HD inline
#endif

namespace bdk {

// ---- IEEE fp32 without contraction (the reference is compiled for x86-64 without FMA) -------
BDK_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
BDK_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
BDK_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
BDK_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}

// ---- per-library constants as the kernels see them -------------------------------------------
struct alignas(16) LibDev {   // one 16-byte load per record
    float upper, lower;
    int32_t min_mapq;   // effective: library override or -q
    int32_t key;        // copy-number key: library index (-a) or the library's bam index
};

// ---- classification --------------------------------------------------------------------------
// Result bits of classify_record()
enum : uint32_t {
    CR_FLAG_MASK = 0xF,        // pass-2 ReadFlag (after -l re-flag and RR->FF), valid if CR_KEPT
    CR_KEPT = 1u << 4,         // survives push_read's filters (BreakDancer.cpp:159-167)
    CR_MPROPER = 1u << 5,      // kept && proper_pair(): counts into nread_ROI/nread_FR (:172-175)
    CR_ANOM = 1u << 6,         // kept && not NORMAL_FR/NORMAL_RF: enters the region builder
    CR_SPROPER = 1u << 7,      // pass 1: bdqual > min_mapq && proper_pair() (BamSummary.cpp:80-87)
    CR_HIST_SHIFT = 8,         // bits 8-11: pass-1 flag to count in read_counts_by_flag, 0 = none
    CR_REV = 1u << 12          // reverse strand (ori() == REV)
};

BDK_HD int long_insert_reflag(int flag, bool gt_upper, bool lt_upper, bool lt_lower) {
    if (gt_upper && flag == BDK_NORMAL_RF) flag = BDK_ARP_RF;
    if (lt_upper && flag == BDK_ARP_RF) flag = BDK_NORMAL_RF;
    if (lt_lower && flag == BDK_NORMAL_RF) flag = BDK_ARP_SMALL_INSERT;
    return flag;
}

struct ClassifyOpts {
    int32_t max_sd;
    int32_t transchr;
    int32_t long_insert;
};

// Branch-free restatement of IlluminaPEReadClassifier::classify + the two filter chains.
// NA (dup || !paired), UNMAPPED and MATE_UNMAPPED reads are dropped by both passes
// (BamSummary.cpp:89-94, BreakDancer.cpp:159), so only the mapped-pair classes are materialised.
BDK_HD uint32_t classify_record(int32_t pos, int32_t mpos, int32_t tid, int32_t mtid, int32_t isize, uint32_t flag,
                                uint32_t bdqual, float upper, float lower, int32_t min_mapq, const ClassifyOpts& o) {
    const int32_t a = isize < 0 ? -isize : isize;   // abs(core.isize)
    const float af = (float)a;                      // the reference compares int against float cut-offs
    const bool gt_upper = af > upper, lt_lower = af < lower;
    const bool inter = tid != mtid;
    const bool rr = (flag & 0x10u) != 0, mr = (flag & 0x20u) != 0;
    int cls = gt_upper ? BDK_ARP_LARGE_INSERT : (lt_lower ? BDK_ARP_SMALL_INSERT : BDK_NORMAL_FR);
    cls = ((pos < mpos) == rr) ? BDK_ARP_RF : cls;
    cls = (rr == mr) ? (rr ? BDK_ARP_RR : BDK_ARP_FF) : cls;
    cls = inter ? BDK_ARP_CTX : cls;
    const bool proper = (flag & 0x40Fu) == 0x3u;                 // paired, proper, both mapped, not dup
    const bool base_ok = (flag & 0x40Du) == 0x1u                 // paired, not dup, neither mate unmapped
                         && !(o.transchr && !inter);
    const bool mapq_ok = (int32_t)bdqual > min_mapq;
    int cls2 = cls;
    if (o.long_insert) cls2 = long_insert_reflag(cls, gt_upper, af < upper, lt_lower);
    const bool normal2 = cls2 == BDK_NORMAL_FR || cls2 == BDK_NORMAL_RF;
    const int cls3 = cls2 == BDK_ARP_RR ? BDK_ARP_FF : cls2;
    const bool kept = mapq_ok && base_ok && !(cls != BDK_ARP_CTX && a > o.max_sd);
    uint32_t r = rr ? CR_REV : 0u;
    r |= (mapq_ok && proper) ? CR_SPROPER : 0u;
    r |= (mapq_ok && base_ok && !normal2) ? (uint32_t)cls2 << CR_HIST_SHIFT : 0u;
    r |= kept ? (CR_KEPT | (uint32_t)cls3) : 0u;
    r |= (kept && proper) ? CR_MPROPER : 0u;
    r |= (kept && !normal2) ? CR_ANOM : 0u;
    return r;
}
BDK_HD uint32_t classify_record(int32_t pos, int32_t mpos, int32_t tid, int32_t mtid, int32_t isize, uint32_t flag,
                                uint32_t bdqual, const LibDev& L, const ClassifyOpts& o) {
    return classify_record(pos, mpos, tid, mtid, isize, flag, bdqual, L.upper, L.lower, L.min_mapq, o);
}

// The four decisions the streaming pass needs from every record, without materialising the class:
//   CH_ANOM = CR_ANOM, CH_MPROPER = CR_MPROPER, CH_HIST = (pass-1 histogram nibble != 0), CH_SPROPER = CR_SPROPER.
// Records with CH_HIST set (a superset of CH_ANOM, ~1-3 % of the stream) are classified in full
// by classify_record() afterwards. Equivalence is checked exhaustively in tests (hostsim_classify_hot_check).
enum : uint32_t { CH_ANOM = 1u, CH_MPROPER = 2u, CH_HIST = 4u, CH_SPROPER = 8u };
BDK_HD uint32_t classify_hot(int32_t pos, int32_t mpos, int32_t tid, int32_t mtid, int32_t isize, uint32_t flag,
                             uint32_t bdqual, float upper, float lower, int32_t min_mapq, const ClassifyOpts& o) {
    const int32_t a = isize < 0 ? -isize : isize;
    const float af = (float)a;
    const bool inter = tid != mtid;
    const bool rr = (flag & 0x10u) != 0;
    const bool opposite = ((flag ^ (flag >> 1)) & 0x10u) != 0;              // read and mate on different strands (bits 4 and 5 differ)
    const bool rf = (pos < mpos) == rr;                                      // reverse read leftmost
    const bool lt_lower = af < lower;
    const bool fr_pair = !inter && opposite;
    bool normal2 = fr_pair && !rf && !(af > upper) && !lt_lower;             // NORMAL_FR
    if (o.long_insert) normal2 = normal2 || (fr_pair && rf && af < upper && !lt_lower);   // ARP_RF re-flagged NORMAL_RF
    const bool mapped_pair = (flag & 0x40Du) == 0x1u;                        // paired, read and mate mapped, not a duplicate
    const bool proper = mapped_pair && (flag & 0x2u) != 0;                   // (flag & 0x40F) == 0x3
    const bool base_ok = mapped_pair && !(o.transchr && !inter);
    const bool mapq_ok = (int32_t)bdqual > min_mapq;
    const bool pre = mapq_ok && base_ok;
    const bool kept = pre && (inter || a <= o.max_sd);
    return (kept && !normal2 ? CH_ANOM : 0u) | (kept && proper ? CH_MPROPER : 0u) | (pre && !normal2 ? CH_HIST : 0u) |
           (mapq_ok && proper ? CH_SPROPER : 0u);
}

// meta word of bdk_aread
BDK_HD uint32_t make_meta(uint32_t cr, int lib, uint32_t bdqual) {
    return (cr & CR_FLAG_MASK) | ((cr & CR_REV) ? 16u : 0u) | ((uint32_t)lib << 8) | ((bdqual & 0xFFu) << 16);
}
BDK_HD int meta_flag(uint32_t m) { return (int)(m & 0xF); }
BDK_HD int meta_rev(uint32_t m) { return (int)((m >> 4) & 1); }
BDK_HD int meta_lib(uint32_t m) { return (int)((m >> 8) & 0xFF); }

// ---- Poisson / gamma -------------------------------------------------------------------------
// Regularised lower incomplete gamma P(a, x) in double: series for x < a + 1, modified Lentz
// continued fraction of Q otherwise. log P[Pois(lambda) > k] = log P(k + 1, lambda).
// Each of P and Q is computed directly where it is the small one.
BDK_HD void gamma_pq_d(double a, double x, double* p, double* q) {
    if (!(x > 0.0)) { *p = 0.0; *q = 1.0; return; }
    double lg = lgamma(a);
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 10000; ++n) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        *p = sum * exp(-x + a * log(x) - lg);
        *q = 1.0 - *p;
        return;
    }
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 10000; ++i) {
        double an = -(double)i * ((double)i - a);
        b += 2.0;
        d = an * d + b; if (fabs(d) < tiny) d = tiny;
        c = b + an / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    *q = exp(-x + a * log(x) - lg) * h;
    *p = 1.0 - *q;
}
BDK_HD double gamma_p_d(double a, double x) { double p, q; gamma_pq_d(a, x, &p, &q); return p; }
BDK_HD double gamma_q_d(double a, double x) { double p, q; gamma_pq_d(a, x, &p, &q); return q; }
BDK_HD double poisson_log_sf(double lambda, int k) { return log(gamma_p_d((double)k + 1.0, lambda)); }

// ---- region building (K2) ----------------------------------------------------------------------
// push_read's break test (BreakDancer.cpp:216)
BDK_HD bool k2_is_break(int32_t prev_tid, int32_t prev_pos, int32_t tid, int32_t pos, int32_t window) {
    return tid != prev_tid || pos - prev_pos > window;
}

struct CandAgg { int32_t ntot, maxlen, fwd, rev, nonctx; };

// Candidate region = anomalous reads [s, e]. _ntotal_nucleotides / _max_readlen skip the
// candidate's first read and include the read that triggers the break, i.e. range (s, e + 1]
// (BreakDancer.cpp:209-212 run before the break test); the stream's last candidate has no
// break read. fwd/rev/non-CTX counts are over [s, e] (ReadRegionData.cpp:99-107).
BDK_HD CandAgg k2_cand_aggregate(const bdk_aread* ar, int64_t s, int64_t e, int64_t A) {
    CandAgg g; g.ntot = 0; g.maxlen = 0; g.fwd = 0; g.rev = 0; g.nonctx = 0;
    for (int64_t j = s; j <= e; ++j) {
        uint32_t m = ar[j].meta;
        if ((m & 0xF) != BDK_ARP_CTX) ++g.nonctx;
        if (m & 16u) ++g.rev; else ++g.fwd;
    }
    int64_t hi = e + 1 < A ? e + 1 : A - 1;
    for (int64_t j = s + 1; j <= hi; ++j) {
        int32_t q = ar[j].qlen;
        g.ntot += q;
        if (q > g.maxlen) g.maxlen = q;
    }
    return g;
}

// process_breakpoint's acceptance test (BreakDancer.cpp:245-247)
BDK_HD bool k2_accept(int32_t start_pos, int32_t end_pos, const CandAgg& g, int32_t min_len, int32_t cov_lim) {
    float cov = f_div((float)g.ntot, (float)(end_pos - start_pos + 1 + g.maxlen));
    return (end_pos - start_pos > min_len) && (cov < (float)cov_lim);
}

// ---- connection walk (K4) ----------------------------------------------------------------------
struct RegionRec {            // BasicRegion + bookkeeping
    int32_t tid, start, end;
    int32_t fwd, rev;
    int32_t first_read, n_reads;
    int32_t stored;           // read vector kept (ReadRegionData.cpp:118-121)
    int32_t cand;             // index of the candidate region it came from (time stamp)
};

struct K4Static {                 // what the scoring of a call needs (k4_score_row)
    const bdk_aread* ar;
    const RegionRec* reg;
    const uint32_t* P;            // [A][nkey] inclusive proper-pair prefix counts per key
    const int32_t* cand_maxlen;   // [ncand] _max_readlen when the candidate was closed
    const float* lib_mean;        // [nlib] LibraryConfig::mean_insertsize
    const uint32_t* hist;         // [nlib][BDK_NUM_FLAGS] pass-1 read_counts_by_flag
    const float* density;         // [nkey]
    uint64_t A;
    int32_t nreg, ncand, period, nkey, nlib;
    int32_t chr_restricted, min_read_pair, score_threshold, fisher;
    uint32_t covered_ref_len;
};
constexpr int32_t K4_NEVER = 0x7f7f7f7f;   // byte pattern 0x7f: tables are initialised with memset

struct K4Mut {                    // per call slot
    bdk_sv* rows;           // [nrow_cap]
    int32_t* row_lib_count; // [nrow_cap][nlib]
    int32_t* row_lib_span;  // [nrow_cap][nlib]
    uint32_t* row_cn_count; // [nrow_cap][nkey]
    float* row_cn;          // [nrow_cap][nkey]
    uint8_t* row_emit;      // [nrow_cap] K4_ROW_*
};

struct WindowInfo { int32_t cF; int32_t maxlen; int32_t last_region; int32_t w; };
enum { K4_ROW_NONE = 0, K4_ROW_EMIT = 1, K4_ROW_PENDING = 2 };   // row_emit[]: no call / call to print / pairs counted, not scored yet

BDK_HD WindowInfo k4_window_info(const K4Static& S, int w) {
    WindowInfo wi;
    wi.w = w;
    int64_t trigger = ((int64_t)w + 1) * S.period - 1;
    if (trigger < S.nreg) {          // flush triggered by the registration of region `trigger`
        wi.cF = S.reg[trigger].cand;
        wi.maxlen = S.cand_maxlen[wi.cF];
        wi.last_region = (int32_t)trigger;
    } else {                          // final build_connection (BreakDancer.cpp:536-541)
        wi.cF = 0x7fffffff;
        wi.maxlen = S.ncand > 0 ? S.cand_maxlen[S.ncand - 1] : 0;
        wi.last_region = S.nreg - 1;
    }
    return wi;
}

// ---- execution policy of the connection walk --------------------------------------------------
// The walk over one connected component is sequential (it reproduces build_connection's order), but
// the loops over the reads of a region are independent per read: a "team" spreads them over its
// lanes. SoloTeam = one thread (host simulation, tests); WarpTeam = the 32 lanes of a warp (K4).
// Every lane of a team runs the same control flow on the same values; writes to shared state are
// done by lane 0 between two sync()s.
struct SoloTeam {
    BDK_HD int lane() const { return 0; }
    BDK_HD int width() const { return 1; }
    BDK_HD int sum(int v) const { return v; }
    BDK_HD int min(int v) const { return v; }
    BDK_HD int max(int v) const { return v; }
    BDK_HD bool any(bool p) const { return p; }
    BDK_HD void sync() const {}
    BDK_HD void add(int32_t* p, int v) const { *p += v; }
    BDK_HD uint32_t fetch_inc(uint32_t* p) const { return (*p)++; }
};
#if defined(__CUDACC__)
struct Tile8Team {           // 8 consecutive lanes of a warp: four regions (or calls) per warp at a time, 128 contiguous bytes of ReadInfo2 per load
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 7u); }
    __device__ __forceinline__ int width() const { return 8; }
    __device__ __forceinline__ unsigned mask() const { return 0xffu << (threadIdx.x & 24u); }
    __device__ __forceinline__ int sum(int v) const { const unsigned m = mask(); v += __shfl_xor_sync(m, v, 4); v += __shfl_xor_sync(m, v, 2); v += __shfl_xor_sync(m, v, 1); return v; }
    __device__ __forceinline__ int min(int v) const { const unsigned m = mask(); v = ::min(v, __shfl_xor_sync(m, v, 4)); v = ::min(v, __shfl_xor_sync(m, v, 2)); v = ::min(v, __shfl_xor_sync(m, v, 1)); return v; }
    __device__ __forceinline__ int max(int v) const { const unsigned m = mask(); v = ::max(v, __shfl_xor_sync(m, v, 4)); v = ::max(v, __shfl_xor_sync(m, v, 2)); v = ::max(v, __shfl_xor_sync(m, v, 1)); return v; }
    __device__ __forceinline__ bool any(bool p) const { return __ballot_sync(mask(), p) != 0; }
    __device__ __forceinline__ void sync() const { __syncwarp(mask()); }
    __device__ __forceinline__ void add(int32_t* p, int v) const { atomicAdd(p, v); }
};
struct WarpTeam {
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31u); }
    __device__ __forceinline__ int width() const { return 32; }
    __device__ __forceinline__ int sum(int v) const { return (int)__reduce_add_sync(0xffffffffu, v); }
    __device__ __forceinline__ int min(int v) const { return __reduce_min_sync(0xffffffffu, v); }
    __device__ __forceinline__ int max(int v) const { return __reduce_max_sync(0xffffffffu, v); }
    __device__ __forceinline__ bool any(bool p) const { return __any_sync(0xffffffffu, p) != 0; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ void add(int32_t* p, int v) const { atomicAdd(p, v); }
    __device__ __forceinline__ uint32_t fetch_inc(uint32_t* p) const { return atomicAdd(p, 1u); }
};
#endif

// Second half of process_sv for row slot `row` left PENDING by k4n_call (BreakDancer.cpp:377-497): breakpoint coordinates,
// copy number, size, ComputeProbScore, the -y cut. One thread per row.
BDK_HD void k4_score_row(const K4Static& S, K4Mut& M, int row) {
    bdk_sv& o = M.rows[row];
    const int s0 = o.region[0], s1 = o.region[1], flag = o.flag, nflag = o.num_pairs, w = o.window;
    const int n = s1 >= 0 ? 2 : 1;
    const WindowInfo wi = k4_window_info(S, w);
    const int32_t* lib_count = M.row_lib_count + (int64_t)row * S.nlib;
    const int32_t* lib_span = M.row_lib_span + (int64_t)row * S.nlib;

    const RegionRec& R0 = S.reg[s0];
    int chr0 = R0.tid, chr1, pos0 = R0.start, pos1 = R0.end;
    int fwd0 = R0.fwd, rev0 = R0.rev, fwd1, rev1;
    int ml = wi.maxlen;
    if (n == 2) {
        const RegionRec& R1 = S.reg[s1];
        if (flag == BDK_ARP_RF) pos1 = R1.end + ml - 5;
        else if (flag == BDK_ARP_FF) { pos0 = pos1; pos1 = R1.end + ml - 5; }
        else if (flag == BDK_ARP_RR) pos1 = R1.start;
        else { pos0 = pos1; pos1 = R1.start; }
        chr1 = R1.tid; fwd1 = R1.fwd; rev1 = R1.rev;
    } else {
        fwd1 = fwd0; rev1 = rev0; chr1 = R0.tid; pos1 = R0.end;
    }
    // copy number (accumulate_reads_between_regions telescopes to P[first read of s1] - P[last read of s0])
    uint32_t* cnc = M.row_cn_count + (int64_t)row * S.nkey;
    float* cnv = M.row_cn + (int64_t)row * S.nkey;
    float cn_sum = 0.0f; int cn_n = 0;
    for (int k = 0; k < S.nkey; ++k) {
        uint32_t c = 0;
        if (n == 2) {
            const RegionRec& R1 = S.reg[s1];
            c = S.P[(uint64_t)R1.first_read * S.nkey + k] - S.P[(uint64_t)(R0.first_read + R0.n_reads - 1) * S.nkey + k];
        }
        cnc[k] = c;
        float v = 0.0f;
        if (c) {
            v = f_mul(f_div((float)c, f_mul(S.density[k], (float)(pos1 - pos0))), 2.0f);
            cn_sum = f_add(cn_sum, v);
            ++cn_n;
        }
        cnv[k] = v;
    }
    cn_sum = f_div(cn_sum, f_mul(2.0f, (float)cn_n));
    float af = f_sub(1.0f, cn_sum);

    if (flag != BDK_ARP_RF && flag != BDK_ARP_RR && pos0 + ml - 5 < pos1) pos0 += ml - 5;

    float diff = 0.0f;
    for (int l = 0; l < S.nlib; ++l)
        if (lib_count[l])
            diff = f_add(diff, f_sub((float)lib_span[l], f_mul((float)lib_count[l], S.lib_mean[l])));
    int diffspan = (int)((double)f_div(diff, (float)nflag) + 0.5);

    int total_region_size = (R0.end - R0.start + 1) + (n == 2 ? (S.reg[s1].end - S.reg[s1].start + 1) : 0);
    double logp = 0.0, err = 0.0;
    int nl = 0;
    for (int l = 0; l < S.nlib; ++l) {
        if (!lib_count[l]) continue;
        ++nl;
        uint32_t cnt = S.hist[l * BDK_NUM_FLAGS + flag];
        double lambda = (double)total_region_size * ((double)cnt / (double)S.covered_ref_len);
        lambda = 1.0e-10 < lambda ? lambda : 1.0e-10;      // std::max(1e-10, lambda)
        double tmp_a = poisson_log_sf(lambda, lib_count[l]) - err;
        double tmp_b = logp + tmp_a;
        err = (tmp_b - logp) - tmp_a;
        logp = tmp_b;
    }
    if (S.fisher && logp < 0) {
        double fp = gamma_q_d((double)nl, -logp);
        logp = fp > exp(-99.0) ? log(fp) : -99.0;
    }
    double phred_tmp = -10.0 * logp / log(10.0);
    int phred = phred_tmp > 99 ? 99 : (int)(phred_tmp + 0.5);
    ++pos0; ++pos1;
    if (!(phred > S.score_threshold)) { M.row_emit[row] = K4_ROW_NONE; return; }
    o.chr[0] = chr0; o.chr[1] = chr1; o.pos[0] = pos0; o.pos[1] = pos1;
    o.fwd[0] = fwd0; o.fwd[1] = fwd1; o.rev[0] = rev0; o.rev[1] = rev1;
    o.diffspan = diffspan; o.score = phred;
    o.logp = logp; o.allele_frequency = af; o.cn_present = 0;
    o.order = 0;
    M.row_emit[row] = K4_ROW_EMIT;
}

// =================================================================================================
// K4, second formulation: the connection walk WITHOUT walking.
//
// What the reference's build_connection / process_sv / is_region_final / clear_region do to the reads (the statements
// cited at the top of this file) has a closed form once one fact per region is known -- the flush window in which the
// region is cleared, del[region] (K4_NEVER: never):
//
//  * a pair (x, y) whose reads sit in two DIFFERENT stored regions rx < ry is only ever looked at by process_sv(rx, ry),
//    which is called at most once, in the window of ry, iff the edge is followed (weight >= -r) and both regions still
//    exist then; both reads are still held at that moment (nothing else can remove them), so the call consumes the pair;
//  * a pair with both reads in ONE region is consumed by the first process_sv call that involves the region;
//  * a read whose later mate sat in a rejected (collapsed) candidate is dropped by the first call involving its region
//    after that collapse; a read whose mate is absent, or sat EARLIER in a collapsed candidate, is held for ever;
//  * whether names were freed (the two early returns of process_sv) never matters: freed reads are no longer held.
//
// So "is read j still held at the end of window w, and does it keep its region from being final?" is a function of del[]
// of the mate's region and of the set of windows in which the region takes part in a call, which again depends only on del[]
// of its partners. The table is the fixed point of  del[v] = first active window of v in which no held read blocks
// (k4n_region_deletion); every event depends only on strictly earlier events (windows in order, inside a window the
// active nodes ascending), so the fixed point is unique and any iteration order reaches it. With the table known, the
// calls of every window are independent of all other windows (k4n_window_calls gives build_connection's order, which
// decides who is the FIRST call of a region; k4n_call counts the pairs of one call).
// =================================================================================================
enum : uint32_t { RI_STRONG = 1u << 24,        // the edge (own region, mate's region) has weight >= -r: process_sv is called for it
                  RI_MATE_STORED = 1u << 25 }; // the mate's region kept its reads (ReadRegionData.cpp:118-121)

struct alignas(16) ReadInfo2 {
    int32_t mate;          // index of the mate in the anomalous-read stream, or -1
    int32_t mate_region;   // region of the mate, or -1 (collapsed candidate / no mate)
    int32_t wthr;          // mate later and collapsed: first flush window whose trigger lies behind that collapse
    uint32_t meta;         // bdk_aread::meta | RI_*
};

struct K4N {
    const ReadInfo2* ri;          // [A]
    const bdk_aread* ar;
    const RegionRec* reg;
    int32_t nreg, period, chr_restricted, min_read_pair;
};

BDK_HD int32_t k4n_ld(const int32_t* p) {      // the table is rewritten in place by other warps during a sweep
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
BDK_HD int k4n_last_region(const K4N& S, int w) {          // last_region_idx() at the flush of window w
    const int64_t trigger = ((int64_t)w + 1) * S.period - 1;
    return trigger < S.nreg ? (int)trigger : S.nreg - 1;
}
BDK_HD bool k4n_before(int d, int w, int rm, int v) { return d < w || (d == w && rm < v); }   // rm cleared before (w, v)?

// cand_regs[c] = number of regions registered up to and including candidate c = index g of the first region behind c; the first
// flush window whose trigger lies behind c (the collapse of c has happened by then) is g / period
BDK_HD ReadInfo2 k4n_make_read_info(const bdk_aread* ar, const int32_t* mate, const int32_t* read_region, const int32_t* read_cand,
                                    const RegionRec* reg, const int32_t* cand_regs, int period, int j, bool strong) {
    ReadInfo2 r;
    r.mate = mate[j];
    r.mate_region = r.mate >= 0 ? read_region[r.mate] : -1;
    r.wthr = 0;
    r.meta = ar[j].meta & 0x00ffffffu;
    if (r.mate >= 0 && r.mate_region < 0 && r.mate > j) r.wthr = cand_regs[read_cand[r.mate]] / period;
    if (r.mate_region >= 0) {
        if (strong) r.meta |= RI_STRONG;
        if (reg[r.mate_region].stored) r.meta |= RI_MATE_STORED;
    }
    return r;
}

// The flush window in which region v is cleared, given the deletion windows del[] of the other regions.
// One pass over the region's reads finds its first active window (and whether a read is held for ever with a name entry
// of size 1, which keeps the region from ever being final); then one pass per active window tried.
template <class Team>
BDK_HD int k4n_region_deletion(const Team& T, const K4N& S, const int32_t* del, int v) {
    const RegionRec R = S.reg[v];
    const int j0 = R.first_read, j1 = R.first_read + R.n_reads, wv = v / S.period;
    int w = K4_NEVER, wlast = -1;
    bool never = false;
    for (int j = j0 + T.lane(); j < j1; j += T.width()) {
        const ReadInfo2 I = S.ri[j];
        const int rm = I.mate_region;
        const bool skip = S.chr_restricted && meta_flag(I.meta) == BDK_ARP_CTX;
        if (rm >= 0) {
            const int aw = (rm > v ? rm : v) / S.period;
            if (aw < w) w = aw;
            if (!skip && rm != v && aw > wlast) wlast = aw;
        } else if (R.stored && !skip && (I.mate < 0 || I.mate < j)) never = true;
    }
    if (T.any(never)) return K4_NEVER;
    w = T.min(w);                                      // the windows of v's edges = max(v, mate's region) / period
    // A read whose mate is not registered yet blocks: nothing can happen before the window of the last mate (itself an
    // active window). Without -o that is the region's last active window: it is cleared there or never.
    wlast = T.max(wlast);
    if (w != K4_NEVER && wlast > w) w = wlast;
    while (w != K4_NEVER) {
        // lc: latest window <= w in which process_sv was called with v; maxc: latest collapse a held read waits for
        int wn = K4_NEVER, lc = -1, intra = 0, maxc = -1;
        bool bad = false;
        for (int j = j0 + T.lane(); j < j1; j += T.width()) {
            const ReadInfo2 I = S.ri[j];
            const int rm = I.mate_region;
            const bool skip = S.chr_restricted && meta_flag(I.meta) == BDK_ARP_CTX;     // is_region_final does not look at these
            if (rm < 0) { if (!skip && I.wthr > maxc) maxc = I.wthr; continue; }         // later mate collapsed: held until a call after that drops it
            if (rm == v) { ++intra; continue; }                                           // an unconsumed pair inside v has a name entry of size 2 (its window, wv, is v's first)
            const int aw = (rm > v ? rm : v) / S.period;
            if (aw > w && aw < wn) wn = aw;
            const int d = k4n_ld(del + rm);
            if ((I.meta & RI_STRONG) && aw <= w && d >= aw) {                             // process_sv(v, rm) was called in window aw
                if (aw > lc) lc = aw;
                if (I.meta & RI_MATE_STORED) continue;                                    // ... and consumed the pair
            }
            if (skip) continue;
            if (w < aw) bad = true;                                                       // mate not registered yet
            else if ((I.meta & RI_MATE_STORED) && k4n_before(d, w, rm, v)) bad = true;    // clear_region(rm) took rm out of the name entry
        }
        wn = T.min(wn); lc = T.max(lc); maxc = T.max(maxc); intra = T.sum(intra);
        const bool blocked = T.any(bad);
        if (v != k4n_last_region(S, w)) {
            if (!R.stored) return w;                   // holds no reads: final the first time it is asked
            if (intra / 2 >= S.min_read_pair && wv > lc) lc = wv;                         // the self loop: process_sv(v) in v's own window
            if (!blocked && !(maxc > lc)) return w;
        }
        w = wn;
    }
    return K4_NEVER;
}

// first window in which process_sv is called with region v (K4_NEVER: never), given the final table
template <class Team>
BDK_HD int k4n_first_call(const Team& T, const K4N& S, const int32_t* del, int v) {
    const RegionRec R = S.reg[v];
    const int dv = del[v];
    int c1 = K4_NEVER, intra = 0;
    for (int j = R.first_read + T.lane(); j < R.first_read + R.n_reads; j += T.width()) {
        const ReadInfo2 I = S.ri[j];
        const int rm = I.mate_region;
        if (rm < 0) continue;
        if (rm == v) { ++intra; continue; }
        if (!(I.meta & RI_STRONG)) continue;
        const int wp = (rm > v ? rm : v) / S.period;
        if (wp < c1 && del[rm] >= wp && dv >= wp) c1 = wp;
    }
    c1 = T.min(c1);
    const int wv = v / S.period;
    if (T.sum(intra) / 2 >= S.min_read_pair && wv < c1 && dv >= wv) c1 = wv;
    return c1;
}

// ---- the calls of one flush window, in build_connection's order --------------------------------------------------
// e[0..n): the directed copies (src, dst) of the followed edges (weight >= -r) of window w, sorted by (src, dst);
// fl[0..n): flags; queue: n + 1 entries.
//   k4n_window_prepare  (per edge, independent): an edge one of whose regions is already cleared is as good as erased; a
//                       region with a call in an earlier window (c1[] < w) starts out touched;
//   k4n_window_calls    sequential, no table reads: vertices ascending, BFS from each, a tail's edges ascending, every live
//                       edge followed once (BreakDancer.cpp:280-338). Call k goes to slot0 + k; returns the number of calls.
// A call is FIRST for one of its regions when no earlier call (in an earlier window, or earlier in this one) involved it.
struct SEdge { int32_t dst, src; };   // as one u64: src << 32 | dst
enum : uint8_t { SE_ERASED = 1, SE_VDONE = 2, SE_TOUCHED = 4 };
enum : uint8_t { K4_ROW_CALL = 4, K4_ROW_FIRST0 = 8, K4_ROW_FIRST1 = 16 };   // row_emit[] of a slot between k4n_window_calls and k4n_call

template <class E> BDK_HD int k4n_find_run(const E* e, int n, int src) {
    int a = 0, b = n;
    while (a < b) { const int m = (a + b) >> 1; if (e[m].src < src) a = m + 1; else b = m; }
    return (a < n && e[a].src == src) ? a : -1;
}
template <class E> BDK_HD int k4n_find_edge(const E* e, int n, int src, int dst) {
    int a = 0, b = n;
    while (a < b) { const int m = (a + b) >> 1; if (e[m].src < src || (e[m].src == src && e[m].dst < dst)) a = m + 1; else b = m; }
    return (a < n && e[a].src == src && e[a].dst == dst) ? a : -1;
}

template <class E, class F> BDK_HD void k4n_window_prepare(const int32_t* del, const int32_t* c1, const E* e, int n, F* fl, int w, int k) {
    const int src = e[k].src, dst = e[k].dst;
    uint8_t f = (del[src] < w || del[dst] < w) ? SE_ERASED : 0;      // !region_exists(tail) / !region_exists(s1)
    if ((k == 0 || e[k - 1].src != src) && c1[src] < w) f |= SE_TOUCHED;
    fl[k] = f;
}

template <class E, class F, class Q>
BDK_HD int k4n_window_calls(const E* e, int n, F* fl, Q* queue, int w, int slot0, bdk_sv* rows, uint8_t* row_emit) {
    int slot = slot0;
    for (int vi = 0; vi < n;) {
        const int v = e[vi].src;
        int vend = vi;
        while (vend < n && e[vend].src == v) ++vend;
        if (!(fl[vi] & SE_VDONE)) {
            int qa = 0, qb = 1, qn;
            queue[0] = v;
            while (qa < qb) {
                qn = qb;
                for (int t = qa; t < qb; ++t) {
                    const int tail = queue[t];
                    const int ts = tail == v ? vi : k4n_find_run(e, n, tail);
                    if (ts < 0 || (fl[ts] & SE_VDONE)) continue;               // graph.find(tail) == end
                    for (int k = ts; k < n && e[k].src == tail; ++k) {
                        if (fl[k] & SE_ERASED) continue;
                        const int s1 = e[k].dst;
                        fl[k] |= SE_ERASED;
                        int r1 = ts;
                        if (s1 != tail) {
                            const int rq = k4n_find_edge(e, n, s1, tail);       // erase_edge(s1, tail)
                            if (rq >= 0) fl[rq] |= SE_ERASED;
                            r1 = k4n_find_run(e, n, s1);
                        }
                        queue[qn++] = s1;
                        const bool first_t = !(fl[ts] & SE_TOUCHED);
                        const bool first_s = s1 == tail ? first_t : !(fl[r1] & SE_TOUCHED);
                        fl[ts] |= SE_TOUCHED; fl[r1] |= SE_TOUCHED;
                        const int a = tail < s1 ? tail : s1, b = tail < s1 ? s1 : tail;
                        bdk_sv& o = rows[slot];
                        o.region[0] = a; o.region[1] = s1 != tail ? b : -1; o.window = w;
                        const bool f0 = a == tail ? first_t : first_s, f1 = a == tail ? first_s : first_t;
                        row_emit[slot] = (uint8_t)(K4_ROW_CALL | (f0 ? K4_ROW_FIRST0 : 0) | (s1 != tail && f1 ? K4_ROW_FIRST1 : 0));
                        ++slot;
                    }
                    fl[ts] |= SE_VDONE;                                         // graph.erase(tail)
                }
                qa = qb; qb = qn;
            }
        }
        vi = vend;
    }
    return slot - slot0;
}

// process_sv's pairing (SvBuilder.cpp:101-118, BreakDancer.cpp:363-375) for the call in slot `row`: every pair between the
// two regions, plus the pairs inside a region this call is the first for. Leaves the slot PENDING (to be scored) or NONE.
struct K4NOut {
    int32_t* sv_of_read; bdk_sv* rows; int32_t* row_lib_count; int32_t* row_lib_span; uint8_t* row_emit; int32_t nlib;
};
template <class Team>
BDK_HD void k4n_call(const Team& T, const K4N& S, const K4NOut& M, int row) {
    const uint8_t st = M.row_emit[row];
    const int s0 = M.rows[row].region[0], s1 = M.rows[row].region[1];
    T.sync();
    const int n = s1 >= 0 ? 2 : 1;
    const int sn[2] = {s0, s1};
    const bool first[2] = {(st & K4_ROW_FIRST0) != 0, (st & K4_ROW_FIRST1) != 0};
    int c[BDK_NUM_FLAGS];
    for (int f = 0; f < BDK_NUM_FLAGS; ++f) c[f] = 0;
    auto pairs = [&](int i, int y, const ReadInfo2& I) -> bool {        // is y the later read of a pair this call consumes?
        const int x = I.mate;
        if (x < 0 || x >= y) return false;
        const int rx = I.mate_region;
        if (rx < 0) return false;
        if (rx == sn[i]) return first[i];
        return i == 1 && rx == s0 && (I.meta & RI_MATE_STORED);
    };
    for (int i = 0; i < n; ++i) {
        const RegionRec R = S.reg[sn[i]];
        if (!R.stored) continue;
        for (int y = R.first_read + T.lane(); y < R.first_read + R.n_reads; y += T.width()) {
            const ReadInfo2 I = S.ri[y];
            if (pairs(i, y, I)) ++c[meta_flag(I.meta)];
        }
    }
    int num_pairs = 0, flag = BDK_NA, best = 0;
    for (int f = 0; f < BDK_NUM_FLAGS; ++f) { c[f] = T.sum(c[f]); num_pairs += c[f]; if (c[f] > best) { best = c[f]; flag = f; } }
    const bool early = num_pairs < S.min_read_pair || c[flag] < S.min_read_pair;
    int32_t* lib_count = M.row_lib_count + (int64_t)row * M.nlib;
    int32_t* lib_span = M.row_lib_span + (int64_t)row * M.nlib;
    if (!early) for (int l = T.lane(); l < M.nlib; l += T.width()) { lib_count[l] = 0; lib_span[l] = 0; }
    T.sync();
    if (num_pairs)
        for (int i = 0; i < n; ++i) {
            const RegionRec R = S.reg[sn[i]];
            if (!R.stored) continue;
            for (int y = R.first_read + T.lane(); y < R.first_read + R.n_reads; y += T.width()) {
                const ReadInfo2 I = S.ri[y];
                if (!pairs(i, y, I)) continue;
                M.sv_of_read[I.mate] = row; M.sv_of_read[y] = row;
                if (!early && meta_flag(I.meta) == flag) {
                    const int l = meta_lib(I.meta);
                    T.add(lib_count + l, 1);
                    T.add(lib_span + l, S.ar[y].abs_isize);
                }
            }
        }
    T.sync();
    if (T.lane() != 0) return;
    if (early) { M.row_emit[row] = K4_ROW_NONE; return; }
    bdk_sv& o = M.rows[row];
    o.flag = flag; o.num_pairs = c[flag];
    M.row_emit[row] = K4_ROW_PENDING;
}

}  // namespace bdk
