// bdk_finalize.h -- scalar host-side steps between the kernels: turning the pass-1 accumulators
// into BamSummary numbers, read densities and the region window
// (reference BamSummary.cpp:116-150, BreakDancerMax.cpp:83-116). Plain C++, shared by the
// CUDA orchestration (bdk_core.cu) and the host simulation used in tests.
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "../../include/bdk.h"
#include "bdk_logic.h"

namespace bdk {

inline int nkey_of(const bdk_params& p) { return p.cn_lib ? p.nlib : p.nbam; }
inline int period_of(const bdk_params& p) { return std::max(1, p.buffer_size + 1); }  // BreakDancer.cpp:254-259
// first process_breakpoint() runs on empty state and registers an empty region when 0 > -s
// (BreakDancer.cpp:244-252 with _region_*_pos == -1)
inline int dummy_region_of(const bdk_params& p) { return (0 > p.min_len && 0.0f < (float)p.seq_coverage_lim) ? 1 : 0; }

inline std::vector<LibDev> make_libdev(const bdk_params& p) {
    std::vector<LibDev> v(p.nlib);
    for (int i = 0; i < p.nlib; ++i) {
        const bdk_lib& l = p.libs[i];
        v[i].upper = l.uppercutoff; v[i].lower = l.lowercutoff;
        v[i].min_mapq = l.min_mapping_quality < 0 ? p.min_map_qual : l.min_mapping_quality;
        v[i].key = p.cn_lib ? i : l.bam_index;
    }
    return v;
}

inline std::vector<float> make_lib_mean(const bdk_params& p) {
    std::vector<float> v(p.nlib);
    for (int i = 0; i < p.nlib; ++i) v[i] = p.libs[i].mean_insertsize;
    return v;
}

// Raw pass-1 accumulators, as the classify kernel leaves them.
struct FinalizeIn {
    int32_t nlib, nbam, nrg, ntid, cn_lib, initial_window;
    const bdk_lib* libs;
    const int32_t* rg_lib;
    const int32_t* rg_bam;
    const unsigned long long* rg_sproper;   // [nrg] proper-pair && mapq-pass records per read group
    const uint32_t* hist;                   // [nlib][BDK_NUM_FLAGS]
    const unsigned long long* first;        // [nbam][ntid] (record index << 32 | pos) of the first record
    const unsigned long long* last;         //              ... of the last record; first = ~0 when none
};

// ref_len contribution of (bam, tid): the reference's sum of consecutive same-tid position
// differences (BamSummary.cpp:70-74) telescopes to last - first.
BDK_HD int64_t ref_len_term(unsigned long long first, unsigned long long last) {
    if (first == ~0ull) return 0;
    return (int64_t)(int32_t)(uint32_t)last - (int64_t)(int32_t)(uint32_t)first;
}

// BamSummary::_analyze_bams tail + density/window block of main(). ref_len: [nbam]; density: [nkey].
BDK_HD void finalize_rest(const FinalizeIn& in, const unsigned long long* ref_len, uint64_t n_records, uint64_t n_anom,
                          bdk_summary_t* S, float* density) {
    S->n_records = n_records; S->n_anomalous = n_anom;
    for (int i = 0; i < BDK_MAX_BAMS; ++i) { S->read_count_per_bam[i] = 0; S->ref_len_per_bam[i] = 0; }
    for (int i = 0; i < BDK_MAX_LIBS; ++i) {
        S->lib_read_count[i] = 0; S->seq_coverage[i] = 0; S->read_density[i] = 0;
        for (int f = 0; f < BDK_NUM_FLAGS; ++f) S->read_counts_by_flag[i][f] = 0;
    }
    for (int rg = 0; rg < in.nrg; ++rg) {
        int lib = in.rg_lib[rg], bam = in.rg_bam[rg];
        if (lib >= 0 && lib < in.nlib) S->lib_read_count[lib] += (uint32_t)in.rg_sproper[rg];
        if (bam >= 0 && bam < in.nbam) S->read_count_per_bam[bam] += (uint32_t)in.rg_sproper[rg];
    }
    uint32_t covered = 0;
    for (int b = 0; b < in.nbam; ++b) {
        S->ref_len_per_bam[b] = ref_len[b];
        if (covered < ref_len[b]) covered = (uint32_t)ref_len[b];   // uint32_t _covered_ref_len (BamSummary.cpp:125)
    }
    S->covered_ref_len = covered;
    for (int l = 0; l < in.nlib; ++l)
        for (int f = 0; f < BDK_NUM_FLAGS; ++f) S->read_counts_by_flag[l][f] = in.hist[l * BDK_NUM_FLAGS + f];
    int window = in.initial_window;
    int nkey = in.cn_lib ? in.nlib : in.nbam;
    for (int k = 0; k < nkey; ++k) density[k] = 0.0f;
    for (int i = 0; i < in.nlib; ++i) {
        uint32_t n = S->lib_read_count[i];
        float covg = 0;
        if (n != 0 && covered != 0) covg = f_div(f_mul((float)n, in.libs[i].readlens), (float)covered);
        S->seq_coverage[i] = covg;
        float dens = 0.000001f;
        if (in.cn_lib) { if (n != 0) dens = f_div((float)n, (float)covered); }
        else dens = f_div((float)S->read_count_per_bam[in.libs[i].bam_index], (float)covered);
        density[in.cn_lib ? i : in.libs[i].bam_index] = dens;
        S->read_density[i] = dens;
        int disc = (int)(S->read_counts_by_flag[i][BDK_ARP_LARGE_INSERT] + S->read_counts_by_flag[i][BDK_ARP_SMALL_INSERT]);
        int tmp = (disc > 0) ? (int)f_div((float)covered, (float)disc) : 50;
        window = window < tmp ? window : tmp;
    }
    S->window = window;
}

// Host convenience used by the test harness.
struct SummaryAcc {
    std::vector<unsigned long long> rg_sproper, first, last;
    std::vector<uint32_t> hist;
};

inline void finalize_summary(const bdk_params& p, const SummaryAcc& acc, uint64_t n_records, uint64_t n_anom,
                             bdk_summary_t* S, std::vector<float>* density) {
    FinalizeIn in{p.nlib, p.nbam, p.nrg, p.ntid, p.cn_lib, p.initial_window, p.libs, p.rg_lib, p.rg_bam,
                  acc.rg_sproper.data(), acc.hist.data(), acc.first.data(), acc.last.data()};
    std::vector<unsigned long long> ref_len(std::max(1, p.nbam), 0);
    for (int b = 0; b < p.nbam; ++b)
        for (int t = 0; t < p.ntid; ++t)
            ref_len[b] += (unsigned long long)ref_len_term(acc.first[(size_t)b * p.ntid + t], acc.last[(size_t)b * p.ntid + t]);
    density->assign(std::max(1, nkey_of(p)), 0.0f);
    memset(S, 0, sizeof(*S));
    finalize_rest(in, ref_len.data(), n_records, n_anom, S, density->data());
}

}  // namespace bdk
