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
        v[i].upper = l.uppercutoff; v[i].lower = l.lowercutoff; v[i].mean = l.mean_insertsize;
        v[i].min_mapq = l.min_mapping_quality < 0 ? p.min_map_qual : l.min_mapping_quality;
        v[i].key = p.cn_lib ? i : l.bam_index;
    }
    return v;
}

// Raw pass-1 accumulators, as the classify kernel leaves them.
struct SummaryAcc {
    std::vector<uint64_t> rg_sproper;   // [nrg] proper-pair && mapq-pass records per read group
    std::vector<uint32_t> hist;         // [nlib][BDK_NUM_FLAGS]
    std::vector<uint64_t> first, last;  // [nbam][ntid] (record index << 32 | pos) of the first / last
                                        // record of (bam, tid); first = ~0 when none
};

// BamSummary::_analyze_bams tail + density/window block of main(). density is per key.
inline void finalize_summary(const bdk_params& p, const SummaryAcc& acc, uint64_t n_records, uint64_t n_anom,
                             bdk_summary_t* S, std::vector<float>* density) {
    memset(S, 0, sizeof(*S));
    S->n_records = n_records; S->n_anomalous = n_anom;
    for (int rg = 0; rg < p.nrg; ++rg) {
        int lib = p.rg_lib[rg], bam = p.rg_bam[rg];
        if (lib >= 0) S->lib_read_count[lib] += (uint32_t)acc.rg_sproper[rg];
        if (bam >= 0 && bam < BDK_MAX_BAMS) S->read_count_per_bam[bam] += (uint32_t)acc.rg_sproper[rg];
    }
    uint32_t covered = 0;
    for (int b = 0; b < p.nbam; ++b) {
        uint64_t ref_len = 0;
        for (int t = 0; t < p.ntid; ++t) {
            uint64_t f = acc.first[(size_t)b * p.ntid + t], l = acc.last[(size_t)b * p.ntid + t];
            if (f == ~0ull) continue;
            // sum of consecutive same-tid position differences telescopes to last - first
            ref_len += (int64_t)(int32_t)(uint32_t)l - (int64_t)(int32_t)(uint32_t)f;
        }
        S->ref_len_per_bam[b] = ref_len;
        if (covered < ref_len) covered = (uint32_t)ref_len;   // uint32_t _covered_ref_len (BamSummary.cpp:125)
    }
    S->covered_ref_len = covered;
    for (int l = 0; l < p.nlib; ++l)
        for (int f = 0; f < BDK_NUM_FLAGS; ++f) S->read_counts_by_flag[l][f] = acc.hist[l * BDK_NUM_FLAGS + f];
    int window = p.initial_window;
    density->assign(std::max(1, nkey_of(p)), 0.0f);
    for (int i = 0; i < p.nlib; ++i) {
        uint32_t n = S->lib_read_count[i];
        float covg = 0;
        if (n != 0 && covered != 0) covg = float(n) * p.libs[i].readlens / covered;
        S->seq_coverage[i] = covg;
        float dens = 0.000001f;
        if (p.cn_lib) { if (n != 0) dens = float(n) / covered; }
        else dens = float(S->read_count_per_bam[p.libs[i].bam_index]) / covered;
        (*density)[p.cn_lib ? i : p.libs[i].bam_index] = dens;
        S->read_density[i] = dens;
        int disc = S->read_counts_by_flag[i][BDK_ARP_LARGE_INSERT] + S->read_counts_by_flag[i][BDK_ARP_SMALL_INSERT];
        int tmp = (disc > 0) ? (float)covered / (float)disc : 50;
        window = std::min(window, tmp);
    }
    S->window = window;
}

}  // namespace bdk
