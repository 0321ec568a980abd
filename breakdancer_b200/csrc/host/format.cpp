// TSV header / row formatter of the drop-in executable.
// Mirrors the reference's output statements: header block BreakDancerMax.cpp:75-153, SV row
// BreakDancer.cpp:377-497 (including the sticky `cout << fixed << setprecision(2)` state that
// makes later allele-frequency columns print with two decimals).
#include "host.hpp"

#include <cmath>
#include <cstring>
#include <iomanip>
#include <map>
#include <ostream>
#include <sstream>

namespace bdh {

static const int kFlagValues[BDK_NUM_FLAGS] = {0, 1, 2, 3, 4, 8, 18, 20, 32, 64, 192};  // ReadFlags.cpp:4-14

const char* sv_type_name(int flag, bool long_insert) {  // Options.cpp:105-119
    switch (flag) {
        case BDK_ARP_FF: return "INV";
        case BDK_ARP_LARGE_INSERT: return long_insert ? "" : "DEL";
        case BDK_ARP_SMALL_INSERT: return "INS";
        case BDK_ARP_RF: return long_insert ? "DEL" : "ITX";
        case BDK_ARP_RR: return "INV";
        case BDK_ARP_CTX: return "CTX";
        default: return "";
    }
}

void format_header(std::ostream& out, const bdk_params& p, const bdk_summary_t& S,
                   const std::vector<std::string>& lib_names, const std::vector<std::string>& bam_names,
                   bool print_af) {
    out << "#Library Statistics:" << std::endl;
    for (int i = 0; i < p.nlib; ++i) {
        const bdk_lib& lc = p.libs[i];
        uint32_t covered = S.covered_ref_len;
        uint32_t n = S.lib_read_count[i];
        float physical_coverage = float(n * lc.mean_insertsize) / covered / 2;
        out << "#" << bam_names[lc.bam_index]
            << "\tmean:" << lc.mean_insertsize
            << "\tstd:" << lc.std_insertsize
            << "\tuppercutoff:" << lc.uppercutoff
            << "\tlowercutoff:" << lc.lowercutoff
            << "\treadlen:" << lc.readlens
            << "\tlibrary:" << lib_names[i]
            << "\treflen:" << covered
            << "\tseqcov:" << S.seq_coverage[i]
            << "\tphycov:" << physical_coverage;
        for (int j = 0; j < BDK_NUM_FLAGS; ++j) {
            uint32_t c = S.read_counts_by_flag[i][j];
            if (c) out << "\t" << kFlagValues[j] << ":" << c;
        }
        out << "\n";
    }
    out << "#Chr1\tPos1\tOrientation1\tChr2\tPos2\tOrientation2\tType\tSize\tScore\tnum_Reads\tnum_Reads_lib";
    if (print_af) out << "\tAllele_frequency";
    if (!p.cn_lib) {
        for (auto const& b : bam_names) {
            std::string::size_type t = b.rfind("/");
            out << "\t" << (t != std::string::npos ? b.substr(t + 1) : b);
        }
    }
    out << "\n";
}

// x86 produces the default NaN with the sign bit set for 0/0 (prints "-nan" through glibc); the
// device's canonical NaN is positive. The reference's only NaN source is that 0/0
// (SvBuilder.cpp:85-86 with no counts), so NaNs are printed the way the reference prints them.
static void put_float(std::ostream& out, float v) {
    if (std::isnan(v)) out << "-nan"; else out << v;
}

void format_rows(std::ostream& out, const bdk_params& p, const bdk_result& r,
                 const std::vector<std::string>& lib_names, const std::vector<std::string>& bam_names,
                 const std::vector<std::string>& tid_names, bool print_af) {
    int nkey = r.nkey;
    for (uint64_t k = 0; k < r.n_sv; ++k) {
        const bdk_sv& sv = r.sv[k];
        const int32_t* lc = r.lib_count + k * p.nlib;
        const uint32_t* cc = r.cn_count + k * nkey;
        const float* cn = r.copy_number + k * nkey;
        std::string sptype;
        if (p.cn_lib) {
            for (int l = 0; l < p.nlib; ++l) {
                if (!lc[l]) continue;
                std::string cns = "NA";
                if (sv.flag != BDK_ARP_CTX && cc[l]) {
                    std::stringstream ss;
                    ss << std::fixed;
                    ss << std::setprecision(2) << cn[l];
                    cns = ss.str();
                }
                if (!sptype.empty()) sptype += ":";
                sptype += lib_names[l] + "|" + std::to_string(lc[l]) + "," + cns;
            }
        } else {
            std::map<std::string, int> per_bam;
            for (int l = 0; l < p.nlib; ++l)
                if (lc[l]) per_bam[bam_names[p.libs[l].bam_index]] += lc[l];
            for (auto const& kv : per_bam) {
                if (!sptype.empty()) sptype += ":";
                sptype += kv.first + "|" + std::to_string(kv.second);
            }
            if (sptype.empty()) sptype = "NA";
        }
        auto tname = [&](int t) -> std::string { return t >= 0 && t < (int)tid_names.size() ? tid_names[t] : std::string(); };
        out << tname(sv.chr[0]) << "\t" << sv.pos[0] << "\t" << sv.fwd[0] << "+" << sv.rev[0] << "-"
            << "\t" << tname(sv.chr[1]) << "\t" << sv.pos[1] << "\t" << sv.fwd[1] << "+" << sv.rev[1] << "-"
            << "\t" << sv_type_name(sv.flag, p.illumina_long_insert) << "\t" << sv.diffspan << "\t" << sv.score
            << "\t" << sv.num_pairs << "\t" << sptype;
        if (print_af) { out << "\t"; put_float(out, sv.allele_frequency); }
        if (!p.cn_lib && sv.flag != BDK_ARP_CTX) {
            for (int b = 0; b < p.nbam; ++b) {
                if (!cc[b]) out << "\tNA";
                else { out << "\t"; out << std::fixed; out << std::setprecision(2); put_float(out, cn[b]); }
            }
        }
        out << "\n";
    }
}

}  // namespace bdh

extern "C" {

static std::vector<std::string> to_vec(const char* const* a, int n) {
    std::vector<std::string> v;
    for (int i = 0; i < n; ++i) v.push_back(a[i] ? a[i] : "");
    return v;
}

// names = { const char* const* lib_names, const char* const* bam_names }
int64_t bdh_format_header(const bdk_params* p, const bdk_summary_t* S, const void* names, int print_af,
                          char* buf, int64_t cap) {
    const char* const* const* nn = (const char* const* const*)names;
    std::ostringstream out;
    bdh::format_header(out, *p, *S, to_vec(nn[0], p->nlib), to_vec(nn[1], p->nbam), print_af != 0);
    std::string s = out.str();
    if ((int64_t)s.size() + 1 <= cap) memcpy(buf, s.c_str(), s.size() + 1);
    return (int64_t)s.size();
}

// sticky: in/out, non-zero once "fixed << setprecision(2)" has been applied to the stream
int64_t bdh_format_rows(const bdk_params* p, const bdk_result* r, const void* names, const void* tid_names,
                        int print_af, int* sticky, char* buf, int64_t cap) {
    const char* const* const* nn = (const char* const* const*)names;
    std::ostringstream out;
    if (sticky && *sticky) out << std::fixed << std::setprecision(2);
    bdh::format_rows(out, *p, *r, to_vec(nn[0], p->nlib), to_vec(nn[1], p->nbam),
                     to_vec((const char* const*)tid_names, p->ntid), print_af != 0);
    if (sticky) *sticky = (out.flags() & std::ios::fixed) ? 1 : 0;
    std::string s = out.str();
    if ((int64_t)s.size() + 1 <= cap) memcpy(buf, s.c_str(), s.size() + 1);
    return (int64_t)s.size();
}

}  // extern "C"
