// Host-side (CPU) pieces of the hot path: config table, BAM decode to struct-of-arrays columns,
// BAM writer for synthetic data, TSV formatter. See include/bdk_host.h for the C ABI.
#pragma once
#include "../../../include/bdk.h"
#include "../../../include/bdk_host.h"

#include <cstdint>
#include <functional>
#include <istream>
#include <ostream>
#include <map>
#include <string>
#include <vector>

namespace bdh {

// BamConfig (reference src/lib/io/BamConfig.hpp:15-48)
struct Config {
    std::vector<bdk_lib> libs;                 // index = rank of library name
    std::vector<std::string> lib_names;
    std::vector<std::string> bam_files;        // sorted unique "map:" values
    std::map<std::string, std::string> readgroup_library;
    std::map<std::string, int> lib_index;
    std::string first_bam_library;             // _bam_library.begin()->second
    int window = 100000000;                    // max_read_window_size()

    static Config parse(std::istream& in, int cut_sd);
    int rg_lib(const std::string& rg) const;   // readgroup_library() -> library index, -1 if none
};


// run fn(i) for i in [0, n) on `threads` std::threads (dynamic chunks of `grain`)
void parallel_for(uint64_t n, uint64_t grain, int threads, const std::function<void(uint64_t, uint64_t)>& fn);
int default_threads();

const char* sv_type_name(int flag, bool long_insert);
void format_header(std::ostream& out, const bdk_params& p, const bdk_summary_t& S,
                   const std::vector<std::string>& lib_names, const std::vector<std::string>& bam_names,
                   bool print_af);
void format_rows(std::ostream& out, const bdk_params& p, const bdk_result& r,
                 const std::vector<std::string>& lib_names, const std::vector<std::string>& bam_names,
                 const std::vector<std::string>& tid_names, bool print_af);

void write_support_reads(bdk_ctx* ctx, const bdh_stream* stream, const bdk_params& p, const bdk_result& res,
                         const std::vector<std::string>& lib_names, const std::vector<std::string>& tid_names,
                         std::ostream* bed, const std::string& fastq_prefix);

}  // namespace bdh

struct bdh_config { bdh::Config cfg; };
