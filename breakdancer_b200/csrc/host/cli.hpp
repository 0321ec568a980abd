// Options of the breakdancer_max executable (reference src/lib/common/Options.hpp:12-71).
#pragma once
#include <string>
#include <vector>

#define BDK_CLI_VERSION "b200-0.1"

namespace bdh {

struct CliOptions {
    std::string chr, cache_file, restore_file, bam_config_path, prefix_fastq, dump_BED, stats_json;
    int min_len = 7, cut_sd = 3, max_sd = 1000000000, min_map_qual = 35, min_read_pair = 2, seq_coverage_lim = 1000,
        buffer_size = 100, score_threshold = 30;
    bool transchr_rearrange = false, fisher = false, Illumina_long_insert = false, CN_lib = false, print_AF = false;
    std::vector<std::string> orig_argv;
};

CliOptions parse_cli(int argc, char** argv);

}  // namespace bdh
