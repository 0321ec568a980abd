// Command line of the drop-in executable. The option letters, their defaults (note -y 30) and the order of the usage lines are
// the reference's contract (src/lib/common/Options.cpp:27-122); they are held in one table from which the getopt string, the
// parser and the usage text are generated. --stats-json FILE (ours) writes the run's timings as JSON.
#include "host.hpp"
#include "cli.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <getopt.h>
#include <stdexcept>

namespace bdh {

namespace {

enum Kind { INT, FLAG, STR };
struct OptSpec {
    char letter; Kind kind;
    int CliOptions::*i; bool CliOptions::*b; std::string CliOptions::*s;
    const char* arg;      // argument placeholder in the usage text, nullptr: not listed there (-f, -C, -R as upstream)
    const char* help;
};
const OptSpec kOpts[] = {
    {'o', STR, nullptr, nullptr, &CliOptions::chr, "STRING", "operate on a single chromosome [all chromosome]"},
    {'s', INT, &CliOptions::min_len, nullptr, nullptr, "INT", "minimum length of a region"},
    {'c', INT, &CliOptions::cut_sd, nullptr, nullptr, "INT", "cutoff in unit of standard deviation"},
    {'m', INT, &CliOptions::max_sd, nullptr, nullptr, "INT", "maximum SV size"},
    {'q', INT, &CliOptions::min_map_qual, nullptr, nullptr, "INT", "minimum alternative mapping quality"},
    {'r', INT, &CliOptions::min_read_pair, nullptr, nullptr, "INT", "minimum number of read pairs required to establish a connection"},
    {'x', INT, &CliOptions::seq_coverage_lim, nullptr, nullptr, "INT", "maximum threshold of haploid sequence coverage for regions to be ignored"},
    {'b', INT, &CliOptions::buffer_size, nullptr, nullptr, "INT", "buffer size for building connection"},
    {'t', FLAG, nullptr, &CliOptions::transchr_rearrange, nullptr, "", "only detect transchromosomal rearrangement, by default off"},
    {'f', FLAG, nullptr, &CliOptions::fisher, nullptr, nullptr, nullptr},
    {'d', STR, nullptr, nullptr, &CliOptions::prefix_fastq, "STRING", "prefix of fastq files that SV supporting reads will be saved by library"},
    {'g', STR, nullptr, nullptr, &CliOptions::dump_BED, "STRING", "dump SVs and supporting reads in BED format for GBrowse"},
    {'l', FLAG, nullptr, &CliOptions::Illumina_long_insert, nullptr, "", "analyze Illumina long insert (mate-pair) library"},
    {'a', FLAG, nullptr, &CliOptions::CN_lib, nullptr, "", "print out copy number and support reads per library rather than per bam, by default off"},
    {'h', FLAG, nullptr, &CliOptions::print_AF, nullptr, "", "print out Allele Frequency column, by default off"},
    {'y', INT, &CliOptions::score_threshold, nullptr, nullptr, "INT", "output score filter"},
    {'C', STR, nullptr, nullptr, &CliOptions::cache_file, nullptr, nullptr},
    {'R', STR, nullptr, nullptr, &CliOptions::restore_file, nullptr, nullptr},
};

void usage(const CliOptions& o) {
    fprintf(stderr, "\nbreakdancer-max (B200) version %s\n\nUsage: breakdancer-max <analysis.config>\n\nOptions: \n", BDK_CLI_VERSION);
    for (const OptSpec& s : kOpts) {
        if (!s.arg) continue;
        if (s.kind == INT) fprintf(stderr, "       -%c %-13s%s [%d]\n", s.letter, s.arg, s.help, o.*(s.i));
        else fprintf(stderr, "       -%c %-13s%s\n", s.letter, s.arg, s.help);
    }
    fprintf(stderr, "       --stats-json FILE  write the run's stage timings and throughput as JSON\n\n");
}

}  // namespace

CliOptions parse_cli(int argc, char** argv) {
    CliOptions o;
    o.orig_argv.assign(argv, argv + argc);
    std::string letters;
    for (const OptSpec& s : kOpts) { letters += s.letter; if (s.kind != FLAG) letters += ':'; }
    static const struct option longopts[] = {{"stats-json", required_argument, nullptr, 1000}, {nullptr, 0, nullptr, 0}};
    optind = 1;
    int c;
    while ((c = getopt_long(argc, argv, letters.c_str(), longopts, nullptr)) >= 0) {
        if (c == 1000) { o.stats_json = optarg; continue; }
        const OptSpec* spec = nullptr;
        for (const OptSpec& s : kOpts) if (s.letter == c) spec = &s;
        if (!spec) {
            fprintf(stderr, "Unrecognized option '-%c'.\n", c);
            exit(1);
        }
        if (c == 'R' && argc != 3) throw std::runtime_error("When using -R, no other options are allowed");
        if (spec->kind == INT) o.*(spec->i) = atoi(optarg);
        else if (spec->kind == FLAG) o.*(spec->b) = true;
        else o.*(spec->s) = optarg;
        if (c == 'R') return o;
    }
    if (optind == argc) {
        usage(o);
        exit(1);
    }
    o.bam_config_path = argv[optind];
    return o;
}

}  // namespace bdh
