// bam2cfg -- stands in for perl/bam2cfg.pl (same options, same output grammar): prints a BreakDancer configuration
// for the given position-sorted BAM files.
#include "../../../include/bdk_host.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

static void usage() {
    fprintf(stderr, "\nUsage:   bam2cfg <bam files>\nOptions:\n"
                    "         -q INT    Minimum mapping quality [35]\n"
                    "         -m        Using mapping quality instead of alternative mapping quality\n"
                    "         -s        Minimal mean insert size [50]\n"
                    "         -C        Change default system from Illumina to SOLiD\n"
                    "         -c FLOAT  Cutoff in unit of standard deviation [4]\n"
                    "         -n INT    Number of observation required to estimate mean and s.d. insert size [10000]\n"
                    "         -v FLOAT  Cutoff on coefficients of variation [1]\n"
                    "         -f STRING A two column tab-delimited text file (RG, LIB) specify the RG=>LIB mapping, useful when BAM header is incomplete\n"
                    "         -g        Output mapping flag distribution\n\n");
}

int main(int argc, char** argv) {
    bdh_bam2cfg_opts o;
    bdh_bam2cfg_defaults(&o);
    std::string fmap;
    int ch;
    while ((ch = getopt(argc, argv, "q:n:c:b:p:s:hmf:gCv:")) != -1) {
        switch (ch) {
            case 'q': o.min_mapq = atoi(optarg); break;
            case 'n': o.n_obs = atoi(optarg); break;
            case 'c': o.cut_sd = atof(optarg); break;
            case 's': o.min_mean = atof(optarg); break;
            case 'v': o.max_cv = atof(optarg); break;
            case 'm': o.use_mapq = 1; break;
            case 'C': o.solid = 1; break;
            case 'g': o.flag_hist = 1; break;
            case 'f': fmap = optarg; o.rg_lib_file = fmap.c_str(); break;
            case 'b': case 'p': break;                       // histogram options of -h
            case 'h': fprintf(stderr, "bam2cfg: -h (insert size histogram plots) is not supported\n"); return 1;
            default: usage(); return 1;
        }
    }
    if (optind >= argc) { usage(); return 1; }
    std::vector<const char*> bams(argv + optind, argv + argc);
    char err[512] = "";
    int64_t n = bdh_bam2cfg(bams.data(), (int)bams.size(), &o, nullptr, 0, err, sizeof err);
    if (n < 0) { fprintf(stderr, "%s\n", err); return 1; }
    std::vector<char> buf((size_t)n + 1);
    bdh_bam2cfg(bams.data(), (int)bams.size(), &o, buf.data(), (int64_t)buf.size(), err, sizeof err);
    fwrite(buf.data(), 1, (size_t)n, stdout);
    return 0;
}
