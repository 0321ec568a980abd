// breakdancer_max -- drop-in for the reference's breakdancer-max executable
// (src/exe/breakdancer-max/BreakDancerMax.cpp:38-163): same command line, same bam2cfg config
// file, same stdout. Host side: parse options + config, decode/merge the BAMs into pinned
// struct-of-arrays columns (all cores), hand them to the bdk context (GPU), print.
#include "host.hpp"
#include "cli.hpp"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <unistd.h>

using namespace bdh;

static void check(bdk_ctx* ctx, int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + bdk_last_error(ctx));
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    bdk_ctx* ctx = nullptr;
    bdh_stream* stream = nullptr;
    int rv = 0;
    const double t_start = now_s();
    std::thread cuda_warmup;
    try {
        CliOptions o = parse_cli(argc, argv);
        // the CUDA context takes a few hundred milliseconds to come up: let that happen while the BAMs are being inflated
        cuda_warmup = std::thread([] { bdk_host_free(bdk_host_alloc(64)); });
        if (!o.restore_file.empty() || !o.cache_file.empty())
            throw std::runtime_error("-C/-R (boost XML summary cache) are not supported by this build");
        bdh_config cfgh;
        {
            std::ifstream in(o.bam_config_path.c_str());
            cfgh.cfg = Config::parse(in, o.cut_sd);
        }
        const Config& cfg = cfgh.cfg;
        if (cfg.bam_files.empty()) {
            std::cout << "Error: no bams files in config file!\n";
            cuda_warmup.join();
            return 1;
        }
        const bool want_reads = !o.prefix_fastq.empty() || !o.dump_BED.empty();
        char err[512] = {0};
        // pageable columns: pinning hundreds of megabytes costs more than the staged copy of a one-shot run saves, and the decoder
        // would have to wait for the CUDA context
        stream = bdh_stream_open(&cfgh, nullptr, 0, o.chr.c_str(), 0, getenv("BDK_CLI_PINNED") ? 1 : 0, want_reads ? 1 : 0, err, sizeof err);
        if (!stream) throw std::runtime_error(err);
        if (!bdh_stream_sorted(stream))
            std::cerr << "WARNING: the input is not sorted by reference sequence and position; the covered reference length, the window and "
                         "the regions assume a coordinate-sorted bam (samtools sort).\n";
        const double t_decoded = now_s();
        if (cuda_warmup.joinable()) cuda_warmup.join();

        bdk_params p;
        memset(&p, 0, sizeof p);
        p.min_len = o.min_len; p.max_sd = o.max_sd; p.min_map_qual = o.min_map_qual; p.min_read_pair = o.min_read_pair;
        p.seq_coverage_lim = o.seq_coverage_lim; p.buffer_size = o.buffer_size; p.score_threshold = o.score_threshold;
        p.transchr_rearrange = o.transchr_rearrange; p.fisher = o.fisher; p.illumina_long_insert = o.Illumina_long_insert;
        p.cn_lib = o.CN_lib; p.chr_restricted = !o.chr.empty(); p.initial_window = cfg.window;
        p.nlib = (int)cfg.libs.size(); p.nbam = (int)cfg.bam_files.size();
        p.nrg = std::max(1, bdh_stream_nrg(stream)); p.ntid = std::max(1, bdh_stream_ntid(stream));
        p.libs = cfg.libs.data();
        static const int32_t zero = 0;
        p.rg_lib = bdh_stream_nrg(stream) ? bdh_stream_rg_lib(stream) : &zero;
        p.rg_bam = bdh_stream_nrg(stream) ? bdh_stream_rg_bam(stream) : &zero;
        check(nullptr, bdk_create(&ctx, 0, &p), "bdk_create");

        bdk_soa cols;
        bdh_stream_cols(stream, &cols);
        const double t_created = now_s();
        check(ctx, bdk_push(ctx, &cols, bdh_stream_n(stream)), "bdk_push");
        bdk_summary_t S;
        check(ctx, bdk_summary(ctx, &S), "bdk_summary");
        for (int b = 0; b < p.nbam; ++b)
            if (S.ref_len_per_bam[b] == 0)
                std::cerr << "Input file " << cfg.bam_files[b] << (o.chr.empty() ? "" : " (region: " + o.chr + ")")
                          << " does not contain legitimate paired end alignment. Please check that you have the correct paths"
                             " and the map/bam files are properly formated and indexed.\n";

        std::cout << "#Software: " << BDK_CLI_VERSION << " (commit " << bdk_version() << ")" << std::endl;
        std::cout << "#Command: ";
        for (auto const& a : o.orig_argv) std::cout << a << " ";
        std::cout << std::endl;
        format_header(std::cout, p, S, cfg.lib_names, cfg.bam_files, o.print_AF);

        const double t_pushed = now_s();
        bdk_result res;
        check(ctx, bdk_finish(ctx, &res), "bdk_finish");
        const double t_finished = now_s();
        if (uint32_t nd = bdk_duplicate_names(ctx))
            std::cerr << "WARNING: " << nd << " anomalous read(s) share their name with two or more others (bams with overlapping read names?); "
                         "the reads of such names are left unpaired.\n";
        std::vector<std::string> tid_names;
        for (int t = 0; t < bdh_stream_ntid(stream); ++t) tid_names.push_back(bdh_stream_tid_name(stream, t));
        format_rows(std::cout, p, res, cfg.lib_names, cfg.bam_files, tid_names, o.print_AF);

        if (want_reads) {
            std::unique_ptr<std::ofstream> bed;
            if (!o.dump_BED.empty()) bed.reset(new std::ofstream(o.dump_BED.c_str()));
            write_support_reads(ctx, stream, p, res, cfg.lib_names, tid_names, bed.get(), o.prefix_fastq);
        }
        if (!o.stats_json.empty()) {        // SURVEY section 5, metrics row
            std::cout.flush();
            const double t_end = now_s();
            const uint64_t n = bdh_stream_n(stream);
            double stage[3] = {0, 0, 0};
            bdh_stream_timings(stream, &stage[0], &stage[1], &stage[2]);
            std::ofstream js(o.stats_json.c_str());
            js << "{\"records\": " << n << ", \"read_pairs\": " << n / 2 << ", \"sv_calls\": " << res.n_sv
               << ", \"anomalous_reads\": " << S.n_anomalous << ", \"total_s\": " << t_end - t_start
               << ", \"decode_s\": " << t_decoded - t_start << ", \"decode_inflate_s\": " << stage[0] << ", \"decode_extract_s\": " << stage[1]
               << ", \"decode_merge_s\": " << stage[2] << ", \"context_s\": " << t_created - t_decoded << ", \"push_s\": " << t_pushed - t_created
               << ", \"finish_s\": " << t_finished - t_pushed << ", \"output_s\": " << t_end - t_finished
               << ", \"h2d_bytes\": " << bdk_h2d_bytes(ctx) << ", \"d2h_bytes\": " << bdk_d2h_bytes(ctx)
               << ", \"read_pairs_per_s\": " << (double)(n / 2) / (t_end - t_start) << "}\n";
        }
    } catch (std::exception const& e) {
        std::cerr << "ERROR: " << e.what() << "\n";
        rv = 1;
    }
    if (cuda_warmup.joinable()) cuda_warmup.join();
    if (!getenv("BDK_CLEAN_EXIT")) {      // everything is written: leave without unmapping buffers and tearing the CUDA context down piece by piece
        std::cout.flush(); std::cerr.flush();
        fflush(nullptr);
        _exit(rv);
    }
    if (ctx) bdk_destroy(ctx);
    if (stream) bdh_stream_free(stream);
    return rv;
}
