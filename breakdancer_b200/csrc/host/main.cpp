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
        bdk_params p;
        memset(&p, 0, sizeof p);
        p.min_len = o.min_len; p.max_sd = o.max_sd; p.min_map_qual = o.min_map_qual; p.min_read_pair = o.min_read_pair;
        p.seq_coverage_lim = o.seq_coverage_lim; p.buffer_size = o.buffer_size; p.score_threshold = o.score_threshold;
        p.transchr_rearrange = o.transchr_rearrange; p.fisher = o.fisher; p.illumina_long_insert = o.Illumina_long_insert;
        p.cn_lib = o.CN_lib; p.chr_restricted = !o.chr.empty(); p.initial_window = cfg.window;
        p.nlib = (int)cfg.libs.size(); p.nbam = (int)cfg.bam_files.size();
        p.libs = cfg.libs.data();
        static const int32_t zero = 0;
        std::vector<std::string> tid_names;
        uint64_t n_records = 0;
        double t_decoded = 0, t_created = 0;
        bdk_bam_stats bstats;
        memset(&bstats, 0, sizeof bstats);
        bool on_device = false;
        // Up to 16 bams and no read dump: the files are decoded on the GPU (two: merged there in BamMerger's order; three or more: the
        // order from the priority queue on the host, the gather on the GPU) (bdk_push_bam: only the compressed bytes cross PCIe,
        // inflate / record parsing / classification of consecutive windows overlap). BDK_GPU_DECODE=0 keeps the host decoder. A file the
        // device path refuses (damaged member, truncated record) goes through the host decoder, which reports what is wrong.
        const char* gd = getenv("BDK_GPU_DECODE");
        if (cfg.bam_files.size() <= 16 && !want_reads && !(gd && atoi(gd) == 0)) {
            const size_t nb = cfg.bam_files.size();
            std::vector<bdh_bamdev*> devs;
            for (size_t b = 0; b < nb; ++b) {
                bdh_bamdev* d = b == 0 ? bdh_bamdev_open(&cfgh, cfg.bam_files[0].c_str(), o.chr.c_str(), err, sizeof err)
                                       : bdh_bamdev_open_next(&cfgh, devs.back(), cfg.bam_files[b].c_str(), o.chr.c_str(), err, sizeof err);
                if (!d) break;
                devs.push_back(d);
            }
            if (devs.size() == nb) {
                t_decoded = now_s();
                if (cuda_warmup.joinable()) cuda_warmup.join();
                std::vector<int32_t> rg_lib, rg_bam;      // the read-group ids of a bam follow those of the bam before it
                for (bdh_bamdev* d : devs) {
                    rg_lib.insert(rg_lib.end(), bdh_bamdev_rg_lib(d), bdh_bamdev_rg_lib(d) + bdh_bamdev_nrg(d));
                    rg_bam.insert(rg_bam.end(), bdh_bamdev_rg_bam(d), bdh_bamdev_rg_bam(d) + bdh_bamdev_nrg(d));
                }
                bdh_bamdev* dev = devs[0];
                p.nrg = (int)rg_lib.size(); p.ntid = std::max(1, bdh_bamdev_ntid(dev));
                p.rg_lib = rg_lib.data(); p.rg_bam = rg_bam.data();
                check(nullptr, bdk_create(&ctx, 0, &p), "bdk_create");      // (copies the tables)
                p.rg_lib = nullptr; p.rg_bam = nullptr;
                t_created = now_s();
                std::vector<bdk_bam_stats> stn(nb);
                memset(stn.data(), 0, nb * sizeof(bdk_bam_stats));
                const int rc = nb == 1 ? bdh_bamdev_push(dev, ctx, &stn[0]) : bdh_bamdev_pushn(devs.data(), (int)nb, ctx, stn.data());
                if (rc == 0) {
                    on_device = true;
                    bstats = stn[0];
                    for (size_t b = 1; b < nb; ++b) {
                        bstats.kept += stn[b].kept; bstats.records += stn[b].records; bstats.h2d_bytes += stn[b].h2d_bytes; bstats.inflated_bytes += stn[b].inflated_bytes;
                        bstats.windows += stn[b].windows; bstats.inflate_ms += stn[b].inflate_ms; bstats.sorted = bstats.sorted && stn[b].sorted;
                        bstats.chain_ms = stn[b].chain_ms; bstats.extract_ms = stn[b].extract_ms;       // (timers accumulate over the job)
                    }
                    n_records = bstats.kept;
                    for (int t = 0; t < bdh_bamdev_ntid(dev); ++t) tid_names.push_back(bdh_bamdev_tid_name(dev, t));      // BamMerger: the first stream's header
                    if (!bstats.sorted)
                        std::cerr << "WARNING: the input is not sorted by reference sequence and position; the covered reference length, the window and "
                                     "the regions assume a coordinate-sorted bam (samtools sort).\n";
                } else {
                    const std::string why = bdk_last_error(ctx);
                    if (getenv("BDK_DECODE_TRACE")) fprintf(stderr, "[decode] device decode refused the input (%s); host decoder\n", why.c_str());
                    bdk_destroy(ctx); ctx = nullptr;
                    if (rc != BDK_ERR_DATA) { for (bdh_bamdev* d : devs) bdh_bamdev_free(d); throw std::runtime_error("bdk_push_bam: " + why); }
                }
            } else if (getenv("BDK_DECODE_TRACE")) fprintf(stderr, "[decode] device decode: %s; host decoder\n", err);
            for (bdh_bamdev* d : devs) bdh_bamdev_free(d);
        }
        if (!on_device) {
            // pageable columns: pinning hundreds of megabytes costs more than the staged copy of a one-shot run saves, and the decoder
            // would have to wait for the CUDA context
            stream = bdh_stream_open(&cfgh, nullptr, 0, o.chr.c_str(), 0, getenv("BDK_CLI_PINNED") ? 1 : 0, want_reads ? 1 : 0, err, sizeof err);
            if (!stream) throw std::runtime_error(err);
            if (!bdh_stream_sorted(stream))
                std::cerr << "WARNING: the input is not sorted by reference sequence and position; the covered reference length, the window and "
                             "the regions assume a coordinate-sorted bam (samtools sort).\n";
            t_decoded = now_s();
            if (cuda_warmup.joinable()) cuda_warmup.join();
            p.nrg = std::max(1, bdh_stream_nrg(stream)); p.ntid = std::max(1, bdh_stream_ntid(stream));
            p.rg_lib = bdh_stream_nrg(stream) ? bdh_stream_rg_lib(stream) : &zero;
            p.rg_bam = bdh_stream_nrg(stream) ? bdh_stream_rg_bam(stream) : &zero;
            check(nullptr, bdk_create(&ctx, 0, &p), "bdk_create");
            bdk_soa cols;
            bdh_stream_cols(stream, &cols);
            t_created = now_s();
            n_records = bdh_stream_n(stream);
            check(ctx, bdk_push(ctx, &cols, n_records), "bdk_push");
            for (int t = 0; t < bdh_stream_ntid(stream); ++t) tid_names.push_back(bdh_stream_tid_name(stream, t));
        }
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
        format_rows(std::cout, p, res, cfg.lib_names, cfg.bam_files, tid_names, o.print_AF);

        if (want_reads) {
            std::unique_ptr<std::ofstream> bed;
            if (!o.dump_BED.empty()) bed.reset(new std::ofstream(o.dump_BED.c_str()));
            write_support_reads(ctx, stream, p, res, cfg.lib_names, tid_names, bed.get(), o.prefix_fastq);
        }
        if (!o.stats_json.empty()) {        // SURVEY section 5, metrics row
            std::cout.flush();
            const double t_end = now_s();
            const uint64_t n = n_records;
            double stage[3] = {0, 0, 0};
            if (stream) bdh_stream_timings(stream, &stage[0], &stage[1], &stage[2]);
            std::ofstream js(o.stats_json.c_str());
            js << "{\"records\": " << n << ", \"read_pairs\": " << n / 2 << ", \"sv_calls\": " << res.n_sv
               << ", \"anomalous_reads\": " << S.n_anomalous << ", \"total_s\": " << t_end - t_start
               << ", \"decode_s\": " << t_decoded - t_start << ", \"decode_inflate_s\": " << stage[0] << ", \"decode_extract_s\": " << stage[1]
               << ", \"decode_merge_s\": " << stage[2] << ", \"context_s\": " << t_created - t_decoded << ", \"push_s\": " << t_pushed - t_created
               << ", \"finish_s\": " << t_finished - t_pushed << ", \"output_s\": " << t_end - t_finished
               << ", \"device_decode\": " << (on_device ? 1 : 0) << ", \"device_inflate_ms\": " << bstats.inflate_ms << ", \"device_chain_ms\": " << bstats.chain_ms
               << ", \"device_extract_ms\": " << bstats.extract_ms << ", \"device_windows\": " << bstats.windows << ", \"inflated_bytes\": " << bstats.inflated_bytes
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
