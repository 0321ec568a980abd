/*
 * bdk_host.h -- C ABI of the host side of the hot path: bam2cfg config parsing, BAM -> pinned
 * struct-of-arrays decoding (the producer of bdk_soa), the k-way merge of several BAMs, the
 * BAM writer used by the synthetic-data tools, and the TSV formatter.  It mirrors the
 * reference's L2 "io" layer for this path; each entry point cites
 * the reference code it stands in for (paths relative to the reference root).
 *
 * These helpers are plain host code (no GPU needed) and live in the same shared library as
 * the bdk_* kernels entry points.
 */
#ifndef BDK_HOST_H
#define BDK_HOST_H

#include <stdint.h>
#include "bdk.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- bam2cfg configuration file: BamConfig / BamConfigEntry --------------------------------
 * src/lib/io/BamConfig.cpp:19-122, BamConfigEntry.cpp:31-86.  Library index = rank of the
 * library name; bam list = sorted unique "map:" paths. */
typedef struct bdh_config bdh_config;
bdh_config* bdh_config_parse(const char* text, int cut_sd, char* err, int errcap);
bdh_config* bdh_config_load(const char* path, int cut_sd, char* err, int errcap);
void bdh_config_free(bdh_config* c);
int bdh_config_nlib(const bdh_config* c);
int bdh_config_nbam(const bdh_config* c);
int bdh_config_window(const bdh_config* c);                    /* max_read_window_size() */
const bdk_lib* bdh_config_libs(const bdh_config* c);           /* [nlib] */
const char* bdh_config_lib_name(const bdh_config* c, int i);
const char* bdh_config_bam_name(const bdh_config* c, int i);
/* readgroup_library(rg) -> library index (BamConfig.hpp:63-72): unknown read groups fall back
 * to the library of the first bam; returns -1 when that library name is empty. */
int bdh_config_rg_lib(const bdh_config* c, const char* rg);

/* ---- decoded, merged record stream --------------------------------------------------------
 * openBams + BamMerger + AlignmentSource::next (src/lib/io/BamIo.cpp:6-31, BamMerger.cpp:40-126,
 * AlignmentSource.hpp:48-65, Alignment.cpp:12-64): every primary record with tid >= 0 of every
 * bam in the config (optionally limited to a region, RegionLimitedBamReader.hpp:36-71), merged
 * by (tid, pos, strand), as struct-of-arrays columns. */
typedef struct bdh_stream bdh_stream;
/* paths == NULL: open the config's bam files (relative to the current directory). threads <= 0:
 * all cores. pinned != 0: allocate the columns with cudaHostAlloc. keep_records != 0: keep the
 * raw records so names / sequences can be fetched afterwards (-g / -d). */
bdh_stream* bdh_stream_open(const bdh_config* cfg, const char* const* paths, int npaths,
                            const char* region, int threads, int pinned, int keep_records,
                            char* err, int errcap);
void bdh_stream_free(bdh_stream* s);
uint64_t bdh_stream_n(const bdh_stream* s);
void bdh_stream_cols(const bdh_stream* s, bdk_soa* out);
int bdh_stream_nrg(const bdh_stream* s);
const int32_t* bdh_stream_rg_lib(const bdh_stream* s);
const int32_t* bdh_stream_rg_bam(const bdh_stream* s);
int bdh_stream_ntid(const bdh_stream* s);
const char* bdh_stream_tid_name(const bdh_stream* s, int tid);
/* Raw record access (keep_records): query name, and FASTQ text of record i
 * (Alignment::to_fastq, src/lib/io/Alignment.cpp:66-84). Returns bytes written or -1. */
const char* bdh_stream_qname(const bdh_stream* s, uint64_t i);
int bdh_stream_fastq(const bdh_stream* s, uint64_t i, char* buf, int cap);
/* 1 if the merged stream is ordered by (reference sequence, position), as BamMerger delivers sorted bams (BamMerger.cpp:40-61);
 * 0: an input bam is not coordinate-sorted -- the covered reference length, the window and the regions are then not what a
 * sorted file would give (the reference does not check either). */
int bdh_stream_sorted(const bdh_stream* s);
/* seconds spent in (inflate, parse+extract, merge) by the last open */
void bdh_stream_timings(const bdh_stream* s, double* inflate_s, double* extract_s, double* merge_s);
/* What `<bam>.bai` (or `<name>.bai`) says about every reference sequence, without touching the bam: records[t] = mapped + unmapped
 * records placed on sequence t (samtools' pseudo-bin; -1 if the index has none), bytes[t] = compressed bytes its chunks span.
 * Returns the number of sequences in the index (fills at most `cap`), -1 if there is no usable index, -2 on a damaged one
 * (message in err). The planner of per-chromosome shards uses it (breakdancer_b200/shard.py); the reader itself uses the same
 * index for -o regions (RegionLimitedBamReader.hpp:40-66 in the reference). */
int bdh_bai_reference_stats(const char* bam_path, int64_t* records, int64_t* bytes, int cap, char* err, int errcap);
/* Process-wide counters of the BGZF stage: members the host's table-driven decoder handed to zlib, and (BDK_GPU_INFLATE=1)
 * members the GPU decoder refused or got wrong and the host decoded again. */
void bdh_inflate_counters(uint64_t* host_fallbacks, uint64_t* gpu_redone);

/* ---- device-resident decode: the host's share ------------------------------------------------------------------------------
 * bdk_push_bam (include/bdk.h) inflates, parses and classifies a BAM file on the GPU. What stays on the host is opening the
 * file: mapping it, listing its BGZF members (with a region and a .bai only those of the reference sequence,
 * RegionLimitedBamReader.hpp:40-66), inflating the few members of the BAM header and numbering the read groups -- the config's
 * read groups in its own order, then one id for every other read-group string (BamConfig.hpp:63-72: unknown read groups get the
 * first bam's library). path NULL/"" = the config's only bam. Order of use: bdh_bamdev_open, bdk_create with nrg / rg_lib /
 * rg_bam / ntid from here, bdh_bamdev_push, bdk_summary / bdk_finish. The records pushed are those bdh_stream_open delivers for
 * the same file and region, in the same order. One file per call; several bams (BamMerger): bdh_bamdev_open_next and
 * bdh_bamdev_push2 / bdh_bamdev_pushn below (or the host decoder, bdh_stream_open). */
typedef struct bdh_bamdev bdh_bamdev;
bdh_bamdev* bdh_bamdev_open(const bdh_config* cfg, const char* path, const char* region, char* err, int errcap);
/* The NEXT bam of a config with several (tumor / normal, one bam per lane ...): its read-group ids follow those of the bam opened
 * before it (`first`: the config's first bam for the second one, the second for the third ...), so the sources can be handed to
 * bdk_push_bams together. bdk_create then takes nrg = the sum of bdh_bamdev_nrg over the bams and their rg_lib / rg_bam arrays one
 * after the other, in the order of the config's (sorted) bam list. */
bdh_bamdev* bdh_bamdev_open_next(const bdh_config* cfg, const bdh_bamdev* first, const char* path, const char* region, char* err, int errcap);
void bdh_bamdev_free(bdh_bamdev* d);
int bdh_bamdev_nrg(const bdh_bamdev* d);
const int32_t* bdh_bamdev_rg_lib(const bdh_bamdev* d);
const int32_t* bdh_bamdev_rg_bam(const bdh_bamdev* d);
int bdh_bamdev_ntid(const bdh_bamdev* d);
const char* bdh_bamdev_tid_name(const bdh_bamdev* d, int tid);
uint64_t bdh_bamdev_members(const bdh_bamdev* d);
uint64_t bdh_bamdev_file_bytes(const bdh_bamdev* d);
int bdh_bamdev_push(bdh_bamdev* d, bdk_ctx* ctx, bdk_bam_stats* stats);
/* bdk_push_bams / bdk_decode_bams of two opened files: decoded on the device, merged there in BamMerger's order, classified (or the
 * merged columns copied to the host). stats2 = two entries or NULL. */
int bdh_bamdev_push2(bdh_bamdev* first, bdh_bamdev* second, bdk_ctx* ctx, bdk_bam_stats* stats2);
int bdh_bamdev_decode2(bdh_bamdev* first, bdh_bamdev* second, bdk_ctx* ctx, const bdk_soa* host_out, uint64_t cap, bdk_bam_stats* stats2);
/* The same for n opened files in the config's order (n = 1, 2, or 3 .. 16: for three or more the merge order is computed by the
 * reference's priority queue on the host from keys the device hands back, BamMerger.cpp:40-126, and the columns are gathered
 * through it on the device). stats = n entries or NULL. */
int bdh_bamdev_pushn(bdh_bamdev* const* devs, int n, bdk_ctx* ctx, bdk_bam_stats* stats);
int bdh_bamdev_decoden(bdh_bamdev* const* devs, int n, bdk_ctx* ctx, const bdk_soa* host_out, uint64_t cap, bdk_bam_stats* stats);
/* bdk_decode_bam of the opened file: the decoded columns into the caller's host arrays (cap records each). */
int bdh_bamdev_decode(bdh_bamdev* d, bdk_ctx* ctx, const bdk_soa* host_out, uint64_t cap, bdk_bam_stats* stats);

/* ---- BAM writer for synthetic inputs (stands in for samtools' bam_write1) ------------------
 * Writes n records from struct-of-arrays columns as a BGZF-compressed BAM with query names
 * "<prefix><qid>", one RG:Z tag per record (rg_names[rgid]), AM:i = mapq when write_am != 0,
 * CIGAR <qlen>M, and a deterministic base/quality pattern. */
int bdh_write_bam(const char* path, int ntid, const char* const* tid_names, const uint32_t* tid_lens,
                  int nrg, const char* const* rg_names, const bdk_soa* cols, uint64_t n,
                  const char* name_prefix, int write_am, int level, int threads, char* err, int errcap);

/* ---- bam2cfg: the configuration file from the BAMs themselves ------------------------------
 * perl/bam2cfg.pl:48-247 (+ perl/AlnParser.pm:31-124, Shapiro-Wilk figure bam2cfg.pl:284-744): per library the
 * insert-size mean / s.d. / lower / upper from the first n_obs proper FR pairs with mapping quality > min_mapq,
 * mean read length; one line per read group in the grammar bdh_config_parse reads. Options are the script's
 * (-q -n -c -s -v -m -C -g -f); -h (histogram plots) is not carried over. Lines come in @RG header order (the script
 * prints in Perl hash order). Returns the text length (stored if it fits in cap, with NUL), -1 on error. */
typedef struct bdh_bam2cfg_opts {
    int32_t min_mapq;          /* -q 35 */
    int32_t n_obs;             /* -n 10000 */
    double cut_sd;             /* -c 4 */
    double min_mean;           /* -s 50 */
    double max_cv;             /* -v 1 */
    int32_t use_mapq;          /* -m: MAPQ instead of the Aq / AM tag */
    int32_t solid;             /* -C */
    int32_t flag_hist;         /* -g */
    const char* rg_lib_file;   /* -f: two-column RG -> LIB table, or NULL */
} bdh_bam2cfg_opts;
void bdh_bam2cfg_defaults(bdh_bam2cfg_opts* o);
int64_t bdh_bam2cfg(const char* const* bam_paths, int nbam, const bdh_bam2cfg_opts* opts, char* buf, int64_t cap,
                    char* err, int errcap);

/* ---- output --------------------------------------------------------------------------------
 * Text of the reference's stdout: the "#Library Statistics" header block
 * (src/exe/breakdancer-max/BreakDancerMax.cpp:82-153) and the SV rows
 * (src/lib/breakdancer/BreakDancer.cpp:377-497). names = { lib_names, bam_names } (two arrays of
 * C strings); sticky carries the stream's "fixed, precision 2" state between calls. Both return
 * the text length; the text is stored only if it fits in cap (with NUL). */
int64_t bdh_format_header(const bdk_params* p, const bdk_summary_t* s, const void* names, int print_af,
                          char* buf, int64_t cap);
int64_t bdh_format_rows(const bdk_params* p, const bdk_result* r, const void* names, const void* tid_names,
                        int print_af, int* sticky, char* buf, int64_t cap);
/* BamConfigEntry::translate_token (src/lib/io/BamConfigEntry.cpp:31-59): Field ordinal, 10 = UNKNOWN */
int bdh_config_translate_token(const char* key);

#ifdef __cplusplus
}
#endif
#endif /* BDK_HOST_H */
