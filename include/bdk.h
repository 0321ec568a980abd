/*
 * bdk.h -- C ABI of the B200-native BreakDancerMax hot path ("bdk" = BreakDancer kernels).
 *
 * The reference (genome/breakdancer) has no FFI of its own for this path: the hot path is the
 * in-process call chain  main -> BamSummary::_analyze_bam -> BreakDancer::run -> push_read ->
 * process_breakpoint -> build_connection -> process_sv  (SURVEY.md section 8a).  This header is
 * the seam a maintainer would cut there: the reference's reader keeps producing records, they
 * are laid out as position-sorted struct-of-arrays columns, and everything from
 * "classify a record" to "one scored SV row" happens behind these entry points on the GPU.
 * Each declaration cites the reference code it replaces (paths relative to the reference
 * root, lines as of commit 4e44b43).
 *
 * Conventions: plain C, no C++/torch types; every function returns 0 on success or a negative
 * bdk_status; bdk_last_error() gives the message.  One context per GPU, calls on one context
 * are serialised by the caller.  Inputs must be sorted by (tid, pos) the way the reference's
 * BamMerger (src/lib/io/BamMerger.cpp:40-61) delivers them, contain only primary records with
 * tid >= 0 (src/lib/io/BamIo.cpp:11-18); a read name key (qid) is expected at most twice (see bdk_duplicate_names).
 */
#ifndef BDK_H
#define BDK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ReadFlag enum values, src/lib/common/ReadFlags.hpp:14-27 */
enum {
    BDK_NA = 0, BDK_ARP_FF = 1, BDK_ARP_LARGE_INSERT = 2, BDK_ARP_SMALL_INSERT = 3,
    BDK_ARP_RF = 4, BDK_ARP_RR = 5, BDK_NORMAL_FR = 6, BDK_NORMAL_RF = 7, BDK_ARP_CTX = 8,
    BDK_MATE_UNMAPPED = 9, BDK_UNMAPPED = 10, BDK_NUM_FLAGS = 11
};

enum bdk_status {
    BDK_OK = 0,
    BDK_ERR_ARG = -1,        /* bad argument / unsupported option value */
    BDK_ERR_CUDA = -2,       /* CUDA runtime error (message has the cudaError string) */
    BDK_ERR_NOMEM = -3,
    BDK_ERR_STATE = -4,      /* call order violated */
    BDK_ERR_DATA = -5,       /* input violates the contract (unknown rgid, library-less read group) */
    BDK_ERR_NCCL = -6
};

/* One library: LibraryConfig, src/lib/io/LibraryConfig.hpp:11-27 (index = position in the array,
 * i.e. rank of the library name, src/lib/io/BamConfig.cpp:97-101). */
typedef struct bdk_lib {
    float mean_insertsize;
    float std_insertsize;
    float uppercutoff;
    float lowercutoff;
    float readlens;
    int32_t min_mapping_quality;   /* -1: use bdk_params.min_map_qual (BreakDancer.cpp:155-156) */
    int32_t bam_index;             /* LibraryConfig::bam_file_index (sorted bam list) */
} bdk_lib;

/* Options (src/lib/common/Options.hpp:12-71) + the config-derived tables the kernels need. */
typedef struct bdk_params {
    int32_t min_len;              /* -s */
    int32_t max_sd;               /* -m */
    int32_t min_map_qual;         /* -q */
    int32_t min_read_pair;        /* -r  (must be >= 1) */
    int32_t seq_coverage_lim;     /* -x */
    int32_t buffer_size;          /* -b */
    int32_t score_threshold;      /* -y */
    int32_t transchr_rearrange;   /* -t */
    int32_t fisher;               /* -f */
    int32_t illumina_long_insert; /* -l */
    int32_t cn_lib;               /* -a */
    int32_t chr_restricted;       /* 1 iff -o was given (ReadRegionData.cpp:118,133) */
    int32_t initial_window;       /* BamConfig::max_read_window_size(), BamConfig.cpp:92-93,121 */
    int32_t nlib;                 /* <= BDK_MAX_LIBS */
    int32_t nbam;                 /* <= BDK_MAX_BAMS */
    int32_t nrg;                  /* number of read-group ids, <= 65536 */
    int32_t ntid;                 /* number of reference sequences (bam_header_t::n_targets) */
    const bdk_lib* libs;          /* [nlib] */
    const int32_t* rg_lib;        /* [nrg] read-group id -> library index (AlignmentSource.hpp:57-63);
                                     -1 = read group without a library: the reference throws
                                     "library index out of range" (BamConfig.hpp:51-55) -> BDK_ERR_DATA */
    const int32_t* rg_bam;        /* [nrg] read-group id -> index of the BAM the record was read from
                                     (BamSummary.cpp:120: counts are per reader) */
} bdk_params;

#define BDK_MAX_LIBS 255
#define BDK_MAX_BAMS 64

/* Position-sorted struct-of-arrays record columns: what Alignment's constructor extracts from a
 * bam1_t (src/lib/io/Alignment.cpp:45-64).  25 bytes per record in the hot columns; qlen and qid
 * are only read for anomalous records. */
typedef struct bdk_soa {
    const int32_t* pos;     /* core.pos   */
    const int32_t* mpos;    /* core.mpos  */
    const int32_t* tid;     /* core.tid   */
    const int32_t* mtid;    /* core.mtid  */
    const int32_t* isize;   /* core.isize (signed; abs() taken on the device) */
    const uint16_t* flag;   /* core.flag  */
    const uint8_t* mapq;    /* bdqual: AM aux tag if present else core.qual (Alignment.cpp:12-23) */
    const uint16_t* rgid;   /* read-group id assigned by the decoder (index into rg_lib / rg_bam) */
    const int32_t* qlen;    /* core.l_qseq */
    const uint64_t* qid;    /* read-name key: equal iff the query names are equal */
} bdk_soa;

/* BamSummary (src/lib/io/BamSummary.cpp:47-150) + the window/density block of main()
 * (src/exe/breakdancer-max/BreakDancerMax.cpp:83-116). */
typedef struct bdk_summary_t {
    uint32_t covered_ref_len;                               /* BamSummary.cpp:125-126 */
    int32_t window;                                         /* final _max_read_window_size */
    uint64_t n_records;
    uint64_t n_anomalous;                                   /* records entering regions */
    uint32_t read_count_per_bam[BDK_MAX_BAMS];              /* BamSummary.cpp:120 */
    uint64_t ref_len_per_bam[BDK_MAX_BAMS];                 /* size_t ref_len, BamSummary.cpp:56 */
    uint32_t lib_read_count[BDK_MAX_LIBS];                  /* BamSummary.cpp:84 */
    uint32_t read_counts_by_flag[BDK_MAX_LIBS][BDK_NUM_FLAGS]; /* BamSummary.cpp:113 */
    float seq_coverage[BDK_MAX_LIBS];                       /* BamSummary.cpp:140-145 */
    float read_density[BDK_MAX_LIBS];                       /* keyed by library index; -a off: the
                                                               density of the library's bam */
} bdk_summary_t;

/* One SV call = one process_sv() that reached the output statement
 * (src/lib/breakdancer/BreakDancer.cpp:348-512, SvBuilder.cpp:18-118). */
typedef struct bdk_sv {
    int32_t chr[2];         /* tid of the two breakpoints */
    int32_t pos[2];         /* 1-based, after padding (BreakDancer.cpp:387-388,463-464) */
    int32_t fwd[2];         /* region fwd_read_count */
    int32_t rev[2];
    int32_t flag;           /* ReadFlag of the call (SV type = Options::SVtype[flag]) */
    int32_t diffspan;       /* Size column */
    int32_t score;          /* PhredQ */
    int32_t num_pairs;      /* flag_counts[flag] */
    double logp;            /* ComputeProbScore result (natural log) */
    float allele_frequency; /* SvBuilder::allele_frequency */
    uint32_t cn_present;    /* reserved */
    int32_t region[2];      /* region indices (snodes); region[1] = -1 for a one-region call */
    int32_t window;         /* flush window the call was made in */
    int32_t order;          /* position in the reference's output order */
} bdk_sv;

typedef struct bdk_result {
    uint64_t n_sv;
    const bdk_sv* sv;              /* [n_sv], in the reference's output order */
    const int32_t* lib_count;      /* [n_sv][nlib] type_library_readcount[flag][lib] (0 = absent) */
    const uint32_t* cn_count;      /* [n_sv][nkey] proper-pair reads between the regions per key
                                      (key = library if cn_lib else bam); 0 = key absent */
    const float* copy_number;      /* [n_sv][nkey] SvBuilder::copy_number (valid where cn_count>0) */
    int32_t nkey;
} bdk_result;

/* Registered region = BasicRegion (src/lib/breakdancer/BasicRegion.hpp:24-45) plus the flush
 * window it was registered in. Debug/inspection view (BD_DUMP_REGION_SUMMARY equivalent). */
typedef struct bdk_region {
    int32_t tid, start, end;
    int32_t fwd, rev;
    int32_t first_read, n_reads;   /* range in the anomalous-read stream */
    int32_t stored;                /* reads kept (ReadRegionData.cpp:118-121) */
    int32_t window;
} bdk_region;

/* One anomalous read as it enters push_read()'s region builder (BreakDancer.cpp:209-241). */
typedef struct bdk_aread {
    int32_t pos, tid, qlen, abs_isize;
    uint32_t meta;        /* bits 0-3 ReadFlag after re-flagging, bit 4 reverse strand,
                             bits 8-15 library index, bits 16-23 bdqual */
    uint32_t record;      /* index of the record in the pushed stream */
    uint64_t qid;
} bdk_aread;

typedef struct bdk_ctx bdk_ctx;

/* Construct the per-GPU context (replaces the BreakDancer / ReadRegionData / BamSummary object
 * graph wired in BreakDancerMax.cpp:38-73). device = CUDA ordinal. */
int bdk_create(bdk_ctx** out, int device, const bdk_params* params);
void bdk_destroy(bdk_ctx* ctx);
/* Message of the last failure on ctx (or of the last failed bdk_create when ctx is NULL). */
const char* bdk_last_error(const bdk_ctx* ctx);
/* Run on a caller-provided CUDA stream (cudaStream_t); default is the context's own stream. */
int bdk_set_stream(bdk_ctx* ctx, void* cuda_stream);
/* Forget all pushed records and results; keeps the allocations (one context per job step). */
int bdk_reset(bdk_ctx* ctx);

/* Feed n position-sorted records from HOST memory (pinned memory makes the copies asynchronous
 * and overlapped with the classify kernel; pinned qlen / qid columns are not copied at all -- the
 * kernel reads the few values it needs, those of anomalous reads, in place over PCIe). Replaces the per-record loops
 * BamSummary::_analyze_bam (BamSummary.cpp:69-114) and BreakDancer::run/push_read up to the
 * point where a read is found anomalous (BreakDancer.cpp:139-207). May be called repeatedly. */
int bdk_push(bdk_ctx* ctx, const bdk_soa* host_cols, uint64_t n);
/* Same, columns already resident in this GPU's memory (16-byte aligned). */
int bdk_push_device(bdk_ctx* ctx, const bdk_soa* dev_cols, uint64_t n);

/* The same records in the decoder's WIRE FORMAT: 12 bytes per record cross the bus instead of 25. A run holds position-sorted
 * records of ONE reference sequence (tid; the mate is on the same sequence unless excepted):
 *   pos[i]   core.pos
 *   meta[i]  core.flag (bits 0-11) | bdqual (bits 12-19) | read-group id (bits 20-31; BDK_PACKED_EXCEPT: see below)
 *   rel[i]   core.isize (low 16 bits, signed) | core.mpos - core.pos (high 16 bits, signed)
 * A record that does not fit (mate on another sequence, |isize| or |mpos - pos| >= 32768, flag >= 4096, read group >= 4095)
 * carries BDK_PACKED_EXCEPT in the read-group field and its exact fields in the exception arrays (x_index ascending = index of
 * the record inside the run). The device expands a chunk back into the 25-byte columns (csrc/k1_classify.cuh: k1_expand_kernel)
 * and classifies it as usual, so results are identical to bdk_push of the unpacked records. qlen / qid as in bdk_soa.
 * bdk_pack builds a run from columns (host, multi-threaded; the BAM decoder writes the same layout); all arrays it returns are
 * pinned and owned by the bdk_packed_buf. Replaces, like bdk_push, the per-record hand-over of AlignmentSource::next
 * (src/lib/io/AlignmentSource.hpp:48-65). */
#define BDK_PACKED_EXCEPT 0xFFFu
typedef struct bdk_packed {
    const int32_t* pos;
    const uint32_t* meta;
    const uint32_t* rel;
    const int32_t* qlen;
    const uint64_t* qid;
    int32_t tid;
    uint32_t reserved;
    uint64_t nx;
    const uint32_t* x_index;
    const int32_t* x_mpos;
    const int32_t* x_mtid;
    const int32_t* x_isize;
    const uint16_t* x_flag;
    const uint16_t* x_rgid;
} bdk_packed;
typedef struct bdk_packed_buf bdk_packed_buf;
/* Pack n records (all of one tid, else BDK_ERR_ARG) given as columns; *view then describes the run. threads <= 0: all cores. */
int bdk_pack(const bdk_soa* host_cols, uint64_t n, int threads, bdk_packed_buf** out, bdk_packed* view);
void bdk_pack_free(bdk_packed_buf* buf);
/* Feed a packed run of n records from HOST memory (pinned: asynchronous copies overlapped with the kernels). */
int bdk_push_packed(bdk_ctx* ctx, const bdk_packed* run, uint64_t n);

/* Pass-1 statistics and the derived window (BamSummary + BreakDancerMax.cpp:83-116). */
int bdk_summary(bdk_ctx* ctx, bdk_summary_t* out);

/* Region building, link graph, connection walk, SV evaluation and scoring
 * (BreakDancer.cpp:209-512, ReadRegionData.cpp:70-227, SvBuilder.cpp, Graph.hpp). The result
 * arrays are owned by ctx and stay valid until the next bdk_reset/bdk_finish/bdk_destroy. */
int bdk_finish(bdk_ctx* ctx, bdk_result* out);

/* Inspection of intermediate stages (valid after bdk_finish). */
int bdk_get_regions(bdk_ctx* ctx, const bdk_region** regions, uint64_t* n);
int bdk_get_areads(bdk_ctx* ctx, const bdk_aread** reads, const int32_t** region_of_read, uint64_t* n);
/* For every anomalous read: order index of the SV call whose process_sv() consumed it as part
 * of a pair, or -1 (BreakDancer.cpp:367-368; feeds the -g BED / -d FASTQ writers). */
int bdk_get_support(bdk_ctx* ctx, const int32_t** sv_of_read, uint64_t* n);

/* Device time of the kernels of the last push/finish, milliseconds, measured with CUDA events
 * on the context's stream. names[i] are static strings. Returns the number of entries. */
int bdk_kernel_times(bdk_ctx* ctx, const char** names, float* ms, int* launches, int cap);

/* Number of kernels this context has launched since the last bdk_reset / bdk_create. */
uint64_t bdk_kernel_launches(bdk_ctx* ctx);
/* Sweeps over the connected components the last bdk_finish needed until the connection walk was stable
 * (csrc/bdk_logic.h, K4Static: 1 = no component depends on another one). */
uint32_t bdk_k4_sweeps(bdk_ctx* ctx);
/* Reads beyond the second one of a read-name key among the anomalous reads of the last bdk_finish (bams with overlapping read
 * names, a key collision). The reference keeps a list per name (ReadRegionData.cpp:99-114) and pairs whichever two a call meets
 * first; here every read of such a name stays unpaired (never consumed, its region is never cleared) and the job goes on. 0 on
 * well-formed input. */
uint32_t bdk_duplicate_names(bdk_ctx* ctx);
/* Bytes the last bdk_push copied host -> device with the copy engine. */
uint64_t bdk_h2d_bytes(bdk_ctx* ctx);
/* Bytes the last bdk_finish copied device -> host (ordered SV table + summary). */
uint64_t bdk_d2h_bytes(bdk_ctx* ctx);

/* Pinned host memory helpers for callers without their own CUDA binding. */
void* bdk_host_alloc(uint64_t bytes);
void bdk_host_free(void* p);

/* Multi-GPU, whole-genome semantics (one job over several GPUs; the reference's non-"-o" run, where regions on
 * different chromosomes are linked: BreakDancer.cpp:161,254-259, ReadRegionData.cpp:109-113). Per-chromosome
 * sharding ("-o chr" runs, README:31) needs no communicator: use one independent context per chromosome set.
 *
 * With a communicator attached, rank r pushes the r-th CONTIGUOUS SLICE of the globally (tid, pos)-sorted
 * record stream (rank order = stream order; a cut may fall anywhere). bdk_push / bdk_push_device stay local.
 * bdk_summary, bdk_finish and bdk_get_support become COLLECTIVE: every rank must call them, in the same order,
 * and every rank receives the same complete result (summary of the whole stream, the whole SV table, region and
 * read indices global). The exchanges run over NCCL on the context's stream: an all-gather of the compacted
 * anomalous reads (1-3 % of the records), a few-KB all-reduce of the pass-1 statistics, and an all-gather of
 * the SV rows each GPU produced for the connected components it walked (csrc/comm.cuh).
 * Attach before the first push of a job; a caller that saw a push fail on one rank must not enter the
 * collectives on the others.
 *
 * bdk_set_comm adopts the caller's ncclComm_t (not destroyed by bdk_destroy); nccl_comm = NULL detaches.
 * bdk_comm_unique_id + bdk_comm_init create one inside the library for callers without an NCCL binding
 * (rank 0 obtains the 128-byte id, hands it to the other ranks by any means, all call bdk_comm_init). */
int bdk_set_comm(bdk_ctx* ctx, void* nccl_comm, int rank, int nranks);
int bdk_comm_unique_id(void* out128, int cap);
int bdk_comm_init(bdk_ctx* ctx, const void* unique_id128, int rank, int nranks);
/* Bytes this rank received over NVLink in the exchanges of the last job. */
uint64_t bdk_comm_bytes(bdk_ctx* ctx);

/* Poisson tail used by ComputeProbScore (BreakDancer.cpp:62-68): log P[Pois(lambda) > k],
 * evaluated on the device (for known-answer tests). */
int bdk_poisson_logsf(bdk_ctx* ctx, const double* lambda, const int32_t* k, double* out, uint64_t n);

/* BGZF members inflated on the GPU (csrc/bgzf_inflate_warp.cuh, one warp per member, the decoder of bdk_push_bam; replaces the zlib calls under the reference's samtools reader,
 * `inflate_block` in bgzf.c of vendor/samtools-0.1.19.tar.gz, reached through `samread` in src/lib/io/BamReader.hpp:65,
 * for whole files at once). `file` is the host image of the BGZF file, `members`
 * its DEFLATE streams (offsets into `file` and into `out`; out_len from the member's ISIZE footer), `out` a host buffer of
 * out_bytes. status[i] = 0 if member i decoded to exactly out_len bytes whose CRC-32 equals the 4 bytes behind the stream (the
 * member's footer; the input must be readable 8 bytes beyond every stream, as in a BGZF file), else a positive error code: the
 * caller re-inflates whatever failed. Needs no context; returns 0 or BDK_ERR_*. kernel_ms (may be NULL) receives
 * the summed duration of the inflate kernel launches. */
typedef struct bdk_bgzf_member {
    uint64_t in_off;
    uint64_t out_off;
    uint32_t in_len;
    uint32_t out_len;
} bdk_bgzf_member;
int bdk_bgzf_inflate(int device, const uint8_t* file, uint64_t file_bytes, const bdk_bgzf_member* members, uint64_t n_members,
                     uint8_t* out, uint64_t out_bytes, int32_t* status, float* kernel_ms);

/* One BAM file decoded ON THE DEVICE and classified: the compressed BGZF members cross PCIe, the GPU inflates them (one warp per
 * member, CRC32 of every member checked against its footer), finds the record boundaries, applies the reader's filter, extracts
 * the columns of bdk_soa and runs the classify kernel on them, window of members by window, with the copy, the inflate and the
 * decode of consecutive windows overlapped (csrc/bdk_bam.inl, bgzf_inflate_warp.cuh, bam_decode.cuh). Replaces, for a run from
 * one file, samread + inflate_block of the vendored samtools reached through src/lib/io/BamReader.hpp:64-70 (or
 * RegionLimitedBamReader.hpp:40-66 with a region), the reader filter src/lib/io/BamIo.cpp:11-18, the Alignment constructor
 * src/lib/io/Alignment.cpp:12-64 and the read-group look-up src/lib/io/AlignmentSource.hpp:48-65 -- and then does what bdk_push does.
 * The records it pushes are the ones bdh_stream_open (host decoder) delivers for the same file, in the same order; read-group ids are
 * the caller's: rg_id[i] for a record whose RG:Z string hashes (bdk_hash_bytes) to rg_hash[i], rg_other for every other string or
 * no RG tag (the reference maps read groups its config does not know to the first bam's library, BamConfig.hpp:63-72).
 * The caller parses the BAM header (text, reference names) itself and passes the members from the one that holds the first record.
 * Errors: BDK_ERR_DATA for a member that does not inflate or fails its CRC, a truncated record, a record longer than 16 MiB; the
 * job then holds an unspecified prefix of the file (bdk_reset before doing anything else with the context). */
typedef struct bdk_bam_source {
    const uint8_t* file;              /* host image of the BGZF file (for instance an mmap) */
    uint64_t file_bytes;
    const bdk_bgzf_member* members;   /* DEFLATE streams in file order, each followed in the file by its CRC32 + ISIZE footer;
                                         out_off = running sum of out_len, 0 for members[0]; members without output may be left out */
    uint64_t n_members;
    uint64_t first_record;            /* offset, in the members' concatenated output, of the first record's block_size field;
                                         must lie inside members[0] (or at least inside the first window) */
    uint64_t end_offset;              /* where the records end (a record boundary), 0 = with the last member */
    int32_t n_ref;                    /* reference sequences in the BAM header */
    int32_t region_on, region_tid, region_beg, region_end;   /* -o: keep records overlapping [beg, end) of tid (0-based) */
    uint32_t n_rg;
    const uint64_t* rg_hash;          /* [n_rg] */
    const uint16_t* rg_id;            /* [n_rg] */
    uint16_t rg_other;
    uint16_t reserved;
    uint64_t window_bytes;            /* compressed bytes per pipeline step, 0 = default (32 MiB) */
} bdk_bam_source;
typedef struct bdk_bam_stats {
    uint64_t records;                 /* records in the stream */
    uint64_t kept;                    /* records that passed the reader's filter and were classified */
    uint64_t h2d_bytes;               /* bytes copied host -> device */
    uint64_t inflated_bytes;
    uint32_t windows;
    uint32_t guess_misses;            /* record-boundary guesses that were wrong or missing (resolved exactly, see bam_decode.cuh) */
    int32_t sorted;                   /* 1 iff the kept records are ordered by (reference sequence, position) */
    float inflate_ms, chain_ms, extract_ms;   /* device time: first inflate launch to the end of the last one (windows overlap);
                                                 the boundary search; the filter + extraction */
    float stage_ms;                   /* host time of the producer thread copying file bytes into its pinned staging buffers */
    float wall_ms;                    /* the whole call */
    uint32_t merge_parts;             /* bdk_push_bams with two bams (in stats[0]): independent parts the merge was cut into ... */
    uint32_t merge_longest_part;      /* ... and the records of the longest one (merged by one thread) */
} bdk_bam_stats;
int bdk_push_bam(bdk_ctx* ctx, const bdk_bam_source* src, bdk_bam_stats* stats);
/* The bams of one config, up to 16: each is decoded on the device as above, the record streams are merged in the order the
 * reference's BamMerger delivers them (src/lib/io/BamMerger.cpp:40-126: a priority queue over the streams' heads by
 * (tid, pos, strand)) and classified. TWO bams: merged on the device -- for two streams the queue's tie behaviour has a closed form,
 * csrc/bam_merge.cuh; both bams must be sorted by (reference sequence, position), else BDK_ERR_DATA. THREE OR MORE: the tie order
 * depends on the heap's history, so the order is computed by the queue itself on the host from the packed keys the device hands
 * back (csrc/host/nway_merge.hpp; 12 bytes per record over PCIe) and the columns are gathered through it on the device. srcs must
 * be in the config's bam order (BamMerger pushes the streams in that order), the read-group ids of the sources must not overlap,
 * stats (may be NULL) has n entries. bdk_decode_bams: the merged columns to the host instead (tests). */
int bdk_push_bams(bdk_ctx* ctx, const bdk_bam_source* srcs, int n, bdk_bam_stats* stats);
int bdk_decode_bams(bdk_ctx* ctx, const bdk_bam_source* srcs, int n, const bdk_soa* host_out, uint64_t cap, bdk_bam_stats* stats);
/* The same decode, but the columns come back to the HOST (arrays of `cap` records the caller owns, written through the const
 * pointers of *host_out) and nothing is classified: the device decoder as a drop-in for bdh_stream_open on one file, and what the
 * parity tests compare with the host decoder column by column. stats->kept = records written. */
int bdk_decode_bam(bdk_ctx* ctx, const bdk_bam_source* src, const bdk_soa* host_out, uint64_t cap, bdk_bam_stats* stats);
/* The 64-bit key of a byte string the decoders use for read names (bdk_soa.qid) and read-group strings. */
uint64_t bdk_hash_bytes(const void* bytes, uint64_t n);

const char* bdk_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BDK_H */
