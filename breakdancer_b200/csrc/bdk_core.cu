// bdk_core.cu -- the per-GPU context behind the bdk C ABI (include/bdk.h): device memory,
// streams, and the launch sequence  K1 classify + ordered compaction (+ span search) -> finalize
// -> K2 regions -> K3 mate join / link sort / run-length / components -> K4 connection walk + score.
// There is no CPU implementation of any stage in this library: every entry point that computes
// needs a CUDA device and fails with BDK_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bdk.h"
#include "bdk_finalize.h"
#include "bgzf_inflate.cuh"
#include "bgzf_inflate_warp.cuh"
#include "bam_decode.cuh"
#include "bam_merge.cuh"
#include "comm.cuh"
#include "k1_classify.cuh"
#include "k234_regions_links_sv.cuh"
#include "scan_sort.cuh"

using namespace bdk;

namespace {

std::string g_create_error;
std::mutex g_err_mu;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct StageTimer {
    const char* name;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms = 0.0f;
    int launches = 0;
    bool pending = false;
};

enum { T_H2D = 0, T_K1, T_SPAN, T_FINALIZE, T_K2, T_K3, T_K4, T_D2H, T_HOST, T_COMM1, T_COMM2, T_INFLATE, T_CHAIN, T_EXTRACT, T_N };
const char* kTimerNames[T_N] = {"h2d_copy", "k1_classify", "k1_span", "finalize_summary", "k2_regions", "k3_links_graph", "k4_sv_score", "d2h_results",
                                "host_order_rows",    // host wall time (output ordering), not a device timer
                                "comm_gather_reads", "comm_gather_rows",    // multi-GPU exchanges (NCCL over NVLink)
                                "bam_inflate", "bam_record_chain", "bam_extract"};   // bdk_push_bam: device-resident BAM decode (bdk_bam.inl)

}  // namespace

struct bdk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr, copy_stream = nullptr;
    bdk_params P;
    std::vector<bdk_lib> libs;
    std::vector<int32_t> rg_lib, rg_bam;
    int nkey = 1, period = 1;
    std::string err;

    // constants on the device
    DevBuf d_rgtab, d_cnt_rg, d_lib_mean, d_blibs, d_rg_lib, d_rg_bam;
    int ncnt = 0;                 // private pass-1 counter columns (0: warp-vote fallback)
    int pad_rg = 0;               // a read group that has a library
    // pass-1 accumulators: one block so it can be snapshotted before a push
    DevBuf d_acc, d_acc_bak;
    size_t acc_bytes = 0, off_first = 0, off_last = 0, off_hist = 0, off_err = 0, off_cursor = 0;
    // per-job state
    uint64_t n_records = 0;       // records pushed so far
    uint32_t A = 0;               // anomalous reads compacted so far
    uint32_t out_cap = 0;         // capacity of d_ar / d_P in reads
    // K1 per-CTA output segments (compacted into d_ar / d_P by k1_compact_kernel), reused by every launch
    DevBuf d_seg_ar, d_seg_P, d_seg_cnt, d_carry_out, d_tile_bams, d_stash;
    uint64_t tile_cap = 0;
    uint32_t seg_cap = 0, seg_cap_min = 8192;
    // chunk buffers for host pushes
    DevBuf d_chunk[2][10];
    DevBuf d_pk[2][3], d_px[2][6];      // packed pushes: staging of the wire-format arrays and of a chunk's exceptions
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    // finish() work space
    DevBuf d_cnt, d_ar, d_P, d_summary, d_density, d_scan_sums, d_read_cand, d_read_region,
        d_mate, d_sv_of_read, d_cand_first, d_cand_maxlen, d_cand_info, d_cand_regs, d_reg, d_table, d_ekeys, d_ecnt,
        d_se, d_se2, d_wstart, d_wdir, d_wund, d_wfill, d_slot_base, d_sefl, d_queue, d_biglist, d_rowpack, d_outpack, d_slot_order, d_pois_l, d_pois_k, d_pois_o;
    uint64_t d2h_bytes = 0;
    uint32_t rows_guess = 0;      // rows the first result copy brings back (a second copy follows only when there are more)
    uint32_t n_slots = 0;
    void* h_pack = nullptr;       // pinned host block the row outputs + summary are copied into
    size_t h_pack_cap = 0;
    bool finished = false, summary_ready = false;
    int k1_blocks_per_sm = 0;
    bool k1_force_general = false;    // BDK_K1_GENERAL (tests): the general multi-key variant also where the four-key one applies
    size_t k1_smem = 0;
    uint64_t launches = 0;       // kernels launched since the last bdk_reset
    uint64_t h2d_bytes = 0;      // bytes the last bdk_push copied host -> device
    // host results
    bdk_summary_t h_summary;
    std::vector<bdk_sv> h_sv;
    std::vector<int32_t> h_lib_count;
    std::vector<uint32_t> h_cn_count;
    std::vector<float> h_copy_number;
    std::vector<bdk_region> h_regions;
    std::vector<bdk_aread> h_areads;
    std::vector<int32_t> h_read_region, h_sv_of_read;
    uint32_t h_cnt[CNT_N] = {0};
    StageTimer timers[T_N];
    // multi-GPU whole-genome mode (comm.cuh)
    ncclComm_t comm = nullptr;
    bool comm_owned = false, exchanged = false;
    int rank = 0, nranks = 1;
    uint32_t A_local = 0;             // anomalous reads of this rank's slice (c->A becomes the global count)
    uint64_t comm_bytes = 0;          // bytes this rank received in the exchanges of the last job
    DevBuf d_hdr, d_hdr_all, d_ar_g, d_P_g, d_koff;
    DevBuf d_del, d_stamp, d_k4sync, d_c1, d_ri;   // K4: table of deletion windows, sweep stamps, barrier / counters
    uint32_t k4_sweeps = 0;                   // sweeps of the last bdk_finish
    uint32_t dup_names = 0;                   // reads beyond the second of a read name among the anomalous reads of the last bdk_finish
    int k4_grid_max = 0;                      // co-resident CTAs of the persistent sweep kernel
    bool k4_host_loop = false;                // BDK_K4_HOST_LOOP (tests): one launch per sweep instead of the persistent kernel
    uint32_t k4w_cap = K4W_CAP;               // BDK_K4W_CAP (tests): directed edges up to which a window is handled by one warp
    uint32_t rows_guess_min = 1024;           // BDK_ROWS_GUESS (tests): floor of the first result copy
    void* bamdev = nullptr;                   // buffers and streams of bdk_push_bam (bdk_bam.inl), created at its first call
};

namespace {

void bamdev_release(void* p);

int fail(bdk_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf;
    else { std::lock_guard<std::mutex> g(g_err_mu); g_create_error = buf; }
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(c, BDK_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

int ensure(bdk_ctx* c, DevBuf& b, size_t bytes, bool preserve = false) {
    if (bytes <= b.cap && b.p) return 0;
    size_t ncap = std::max<size_t>(bytes + bytes / 4, 256);
    void* np = nullptr;
    CU(cudaMalloc(&np, ncap));
    if (preserve && b.p && b.cap) CU(cudaMemcpyAsync(np, b.p, b.cap, cudaMemcpyDeviceToDevice, c->stream));
    if (b.p) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(b.p)); }
    b.p = np; b.cap = ncap;
    return 0;
}
#define ENS(buf, bytes) do { int rc_ = ensure(c, buf, (bytes)); if (rc_) return rc_; } while (0)
#define ENSP(buf, bytes) do { int rc_ = ensure(c, buf, (bytes), true); if (rc_) return rc_; } while (0)

void tstart(bdk_ctx* c, int t) { cudaEventRecord(c->timers[t].e0, c->stream); }
void tstop(bdk_ctx* c, int t) { cudaEventRecord(c->timers[t].e1, c->stream); c->timers[t].pending = true; c->timers[t].launches++; }
void tcollect(bdk_ctx* c) {
    for (int t = 0; t < T_N; ++t)
        if (c->timers[t].pending) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, c->timers[t].e0, c->timers[t].e1) == cudaSuccess) c->timers[t].ms += ms;
            c->timers[t].pending = false;
        }
}

int reset_job(bdk_ctx* c) {
    c->n_records = 0; c->A = 0; c->A_local = 0; c->exchanged = false; c->comm_bytes = 0;
    c->finished = false; c->summary_ready = false; c->launches = 0;
    // accumulators: counts 0, first = ~0, last = 0
    CU(cudaMemsetAsync(c->d_acc.p, 0, c->acc_bytes, c->stream));
    size_t nbt = (size_t)c->P.nbam * c->P.ntid;
    if (nbt) CU(cudaMemsetAsync((char*)c->d_acc.p + c->off_first, 0xff, nbt * 8, c->stream));
    for (int t = 0; t < T_N; ++t) { c->timers[t].ms = 0; c->timers[t].launches = 0; c->timers[t].pending = false; }
    return 0;
}

typedef void (*K1Fn)(const K1Args);
template <int MODE> K1Fn k1_pick(bool smem, bool plain) {
    if (smem) return plain ? k1_classify_kernel<MODE, true, true> : k1_classify_kernel<MODE, true, false>;
    return plain ? k1_classify_kernel<MODE, false, true> : k1_classify_kernel<MODE, false, false>;
}
K1Fn k1_fn(const bdk_ctx* c) {
    const bool single = c->nkey == 1 && c->P.nbam == 1 && c->ncnt == 1, smem = c->P.nrg <= K1_RG_SMEM;
    const bool keys4 = !single && c->nkey <= 4 && c->P.nbam <= 4 && c->ncnt >= 1 && c->ncnt <= 4 && !c->k1_force_general;
    const bool plain = !c->P.transchr_rearrange && !c->P.illumina_long_insert;
    if (single) return k1_pick<K1_FAST>(smem, plain);
    if (keys4) return k1_pick<K1_KEYS4>(smem, plain);
    return k1_pick<K1_GENERAL>(smem, plain);
}

int grow_segments(bdk_ctx* c, uint64_t want);

// one K1 launch (+ segment compaction + the span search) over n records whose columns are on the device;
// qlen / qid may be mapped host memory (only anomalous records touch them)
int launch_k1(bdk_ctx* c, const bdk_soa& cols, uint64_t n, uint32_t base_index, bool timed) {
    const uint64_t ntiles = div_up<uint64_t>(n, K1_TILE);
    if (!ntiles) return 0;
    const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)kNumSMs * c->k1_blocks_per_sm);
    K1Args a;
    a.c = cols; a.n = n; a.base_index = base_index;
    a.tiles_per_cta = (uint32_t)div_up<uint64_t>(ntiles, grid);
    a.rgtab = c->d_rgtab.as<RgDev>();
    a.nrg = c->P.nrg; a.nlib = c->P.nlib; a.nbam = c->P.nbam; a.nkey = c->nkey;
    a.pad_rg = c->pad_rg;
    a.ncnt = c->ncnt; a.cnt_rg = c->d_cnt_rg.as<int32_t>(); a.cn_lib = c->P.cn_lib;
    a.co.max_sd = c->P.max_sd; a.co.transchr = c->P.transchr_rearrange; a.co.long_insert = c->P.illumina_long_insert;
    a.seg_ar = c->d_seg_ar.as<bdk_aread>(); a.seg_P = c->d_seg_P.as<uint32_t>(); a.seg_cap = c->seg_cap;
    a.seg_cnt = c->d_seg_cnt.as<uint32_t>();
    a.stash = c->d_stash.as<K1Stash>();
    a.tile_bams = c->d_tile_bams.as<unsigned long long>();
    char* acc = (char*)c->d_acc.p;
    a.rg_sproper = (unsigned long long*)acc;
    a.hist = (uint32_t*)(acc + c->off_hist); a.err = (uint32_t*)(acc + c->off_err);
    uint32_t* carry = (uint32_t*)(acc + c->off_cursor);
    if (timed) tstart(c, T_K1);
    k1_fn(c)<<<grid, K1_THREADS, c->k1_smem, c->stream>>>(a);
    k1_compact_kernel<<<grid, 256, 0, c->stream>>>(a.seg_ar, a.seg_P, a.seg_cap, a.seg_cnt, c->nkey, carry, c->d_carry_out.as<uint32_t>(),
                                                   c->d_ar.as<bdk_aread>(), c->d_P.as<uint32_t>(), c->out_cap, a.err);
    CU(cudaMemcpyAsync(carry, c->d_carry_out.p, 4 * (1 + (size_t)c->nkey), cudaMemcpyDeviceToDevice, c->stream));
    if (timed) tstop(c, T_K1); else c->timers[T_K1].launches++;     // host pushes: the kernel's time is inside the h2d_copy span
    CU(cudaGetLastError());
    if (timed) tstart(c, T_SPAN);
    const int64_t items = (int64_t)c->P.ntid * c->P.nbam;
    const unsigned sgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(div_up<int64_t>(items, 4), (int64_t)kNumSMs * 8));
    k1_span_kernel<<<sgrid, 128, 0, c->stream>>>(cols.tid, cols.pos, cols.rgid, n, base_index, c->d_rgtab.as<RgDev>(), c->P.nrg, c->P.nbam, c->P.ntid,
                                                 c->d_tile_bams.as<unsigned long long>(), (unsigned long long*)(acc + c->off_first),
                                                 (unsigned long long*)(acc + c->off_last));
    if (timed) tstop(c, T_SPAN); else c->timers[T_SPAN].launches++;
    c->launches += 3;
    CU(cudaGetLastError());
    return 0;
}

// per-tile tables and per-CTA output segments for launches of up to `n` records
int grow_tiles(bdk_ctx* c, uint64_t n) {
    const uint64_t tiles = div_up<uint64_t>(n, K1_TILE);
    if (tiles > c->tile_cap) { ENS(c->d_tile_bams, tiles * 8); c->tile_cap = tiles; }
    const uint64_t grid = (uint64_t)kNumSMs * c->k1_blocks_per_sm;
    const uint64_t want = std::max<uint64_t>(c->seg_cap_min, n / (8 * grid));      // 12.5 % anomalous, evenly spread
    if (want > c->seg_cap) return grow_segments(c, want);
    return 0;
}

int grow_out(bdk_ctx* c, uint64_t want) {
    if (want <= c->out_cap) return 0;
    if (want > 0xfffffff0ull) return fail(c, BDK_ERR_NOMEM, "more than 2^32 anomalous reads in one context");
    ENSP(c->d_ar, want * sizeof(bdk_aread));
    ENSP(c->d_P, want * 4 * c->nkey);
    c->out_cap = (uint32_t)want;
    return 0;
}

// push of one batch whose columns are already device pointers; handles staging overflow by retry
int grow_segments(bdk_ctx* c, uint64_t want) {
    const uint64_t grid = (uint64_t)kNumSMs * c->k1_blocks_per_sm;
    if (want > 0xfffffff0ull) return fail(c, BDK_ERR_NOMEM, "segment too large");
    ENS(c->d_seg_ar, grid * want * sizeof(bdk_aread));
    ENS(c->d_seg_P, grid * want * 4 * c->nkey);
    c->seg_cap = (uint32_t)want;
    return 0;
}

template <class RunFn>
int push_common(bdk_ctx* c, uint64_t n, uint64_t max_launch, RunFn run) {
    if (c->finished) return fail(c, BDK_ERR_STATE, "bdk_push after bdk_finish (call bdk_reset first)");
    if (c->exchanged) return fail(c, BDK_ERR_STATE, "bdk_push after the collective bdk_summary / bdk_finish of a multi-GPU job (call bdk_reset first)");
    if (n == 0) return 0;
    if (c->n_records + n > 0xffffffffull) return fail(c, BDK_ERR_ARG, "more than 2^32 records in one context");
    int rc = grow_tiles(c, std::min(n, max_launch));
    if (rc) return rc;
    rc = grow_out(c, std::max<uint64_t>((uint64_t)c->A + std::max<uint64_t>(n / 16, 1 << 16), c->out_cap));
    if (rc) return rc;
    for (int attempt = 0; attempt < 3; ++attempt) {
        CU(cudaMemcpyAsync(c->d_acc_bak.p, c->d_acc.p, c->acc_bytes, cudaMemcpyDeviceToDevice, c->stream));
        rc = run();
        if (rc) return rc;
        uint32_t tail[3];   // err, largest per-CTA segment, anomalous reads so far
        CU(cudaMemcpyAsync(tail, (char*)c->d_acc.p + c->off_err, 12, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        tcollect(c);
        if (tail[0] & K1_ERR_RG)
            return fail(c, BDK_ERR_DATA, "library index out of range (a record's read group has no library)");
        if (tail[0] & K1_ERR_OVERFLOW) {   // a segment or the output is too small: restore the accumulators, grow, run again
            CU(cudaMemcpyAsync(c->d_acc.p, c->d_acc_bak.p, c->acc_bytes, cudaMemcpyDeviceToDevice, c->stream));
            if (tail[1] > c->seg_cap) { rc = grow_segments(c, (uint64_t)tail[1] + tail[1] / 4 + 64); if (rc) return rc; }
            rc = grow_out(c, (uint64_t)tail[2] + (tail[2] - c->A) / 8 + 1024);
            if (rc) return rc;
            continue;
        }
        c->A = tail[2];
        c->summary_ready = false;
        c->n_records += n;
        return 0;
    }
    return fail(c, BDK_ERR_NOMEM, "output overflow persisted");
}

#define NC(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) return fail(c, BDK_ERR_NCCL, "%s failed: %s", #call, nc->GetErrorString(r_)); \
    } while (0)

// in-place all-gather of per-rank ranges of one array: rank p owns elements [cut[p], cut[p + 1]) of `elt` bytes each. Every rank
// sends its range straight to every peer (point-to-point over NVSwitch; the caller groups the calls).
int gather_ranges(bdk_ctx* c, NcclApi* nc, void* base, size_t elt, const uint64_t* cut) {
    const uint64_t mine = cut[c->rank + 1] - cut[c->rank];
    for (int p = 0; p < c->nranks; ++p) {
        if (p == c->rank) continue;
        const uint64_t cnt = cut[p + 1] - cut[p];
        if (mine) NC(nc->Send((char*)base + cut[c->rank] * elt, mine * elt, ncclUint8, p, c->comm, c->stream));
        if (cnt) { NC(nc->Recv((char*)base + cut[p] * elt, cnt * elt, ncclUint8, p, c->comm, c->stream)); c->comm_bytes += cnt * elt; }
    }
    return 0;
}

// Exchange 1 (comm.cuh): totals, pass-1 accumulators, and the anomalous-read streams of all ranks. Collective.
int comm_exchange(bdk_ctx* c) {
    if (!c->comm || c->exchanged) return 0;
    NcclApi* nc = nccl_api();
    if (!nc) return fail(c, BDK_ERR_NCCL, "%s", nccl_api_error());
    const int nkey = c->nkey, ncomp = 1 + nkey, W = ncomp + 2, N = c->nranks;
    char* acc = (char*)c->d_acc.p;
    tstart(c, T_COMM1);
    // totals of every rank: anomalous reads, kept proper pairs per key, records
    ENS(c->d_hdr, (size_t)W * 4); ENS(c->d_hdr_all, (size_t)N * W * 4); ENS(c->d_koff, (size_t)nkey * 4);
    std::vector<uint32_t> hdr(W, 0), all((size_t)N * W, 0);
    hdr[0] = c->A; hdr[ncomp] = (uint32_t)c->n_records; hdr[ncomp + 1] = (uint32_t)(c->n_records >> 32);
    CU(cudaMemcpyAsync(c->d_hdr.p, hdr.data(), (size_t)W * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_hdr.as<uint32_t>() + 1, acc + c->off_cursor + 4, (size_t)nkey * 4, cudaMemcpyDeviceToDevice, c->stream));
    NC(nc->AllGather(c->d_hdr.p, c->d_hdr_all.p, W, ncclUint32, c->comm, c->stream));
    CU(cudaMemcpyAsync(all.data(), c->d_hdr_all.p, (size_t)N * W * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    std::vector<uint64_t> a_cut(N + 1, 0);
    std::vector<uint32_t> koff(nkey, 0);
    uint64_t rec_off = 0, rec_tot = 0;
    for (int p = 0; p < N; ++p) {
        const uint32_t* h = all.data() + (size_t)p * W;
        a_cut[p + 1] = a_cut[p] + h[0];
        const uint64_t nrec = (uint64_t)h[ncomp] | ((uint64_t)h[ncomp + 1] << 32);
        if (p < c->rank) { rec_off += nrec; for (int k = 0; k < nkey; ++k) koff[k] += h[1 + k]; }
        rec_tot += nrec;
    }
    if (rec_tot > 0xffffffffull) return fail(c, BDK_ERR_ARG, "more than 2^32 records in one multi-GPU job");
    const uint64_t A_tot = a_cut[N];
    if (A_tot > 0xfffffff0ull) return fail(c, BDK_ERR_NOMEM, "more than 2^32 anomalous reads in one multi-GPU job");
    // pass-1 accumulators: global record indices in the first / last keys, then sum / min / max over the ranks
    const uint32_t nbt = (uint32_t)((size_t)c->P.nbam * c->P.ntid);
    unsigned long long* first = (unsigned long long*)(acc + c->off_first);
    unsigned long long* last = (unsigned long long*)(acc + c->off_last);
    comm_rebase_span_kernel<<<std::max(1u, std::min(div_up(nbt, 256u), 1184u)), 256, 0, c->stream>>>(first, last, nbt, (uint32_t)rec_off);
    NC(nc->GroupStart());
    NC(nc->AllReduce(acc, acc, (size_t)c->P.nrg, ncclUint64, ncclSum, c->comm, c->stream));
    NC(nc->AllReduce(first, first, nbt, ncclUint64, ncclMin, c->comm, c->stream));
    NC(nc->AllReduce(last, last, nbt, ncclUint64, ncclMax, c->comm, c->stream));
    NC(nc->AllReduce(acc + c->off_hist, acc + c->off_hist, (size_t)c->P.nlib * BDK_NUM_FLAGS, ncclUint32, ncclSum, c->comm, c->stream));
    NC(nc->GroupEnd());
    // the global anomalous-read stream: every rank writes its rebased slice into place, the slices are all-gathered
    ENS(c->d_ar_g, (A_tot + 2) * sizeof(bdk_aread)); ENS(c->d_P_g, (A_tot + 2) * 4 * nkey);
    CU(cudaMemcpyAsync(c->d_koff.p, koff.data(), (size_t)nkey * 4, cudaMemcpyHostToDevice, c->stream));
    if (c->A)
        comm_rebase_stream_kernel<<<GS_GRID, GS_THREADS, 0, c->stream>>>(c->d_ar.as<bdk_aread>(), c->d_P.as<uint32_t>(), c->A, nkey, (uint32_t)rec_off,
            c->d_koff.as<uint32_t>(), c->d_ar_g.as<bdk_aread>() + a_cut[c->rank], c->d_P_g.as<uint32_t>() + a_cut[c->rank] * nkey);
    CU(cudaGetLastError());
    NC(nc->GroupStart());
    int rc = gather_ranges(c, nc, c->d_ar_g.p, sizeof(bdk_aread), a_cut.data());
    if (!rc) rc = gather_ranges(c, nc, c->d_P_g.p, (size_t)4 * nkey, a_cut.data());
    NC(nc->GroupEnd());
    if (rc) return rc;
    tstop(c, T_COMM1);
    c->launches += 2;
    CU(cudaStreamSynchronize(c->stream));   // koff / hdr are host stack data; also surfaces asynchronous NCCL errors here
    tcollect(c);
    std::swap(c->d_ar, c->d_ar_g); std::swap(c->d_P, c->d_P_g);
    c->out_cap = (uint32_t)std::min<uint64_t>({c->d_ar.cap / sizeof(bdk_aread), c->d_P.cap / ((size_t)4 * nkey), 0xfffffff0ull});
    c->A_local = c->A; c->A = (uint32_t)A_tot;
    c->n_records = rec_tot;
    c->exchanged = true;
    c->summary_ready = false;
    return 0;
}

}  // namespace

extern "C" {

const char* bdk_version(void) { return "breakdancer-b200 0.1 (sm_100a)"; }

const char* bdk_last_error(const bdk_ctx* c) {
    if (c) return c->err.c_str();
    std::lock_guard<std::mutex> g(g_err_mu);
    return g_create_error.c_str();
}

void* bdk_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void bdk_host_free(void* p) { if (p) cudaFreeHost(p); }

void bdk_destroy(bdk_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && c->comm_owned) { if (NcclApi* nc = nccl_api()) nc->CommDestroy(c->comm); }
    c->comm = nullptr;
    DevBuf* all[] = {&c->d_hdr, &c->d_hdr_all, &c->d_ar_g, &c->d_P_g, &c->d_koff, &c->d_del, &c->d_stamp, &c->d_k4sync, &c->d_c1, &c->d_ri,
        &c->d_rgtab, &c->d_cnt_rg, &c->d_lib_mean, &c->d_blibs, &c->d_rg_lib, &c->d_rg_bam, &c->d_acc, &c->d_acc_bak, &c->d_seg_ar,
        &c->d_seg_P, &c->d_seg_cnt, &c->d_carry_out, &c->d_tile_bams, &c->d_stash, &c->d_cnt, &c->d_ar, &c->d_P, &c->d_summary,
        &c->d_density, &c->d_scan_sums, &c->d_read_cand, &c->d_read_region, &c->d_mate, &c->d_sv_of_read,
        &c->d_cand_first, &c->d_cand_maxlen, &c->d_cand_info, &c->d_cand_regs, &c->d_reg, &c->d_table, &c->d_ekeys, &c->d_ecnt,
        &c->d_se, &c->d_se2, &c->d_wstart, &c->d_wdir, &c->d_wund, &c->d_wfill, &c->d_slot_base, &c->d_sefl, &c->d_queue, &c->d_biglist, &c->d_rowpack, &c->d_outpack,
        &c->d_slot_order, &c->d_pois_l, &c->d_pois_k, &c->d_pois_o};
    for (DevBuf* b : all) if (b->p) cudaFree(b->p);
    for (int i = 0; i < 2; ++i) {
        for (int k = 0; k < 10; ++k) if (c->d_chunk[i][k].p) cudaFree(c->d_chunk[i][k].p);
        for (int k = 0; k < 3; ++k) if (c->d_pk[i][k].p) cudaFree(c->d_pk[i][k].p);
        for (int k = 0; k < 6; ++k) if (c->d_px[i][k].p) cudaFree(c->d_px[i][k].p);
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
    }
    for (int t = 0; t < T_N; ++t) { if (c->timers[t].e0) cudaEventDestroy(c->timers[t].e0); if (c->timers[t].e1) cudaEventDestroy(c->timers[t].e1); }
    if (c->h_pack) cudaFreeHost(c->h_pack);
    bamdev_release(c->bamdev);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

int bdk_create(bdk_ctx** out, int device, const bdk_params* p) {
    bdk_ctx* c = nullptr;
    if (!out || !p) return fail(c, BDK_ERR_ARG, "null argument");
    *out = nullptr;
    if (p->nlib < 1 || p->nlib > BDK_MAX_LIBS) return fail(c, BDK_ERR_ARG, "number of libraries must be in [1, %d]", BDK_MAX_LIBS);
    if (p->nbam < 1 || p->nbam > BDK_MAX_BAMS) return fail(c, BDK_ERR_ARG, "number of bams must be in [1, %d]", BDK_MAX_BAMS);
    if (p->nrg < 1 || p->nrg > 65536) return fail(c, BDK_ERR_ARG, "number of read groups must be in [1, 65536]");
    if (p->ntid < 1) return fail(c, BDK_ERR_ARG, "ntid must be >= 1");
    if (p->min_read_pair < 1) return fail(c, BDK_ERR_ARG, "-r (min_read_pair) must be >= 1");
    if (p->cn_lib && p->nlib > K1_MAXK) return fail(c, BDK_ERR_ARG, "-a supports at most %d libraries", K1_MAXK);
    if (!p->libs || !p->rg_lib || !p->rg_bam) return fail(c, BDK_ERR_ARG, "null table in bdk_params");
    for (int i = 0; i < p->nlib; ++i)
        if (p->libs[i].bam_index < 0 || p->libs[i].bam_index >= p->nbam) return fail(c, BDK_ERR_ARG, "library %d: bam_index out of range", i);
    for (int i = 0; i < p->nrg; ++i) {
        if (p->rg_lib[i] >= p->nlib) return fail(c, BDK_ERR_ARG, "rg_lib[%d] out of range", i);
        if (p->rg_bam[i] < 0 || p->rg_bam[i] >= p->nbam) return fail(c, BDK_ERR_ARG, "rg_bam[%d] out of range", i);
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(c, BDK_ERR_CUDA, "no CUDA device available (%s); the bdk hot path has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(c, BDK_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    c = new bdk_ctx;
    c->device = device;
    c->P = *p;
    c->libs.assign(p->libs, p->libs + p->nlib);
    c->rg_lib.assign(p->rg_lib, p->rg_lib + p->nrg);
    c->rg_bam.assign(p->rg_bam, p->rg_bam + p->nrg);
    c->P.libs = c->libs.data(); c->P.rg_lib = c->rg_lib.data(); c->P.rg_bam = c->rg_bam.data();
    c->nkey = nkey_of(c->P); c->period = period_of(c->P);
#define CUC(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            int rc_ = fail(nullptr, BDK_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));  \
            bdk_destroy(c);                                                                         \
            return rc_;                                                                             \
        }                                                                                           \
    } while (0)
    CUC(cudaSetDevice(device));
    {   // highest priority: beside the inflate kernels of bdk_push_bam (lowest) these launches get the warp slots that come free
        int prio_lo = 0, prio_hi = 0;
        CUC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        if (getenv("BDK_BAMDEV_NOPRIO")) prio_hi = 0;
        CUC(cudaStreamCreateWithPriority(&c->own_stream, cudaStreamNonBlocking, prio_hi));
    }
    CUC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (int i = 0; i < 2; ++i) { CUC(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming)); CUC(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming)); }
    for (int t = 0; t < T_N; ++t) { c->timers[t].name = kTimerNames[t]; CUC(cudaEventCreate(&c->timers[t].e0)); CUC(cudaEventCreate(&c->timers[t].e1)); }
    // constant tables: per-read-group constants with the ids a record maps to
    std::vector<LibDev> ld = make_libdev(c->P);
    std::vector<float> lm = make_lib_mean(c->P);
    std::vector<RgDev> rgt(p->nrg + 1);
    std::vector<int32_t> cnt_rg;                       // counter column -> representative read group
    {
        std::vector<std::pair<int, int>> pairs;        // distinct (library, bam) pairs, in order of appearance
        for (int i = 0; i < p->nrg; ++i) {
            if (p->rg_lib[i] < 0) continue;
            std::pair<int, int> pr(p->rg_lib[i], p->rg_bam[i]);
            if (std::find(pairs.begin(), pairs.end(), pr) == pairs.end()) { pairs.push_back(pr); cnt_rg.push_back(i); }
        }
        c->ncnt = (int)pairs.size() <= K1_PRIV_CNT ? std::max<int>(1, (int)pairs.size()) : 0;
        if (!c->ncnt) cnt_rg.clear();
        if (cnt_rg.empty()) cnt_rg.push_back(0);
        for (int i = p->nrg - 1; i >= 0; --i) if (p->rg_lib[i] >= 0) c->pad_rg = i;
        for (int i = 0; i <= p->nrg; ++i) {
            RgDev& r = rgt[i];
            if (i == p->nrg || p->rg_lib[i] < 0) { r.upper = 0; r.lower = 0; r.min_mapq = 0x7fffffff; r.info = RGI_INVALID; continue; }
            const int lib = p->rg_lib[i], bam = p->rg_bam[i];
            const int col = c->ncnt ? (int)(std::find(pairs.begin(), pairs.end(), std::make_pair(lib, bam)) - pairs.begin()) : (int)RGI_CNT_NONE;
            r.upper = ld[lib].upper; r.lower = ld[lib].lower; r.min_mapq = ld[lib].min_mapq;
            r.info = (uint32_t)lib | ((uint32_t)ld[lib].key << RGI_KEY_SHIFT) | ((uint32_t)bam << RGI_BAM_SHIFT) | ((uint32_t)col << RGI_CNT_SHIFT);
        }
    }
    auto up = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e1 = cudaMalloc(&b.p, std::max<size_t>(bytes, 16));
        if (e1 != cudaSuccess) return e1;
        b.cap = bytes;
        return cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice);
    };
    CUC(up(c->d_rgtab, rgt.data(), rgt.size() * sizeof(RgDev)));
    CUC(up(c->d_cnt_rg, cnt_rg.data(), cnt_rg.size() * 4));
    CUC(up(c->d_lib_mean, lm.data(), lm.size() * 4));
    CUC(up(c->d_blibs, c->libs.data(), c->libs.size() * sizeof(bdk_lib)));
    CUC(up(c->d_rg_lib, c->rg_lib.data(), c->rg_lib.size() * 4));
    CUC(up(c->d_rg_bam, c->rg_bam.data(), c->rg_bam.size() * 4));
    // accumulator block
    size_t nbt = (size_t)p->nbam * p->ntid;
    c->off_first = (size_t)p->nrg * 8;
    c->off_last = c->off_first + nbt * 8;
    c->off_hist = c->off_last + nbt * 8;
    c->off_err = c->off_hist + (size_t)p->nlib * BDK_NUM_FLAGS * 4;
    c->off_err = (c->off_err + 7) & ~size_t(7);
    c->off_cursor = c->off_err + 8;                       // err, max segment; then the carry: anomalous reads, kept proper pairs per key
    c->acc_bytes = c->off_cursor + 4 * (1 + (size_t)c->nkey);
    CUC(cudaMalloc(&c->d_acc.p, c->acc_bytes)); c->d_acc.cap = c->acc_bytes;
    CUC(cudaMalloc(&c->d_acc_bak.p, c->acc_bytes)); c->d_acc_bak.cap = c->acc_bytes;
    CUC(cudaMalloc(&c->d_cnt.p, CNT_N * 4)); c->d_cnt.cap = CNT_N * 4;
    CUC(cudaMalloc(&c->d_summary.p, sizeof(bdk_summary_t))); c->d_summary.cap = sizeof(bdk_summary_t);
    CUC(cudaMalloc(&c->d_density.p, (size_t)std::max(1, c->nkey) * 4)); c->d_density.cap = (size_t)std::max(1, c->nkey) * 4;
    CUC(cudaMalloc(&c->d_scan_sums.p, SS_GRID * 4)); c->d_scan_sums.cap = SS_GRID * 4;
    if (const char* e = getenv("BDK_K1_GENERAL")) c->k1_force_general = atoi(e) != 0;
    {   // K1 launch shape: dynamic shared memory and resident CTAs per SM
        c->k1_smem = k1_smem_bytes(p->nrg, p->nlib, c->ncnt, c->nkey, c->nkey == 1, p->nrg <= K1_RG_SMEM);
        int bps = 0;
        CUC(cudaFuncSetAttribute(k1_fn(c), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->k1_smem));
        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k1_fn(c), K1_THREADS, c->k1_smem));
        c->k1_blocks_per_sm = std::max(1, std::min(bps, 2));
        // thread-private stash of the anomalous reads of a tile, two tiles deep, for the largest grid
        const size_t stash_bytes = (size_t)kNumSMs * c->k1_blocks_per_sm * 2 * K1_CTHREADS * K1_STASH * sizeof(K1Stash);
        CUC(cudaMalloc(&c->d_stash.p, stash_bytes)); c->d_stash.cap = stash_bytes;
        const size_t ncomp = 1 + (size_t)c->nkey, grid = (size_t)kNumSMs * c->k1_blocks_per_sm;
        CUC(cudaMalloc(&c->d_seg_cnt.p, grid * ncomp * 4)); c->d_seg_cnt.cap = grid * ncomp * 4;
        CUC(cudaMalloc(&c->d_carry_out.p, ncomp * 4)); c->d_carry_out.cap = ncomp * 4;
        {
            int k4bps = 0, nsm = kNumSMs;
            CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k4bps, k4n_sweeps_kernel, K4_THREADS, 0));
            CUC(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
            c->k4_grid_max = std::max(1, k4bps) * nsm;
            CUC(cudaMalloc(&c->d_k4sync.p, 64 + sizeof(K4Trace))); c->d_k4sync.cap = 64 + sizeof(K4Trace);
        }
        if (const char* e = getenv("BDK_K4_HOST_LOOP")) c->k4_host_loop = atoi(e) != 0;
        if (const char* e = getenv("BDK_K4W_CAP")) c->k4w_cap = (uint32_t)std::max(0, std::min(atoi(e), (int)K4W_CAP));
        if (const char* e = getenv("BDK_ROWS_GUESS")) c->rows_guess_min = (uint32_t)std::max(0, atoi(e));
        if (const char* e = getenv("BDK_SEG_CAP_MIN")) c->seg_cap_min = (uint32_t)std::max(1, atoi(e));   // tests: force the segment-overflow retry
    }
    int rc = reset_job(c);
    if (rc) { g_create_error = c->err; bdk_destroy(c); return rc; }
    CUC(cudaStreamSynchronize(c->stream));
#undef CUC
    *out = c;
    return 0;
}

int bdk_set_stream(bdk_ctx* c, void* s) {
    if (!c) return BDK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return 0;
}

int bdk_reset(bdk_ctx* c) {
    if (!c) return BDK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    return reset_job(c);
}

int bdk_push_device(bdk_ctx* c, const bdk_soa* cols, uint64_t n) {
    if (!c || !cols) return BDK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    const void* ptrs[10] = {cols->pos, cols->mpos, cols->tid, cols->mtid, cols->isize, cols->flag, cols->mapq, cols->rgid, cols->qlen, cols->qid};
    for (int i = 0; i < 10; ++i) {
        if (!ptrs[i] && n) return fail(c, BDK_ERR_ARG, "null column %d", i);
        if ((uintptr_t)ptrs[i] & 15) return fail(c, BDK_ERR_ARG, "device column %d is not 16-byte aligned", i);
    }
    return push_common(c, n, n, [&]() -> int { return launch_k1(c, *cols, n, (uint32_t)c->n_records, true); });
}

int bdk_push(bdk_ctx* c, const bdk_soa* h, uint64_t n) {
    if (!c || !h) return BDK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    const void* src[10] = {h->pos, h->mpos, h->tid, h->mtid, h->isize, h->flag, h->mapq, h->rgid, h->qlen, h->qid};
    static const size_t width[10] = {4, 4, 4, 4, 4, 2, 1, 2, 4, 8};
    for (int i = 0; i < 10; ++i) if (!src[i] && n) return fail(c, BDK_ERR_ARG, "null column %d", i);
    const uint64_t CH = (uint64_t)K1_TILE * 1024;   // 8 Mi records per chunk (multiple of the tile size)
    const uint64_t chunk_cap = std::min<uint64_t>(CH, div_up<uint64_t>(std::max<uint64_t>(n, 1), K1_TILE) * K1_TILE);
    // qlen / qid are read only for the anomalous 1-3 % of the records: when the caller's columns are pinned
    // (mapped) host memory the kernel reads those few values in place instead of copying 12 bytes per record.
    const void* side_dev[2] = {nullptr, nullptr};
    bool zero_copy = n > 0 && !getenv("BDK_NO_ZEROCOPY");
    for (int k = 0; k < 2 && zero_copy; ++k) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, src[8 + k]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
            cudaGetLastError();
            zero_copy = false;
        } else side_dev[k] = at.devicePointer;
    }
    const int ncopy = zero_copy ? 8 : 10;
    for (int b = 0; b < 2; ++b)
        for (int k = 0; k < ncopy; ++k) ENS(c->d_chunk[b][k], chunk_cap * width[k]);
    c->h2d_bytes = 0;
    return push_common(c, n, CH, [&]() -> int {
        // double-buffered: the copy stream fills chunk i+1 while K1 runs on chunk i
        CU(cudaEventRecord(c->ev_done[0], c->stream));
        CU(cudaEventRecord(c->ev_done[1], c->stream));
        tstart(c, T_H2D);   // spans copies + kernels of this push on the compute stream
        uint64_t off = 0; int i = 0;
        c->h2d_bytes = 0;
        while (off < n) {
            const uint64_t m = std::min<uint64_t>(CH, n - off);
            const int b = i & 1;
            CU(cudaStreamWaitEvent(c->copy_stream, c->ev_done[b], 0));
            for (int k = 0; k < ncopy; ++k) {
                CU(cudaMemcpyAsync(c->d_chunk[b][k].p, (const char*)src[k] + off * width[k], m * width[k], cudaMemcpyHostToDevice, c->copy_stream));
                c->h2d_bytes += m * width[k];
            }
            CU(cudaEventRecord(c->ev_copied[b], c->copy_stream));
            CU(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
            bdk_soa d;
            d.pos = c->d_chunk[b][0].as<int32_t>(); d.mpos = c->d_chunk[b][1].as<int32_t>(); d.tid = c->d_chunk[b][2].as<int32_t>();
            d.mtid = c->d_chunk[b][3].as<int32_t>(); d.isize = c->d_chunk[b][4].as<int32_t>(); d.flag = c->d_chunk[b][5].as<uint16_t>();
            d.mapq = c->d_chunk[b][6].as<uint8_t>(); d.rgid = c->d_chunk[b][7].as<uint16_t>();
            if (zero_copy) { d.qlen = (const int32_t*)side_dev[0] + off; d.qid = (const uint64_t*)side_dev[1] + off; }
            else { d.qlen = c->d_chunk[b][8].as<int32_t>(); d.qid = c->d_chunk[b][9].as<uint64_t>(); }
            int rc = launch_k1(c, d, m, (uint32_t)(c->n_records + off), false);
            if (rc) return rc;
            CU(cudaEventRecord(c->ev_done[b], c->stream));
            off += m; ++i;
        }
        tstop(c, T_H2D);
        return 0;
    });
}

struct bdk_packed_buf {
    void* meta = nullptr; void* rel = nullptr; void* x = nullptr;      // pinned
};

void bdk_pack_free(bdk_packed_buf* b) {
    if (!b) return;
    if (b->meta) cudaFreeHost(b->meta);
    if (b->rel) cudaFreeHost(b->rel);
    if (b->x) cudaFreeHost(b->x);
    delete b;
}

int bdk_pack(const bdk_soa* h, uint64_t n, int threads, bdk_packed_buf** out, bdk_packed* view) {
    bdk_ctx* c = nullptr;
    if (!h || !out || !view) return fail(c, BDK_ERR_ARG, "null argument");
    *out = nullptr;
    memset(view, 0, sizeof *view);
    if (n > 0xffffffffull) return fail(c, BDK_ERR_ARG, "more than 2^32 records in one run");
    bdk_packed_buf* b = new bdk_packed_buf;
    if (cudaHostAlloc(&b->meta, std::max<uint64_t>(n, 4) * 4, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc(&b->rel, std::max<uint64_t>(n, 4) * 4, cudaHostAllocDefault) != cudaSuccess) {
        bdk_pack_free(b);
        return fail(c, BDK_ERR_CUDA, "cudaHostAlloc failed in bdk_pack: %s", cudaGetErrorString(cudaGetLastError()));
    }
    uint32_t* meta = (uint32_t*)b->meta; uint32_t* rel = (uint32_t*)b->rel;
    const int32_t tid = n ? h->tid[0] : 0;
    int T = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::max<uint64_t>(1, std::min<uint64_t>(T, n / 65536 + 1));
    struct Exc { uint32_t index; int32_t mpos, mtid, isize; uint16_t flag, rgid; };
    std::vector<std::vector<Exc>> exc(T);
    std::vector<int> bad(T, 0);
    auto work = [&](int t) {
        const uint64_t lo = n * t / T, hi = n * (t + 1) / T;
        for (uint64_t i = lo; i < hi; ++i) {
            if (h->tid[i] != tid) { bad[t] = 1; return; }
            const int32_t is = h->isize[i];
            const int64_t dm = (int64_t)h->mpos[i] - h->pos[i];
            const uint32_t fl = h->flag[i], rg = h->rgid[i];
            const bool fits = h->mtid[i] == tid && is >= -32768 && is < 32768 && dm >= -32768 && dm < 32768 && fl < 4096u && rg < BDK_PACKED_EXCEPT;
            meta[i] = (fl & 0xfffu) | ((uint32_t)h->mapq[i] << 12) | ((fits ? rg : BDK_PACKED_EXCEPT) << 20);
            rel[i] = fits ? (((uint32_t)is & 0xffffu) | ((uint32_t)dm << 16)) : 0u;
            if (!fits) exc[t].push_back(Exc{(uint32_t)i, h->mpos[i], h->mtid[i], is, (uint16_t)fl, (uint16_t)rg});
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < T; ++t) if (bad[t]) { bdk_pack_free(b); return fail(c, BDK_ERR_ARG, "bdk_pack: the records of a run must lie on one reference sequence"); }
    uint64_t nx = 0;
    for (auto& e : exc) nx += e.size();
    const uint64_t nxa = (nx + 3) & ~3ull;                   // every exception array 16-byte aligned
    if (cudaHostAlloc(&b->x, std::max<uint64_t>(nxa, 4) * 20, cudaHostAllocDefault) != cudaSuccess) {
        bdk_pack_free(b);
        return fail(c, BDK_ERR_CUDA, "cudaHostAlloc failed in bdk_pack");
    }
    uint32_t* xi = (uint32_t*)b->x; int32_t* xm = (int32_t*)(xi + nxa); int32_t* xt = xm + nxa; int32_t* xs = xt + nxa;
    uint16_t* xf = (uint16_t*)(xs + nxa); uint16_t* xr = xf + nxa;
    uint64_t k = 0;
    for (auto& ev : exc) for (auto& e : ev) { xi[k] = e.index; xm[k] = e.mpos; xt[k] = e.mtid; xs[k] = e.isize; xf[k] = e.flag; xr[k] = e.rgid; ++k; }
    view->pos = h->pos; view->meta = meta; view->rel = rel; view->qlen = h->qlen; view->qid = h->qid; view->tid = tid; view->nx = nx;
    view->x_index = xi; view->x_mpos = xm; view->x_mtid = xt; view->x_isize = xs; view->x_flag = xf; view->x_rgid = xr;
    *out = b;
    return 0;
}

int bdk_push_packed(bdk_ctx* c, const bdk_packed* h, uint64_t n) {
    if (!c || !h) return BDK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    if (n && (!h->pos || !h->meta || !h->rel || !h->qlen || !h->qid)) return fail(c, BDK_ERR_ARG, "null array in bdk_packed");
    if (h->nx && (!h->x_index || !h->x_mpos || !h->x_mtid || !h->x_isize || !h->x_flag || !h->x_rgid)) return fail(c, BDK_ERR_ARG, "null exception array in bdk_packed");
    if (h->tid < 0 || h->tid >= c->P.ntid) return fail(c, BDK_ERR_ARG, "bdk_packed.tid out of range");
    const uint64_t CH = (uint64_t)K1_TILE * 1024;
    const uint64_t chunk_cap = std::min<uint64_t>(CH, div_up<uint64_t>(std::max<uint64_t>(n, 1), K1_TILE) * K1_TILE);
    static const size_t width[10] = {4, 4, 4, 4, 4, 2, 1, 2, 4, 8};
    const void* side[2] = {h->qlen, h->qid};
    const void* side_dev[2] = {nullptr, nullptr};
    bool zero_copy = n > 0 && !getenv("BDK_NO_ZEROCOPY");
    for (int k = 0; k < 2 && zero_copy; ++k) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, side[k]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) { cudaGetLastError(); zero_copy = false; }
        else side_dev[k] = at.devicePointer;
    }
    // exceptions of every chunk: ranges of the ascending x_index
    uint64_t max_x = 0;
    for (uint64_t off = 0; off < n; off += CH) {
        const uint32_t* lo = std::lower_bound(h->x_index, h->x_index + h->nx, (uint32_t)off);
        const uint32_t* hi = std::lower_bound(lo, h->x_index + h->nx, (uint32_t)std::min<uint64_t>(off + CH, 0xffffffffull));
        max_x = std::max<uint64_t>(max_x, (uint64_t)(hi - lo));
    }
    static const size_t xwidth[6] = {4, 4, 4, 4, 2, 2};
    for (int b = 0; b < 2; ++b) {
        for (int k = 0; k < (zero_copy ? 8 : 10); ++k) ENS(c->d_chunk[b][k], chunk_cap * width[k]);
        for (int k = 0; k < 3; ++k) ENS(c->d_pk[b][k], chunk_cap * 4);
        for (int k = 0; k < 6; ++k) ENS(c->d_px[b][k], (max_x + 8) * xwidth[k]);
    }
    c->h2d_bytes = 0;
    return push_common(c, n, CH, [&]() -> int {
        CU(cudaEventRecord(c->ev_done[0], c->stream));
        CU(cudaEventRecord(c->ev_done[1], c->stream));
        tstart(c, T_H2D);
        uint64_t off = 0; int i = 0;
        c->h2d_bytes = 0;
        const void* pk[3] = {h->pos, h->meta, h->rel};
        const void* xs[6] = {h->x_index, h->x_mpos, h->x_mtid, h->x_isize, h->x_flag, h->x_rgid};
        while (off < n) {
            const uint64_t m = std::min<uint64_t>(CH, n - off);
            const int b = i & 1;
            const uint32_t* xlo = std::lower_bound(h->x_index, h->x_index + h->nx, (uint32_t)off);
            const uint32_t* xhi = std::lower_bound(xlo, h->x_index + h->nx, (uint32_t)std::min<uint64_t>(off + m, 0xffffffffull));
            const uint64_t x0 = (uint64_t)(xlo - h->x_index), nxc = (uint64_t)(xhi - xlo);
            CU(cudaStreamWaitEvent(c->copy_stream, c->ev_done[b], 0));
            for (int k = 0; k < 3; ++k) {
                CU(cudaMemcpyAsync(c->d_pk[b][k].p, (const char*)pk[k] + off * 4, m * 4, cudaMemcpyHostToDevice, c->copy_stream));
                c->h2d_bytes += m * 4;
            }
            if (nxc)
                for (int k = 0; k < 6; ++k) {
                    CU(cudaMemcpyAsync(c->d_px[b][k].p, (const char*)xs[k] + x0 * xwidth[k], nxc * xwidth[k], cudaMemcpyHostToDevice, c->copy_stream));
                    c->h2d_bytes += nxc * xwidth[k];
                }
            if (!zero_copy)
                for (int k = 8; k < 10; ++k) {
                    CU(cudaMemcpyAsync(c->d_chunk[b][k].p, (const char*)side[k - 8] + off * width[k], m * width[k], cudaMemcpyHostToDevice, c->copy_stream));
                    c->h2d_bytes += m * width[k];
                }
            CU(cudaEventRecord(c->ev_copied[b], c->copy_stream));
            CU(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
            bdk_soa d;
            d.pos = c->d_chunk[b][0].as<int32_t>(); d.mpos = c->d_chunk[b][1].as<int32_t>(); d.tid = c->d_chunk[b][2].as<int32_t>();
            d.mtid = c->d_chunk[b][3].as<int32_t>(); d.isize = c->d_chunk[b][4].as<int32_t>(); d.flag = c->d_chunk[b][5].as<uint16_t>();
            d.mapq = c->d_chunk[b][6].as<uint8_t>(); d.rgid = c->d_chunk[b][7].as<uint16_t>();
            if (zero_copy) { d.qlen = (const int32_t*)side_dev[0] + off; d.qid = (const uint64_t*)side_dev[1] + off; }
            else { d.qlen = c->d_chunk[b][8].as<int32_t>(); d.qid = c->d_chunk[b][9].as<uint64_t>(); }
            k1_expand_kernel<<<kNumSMs * 8, 256, 0, c->stream>>>(c->d_pk[b][0].as<int32_t>(), c->d_pk[b][1].as<uint32_t>(), c->d_pk[b][2].as<uint32_t>(), m, h->tid,
                (uint16_t)c->pad_rg, c->d_chunk[b][0].as<int32_t>(), c->d_chunk[b][1].as<int32_t>(), c->d_chunk[b][2].as<int32_t>(), c->d_chunk[b][3].as<int32_t>(),
                c->d_chunk[b][4].as<int32_t>(), c->d_chunk[b][5].as<uint16_t>(), c->d_chunk[b][6].as<uint8_t>(), c->d_chunk[b][7].as<uint16_t>());
            if (nxc)
                k1_expand_exceptions_kernel<<<(unsigned)std::min<uint64_t>(div_up<uint64_t>(nxc, 256), kNumSMs * 4), 256, 0, c->stream>>>(
                    c->d_px[b][0].as<uint32_t>(), c->d_px[b][1].as<int32_t>(), c->d_px[b][2].as<int32_t>(), c->d_px[b][3].as<int32_t>(), c->d_px[b][4].as<uint16_t>(),
                    c->d_px[b][5].as<uint16_t>(), nxc, (uint32_t)off, c->d_chunk[b][1].as<int32_t>(), c->d_chunk[b][3].as<int32_t>(), c->d_chunk[b][4].as<int32_t>(),
                    c->d_chunk[b][5].as<uint16_t>(), c->d_chunk[b][7].as<uint16_t>());
            c->launches += nxc ? 2 : 1;
            CU(cudaGetLastError());
            int rc = launch_k1(c, d, m, (uint32_t)(c->n_records + off), false);
            if (rc) return rc;
            CU(cudaEventRecord(c->ev_done[b], c->stream));
            off += m; ++i;
        }
        tstop(c, T_H2D);
        return 0;
    });
}

static int run_finalize(bdk_ctx* c) {
    if (c->comm && !c->exchanged) { int rc = comm_exchange(c); if (rc) return rc; }
    if (c->summary_ready) return 0;
    CU(cudaMemsetAsync(c->d_cnt.p, 0, CNT_N * 4, c->stream));
    CU(cudaMemcpyAsync(c->d_cnt.as<uint32_t>() + CNT_A, &c->A, 4, cudaMemcpyHostToDevice, c->stream));
    char* acc = (char*)c->d_acc.p;
    FinalizeIn in{c->P.nlib, c->P.nbam, c->P.nrg, c->P.ntid, c->P.cn_lib, c->P.initial_window, c->d_blibs.as<bdk_lib>(),
                  c->d_rg_lib.as<int32_t>(), c->d_rg_bam.as<int32_t>(), (unsigned long long*)acc,
                  (uint32_t*)(acc + c->off_hist), (unsigned long long*)(acc + c->off_first), (unsigned long long*)(acc + c->off_last)};
    tstart(c, T_FINALIZE);
    finalize_kernel<<<1, 256, 0, c->stream>>>(in, c->n_records, c->d_cnt.as<uint32_t>(), c->d_summary.as<bdk_summary_t>(), c->d_density.as<float>());
    c->launches += 1;
    tstop(c, T_FINALIZE);
    CU(cudaGetLastError());
    c->summary_ready = true;
    return 0;
}

int bdk_summary(bdk_ctx* c, bdk_summary_t* out) {
    if (!c || !out) return BDK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    int rc = run_finalize(c);
    if (rc) return rc;
    CU(cudaMemcpyAsync(&c->h_summary, c->d_summary.p, sizeof(bdk_summary_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    tcollect(c);
    *out = c->h_summary;
    return 0;
}

int bdk_finish(bdk_ctx* c, bdk_result* out) {
    if (!c || !out) return BDK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    memset(out, 0, sizeof(*out));
    out->nkey = c->nkey;
    const int nkey = c->nkey, nlib = c->P.nlib, period = c->period;
    cudaStream_t st = c->stream;
    uint32_t* d_cnt = c->d_cnt.as<uint32_t>();
    int rc = run_finalize(c);     // multi-GPU: exchange 1 happens here, c->A becomes the global count
    if (rc) return rc;
    const uint32_t A = c->A;
    c->h_sv.clear(); c->h_lib_count.clear(); c->h_cn_count.clear(); c->h_copy_number.clear();
    c->h_regions.clear(); c->h_areads.clear(); c->h_read_region.clear(); c->h_sv_of_read.clear();
    memset(c->h_cnt, 0, sizeof(c->h_cnt));
    if (A == 0) {
        CU(cudaMemcpyAsync(&c->h_summary, c->d_summary.p, sizeof(bdk_summary_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        tcollect(c);
        c->finished = true;
        out->sv = c->h_sv.data(); out->lib_count = c->h_lib_count.data(); out->cn_count = c->h_cn_count.data(); out->copy_number = c->h_copy_number.data();
        return 0;
    }
    // Everything downstream is sized from A alone (regions <= A + 1, followed edges <= A / 2, call slots <= A / 2): no count
    // comes back to the host before the result does.
    const size_t A1 = (size_t)A + 2;
    { int rc2 = grow_out(c, A1); if (rc2) return rc2; }
    ENS(c->d_read_cand, A1 * 4); ENS(c->d_read_region, A1 * 4);
    ENS(c->d_mate, A1 * 4); ENS(c->d_sv_of_read, A1 * 4); ENS(c->d_cand_first, A1 * 4); ENS(c->d_cand_maxlen, A1 * 4);
    ENS(c->d_cand_info, A1 * sizeof(CandInfo)); ENS(c->d_cand_regs, A1 * 4); ENS(c->d_reg, A1 * sizeof(RegionRec));
    uint32_t tsize = 1024; while (tsize < 2 * (uint64_t)A) tsize <<= 1;
    ENS(c->d_table, (size_t)tsize * 4);
    const size_t nwin_cap = A1 / (size_t)period + 2;
    ENS(c->d_del, A1 * 4); ENS(c->d_stamp, A1 * 4); ENS(c->d_c1, A1 * 4); ENS(c->d_ri, A1 * sizeof(ReadInfo2));
    ENS(c->d_se, A1 * 8); ENS(c->d_se2, 2 * A1 * 8); ENS(c->d_sefl, A1); ENS(c->d_queue, (A1 + nwin_cap) * 4);
    ENS(c->d_wstart, nwin_cap * 4); ENS(c->d_wdir, nwin_cap * 4); ENS(c->d_wund, nwin_cap * 4); ENS(c->d_wfill, nwin_cap * 4);
    ENS(c->d_slot_base, nwin_cap * 4); ENS(c->d_biglist, nwin_cap * 4);

    // ---- K2 ----------------------------------------------------------------------------------
    const int dummy = dummy_region_of(c->P);
    ScanScratch ssc{c->d_scan_sums.as<uint32_t>()};
    tstart(c, T_K2);
    device_scan(st, BreakFlag{c->d_ar.as<bdk_aread>(), c->d_summary.as<bdk_summary_t>()},
                BreakOut{c->d_read_cand.as<int32_t>(), c->d_cand_first.as<uint32_t>()}, d_cnt + CNT_A, d_cnt + CNT_NCAND, 0, ssc);
    k2_candidates_kernel<<<GS_GRID, GS_THREADS, 0, st>>>(c->d_ar.as<bdk_aread>(), c->d_cand_first.as<uint32_t>(), d_cnt, c->P.min_len,
                                                         c->P.seq_coverage_lim, c->d_cand_maxlen.as<int32_t>(), c->d_cand_info.as<CandInfo>());
    device_scan(st, AcceptFlag{c->d_cand_info.as<CandInfo>()},
                RegionOut{c->d_ar.as<bdk_aread>(), c->d_cand_first.as<uint32_t>(), c->d_cand_info.as<CandInfo>(), d_cnt, c->d_reg.as<RegionRec>(),
                          c->d_read_region.as<int32_t>(), c->d_cand_regs.as<int32_t>(), dummy, c->P.chr_restricted, c->P.min_read_pair},
                d_cnt + CNT_NCAND, d_cnt + CNT_NREG, (uint32_t)dummy, ssc);
    c->launches += 3 + 1 + 3;   // two scans (3 kernels each) + the candidate kernel
    tstop(c, T_K2);
    CU(cudaGetLastError());

    // ---- K3 ----------------------------------------------------------------------------------
    tstart(c, T_K3);
    uint32_t esize = 1024; while (esize < (uint64_t)A + 2) esize <<= 1;    // edge table: at most A / 2 distinct (region, region) keys
    ENS(c->d_ekeys, (size_t)esize * 8); ENS(c->d_ecnt, (size_t)esize * 4);
    CU(cudaMemsetAsync(c->d_table.p, 0xff, (size_t)tsize * 4, st));
    CU(cudaMemsetAsync(c->d_mate.p, 0xff, (size_t)A * 4, st));
    CU(cudaMemsetAsync(c->d_ekeys.p, 0xff, (size_t)esize * 8, st));
    CU(cudaMemsetAsync(c->d_ecnt.p, 0, (size_t)esize * 4, st));
    CU(cudaMemsetAsync(c->d_wdir.p, 0, nwin_cap * 4, st));
    CU(cudaMemsetAsync(c->d_wund.p, 0, nwin_cap * 4, st));
    unsigned long long* ekeys = c->d_ekeys.as<unsigned long long>();
    uint32_t* ecnt = c->d_ecnt.as<uint32_t>();
    CU(cudaMemsetAsync(c->d_sefl.p, 0, A1, st));            // (flags of names with more than two reads; the big-window kernel reuses the bytes later)
    k3_mate_join_kernel<<<GS_GRID, GS_THREADS, 0, st>>>(c->d_ar.as<bdk_aread>(), A, c->d_table.as<uint32_t>(), tsize - 1, c->d_mate.as<int32_t>(), c->d_sefl.as<uint8_t>(), d_cnt);
    k3_links_kernel<<<GS_GRID, GS_THREADS, 0, st>>>(c->d_mate.as<int32_t>(), c->d_sefl.as<uint8_t>(), c->d_read_region.as<int32_t>(), A, ekeys, ecnt, esize - 1);
    k3_read_info_kernel<<<GS_GRID, GS_THREADS, 0, st>>>(c->d_ar.as<bdk_aread>(), c->d_mate.as<int32_t>(), c->d_read_region.as<int32_t>(), c->d_read_cand.as<int32_t>(),
        c->d_reg.as<RegionRec>(), c->d_cand_regs.as<int32_t>(), d_cnt, period, c->P.min_read_pair, ekeys, ecnt, esize - 1, A, c->d_ri.as<ReadInfo2>(), c->d_sv_of_read.as<int32_t>());
    // the followed edges, bucketed by flush window
    k3_strong_count_kernel<<<GS_GRID, GS_THREADS, 0, st>>>(ekeys, ecnt, esize, c->P.min_read_pair, period, c->d_wdir.as<uint32_t>(), c->d_wund.as<uint32_t>());
    k3_window_scan_kernel<<<1, K3S_THREADS, 0, st>>>(c->d_wdir.as<uint32_t>(), c->d_wund.as<uint32_t>(), period, c->d_wstart.as<uint32_t>(),
                                                     c->d_slot_base.as<uint32_t>(), c->d_wfill.as<uint32_t>(), d_cnt);
    k3_strong_scatter_kernel<<<GS_GRID, GS_THREADS, 0, st>>>(ekeys, ecnt, esize, c->P.min_read_pair, period, c->d_wstart.as<uint32_t>(), c->d_wfill.as<uint32_t>(),
                                                             c->d_se.as<unsigned long long>());
    c->launches += 6;
    tstop(c, T_K3);
    CU(cudaGetLastError());

    // ---- K4 ----------------------------------------------------------------------------------
    // per-call outputs by slot (K4 writes them) and, after compaction, by output position (copied to a pinned host
    // block that the result pointers then refer to: the host never touches the rows)
    const size_t R1 = (size_t)A / 2 + 2;                       // call slots: one per followed edge
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t o_rows = 0, o_lc = al(o_rows + R1 * sizeof(bdk_sv)), o_cc = al(o_lc + R1 * 4 * nlib), o_cn = al(o_cc + R1 * 4 * nkey),
                 out_bytes = al(o_cn + R1 * 4 * nkey);                                                             // compacted block (device)
    const size_t w_emit = al(out_bytes), w_span = al(w_emit + R1), work_bytes = al(w_span + R1 * 4 * nlib);   // slot-indexed block
    ENS(c->d_rowpack, work_bytes); ENS(c->d_outpack, out_bytes); ENS(c->d_slot_order, R1 * 4);
    char* dp = (char*)c->d_rowpack.p;
    char* op = (char*)c->d_outpack.p;
    CU(cudaMemsetAsync(dp + w_emit, 0, R1, st));
    K4N KS;
    KS.ri = c->d_ri.as<ReadInfo2>(); KS.ar = c->d_ar.as<bdk_aread>(); KS.reg = c->d_reg.as<RegionRec>();
    KS.nreg = 0; KS.period = period; KS.chr_restricted = c->P.chr_restricted; KS.min_read_pair = c->P.min_read_pair;
    K4Tab Tb;
    Tb.del = c->d_del.as<int32_t>(); Tb.stamp = c->d_stamp.as<uint32_t>(); Tb.c1 = c->d_c1.as<int32_t>(); Tb.d_cnt = d_cnt;
    Tb.count_changes = getenv("BDK_K4_TRACE") ? 1 : 0;
    K4Static S;
    memset(&S, 0, sizeof S);
    S.ar = c->d_ar.as<bdk_aread>(); S.reg = c->d_reg.as<RegionRec>(); S.P = c->d_P.as<uint32_t>();
    S.cand_maxlen = c->d_cand_maxlen.as<int32_t>(); S.lib_mean = c->d_lib_mean.as<float>();
    S.hist = (uint32_t*)((char*)c->d_acc.p + c->off_hist); S.density = c->d_density.as<float>();
    S.A = A; S.period = period; S.nkey = nkey; S.nlib = nlib;
    S.chr_restricted = c->P.chr_restricted; S.min_read_pair = c->P.min_read_pair; S.score_threshold = c->P.score_threshold; S.fisher = c->P.fisher;
    K4Mut M;
    M.rows = (bdk_sv*)(dp + o_rows); M.row_lib_count = (int32_t*)(dp + o_lc); M.row_lib_span = (int32_t*)(dp + w_span);
    M.row_cn_count = (uint32_t*)(dp + o_cc); M.row_cn = (float*)(dp + o_cn); M.row_emit = (uint8_t*)(dp + w_emit);
    K4NOut KO{c->d_sv_of_read.as<int32_t>(), M.rows, M.row_lib_count, M.row_lib_span, M.row_emit, nlib};
    tstart(c, T_K4);
    c->k4_sweeps = 0;
    bool sweeps_on_device = false;
    uint32_t* sync = c->d_k4sync.as<uint32_t>();
    K4Trace* trace = getenv("BDK_K4_TRACE") ? (K4Trace*)((char*)c->d_k4sync.p + 64) : nullptr;
    {
        const unsigned rgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(div_up<uint64_t>(A1, 32 * (K4_THREADS / 32)), (uint64_t)kNumSMs * 8));
        bool host_loop = c->k4_host_loop;
        CU(cudaMemsetAsync(sync, 0, 64, st));
        CU(cudaMemsetAsync(c->d_del.p, 0x7f, A1 * 4, st));           // K4_NEVER everywhere: the sweeps start from an empty table
        CU(cudaMemsetAsync(c->d_stamp.p, 0, A1 * 4, st));
        if (trace) CU(cudaMemsetAsync(trace, 0, sizeof(K4Trace), st));
        if (!host_loop) {           // one persistent cooperative kernel, a grid-wide barrier between the sweeps
            const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(div_up<uint64_t>(A1, 32 * (K4_THREADS / 32)), (uint64_t)c->k4_grid_max));
            void* args[] = {&KS, &Tb, &sync, &trace};
            const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)k4n_sweeps_kernel, dim3(grid), dim3(K4_THREADS), args, 0, st);
            if (ce == cudaSuccess) { c->launches += 1; sweeps_on_device = true; }
            else if (ce == cudaErrorCooperativeLaunchTooLarge || ce == cudaErrorNotSupported || ce == cudaErrorLaunchOutOfResources) {
                cudaGetLastError();      // the GPU is shared (MPS, another context): the CTAs cannot all be resident. One launch per sweep instead.
                host_loop = true;
            } else return fail(c, BDK_ERR_CUDA, "cudaLaunchCooperativeKernel(k4n_sweeps_kernel) failed: %s", cudaGetErrorString(ce));
        }
        if (host_loop) {
            for (uint32_t sweep = 0;; ++sweep) {
                if (sweep > 1000000) return fail(c, BDK_ERR_STATE, "the table of deletion windows did not reach a fixed point");
                CU(cudaMemsetAsync(d_cnt + CNT_K4_CHANGED, 0, 4, st));
                k4n_sweep_kernel<<<rgrid, K4_THREADS, 0, st>>>(KS, Tb, sweep, d_cnt + CNT_K4_CHANGED);
                c->launches += 1;
                uint32_t nchanged = 0;
                CU(cudaMemcpyAsync(&nchanged, d_cnt + CNT_K4_CHANGED, 4, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                c->k4_sweeps = sweep + 1;
                if (!nchanged) break;
            }
        }
        k4n_first_call_kernel<<<rgrid, K4_THREADS, 0, st>>>(KS, Tb);
        K4Windows W;
        W.se = c->d_se.as<unsigned long long>(); W.wstart = c->d_wstart.as<uint32_t>(); W.wdir = c->d_wdir.as<uint32_t>(); W.slot_base = c->d_slot_base.as<uint32_t>();
        W.scratch = c->d_se2.as<unsigned long long>(); W.fl = c->d_sefl.as<uint8_t>(); W.queue = c->d_queue.as<int32_t>();
        W.big_list = c->d_biglist.as<uint32_t>(); W.big_count = d_cnt + CNT_K4_NBIGWIN;
        W.rows = M.rows; W.row_emit = M.row_emit; W.period = period; W.cap = c->k4w_cap;
        const unsigned wgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(div_up<uint64_t>(nwin_cap, K4W_WARPS), (uint64_t)kNumSMs * 16));
        k4n_windows_kernel<<<wgrid, K4W_WARPS * 32, 0, st>>>(Tb, W);
        k4n_big_windows_kernel<<<kNumSMs, K4WB_THREADS, 0, st>>>(Tb, W);
        const unsigned cgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(div_up<uint64_t>(R1, K4_THREADS / 8), (uint64_t)kNumSMs * 8));
        k4n_calls_kernel<<<cgrid, K4_THREADS, 0, st>>>(KS, KO, d_cnt);
        k4_score_kernel<<<GS_GRID, GS_THREADS, 0, st>>>(S, M, c->d_summary.as<bdk_summary_t>(), d_cnt);
        c->launches += 5;
    }
    tstop(c, T_K4);
    CU(cudaGetLastError());

    // ---- output order + results to the host -----------------------------------------------------------
    tstart(c, T_D2H);
    RowPack in{M.rows, M.row_lib_count, M.row_cn_count, M.row_cn};
    RowPack outp{(bdk_sv*)(op + o_rows), (int32_t*)(op + o_lc), (uint32_t*)(op + o_cc), (float*)(op + o_cn)};
    device_scan(st, EmitFlag{M.row_emit}, GatherOut{in, outp, c->d_slot_order.as<int32_t>(), nlib, nkey}, d_cnt + CNT_NROW, d_cnt + CNT_NEMIT, 0, ssc);
    c->launches += 3;
    CU(cudaGetLastError());
    // The number of rows is only known on the device: the first copy brings the counters, the summary and the first `guess`
    // rows (sized from the previous job on this context, else from A); a second copy follows only if there are more.
    const size_t guess = std::min<size_t>(R1, std::max<size_t>(c->rows_guess, c->rows_guess_min == 1024 ? (size_t)A / 64 + 1024 : (size_t)c->rows_guess_min));
    auto host_layout = [&](size_t cap, size_t* h_lc, size_t* h_cc, size_t* h_cn, size_t* h_cnt, size_t* h_sum, size_t* h_sync) {
        *h_lc = al(cap * sizeof(bdk_sv)); *h_cc = al(*h_lc + cap * 4 * nlib); *h_cn = al(*h_cc + cap * 4 * nkey);
        *h_cnt = al(*h_cn + cap * 4 * nkey); *h_sum = al(*h_cnt + CNT_N * 4); *h_sync = al(*h_sum + sizeof(bdk_summary_t));
        return al(*h_sync + 64);
    };
    size_t h_lc, h_cc, h_cn, h_cnt, h_sum, h_sync;
    size_t cap = guess;
    size_t host_bytes = host_layout(cap, &h_lc, &h_cc, &h_cn, &h_cnt, &h_sum, &h_sync);
    auto ensure_host = [&](size_t bytes) -> int {
        if (c->h_pack_cap >= bytes) return 0;
        if (c->h_pack) cudaFreeHost(c->h_pack);
        c->h_pack = nullptr; c->h_pack_cap = 0;
        const size_t want = bytes * 5 / 4 + 4096;
        CU(cudaHostAlloc(&c->h_pack, want, cudaHostAllocDefault));
        c->h_pack_cap = want;
        return 0;
    };
    { int rc2 = ensure_host(host_bytes); if (rc2) return rc2; }
    char* hp = (char*)c->h_pack;
    auto copy_rows = [&](size_t from, size_t to) -> int {       // rows [from, to) of the four arrays
        if (to <= from) return 0;
        CU(cudaMemcpyAsync(hp + from * sizeof(bdk_sv), op + o_rows + from * sizeof(bdk_sv), (to - from) * sizeof(bdk_sv), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hp + h_lc + from * 4 * nlib, op + o_lc + from * 4 * nlib, (to - from) * 4 * nlib, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hp + h_cc + from * 4 * nkey, op + o_cc + from * 4 * nkey, (to - from) * 4 * nkey, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hp + h_cn + from * 4 * nkey, op + o_cn + from * 4 * nkey, (to - from) * 4 * nkey, cudaMemcpyDeviceToHost, st));
        c->d2h_bytes += (to - from) * (sizeof(bdk_sv) + 4 * (size_t)nlib + 8 * (size_t)nkey);
        return 0;
    };
    c->d2h_bytes = CNT_N * 4 + sizeof(bdk_summary_t) + 64;
    CU(cudaMemcpyAsync(hp + h_cnt, d_cnt, CNT_N * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hp + h_sum, c->d_summary.p, sizeof(bdk_summary_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hp + h_sync, sync, 64, cudaMemcpyDeviceToHost, st));
    { int rc2 = copy_rows(0, guess); if (rc2) return rc2; }
    tstop(c, T_D2H);
    CU(cudaStreamSynchronize(st));
    tcollect(c);
    memcpy(c->h_cnt, hp + h_cnt, CNT_N * 4);
    memcpy(&c->h_summary, hp + h_sum, sizeof(bdk_summary_t));
    if (sweeps_on_device) c->k4_sweeps = ((const uint32_t*)(hp + h_sync))[7];
    c->dup_names = c->h_cnt[CNT_NDUP];
    const size_t ns = c->h_cnt[CNT_NEMIT];
    if (c->h_cnt[CNT_NROW] > R1 || c->h_cnt[CNT_NSE] > A1) return fail(c, BDK_ERR_STATE, "internal: more followed edges than reads");
    if (ns > cap) {            // more rows than the first copy brought: lay the host block out again and fetch everything
        cap = ns;
        host_bytes = host_layout(cap, &h_lc, &h_cc, &h_cn, &h_cnt, &h_sum, &h_sync);
        { int rc2 = ensure_host(host_bytes); if (rc2) return rc2; }
        hp = (char*)c->h_pack;
        { int rc2 = copy_rows(0, ns); if (rc2) return rc2; }
        CU(cudaStreamSynchronize(st));
    }
    c->rows_guess = (uint32_t)std::min<size_t>(ns + ns / 8 + 64, 0xffffffffu);
    if (sweeps_on_device && trace) {
        K4Trace tr;
        CU(cudaMemcpy(&tr, trace, sizeof tr, cudaMemcpyDeviceToHost));
        fprintf(stderr, "bdk K4 trace: %u regions, %u directed followed edges, %u call slots, %u sweeps\n", c->h_cnt[CNT_NREG], c->h_cnt[CNT_NSE], c->h_cnt[CNT_NROW], c->k4_sweeps);
        for (uint32_t sw = 0; sw < std::min<uint32_t>(c->k4_sweeps, K4_TRACE_SWEEPS); ++sw)
            fprintf(stderr, "  sweep %2u: %7.1f us -> %u regions changed\n", sw, (tr.t[1 + sw] - tr.t[sw]) / 1e3, tr.nchanged[sw]);
    }
    c->h_sv_of_read.assign(1, -2);   // marker: not fetched yet
    c->n_slots = c->h_cnt[CNT_NROW];
    c->finished = true;
    out->n_sv = ns;
    out->sv = (const bdk_sv*)hp; out->lib_count = (const int32_t*)(hp + h_lc); out->cn_count = (const uint32_t*)(hp + h_cc);
    out->copy_number = (const float*)(hp + h_cn); out->nkey = nkey;
    return 0;
}

int bdk_get_regions(bdk_ctx* c, const bdk_region** regions, uint64_t* n) {
    if (!c || !regions || !n) return BDK_ERR_ARG;
    if (!c->finished) return fail(c, BDK_ERR_STATE, "bdk_get_regions before bdk_finish");
    CU(cudaSetDevice(c->device));
    const uint32_t nreg = c->A ? c->h_cnt[CNT_NREG] : 0;
    std::vector<RegionRec> reg(nreg);
    if (nreg) CU(cudaMemcpy(reg.data(), c->d_reg.p, (size_t)nreg * sizeof(RegionRec), cudaMemcpyDeviceToHost));
    c->h_regions.resize(nreg);
    for (uint32_t r = 0; r < nreg; ++r) {
        bdk_region& o = c->h_regions[r];
        o.tid = reg[r].tid; o.start = reg[r].start; o.end = reg[r].end; o.fwd = reg[r].fwd; o.rev = reg[r].rev;
        o.first_read = reg[r].first_read; o.n_reads = reg[r].n_reads; o.stored = reg[r].stored; o.window = (int32_t)(r / c->period);
    }
    *regions = c->h_regions.data(); *n = nreg;
    return 0;
}

int bdk_get_areads(bdk_ctx* c, const bdk_aread** reads, const int32_t** region_of_read, uint64_t* n) {
    if (!c || !reads || !region_of_read || !n) return BDK_ERR_ARG;
    if (!c->finished) return fail(c, BDK_ERR_STATE, "bdk_get_areads before bdk_finish");
    CU(cudaSetDevice(c->device));
    c->h_areads.resize(c->A); c->h_read_region.resize(c->A);
    if (c->A) {
        CU(cudaMemcpy(c->h_areads.data(), c->d_ar.p, (size_t)c->A * sizeof(bdk_aread), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(c->h_read_region.data(), c->d_read_region.p, (size_t)c->A * 4, cudaMemcpyDeviceToHost));
    }
    *reads = c->h_areads.data(); *region_of_read = c->h_read_region.data(); *n = c->A;
    return 0;
}

int bdk_get_support(bdk_ctx* c, const int32_t** sv_of_read, uint64_t* n) {
    if (!c || !sv_of_read || !n) return BDK_ERR_ARG;
    if (!c->finished) return fail(c, BDK_ERR_STATE, "bdk_get_support before bdk_finish");
    CU(cudaSetDevice(c->device));
    if (c->h_sv_of_read.size() == 1 && c->h_sv_of_read[0] == -2) {
        std::vector<int32_t> slots(c->A), slot_order(c->n_slots);
        if (c->A) CU(cudaMemcpy(slots.data(), c->d_sv_of_read.p, (size_t)c->A * 4, cudaMemcpyDeviceToHost));
        if (c->n_slots) CU(cudaMemcpy(slot_order.data(), c->d_slot_order.p, (size_t)c->n_slots * 4, cudaMemcpyDeviceToHost));
        for (auto& s : slots) s = (s >= 0 && (size_t)s < slot_order.size()) ? slot_order[s] : -1;
        c->h_sv_of_read.swap(slots);
    }
    *sv_of_read = c->h_sv_of_read.data(); *n = c->A;
    return 0;
}

int bdk_kernel_times(bdk_ctx* c, const char** names, float* ms, int* launches, int cap) {
    if (!c) return 0;
    int k = 0;
    for (int t = 0; t < T_N && k < cap; ++t, ++k) { names[k] = c->timers[t].name; ms[k] = c->timers[t].ms; launches[k] = c->timers[t].launches; }
    return k;
}

uint64_t bdk_kernel_launches(bdk_ctx* c) { return c ? c->launches : 0; }

uint64_t bdk_h2d_bytes(bdk_ctx* c) { return c ? c->h2d_bytes : 0; }
uint64_t bdk_d2h_bytes(bdk_ctx* c) { return c ? c->d2h_bytes : 0; }

static int attach_comm(bdk_ctx* c, ncclComm_t comm, int rank, int nranks, bool owned) {
    if (c->n_records || c->exchanged) return fail(c, BDK_ERR_STATE, "attach the communicator before the first bdk_push of a job");
    if (c->comm && c->comm_owned) { if (NcclApi* nc = nccl_api()) nc->CommDestroy(c->comm); }
    c->comm = comm; c->comm_owned = owned; c->rank = comm ? rank : 0; c->nranks = comm ? nranks : 1;
    return 0;
}

int bdk_set_comm(bdk_ctx* c, void* nccl_comm, int rank, int nranks) {
    if (!c) return BDK_ERR_ARG;
    if (nccl_comm && (nranks < 1 || rank < 0 || rank >= nranks || nranks > 1024)) return fail(c, BDK_ERR_ARG, "bad rank %d / nranks %d", rank, nranks);
    if (nccl_comm && !nccl_api()) return fail(c, BDK_ERR_NCCL, "%s", nccl_api_error());
    return attach_comm(c, (ncclComm_t)nccl_comm, rank, nranks, false);
}

int bdk_comm_unique_id(void* out, int cap) {
    bdk_ctx* c = nullptr;
    if (!out || cap < (int)sizeof(ncclUniqueId)) return fail(c, BDK_ERR_ARG, "bdk_comm_unique_id needs a %d-byte buffer", (int)sizeof(ncclUniqueId));
    NcclApi* nc = nccl_api();
    if (!nc) return fail(c, BDK_ERR_NCCL, "%s", nccl_api_error());
    ncclUniqueId id;
    NC(nc->GetUniqueId(&id));
    memcpy(out, &id, sizeof id);
    return 0;
}

int bdk_comm_init(bdk_ctx* c, const void* unique_id, int rank, int nranks) {
    if (!c || !unique_id) return BDK_ERR_ARG;
    if (nranks < 1 || rank < 0 || rank >= nranks || nranks > 1024) return fail(c, BDK_ERR_ARG, "bad rank %d / nranks %d", rank, nranks);
    NcclApi* nc = nccl_api();
    if (!nc) return fail(c, BDK_ERR_NCCL, "%s", nccl_api_error());
    CU(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    ncclComm_t comm = nullptr;
    NC(nc->CommInitRank(&comm, nranks, id, rank));
    return attach_comm(c, comm, rank, nranks, true);
}

uint64_t bdk_comm_bytes(bdk_ctx* c) { return c ? c->comm_bytes : 0; }
uint32_t bdk_k4_sweeps(bdk_ctx* c) { return c ? c->k4_sweeps : 0; }
uint32_t bdk_duplicate_names(bdk_ctx* c) { return c ? c->dup_names : 0; }

int bdk_poisson_logsf(bdk_ctx* c, const double* lambda, const int32_t* k, double* out, uint64_t n) {
    if (!c || !lambda || !k || !out) return BDK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    if (!n) return 0;
    ENS(c->d_pois_l, n * 8); ENS(c->d_pois_k, n * 4); ENS(c->d_pois_o, n * 8);
    CU(cudaMemcpyAsync(c->d_pois_l.p, lambda, n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_pois_k.p, k, n * 4, cudaMemcpyHostToDevice, c->stream));
    poisson_logsf_kernel<<<(unsigned)std::min<uint64_t>(div_up<uint64_t>(n, 128), 1184), 128, 0, c->stream>>>(c->d_pois_l.as<double>(), c->d_pois_k.as<int32_t>(), c->d_pois_o.as<double>(), n);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, c->d_pois_o.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- BGZF members inflated on the device (opt-in path of the BAM reader, csrc/host/bam_io.cpp) ---------------------------
static_assert(sizeof(bdk_bgzf_member) == sizeof(bgz::Member), "bdk_bgzf_member and bgz::Member must have one layout");
int bdk_bgzf_inflate(int device, const uint8_t* file, uint64_t file_bytes, const bdk_bgzf_member* members, uint64_t n_members,
                     uint8_t* out, uint64_t out_bytes, int32_t* status, float* kernel_ms) {
    if (!file || !members || !out || !status) return BDK_ERR_ARG;
    if (kernel_ms) *kernel_ms = 0.f;
    if (!n_members) return 0;
    for (uint64_t i = 0; i < n_members; ++i) {
        const bdk_bgzf_member& m = members[i];
        if (m.in_off > file_bytes || file_bytes - m.in_off < m.in_len || m.out_off > out_bytes || out_bytes - m.out_off < m.out_len) return BDK_ERR_ARG;
        if (i && (m.in_off < members[i - 1].in_off || m.out_off < members[i - 1].out_off)) return BDK_ERR_ARG;     // file order
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return BDK_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return BDK_ERR_CUDA;
    // the warp-per-member decoder of the device-resident BAM decode (bgzf_inflate_warp.cuh) and its CRC-32 check
    const size_t smem = sizeof(bgzw::Tables) * bgzw::WARPS_PER_CTA;
    if (cudaFuncSetAttribute(bgzw::bgzf_inflate_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return BDK_ERR_CUDA;
    // batches of members: at most 256 MiB of input and 1 GiB of output on the device at a time
    const uint64_t IN_CAP = 256ull << 20, OUT_CAP = 1ull << 30;
    uint8_t *d_in = 0, *d_out = 0; bgz::Member* d_mem = 0; uint32_t* d_counter = 0; int32_t* d_st = 0;
    cudaEvent_t e0 = 0, e1 = 0;
    cudaStream_t st = 0;
    std::vector<bgz::Member> batch;
    int rc = 0;
    auto ck = [&](cudaError_t e) { if (e != cudaSuccess && !rc) rc = BDK_ERR_CUDA; return e == cudaSuccess; };
    uint64_t max_members = 0;
    {   // size the buffers for the largest batch
        uint64_t i = 0;
        while (i < n_members) {
            uint64_t j = i, out_b = 0;
            while (j < n_members && (j == i || (members[j].in_off + members[j].in_len - members[i].in_off <= IN_CAP && out_b + members[j].out_len <= OUT_CAP))) {
                out_b += members[j].out_len; ++j;
            }
            max_members = std::max(max_members, j - i);
            i = j;
        }
    }
    ck(cudaStreamCreate(&st)); ck(cudaEventCreate(&e0)); ck(cudaEventCreate(&e1));
    ck(cudaMalloc(&d_in, std::min<uint64_t>(IN_CAP, file_bytes) + (1 << 17))); ck(cudaMalloc(&d_out, std::min<uint64_t>(OUT_CAP, out_bytes) + (1 << 17)));
    ck(cudaMalloc(&d_mem, max_members * sizeof(bgz::Member))); ck(cudaMalloc(&d_st, max_members * sizeof(int32_t)));
    ck(cudaMalloc(&d_counter, 4));
    uint64_t i = 0;
    while (!rc && i < n_members) {
        uint64_t j = i, in_b = 0, out_b = 0;
        batch.clear();
        while (j < n_members && (j == i || (members[j].in_off + members[j].in_len - members[i].in_off <= IN_CAP && out_b + members[j].out_len <= OUT_CAP))) {
            bgz::Member m;
            m.in_off = members[j].in_off - members[i].in_off; m.out_off = out_b; m.in_len = members[j].in_len; m.out_len = members[j].out_len;
            batch.push_back(m);
            in_b = members[j].in_off + members[j].in_len - members[i].in_off; out_b += members[j].out_len; ++j;
        }
        if (in_b > IN_CAP + (1 << 17) || out_b > OUT_CAP + (1 << 17)) { rc = BDK_ERR_ARG; break; }      // a member larger than BGZF allows
        // members are in file order but need not be contiguous in `out`: copy back member by member only if they are not
        bool contiguous = true;
        for (uint64_t k = i + 1; k < j; ++k) contiguous &= members[k].out_off == members[k - 1].out_off + members[k - 1].out_len;
        // (with the 8 bytes behind the last stream: the CRC-32 the device checks, and the words the bit reader looks ahead to)
        if (!ck(cudaMemcpyAsync(d_in, file + members[i].in_off, std::min<uint64_t>(in_b + 8, file_bytes - members[i].in_off), cudaMemcpyHostToDevice, st))) break;
        if (!ck(cudaMemcpyAsync(d_mem, batch.data(), batch.size() * sizeof(bgz::Member), cudaMemcpyHostToDevice, st))) break;
        ck(cudaEventRecord(e0, st));
        ck(cudaMemsetAsync(d_counter, 0, 4, st));
        const uint32_t nb = (uint32_t)batch.size();
        const unsigned grid = (unsigned)std::min<uint64_t>((nb + bgzw::WARPS_PER_CTA - 1) / bgzw::WARPS_PER_CTA, (uint64_t)kNumSMs * bgzw::CTAS_PER_SM);
        bgzw::bgzf_inflate_warp_kernel<<<grid, bgzw::CTA_THREADS, smem, st>>>(d_in, d_mem, nb, d_out, d_st, d_counter);
        bgzw::bgzf_crc_kernel<<<(unsigned)std::min<uint64_t>((nb + 7) / 8, (uint64_t)kNumSMs * 8), 256, 0, st>>>(d_in, d_mem, nb, d_out, d_st);
        ck(cudaGetLastError());
        ck(cudaEventRecord(e1, st));
        if (contiguous) ck(cudaMemcpyAsync(out + members[i].out_off, d_out, out_b, cudaMemcpyDeviceToHost, st));
        else for (uint64_t k = i; k < j && !rc; ++k) ck(cudaMemcpyAsync(out + members[k].out_off, d_out + batch[k - i].out_off, members[k].out_len, cudaMemcpyDeviceToHost, st));
        ck(cudaMemcpyAsync(status + i, d_st, batch.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (!ck(cudaStreamSynchronize(st))) break;
        float ms = 0.f;
        if (ck(cudaEventElapsedTime(&ms, e0, e1)) && kernel_ms) *kernel_ms += ms;
        i = j;
    }
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_mem); cudaFree(d_st); cudaFree(d_counter);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    return rc;
}

}  // extern "C"

#include "bdk_bam.inl"
