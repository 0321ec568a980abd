// bdk_bam.inl -- bdk_push_bam / bdk_decode_bam: a BAM file decoded on the device, window of BGZF members by window, and
// classified (part of bdk_core.cu). Only the COMPRESSED file crosses PCIe; inflated bytes, record offsets and the record columns
// live and die in HBM.
//
//   producer thread:  file bytes -> pinned staging buffer -> H2D (copy stream) -> bgzf_inflate_warp_kernel + bgzf_crc_kernel,
//                     CHUNK by chunk (16 MiB of compressed members) on a ring of inflate streams, up to WSLOTS windows ahead of
//   caller's thread:  per WINDOW (64 MiB of compressed members, ~200 MB inflated): chain_guess / chain_resolve (record
//                     boundaries) -> chain_write -> scan(keep) + extract (columns) -> sorted check -> K1 (the launch
//                     bdk_push_device makes), on the context's stream.
//
// Two granularities because the two sides want different ones: staging buffers are pinned memory (small, a few in flight) and an
// inflate launch should start as soon as its bytes are there -- several chunks run concurrently, so the GPU's 4736 warp slots
// are filled although one chunk has ~800 members; the record-boundary search and the scans are latency-bound launches with a
// fixed cost, so they want windows as large as memory allows. The inflate of window w + 1 overlaps the decode and classification
// of window w and the copies of window w + 2. A window's inflated bytes are written at a fixed offset of its buffer; the bytes of
// a record cut off by the end of window w are copied in front of window w + 1's (`carry`), so records are parsed where they lie.
#include "host/nway_merge.hpp"

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace {

struct BamDev {
    static constexpr int WSLOTS = 6;                        // most windows resident on the device (compressed + inflated); `wslots` are used
    static constexpr int PSLOTS = 4;                        // pinned staging buffers (one piece of PIECE bytes each)
    static constexpr size_t PIECE = 16u << 20;              // what is staged and copied at a time (a chunk is one or more pieces)
    cudaEvent_t ev_piece[PSLOTS] = {};                      // the copy out of a staging buffer is through
    static constexpr int NSTREAMS = 32;                     // inflate streams created; `nstreams` of them are used (chunks inflated concurrently)
    static constexpr size_t CARRY_CAP = 16u << 20;          // longest partial record carried between windows
    cudaStream_t inflate_stream[NSTREAMS] = {}, copy_stream = nullptr;
    void* h_slot[PSLOTS] = {};
    size_t h_cap[PSLOTS] = {};
    void* h_tab[WSLOTS] = {};                               // pinned member tables of the resident windows
    size_t h_tab_cap[WSLOTS] = {};
    DevBuf d_slot[WSLOTS], d_raw[WSLOTS], d_status[WSLOTS];
    std::vector<cudaEvent_t> ev_copied, ev_inflated;        // per chunk, rings
    cudaEvent_t ev_decoded[WSLOTS] = {}, ev_first = nullptr, ev_last = nullptr;
    DevBuf d_seg, d_base, d_recoff, d_cols[10], d_info, d_rgtab, d_prev, d_counter;
    bamdev::WinInfo* h_info = nullptr;                      // pinned: WinInfo + kept count
    bool ready = false;
};

void bamdev_free(BamDev* B) {
    if (!B) return;
    for (int s = 0; s < BamDev::PSLOTS; ++s) { if (B->h_slot[s]) cudaFreeHost(B->h_slot[s]); if (B->ev_piece[s]) cudaEventDestroy(B->ev_piece[s]); }
    for (int s = 0; s < BamDev::WSLOTS; ++s) {
        if (B->h_tab[s]) cudaFreeHost(B->h_tab[s]);
        for (DevBuf* b : {&B->d_slot[s], &B->d_raw[s], &B->d_status[s]}) if (b->p) cudaFree(b->p);
        if (B->ev_decoded[s]) cudaEventDestroy(B->ev_decoded[s]);
    }
    for (cudaEvent_t e : B->ev_copied) cudaEventDestroy(e);
    for (cudaEvent_t e : B->ev_inflated) cudaEventDestroy(e);
    if (B->ev_first) cudaEventDestroy(B->ev_first);
    if (B->ev_last) cudaEventDestroy(B->ev_last);
    for (DevBuf* b : {&B->d_seg, &B->d_base, &B->d_recoff, &B->d_info, &B->d_rgtab, &B->d_prev, &B->d_counter}) if (b->p) cudaFree(b->p);
    for (int k = 0; k < 10; ++k) if (B->d_cols[k].p) cudaFree(B->d_cols[k].p);
    if (B->h_info) cudaFreeHost(B->h_info);
    for (int k = 0; k < BamDev::NSTREAMS; ++k) if (B->inflate_stream[k]) cudaStreamDestroy(B->inflate_stream[k]);
    if (B->copy_stream) cudaStreamDestroy(B->copy_stream);
    delete B;
}

struct BamChunk { uint64_t m0, m1, in_begin, in_bytes; uint32_t window; };
struct BamWindow { uint64_t m0, m1, in_begin, in_bytes, out_begin, out_bytes; uint32_t c0, c1; };

void bamdev_release(void* p) { bamdev_free((BamDev*)p); }

// the decoded records of one whole bam, kept on the device (two-bam runs: bdk_push_bams)
struct BamCollect {
    DevBuf cols[10];
    uint64_t n = 0;
    bamdev::Columns view() const {
        return bamdev::Columns{cols[0].as<int32_t>(), cols[1].as<int32_t>(), cols[2].as<int32_t>(), cols[3].as<int32_t>(), cols[4].as<int32_t>(),
                               cols[8].as<int32_t>(), cols[5].as<uint16_t>(), cols[7].as<uint16_t>(), cols[6].as<uint8_t>(), cols[9].as<uint64_t>()};
    }
    void release() { for (auto& b : cols) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; } n = 0; }
};

// A few threads that copy ranges of one buffer side by side (the staging of the compressed file into pinned memory): started
// once per pipeline, woken per piece.
class CopyPool {
public:
    explicit CopyPool(int n) : n_(std::max(1, n)) {
        for (int t = 1; t < n_; ++t) th_.emplace_back([this, t] { work(t); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> g(m_); quit_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    // fn(lo, hi) over [0, bytes) cut into one range per thread (fewer for small sizes); returns when all are done
    template <class F> void run(uint64_t bytes, F fn) {
        const int parts = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)n_, bytes >> 20));
        job_ = [&fn, bytes, parts](int t) { if (t < parts) fn(bytes * (uint64_t)t / parts, bytes * (uint64_t)(t + 1) / parts); };
        { std::lock_guard<std::mutex> g(m_); pending_ = n_ - 1; ++gen_; }
        cv_.notify_all();
        job_(0);
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return pending_ == 0; });
    }
private:
    void work(int t) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (quit_) return;
            }
            job_(t);
            { std::lock_guard<std::mutex> g(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::function<void(int)> job_;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool quit_ = false;
};

// host_out == nullptr && collect == nullptr: classify the records (bdk_push_bam); host_out: copy the decoded columns into the
// caller's host arrays of `cap` records and classify nothing (bdk_decode_bam); collect: keep them on the device (bdk_push_bams).
int bam_pipeline(bdk_ctx* c, const bdk_bam_source* src, bdk_bam_stats* stats, const bdk_soa* host_out, uint64_t cap, BamCollect* collect = nullptr) {
    if (!c || !src) return BDK_ERR_ARG;
    if (stats) memset(stats, 0, sizeof *stats);
    if (c->finished && !host_out && !collect) return fail(c, BDK_ERR_STATE, "bdk_push_bam after bdk_finish (call bdk_reset first)");
    if (!src->file || (!src->members && src->n_members)) return fail(c, BDK_ERR_ARG, "null file image or member list");
    if (src->n_rg && (!src->rg_hash || !src->rg_id)) return fail(c, BDK_ERR_ARG, "null read-group table");
    if (src->n_ref < 0) return fail(c, BDK_ERR_ARG, "n_ref < 0");
    CU(cudaSetDevice(c->device));
    const auto wall0 = std::chrono::steady_clock::now();
    const bdk_bgzf_member* M = src->members;
    const uint64_t nm = src->n_members;
    uint64_t total_out = 0;
    for (uint64_t i = 0; i < nm; ++i) {
        if (M[i].in_off > src->file_bytes || src->file_bytes - M[i].in_off < (uint64_t)M[i].in_len + 8) return fail(c, BDK_ERR_ARG, "member %llu lies outside the file image", (unsigned long long)i);
        if (M[i].out_off != total_out) return fail(c, BDK_ERR_ARG, "member %llu: out_off must be the running sum of the output lengths", (unsigned long long)i);
        if (i && M[i].in_off < M[i - 1].in_off + M[i - 1].in_len) return fail(c, BDK_ERR_ARG, "members must be in file order");
        if (M[i].out_len > (1u << 16) || M[i].in_len > (1u << 17)) return fail(c, BDK_ERR_ARG, "member %llu is larger than BGZF allows", (unsigned long long)i);
        total_out += M[i].out_len;
    }
    const uint64_t end_off = src->end_offset ? src->end_offset : total_out;
    if (src->first_record > end_off || end_off > total_out) return fail(c, BDK_ERR_ARG, "first_record / end_offset outside the inflated stream");
    if (end_off == src->first_record || nm == 0) { if (stats) stats->sorted = 1; return 0; }

    // ---- windows and their chunks --------------------------------------------------------------------------------
    // Sizes (compressed bytes). The decode of a window is a dozen dependent launches and two host round trips on the context's
    // stream while the inflate kernels of the next windows hold every warp slot: about 5 ms per window whatever its size, and
    // the windows are decoded one after the other -- with 64 MiB windows that chain, not the inflate, bounded the run (trace:
    // the producer waited 82 of 175 ms for window slots). So windows are as large as the file allows with a dozen of them left
    // to overlap, up to 256 MiB; an inflate launch is a 32 MiB chunk of a window (1600 members: eight in flight fill the warp
    // slots, and a chunk starts as soon as its bytes are there). Measured on a 2.7 GB file (sweeps b14 / b15 / b22 in profiles/).
    uint64_t WIN_IN = 64ull << 20, WIN_OUT = 1024ull << 20, CHUNK_IN = 0;
    {
        const uint64_t total_in = nm ? M[nm - 1].in_off + M[nm - 1].in_len - M[0].in_off : 0;
        WIN_IN = std::min<uint64_t>(256ull << 20, std::max<uint64_t>(16ull << 20, (total_in / 12 + (1ull << 20)) & ~((1ull << 20) - 1)));
    }
    if (src->window_bytes) WIN_IN = src->window_bytes;
    if (const char* e = getenv("BDK_BAMDEV_WINDOW_KB")) if (atoll(e) > 0) WIN_IN = (uint64_t)atoll(e) << 10;      // tests: many small windows
    CHUNK_IN = std::min<uint64_t>(32ull << 20, std::max<uint64_t>(4ull << 20, WIN_IN / 4));
    if (const char* e = getenv("BDK_BAMDEV_CHUNK_KB")) if (atoll(e) > 0) CHUNK_IN = (uint64_t)atoll(e) << 10;
    CHUNK_IN = std::min(CHUNK_IN, WIN_IN);
    std::vector<BamWindow> wins;
    std::vector<BamChunk> chunks;
    const bool sizes_given = src->window_bytes || getenv("BDK_BAMDEV_WINDOW_KB") || getenv("BDK_BAMDEV_CHUNK_KB");
    for (uint64_t i = 0; i < nm;) {
        if (M[i].out_off >= end_off) break;                  // members behind the end of the records
        BamWindow w{i, i, M[i].in_off, 0, M[i].out_off, 0, (uint32_t)chunks.size(), 0};
        // nothing overlaps the copy and inflate of the first window: the first two are small (16 and 64 MiB)
        uint64_t win_in = WIN_IN, chunk_in = CHUNK_IN;
        if (!sizes_given && wins.size() < 2) {
            win_in = std::min<uint64_t>(WIN_IN, (16ull << 20) << (2 * wins.size()));
            chunk_in = std::min<uint64_t>(CHUNK_IN, std::max<uint64_t>(4ull << 20, win_in / 4));
        }
        while (w.m1 < nm && M[w.m1].out_off < end_off &&
               (w.m1 == w.m0 || (M[w.m1].in_off + M[w.m1].in_len + 8 - w.in_begin <= win_in && w.out_bytes + M[w.m1].out_len <= WIN_OUT))) {
            w.in_bytes = M[w.m1].in_off + M[w.m1].in_len + 8 - w.in_begin;         // the CRC32 + ISIZE footer travels along
            w.out_bytes += M[w.m1].out_len;
            ++w.m1;
        }
        for (uint64_t j = w.m0; j < w.m1;) {
            BamChunk ch{j, j, M[j].in_off, 0, (uint32_t)wins.size()};
            while (ch.m1 < w.m1 && (ch.m1 == ch.m0 || M[ch.m1].in_off + M[ch.m1].in_len + 8 - ch.in_begin <= chunk_in)) {
                ch.in_bytes = M[ch.m1].in_off + M[ch.m1].in_len + 8 - ch.in_begin;
                ++ch.m1;
            }
            chunks.push_back(ch);
            j = ch.m1;
        }
        w.c1 = (uint32_t)chunks.size();
        wins.push_back(w);
        i = w.m1;
    }
    const size_t nwin = wins.size(), nchunk = chunks.size();
    if (nwin == 0 || src->first_record < wins[0].out_begin || src->first_record - wins[0].out_begin >= wins[0].out_bytes)
        return fail(c, BDK_ERR_ARG, "first_record must lie in the first window of members (pass the members from the one that holds it)");
    uint64_t max_in = 0, max_out = 0, max_members = 0, max_chunk = 0, max_chunks_per_win = 0;
    for (auto const& w : wins) {
        max_in = std::max(max_in, w.in_bytes); max_out = std::max(max_out, w.out_bytes); max_members = std::max(max_members, w.m1 - w.m0);
        max_chunks_per_win = std::max<uint64_t>(max_chunks_per_win, w.c1 - w.c0);
    }
    for (auto const& ch : chunks) max_chunk = std::max(max_chunk, ch.in_bytes);
    if (max_in > (1ull << 31) || max_out + BamDev::CARRY_CAP > 0xfff00000ull) return fail(c, BDK_ERR_ARG, "decode window too large");

    // ---- buffers -------------------------------------------------------------------------------------------------
    if (!c->bamdev) c->bamdev = new BamDev;
    BamDev* B = (BamDev*)c->bamdev;
    if (!B->ready) {
        // the inflate kernels fill every warp slot for as long as there are members; the record-boundary / extraction / classify
        // launches of the context's stream (highest priority, bdk_create) take the slots that come free first
        int prio_lo = 0, prio_hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        if (getenv("BDK_BAMDEV_NOPRIO")) prio_lo = prio_hi = 0;
        for (int k = 0; k < BamDev::NSTREAMS; ++k) CU(cudaStreamCreateWithPriority(&B->inflate_stream[k], cudaStreamNonBlocking, prio_lo));
        CU(cudaStreamCreateWithFlags(&B->copy_stream, cudaStreamNonBlocking));
        for (int s = 0; s < BamDev::WSLOTS; ++s) CU(cudaEventCreateWithFlags(&B->ev_decoded[s], cudaEventDisableTiming));
        for (int s = 0; s < BamDev::PSLOTS; ++s) CU(cudaEventCreateWithFlags(&B->ev_piece[s], cudaEventDisableTiming));
        CU(cudaEventCreate(&B->ev_first)); CU(cudaEventCreate(&B->ev_last));
        CU(cudaHostAlloc((void**)&B->h_info, 256, cudaHostAllocDefault));
        CU(cudaFuncSetAttribute(bgzw::bgzf_inflate_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(bgzw::Tables) * bgzw::WARPS_PER_CTA)));
        B->ready = true;
    }
    // chunk events: a ring long enough that an event is not recorded again before the window it belonged to is decoded
    const size_t ring = (size_t)(BamDev::WSLOTS + 1) * max_chunks_per_win + BamDev::PSLOTS;
    while (B->ev_copied.size() < ring) {
        cudaEvent_t a = nullptr, b = nullptr;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); B->ev_copied.push_back(a);
        CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming)); B->ev_inflated.push_back(b);
    }
    const size_t tab_bytes = (max_members * sizeof(bgz::Member) + 255) & ~size_t(255);
    int wslots = 4;               // windows resident: one being decoded, the others being inflated / copied
    if (const char* e = getenv("BDK_BAMDEV_WSLOTS")) wslots = std::max(2, std::min(atoi(e), (int)BamDev::WSLOTS));
    const int nslots = (int)std::min<size_t>((size_t)wslots, nwin);
    const uint64_t max_n = BamDev::CARRY_CAP + max_out;                 // bytes of a window incl. the carry
    const uint64_t max_seg = max_n / bamdev::SEG_BYTES + 2, max_rec = max_n / 36 + 2;
    for (int s = 0; s < nslots; ++s) {
        ENS(B->d_slot[s], tab_bytes + max_in + 64); ENS(B->d_raw[s], max_n + 1024); ENS(B->d_status[s], max_members * 4);   // (+1024: a damaged member may be written up to 512 bytes beyond its output)
        if (B->h_tab_cap[s] < tab_bytes) {
            if (B->h_tab[s]) cudaFreeHost(B->h_tab[s]);
            B->h_tab[s] = nullptr; B->h_tab_cap[s] = 0;
            CU(cudaHostAlloc(&B->h_tab[s], tab_bytes, cudaHostAllocDefault));
            B->h_tab_cap[s] = tab_bytes;
        }
    }
    ENS(B->d_seg, max_seg * sizeof(brec::Segment)); ENS(B->d_base, max_seg * 4);
    ENS(B->d_info, 256); ENS(B->d_prev, 16); ENS(B->d_counter, 4 * (nchunk + 1));
    static const size_t width[10] = {4, 4, 4, 4, 4, 2, 1, 2, 4, 8};
    {   // record offsets and columns: records per byte are not known yet, so start from a typical density; the window loop grows
        // them (it knows a window's record count before it extracts)
        const uint64_t guess = std::min<uint64_t>(max_rec, max_n / 96 + 1024);
        ENS(B->d_recoff, guess * 4);
        if (!collect) for (int k = 0; k < 10; ++k) ENS(B->d_cols[k], guess * width[k]);
    }
    // read-group table: open addressing, at most half full
    uint32_t rg_slots = 16;
    while (rg_slots < 2 * src->n_rg + 2) rg_slots <<= 1;
    {
        std::vector<bamdev::RgEntry> tab(rg_slots, bamdev::RgEntry{0, 0, 0});
        for (uint32_t i = 0; i < src->n_rg; ++i) {
            uint32_t p = (uint32_t)(src->rg_hash[i] % rg_slots);
            while (tab[p].used) {
                if (tab[p].hash == src->rg_hash[i]) return fail(c, BDK_ERR_ARG, "two read groups with one key");
                p = p + 1 == rg_slots ? 0 : p + 1;
            }
            tab[p] = bamdev::RgEntry{src->rg_hash[i], src->rg_id[i], 1};
            if ((int)src->rg_id[i] >= c->P.nrg) return fail(c, BDK_ERR_ARG, "rg_id[%u] out of range", i);
        }
        if ((int)src->rg_other >= c->P.nrg) return fail(c, BDK_ERR_ARG, "rg_other out of range");
        ENS(B->d_rgtab, rg_slots * sizeof(bamdev::RgEntry));
        CU(cudaMemcpyAsync(B->d_rgtab.p, tab.data(), rg_slots * sizeof(bamdev::RgEntry), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemsetAsync(B->d_counter.p, 0, 4 * (nchunk + 1), c->stream));
        const int32_t prev0[4] = {INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN};
        CU(cudaMemcpyAsync(B->d_prev.p, prev0, 16, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));                // tab / prev0 are locals; the producer's kernels need the zeroed counters
    }

    const bool trace = getenv("BDK_DECODE_TRACE") != nullptr;
    auto since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count() * 1e3; };
    if (trace) fprintf(stderr, "[bamdev] %.1f ms: %zu windows, %zu chunks planned, buffers ready\n", since(), nwin, nchunk);
    // ---- producer ----------------------------------------------------------------------------------------------------
    int nstreams = 8;             // chunks inflated concurrently (measured: 4 streams lose 20 % to empty warp slots, 16 change nothing)
    if (const char* e = getenv("BDK_BAMDEV_STREAMS")) nstreams = std::max(1, std::min(atoi(e), (int)BamDev::NSTREAMS));
    std::atomic<int64_t> launched(0), decoded(0);            // chunks whose inflate is launched; windows decoded
    std::atomic<int> producer_rc(0);
    std::atomic<bool> stop(false);
    std::string producer_err;
    double stage_s = 0, wait_pinned_s = 0, wait_slot_s = 0;
    int copy_threads = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("BDK_BAMDEV_COPY_THREADS")) copy_threads = std::max(1, std::min(atoi(e), 64));
    CopyPool copiers(copy_threads);
    uint64_t piece = 0;
    std::thread producer([&]() {
        auto bad = [&](const char* what, cudaError_t e) { producer_err = std::string(what) + ": " + cudaGetErrorString(e); producer_rc = BDK_ERR_CUDA; launched = (int64_t)nchunk + 1; };
        cudaError_t e = cudaSetDevice(c->device);
        if (e != cudaSuccess) return bad("cudaSetDevice", e);
        int last_on_stream[BamDev::NSTREAMS];
        for (int k = 0; k < BamDev::NSTREAMS; ++k) last_on_stream[k] = -1;
        for (size_t g = 0; g < nchunk && !stop; ++g) {
            BamChunk const& C = chunks[g];
            BamWindow const& W = wins[C.window];
            const int ws = (int)(C.window % wslots);
            cudaStream_t ist = B->inflate_stream[g % nstreams];
            const size_t ev = g % ring;
            const auto tw1 = std::chrono::steady_clock::now();
            const bool first_of_window = g == W.c0;
            if (first_of_window && C.window >= (uint32_t)wslots) {      // the device slot is free once window w - wslots is decoded
                while (decoded.load(std::memory_order_acquire) < (int64_t)C.window - wslots + 1 && !stop) std::this_thread::sleep_for(std::chrono::microseconds(50));
                if (stop) break;
                if ((e = cudaStreamWaitEvent(B->copy_stream, B->ev_decoded[ws], 0)) != cudaSuccess) return bad("cudaStreamWaitEvent", e);
                wait_slot_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - tw1).count();
            }
            if (first_of_window) {                            // the window's member table
                bgz::Member* tab = (bgz::Member*)B->h_tab[ws];
                const uint64_t nmem = W.m1 - W.m0;
                for (uint64_t i = 0; i < nmem; ++i) {
                    const bdk_bgzf_member& m = M[W.m0 + i];
                    tab[i].in_off = m.in_off - W.in_begin; tab[i].out_off = m.out_off - W.out_begin; tab[i].in_len = m.in_len; tab[i].out_len = m.out_len;
                }
                if ((e = cudaMemcpyAsync(B->d_slot[ws].p, tab, nmem * sizeof(bgz::Member), cudaMemcpyHostToDevice, B->copy_stream)) != cudaSuccess) return bad("cudaMemcpyAsync (member table)", e);
            }
            // file bytes -> pinned buffer (a few threads; page-cache reads through the caller's mapping) -> device, piece by piece
            uint8_t* dcomp = (uint8_t*)B->d_slot[ws].p + tab_bytes;
            for (uint64_t p0 = 0; p0 < C.in_bytes; p0 += BamDev::PIECE, ++piece) {
                const uint64_t pn = std::min<uint64_t>(BamDev::PIECE, C.in_bytes - p0);
                const int ps = (int)(piece % BamDev::PSLOTS);
                const auto tp0 = std::chrono::steady_clock::now();
                if (piece >= (uint64_t)BamDev::PSLOTS)      // the buffer is free once the copy of piece - PSLOTS is through
                    if ((e = cudaEventSynchronize(B->ev_piece[ps])) != cudaSuccess) return bad("cudaEventSynchronize", e);
                if (B->h_cap[ps] < BamDev::PIECE + 64) {
                    if (B->h_slot[ps]) cudaFreeHost(B->h_slot[ps]);
                    B->h_slot[ps] = nullptr; B->h_cap[ps] = 0;
                    if ((e = cudaHostAlloc(&B->h_slot[ps], BamDev::PIECE + 64, cudaHostAllocDefault)) != cudaSuccess) return bad("cudaHostAlloc (staging buffer)", e);
                    B->h_cap[ps] = BamDev::PIECE + 64;
                }
                uint8_t* hs = (uint8_t*)B->h_slot[ps];
                const auto ts0 = std::chrono::steady_clock::now();
                wait_pinned_s += std::chrono::duration<double>(ts0 - tp0).count();
                const uint8_t* from = src->file + C.in_begin + p0;
                copiers.run(pn, [&](uint64_t lo, uint64_t hi) { memcpy(hs + lo, from + lo, hi - lo); });
                stage_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - ts0).count();
                if ((e = cudaMemcpyAsync(dcomp + (C.in_begin - W.in_begin) + p0, hs, pn, cudaMemcpyHostToDevice, B->copy_stream)) != cudaSuccess) return bad("cudaMemcpyAsync (compressed bytes)", e);
                cudaEventRecord(B->ev_piece[ps], B->copy_stream);
            }
            cudaEventRecord(B->ev_copied[ev], B->copy_stream);
            cudaStreamWaitEvent(ist, B->ev_copied[ev], 0);
            const uint64_t k0 = C.m0 - W.m0, nmem = C.m1 - C.m0;
            const bgz::Member* dm = (const bgz::Member*)B->d_slot[ws].p + k0;
            uint8_t* dout = (uint8_t*)B->d_raw[ws].p + BamDev::CARRY_CAP;
            int32_t* dst = B->d_status[ws].as<int32_t>() + k0;
            if (g == 0) cudaEventRecord(B->ev_first, ist);
            const unsigned grid = (unsigned)std::min<uint64_t>((nmem + bgzw::WARPS_PER_CTA - 1) / bgzw::WARPS_PER_CTA, (uint64_t)kNumSMs * bgzw::CTAS_PER_SM);
            bgzw::bgzf_inflate_warp_kernel<<<grid, bgzw::CTA_THREADS, sizeof(bgzw::Tables) * bgzw::WARPS_PER_CTA, ist>>>(
                dcomp, dm, (uint32_t)nmem, dout, dst, B->d_counter.as<uint32_t>() + g);
            bgzw::bgzf_crc_kernel<<<(unsigned)std::min<uint64_t>((nmem + 7) / 8, (uint64_t)kNumSMs * 8), 256, 0, ist>>>(dcomp, dm, (uint32_t)nmem, dout, dst);
            last_on_stream[g % nstreams] = (int)ev;
            if ((e = cudaEventRecord(B->ev_inflated[ev], ist)) != cudaSuccess) return bad("bgzf inflate launch", e);
            if (g + 1 == nchunk) {                            // the end of the inflate work: after the last kernel of every stream
                for (int k = 0; k < BamDev::NSTREAMS; ++k)
                    if (last_on_stream[k] >= 0 && B->inflate_stream[k] != ist) cudaStreamWaitEvent(ist, B->ev_inflated[last_on_stream[k]], 0);
                cudaEventRecord(B->ev_last, ist);
            }
            if ((e = cudaGetLastError()) != cudaSuccess) return bad("bgzf inflate launch", e);
            launched.store((int64_t)g + 1, std::memory_order_release);
            if (trace && (g == 0 || g + 1 == nchunk))
                fprintf(stderr, "[bamdev] %.1f ms: chunk %zu of %zu launched (so far: staging %.1f ms, waiting for a pinned buffer %.1f ms, for a window slot %.1f ms)\n",
                        since(), g + 1, nchunk, stage_s * 1e3, wait_pinned_s * 1e3, wait_slot_s * 1e3);
        }
    });

    // ---- consumer ----------------------------------------------------------------------------------------------------
    uint64_t carry = 0, records = 0, kept_total = 0, h2d = 0;
    uint32_t misses = 0;
    float ms_inflate = 0.f;
    bamdev::WinInfo* hinfo = B->h_info;
    uint32_t* hkept = (uint32_t*)(hinfo + 1);
    bamdev::WinInfo* dinfo = B->d_info.as<bamdev::WinInfo>();
    uint32_t* dnrec = (uint32_t*)(dinfo + 1);
    uint32_t* dkept = dnrec + 1;
    uint32_t* dunsorted = dnrec + 2;
    const bool inflate_only = getenv("BDK_BAMDEV_INFLATE_ONLY") != nullptr;
    auto consume = [&]() -> int {
        CU(cudaMemsetAsync(dunsorted, 0, 4, c->stream));
        for (size_t w = 0; w < nwin; ++w) {
            const int s = (int)(w % wslots);
            BamWindow const& W = wins[w];
            while (launched.load(std::memory_order_acquire) < (int64_t)W.c1) std::this_thread::sleep_for(std::chrono::microseconds(20));
            if (producer_rc) return fail(c, producer_rc.load(), "bdk_push_bam: %s", producer_err.c_str());
            for (uint32_t g = W.c0; g < W.c1; ++g) CU(cudaStreamWaitEvent(c->stream, B->ev_inflated[g % ring], 0));
            const bool last = w + 1 == nwin;
            const uint64_t nmem = W.m1 - W.m0;
            if (inflate_only) {                              // measurement aid (BDK_BAMDEV_INFLATE_ONLY): the producer's side alone
                CU(cudaEventRecord(B->ev_decoded[s], c->stream));
                decoded.store((int64_t)w + 1, std::memory_order_release);
                continue;
            }
            // the window's bytes: [carry][members' output], the first window from the first record, the last one up to end_off
            const uint64_t skip = w == 0 ? src->first_record - W.out_begin : 0;
            const uint8_t* raw = (const uint8_t*)B->d_raw[s].p + BamDev::CARRY_CAP + skip - carry;
            uint64_t n = carry + W.out_bytes - skip;
            if (last) n -= W.out_begin + W.out_bytes - end_off;
            const uint32_t nseg = (uint32_t)div_up<uint64_t>(std::max<uint64_t>(n, 1), bamdev::SEG_BYTES);
            brec::Segment* seg = B->d_seg.as<brec::Segment>();
            tstart(c, T_CHAIN);
            bamdev::chain_guess_kernel<<<div_up<uint32_t>(nseg, 128), 128, 0, c->stream>>>(raw, n, nseg, src->n_ref, seg);
            bamdev::chain_resolve_kernel<<<1, 1024, 0, c->stream>>>(raw, n, nseg, (uint32_t)nmem, last ? 1 : 0, B->d_status[s].as<int32_t>(), seg,
                                                                     B->d_base.as<uint32_t>(), dinfo, dnrec);
            tstop(c, T_CHAIN);
            c->launches += 2;
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(hinfo, dinfo, sizeof(bamdev::WinInfo), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            tcollect(c);
            h2d += nmem * sizeof(bgz::Member) + W.in_bytes;
            if (hinfo->err & bamdev::E_MEMBER)
                return fail(c, BDK_ERR_DATA, "BGZF member %llu did not inflate on the device (status %d, %u members of the window refused)",
                            (unsigned long long)(W.m0 + hinfo->first_bad_member), hinfo->first_bad_status, hinfo->bad_members);
            if (hinfo->err & (bamdev::E_RECORD | bamdev::E_TRUNCATED)) return fail(c, BDK_ERR_DATA, "truncated BAM record");
            const uint32_t nrec = hinfo->nrec;
            misses += hinfo->guess_misses;
            records += nrec;
            uint32_t kept = 0;
            if (nrec) {
                ENS(B->d_recoff, (size_t)nrec * 4);
                bamdev::Columns cols;
                if (collect) {              // behind what the bam's columns hold; sized once from the first window's record density
                    uint64_t want = kept_total + nrec;
                    if (want * 4 > collect->cols[0].cap && n) want = std::max<uint64_t>(want, (uint64_t)((double)nrec / (double)n * (double)(end_off - src->first_record) * 1.03) + 65536);
                    for (int k = 0; k < 10; ++k) { int erc = ensure(c, collect->cols[k], (size_t)want * width[k], true); if (erc) return erc; }
                    const bamdev::Columns v = collect->view();
                    cols = bamdev::Columns{v.pos + kept_total, v.mpos + kept_total, v.tid + kept_total, v.mtid + kept_total, v.isize + kept_total, v.qlen + kept_total,
                                           v.flag + kept_total, v.rgid + kept_total, v.mapq + kept_total, v.qid + kept_total};
                } else {
                    for (int k = 0; k < 10; ++k) ENS(B->d_cols[k], (size_t)nrec * width[k]);
                    cols = bamdev::Columns{B->d_cols[0].as<int32_t>(), B->d_cols[1].as<int32_t>(), B->d_cols[2].as<int32_t>(), B->d_cols[3].as<int32_t>(), B->d_cols[4].as<int32_t>(),
                                           B->d_cols[8].as<int32_t>(), B->d_cols[5].as<uint16_t>(), B->d_cols[7].as<uint16_t>(), B->d_cols[6].as<uint8_t>(), B->d_cols[9].as<uint64_t>()};
                }
                tstart(c, T_EXTRACT);
                bamdev::chain_write_kernel<<<div_up<uint32_t>(nseg, 128), 128, 0, c->stream>>>(raw, seg, B->d_base.as<uint32_t>(), nseg, B->d_recoff.as<uint32_t>());
                const brec::RegionSel sel{src->region_on, src->region_tid, src->region_beg, src->region_end};
                device_scan(c->stream, bamdev::KeepFlag{raw, B->d_recoff.as<uint32_t>(), sel},
                            bamdev::ExtractOut{raw, B->d_recoff.as<uint32_t>(), B->d_rgtab.as<bamdev::RgEntry>(), rg_slots, src->rg_other, cols},
                            dnrec, dkept, 0, ScanScratch{c->d_scan_sums.as<uint32_t>()});
                bamdev::sorted_check_kernel<<<kNumSMs * 2, 256, 0, c->stream>>>(cols.tid, cols.pos, dkept, B->d_prev.as<int32_t>(), (uint32_t)w, dunsorted);
                tstop(c, T_EXTRACT);
                c->launches += 5;
                CU(cudaGetLastError());
                CU(cudaMemcpyAsync(hkept, dkept, 8, cudaMemcpyDeviceToHost, c->stream));      // kept, unsorted
                CU(cudaStreamSynchronize(c->stream));
                tcollect(c);
                kept = hkept[0];
                if (kept && host_out) {
                    if (kept_total + kept > cap) return fail(c, BDK_ERR_ARG, "bdk_decode_bam: more than %llu records", (unsigned long long)cap);
                    void* dst[10] = {(void*)host_out->pos, (void*)host_out->mpos, (void*)host_out->tid, (void*)host_out->mtid, (void*)host_out->isize,
                                     (void*)host_out->flag, (void*)host_out->mapq, (void*)host_out->rgid, (void*)host_out->qlen, (void*)host_out->qid};
                    for (int k = 0; k < 10; ++k)
                        CU(cudaMemcpyAsync((char*)dst[k] + kept_total * width[k], B->d_cols[k].p, (size_t)kept * width[k], cudaMemcpyDeviceToHost, c->stream));
                    CU(cudaStreamSynchronize(c->stream));
                } else if (kept && collect) {
                    collect->n = kept_total + kept;
                } else if (kept) {
                    bdk_soa d;
                    d.pos = cols.pos; d.mpos = cols.mpos; d.tid = cols.tid; d.mtid = cols.mtid; d.isize = cols.isize; d.flag = cols.flag; d.mapq = cols.mapq;
                    d.rgid = cols.rgid; d.qlen = cols.qlen; d.qid = cols.qid;
                    int prc = push_common(c, kept, kept, [&]() -> int { return launch_k1(c, d, kept, (uint32_t)c->n_records, true); });
                    if (prc) return prc;
                }
            } else {
                CU(cudaMemsetAsync(dkept, 0, 4, c->stream));
                bamdev::sorted_check_kernel<<<1, 32, 0, c->stream>>>(nullptr, nullptr, dkept, B->d_prev.as<int32_t>(), (uint32_t)w, dunsorted);
                c->launches += 1;
            }
            kept_total += kept;
            const uint64_t new_carry = n - hinfo->tail;
            if (!last) {
                if (new_carry > BamDev::CARRY_CAP) return fail(c, BDK_ERR_DATA, "a BAM record longer than %zu bytes: not supported by the device decode", BamDev::CARRY_CAP);
                const int s2 = (int)((w + 1) % wslots);
                if (new_carry)
                    CU(cudaMemcpyAsync((uint8_t*)B->d_raw[s2].p + BamDev::CARRY_CAP - new_carry, raw + hinfo->tail, new_carry, cudaMemcpyDeviceToDevice, c->stream));
            }
            carry = new_carry;
            if (trace && (w == 0 || w + 1 == nwin)) fprintf(stderr, "[bamdev] %.1f ms: window %zu of %zu decoded (%u records)\n", since(), w + 1, nwin, nrec);
            CU(cudaEventRecord(B->ev_decoded[s], c->stream));
            decoded.store((int64_t)w + 1, std::memory_order_release);
        }
        CU(cudaMemcpyAsync(hkept, dkept, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        return 0;
    };
    const int rc = consume();
    stop = true;
    producer.join();
    for (int k = 0; k < BamDev::NSTREAMS; ++k) cudaStreamSynchronize(B->inflate_stream[k]);
    cudaStreamSynchronize(B->copy_stream);
    if (rc) return rc;
    if (producer_rc) return fail(c, producer_rc.load(), "bdk_push_bam: %s", producer_err.c_str());
    cudaEventElapsedTime(&ms_inflate, B->ev_first, B->ev_last);      // first inflate launch to the end of the last one (chunks overlap)
    c->timers[T_INFLATE].ms += ms_inflate; c->timers[T_INFLATE].launches += (int)(2 * nchunk);
    c->launches += 2 * nchunk;
    c->h2d_bytes = h2d;
    if (stats) {
        stats->records = records; stats->kept = kept_total; stats->windows = (uint32_t)nwin; stats->sorted = hkept[1] ? 0 : 1;
        stats->guess_misses = misses; stats->h2d_bytes = h2d; stats->inflated_bytes = end_off - src->first_record;
        stats->stage_ms = (float)(stage_s * 1e3);
        stats->wall_ms = (float)(std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count() * 1e3);
        stats->inflate_ms = ms_inflate; stats->chain_ms = c->timers[T_CHAIN].ms; stats->extract_ms = c->timers[T_EXTRACT].ms;
    }
    return 0;
}

}  // namespace

extern "C" {

uint64_t bdk_hash_bytes(const void* p, uint64_t n) { return brec::hash_bytes((const uint8_t*)p, (size_t)n); }

int bdk_push_bam(bdk_ctx* c, const bdk_bam_source* src, bdk_bam_stats* stats) { return bam_pipeline(c, src, stats, nullptr, 0); }

// Two bams of one config, decoded on the device one after the other, merged in the reference's order (bam_merge.cuh) and
// classified; host_out != nullptr: the merged columns go to the host instead (tests). srcs[0] must be the config's first bam.
static int push_two_bams(bdk_ctx* c, const bdk_bam_source* srcs, bdk_bam_stats* stats, const bdk_soa* host_out, uint64_t cap) {
    static const size_t width[10] = {4, 4, 4, 4, 4, 2, 1, 2, 4, 8};
    BamCollect col[2], merged;
    DevBuf d_k[2], d_cutx, d_cuty, d_partx, d_party, d_order, d_small;
    auto cleanup = [&]() {
        col[0].release(); col[1].release(); merged.release();
        for (DevBuf* b : {&d_k[0], &d_k[1], &d_cutx, &d_cuty, &d_partx, &d_party, &d_order, &d_small}) if (b->p) { cudaFree(b->p); b->p = nullptr; }
    };
    bdk_bam_stats local_stats[2];
    if (!stats) stats = local_stats;
    auto run = [&]() -> int {
        for (int b = 0; b < 2; ++b) {
            const int rc = bam_pipeline(c, &srcs[b], &stats[b], nullptr, 0, &col[b]);
            if (rc) return rc;
            if (!stats[b].sorted) return fail(c, BDK_ERR_DATA, "bam %d is not sorted by reference sequence and position: the two-bam device merge needs sorted input", b);
        }
        const uint64_t n0 = col[0].n, n1 = col[1].n, n = n0 + n1;
        if (n > 0x7ffffff0ull) return fail(c, BDK_ERR_ARG, "more than 2^31 records in a two-bam device merge");
        if (n == 0) return 0;
        cudaStream_t st = c->stream;
        tstart(c, T_EXTRACT);
        for (int k = 0; k < 10; ++k) ENS(merged.cols[k], (size_t)n * width[k]);
        for (int b = 0; b < 2; ++b) {
            ENS(d_k[b], (col[b].n + 1) * 8);
            if (col[b].n) {
                const bamdev::Columns v = col[b].view();
                bammerge::keys_kernel<<<kNumSMs * 8, 256, 0, st>>>(v.tid, v.pos, v.flag, (uint32_t)col[b].n, d_k[b].as<unsigned long long>());
            }
        }
        // cuts from the larger bam
        const int X = n1 > n0 ? 1 : 0, Y = 1 - X;
        const uint32_t nx = (uint32_t)col[X].n, ny = (uint32_t)col[Y].n;
        const uint32_t nblocks = div_up<uint32_t>(std::max<uint32_t>(nx, 1), bammerge::CUT_EVERY);
        ENS(d_cutx, (size_t)nblocks * 4); ENS(d_cuty, (size_t)nblocks * 4); ENS(d_partx, ((size_t)nblocks + 2) * 4); ENS(d_party, ((size_t)nblocks + 2) * 4);
        ENS(d_order, (size_t)n * 4); ENS(d_small, 64);
        uint32_t* small = d_small.as<uint32_t>();          // [0] blocks, [1] cuts found, [2] longest part
        const uint32_t init[3] = {nblocks, 0, 0};
        CU(cudaMemcpyAsync(small, init, 12, cudaMemcpyHostToDevice, st));
        bammerge::find_cuts_kernel<<<div_up<uint32_t>(nblocks, 128), 128, 0, st>>>(d_k[X].as<unsigned long long>(), nx, d_k[Y].as<unsigned long long>(), ny, nblocks,
                                                                                   d_cutx.as<uint32_t>(), d_cuty.as<uint32_t>());
        device_scan(st, bammerge::CutFlag{d_cutx.as<uint32_t>()}, bammerge::CutOut{d_cutx.as<uint32_t>(), d_cuty.as<uint32_t>(), d_partx.as<uint32_t>(), d_party.as<uint32_t>()},
                    small, small + 1, 0, ScanScratch{c->d_scan_sums.as<uint32_t>()});
        const uint32_t* part_a = X == 0 ? d_partx.as<uint32_t>() : d_party.as<uint32_t>();
        const uint32_t* part_b = X == 0 ? d_party.as<uint32_t>() : d_partx.as<uint32_t>();
        bammerge::merge_parts_kernel<<<div_up<uint32_t>(nblocks + 1, 128), 128, 0, st>>>(d_k[0].as<unsigned long long>(), (uint32_t)n0, d_k[1].as<unsigned long long>(), (uint32_t)n1,
                                                                                        part_a, part_b, small + 1, d_order.as<uint32_t>(), small + 2);
        bammerge::gather_kernel<<<kNumSMs * 16, 256, 0, st>>>(d_order.as<uint32_t>(), (uint32_t)n, col[0].view(), col[1].view(), merged.view());
        tstop(c, T_EXTRACT);
        c->launches += 8;
        CU(cudaGetLastError());
        uint32_t hs[3];
        CU(cudaMemcpyAsync(hs, small, 12, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        tcollect(c);
        stats[0].merge_parts = hs[1] + 1; stats[0].merge_longest_part = hs[2];
        col[0].release(); col[1].release();
        const bamdev::Columns m = merged.view();
        if (host_out) {
            if (n > cap) return fail(c, BDK_ERR_ARG, "bdk_decode_bams: more than %llu records", (unsigned long long)cap);
            void* dst[10] = {(void*)host_out->pos, (void*)host_out->mpos, (void*)host_out->tid, (void*)host_out->mtid, (void*)host_out->isize,
                             (void*)host_out->flag, (void*)host_out->mapq, (void*)host_out->rgid, (void*)host_out->qlen, (void*)host_out->qid};
            for (int k = 0; k < 10; ++k) CU(cudaMemcpyAsync(dst[k], merged.cols[k].p, (size_t)n * width[k], cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            return 0;
        }
        bdk_soa d;
        d.pos = m.pos; d.mpos = m.mpos; d.tid = m.tid; d.mtid = m.mtid; d.isize = m.isize; d.flag = m.flag; d.mapq = m.mapq; d.rgid = m.rgid; d.qlen = m.qlen; d.qid = m.qid;
        return push_common(c, n, n, [&]() -> int { return launch_k1(c, d, n, (uint32_t)c->n_records, true); });
    };
    const int rc = run();
    cudaStreamSynchronize(c->stream);
    cleanup();
    return rc;
}

// Three or more bams (up to bammerge::MAX_BAMS): each decoded on the device into columns that stay there, the packed
// (tid, pos, strand) keys copied to the host, the reference's merge order computed there by the priority queue itself
// (csrc/host/nway_merge.hpp: for three or more streams its tie order depends on the heap's history, so it cannot be cut into
// independent parts the way the two-stream merge is), the order copied back and the columns gathered through it. 12 bytes per record
// cross PCIe on top of the compressed files. Unsorted bams are fine here: the queue defines the order whatever the keys are.
int push_many_bams(bdk_ctx* c, const bdk_bam_source* srcs, int nb, bdk_bam_stats* stats, const bdk_soa* host_out, uint64_t cap) {
    static const size_t width[10] = {4, 4, 4, 4, 4, 2, 1, 2, 4, 8};
    std::vector<BamCollect> col((size_t)nb);
    BamCollect merged;
    std::vector<DevBuf> d_k((size_t)nb);
    DevBuf d_order;
    auto cleanup = [&]() {
        for (auto& x : col) x.release();
        merged.release();
        for (auto& b : d_k) if (b.p) { cudaFree(b.p); b.p = nullptr; }
        if (d_order.p) { cudaFree(d_order.p); d_order.p = nullptr; }
    };
    std::vector<bdk_bam_stats> local_stats((size_t)nb);
    if (!stats) stats = local_stats.data();
    auto run = [&]() -> int {
        uint64_t n = 0;
        for (int b = 0; b < nb; ++b) {
            const int rc = bam_pipeline(c, &srcs[b], &stats[b], nullptr, 0, &col[b]);
            if (rc) return rc;
            if (col[b].n >= (1ull << bammerge::BAM_SHIFT)) return fail(c, BDK_ERR_ARG, "bam %d has 2^%d records or more: too many for the device merge of several bams", b, bammerge::BAM_SHIFT);
            n += col[b].n;
        }
        if (n > 0x7ffffff0ull) return fail(c, BDK_ERR_ARG, "more than 2^31 records in a device merge");
        if (n == 0) return 0;
        cudaStream_t st = c->stream;
        tstart(c, T_EXTRACT);
        std::vector<std::vector<uint64_t>> hk((size_t)nb);
        for (int b = 0; b < nb; ++b) {
            if (!col[b].n) continue;
            ENS(d_k[b], col[b].n * 8);
            const bamdev::Columns v = col[b].view();
            bammerge::keys_kernel<<<kNumSMs * 8, 256, 0, st>>>(v.tid, v.pos, v.flag, (uint32_t)col[b].n, d_k[b].as<unsigned long long>());
            hk[b].resize(col[b].n);
            CU(cudaMemcpyAsync(hk[b].data(), d_k[b].p, col[b].n * 8, cudaMemcpyDeviceToHost, st));
            c->launches += 1;
        }
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(st));
        std::vector<uint32_t> order(n);
        {
            std::vector<const uint64_t*> kp((size_t)nb);
            std::vector<uint64_t> counts((size_t)nb);
            for (int b = 0; b < nb; ++b) { kp[b] = hk[b].data(); counts[b] = hk[b].size(); }
            // up to five sorted bams and enough cores: the parts of the merge in parallel, one run per valid heap layout at a cut
            // (exact: nway_merge.hpp); else -- or when it declines -- the queue on one thread
            bool all_sorted = true;
            for (int b = 0; b < nb; ++b) all_sorted = all_sorted && stats[b].sorted;
            const int hw = (int)std::thread::hardware_concurrency();
            int mthreads = hw >= 8 && !getenv("BDK_MERGE_HEAP") ? std::min(hw, 32) : 1;
            if (const char* e = getenv("BDK_MERGE_THREADS")) mthreads = std::max(1, atoi(e));
            if (!(all_sorted && mthreads >= 2 && bdh::nway_merge_order_parallel(kp.data(), counts.data(), nb, bammerge::BAM_SHIFT, order.data(), mthreads)))
                bdh::nway_merge_order(kp.data(), counts.data(), nb, bammerge::BAM_SHIFT, order.data());
        }
        ENS(d_order, (size_t)n * 4);
        CU(cudaMemcpyAsync(d_order.p, order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        for (int k = 0; k < 10; ++k) ENS(merged.cols[k], (size_t)n * width[k]);
        bammerge::ColumnsN all;
        memset(&all, 0, sizeof all);
        for (int b = 0; b < nb; ++b) all.c[b] = col[b].view();
        bammerge::gather_n_kernel<<<kNumSMs * 16, 256, 0, st>>>(d_order.as<uint32_t>(), (uint32_t)n, all, merged.view());
        tstop(c, T_EXTRACT);
        c->launches += 1;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(st));           // `order` is a local
        tcollect(c);
        stats[0].merge_parts = 1; stats[0].merge_longest_part = (uint32_t)std::min<uint64_t>(n, 0xffffffffu);
        for (auto& x : col) x.release();
        const bamdev::Columns m = merged.view();
        if (host_out) {
            if (n > cap) return fail(c, BDK_ERR_ARG, "bdk_decode_bams: more than %llu records", (unsigned long long)cap);
            void* dst[10] = {(void*)host_out->pos, (void*)host_out->mpos, (void*)host_out->tid, (void*)host_out->mtid, (void*)host_out->isize,
                             (void*)host_out->flag, (void*)host_out->mapq, (void*)host_out->rgid, (void*)host_out->qlen, (void*)host_out->qid};
            for (int k = 0; k < 10; ++k) CU(cudaMemcpyAsync(dst[k], merged.cols[k].p, (size_t)n * width[k], cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            return 0;
        }
        bdk_soa d;
        d.pos = m.pos; d.mpos = m.mpos; d.tid = m.tid; d.mtid = m.mtid; d.isize = m.isize; d.flag = m.flag; d.mapq = m.mapq; d.rgid = m.rgid; d.qlen = m.qlen; d.qid = m.qid;
        return push_common(c, n, n, [&]() -> int { return launch_k1(c, d, n, (uint32_t)c->n_records, true); });
    };
    const int rc = run();
    cudaStreamSynchronize(c->stream);
    cleanup();
    return rc;
}

int bdk_push_bams(bdk_ctx* c, const bdk_bam_source* srcs, int n, bdk_bam_stats* stats) {
    if (!c || !srcs) return BDK_ERR_ARG;
    if (n == 1) return bam_pipeline(c, srcs, stats, nullptr, 0);
    if (n < 1 || n > bammerge::MAX_BAMS) return fail(c, BDK_ERR_ARG, "bdk_push_bams merges up to %d bams on the device (%d given): use the host reader for more", bammerge::MAX_BAMS, n);
    if (c->finished) return fail(c, BDK_ERR_STATE, "bdk_push_bams after bdk_finish (call bdk_reset first)");
    if (n == 2) return push_two_bams(c, srcs, stats, nullptr, 0);
    return push_many_bams(c, srcs, n, stats, nullptr, 0);
}

int bdk_decode_bams(bdk_ctx* c, const bdk_bam_source* srcs, int n, const bdk_soa* host_out, uint64_t cap, bdk_bam_stats* stats) {
    if (!c || !srcs || !host_out) return BDK_ERR_ARG;
    if (n == 1) return bam_pipeline(c, srcs, stats, host_out, cap);
    if (n < 1 || n > bammerge::MAX_BAMS) return fail(c, BDK_ERR_ARG, "bdk_decode_bams: 1 to %d bams", bammerge::MAX_BAMS);
    if (n == 2) return push_two_bams(c, srcs, stats, host_out, cap);
    return push_many_bams(c, srcs, n, stats, host_out, cap);
}

int bdk_decode_bam(bdk_ctx* c, const bdk_bam_source* src, const bdk_soa* host_out, uint64_t cap, bdk_bam_stats* stats) {
    if (!host_out) return BDK_ERR_ARG;
    const void* ptrs[10] = {host_out->pos, host_out->mpos, host_out->tid, host_out->mtid, host_out->isize, host_out->flag, host_out->mapq, host_out->rgid, host_out->qlen, host_out->qid};
    for (int i = 0; i < 10; ++i) if (!ptrs[i] && cap) return fail(c, BDK_ERR_ARG, "null column %d", i);
    return bam_pipeline(c, src, stats, host_out, cap);
}

}  // extern "C"
