// bdk_bam.inl -- bdk_push_bam: a BAM file decoded on the device, window of BGZF members by window, and classified (part of
// bdk_core.cu). Only the COMPRESSED file crosses PCIe; inflated bytes, record offsets and the record columns live and die in HBM.
//
//   producer thread:  file bytes -> pinned slot -> H2D (copy stream) -> bgzf_inflate_warp_kernel + bgzf_crc_kernel (inflate
//                     stream), up to SLOTS windows ahead of
//   caller's thread:  chain_guess / chain_resolve (record boundaries) -> chain_write -> scan(keep) + extract (columns) ->
//                     sorted check -> K1 (the same launch bdk_push_device makes), on the context's stream.
//
// The inflate of window w + 1 therefore overlaps the decode and classification of window w, and the H2D copy of window w + 2.
// A window's inflated bytes are written at a fixed offset of its buffer; the bytes of a record cut off by the end of window w
// are copied in front of window w + 1's (`carry`), so records are parsed where they lie.
#include <atomic>
#include <thread>

namespace {

struct BamDev {
    static constexpr int SLOTS = 8;                         // windows resident on the device (compressed + inflated)
    static constexpr int PSLOTS = 3;                        // pinned staging buffers (one H2D copy each in flight)
    static constexpr int NSTREAMS = 4;                      // windows inflated concurrently: a window has fewer members than the GPU has warps
    static constexpr size_t CARRY_CAP = 16u << 20;          // longest partial record carried between windows
    cudaStream_t inflate_stream[NSTREAMS] = {nullptr, nullptr, nullptr, nullptr}, copy_stream = nullptr;
    void* h_slot[PSLOTS] = {nullptr, nullptr, nullptr};
    size_t h_cap[PSLOTS] = {0, 0, 0};
    DevBuf d_slot[SLOTS], d_raw[SLOTS], d_status[SLOTS];
    cudaEvent_t ev_copied[SLOTS], ev_inflated[SLOTS], ev_decoded[SLOTS], ev_first = nullptr, ev_last = nullptr;
    DevBuf d_seg, d_base, d_recoff, d_cols[10], d_info, d_rgtab, d_prev, d_counter;
    bamdev::WinInfo* h_info = nullptr;                      // pinned: WinInfo + kept count
    bool events = false;
};

void bamdev_free(BamDev* B) {
    if (!B) return;
    for (int s = 0; s < BamDev::PSLOTS; ++s) if (B->h_slot[s]) cudaFreeHost(B->h_slot[s]);
    for (int s = 0; s < BamDev::SLOTS; ++s) {
        for (DevBuf* b : {&B->d_slot[s], &B->d_raw[s], &B->d_status[s]}) if (b->p) cudaFree(b->p);
        if (B->events) { cudaEventDestroy(B->ev_copied[s]); cudaEventDestroy(B->ev_inflated[s]); cudaEventDestroy(B->ev_decoded[s]); }
    }
    if (B->events) { cudaEventDestroy(B->ev_first); cudaEventDestroy(B->ev_last); }
    for (DevBuf* b : {&B->d_seg, &B->d_base, &B->d_recoff, &B->d_info, &B->d_rgtab, &B->d_prev, &B->d_counter}) if (b->p) cudaFree(b->p);
    for (int k = 0; k < 10; ++k) if (B->d_cols[k].p) cudaFree(B->d_cols[k].p);
    if (B->h_info) cudaFreeHost(B->h_info);
    for (int k = 0; k < BamDev::NSTREAMS; ++k) if (B->inflate_stream[k]) cudaStreamDestroy(B->inflate_stream[k]);
    if (B->copy_stream) cudaStreamDestroy(B->copy_stream);
    delete B;
}

struct BamWindow { uint64_t m0, m1, in_begin, in_bytes, out_begin, out_bytes; };

void bamdev_release(void* p) { bamdev_free((BamDev*)p); }

// host_out == nullptr: classify the records (bdk_push_bam); else copy the decoded columns into the caller's host arrays of
// `cap` records and classify nothing (bdk_decode_bam).
int bam_pipeline(bdk_ctx* c, const bdk_bam_source* src, bdk_bam_stats* stats, const bdk_soa* host_out, uint64_t cap) {
    if (!c || !src) return BDK_ERR_ARG;
    if (stats) memset(stats, 0, sizeof *stats);
    if (c->finished && !host_out) return fail(c, BDK_ERR_STATE, "bdk_push_bam after bdk_finish (call bdk_reset first)");
    if (!src->file || (!src->members && src->n_members)) return fail(c, BDK_ERR_ARG, "null file image or member list");
    if (src->n_rg && (!src->rg_hash || !src->rg_id)) return fail(c, BDK_ERR_ARG, "null read-group table");
    if (src->n_ref < 0) return fail(c, BDK_ERR_ARG, "n_ref < 0");
    CU(cudaSetDevice(c->device));
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

    // ---- windows -------------------------------------------------------------------------------------------------
    uint64_t WIN_IN = 16ull << 20, WIN_OUT = 128ull << 20;
    if (src->window_bytes) WIN_IN = src->window_bytes;
    if (const char* e = getenv("BDK_BAMDEV_WINDOW_KB")) if (atoll(e) > 0) WIN_IN = (uint64_t)atoll(e) << 10;      // tests: many small windows
    std::vector<BamWindow> wins;
    for (uint64_t i = 0; i < nm;) {
        if (M[i].out_off >= end_off) break;                  // members behind the end of the records
        BamWindow w{i, i, M[i].in_off, 0, M[i].out_off, 0};
        while (w.m1 < nm && M[w.m1].out_off < end_off &&
               (w.m1 == w.m0 || (M[w.m1].in_off + M[w.m1].in_len + 8 - w.in_begin <= WIN_IN && w.out_bytes + M[w.m1].out_len <= WIN_OUT))) {
            w.in_bytes = M[w.m1].in_off + M[w.m1].in_len + 8 - w.in_begin;         // the CRC32 + ISIZE footer travels along
            w.out_bytes += M[w.m1].out_len;
            ++w.m1;
        }
        wins.push_back(w);
        i = w.m1;
    }
    const size_t nwin = wins.size();
    if (nwin == 0 || src->first_record < wins[0].out_begin || src->first_record - wins[0].out_begin >= wins[0].out_bytes)
        return fail(c, BDK_ERR_ARG, "first_record must lie in the first window of members (pass the members from the one that holds it)");
    uint64_t max_in = 0, max_out = 0, max_members = 0;
    for (auto const& w : wins) { max_in = std::max(max_in, w.in_bytes); max_out = std::max(max_out, w.out_bytes); max_members = std::max(max_members, w.m1 - w.m0); }
    if (max_in > (1ull << 31) || max_out + BamDev::CARRY_CAP > 0xfff00000ull) return fail(c, BDK_ERR_ARG, "decode window too large");

    // ---- buffers -------------------------------------------------------------------------------------------------
    if (!c->bamdev) c->bamdev = new BamDev;
    BamDev* B = (BamDev*)c->bamdev;
    if (!B->events) {
        for (int k = 0; k < BamDev::NSTREAMS; ++k) CU(cudaStreamCreateWithFlags(&B->inflate_stream[k], cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&B->copy_stream, cudaStreamNonBlocking));
        for (int s = 0; s < BamDev::SLOTS; ++s) {
            CU(cudaEventCreateWithFlags(&B->ev_copied[s], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&B->ev_inflated[s], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&B->ev_decoded[s], cudaEventDisableTiming));
        }
        CU(cudaEventCreate(&B->ev_first)); CU(cudaEventCreate(&B->ev_last));
        B->events = true;
        CU(cudaHostAlloc((void**)&B->h_info, 256, cudaHostAllocDefault));
        CU(cudaFuncSetAttribute(bgzw::bgzf_inflate_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(bgzw::Tables) * bgzw::WARPS_PER_CTA)));
    }
    const size_t tab_bytes = (max_members * sizeof(bgz::Member) + 255) & ~size_t(255);
    const size_t slot_bytes = tab_bytes + max_in + 64;
    const int nslots = (int)std::min<size_t>(BamDev::SLOTS, nwin);
    for (int s = 0; s < nslots; ++s) {
        ENS(B->d_slot[s], slot_bytes); ENS(B->d_raw[s], BamDev::CARRY_CAP + max_out + 64); ENS(B->d_status[s], max_members * 4);
    }
    ENS(B->d_seg, max_members * sizeof(brec::Segment)); ENS(B->d_base, max_members * 4);
    ENS(B->d_info, 256); ENS(B->d_prev, 16); ENS(B->d_counter, 4 * (nwin + 1));
    {   // read-group table: open addressing, at most half full
        uint32_t slots = 16;
        while (slots < 2 * src->n_rg + 2) slots <<= 1;
        std::vector<bamdev::RgEntry> tab(slots, bamdev::RgEntry{0, 0, 0});
        for (uint32_t i = 0; i < src->n_rg; ++i) {
            uint32_t p = (uint32_t)(src->rg_hash[i] % slots);
            while (tab[p].used) {
                if (tab[p].hash == src->rg_hash[i]) return fail(c, BDK_ERR_ARG, "two read groups with one key");
                p = p + 1 == slots ? 0 : p + 1;
            }
            tab[p] = bamdev::RgEntry{src->rg_hash[i], src->rg_id[i], 1};
            if ((int)src->rg_id[i] >= c->P.nrg) return fail(c, BDK_ERR_ARG, "rg_id[%u] out of range", i);
        }
        if ((int)src->rg_other >= c->P.nrg) return fail(c, BDK_ERR_ARG, "rg_other out of range");
        ENS(B->d_rgtab, slots * sizeof(bamdev::RgEntry));
        CU(cudaMemcpyAsync(B->d_rgtab.p, tab.data(), slots * sizeof(bamdev::RgEntry), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemsetAsync(B->d_counter.p, 0, 4 * (nwin + 1), c->stream));
        const int32_t prev0[4] = {INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN};
        CU(cudaMemcpyAsync(B->d_prev.p, prev0, 16, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));                // tab / prev0 are locals; the producer's kernels need the zeroed counters
        const uint32_t rg_slots = slots;
        // ---- producer ----------------------------------------------------------------------------------------------
        std::atomic<int64_t> inflated(0), decoded(0);
        std::atomic<int> producer_rc(0);
        std::atomic<bool> stop(false);
        std::string producer_err;
        const int copy_threads = (int)std::max(1u, std::min(4u, std::thread::hardware_concurrency()));
        std::thread producer([&]() {
            auto bad = [&](const char* what, cudaError_t e) { producer_err = std::string(what) + ": " + cudaGetErrorString(e); producer_rc = BDK_ERR_CUDA; inflated = (int64_t)nwin + 1; };
            cudaError_t e = cudaSetDevice(c->device);
            if (e != cudaSuccess) return bad("cudaSetDevice", e);
            for (size_t w = 0; w < nwin && !stop; ++w) {
                const int s = (int)(w % BamDev::SLOTS), ps = (int)(w % BamDev::PSLOTS);
                cudaStream_t ist = B->inflate_stream[w % BamDev::NSTREAMS];
                BamWindow const& W = wins[w];
                if (w >= (size_t)BamDev::PSLOTS)          // the pinned buffer is free once the copy of window w - PSLOTS is through
                    if ((e = cudaEventSynchronize(B->ev_copied[(w - BamDev::PSLOTS) % BamDev::SLOTS])) != cudaSuccess) return bad("cudaEventSynchronize", e);
                if (w >= (size_t)BamDev::SLOTS) {         // the device slot is free once window w - SLOTS is decoded
                    while (decoded.load(std::memory_order_acquire) < (int64_t)(w - BamDev::SLOTS + 1) && !stop) std::this_thread::sleep_for(std::chrono::microseconds(50));
                    if (stop) break;
                    if ((e = cudaStreamWaitEvent(B->copy_stream, B->ev_decoded[s], 0)) != cudaSuccess) return bad("cudaStreamWaitEvent", e);
                }
                if (B->h_cap[ps] < slot_bytes) {
                    if (B->h_slot[ps]) cudaFreeHost(B->h_slot[ps]);
                    B->h_slot[ps] = nullptr; B->h_cap[ps] = 0;
                    if ((e = cudaHostAlloc(&B->h_slot[ps], slot_bytes, cudaHostAllocDefault)) != cudaSuccess) return bad("cudaHostAlloc (staging slot)", e);
                    B->h_cap[ps] = slot_bytes;
                }
                uint8_t* hs = (uint8_t*)B->h_slot[ps];
                bgz::Member* tab = (bgz::Member*)hs;
                const uint64_t nmem = W.m1 - W.m0;
                for (uint64_t i = 0; i < nmem; ++i) {
                    const bdk_bgzf_member& m = M[W.m0 + i];
                    tab[i].in_off = m.in_off - W.in_begin; tab[i].out_off = m.out_off - W.out_begin; tab[i].in_len = m.in_len; tab[i].out_len = m.out_len;
                }
                {   // file bytes -> pinned slot, a few threads (page-cache reads through the caller's mapping)
                    std::vector<std::thread> th;
                    const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)copy_threads, W.in_bytes >> 20));
                    auto part = [&](int t) {
                        const uint64_t lo = W.in_bytes * t / T, hi = W.in_bytes * (t + 1) / T;
                        memcpy(hs + tab_bytes + lo, src->file + W.in_begin + lo, hi - lo);
                    };
                    for (int t = 1; t < T; ++t) th.emplace_back(part, t);
                    part(0);
                    for (auto& x : th) x.join();
                }
                if ((e = cudaMemcpyAsync(B->d_slot[s].p, hs, tab_bytes + W.in_bytes, cudaMemcpyHostToDevice, B->copy_stream)) != cudaSuccess) return bad("cudaMemcpyAsync (compressed window)", e);
                cudaEventRecord(B->ev_copied[s], B->copy_stream);
                cudaStreamWaitEvent(ist, B->ev_copied[s], 0);
                const bgz::Member* dm = (const bgz::Member*)B->d_slot[s].p;
                const uint8_t* dcomp = (const uint8_t*)B->d_slot[s].p + tab_bytes;
                uint8_t* dout = (uint8_t*)B->d_raw[s].p + BamDev::CARRY_CAP;
                if (w == 0) cudaEventRecord(B->ev_first, ist);
                const unsigned grid = (unsigned)std::min<uint64_t>((nmem + bgzw::WARPS_PER_CTA - 1) / bgzw::WARPS_PER_CTA, (uint64_t)kNumSMs * 4);
                bgzw::bgzf_inflate_warp_kernel<<<grid, bgzw::CTA_THREADS, sizeof(bgzw::Tables) * bgzw::WARPS_PER_CTA, ist>>>(
                    dcomp, dm, (uint32_t)nmem, dout, B->d_status[s].as<int32_t>(), B->d_counter.as<uint32_t>() + w);
                bgzw::bgzf_crc_kernel<<<(unsigned)std::min<uint64_t>((nmem + 7) / 8, (uint64_t)kNumSMs * 8), 256, 0, ist>>>(
                    dcomp, dm, (uint32_t)nmem, dout, B->d_status[s].as<int32_t>());
                if (w + 1 == nwin) {                      // the last window's kernels end after everything the other streams still run
                    for (int k = 0; k < BamDev::NSTREAMS; ++k)
                        if (B->inflate_stream[k] != ist && w >= (size_t)((w % BamDev::NSTREAMS + BamDev::NSTREAMS - k) % BamDev::NSTREAMS)) {
                            const size_t wk = w - (size_t)((w % BamDev::NSTREAMS + BamDev::NSTREAMS - k) % BamDev::NSTREAMS);
                            cudaStreamWaitEvent(ist, B->ev_inflated[wk % BamDev::SLOTS], 0);
                        }
                    cudaEventRecord(B->ev_last, ist);
                }
                if ((e = cudaEventRecord(B->ev_inflated[s], ist)) != cudaSuccess) return bad("bgzf inflate launch", e);
                if ((e = cudaGetLastError()) != cudaSuccess) return bad("bgzf inflate launch", e);
                inflated.store((int64_t)w + 1, std::memory_order_release);
            }
        });
        // ---- consumer ----------------------------------------------------------------------------------------------
        int rc = 0;
        uint64_t carry = 0, records = 0, kept_total = 0, h2d = 0;
        uint32_t misses = 0;
        float ms_inflate = 0.f;
        bamdev::WinInfo* hinfo = B->h_info;
        uint32_t* hkept = (uint32_t*)(hinfo + 1);
        bamdev::WinInfo* dinfo = B->d_info.as<bamdev::WinInfo>();
        uint32_t* dnrec = (uint32_t*)(dinfo + 1);
        uint32_t* dkept = dnrec + 1;
        uint32_t* dunsorted = dnrec + 2;
        auto consume = [&]() -> int {
            CU(cudaMemsetAsync(dunsorted, 0, 4, c->stream));
            for (size_t w = 0; w < nwin; ++w) {
                const int s = (int)(w % BamDev::SLOTS);
                BamWindow const& W = wins[w];
                while (inflated.load(std::memory_order_acquire) < (int64_t)w + 1) std::this_thread::sleep_for(std::chrono::microseconds(20));
                if (producer_rc) return fail(c, producer_rc.load(), "bdk_push_bam: %s", producer_err.c_str());
                CU(cudaStreamWaitEvent(c->stream, B->ev_inflated[s], 0));
                const bool last = w + 1 == nwin;
                const uint64_t nmem = W.m1 - W.m0;
                // the window's bytes: [carry][members' output], the first window from the first record, the last one up to end_off
                const int64_t skip = w == 0 ? (int64_t)(src->first_record - W.out_begin) : 0;
                const uint8_t* raw = (const uint8_t*)B->d_raw[s].p + BamDev::CARRY_CAP + skip - carry;
                const int64_t carry_arg = w == 0 ? -skip : (int64_t)carry;
                uint64_t n = carry + W.out_bytes - (uint64_t)skip;
                if (last) n -= W.out_begin + W.out_bytes - end_off;
                const bgz::Member* dm = (const bgz::Member*)B->d_slot[s].p;
                brec::Segment* seg = B->d_seg.as<brec::Segment>();
                tstart(c, T_CHAIN);
                bamdev::chain_guess_kernel<<<(unsigned)div_up<uint64_t>(nmem, 64), 64, 0, c->stream>>>(raw, n, dm, (uint32_t)nmem, carry_arg, src->n_ref, seg);
                bamdev::chain_resolve_kernel<<<1, 1024, 0, c->stream>>>(raw, n, dm, (uint32_t)nmem, carry_arg, last ? 1 : 0, B->d_status[s].as<int32_t>(), seg,
                                                                         B->d_base.as<uint32_t>(), dinfo, dnrec);
                tstop(c, T_CHAIN);
                c->launches += 2;
                CU(cudaGetLastError());
                CU(cudaMemcpyAsync(hinfo, dinfo, sizeof(bamdev::WinInfo), cudaMemcpyDeviceToHost, c->stream));
                CU(cudaStreamSynchronize(c->stream));
                tcollect(c);
                h2d += tab_bytes + W.in_bytes;
                if (hinfo->err & bamdev::E_MEMBER)
                    return fail(c, BDK_ERR_DATA, "BGZF member %llu did not inflate on the device (status %d, %u members of the window refused)",
                                (unsigned long long)(W.m0 + hinfo->first_bad_member), hinfo->first_bad_status, hinfo->bad_members);
                if (hinfo->err & bamdev::E_RECORD) return fail(c, BDK_ERR_DATA, "truncated BAM record");
                if (hinfo->err & bamdev::E_TRUNCATED) return fail(c, BDK_ERR_DATA, "truncated BAM record");
                const uint32_t nrec = hinfo->nrec;
                misses += hinfo->guess_misses;
                records += nrec;
                uint32_t kept = 0;
                if (nrec) {
                    static const size_t width[10] = {4, 4, 4, 4, 4, 2, 1, 2, 4, 8};
                    ENS(B->d_recoff, (size_t)nrec * 4);
                    for (int k = 0; k < 10; ++k) ENS(B->d_cols[k], (size_t)nrec * width[k]);
                    tstart(c, T_EXTRACT);
                    bamdev::chain_write_kernel<<<(unsigned)div_up<uint64_t>(nmem, 64), 64, 0, c->stream>>>(raw, seg, B->d_base.as<uint32_t>(), (uint32_t)nmem, B->d_recoff.as<uint32_t>());
                    bamdev::Columns cols{B->d_cols[0].as<int32_t>(), B->d_cols[1].as<int32_t>(), B->d_cols[2].as<int32_t>(), B->d_cols[3].as<int32_t>(), B->d_cols[4].as<int32_t>(),
                                         B->d_cols[8].as<int32_t>(), B->d_cols[5].as<uint16_t>(), B->d_cols[7].as<uint16_t>(), B->d_cols[6].as<uint8_t>(), B->d_cols[9].as<uint64_t>()};
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
                    const int s2 = (int)((w + 1) % BamDev::SLOTS);
                    if (new_carry)
                        CU(cudaMemcpyAsync((uint8_t*)B->d_raw[s2].p + BamDev::CARRY_CAP - new_carry, raw + hinfo->tail, new_carry, cudaMemcpyDeviceToDevice, c->stream));
                }
                carry = new_carry;
                CU(cudaEventRecord(B->ev_decoded[s], c->stream));
                decoded.store((int64_t)w + 1, std::memory_order_release);
            }
            CU(cudaMemcpyAsync(hkept, dkept, 8, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            return 0;
        };
        rc = consume();
        stop = true;
        producer.join();
        for (int k = 0; k < BamDev::NSTREAMS; ++k) cudaStreamSynchronize(B->inflate_stream[k]);
        cudaStreamSynchronize(B->copy_stream);
        if (rc) return rc;
        if (producer_rc) return fail(c, producer_rc.load(), "bdk_push_bam: %s", producer_err.c_str());
        cudaEventElapsedTime(&ms_inflate, B->ev_first, B->ev_last);      // first inflate launch to the end of the last one (windows overlap)
        c->timers[T_INFLATE].ms += ms_inflate; c->timers[T_INFLATE].launches += (int)(2 * nwin);
        c->launches += 2 * nwin;
        c->h2d_bytes = h2d;
        if (stats) {
            stats->records = records; stats->kept = kept_total; stats->windows = (uint32_t)nwin; stats->sorted = hkept[1] ? 0 : 1;
            stats->guess_misses = misses; stats->h2d_bytes = h2d; stats->inflated_bytes = end_off - src->first_record;
            stats->inflate_ms = ms_inflate; stats->chain_ms = c->timers[T_CHAIN].ms; stats->extract_ms = c->timers[T_EXTRACT].ms;
        }
    }
    return 0;
}

}  // namespace

extern "C" {

uint64_t bdk_hash_bytes(const void* p, uint64_t n) { return brec::hash_bytes((const uint8_t*)p, (size_t)n); }

int bdk_push_bam(bdk_ctx* c, const bdk_bam_source* src, bdk_bam_stats* stats) { return bam_pipeline(c, src, stats, nullptr, 0); }

int bdk_decode_bam(bdk_ctx* c, const bdk_bam_source* src, const bdk_soa* host_out, uint64_t cap, bdk_bam_stats* stats) {
    if (!host_out) return BDK_ERR_ARG;
    const void* ptrs[10] = {host_out->pos, host_out->mpos, host_out->tid, host_out->mtid, host_out->isize, host_out->flag, host_out->mapq, host_out->rgid, host_out->qlen, host_out->qid};
    for (int i = 0; i < 10; ++i) if (!ptrs[i] && cap) return fail(c, BDK_ERR_ARG, "null column %d", i);
    return bam_pipeline(c, src, stats, host_out, cap);
}

}  // extern "C"
