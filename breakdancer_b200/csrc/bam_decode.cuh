// bam_decode.cuh -- record boundaries and field extraction on inflated BAM bytes that stay in HBM: the stages between the GPU
// inflate (bgzf_inflate_warp.cuh) and K1, so that a run from a BAM file moves the COMPRESSED file over PCIe and nothing else.
// Launched by bdk_push_bam (bdk_bam.inl), window of BGZF members by window. The per-record logic is csrc/bam_records.h, the
// code the host decoder runs and its tests pin. Replaces, for a run from files, samread + the Alignment constructor
// (src/lib/io/BamReader.hpp:64-70, src/lib/io/Alignment.cpp:12-64) and the reader filter (src/lib/io/BamIo.cpp:11-18).
//
// Record boundaries, as on the host (bam_io.cpp find_records): a window is cut into segments of 8 KiB; every segment guesses its first record (three plausible records in a row) and follows the
// chain of block_size prefixes to its end, all segments in parallel (chain_guess_kernel). chain_resolve_kernel then checks the
// guesses against the chain that really arrives: if every chain ends exactly on the next segment's guess the guessed chains ARE
// the serial chain (induction from the window's first byte, which is a true record start); wherever a guess is wrong or missing
// one thread follows the true chain through that segment. Exactness never rests on a guess. A record cut off by the end of the
// window is carried into the next one (`tail`).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bam_records.h"
#include "bgzf_inflate.cuh"
#include "scan_sort.cuh"

namespace bamdev {

using brec::Segment;
using brec::NO_GUESS;
using bgz::Member;

enum { E_MEMBER = 1, E_RECORD = 2, E_TRUNCATED = 4 };

struct WinInfo {                 // what the host needs to know about a window before it can go on
    uint64_t tail;               // offset (in the window's bytes) of the first byte that belongs to the next window
    uint32_t nrec;               // records that start (and end) in the window
    uint32_t err;                // E_*
    uint32_t bad_members, first_bad_member;
    int32_t first_bad_status;
    uint32_t guess_misses;       // segments the true chain had to be followed through
};

// segment k of a window = bytes [k * SEG_BYTES, (k + 1) * SEG_BYTES) of it (the last one shorter)
constexpr uint32_t SEG_BYTES = 8192;
__device__ __forceinline__ uint64_t seg_lo(uint32_t k) { return (uint64_t)k * SEG_BYTES; }
__device__ __forceinline__ uint64_t seg_hi(uint32_t k, uint64_t n) { const uint64_t h = (uint64_t)(k + 1) * SEG_BYTES; return h < n ? h : n; }

// One thread per segment.
__global__ void __launch_bounds__(128) chain_guess_kernel(const uint8_t* __restrict__ raw, uint64_t n, uint32_t nseg, int32_t nref, Segment* __restrict__ seg) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nseg) return;
    seg[k] = brec::segment_guess(raw, n, seg_lo(k), seg_hi(k, n), k == 0, nref);
}

// neighbours read a segment while its owner rewrites it (rounds of chain_resolve_kernel): two 16-byte halves, each consistent
__device__ __forceinline__ void seg_store(Segment* p, const Segment& s) {
    volatile uint64_t* q = reinterpret_cast<volatile uint64_t*>(p);
    q[0] = s.guess; q[1] = s.end; q[2] = (uint64_t)s.count | ((uint64_t)s.bad << 32);
}

// One CTA. Member verdicts, the check of the guesses, the records before every segment (base[k], exclusive scan of the counts)
// and the window's WinInfo.
// The chain enters segment k where it left segment k - 1 (segment 0 at byte 0, a record start by construction). Rounds over all
// segments in parallel: a segment that does not stand where its predecessor's chain ends is walked again from there (at most a
// segment's worth of records) -- but only once the predecessor itself stands where ITS predecessor ends, so a wrong guess never
// sends its neighbours off. A round that finds nothing to do proves the picture is the serial chain (the lowest segment out of
// place always has a predecessor in place, so it moves). Right guesses need one round; a wrong or missing guess two; a run of
// m wrong segments in a row (decoy records, a record spanning m segments) m + 1. If the rounds do not settle, one thread follows
// the chain through the window. Checked against the serial chain on the host (tests/hostsim/bam_chain_rounds.cpp).
__global__ void __launch_bounds__(1024) chain_resolve_kernel(const uint8_t* __restrict__ raw, uint64_t n, uint32_t nseg, uint32_t nmem, int last_window,
                                                             const int32_t* __restrict__ status, Segment* __restrict__ seg, uint32_t* __restrict__ base,
                                                             WinInfo* __restrict__ info, uint32_t* __restrict__ nrec_out) {
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_changed, s_bad, s_first_bad, s_carry, s_misses, s_flags;
    __shared__ WinInfo s_info;
    const uint32_t t = threadIdx.x;
    if (t == 0) { s_changed = 0; s_bad = 0; s_first_bad = 0xffffffffu; s_carry = 0; s_misses = 0; s_flags = 0; s_info.tail = n; s_info.err = 0; s_info.guess_misses = 0; }
    __syncthreads();
    for (uint32_t k = t; k < nmem; k += blockDim.x)
        if (status[k] != 0) { atomicAdd(&s_bad, 1u); atomicMin(&s_first_bad, k); }
    __syncthreads();
    if (s_bad) {                                    // a member did not inflate: nothing of this window is used
        if (t == 0) {
            s_info.err = E_MEMBER; s_info.nrec = 0; s_info.bad_members = s_bad; s_info.first_bad_member = s_first_bad;
            s_info.first_bad_status = status[s_first_bad];
            *info = s_info; *nrec_out = 0;
        }
        return;
    }
    bool settled = false;
    for (int round = 0; round < 96 && !settled; ++round) {
        for (uint32_t k = 1 + t; k < nseg; k += blockDim.x) {
            const Segment prev = seg[k - 1];
            const Segment s = seg[k];
            if (brec::segment_in_place(s, prev)) continue;
            s_changed = 1;
            // repair from the predecessor's end only if the predecessor itself stands where its own predecessor left the chain:
            // the end of a wrongly guessed chain can lie anywhere, and walking on from it would push the error down the window
            if (k > 1 && !brec::segment_in_place(prev, seg[k - 2])) continue;
            seg_store(seg + k, brec::segment_after(raw, n, seg_hi(k, n), prev));
            atomicAdd(&s_misses, 1u);
        }
        __syncthreads();
        settled = s_changed == 0;
        __syncthreads();
        if (t == 0) s_changed = 0;
        __syncthreads();
    }
    if (!settled && t == 0) {                       // one thread, the chain from the start
        uint64_t cur = 0;
        uint32_t ended = 0;
        for (uint32_t k = 0; k < nseg; ++k) {
            const uint64_t hi = seg_hi(k, n);
            Segment s;
            s.guess = cur; s.count = 0; s.bad = ended;
            uint64_t o = cur;
            while (!ended && o + 4 <= n && o < hi) {
                const uint32_t bs = brec::ld32(raw + o);
                if (bs < 32) { ended = 1; break; }
                if (o + 4 + (uint64_t)bs > n) { ended = 2; break; }
                ++s.count;
                o += 4 + (uint64_t)bs;
            }
            s.end = o; s.bad = ended;
            seg[k] = s;
            cur = o;
        }
        s_misses += nseg;
    }
    __syncthreads();
    // what the chain met, and the exclusive scan of the counts
    for (uint32_t k0 = 0; k0 < nseg; k0 += blockDim.x) {
        const uint32_t k = k0 + t;
        uint32_t v = 0;
        if (k < nseg) { const Segment s = seg[k]; v = s.count; if (s.bad) atomicOr(&s_flags, s.bad == 1 ? 1u : 2u); }
        uint32_t total;
        const uint32_t inc = bdk::ss_block_scan_any(v, s_warp, &total);
        if (k < nseg) base[k] = s_carry + inc - v;
        __syncthreads();
        if (t == 0) s_carry += total;
        __syncthreads();
    }
    if (t == 0) {
        s_info.tail = nseg ? seg[nseg - 1].end : 0;  // a cut-off record's start, or (fewer than 4 stray bytes aside) the end of the data
        if (s_flags & 1u) s_info.err |= E_RECORD;
        else if ((s_flags & 2u) && last_window) s_info.err |= E_TRUNCATED;
        s_info.guess_misses = s_misses;
        s_info.nrec = s_carry; s_info.bad_members = 0; s_info.first_bad_member = 0; s_info.first_bad_status = 0;
        *info = s_info; *nrec_out = s_carry;
    }
}

// One thread per segment: the offsets of the records' cores (behind block_size), in stream order.
__global__ void __launch_bounds__(128) chain_write_kernel(const uint8_t* __restrict__ raw, const Segment* __restrict__ seg, const uint32_t* __restrict__ base,
                                                          uint32_t nseg, uint32_t* __restrict__ rec_off) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nseg) return;
    const Segment s = seg[k];
    if (s.guess == NO_GUESS) return;
    uint64_t o = s.guess;
    uint32_t i = base[k];
    for (uint32_t r = 0; r < s.count; ++r) {
        rec_off[i++] = (uint32_t)(o + 4);
        o += 4 + (uint64_t)brec::ld32(raw + o);
    }
}

struct RgEntry { uint64_t hash; uint32_t id; uint32_t used; };     // read-group string (hashed) -> rgid, filled by the host

struct Columns {
    int32_t *pos, *mpos, *tid, *mtid, *isize, *qlen;
    uint16_t *flag, *rgid;
    uint8_t* mapq;
    uint64_t* qid;
};

// The reader's filter as the input of a device-wide scan, and the field extraction as its output: record i of the window goes to
// position (kept records before it) of the columns.
struct KeepFlag {
    const uint8_t* raw; const uint32_t* rec_off; brec::RegionSel sel;
    __device__ uint32_t operator()(uint32_t i, uint32_t) const { return brec::keep_record(raw + rec_off[i], sel) ? 1u : 0u; }
};
struct ExtractOut {
    const uint8_t* raw; const uint32_t* rec_off; const RgEntry* rg_table; uint32_t rg_slots; uint32_t rg_other; Columns out;
    __device__ void operator()(uint32_t i, uint32_t inc, uint32_t v, uint32_t) const {
        if (!v) return;
        const brec::Fields f = brec::record_fields(raw + rec_off[i]);
        const uint32_t o = inc - 1;
        out.pos[o] = f.pos; out.mpos[o] = f.mpos; out.tid[o] = f.tid; out.mtid[o] = f.mtid; out.isize[o] = f.isize; out.qlen[o] = f.qlen;
        out.flag[o] = f.flag; out.mapq[o] = f.mapq; out.qid[o] = f.qid;
        const uint64_t h = brec::hash_bytes(f.rg, f.rg_len);
        uint32_t id = rg_other;
        for (uint32_t p = (uint32_t)(h % rg_slots), tries = 0; tries < rg_slots; ++tries, p = p + 1 == rg_slots ? 0 : p + 1) {
            if (!rg_table[p].used) break;
            if (rg_table[p].hash == h) { id = rg_table[p].id; break; }
        }
        out.rgid[o] = (uint16_t)id;
    }
};

// Is the kept stream ordered by (reference sequence, position)? prev[2 * (w & 1)] holds the last record of window w - 1,
// prev[2 * ((w + 1) & 1)] receives this window's. unsorted[0] is set and stays set.
__global__ void __launch_bounds__(256) sorted_check_kernel(const int32_t* __restrict__ tid, const int32_t* __restrict__ pos, const uint32_t* __restrict__ n_ptr,
                                                           int32_t* __restrict__ prev, uint32_t window, uint32_t* __restrict__ unsorted) {
    const uint32_t n = *n_ptr;
    const int32_t* before = prev + 2 * (window & 1);
    int32_t* after = prev + 2 * ((window + 1) & 1);
    bool bad = false;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int32_t t0 = i ? tid[i - 1] : before[0], p0 = i ? pos[i - 1] : before[1];
        bad |= t0 > tid[i] || (t0 == tid[i] && p0 > pos[i]);
        if (i + 1 == n) { after[0] = tid[i]; after[1] = pos[i]; }
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) { after[0] = before[0]; after[1] = before[1]; }
    if (bad) atomicOr(unsorted, 1u);
}

}  // namespace bamdev
