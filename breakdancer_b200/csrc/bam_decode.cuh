// bam_decode.cuh -- record boundaries and field extraction on inflated BAM bytes that stay in HBM: the stages between the GPU
// inflate (bgzf_inflate.cuh) and K1, so that a run from BAM files moves the COMPRESSED file over PCIe and nothing else
// (DESIGN.md section 9 item 4). The per-record logic is csrc/bam_records.h, the code the host decoder runs and its tests pin.
//
// STATUS: kernels only. They compile for sm_100a and are not launched by anything yet (no B200 time was left in the round
// they were written in); the orchestration -- header on the host, read-group table, bdk_push_device of the columns -- and the
// GPU tests come with the first measurement. Nothing in the product path depends on this file.
//
// Record boundaries, as on the host (bam_io.cpp find_records): the stream is cut into segments (one per BGZF member is the
// natural choice: the inflate kernel already knows their output offsets); every segment guesses its first record (three
// plausible records in a row) and follows the chain to its end; the guesses are right iff every segment's chain ends exactly
// on the next segment's guess, which one comparison per segment checks. A wrong or missing guess (records that contain
// record-like bytes, or a record longer than a segment) sends the file to the host path -- exactness never rests on the guess.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bam_records.h"

namespace bamdev {

using brec::Segment;
using brec::NO_GUESS;

// One thread per segment k = [cut[k], cut[k + 1]); cut[0] is the first record (after the BAM header), cut[nseg] = n.
__global__ void __launch_bounds__(128) chain_guess_kernel(const uint8_t* __restrict__ raw, uint64_t n, const uint64_t* __restrict__ cut, uint32_t nseg,
                                                          int32_t nref, Segment* __restrict__ seg) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nseg) seg[k] = brec::segment_guess(raw, n, cut[k], cut[k + 1], k == 0, nref);
}

// flags[0] != 0 afterwards: some guess was missing or wrong, or a chain hit a broken record -> the host decodes this file
__global__ void __launch_bounds__(256) chain_check_kernel(const Segment* __restrict__ seg, uint32_t nseg, uint64_t n, uint32_t* __restrict__ flags) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nseg && !brec::segment_consistent(seg, k, nseg, n)) atomicOr(flags, 1u);
}

// One thread per segment again, now with base[k] = number of records before the segment: writes the record offsets (of the
// cores, i.e. after block_size), the order of the stream.
__global__ void __launch_bounds__(128) chain_write_kernel(const uint8_t* __restrict__ raw, uint64_t n, const uint64_t* __restrict__ cut, uint32_t nseg,
                                                          const Segment* __restrict__ seg, const uint64_t* __restrict__ base, uint64_t* __restrict__ rec_off) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nseg) return;
    const uint64_t hi = cut[k + 1];
    uint64_t o = seg[k].guess, i = base[k];
    while (o + 4 <= n && o < hi) {
        rec_off[i++] = o + 4;
        o += 4 + (uint64_t)brec::ld32(raw + o);
    }
}

// One thread per record: the reader's filter (primary, placed, -o overlap).
__global__ void __launch_bounds__(256) keep_kernel(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ rec_off, uint64_t nrec, brec::RegionSel sel,
                                                   uint8_t* __restrict__ keep) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrec) return;
    keep[i] = brec::keep_record(raw + rec_off[i], sel) ? 1 : 0;
}

struct RgEntry { uint64_t hash; uint32_t id; uint32_t used; };     // read-group byte string (hashed) -> rgid, filled by the host

struct Columns {
    int32_t *pos, *mpos, *tid, *mtid, *isize, *qlen;
    uint16_t *flag, *rgid;
    uint8_t* mapq;
    uint64_t* qid;
};

// One thread per record; out_idx[i] = position of record i among the kept ones (exclusive scan of keep). Read groups the table
// does not know are reported (unknown[0] = count, unknown[1] = a record offset to look at) and get id 0xffff: the host adds
// them and the kernel runs again (a handful of distinct read groups per file).
__global__ void __launch_bounds__(256) extract_kernel(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ rec_off, uint64_t nrec,
                                                      const uint8_t* __restrict__ keep, const uint64_t* __restrict__ out_idx,
                                                      const RgEntry* __restrict__ rg_table, uint32_t rg_slots, Columns out,
                                                      unsigned long long* __restrict__ unknown) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrec || !keep[i]) return;
    const brec::Fields f = brec::record_fields(raw + rec_off[i]);
    const uint64_t o = out_idx[i];
    out.pos[o] = f.pos; out.mpos[o] = f.mpos; out.tid[o] = f.tid; out.mtid[o] = f.mtid; out.isize[o] = f.isize; out.qlen[o] = f.qlen;
    out.flag[o] = f.flag; out.mapq[o] = f.mapq; out.qid[o] = f.qid;
    const uint64_t h = brec::hash_bytes(f.rg, f.rg_len);
    uint32_t id = 0xffffu;
    for (uint32_t p = (uint32_t)(h % rg_slots), tries = 0; tries < rg_slots; ++tries, p = p + 1 == rg_slots ? 0 : p + 1) {
        if (!rg_table[p].used) break;
        if (rg_table[p].hash == h) { id = rg_table[p].id; break; }
    }
    if (id == 0xffffu) { atomicAdd(&unknown[0], 1ull); unknown[1] = rec_off[i]; }
    out.rgid[o] = (uint16_t)id;
}

}  // namespace bamdev
