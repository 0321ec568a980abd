// bgzf_inflate.cuh -- what the BGZF decoders share: the member descriptor, the status codes and the checked byte-wise bit reader
// (the reader of the host build of bgzf_inflate_warp.cuh, where the fuzz harness runs the decoder on exact-size buffers under
// AddressSanitizer). The thread-per-member kernel that used to live here (round 1: 6.2 of 32 lanes active, 14 GB/s) is gone:
// every device inflate -- bdk_push_bam and the opt-in BDK_GPU_INFLATE=1 stage of the host reader (bdk_bgzf_inflate) -- runs the
// warp-per-member decoder of bgzf_inflate_warp.cuh.
//
// BGZF: a BAM file is a sequence of independent raw-DEFLATE members of at most 64 KiB of output each, with the compressed and
// uncompressed sizes in the member's header and footer (RFC 1951 / RFC 1952).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define BGZ_HD __host__ __device__ __forceinline__
#else
#define BGZ_HD inline
#endif

namespace bgz {

struct Member {                 // one BGZF member: where its DEFLATE stream is in the file, where its output goes
    uint64_t in_off;
    uint64_t out_off;
    uint32_t in_len;
    uint32_t out_len;
};

enum Status { OK = 0, ERR_HEADER = 1, ERR_CODES = 2, ERR_SYMBOL = 3, ERR_DISTANCE = 4, ERR_OUTPUT = 5, ERR_INPUT = 6 };

struct Bits {
    const uint8_t* in;
    uint32_t n, pos;
    uint64_t buf;
    int cnt;
    uint32_t ahead;             // the aligned word at `pos`, loaded when the previous one was consumed (its latency hides behind
    bool has_ahead;             // the decoding of the bits in between)
};

// at least 33 valid bits afterwards (bytes past the end read as zero; running past the end is detected when the member ends)
BGZ_HD void refill(Bits& b) {
    while (b.cnt <= 32) {
        if (b.pos + 4 <= b.n && (((uintptr_t)(b.in + b.pos)) & 3) == 0) {
            const uint32_t w = b.has_ahead ? b.ahead : *(const uint32_t*)(b.in + b.pos);
            b.buf |= (uint64_t)w << b.cnt;
            b.cnt += 32; b.pos += 4;
            b.has_ahead = b.pos + 4 <= b.n;
            if (b.has_ahead) b.ahead = *(const uint32_t*)(b.in + b.pos);
        } else {
            b.has_ahead = false;
            const uint64_t v = b.pos < b.n ? b.in[b.pos] : 0;
            b.buf |= v << b.cnt;
            b.cnt += 8; b.pos += 1;
        }
    }
}
BGZ_HD uint32_t take(Bits& b, int n) {
    const uint32_t v = (uint32_t)(b.buf & ((1ull << n) - 1));
    b.buf >>= n; b.cnt -= n;
    return v;
}

}  // namespace bgz
