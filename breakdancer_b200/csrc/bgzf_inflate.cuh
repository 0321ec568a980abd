// bgzf_inflate.cuh -- BGZF blocks inflated on the GPU (opt-in: BDK_GPU_INFLATE=1), the first stage of moving the BAM decode
// of the drop-in executable off the host (from BAM files the host's inflate is what bounds the
// whole run; the reference spends ~47 % of its wall time in zlib).
//
// BGZF is made for this: a BAM file is a sequence of independent raw-DEFLATE members of at most 64 KiB of output each, with
// the compressed and uncompressed sizes in the member's header and footer. One THREAD decodes one member (a 220 MB BAM has
// ~10^4 members, a 30x genome ~10^6: more members than the GPU has thread slots), so there is no cooperation inside a member
// and the decoder below is plain sequential code, cut into steps so that the lanes of a warp stay together (see member_step)
// -- `inflate_member`, the same steps in a plain loop, is compiled for the host as well and is tested there against zlib
// (tests/hostsim/gpu_inflate_host.cpp). What is GPU-specific is where the decoding tables live:
//   * the first-level lookup tables (9 bits literal/length, 6 bits distance; entry = symbol << 4 | code length) of all threads
//     of a CTA sit in shared memory, interleaved by thread ([entry][thread]), 1152 bytes per thread, 192 threads per CTA;
//   * codes longer than the first level are rare and are decoded canonically (count / sorted-symbol arrays as in the DEFLATE
//     specification's own decoding procedure) from a per-thread scratch block in global memory.
// Every access is bounds-checked (a damaged member stops with an error code, it never writes outside its own output range);
// the caller checks each member's CRC32 on the host and re-inflates with zlib whatever failed, so a defect here costs time,
// never correctness -- the same contract as the host's table-driven decoder (csrc/host/fast_inflate.hpp).
//
// Written from RFC 1951. First measurement: DESIGN.md section 8.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define BGZ_HD __host__ __device__ __forceinline__
#else
#define BGZ_HD inline
#endif

namespace bgz {

constexpr int LL_BITS = 9, D_BITS = 6;
constexpr int LL_LUT = 1 << LL_BITS, D_LUT = 1 << D_BITS;
constexpr int LUT_PER_THREAD = LL_LUT + D_LUT;           // uint16 entries
constexpr int CTA_THREADS = 192;

struct Member {                 // one BGZF member: where its DEFLATE stream is in the file, where its output goes
    uint64_t in_off;
    uint64_t out_off;
    uint32_t in_len;
    uint32_t out_len;
};

struct Scratch {                // per thread, global memory
    uint16_t cnt_ll[16], cnt_d[16];
    uint16_t sym_ll[288], sym_d[32];
    uint8_t lens[320];
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

BGZ_HD uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) { r = (r << 1) | (code & 1); code >>= 1; }
    return r;
}

// Canonical code from code lengths: counts per length, symbols sorted by (length, symbol), and the first-level table.
// Over-subscribed codes are refused; incomplete ones only if more than one code is present.
BGZ_HD bool build(const uint8_t* lens, int nsym, uint16_t* cnt, uint16_t* sym, uint16_t* lut, int lut_stride, int lut_bits) {
    for (int l = 0; l < 16; ++l) cnt[l] = 0;
    for (int s = 0; s < nsym; ++s) cnt[lens[s]]++;
    cnt[0] = 0;
    int left = 1, used = 0;
    for (int l = 1; l < 16; ++l) { left = (left << 1) - cnt[l]; if (left < 0) return false; used += cnt[l]; }
    if (left > 0 && used > 1) return false;
    uint16_t offs[16];
    uint32_t next_code[16];
    offs[1] = 0;
    uint32_t code = 0;
    next_code[0] = 0;
    for (int l = 1; l < 16; ++l) {
        if (l > 1) offs[l] = offs[l - 1] + cnt[l - 1];
        code = (code + (l > 1 ? cnt[l - 1] : 0)) << 1;
        next_code[l] = code;
    }
    const int lut_size = 1 << lut_bits;
    for (int i = 0; i < lut_size; ++i) lut[i * lut_stride] = 0;
    for (int s = 0; s < nsym; ++s) {
        const int l = lens[s];
        if (!l) continue;
        sym[offs[l]++] = (uint16_t)s;
        const uint32_t c = next_code[l]++;
        if (l <= lut_bits) {
            const uint16_t e = (uint16_t)(s << 4 | l);
            for (uint32_t i = reverse_bits(c, l); i < (uint32_t)lut_size; i += 1u << l) lut[i * lut_stride] = e;
        }
    }
    return true;
}

// one symbol; needs >= 15 valid bits. -1: no such code
BGZ_HD int decode(Bits& b, const uint16_t* lut, int lut_stride, int lut_bits, const uint16_t* cnt, const uint16_t* sym) {
    const uint16_t e = lut[(uint32_t)(b.buf & ((1u << lut_bits) - 1)) * lut_stride];
    if (e) { const int l = e & 15; b.buf >>= l; b.cnt -= l; return e >> 4; }
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; ++l) {
        code |= (int)((b.buf >> (l - 1)) & 1);
        const int c = cnt[l];
        if (code - c < first) { b.buf >>= l; b.cnt -= l; return sym[index + (code - first)]; }
        index += c; first += c;
        first <<= 1; code <<= 1;
    }
    return -1;
}

// ---- one member as a resumable state machine ------------------------------------------------------------------------------
// The 32 threads of a warp decode 32 different members. Written as one long loop per thread, the threads of a warp drift apart
// for good (measured: r03a/r03b, ~2*10^4 cycles per symbol -- 32 separate instruction streams on one scheduler slot). So the
// work is cut into STEPS -- read a block header and build its tables, or decode one symbol and do its copy -- and the kernel
// runs "every unfinished lane takes one step, then the warp meets again" (member_step + __syncwarp). Lanes that take different
// branches inside a step cost the sum of the branches, once per step, instead of never running together again.
struct MemberState {
    Bits b;
    uint32_t op;
    uint32_t copy_len, copy_dist;   // phase 3: what is left of a long match (a step copies at most 32 bytes, so that one lane's
                                    // 258-byte match does not hold the other 31 for 33 rounds of the copy loop)
    int phase;                  // 0: at a block header, 1: inside a block's symbols, 2: finished (status is final), 3: inside a match
    int status;
    bool last;
};

BGZ_HD void member_begin(MemberState& st, const uint8_t* in, uint32_t in_len, uint32_t out_len) {
    st.b.in = in; st.b.n = in_len; st.b.pos = 0; st.b.buf = 0; st.b.cnt = 0; st.b.ahead = 0; st.b.has_ahead = false;
    st.op = 0; st.last = false; st.status = OK; st.copy_len = 0; st.copy_dist = 0;
    st.phase = 0;
}

// lut: LL_LUT entries for literal/length codes followed by D_LUT entries for distance codes, entry i at lut[i * lut_stride]
BGZ_HD void member_step(MemberState& st, uint8_t* out, uint32_t out_len, Scratch& sc, uint16_t* lut, int lut_stride) {
    Bits& b = st.b;
    const uint32_t in_len = b.n;
    const uint8_t* in = b.in;
    uint16_t* lut_ll = lut;
    uint16_t* lut_d = lut + (size_t)LL_LUT * lut_stride;
#define BGZ_FAIL(code) do { st.status = (code); st.phase = 2; return; } while (0)
    if (st.phase == 0) {
        if (st.last) {                                               // after the final block
            if (b.pos - (uint32_t)(b.cnt >> 3) > in_len) BGZ_FAIL(ERR_INPUT);
            st.status = st.op == out_len ? OK : ERR_OUTPUT;
            st.phase = 2;
            return;
        }
        refill(b);
        st.last = take(b, 1) != 0;
        const uint32_t type = take(b, 2);
        if (type == 3) BGZ_FAIL(ERR_HEADER);
        if (type == 0) {                                             // stored: the whole block in this step
            take(b, b.cnt & 7);
            refill(b);
            const uint32_t len = take(b, 16);
            refill(b);
            const uint32_t nlen = take(b, 16);
            if ((len ^ 0xffffu) != nlen) BGZ_FAIL(ERR_HEADER);
            // whole bytes still in the bit buffer belong to the stored data
            const uint32_t src = b.pos - (uint32_t)(b.cnt >> 3);
            b.buf = 0; b.cnt = 0;
            if (src > in_len || in_len - src < len) BGZ_FAIL(ERR_INPUT);
            if (out_len - st.op < len) BGZ_FAIL(ERR_OUTPUT);
            for (uint32_t i = 0; i < len; ++i) out[st.op + i] = in[src + i];
            st.op += len; b.pos = src + len; b.has_ahead = false;
            return;                                                  // phase stays 0: next header, or the end
        }
        if (type == 1) {                                             // fixed codes
            for (int i = 0; i < 144; ++i) sc.lens[i] = 8;
            for (int i = 144; i < 256; ++i) sc.lens[i] = 9;
            for (int i = 256; i < 280; ++i) sc.lens[i] = 7;
            for (int i = 280; i < 288; ++i) sc.lens[i] = 8;
            if (!build(sc.lens, 288, sc.cnt_ll, sc.sym_ll, lut_ll, lut_stride, LL_BITS)) BGZ_FAIL(ERR_CODES);
            for (int i = 0; i < 32; ++i) sc.lens[i] = 5;
            if (!build(sc.lens, 32, sc.cnt_d, sc.sym_d, lut_d, lut_stride, D_BITS)) BGZ_FAIL(ERR_CODES);
        } else {                                                     // dynamic codes
            const uint8_t cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            refill(b);
            const int hlit = (int)take(b, 5) + 257, hdist = (int)take(b, 5) + 1, hclen = (int)take(b, 4) + 4;
            if (hlit > 286 || hdist > 30) BGZ_FAIL(ERR_HEADER);
            uint8_t cl[19];
            for (int i = 0; i < 19; ++i) cl[i] = 0;
            for (int i = 0; i < hclen; ++i) { refill(b); cl[cl_order[i]] = (uint8_t)take(b, 3); }
            // the code-length code (at most 7 bits, 19 symbols) borrows the distance tables
            if (!build(cl, 19, sc.cnt_d, sc.sym_d, lut_d, lut_stride, D_BITS)) BGZ_FAIL(ERR_CODES);
            int n = 0;
            while (n < hlit + hdist) {
                refill(b);
                const int s = decode(b, lut_d, lut_stride, D_BITS, sc.cnt_d, sc.sym_d);
                if (s < 0 || s > 18) BGZ_FAIL(ERR_CODES);
                if (s < 16) { sc.lens[n++] = (uint8_t)s; continue; }
                int rep, val = 0;
                if (s == 16) { if (n == 0) BGZ_FAIL(ERR_CODES); val = sc.lens[n - 1]; rep = 3 + (int)take(b, 2); }
                else if (s == 17) rep = 3 + (int)take(b, 3);
                else rep = 11 + (int)take(b, 7);
                if (n + rep > hlit + hdist) BGZ_FAIL(ERR_CODES);
                while (rep--) sc.lens[n++] = (uint8_t)val;
            }
            if (b.pos - (uint32_t)(b.cnt >> 3) > in_len) BGZ_FAIL(ERR_INPUT);
            if (sc.lens[256] == 0) BGZ_FAIL(ERR_CODES);
            // distance lengths first (they sit behind the literal/length ones), then the literal/length code
            if (!build(sc.lens + hlit, hdist, sc.cnt_d, sc.sym_d, lut_d, lut_stride, D_BITS)) BGZ_FAIL(ERR_CODES);
            if (!build(sc.lens, hlit, sc.cnt_ll, sc.sym_ll, lut_ll, lut_stride, LL_BITS)) BGZ_FAIL(ERR_CODES);
        }
        st.phase = 1;
        return;
    }
    // ---- one symbol -------------------------------------------------------------------------------------------------------
  if (st.phase != 3) {
    refill(b);
    int s = decode(b, lut_ll, lut_stride, LL_BITS, sc.cnt_ll, sc.sym_ll);
    if (s < 0) BGZ_FAIL(ERR_SYMBOL);
    if (s < 256) {
        if (st.op >= out_len) BGZ_FAIL(ERR_OUTPUT);
        out[st.op++] = (uint8_t)s;
        return;
    }
    if (s == 256) { st.phase = 0; return; }
    s -= 257;
    if (s >= 29) BGZ_FAIL(ERR_SYMBOL);
    uint32_t len;
    if (s < 8) len = 3 + s;
    else if (s == 28) len = 258;
    else { const int x = (s >> 2) - 1; len = ((4u + (s & 3)) << x) + 3 + take(b, x); }
    refill(b);
    const int d = decode(b, lut_d, lut_stride, D_BITS, sc.cnt_d, sc.sym_d);
    if (d < 0 || d >= 30) BGZ_FAIL(ERR_DISTANCE);
    uint32_t dist;
    if (d < 4) dist = 1 + d;
    else { const int x = (d >> 1) - 1; dist = ((2u + (d & 1)) << x) + 1 + take(b, x); }
    if (dist > st.op) BGZ_FAIL(ERR_DISTANCE);
    if (len > out_len - st.op) BGZ_FAIL(ERR_OUTPUT);
    if (b.pos - (uint32_t)(b.cnt >> 3) > in_len) BGZ_FAIL(ERR_INPUT);         // ran past the end of the input a while ago
    st.copy_len = len; st.copy_dist = dist;
  }
    // up to eight bytes are loaded before the first of them is stored (never more than `dist`, so no load needs a byte of the
    // same group): the loads are independent and their latencies overlap, instead of one round trip per byte
    const uint32_t dist = st.copy_dist;
    uint32_t len = st.copy_len < 32 ? st.copy_len : 32;
    const uint8_t* src = out + st.op - dist;
    uint8_t* dst = out + st.op;
    st.op += len;
    st.copy_len -= len;
    st.phase = st.copy_len ? 3 : 1;
    const uint32_t group = dist < 8 ? dist : 8;
    while (len) {
        const uint32_t k = len < group ? len : group;
        uint8_t v[8];
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) if (i < k) v[i] = src[i];
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) if (i < k) dst[i] = v[i];
        src += k; dst += k; len -= k;
    }
#undef BGZ_FAIL
}

// Inflate one member: exactly out_len bytes from in[0 .. in_len). The sequential driver of the state machine (host, tests).
BGZ_HD int inflate_member(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len, Scratch& sc, uint16_t* lut, int lut_stride) {
    MemberState st;
    member_begin(st, in, in_len, out_len);
    while (st.phase != 2) member_step(st, out, out_len, sc, lut, lut_stride);
    return st.status;
}

#ifdef __CUDACC__
// One thread per member. Consecutive members go to different CTAs (member m of a round to CTA m % grid, thread m / grid), so a
// file with fewer members than thread slots still spreads over all SMs. Every warp runs the same number of rounds, and inside
// a round its lanes advance step by step together. Dynamic shared memory: CTA_THREADS * LUT_PER_THREAD * 2 bytes.
__global__ void __launch_bounds__(CTA_THREADS, 1)
bgzf_inflate_kernel(const uint8_t* __restrict__ file, const Member* __restrict__ members, uint64_t n_members, uint8_t* __restrict__ out,
                    Scratch* __restrict__ scratch, int32_t* __restrict__ status) {
    extern __shared__ uint16_t bgz_lut[];
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    Scratch& sc = scratch[(uint64_t)blockIdx.x * blockDim.x + threadIdx.x];
    uint16_t* lut = bgz_lut + threadIdx.x;
    const uint64_t rounds = (n_members + nthreads - 1) / nthreads;
    for (uint64_t r = 0; r < rounds; ++r) {
        const uint64_t m = r * nthreads + (uint64_t)threadIdx.x * gridDim.x + blockIdx.x;
        const bool have = m < n_members;
        Member mb;
        mb.in_off = 0; mb.out_off = 0; mb.in_len = 0; mb.out_len = 0;
        if (have) mb = members[m];
        MemberState st;
        member_begin(st, file + mb.in_off, mb.in_len, mb.out_len);
        if (!have || mb.out_len == 0) st.phase = 2;
        uint8_t* dst = out + mb.out_off;
        while (__any_sync(0xffffffffu, st.phase != 2)) {
            if (st.phase != 2) member_step(st, dst, mb.out_len, sc, lut, (int)blockDim.x);
            __syncwarp();
        }
        if (have) status[m] = st.status;
    }
}
#endif

}  // namespace bgz
