// bgzf_inflate_warp.cuh -- BGZF members inflated on the GPU, one WARP per member (v3; bgzf_inflate.cuh is the thread-per-member
// decoder v2 of the opt-in host reader path). This is the decoder of the device-resident BAM decode (bam_device.cuh): inflated
// bytes stay in HBM, so only the compressed file crosses PCIe.
//
// Why a warp per member: v2's ncu capture (profiles/bgzf_inflate_r03d.md) shows 6.2 of 32 lanes active per instruction and 6
// warps per SM -- 32 members per warp diverge at every symbol, and the per-thread tables (36 KB of shared memory per warp) leave
// nothing to hide latency with. Here ALL 32 LANES DECODE THE SAME BIT STREAM REDUNDANTLY (same registers, same branches: no
// divergence, table look-ups and input loads are broadcasts), which costs the issue slots one decoding lane would, and
//   * a match is copied by the whole warp (a byte per lane and round: one load + one store for matches up to 32 bytes),
//   * the tables of a block are built by the whole warp (canonical decode of every table index, 32 indices at a time),
//   * tables are per WARP: 6 KB, so 32 warps (4 CTAs of 8) are resident per SM and hide the look-up / load latencies.
// Table entries carry base value and extra-bit count, so a length or distance needs no arithmetic on the symbol number.
// Codes longer than the first-level index (10 bits literal/length, 8 bits distance) are decoded canonically (count / sorted
// symbols, RFC 1951 section 3.2.2), rare by construction. Every access is bounds-checked: a damaged member stops with an error
// code and never writes outside its own output range. bgzf_crc_kernel then checks every member's CRC32 against its footer.
//
// The member decoder is written as "phases": sections whose 32 lanes work on different data, separated by warp barriers. On
// the host the same source runs the lanes of a phase one after the other (BGZW_PHASE), which is how tests/hostsim/
// gpu_inflate_warp_host.cpp fuzzes the decoder against zlib under AddressSanitizer. Written from RFC 1951 / RFC 1952.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "bgzf_inflate.cuh"     // bgz::Member, bgz::Bits / refill / take, bgz::Status

namespace bgzw {

using bgz::Bits;
using bgz::Member;

constexpr int LL_BITS = 11, D_BITS = 9, CL_BITS = 7;
constexpr int LL_SIZE = 1 << LL_BITS, D_SIZE = 1 << D_BITS;
#ifndef BGZW_WARPS_PER_CTA
#define BGZW_WARPS_PER_CTA 8
#endif
#ifndef BGZW_CTAS_PER_SM
#define BGZW_CTAS_PER_SM 4
#endif
constexpr int WARPS_PER_CTA = BGZW_WARPS_PER_CTA, CTA_THREADS = WARPS_PER_CTA * 32, CTAS_PER_SM = BGZW_CTAS_PER_SM;
constexpr int ERR_CRC = 7;

// First-level table entries, 16 bits. Bits 0-3: the code length; 0 = the first level does not settle it (LL_LONGER / D_LONGER: the
// code is longer than the index, or there is none: canonical_tail(); *_ILLEGAL: a code of at most the index width for a symbol
// that may not occur -- literal/length 286, 287, distance 30, 31). The other bits hold what the symbol MEANS, so that no
// arithmetic on symbol numbers and no second table is needed (RFC 1951 3.2.5):
//   literal/length  bit 12 clear: a literal, the byte in bits 4-11
//                   bit 12 set:   a length, 3 + B + extra with B = base - 3 (0..255) in bits 4-11 and the extra-bit count in bits
//                                 13-15; end of block: B = 255 with 7 extra bits (no length has more than 5), so a decoder that
//                                 treats it as a length sees one of 258 or more. Entries without a code length have bit 12 set
//                                 as well: one test separates the literals from everything else.
//   distance        1 + (m << x) + extra with x in bits 4-7 and m in bits 8-9 (symbols 0-3: m = symbol, x = 0; others: m = 2 + symbol % 2,
//                   x = symbol / 2 - 1)
constexpr uint32_t LL_NOT_LITERAL = 0x1000, LL_EOB = 0xfff0, LL_LONGER = 0x1000, LL_ILLEGAL = 0x1010, D_LONGER = 0, D_ILLEGAL = 0x10;
struct Tables {                    // one per warp, shared memory: 6.1 KB
    uint16_t ll[LL_SIZE];
    uint16_t d[D_SIZE];            // also the code-length code while a dynamic header is read
    uint16_t sym_ll[288], sym_d[32];
    uint16_t cnt_ll[16], cnt_d[16];
    uint8_t lens[320];
    uint16_t tail_first_ll, tail_index_ll, tail_first_d, tail_index_d;   // canonical() state behind the first-level widths (canonical_tail)
    uint32_t ok;                   // build verdict of the phase that checks the code lengths
};

#ifdef __CUDA_ARCH__
#define BGZW_HD __device__ __forceinline__
#define BGZW_LANE() ((int)(threadIdx.x & 31u))
#define BGZW_PHASE_BEGIN(lane) __syncwarp(); { const int lane = BGZW_LANE();
#define BGZW_PHASE_END() } __syncwarp();
#define BGZW_PHASE_END_OPEN() }
#define BGZW_LANE0(stmt) do { if (BGZW_LANE() == 0) { stmt; } } while (0)
#else
#define BGZW_HD inline
#define BGZW_PHASE_BEGIN(lane) for (int lane = 0; lane < 32; ++lane) {
#define BGZW_PHASE_END() }
#define BGZW_PHASE_END_OPEN() }
#define BGZW_LANE0(stmt) do { stmt; } while (0)
#endif

BGZW_HD uint32_t ll_entry(uint32_t sym, uint32_t len) {          // sym <= 285
    if (sym < 256) return sym << 4 | len;
    if (sym == 256) return LL_EOB | len;
    const uint32_t idx = sym - 257;
    uint32_t base, x;
    if (idx < 8) { base = idx; x = 0; }
    else if (idx == 28) { base = 255; x = 0; }
    else { x = (idx >> 2) - 1; base = (4 + (idx & 3)) << x; }
    return x << 13 | LL_NOT_LITERAL | base << 4 | len;
}
BGZW_HD uint32_t d_entry(uint32_t sym, uint32_t len) {           // sym <= 29
    const uint32_t x = sym < 4 ? 0 : (sym >> 1) - 1, m = sym < 4 ? sym : 2 + (sym & 1);
    return m << 8 | x << 4 | len;
}

// Bit reader of the warp decoder. Device: a window of three consecutive aligned words (lo, hi, hi2) of the input plus two more
// already requested (ahead, ahead2: a word is asked for two slides -- 64 input bits, half a dozen symbols -- before it is used;
// with one word in flight the warps sat 16 % of their stall samples on that load, profiles/inflate_warp_b20.md); the read
// position is a bit offset `off` into lo. normalize() slides the window while off >= 32, so behind it 64 + (32 - off) > 64 bits are valid: a whole literal/length code with its extra bits (<= 20 bits, peek()) and a whole
// distance code with its (<= 28 bits, peek2()) are read without looking at the window again. Extracting is one funnel shift, the
// 64-bit shift-and-or of a classic bit buffer is gone, and a load has two words of decoding to arrive. The
// compressed bytes sit in a buffer that is readable a few words beyond every member (its CRC32 / ISIZE footer and the next
// member follow; bdk_bam.inl pads the last one): words beyond `wlim` are not loaded, a damaged stream that runs on decodes the
// last word again until one of the over_end() checks or the output bound stops it. Host build (fuzz harness, exact-size buffers
// under AddressSanitizer): the checked byte-wise reader of bgzf_inflate.cuh behind the same interface.
#if defined(__CUDA_ARCH__) || defined(BGZW_WINDOW_READER_ON_HOST)
#define BGZW_WINDOW_READER 1
#endif
#ifndef __CUDA_ARCH__
inline uint32_t bgzw_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
inline uint32_t bgzw_min(uint32_t a, uint32_t b) { return a < b ? a : b; }
#else
__device__ __forceinline__ uint32_t bgzw_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
__device__ __forceinline__ uint32_t bgzw_min(uint32_t a, uint32_t b) { return min(a, b); }
#endif
struct Reader {
#ifdef BGZW_WINDOW_READER
    const uint32_t* w0;
    uint32_t lo, hi, hi2, ahead, ahead2;
    uint32_t off;               // bit offset of the read position in lo
    uint32_t widx, wlim;        // index (from w0) of the word in `ahead2`; last index that may be loaded
    uint32_t bits_skip;         // 32 * 4 + bits in front of the byte the window was opened at: the bits consumed since then are
                                // 32 * widx + off - bits_skip (nothing to keep up to date when the window slides)
    uint32_t origin;            // that byte's offset in the member (0, or where a stored block ended)
    BGZW_HD void init(const uint8_t* in, uint32_t n, uint32_t at = 0) {
        origin = at;
        const uint32_t mis = (uint32_t)((uintptr_t)in & 3);
        w0 = (const uint32_t*)(in - mis);
        wlim = (mis + n + 4) >> 2;                                  // the word that holds in[n + 4], inside the 8-byte footer
        lo = w0[0]; hi = w0[bgzw_min(1u, wlim)]; hi2 = w0[bgzw_min(2u, wlim)]; ahead = w0[bgzw_min(3u, wlim)]; ahead2 = w0[bgzw_min(4u, wlim)];
        widx = 4;
        off = 8 * mis;
        bits_skip = 128 + 8 * mis;
    }
    BGZW_HD void slide() {
        lo = hi; hi = hi2; hi2 = ahead; ahead = ahead2;
        ++widx;
        ahead2 = w0[bgzw_min(widx, wlim)];
        off -= 32;
    }
    // at most two slides are ever due: off < 32 after a normalize, and no more than 48 bits are dropped before the next one
    // (written as two tests: left as a loop, the compiler unrolls it sixteen-fold with look-ahead loads)
    BGZW_HD void normalize() { if (off >= 32) { slide(); if (__builtin_expect(off >= 32, 0)) slide(); } }
    BGZW_HD void normalize1() { if (off >= 32) slide(); }                                       // where off < 64 is known
    BGZW_HD uint32_t peek() const { return bgzw_funnel_r(lo, hi, off); }                      // off < 32
    BGZW_HD uint32_t peek2() const { return off < 32 ? bgzw_funnel_r(lo, hi, off) : bgzw_funnel_r(hi, hi2, off); }   // off < 64
    BGZW_HD void drop(uint32_t n) { off += n; }
    BGZW_HD uint32_t take(uint32_t n) { normalize(); const uint32_t v = peek() & ((1u << n) - 1u); off += n; return v; }   // n <= 16
    BGZW_HD bool over_end(uint32_t in_len) const { return 32 * widx + off > 8 * (in_len - origin) + bits_skip; }
    BGZW_HD uint32_t consumed() const { return origin + ((32 * widx + off - bits_skip + 7) >> 3); }    // bytes of the member, a started byte counts
    BGZW_HD void align_byte() { off = (off + 7) & ~7u; }
    BGZW_HD void seek(const uint8_t* in, uint32_t at, uint32_t n) { init(in + at, n - at, at); }
#else
    Bits b;
    void init(const uint8_t* in, uint32_t n) { b.in = in; b.n = n; b.pos = 0; b.buf = 0; b.cnt = 0; b.ahead = 0; b.has_ahead = false; }
    void normalize() { bgz::refill(b); }
    void normalize1() { bgz::refill(b); }
    uint32_t peek() const { return (uint32_t)b.buf; }
    uint32_t peek2() { bgz::refill(b); return (uint32_t)b.buf; }
    void drop(uint32_t n) { b.buf >>= n; b.cnt -= (int)n; }
    uint32_t take(uint32_t n) { bgz::refill(b); return bgz::take(b, (int)n); }
    bool over_end(uint32_t in_len) const { return b.pos - (uint32_t)(b.cnt >> 3) > in_len; }
    uint32_t consumed() const { return b.pos - (uint32_t)(b.cnt >> 3); }
    void align_byte() { drop((uint32_t)(b.cnt & 7)); }
    void seek(const uint8_t*, uint32_t at, uint32_t) { b.pos = at; b.buf = 0; b.cnt = 0; b.has_ahead = false; }
#endif
};

// j mod d for the offsets of an overlapping match (j < 258, 1 <= d < 258) without an integer division: (j + 0.5) / d is at
// least 0.5 / 258 away from every integer, the quotient is below 258, so a division that is good to a few ulp (the device's
// approximate one: 2 ulp) truncates to floor(j / d).
BGZW_HD uint32_t small_mod(uint32_t j, uint32_t d) {
#ifdef __CUDA_ARCH__
    const uint32_t qt = (uint32_t)__fdividef((float)j + 0.5f, (float)d);
#else
    const uint32_t qt = (uint32_t)(((float)j + 0.5f) / (float)d);
#endif
    return j - qt * d;
}

// Canonical decode of the code that starts at bit 0 of `bits` (first bit of the code = bit 0), looking at code lengths up to
// maxlen. Returns the symbol and its length, or -1.
BGZW_HD int canonical(uint32_t bits, int maxlen, const uint16_t* cnt, const uint16_t* sym, int* len_out) {
    int code = 0, first = 0, index = 0;
#pragma unroll 1
    for (int l = 1; l <= maxlen; ++l) {
        code |= (int)((bits >> (l - 1)) & 1);
        const int c = cnt[l];
        if (code - c < first) { *len_out = l; return sym[index + (code - first)]; }
        index += c; first += c;
        first <<= 1; code <<= 1;
    }
    return -1;
}

// The same decode entered behind the first l0 lengths, for a code the first-level table has no entry for (so none of the first l0
// lengths is a hit): `first` and `index` as canonical() has them after l0 rounds (Tables::tail_*), the code so far is the first l0
// bits, first bit most significant.
BGZW_HD uint32_t bit_reverse(uint32_t v, int n) {                // the low n bits of v, reversed
#ifdef __CUDA_ARCH__
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}
BGZW_HD int canonical_tail(uint32_t bits, int l0, int first, int index, const uint16_t* cnt, const uint16_t* sym, int* len_out) {
    int code = (int)(bit_reverse(bits, l0) << 1);
#pragma unroll 1
    for (int l = l0 + 1; l <= 15; ++l) {
        code |= (int)((bits >> (l - 1)) & 1);
        const int c = cnt[l];
        if (code - c < first) { *len_out = l; return sym[index + (code - first)]; }
        index += c; first += c;
        first <<= 1; code <<= 1;
    }
    return -1;
}

// Counts per code length and symbols sorted by (length, symbol) for the literal/length code (lanes 0-15: lane = length) and the
// distance code (lanes 16-31) at once, then the over-subscription / completeness verdict, then the first-level tables.
// `cl`: build only a code-length code of 19 symbols from T.lens into the distance table.
BGZW_HD bool build_tables(Tables& T, int hlit, int hdist, bool cl) {
    BGZW_PHASE_BEGIN(lane)
        const int l = lane & 15;
        const bool dist = lane >= 16;
        if (!(cl && !dist)) {
            const uint8_t* lens = T.lens + (dist && !cl ? hlit : 0);
            const int nsym = cl ? 19 : dist ? hdist : hlit;
            uint16_t* cnt = dist ? T.cnt_d : T.cnt_ll;
            int c = 0;
            if (l) for (int s = 0; s < nsym; ++s) c += lens[s] == l;
            cnt[l] = (uint16_t)c;
        }
    BGZW_PHASE_END()
    BGZW_PHASE_BEGIN(lane)
        const int l = lane & 15;
        const bool dist = lane >= 16;
        if (l && !(cl && !dist)) {
            const uint8_t* lens = T.lens + (dist && !cl ? hlit : 0);
            const int nsym = cl ? 19 : dist ? hdist : hlit;
            const uint16_t* cnt = dist ? T.cnt_d : T.cnt_ll;
            uint16_t* sym = dist ? T.sym_d : T.sym_ll;
            int o = 0;
            for (int k = 1; k < l; ++k) o += cnt[k];
            if (cnt[l]) for (int s = 0; s < nsym; ++s) if (lens[s] == l) sym[o++] = (uint16_t)s;
        }
        if (lane == 0) {
            bool ok = true;
            for (int t = cl ? 1 : 0; t < 2; ++t) {
                const uint16_t* cnt = t ? T.cnt_d : T.cnt_ll;
                int left = 1, used = 0;
                for (int k = 1; k < 16; ++k) { left = (left << 1) - cnt[k]; if (left < 0) { ok = false; break; } used += cnt[k]; }
                if (left > 0 && used > 1) ok = false;      // incomplete codes only with a single code (RFC 1951 3.2.7, as zlib)
            }
            T.ok = ok ? 1u : 0u;
            if (!cl) {
                int first = 0, index = 0;
                for (int k = 1; k <= LL_BITS; ++k) { index += T.cnt_ll[k]; first = (first + T.cnt_ll[k]) << 1; }
                T.tail_first_ll = (uint16_t)first; T.tail_index_ll = (uint16_t)index;
                first = 0; index = 0;
                for (int k = 1; k <= D_BITS; ++k) { index += T.cnt_d[k]; first = (first + T.cnt_d[k]) << 1; }
                T.tail_first_d = (uint16_t)first; T.tail_index_d = (uint16_t)index;
            }
        }
    BGZW_PHASE_END()
    if (!T.ok) return false;
    BGZW_PHASE_BEGIN(lane)
        if (cl) {
            for (int i = lane; i < (1 << CL_BITS); i += 32) {
                int len = 0;
                const int s = canonical((uint32_t)i, CL_BITS, T.cnt_d, T.sym_d, &len);
                T.d[i] = s < 0 ? (uint16_t)0 : (uint16_t)(s << 4 | len);
            }
        } else {
            for (int i = lane; i < LL_SIZE; i += 32) {
                int len = 0;
                const int s = canonical((uint32_t)i, LL_BITS, T.cnt_ll, T.sym_ll, &len);
                T.ll[i] = (uint16_t)(s < 0 ? LL_LONGER : s > 285 ? LL_ILLEGAL : ll_entry((uint32_t)s, (uint32_t)len));
            }
            for (int i = lane; i < D_SIZE; i += 32) {
                int len = 0;
                const int s = canonical((uint32_t)i, D_BITS, T.cnt_d, T.sym_d, &len);
                T.d[i] = (uint16_t)(s < 0 ? D_LONGER : s > 29 ? D_ILLEGAL : d_entry((uint32_t)s, (uint32_t)len));
            }
        }
    BGZW_PHASE_END()
    return true;
}

#if defined(__CUDA_ARCH__) && !defined(BGZW_NO_FAST_LOOP)
#define BGZW_FAST_LOOP 1
// The symbols the first-level tables settle -- literals, and matches of up to 32 bytes whose length and distance codes are in
// the tables -- decoded in a loop written in PTX, for two things the compiler will not do with the C++ below: every branch is
// `bra.uni` (all 32 lanes hold the same decoder state, so no branch here ever diverges; compiled from C++ each `if` is wrapped
// in reconvergence instructions and every way out of the loop costs a chain of them at every check), and the window registers
// are rotated in place. It stops IN FRONT of the first symbol it leaves to the C++ step that follows it (a longer or illegal
// code, end of block, a long match, anything that fails a check) with the reader state and `op` as that step expects them, so
// what the loop does NOT handle needs no cases here. Invariants as in the C++ loop: off < 48 on entry and at the top, one slide
// there (with the bounds check of every 32 input bits), off <= 51 where the distance code is looked at, <= 79 behind a match
// (one slide there). 17 instructions per literal, about 60 per match (the C++ loop: 31 and 95).
__device__ __forceinline__ void fast_symbols(Reader& b, uint32_t& op, uint8_t* out, uint32_t out_len, uint32_t in_len, const Tables& T) {
    const uint32_t lane = threadIdx.x & 31u;
    const float lane_half = (float)lane + 0.5f;
    const uint32_t t_ll = (uint32_t)__cvta_generic_to_shared(T.ll), t_d = (uint32_t)__cvta_generic_to_shared(T.d);
    const uint32_t bitlim = 8 * (in_len - b.origin) + b.bits_skip;     // over_end(): 32 * widx + off > bitlim
    asm volatile(
        "{\n\t"
        ".reg .pred p, pq, ppend;\n\t"
        ".reg .b32 v, a, e, l, t, u, lx, len, f, dl, dx, dist, noff, so, j, pval;\n\t"
        ".reg .b64 ad, pad;\n\t"
        "setp.eq.u32 ppend, %15, 0xffffffff;\n\t"           // no store waiting
        "mov.b32 pval, 0;\n\t"
        "mov.b64 pad, %10;\n\t"
        ".reg .f32 fd, fq, fj;\n\t"
        "setp.lt.u32 p, %4, 32;\n\t"
        "@p bra.uni BGZW_LOOK;\n\t"
        // slide the window by one word; the bounds of input and output, once per 32 bits
        "BGZW_SLIDE:\n\t"
        "mov.b32 %0, %1;\n\t"
        "mov.b32 %1, %2;\n\t"
        "mov.b32 %2, %3;\n\t"
        "mov.b32 %3, %7;\n\t"
        "add.u32 %5, %5, 1;\n\t"
        "min.u32 t, %5, %8;\n\t"
        "mad.wide.u32 ad, t, 4, %9;\n\t"
        "ld.global.u32 %7, [ad];\n\t"
        "sub.u32 %4, %4, 32;\n\t"
        "setp.gt.u32 p, %6, %11;\n\t"
        "shl.b32 t, %5, 5;\n\t"
        "add.u32 t, t, %4;\n\t"
        "setp.gt.or.u32 p, t, %12, p;\n\t"
        "@p bra.uni BGZW_OUT;\n\t"
        "BGZW_LOOK:\n\t"
        "shf.r.wrap.b32 v, %0, %1, %4;\n\t"
        "and.b32 a, v, 2047;\n\t"
        "mad.lo.u32 a, a, 2, %13;\n\t"
        "ld.shared.u16 e, [a];\n\t"
        "and.b32 l, e, 15;\n\t"
        "and.b32 t, e, 0x1000;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "@p bra.uni BGZW_MATCH;\n\t"
        // a literal (every lane stores the same byte to the same place)
        "add.u32 %4, %4, l;\n\t"
        "shr.u32 t, e, 4;\n\t"
        "mad.wide.u32 ad, %6, 1, %10;\n\t"
        "st.global.u8 [ad], t;\n\t"
        "add.u32 %6, %6, 1;\n\t"
        "setp.lt.u32 p, %4, 32;\n\t"
        "@p bra.uni BGZW_LOOK;\n\t"
        "bra.uni BGZW_SLIDE;\n\t"
        "BGZW_MATCH:\n\t"
        // no code length in the entry (a longer code, an illegal symbol) or the end of the block (7 "extra bits"): the long way
        "setp.eq.u32 p, l, 0;\n\t"
        "setp.ge.or.u32 p, e, 0xf000, p;\n\t"
        "@p bra.uni BGZW_OUT;\n\t"
        "shr.u32 lx, e, 13;\n\t"
        "shr.u32 len, e, 4;\n\t"
        "and.b32 len, len, 255;\n\t"
        "shr.u32 t, v, l;\n\t"
        "shl.b32 u, 0xffffffff, lx;\n\t"
        "not.b32 u, u;\n\t"
        "and.b32 t, t, u;\n\t"
        "add.u32 len, len, t;\n\t"
        "add.u32 len, len, 3;\n\t"
        "add.u32 noff, %4, l;\n\t"
        "add.u32 noff, noff, lx;\n\t"
        // the distance code: noff < 64
        "setp.lt.u32 p, noff, 32;\n\t"
        "@p shf.r.wrap.b32 v, %0, %1, noff;\n\t"
        "@!p shf.r.wrap.b32 v, %1, %2, noff;\n\t"
        "and.b32 a, v, 511;\n\t"
        "mad.lo.u32 a, a, 2, %14;\n\t"
        "ld.shared.u16 f, [a];\n\t"
        "and.b32 dl, f, 15;\n\t"
        "shr.u32 dx, f, 4;\n\t"
        "and.b32 dx, dx, 15;\n\t"
        "shr.u32 dist, f, 8;\n\t"
        "shl.b32 dist, dist, dx;\n\t"
        "shr.u32 t, v, dl;\n\t"
        "shl.b32 u, 0xffffffff, dx;\n\t"
        "not.b32 u, u;\n\t"
        "and.b32 t, t, u;\n\t"
        "add.u32 dist, dist, t;\n\t"
        "add.u32 dist, dist, 1;\n\t"
        "add.u32 noff, noff, dl;\n\t"
        "add.u32 noff, noff, dx;\n\t"
        // no entry for the distance, a distance beyond the start, a match beyond the end: the long way
        "setp.eq.u32 p, dl, 0;\n\t"
        "setp.gt.or.u32 p, dist, %6, p;\n\t"
        "add.u32 t, %6, len;\n\t"
        "setp.gt.or.u32 p, t, %11, p;\n\t"
        "@p bra.uni BGZW_OUT;\n\t"
        // the copy: byte j of the match is byte (j mod dist) of the dist bytes in front of it, 32 bytes a round, lane = j mod 32.
        // The store of a match's last round waits in registers (pad, pval, ppend) until the next match begins (or the loop
        // ends): issued right behind its load it would hold the warp for the load's whole latency at every match; this way
        // the next symbols are decoded meanwhile. Nothing reads those bytes before the next match's loads.
        "bar.warp.sync 0xffffffff;\n\t"
        "@ppend st.global.u8 [pad], pval;\n\t"
        "sub.u32 so, %6, dist;\n\t"
        "mov.b32 j, %15;\n\t"
        "setp.ge.u32 pq, dist, len;\n\t"
        "@pq bra.uni BGZW_PLAIN;\n\t"
        // source and destination overlap (dist < len <= 258): (j + 0.5) / dist truncates to floor(j / dist) with a quotient good
        // to 2 ulp (small_mod())
        "cvt.rn.f32.u32 fd, dist;\n\t"
        "rcp.approx.ftz.f32 fd, fd;\n\t"
        "mov.f32 fq, %16;\n\t"
        "BGZW_OVER:\n\t"
        "setp.lt.u32 ppend, j, len;\n\t"
        "mul.ftz.f32 fj, fq, fd;\n\t"
        "cvt.rzi.u32.f32 u, fj;\n\t"
        "mul.lo.u32 u, u, dist;\n\t"
        "sub.u32 u, j, u;\n\t"
        "add.u32 u, u, so;\n\t"
        "mad.wide.u32 ad, u, 1, %10;\n\t"
        "ld.global.u8 pval, [ad];\n\t"              // (lanes beyond the match read a byte in front of it: valid, unused)
        "add.u32 a, %6, j;\n\t"
        "mad.wide.u32 pad, a, 1, %10;\n\t"
        "add.u32 j, j, 32;\n\t"
        "add.f32 fq, fq, 0f42000000;\n\t"
        "sub.u32 a, j, %15;\n\t"
        "setp.lt.u32 p, a, len;\n\t"
        "@!p bra.uni BGZW_COPIED;\n\t"
        "@ppend st.global.u8 [pad], pval;\n\t"
        "bra.uni BGZW_OVER;\n\t"
        "BGZW_PLAIN:\n\t"
        "setp.lt.u32 ppend, j, len;\n\t"
        "add.u32 u, so, j;\n\t"
        "mad.wide.u32 ad, u, 1, %10;\n\t"
        "ld.global.u8 pval, [ad];\n\t"
        "add.u32 a, %6, j;\n\t"
        "mad.wide.u32 pad, a, 1, %10;\n\t"
        "add.u32 j, j, 32;\n\t"
        "sub.u32 a, j, %15;\n\t"
        "setp.lt.u32 p, a, len;\n\t"
        "@!p bra.uni BGZW_COPIED;\n\t"
        "@ppend st.global.u8 [pad], pval;\n\t"
        "bra.uni BGZW_PLAIN;\n\t"
        "BGZW_COPIED:\n\t"
        "mov.b32 %6, t;\n\t"
        "mov.b32 %4, noff;\n\t"
        "setp.lt.u32 p, noff, 32;\n\t"
        "@p bra.uni BGZW_LOOK;\n\t"
        "mov.b32 %0, %1;\n\t"
        "mov.b32 %1, %2;\n\t"
        "mov.b32 %2, %3;\n\t"
        "mov.b32 %3, %7;\n\t"
        "add.u32 %5, %5, 1;\n\t"
        "min.u32 t, %5, %8;\n\t"
        "mad.wide.u32 ad, t, 4, %9;\n\t"
        "ld.global.u32 %7, [ad];\n\t"
        "sub.u32 %4, %4, 32;\n\t"
        "setp.lt.u32 p, %4, 32;\n\t"
        "@p bra.uni BGZW_LOOK;\n\t"
        "bra.uni BGZW_SLIDE;\n\t"
        "BGZW_OUT:\n\t"
        "@ppend st.global.u8 [pad], pval;\n\t"
        "}"
        : "+r"(b.lo), "+r"(b.hi), "+r"(b.hi2), "+r"(b.ahead), "+r"(b.off), "+r"(b.widx), "+r"(op), "+r"(b.ahead2)
        : "r"(b.wlim), "l"(b.w0), "l"(out), "r"(out_len), "r"(bitlim), "r"(t_ll), "r"(t_d), "r"(lane), "f"(lane_half)
        : "memory");
}
#endif

// Inflate one member: exactly out_len bytes from in[0 .. in_len). Called by all 32 lanes of a warp with the same arguments.
// `out` must be writable for 512 bytes beyond out_len: a damaged stream is stopped at the next window slide or match, not at
// every literal (the member is then refused; what it scribbled over belongs to a job that fails).
BGZW_HD int inflate_member(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len, Tables& T) {
    Reader b;
    b.init(in, in_len);
    uint32_t op = 0;
    bool last = false;
    while (!last) {
        if (b.over_end(in_len)) return bgz::ERR_INPUT;
        last = b.take(1) != 0;
        const uint32_t type = b.take(2);
        if (type == 3) return bgz::ERR_HEADER;
        if (type == 0) {                                             // stored
            b.align_byte();
            const uint32_t len = b.take(16);
            const uint32_t nlen = b.take(16);
            if ((len ^ 0xffffu) != nlen) return bgz::ERR_HEADER;
            const uint32_t src = b.consumed();                       // whole bytes still in the bit buffer belong to the data
            if (src > in_len || in_len - src < len) return bgz::ERR_INPUT;
            if (op > out_len || out_len - op < len) return bgz::ERR_OUTPUT;
            BGZW_PHASE_BEGIN(lane)
                for (uint32_t i = (uint32_t)lane; i < len; i += 32) out[op + i] = in[src + i];
            BGZW_PHASE_END()
            op += len;
            b.seek(in, src + len, in_len);
            continue;
        }
        int hlit = 288, hdist = 32;
        if (type == 1) {                                             // fixed codes
            BGZW_PHASE_BEGIN(lane)
                for (int i = lane; i < 320; i += 32) T.lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5);
            BGZW_PHASE_END()
        } else {                                                     // dynamic codes
            hlit = (int)b.take(5) + 257; hdist = (int)b.take(5) + 1;
            const int hclen = (int)b.take(4) + 4;
            if (hlit > 286 || hdist > 30) return bgz::ERR_HEADER;
            BGZW_PHASE_BEGIN(lane)
                if (lane < 19) T.lens[lane] = 0;
            BGZW_PHASE_END()
            for (int i = 0; i < hclen; ++i) {
                // 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 (RFC 1951 3.2.7), five bits each
                const uint64_t order_lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 | 10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
                const uint64_t order_hi = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
                const int which = (int)((i < 12 ? order_lo >> (5 * i) : order_hi >> (5 * (i - 12))) & 31);
                const uint8_t v = (uint8_t)b.take(3);
                BGZW_LANE0(T.lens[which] = v);
            }
            if (!build_tables(T, 0, 0, true)) return bgz::ERR_CODES;
            int n = 0, prev = 0;
            while (n < hlit + hdist) {
                b.normalize();
                const uint32_t e = T.d[b.peek() & ((1u << CL_BITS) - 1)];
                if (!e) return bgz::ERR_CODES;
                b.drop(e & 15);
                const int s = (int)(e >> 4);
                if (s < 16) { BGZW_LANE0(T.lens[n] = (uint8_t)s); prev = s; ++n; continue; }
                int rep, val = 0;
                if (s == 16) { if (n == 0) return bgz::ERR_CODES; val = prev; rep = 3 + (int)b.take(2); }
                else if (s == 17) rep = 3 + (int)b.take(3);
                else rep = 11 + (int)b.take(7);
                if (n + rep > hlit + hdist) return bgz::ERR_CODES;
                BGZW_PHASE_BEGIN(lane)
                    for (int i = lane; i < rep; i += 32) T.lens[n + i] = (uint8_t)val;
                BGZW_PHASE_END()
                n += rep; prev = val;
            }
            if (b.over_end(in_len)) return bgz::ERR_INPUT;
        }
        // lens[256] is read by every lane after the barrier that build_tables starts with
        if (!build_tables(T, hlit, hdist, false)) return bgz::ERR_CODES;
        if (T.lens[256] == 0) return bgz::ERR_CODES;
        // ---- the symbols of the block ----
        // One loop, one way out: what goes wrong sets `st` and leaves the loop (early returns from inside it cost a chain of
        // predicated-off reconvergence instructions at every check, every symbol). Window reader: off < 32 at the top, <= 47
        // behind a literal, <= 51 when the distance code is looked at (peek2), <= 79 behind a match (one slide there, one at
        // the top); the bounds are looked at once per slide at the top and at every match.
        int st = bgz::OK;
        for (;;) {
#ifdef BGZW_FAST_LOOP
            fast_symbols(b, op, out, out_len, in_len, T);            // everything the tables settle; then one symbol the long way
#endif
#ifdef BGZW_WINDOW_READER
            {
#ifdef BGZW_FAST_LOOP
                bool look = true;                                    // the fast loop may have stopped because of the bounds
#else
                bool look = false;
#endif
                if (b.off >= 32) { b.slide(); look = true; }
                if (look && (op > out_len || b.over_end(in_len))) { st = op > out_len ? bgz::ERR_OUTPUT : bgz::ERR_INPUT; break; }
            }
#else
            b.normalize();
            if (op > out_len) { st = bgz::ERR_OUTPUT; break; }
#endif
            const uint32_t p = b.peek();
            uint32_t e = T.ll[p & (LL_SIZE - 1)];
            if ((e & 15) == 0) {                                     // a code longer than the table index, an illegal symbol, or no code
                int len = 0;
                const int s = e != LL_LONGER ? -1 : canonical_tail(p, LL_BITS, T.tail_first_ll, T.tail_index_ll, T.cnt_ll, T.sym_ll, &len);
                if (s < 0 || s > 285) { st = bgz::ERR_SYMBOL; break; }
                e = ll_entry((uint32_t)s, (uint32_t)len);
            }
            const uint32_t l = e & 15;
            if (!(e & LL_NOT_LITERAL)) {
                b.drop(l);
#ifndef BGZW_WINDOW_READER
                if (op >= out_len) { st = bgz::ERR_OUTPUT; break; }  // checked reader: exact buffers
#endif
                BGZW_LANE0(out[op] = (uint8_t)(e >> 4));
                ++op;
                continue;
            }
            if (e >= LL_EOB) { b.drop(l); break; }
            const uint32_t lx = e >> 13;
            const uint32_t len = 3 + ((e >> 4) & 255u) + ((p >> l) & ((1u << lx) - 1u));
            b.drop(l + lx);
            const uint32_t q = b.peek2();
            uint32_t f = T.d[q & (D_SIZE - 1)];
            if ((f & 15) == 0) {
                int dl = 0;
                const int s = f != D_LONGER ? -1 : canonical_tail(q, D_BITS, T.tail_first_d, T.tail_index_d, T.cnt_d, T.sym_d, &dl);
                if (s < 0 || s > 29) { st = bgz::ERR_DISTANCE; break; }
                f = d_entry((uint32_t)s, (uint32_t)dl);
            }
            const uint32_t dl = f & 15, dx = (f >> 4) & 15;
            const uint32_t dist = 1 + ((f >> 8) << dx) + ((q >> dl) & ((1u << dx) - 1u));
            b.drop(dl + dx);
            if (dist > op || op + len > out_len) { st = dist > op ? bgz::ERR_DISTANCE : bgz::ERR_OUTPUT; break; }
            // byte j of the match is byte (j mod dist) of the `dist` bytes before it: every lane reads bytes that were complete
            // before this match began (the barrier that opens the phase orders them after the stores of all lanes). No barrier
            // behind it: until the next match's, lane 0 only stores literals beyond this match, where nothing is read.
            BGZW_PHASE_BEGIN(lane)
                const uint32_t so = op - dist;
                if (len <= 32) {
                    if ((uint32_t)lane < len) out[op + (uint32_t)lane] = out[so + (dist < len ? small_mod((uint32_t)lane, dist) : (uint32_t)lane)];
                } else if (dist >= len) {
                    for (uint32_t j = (uint32_t)lane; j < len; j += 32) out[op + j] = out[so + j];
                } else {
                    for (uint32_t j = (uint32_t)lane; j < len; j += 32) out[op + j] = out[so + small_mod(j, dist)];
                }
            BGZW_PHASE_END_OPEN()
            op += len;
            b.normalize1();
        }
        if (st != bgz::OK) return st;
    }
    if (b.over_end(in_len)) return bgz::ERR_INPUT;
    return op == out_len ? bgz::OK : bgz::ERR_OUTPUT;
}

// ---- CRC-32 (RFC 1952 section 8; reflected polynomial 0xEDB88320) of a member's output, 32 slices combined -------------------
// The CRC of a concatenation A|B is crc(A) * x^(8|B|) + crc(B) in GF(2)[x] modulo the polynomial (init and final complement
// cancel), so every lane takes a slice, multiplies its CRC by x^(8 * bytes behind the slice), and the warp XORs the products.
constexpr uint32_t CRC_POLY = 0xEDB88320u;
BGZW_HD uint32_t crc_mulmod(uint32_t a, uint32_t b) {            // a * b mod P, bit 31 = x^0
    uint32_t p = 0;
    for (uint32_t m = 1u << 31; m; m >>= 1) {
        if (a & m) p ^= b;
        b = (b & 1) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}
BGZW_HD uint32_t crc_x_pow8n(uint32_t nbytes) {                   // x^(8 * nbytes) mod P by square and multiply
    uint32_t r = 1u << 31;                                         // x^0
    uint32_t sq = 1u << 23;                                        // x^8
    while (nbytes) {
        if (nbytes & 1) r = crc_mulmod(sq, r);
        sq = crc_mulmod(sq, sq);
        nbytes >>= 1;
    }
    return r;
}
BGZW_HD uint32_t crc_table_entry(uint32_t k) {
    uint32_t c = k;
    for (int i = 0; i < 8; ++i) c = (c & 1) ? (c >> 1) ^ CRC_POLY : c >> 1;
    return c;
}
BGZW_HD uint32_t crc_slice(const uint32_t* table, const uint8_t* p, uint32_t n) {
    uint32_t c = 0xffffffffu;
    for (uint32_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 255] ^ (c >> 8);
    return c ^ 0xffffffffu;
}
// lane's share of crc32(p[0 .. n)): XOR over the 32 lanes gives the CRC
BGZW_HD uint32_t crc_lane_part(const uint32_t* table, const uint8_t* p, uint32_t n, int lane) {
    const uint32_t chunk = (n + 31) / 32;
    const uint32_t lo = (uint32_t)lane * chunk < n ? (uint32_t)lane * chunk : n;
    const uint32_t hi = lo + chunk < n ? lo + chunk : n;
    if (hi == lo) return 0;
    return crc_mulmod(crc_x_pow8n(n - hi), crc_slice(table, p + lo, hi - lo));
}

#ifdef __CUDACC__
// One warp per member, members handed out by an atomic counter (counter[0], zeroed by the caller). status[m] = 0 or bgz::Status.
// 56 registers (no spills; 54 used): measured 155 ms for a 2.7 GB file against 166 ms with __launch_bounds__(256, 4) and its 62
// registers (variants b24 / b25: 48, 52, 56, 60 registers all within 2 %; four CTAs of 8 warps are resident per SM either way).
#ifndef BGZW_MAXNREG
#define BGZW_MAXNREG 56
#endif
__global__ void __maxnreg__(BGZW_MAXNREG)
bgzf_inflate_warp_kernel(const uint8_t* __restrict__ comp, const Member* __restrict__ members, uint32_t n_members, uint8_t* __restrict__ out,
                         int32_t* __restrict__ status, uint32_t* __restrict__ counter) {
    extern __shared__ __align__(16) uint8_t bgzw_smem[];
    Tables& T = reinterpret_cast<Tables*>(bgzw_smem)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t m = 0;
        if (lane == 0) m = atomicAdd(counter, 1u);
        m = __shfl_sync(0xffffffffu, m, 0);
        if (m >= n_members) break;
        const Member mb = members[m];
        int st = bgz::OK;
        if (mb.out_len) st = inflate_member(comp + mb.in_off, mb.in_len, out + mb.out_off, mb.out_len, T);
        if (lane == 0) status[m] = st;
        __syncwarp();
    }
}

// One warp per member: CRC-32 of the inflated bytes against the member's footer (the 4 bytes behind its DEFLATE stream).
// Members whose status is already non-zero are skipped; a mismatch sets ERR_CRC.
__global__ void __launch_bounds__(256)
bgzf_crc_kernel(const uint8_t* __restrict__ comp, const Member* __restrict__ members, uint32_t n_members, const uint8_t* __restrict__ out,
                int32_t* __restrict__ status) {
    __shared__ uint32_t table[256];
    table[threadIdx.x] = crc_table_entry(threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t m = warp; m < n_members; m += nwarps) {
        if (status[m] != 0) continue;
        const Member mb = members[m];
        uint32_t part = crc_lane_part(table, out + mb.out_off, mb.out_len, lane);
#pragma unroll
        for (int d = 16; d; d >>= 1) part ^= __shfl_xor_sync(0xffffffffu, part, d);
        const uint8_t* f = comp + mb.in_off + mb.in_len;
        const uint32_t want = (uint32_t)f[0] | (uint32_t)f[1] << 8 | (uint32_t)f[2] << 16 | (uint32_t)f[3] << 24;
        if (lane == 0 && part != want) status[m] = ERR_CRC;
    }
}
#endif

}  // namespace bgzw
