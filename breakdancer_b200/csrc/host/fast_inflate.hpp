// fast_inflate.hpp -- raw DEFLATE (RFC 1951) decoder for BGZF blocks: input and output are whole in memory and the output
// size is known, so the decoder can refill a 64-bit bit buffer with unaligned word loads, decode through two-level
// tables with an 11-bit first level (up to three literals per refill), and copy matches in words without bounds checks
// while both buffers have slack. Measured 1.3-1.4x zlib 1.3's inflate on BAM blocks (420 against 300 MB/s of output on the
// bundled chr21 BAMs, one core); inflate is what bounds the drop-in executable on real files (the reference spends ~47 % of
// its wall time in zlib: SURVEY.md section 8f-1). The caller (bam_io.cpp; BDK_FAST_INFLATE=0 turns it off) checks the BGZF CRC32 of every block and falls
// back to zlib if this decoder refuses a block or the checksum differs, so a defect here can cost time but never
// correctness; tests/hostsim/inflate_fuzz.cpp compares it with zlib on generated streams and feeds it corrupted ones under
// AddressSanitizer.
//
// Written from RFC 1951: stored / fixed / dynamic blocks, canonical Huffman codes (codes packed most-significant bit first
// into a least-significant-bit-first bit stream, hence the bit reversal when the tables are filled).
#pragma once
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

namespace bdh {
namespace finf {

constexpr int LL_BITS = 11, D_BITS = 8;              // first-level table widths
constexpr int LL_SYMS = 288, D_SYMS = 32, MAX_LEN = 15;
// table entry, 4 bytes: value (literal byte, length / distance base, or subtable offset), bits consumed by this lookup,
// kind and number of extra bits
struct Entry {
    uint16_t val;
    uint8_t nbits;      // bits of the code consumed by this lookup (first level: up to LL_BITS/D_BITS; subtable: the remainder)
    uint8_t kx;         // kind << 4 | extra.  kind: 0 literal, 1 length / distance, 2 end of block, 3 subtable link, 4 invalid
};
#define FINF_KIND(e) ((e).kx >> 4)
#define FINF_EXTRA(e) ((e).kx & 15)
inline Entry make_entry(uint16_t val, int nbits, int kind, int extra) { Entry e; e.val = val; e.nbits = (uint8_t)nbits; e.kx = (uint8_t)((kind << 4) | extra); return e; }

struct Tables {
    Entry ll[(1 << LL_BITS) + 2048];       // first level + subtables (worst case well below this)
    Entry d[(1 << D_BITS) + 1024];
};

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                                       8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

struct Rev8 { uint8_t t[256]; Rev8() { for (int i = 0; i < 256; ++i) { int r = 0; for (int b = 0; b < 8; ++b) r |= ((i >> b) & 1) << (7 - b); t[i] = (uint8_t)r; } } };
inline uint32_t reverse_bits(uint32_t code, int len) {          // len <= 15: reverse 16 bits through a byte table, drop the rest
    static const Rev8 R;
    return (((uint32_t)R.t[code & 0xff] << 8) | R.t[(code >> 8) & 0xff]) >> (16 - len);
}

// Build a two-level decoding table from code lengths. is_dist selects the symbol semantics. Returns false for an
// over-subscribed or (except for the single-code case RFC 1951 allows) incomplete code.
inline bool build_table(const uint8_t* lens, int nsym, bool is_dist, int first_bits, Entry* table, int table_cap) {
    int count[MAX_LEN + 1] = {0};
    for (int i = 0; i < nsym; ++i) count[lens[i]]++;
    count[0] = 0;
    int left = 1, used = 0;
    for (int l = 1; l <= MAX_LEN; ++l) { left = (left << 1) - count[l]; if (left < 0) return false; used += count[l]; }
    if (used == 0) {                                   // no codes at all (a block without distances): every lookup is invalid
        for (int i = 0; i < (1 << first_bits); ++i) table[i] = make_entry(0, 1, 4, 0);
        return true;
    }
    if (left > 0 && !(used == 1)) return false;        // incomplete code: only the one-code case is legal
    uint32_t next_code[MAX_LEN + 2];
    uint32_t code = 0;
    for (int l = 1; l <= MAX_LEN; ++l) { code = (code + count[l - 1]) << 1; next_code[l] = code; }
    const int first_size = 1 << first_bits;
    // a complete code covers every first-level index with an entry or a link; only an incomplete one leaves holes
    if (left > 0) for (int i = 0; i < first_size; ++i) table[i] = make_entry(0, 1, 4, 0);
    int max_len = MAX_LEN;
    while (max_len > 1 && !count[max_len]) --max_len;
    // longest code behind every first-level prefix that needs a subtable
    int sub_len[1 << LL_BITS];
    if (max_len > first_bits) {
        memset(sub_len, 0, sizeof(int) * first_size);
        uint32_t nc[MAX_LEN + 2];
        memcpy(nc, next_code, sizeof nc);
        for (int s = 0; s < nsym; ++s) {
            const int l = lens[s];
            if (l <= first_bits) { if (l) nc[l]++; continue; }
            const uint32_t rev = reverse_bits(nc[l]++, l);
            const int prefix = rev & (first_size - 1);
            if (l - first_bits > sub_len[prefix]) sub_len[prefix] = l - first_bits;
        }
    }
    int next_free = first_size;
    for (int p = 0; max_len > first_bits && p < first_size; ++p) {
        if (!sub_len[p]) continue;
        const int w = sub_len[p];
        if (next_free + (1 << w) > table_cap) return false;
        table[p] = make_entry((uint16_t)next_free, first_bits, 3, w);
        for (int i = 0; i < (1 << w); ++i) table[next_free + i] = make_entry(0, 1, 4, 0);
        next_free += 1 << w;
    }
    for (int s = 0; s < nsym; ++s) {
        const int l = lens[s];
        if (!l) continue;
        int kind, extra = 0; uint16_t val = 0;
        if (is_dist) {
            if (s >= 30) kind = 4;
            else { kind = 1; val = kDistBase[s]; extra = kDistExtra[s]; }
        } else if (s < 256) { kind = 0; val = (uint16_t)s; }
        else if (s == 256) kind = 2;
        else if (s <= 285) { kind = 1; val = kLenBase[s - 257]; extra = kLenExtra[s - 257]; }
        else kind = 4;
        const uint32_t rev = reverse_bits(next_code[l]++, l);
        if (l <= first_bits) {
            const Entry e = make_entry(val, l, kind, extra);
            for (uint32_t i = rev; i < (uint32_t)first_size; i += 1u << l) table[i] = e;
        } else {
            const int prefix = rev & (first_size - 1);
            const Entry link = table[prefix];
            const int w = FINF_EXTRA(link);
            const Entry e = make_entry(val, l - first_bits, kind, extra);
            for (uint32_t i = rev >> first_bits; i < (1u << w); i += 1u << (l - first_bits)) table[link.val + i] = e;
        }
    }
    return true;
}

struct BitReader {
    const uint8_t* in;
    const uint8_t* end;
    uint64_t buf = 0;
    int cnt = 0;            // valid bits in buf
    // refill to at least 56 bits while 8 input bytes are readable; byte-wise near the end (missing bytes read as zero)
    inline void refill() {
        if (end - in >= 8) {
            uint64_t w;
            memcpy(&w, in, 8);
            buf |= w << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56 && in < end) { buf |= (uint64_t)*in++ << cnt; cnt += 8; }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    inline void drop(int n) { buf >>= n; cnt -= n; }
};

static const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// Decodes exactly out_len bytes from in[0 .. in_len). Returns true iff the stream is well formed, ends with its final block
// and produces exactly out_len bytes.
inline bool inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, Tables& T) {
    BitReader br;
    br.in = in; br.end = in + in_len;
    uint8_t* op = out;
    uint8_t* const oend = out + out_len;
    bool last = false;
    while (!last) {
        br.refill();
        if (br.cnt < 3) return false;
        last = br.peek(1); br.drop(1);
        const uint32_t type = br.peek(2); br.drop(2);
        if (type == 0) {                                  // stored
            br.drop(br.cnt & 7);                          // to the byte boundary
            br.refill();
            if (br.cnt < 32) return false;
            const uint32_t len = br.peek(16); br.drop(16);
            const uint32_t nlen = br.peek(16); br.drop(16);
            if ((len ^ 0xffffu) != nlen) return false;
            // bytes still in the bit buffer belong to the stored data: step the input pointer back
            const uint8_t* src = br.in - (br.cnt >> 3);
            br.buf = 0; br.cnt = 0;
            if ((size_t)(br.end - src) < len || (size_t)(oend - op) < len) return false;
            memcpy(op, src, len);
            op += len; br.in = src + len;
            continue;
        }
        if (type == 3) return false;
        if (type == 1) {                                  // fixed codes
            uint8_t lens[LL_SYMS];
            for (int i = 0; i < 144; ++i) lens[i] = 8;
            for (int i = 144; i < 256; ++i) lens[i] = 9;
            for (int i = 256; i < 280; ++i) lens[i] = 7;
            for (int i = 280; i < 288; ++i) lens[i] = 8;
            uint8_t dl[D_SYMS];
            for (int i = 0; i < 32; ++i) dl[i] = 5;
            if (!build_table(lens, 288, false, LL_BITS, T.ll, (int)(sizeof T.ll / sizeof(Entry)))) return false;
            if (!build_table(dl, 32, true, D_BITS, T.d, (int)(sizeof T.d / sizeof(Entry)))) return false;
        } else {                                          // dynamic codes
            br.refill();
            if (br.cnt < 14) return false;
            const int hlit = (int)br.peek(5) + 257; br.drop(5);
            const int hdist = (int)br.peek(5) + 1; br.drop(5);
            const int hclen = (int)br.peek(4) + 4; br.drop(4);
            if (hlit > 286 || hdist > 30) return false;
            uint8_t cl[19] = {0};
            for (int i = 0; i < hclen; ++i) {
                br.refill();
                if (br.cnt < 3) return false;
                cl[kClOrder[i]] = (uint8_t)br.peek(3); br.drop(3);
            }
            Entry clt[1 << 7];
            if (!build_table(cl, 19, false, 7, clt, 1 << 7)) return false;     // code-length codes are at most 7 bits: no subtables
            uint8_t lens[LL_SYMS + D_SYMS];
            int n = 0;
            while (n < hlit + hdist) {
                br.refill();
                const Entry e = clt[br.peek(7)];
                if (FINF_KIND(e) == 4 || e.nbits > br.cnt) return false;
                br.drop(e.nbits);
                // build_table stored the code-length symbols 0..18 as "literals" (kind 0, val = symbol)
                const int sym = e.val;
                if (FINF_KIND(e) != 0) return false;
                if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
                int rep, val = 0;
                if (sym == 16) { if (n == 0) return false; val = lens[n - 1]; rep = 3 + (int)br.peek(2); br.drop(2); }
                else if (sym == 17) { rep = 3 + (int)br.peek(3); br.drop(3); }
                else { rep = 11 + (int)br.peek(7); br.drop(7); }
                if (br.cnt < 0 || n + rep > hlit + hdist) return false;
                while (rep--) lens[n++] = (uint8_t)val;
            }
            if (lens[256] == 0) return false;             // no end-of-block code
            uint8_t ll[LL_SYMS] = {0}, dl[D_SYMS] = {0};
            memcpy(ll, lens, hlit);
            memcpy(dl, lens + hlit, hdist);
            if (!build_table(ll, 288, false, LL_BITS, T.ll, (int)(sizeof T.ll / sizeof(Entry)))) return false;
            if (!build_table(dl, 32, true, D_BITS, T.d, (int)(sizeof T.d / sizeof(Entry)))) return false;
        }
        // ---- the symbols of this block ----------------------------------------------------------------------
        // Fast loop while both buffers have slack: one refill (>= 56 bits) covers two literals, or the rest of a match after a
        // second refill; nothing is bounds-checked inside (a match may write up to 15 bytes past its end, into the slack).
        bool block_done = false;
        while (br.end - br.in >= 16 && oend - op >= 258 + 16) {
            br.refill();
            Entry e = T.ll[br.peek(LL_BITS)];
            if (FINF_KIND(e) == 0) {                            // literal, and very often another one
                br.drop(e.nbits);
                *op++ = (uint8_t)e.val;
                e = T.ll[br.peek(LL_BITS)];
                if (FINF_KIND(e) == 0) {
                    br.drop(e.nbits);
                    *op++ = (uint8_t)e.val;
                    e = T.ll[br.peek(LL_BITS)];
                    if (FINF_KIND(e) == 0) { br.drop(e.nbits); *op++ = (uint8_t)e.val; continue; }      // 3 x 11 bits at most so far
                }
                br.refill();
            }
            if (FINF_KIND(e) == 3) { br.drop(LL_BITS); e = T.ll[e.val + br.peek(FINF_EXTRA(e))]; }
            br.drop(e.nbits);
            if (FINF_KIND(e) == 0) { *op++ = (uint8_t)e.val; continue; }
            if (FINF_KIND(e) == 2) { block_done = true; break; }
            if (FINF_KIND(e) != 1) return false;
            const uint32_t len = e.val + br.peek(FINF_EXTRA(e));
            br.drop(FINF_EXTRA(e));
            Entry de = T.d[br.peek(D_BITS)];
            if (FINF_KIND(de) == 3) { br.drop(D_BITS); de = T.d[de.val + br.peek(FINF_EXTRA(de))]; }
            if (FINF_KIND(de) != 1) return false;
            br.drop(de.nbits);
            const uint32_t dist = de.val + br.peek(FINF_EXTRA(de));
            br.drop(FINF_EXTRA(de));
            if (dist > (size_t)(op - out)) return false;
            const uint8_t* src = op - dist;
            uint8_t* dst = op;
            op += len;
            if (dist >= 16) {                                   // most matches: at most two 16-byte moves, no loop
                _mm_storeu_si128((__m128i*)dst, _mm_loadu_si128((const __m128i*)src));
                if (len > 16) {
                    _mm_storeu_si128((__m128i*)(dst + 16), _mm_loadu_si128((const __m128i*)(src + 16)));
                    if (len > 32) {
                        src += 32; dst += 32;
                        do { _mm_storeu_si128((__m128i*)dst, _mm_loadu_si128((const __m128i*)src)); src += 16; dst += 16; } while (dst < op);
                    }
                }
            } else if (dist >= 8) {
                do { uint64_t w0, w1; memcpy(&w0, src, 8); memcpy(dst, &w0, 8); memcpy(&w1, src + 8, 8); memcpy(dst + 8, &w1, 8); src += 16; dst += 16; } while (dst < op);
            } else if (dist == 1) {
                uint64_t w = 0x0101010101010101ull * src[0];
                do { memcpy(dst, &w, 8); dst += 8; } while (dst < op);
            } else {
                do { *dst++ = *src++; } while (dst < op);
            }
        }
        // careful loop: near the end of either buffer
        while (!block_done) {
            br.refill();
            Entry e = T.ll[br.peek(LL_BITS)];
            if (FINF_KIND(e) == 3) { br.drop(LL_BITS); e = T.ll[e.val + br.peek(FINF_EXTRA(e))]; }
            if (e.nbits > br.cnt) return false;
            br.drop(e.nbits);
            if (FINF_KIND(e) == 0) {
                if (op >= oend) return false;
                *op++ = (uint8_t)e.val;
                continue;
            }
            if (FINF_KIND(e) == 2) break;
            if (FINF_KIND(e) != 1) return false;
            uint32_t len = e.val + br.peek(FINF_EXTRA(e));
            br.drop(FINF_EXTRA(e));
            Entry de = T.d[br.peek(D_BITS)];
            if (FINF_KIND(de) == 3) { br.drop(D_BITS); de = T.d[de.val + br.peek(FINF_EXTRA(de))]; }
            if (FINF_KIND(de) != 1 || de.nbits > br.cnt) return false;
            br.drop(de.nbits);
            if (br.cnt < FINF_EXTRA(de)) { br.refill(); if (br.cnt < FINF_EXTRA(de)) return false; }
            const uint32_t dist = de.val + br.peek(FINF_EXTRA(de));
            br.drop(FINF_EXTRA(de));
            if (br.cnt < 0) return false;
            if (dist > (size_t)(op - out) || len > (size_t)(oend - op)) return false;
            const uint8_t* src = op - dist;
            for (uint32_t i = 0; i < len; ++i) op[i] = src[i];
            op += len;
        }
    }
    return op == oend;
}

// ---- CRC-32 of a decoded block ------------------------------------------------------------------------------------------
// CRC-32 (IEEE 802.3, the zlib polynomial, reflected) by carry-less multiplication: four 128-bit lanes folded by 512 bits per
// step, then reduced 512 -> 128 -> 64 -> 32 bits (Barrett). Needs len >= 64 and len % 16 == 0; the caller handles the rest.
__attribute__((target("pclmul,sse4.1")))
inline uint32_t crc32_clmul(const uint8_t* buf, size_t len, uint32_t crc) {
    static const uint64_t __attribute__((aligned(16))) k1k2[] = {0x0154442bd4ull, 0x01c6e41596ull};
    static const uint64_t __attribute__((aligned(16))) k3k4[] = {0x01751997d0ull, 0x00ccaa009eull};
    static const uint64_t __attribute__((aligned(16))) k5k0[] = {0x0163cd6124ull, 0x0000000000ull};
    static const uint64_t __attribute__((aligned(16))) poly[] = {0x01db710641ull, 0x01f7011641ull};
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
    x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
    x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = _mm_load_si128((const __m128i*)k1k2);
    buf += 64; len -= 64;
    while (len >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00)); y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
        y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20)); y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5); x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7); x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64; len -= 64;
    }
    x0 = _mm_load_si128((const __m128i*)k3k4);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128((const __m128i*)buf);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16; len -= 16;
    }
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64((const __m128i*)k5k0);
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_load_si128((const __m128i*)poly);
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}


// zlib's crc32(0, p, n): carry-less multiplication where the CPU has it (5.8 against 2.7 GB/s here), zlib for the tail / otherwise
inline uint32_t crc32_block(const uint8_t* p, size_t n, uint32_t (*zlib_crc)(uint32_t, const uint8_t*, size_t)) {
    uint32_t c = 0xffffffffu;
    static const bool have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    if (n >= 64 && have) {
        const size_t m = n & ~(size_t)15;
        c = crc32_clmul(p, m, c);
        p += m; n -= m;
    }
    c = ~c;
    return n ? zlib_crc(c, p, n) : c;
}

}  // namespace finf
}  // namespace bdh
