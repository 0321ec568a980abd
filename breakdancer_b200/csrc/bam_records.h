// bam_records.h -- what is read out of one BAM record on the hot path, as host/device functions: the host decoder
// (csrc/host/bam_io.cpp) is built from them today, and the device-resident decode that follows the GPU inflate
// (record chain and field extraction on the inflated bytes in HBM) will run the same code, the
// way bdk_logic.h serves the kernels and the host simulation.  Reference: Alignment ctor / determine_bdqual /
// determine_read_group (src/lib/io/Alignment.cpp:12-64), the reader filter (src/lib/io/BamIo.cpp:11-18) and the region
// overlap test of samtools' bam_iter_read under RegionLimitedBamReader.hpp:40-66.  Written from the SAM/BAM specification:
// a record is block_size (u32) + 32-byte core + read name + cigar + packed sequence + qualities + aux fields.
// All loads tolerate any alignment (records start at arbitrary byte offsets of the inflated stream).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define BREC_HD __host__ __device__ __forceinline__
#else
#define BREC_HD inline
#endif

namespace brec {

// Unaligned little-endian loads. Device: the two aligned words around p and one funnel shift (records lie at any alignment in the
// inflated stream; four byte loads and three shifts per field made the extraction pass load-issue bound). This reads up to 7
// bytes beyond p + 4: the device buffers of bdk_bam.inl end in slack for that (d_raw: + 1024 bytes).
BREC_HD uint32_t ld32(const uint8_t* p) {
#ifdef __CUDA_ARCH__
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    return __funnelshift_r(w[0], w[1], (uint32_t)(a & 3) * 8);
#else
    uint32_t v; memcpy(&v, p, 4); return v;
#endif
}
BREC_HD uint32_t ld16(const uint8_t* p) {
#ifdef __CUDA_ARCH__
    return ld32(p) & 0xffffu;
#else
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8);
#endif
}
BREC_HD int32_t ldi32(const uint8_t* p) { return (int32_t)ld32(p); }
BREC_HD uint64_t ld64(const uint8_t* p) { return (uint64_t)ld32(p) | ((uint64_t)ld32(p + 4) << 32); }

// 32-byte BAM core (little-endian on disk)
struct Core {
    int32_t tid, pos; uint8_t l_qname, mapq; uint16_t bin; uint16_t n_cigar, flag; int32_t l_qseq, mtid, mpos, isize;
};
BREC_HD Core read_core(const uint8_t* r) {
    Core c;
    c.tid = ldi32(r); c.pos = ldi32(r + 4);
    c.l_qname = r[8]; c.mapq = r[9]; c.bin = (uint16_t)ld16(r + 10);
    c.n_cigar = (uint16_t)ld16(r + 12); c.flag = (uint16_t)ld16(r + 14);
    c.l_qseq = ldi32(r + 16); c.mtid = ldi32(r + 20); c.mpos = ldi32(r + 24); c.isize = ldi32(r + 28);
    return c;
}

// samtools bam_calend: reference span from the CIGAR (ops M,D,N,=,X consume the reference)
BREC_HD uint32_t calend(const Core& c, const uint8_t* cigar) {
    uint32_t end = (uint32_t)c.pos;
    for (int k = 0; k < c.n_cigar; ++k) {
        const uint32_t v = ld32(cigar + 4 * k), op = v & 0xf;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += v >> 4;
    }
    return end;
}

// Walk the aux area [s, e); pointer to the value-type byte of the tag, or null (bam_aux_get).
BREC_HD const uint8_t* aux_get(const uint8_t* s, const uint8_t* e, char t0, char t1) {
    while (s + 3 <= e) {
        const uint8_t* v = s + 2;
        if (s[0] == (uint8_t)t0 && s[1] == (uint8_t)t1) return v;
        const uint8_t t = *v++;
        switch (t) {
            case 'A': case 'c': case 'C': v += 1; break;
            case 's': case 'S': v += 2; break;
            case 'i': case 'I': case 'f': v += 4; break;
            case 'd': v += 8; break;
            case 'Z': case 'H': while (v < e && *v) ++v; ++v; break;
            case 'B': {
                if (v + 5 > e) return 0;
                const uint8_t st = *v; const uint32_t cnt = ld32(v + 1); v += 5;
                const size_t sz = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                v += sz * cnt; break;
            }
            default: return 0;
        }
        s = v;
    }
    return 0;
}

// does the integer value at v (type byte first) lie inside [v, e)?
BREC_HD bool aux_value_fits(const uint8_t* v, const uint8_t* e) {
    if (v >= e) return false;
    const size_t sz = (*v == 'c' || *v == 'C') ? 1 : (*v == 's' || *v == 'S') ? 2 : (*v == 'i' || *v == 'I') ? 4 : 0;
    return v + 1 + sz <= e;
}

BREC_HD int32_t aux2i(const uint8_t* v) {  // bam_aux2i
    switch (*v) {
        case 'c': return (int8_t)v[1];
        case 'C': return v[1];
        case 's': return (int16_t)ld16(v + 1);
        case 'S': return (int32_t)ld16(v + 1);
        case 'i': case 'I': return ldi32(v + 1);
        default: return 0;
    }
}

BREC_HD uint64_t mix64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return (a * b) ^ __umul64hi(a, b);
#else
    const __uint128_t r = (__uint128_t)a * b;
    return (uint64_t)r ^ (uint64_t)(r >> 64);
#endif
}

// 64-bit key of a byte string (read names: the mate join; read-group ids): wyhash-style multiply-mix
BREC_HD uint64_t hash_bytes(const uint8_t* s, size_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xA0761D6478BD642Full);
    while (n >= 8) { h = mix64(h ^ ld64(s), 0xE7037ED1A0B428DBull); s += 8; n -= 8; }
    uint64_t w = 0;
    for (size_t i = 0; i < n; ++i) w |= (uint64_t)s[i] << (8 * i);
    h = mix64(h ^ w ^ ((uint64_t)n << 56), 0x8EBC6AF09C88C6E3ull);
    return mix64(h, 0x589965CC75374CC3ull);
}

BREC_HD size_t bounded_strlen(const uint8_t* s, size_t cap) { size_t n = 0; while (n < cap && s[n]) ++n; return n; }

// Could a record start at offset o (its block_size field) of raw[0 .. n)? Only a guess (the search for record boundaries
// checks every guess against the chain that really arrives).
BREC_HD bool plausible(const uint8_t* raw, size_t n, size_t o, int32_t nref) {
    if (o + 36 > n) return false;
    const uint32_t bs = ld32(raw + o);
    if (bs < 32 || bs > (1u << 26) || o + 4 + bs > n) return false;
    const uint8_t* r = raw + o + 4;
    const int32_t tid = ldi32(r), pos = ldi32(r + 4), lq = ldi32(r + 16), mtid = ldi32(r + 20), mpos = ldi32(r + 24);
    const uint32_t l_qname = r[8], n_cigar = ld16(r + 12);
    if (tid < -1 || tid >= nref || mtid < -1 || mtid >= nref || pos < -1 || mpos < -1 || lq < 0 || l_qname == 0) return false;
    const uint64_t fixed = 32 + (uint64_t)l_qname + 4ull * n_cigar + ((uint64_t)lq + 1) / 2 + (uint64_t)lq;
    return fixed <= bs && r[32 + l_qname - 1] == 0;
}

// ---- record boundaries by segments (the device form of bam_io.cpp find_records) ----------------------------------------------
constexpr uint64_t NO_GUESS = ~0ull;
struct Segment {
    uint64_t guess;      // offset of the block_size field of the first record that starts in the segment (NO_GUESS: none found)
    uint64_t end;        // where the chain from `guess` leaves the segment (offset of the first record starting at or after its end)
    uint32_t count;      // records that start in the segment
    uint32_t bad;        // the chain ran into a record that cannot be: 1 = block_size < 32, 2 = it reaches past the end of the data
};

// Segment [lo, hi) of raw[0 .. n): guess its first record (the first segment starts on one: `first`), follow the chain.
BREC_HD Segment segment_guess(const uint8_t* raw, uint64_t n, uint64_t lo, uint64_t hi, bool first, int32_t nref) {
    Segment s;
    s.guess = NO_GUESS; s.end = hi; s.count = 0; s.bad = 0;
    uint64_t o = lo;
    if (!first) {
        for (; o < hi; ++o) {
            uint64_t q = o;
            int depth = 0;
            while (depth < 3 && plausible(raw, n, q, nref)) { q += 4 + (uint64_t)ld32(raw + q); ++depth; }
            if (depth == 3 || (depth > 0 && q + 4 > n)) break;
        }
        if (o >= hi) return s;
    }
    s.guess = o;
    while (o + 4 <= n && o < hi) {
        const uint32_t bs = ld32(raw + o);
        if (bs < 32 || o + 4 + bs > n) { s.bad = bs < 32 ? 1u : 2u; break; }
        ++s.count;
        o += 4 + (uint64_t)bs;
    }
    s.end = o;
    return s;
}

// Is segment k's part of the picture right? Every guess present, no broken record, every chain ending exactly on the next
// segment's guess, the last one with the data (fewer than 4 stray bytes are ignored, as on the host). If this holds for all k,
// the guessed chains ARE the serial chain (induction from the first segment, which starts on a true record); if it fails
// anywhere the file goes to the host decoder.
BREC_HD bool segment_consistent(const Segment* seg, uint32_t k, uint32_t nseg, uint64_t n) {
    if (seg[k].guess == NO_GUESS || seg[k].bad) return false;
    if (k + 1 < nseg) return seg[k + 1].guess == seg[k].end;
    return seg[k].end + 4 > n;
}

// ---- repair rounds (bam_decode.cuh chain_resolve_kernel; tests/hostsim/bam_chain_rounds.cpp runs the same two functions) ----
// Does segment `cur` stand where `prev` left the chain? A chain that met a broken or cut-off record ends there: the segments
// behind it are empty and carry its end and verdict on.
BREC_HD bool segment_in_place(const Segment& cur, const Segment& prev) {
    if (cur.guess != prev.end) return false;
    if (prev.bad) return cur.count == 0 && cur.end == prev.end && cur.bad == prev.bad;
    return true;
}
// The segment that ends at `hi`, entered where `prev` left the chain.
BREC_HD Segment segment_after(const uint8_t* raw, uint64_t n, uint64_t hi, const Segment& prev) {
    Segment s;
    s.guess = prev.end; s.end = prev.end; s.count = 0; s.bad = prev.bad;
    if (prev.bad) return s;
    uint64_t o = prev.end;
    while (o + 4 <= n && o < hi) {
        const uint32_t bs = ld32(raw + o);
        if (bs < 32) { s.bad = 1; break; }
        if (o + 4 + (uint64_t)bs > n) { s.bad = 2; break; }
        ++s.count;
        o += 4 + (uint64_t)bs;
    }
    s.end = o;
    return s;
}

struct RegionSel { int on, tid, beg, end; };

// The reader's filter: primary records placed on a reference sequence (BamIo.cpp:11-18) and, with -o, bam_iter_read's
// is_overlap(): rend > beg && pos < end on the target. r points at the record's core.
// do the variable-length parts the core announces fit inside the record (block_size precedes the core)?
BREC_HD bool record_sizes_ok(const uint8_t* r, const Core& c) {
    const uint64_t bs = ld32(r - 4);
    if (c.l_qseq < 0) return false;
    return 32ull + c.l_qname + 4ull * c.n_cigar + ((uint64_t)c.l_qseq + 1) / 2 + (uint64_t)c.l_qseq <= bs;
}

BREC_HD bool keep_record(const uint8_t* r, const RegionSel& region) {
    const Core c = read_core(r);
    if (!record_sizes_ok(r, c)) return false;          // a corrupt record that still chains: dropped, never read through
    bool ok = !(c.flag & (0x100 | 0x800)) && c.tid >= 0;
    if (ok && region.on) {
        const uint32_t rend = c.n_cigar ? calend(c, r + 32 + c.l_qname) : (uint32_t)c.pos + 1;
        ok = c.tid == region.tid && rend > (uint32_t)region.beg && c.pos < region.end;
    }
    return ok;
}

// The fields of a kept record (Alignment.cpp:12-64). mapq is the breakdancer quality: the AM tag where present, else MAPQ.
// The read group is returned as a byte range (empty without an RG:Z tag); the caller maps it to an id.
struct Fields {
    int32_t tid, pos, mtid, mpos, isize, qlen;
    uint16_t flag;
    uint8_t mapq;
    uint64_t qid;
    const uint8_t* rg;
    size_t rg_len;
};
BREC_HD Fields record_fields(const uint8_t* r) {
    const uint32_t bs = ld32(r - 4);
    const Core c = read_core(r);
    const uint8_t* name = r + 32;
    const uint8_t* aux = name + c.l_qname + 4 * (size_t)c.n_cigar + ((size_t)c.l_qseq + 1) / 2 + (size_t)c.l_qseq;
    const uint8_t* end = r + bs;
    Fields f;
    f.tid = c.tid; f.pos = c.pos; f.mtid = c.mtid; f.mpos = c.mpos; f.isize = c.isize; f.qlen = c.l_qseq; f.flag = c.flag;
    f.mapq = c.mapq;
    f.rg = 0; f.rg_len = 0;
    if (aux <= end) {
        if (const uint8_t* am = aux_get(aux, end, 'A', 'M')) if (aux_value_fits(am, end)) f.mapq = (uint8_t)aux2i(am);       // determine_bdqual
        if (const uint8_t* rg = aux_get(aux, end, 'R', 'G'))
            if (*rg == 'Z' || *rg == 'H') { f.rg = rg + 1; f.rg_len = bounded_strlen(rg + 1, (size_t)(end - (rg + 1))); }
    }
    f.qid = hash_bytes(name, c.l_qname ? bounded_strlen(name, c.l_qname) : 0);
    return f;
}

}  // namespace brec
