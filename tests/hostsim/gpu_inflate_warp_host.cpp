// Differential + corruption fuzz of the warp-per-member GPU decoder (breakdancer_b200/csrc/bgzf_inflate_warp.cuh,
// `bgzw::inflate_member`, what every warp of bgzf_inflate_warp_kernel runs) compiled for the host, against zlib; and of the
// sliced CRC-32 of bgzf_crc_kernel against zlib's crc32. On the host the 32 lanes of a phase run one after the other
// (BGZW_PHASE_BEGIN), so a lane that depended on another lane's work inside a phase would show as a mismatch. Built with
// -fsanitize=address,undefined by tests/test_zz_gpu_inflate.py: every well-formed stream must decode to zlib's bytes, every
// corrupted stream must be refused or decoded without touching memory outside the buffers.
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include "../../breakdancer_b200/csrc/bgzf_inflate_warp.cuh"

#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <vector>

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t>& src, int level, int strategy) {
    z_stream zs{};
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
    std::vector<uint8_t> out(deflateBound(&zs, src.size()) + 64);
    zs.next_in = (Bytef*)src.data(); zs.avail_in = (uInt)src.size();
    zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
    deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}

int main(int argc, char** argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 300;
    std::mt19937_64 rng(4321);
    std::unique_ptr<bgzw::Tables> T(new bgzw::Tables);
#ifdef BGZW_WINDOW_READER_ON_HOST
    // the device's word-window reader, on the host: the member sits at every alignment inside a buffer that has what the device
    // buffer has around it (3 bytes in front, the 8-byte footer and a word behind); the output has the 512 bytes of slack the
    // device buffer has, and must not be touched beyond them
    int align_round = 0;
    auto run = [&](const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
        const size_t mis = (size_t)(align_round++ & 3);
        std::vector<uint32_t> words((in_len + 3 + 12 + 3) / 4 + 1, 0xA5A5A5A5u);
        uint8_t* base = (uint8_t*)words.data() + mis;
        memcpy(base, in, in_len);
        std::vector<uint8_t> o(out_len + 512 + 1, 0xAB);
        const int st = bgzw::inflate_member(base, (uint32_t)in_len, o.data(), (uint32_t)out_len, *T);
        if (o[out_len + 512] != 0xAB) { fprintf(stderr, "WROTE PAST THE SLACK\n"); exit(2); }
        memcpy(out, o.data(), out_len);
        return st == bgz::OK;
    };
#else
    auto run = [&](const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
        std::vector<uint8_t> exact(in, in + in_len);          // exact-size copy: AddressSanitizer sees any read past the member
        return bgzw::inflate_member(exact.data(), (uint32_t)in_len, out, (uint32_t)out_len, *T) == bgz::OK;
    };
#endif
    uint32_t table[256];
    for (uint32_t k = 0; k < 256; ++k) table[k] = bgzw::crc_table_entry(k);
    long ok = 0, refused_good = 0, mismatched = 0, corrupt_accepted_wrong = 0, corrupt_cases = 0, crc_bad = 0;
    for (int r = 0; r < rounds; ++r) {
        const size_t n = r < 8 ? (size_t)r : r < 40 ? (size_t)(rng() % 700 + 1) : (size_t)(rng() % 65536 + 1);
        std::vector<uint8_t> src(n);
        const int kind = r % 7;
        for (size_t i = 0; i < n; ++i) {
            switch (kind) {
                case 0: src[i] = (uint8_t)rng(); break;                                  // incompressible
                case 1: src[i] = (uint8_t)("ACGTN"[rng() % 5]); break;                   // small alphabet
                case 2: src[i] = (uint8_t)(i % 7 == 0 ? rng() : 0); break;               // long zero runs (dist 1 matches)
                case 3: src[i] = (uint8_t)(i >= 3 ? src[i - 3] ^ (rng() % 16 == 0) : rng()); break;   // period-3 overlaps
                case 4: src[i] = (uint8_t)((i / 36) * 131 + (i % 36 < 32 ? (i % 36) * 7 : rng())); break;   // record-like
                case 5: src[i] = (uint8_t)(i >= 40 ? src[i - 40] ^ (rng() % 64 == 0) : rng()); break; // period-40 overlaps (len > dist > 32)
                default: src[i] = (uint8_t)(rng() % 3 ? 'A' + rng() % 4 : rng()); break;
            }
        }
        // the sliced CRC
        uint32_t crc = 0;
        for (int lane = 0; lane < 32; ++lane) crc ^= bgzw::crc_lane_part(table, src.data(), (uint32_t)n, lane);
        if (crc != (uint32_t)crc32(crc32(0L, Z_NULL, 0), src.data(), (uInt)n)) { ++crc_bad; fprintf(stderr, "CRC MISMATCH n=%zu\n", n); }
        const int level = (int)(rng() % 10);
        const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE, Z_FILTERED};
        const int strategy = strategies[rng() % 5];
        std::vector<uint8_t> comp = deflate_raw(src, level, strategy);
        std::vector<uint8_t> out(n + 1, 0xAB);   // (the host build checks the output bound at every symbol)
        if (!run(comp.data(), comp.size(), out.data(), n)) { ++refused_good; fprintf(stderr, "refused a good stream: n=%zu level=%d strategy=%d\n", n, level, strategy); continue; }
        if ((n && memcmp(out.data(), src.data(), n) != 0) || out[n] != 0xAB) { ++mismatched; fprintf(stderr, "MISMATCH n=%zu level=%d strategy=%d\n", n, level, strategy); continue; }
        ++ok;
        for (int c = 0; c < 6 && !comp.empty(); ++c) {
            std::vector<uint8_t> bad(comp);
            size_t want = n;
            if (c < 3) bad[rng() % bad.size()] ^= (uint8_t)(1u << (rng() % 8));
            else if (c == 3) bad.resize(rng() % bad.size());
            else if (c == 4) want = n ? n - 1 : 0;
            else want = n + 1;
            std::vector<uint8_t> o2(want + 1, 0xCD);
            ++corrupt_cases;
            const bool acc = run(bad.data(), bad.size(), o2.data(), want);
            if (o2[want] != 0xCD) { fprintf(stderr, "WROTE PAST THE OUTPUT\n"); return 2; }
            if (acc && (want != n || (n && memcmp(o2.data(), src.data(), n) != 0))) ++corrupt_accepted_wrong;   // legal (the CRC kernel catches it), just counted
        }
    }
    printf("ok=%ld refused_good=%ld mismatched=%ld crc_bad=%ld corrupt_cases=%ld corrupt_accepted_with_other_bytes=%ld\n", ok, refused_good, mismatched, crc_bad, corrupt_cases, corrupt_accepted_wrong);
    return (refused_good || mismatched || crc_bad) ? 1 : 0;
}
