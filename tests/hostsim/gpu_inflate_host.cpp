// Differential + corruption fuzz of the GPU member decoder (breakdancer_b200/csrc/bgzf_inflate.cuh, `inflate_member`, the
// function every thread of bgzf_inflate_kernel runs) compiled for the host, against zlib. Built with
// -fsanitize=address,undefined by tests/test_gpu_inflate.py: every well-formed stream must decode to the same bytes as zlib's,
// every corrupted stream must be refused or decoded without touching memory outside the buffers. The lookup tables are
// addressed with the kernel's stride (entry i of "thread" t at lut[i * stride + t]) to cover the interleaved layout.
#include "../../breakdancer_b200/csrc/bgzf_inflate.cuh"
#include <cstring>

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
    std::mt19937_64 rng(12345);
    const int stride = 5, lane = 3;
    std::vector<uint16_t> lut((size_t)bgz::LUT_PER_THREAD * stride, 0xFFFF);
    std::unique_ptr<bgz::Scratch> sc(new bgz::Scratch);
    auto run = [&](const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
        // exact-size copy of the input so that AddressSanitizer sees any read past the member
        std::vector<uint8_t> exact(in, in + in_len);
        return bgz::inflate_member(exact.data(), (uint32_t)in_len, out, (uint32_t)out_len, *sc, lut.data() + lane, stride) == bgz::OK;
    };
    long ok = 0, refused_good = 0, mismatched = 0, corrupt_accepted_wrong = 0, corrupt_cases = 0;
    for (int r = 0; r < rounds; ++r) {
        const size_t n = r < 8 ? (size_t)r : (size_t)(rng() % 65536 + 1);
        std::vector<uint8_t> src(n);
        const int kind = r % 6;
        for (size_t i = 0; i < n; ++i) {
            switch (kind) {
                case 0: src[i] = (uint8_t)rng(); break;                                  // incompressible
                case 1: src[i] = (uint8_t)("ACGTN"[rng() % 5]); break;                   // small alphabet
                case 2: src[i] = (uint8_t)(i % 7 == 0 ? rng() : 0); break;               // long zero runs (dist 1 matches)
                case 3: src[i] = (uint8_t)(i >= 3 ? src[i - 3] ^ (rng() % 16 == 0) : rng()); break;   // period-3 overlaps
                case 4: src[i] = (uint8_t)((i / 36) * 131 + (i % 36 < 32 ? (i % 36) * 7 : rng())); break;   // record-like
                default: src[i] = (uint8_t)(rng() % 3 ? 'A' + rng() % 4 : rng()); break;
            }
        }
        const int level = (int)(rng() % 10);
        const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE, Z_FILTERED};
        const int strategy = strategies[rng() % 5];
        std::vector<uint8_t> comp = deflate_raw(src, level, strategy);
        std::vector<uint8_t> out(n + 1, 0xAB);
        if (!run(comp.data(), comp.size(), out.data(), n)) { ++refused_good; fprintf(stderr, "refused a good stream: n=%zu level=%d strategy=%d\n", n, level, strategy); continue; }
        if ((n && memcmp(out.data(), src.data(), n) != 0) || out[n] != 0xAB) { ++mismatched; fprintf(stderr, "MISMATCH n=%zu level=%d strategy=%d\n", n, level, strategy); continue; }
        ++ok;
        // corruptions: flipped bits, truncation, wrong output length
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
            if (acc && (want != n || (n && memcmp(o2.data(), src.data(), n) != 0))) ++corrupt_accepted_wrong;   // legal (the caller checks the CRC), just counted
        }
    }
    for (size_t i = 0; i < lut.size(); ++i) if ((int)(i % stride) != lane && lut[i] != 0xFFFF) { fprintf(stderr, "WROTE ANOTHER THREAD'S TABLE ENTRY\n"); return 2; }
    printf("ok=%ld refused_good=%ld mismatched=%ld corrupt_cases=%ld corrupt_accepted_with_other_bytes=%ld\n", ok, refused_good, mismatched, corrupt_cases, corrupt_accepted_wrong);
    return (refused_good || mismatched) ? 1 : 0;
}
