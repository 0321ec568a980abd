// Single-core throughput of the BGZF block decoder (csrc/host/fast_inflate.hpp) against zlib on the blocks of a BAM file.
//   g++ -O3 -std=c++17 scripts/inflate_bench.cpp -o build/inflate_bench -lz && build/inflate_bench file.bam [reps]
// Every block's output is compared with zlib's. CPU only; not part of the product.
#include "../breakdancer_b200/csrc/host/fast_inflate.hpp"
#include <zlib.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
struct Blk { size_t in_off; uint32_t in_len, out_len; size_t out_off; };
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const int reps = argc > 2 ? atoi(argv[2]) : 3;
    FILE* f = fopen(argv[1], "rb"); if (!f) return 2;
    fseek(f, 0, SEEK_END); size_t n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> d(n + 64); if (fread(d.data(), 1, n, f) != n) return 2; fclose(f);
    std::vector<Blk> blks; size_t off = 0, total = 0;
    while (off + 18 <= n) {
        const uint8_t* h = d.data() + off;
        uint32_t xlen = h[10] | (h[11] << 8), bsize = (h[16] | (h[17] << 8)) + 1;       // BC is the first extra field in files written by samtools / us
        Blk b; b.in_off = off + 12 + xlen; b.in_len = bsize - 12 - xlen - 8;
        memcpy(&b.out_len, d.data() + off + bsize - 4, 4); b.out_off = total; total += b.out_len;
        blks.push_back(b); off += bsize;
    }
    std::vector<uint8_t> a(total + 64), b2(total + 64);
    z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, -15);
    double tz = 1e9, tf = 1e9;
    for (int r = 0; r < reps; ++r) {
        double t0 = now();
        for (auto& b : blks) { if (!b.out_len) continue; inflateReset(&zs); zs.next_in = d.data() + b.in_off; zs.avail_in = b.in_len; zs.next_out = a.data() + b.out_off; zs.avail_out = b.out_len; if (inflate(&zs, Z_FINISH) != Z_STREAM_END) { printf("zlib failed\n"); return 1; } }
        double t1 = now(); if (t1 - t0 < tz) tz = t1 - t0;
    }
    auto T = new bdh::finf::Tables; size_t refused = 0;
    for (int r = 0; r < reps; ++r) {
        refused = 0;
        double t0 = now();
        for (auto& b : blks) { if (!b.out_len) continue; if (!bdh::finf::inflate_raw(d.data() + b.in_off, b.in_len, b2.data() + b.out_off, b.out_len, *T)) ++refused; }
        double t1 = now(); if (t1 - t0 < tf) tf = t1 - t0;
    }
    const bool same = memcmp(a.data(), b2.data(), total) == 0;
    printf("%s: %zu blocks, %.1f MB -> %.1f MB; zlib %.3f s (%.0f MB/s), fast %.3f s (%.0f MB/s), ratio %.2f, refused %zu, %s\n", argv[1], blks.size(), n / 1e6, total / 1e6,
           tz, total / 1e6 / tz, tf, total / 1e6 / tf, tz / tf, refused, same ? "identical" : "DIFFERENT");
    return same && !refused ? 0 : 1;
}
