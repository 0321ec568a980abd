// The repair rounds of chain_resolve_kernel (breakdancer_b200/csrc/bam_decode.cuh) on the host: the same per-segment functions
// (bam_records.h: segment_guess, segment_in_place, segment_after) and the same rule -- a segment out of place is re-entered from
// its predecessor's end once the predecessor is in place -- run as synchronous rounds over the inflated record bytes of a BAM
// (argv[1]: file of raw record bytes, argv[2]: reference count, argv[3]: bytes cut off the end, then segment indices whose
// guesses are spoiled). Must reach the serial chain: same record count, same tail, same verdict. Prints the rounds it took.
#include "../../breakdancer_b200/csrc/bam_records.h"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <vector>
using brec::Segment;
int main(int argc, char** argv) {
    if (argc < 4) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<uint8_t> raw((std::istreambuf_iterator<char>(f)), {});
    const int nref = atoi(argv[2]);
    const uint64_t n = raw.size() - (uint64_t)atoll(argv[3]), SEG = 8192;
    const uint32_t nseg = (uint32_t)((n + SEG - 1) / SEG);
    auto hi = [&](uint32_t k) { const uint64_t h = (uint64_t)(k + 1) * SEG; return h < n ? h : n; };
    std::vector<Segment> seg(nseg);
    for (uint32_t k = 0; k < nseg; ++k) seg[k] = brec::segment_guess(raw.data(), n, (uint64_t)k * SEG, hi(k), k == 0, nref);
    for (int i = 4; i < argc; ++i) {                                  // a guess a few bytes off whose chain "ends" far away
        const uint32_t k = (uint32_t)atoi(argv[i]);
        if (k == 0 || k >= nseg) continue;
        seg[k].guess += 7; seg[k].end = seg[k].guess + 123456789ull; seg[k].bad = i & 1;
    }
    int rounds = 0;
    for (; rounds < 1000; ++rounds) {
        const std::vector<Segment> old = seg;
        int changed = 0;
        for (uint32_t k = 1; k < nseg; ++k) {
            if (brec::segment_in_place(old[k], old[k - 1])) continue;
            ++changed;
            if (k > 1 && !brec::segment_in_place(old[k - 1], old[k - 2])) continue;
            seg[k] = brec::segment_after(raw.data(), n, hi(k), old[k - 1]);
        }
        if (!changed) break;
    }
    uint64_t total = 0;
    for (auto& s : seg) total += s.count;
    uint64_t o = 0, cnt = 0;
    uint32_t bad = 0;
    while (o + 4 <= n) {
        const uint32_t bs = brec::ld32(raw.data() + o);
        if (bs < 32) { bad = 1; break; }
        if (o + 4 + (uint64_t)bs > n) { bad = 2; break; }
        ++cnt; o += 4 + bs;
    }
    const bool ok = nseg && total == cnt && seg[nseg - 1].end == o && seg[nseg - 1].bad == bad;
    printf("segments=%u rounds=%d records=%llu tail=%llu verdict=%u serial_records=%llu serial_tail=%llu serial_verdict=%u %s\n", nseg, rounds,
           (unsigned long long)total, (unsigned long long)seg[nseg - 1].end, seg[nseg - 1].bad, (unsigned long long)cnt, (unsigned long long)o, bad, ok ? "OK" : "MISMATCH");
    return ok ? 0 : 1;
}
