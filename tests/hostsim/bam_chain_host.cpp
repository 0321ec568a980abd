// The segment form of the record-boundary search (csrc/bam_records.h: segment_guess / segment_consistent, the bodies of the
// kernels in csrc/bam_decode.cuh) on the host: for a BAM file and several segment sizes, either the segments are consistent
// and then their chains must be exactly the serial chain, or they are not and the file would go to the host decoder.
//   bam_chain_host file.bam  ->  one line per segment size: "seg=<bytes> consistent=<0|1> equal=<0|1> records=<n>"
// exit code 1 iff some size is consistent but not equal (the one thing that must never happen).
#include "../../breakdancer_b200/csrc/bam_records.h"

#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    gzFile g = gzopen(argv[1], "rb");
    if (!g) return 2;
    std::vector<uint8_t> raw;
    std::vector<uint8_t> buf(1 << 20);
    for (int k; (k = gzread(g, buf.data(), (unsigned)buf.size())) > 0;) raw.insert(raw.end(), buf.begin(), buf.begin() + k);
    gzclose(g);
    const uint64_t n = raw.size();
    if (n < 12) return 2;
    uint64_t o = 8 + brec::ld32(raw.data() + 4);
    const int32_t nref = brec::ldi32(raw.data() + o); o += 4;
    for (int32_t i = 0; i < nref; ++i) o += 4 + brec::ld32(raw.data() + o) + 4;
    const uint64_t first = o;
    std::vector<uint64_t> serial;
    while (o + 4 <= n) { const uint32_t bs = brec::ld32(raw.data() + o); if (bs < 32 || o + 4 + bs > n) return 3; serial.push_back(o); o += 4 + (uint64_t)bs; }
    int rc = 0;
    for (uint64_t segsize : {512ull, 4096ull, 65536ull, 1000003ull}) {
        const uint32_t nseg = (uint32_t)std::max<uint64_t>(1, (n - first + segsize - 1) / segsize);
        std::vector<uint64_t> cut(nseg + 1);
        for (uint32_t k = 0; k <= nseg; ++k) cut[k] = std::min<uint64_t>(n, first + k * segsize);
        std::vector<brec::Segment> seg(nseg);
        for (uint32_t k = 0; k < nseg; ++k) seg[k] = brec::segment_guess(raw.data(), n, cut[k], cut[k + 1], k == 0, nref);
        bool consistent = true;
        for (uint32_t k = 0; k < nseg; ++k) consistent &= brec::segment_consistent(seg.data(), k, nseg, n);
        bool equal = true;
        if (consistent) {
            std::vector<uint64_t> got;
            for (uint32_t k = 0; k < nseg; ++k) {
                uint64_t p = seg[k].guess; uint32_t c = 0;
                while (p + 4 <= n && p < cut[k + 1]) { got.push_back(p); p += 4 + (uint64_t)brec::ld32(raw.data() + p); ++c; }
                equal &= c == seg[k].count;
            }
            equal &= got == serial;
            if (!equal) rc = 1;
        }
        printf("seg=%llu consistent=%d equal=%d records=%zu\n", (unsigned long long)segsize, (int)consistent, (int)(consistent && equal), serial.size());
    }
    return rc;
}
