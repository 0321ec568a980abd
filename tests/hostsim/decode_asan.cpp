// The host BAM decoder (csrc/host/bam_io.cpp and what it includes) under AddressSanitizer / UBSan: opens the config's bams
// the way the executable does and prints the record count and a checksum of the columns. Built by tests/test_decode_asan.py
// from the host sources with -fsanitize=address,undefined (the CUDA object is linked as it is; no CUDA call is made).
//   decode_asan <config> <region or ""> <threads> <keep_records 0|1>
#include "../../include/bdk.h"
#include "../../include/bdk_host.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    char err[512] = {0};
    bdh_config* cfg = bdh_config_load(argv[1], 3, err, sizeof err);
    if (!cfg) { printf("config error: %s\n", err); return 3; }
    bdh_stream* s = bdh_stream_open(cfg, nullptr, 0, argv[2], atoi(argv[3]), 0, atoi(argv[4]), err, sizeof err);
    if (!s) { printf("error: %s\n", err); bdh_config_free(cfg); return 0; }
    const uint64_t n = bdh_stream_n(s);
    bdk_soa c;
    bdh_stream_cols(s, &c);
    uint64_t h = 1469598103934665603ull;
    auto mixin = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
    for (uint64_t i = 0; i < n; ++i) {
        mixin((uint32_t)c.pos[i]); mixin((uint32_t)c.mpos[i]); mixin((uint32_t)c.tid[i]); mixin((uint32_t)c.mtid[i]); mixin((uint32_t)c.isize[i]);
        mixin(c.flag[i]); mixin(c.mapq[i]); mixin((uint32_t)c.qlen[i]); mixin(c.qid[i]);
    }
    if (atoi(argv[4]) && n) { char buf[1 << 16]; if (bdh_stream_fastq(s, n / 2, buf, sizeof buf) < 0) { printf("fastq failed\n"); return 4; } mixin(strlen(buf)); }
    printf("n=%llu checksum=%016llx\n", (unsigned long long)n, (unsigned long long)h);
    bdh_stream_free(s);
    bdh_config_free(cfg);
    return 0;
}
