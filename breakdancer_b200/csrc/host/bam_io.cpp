// BAM -> struct-of-arrays decoder, k-way merge, and BAM writer.
//
// Host-side mirror of the reference's reader stack for the hot path:
//   openBam / openBams              src/lib/io/BamIo.cpp:6-31          (primary && tid >= 0 filter)
//   RegionLimitedBamReader          src/lib/io/RegionLimitedBamReader.hpp:36-71   (-o region)
//   BamMerger                       src/lib/io/BamMerger.cpp:40-126    ((tid,pos,strand) merge)
//   Alignment ctor / determine_*    src/lib/io/Alignment.cpp:12-64     (field extraction, AM, RG)
//   AlignmentSource::next           src/lib/io/AlignmentSource.hpp:48-65 (read group -> library)
// Written from the SAM/BAM specification (BGZF = concatenated gzip members with a "BC" extra
// field; BAM records = 32-byte core + name + cigar + seq + qual + aux). Unlike the reference
// (one zlib stream, one record and two heap objects at a time) a file is mapped, its BGZF
// blocks are inflated in parallel into one buffer, and fields are extracted by all cores
// straight into the columns the GPU consumes.
#include "host.hpp"
#include "nway_merge.hpp"
#include "fast_inflate.hpp"
#include "../bam_records.h"

#include <zlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <queue>
#include <stdexcept>
#include <thread>
#include <unordered_map>

#include <cuda_runtime_api.h>

namespace bdh {

int default_threads() {
    unsigned n = std::thread::hardware_concurrency();
    const char* e = getenv("BDK_THREADS");
    if (e && atoi(e) > 0) return atoi(e);
    return n ? (int)n : 1;
}

void parallel_for(uint64_t n, uint64_t grain, int threads, const std::function<void(uint64_t, uint64_t)>& fn) {
    if (n == 0) return;
    if (grain == 0) grain = 1;
    if (threads <= 1 || n <= grain) { fn(0, n); return; }
    std::atomic<uint64_t> next(0);
    auto worker = [&]() {
        for (;;) {
            uint64_t b = next.fetch_add(grain);
            if (b >= n) return;
            fn(b, std::min(n, b + grain));
        }
    };
    std::vector<std::thread> pool;
    int nt = (int)std::min<uint64_t>(threads, (n + grain - 1) / grain);
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
}

namespace {

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t rdi32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

struct MappedFile {
    const uint8_t* data = 0;
    size_t size = 0;
    int fd = -1;
    void open(const std::string& path) {
        fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("Failed to open samfile " + path);
        struct stat st;
        if (fstat(fd, &st) != 0) throw std::runtime_error("Failed to open samfile " + path);
        size = (size_t)st.st_size;
        if (size) {
            void* p = mmap(0, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) throw std::runtime_error("Failed to open samfile " + path);
            data = (const uint8_t*)p;
            madvise(p, size, MADV_SEQUENTIAL);
        }
    }
    ~MappedFile() {
        if (data) munmap((void*)data, size);
        if (fd >= 0) ::close(fd);
    }
};

struct Block { size_t in_off; uint32_t in_len; uint32_t out_len; size_t out_off; size_t file_off; };

// The inflated file: a plain allocation that is NOT zero-filled (std::vector::resize would memset more than a gigabyte on one
// thread before the workers start; here every page is first touched by the worker that inflates into it).
struct RawBuf {
    uint8_t* p = nullptr;
    size_t n = 0;
    RawBuf() {}
    RawBuf(const RawBuf&) = delete;
    RawBuf& operator=(const RawBuf&) = delete;
    RawBuf(RawBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    RawBuf& operator=(RawBuf&& o) noexcept { if (this != &o) { free(p); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~RawBuf() { free(p); }
    void resize(size_t bytes) {
        free(p); p = nullptr; n = 0;
        if (bytes) { p = (uint8_t*)malloc(bytes); if (!p) throw std::bad_alloc(); n = bytes; }
    }
    void release() { free(p); p = nullptr; n = 0; }
    uint8_t* data() { return p; }
    const uint8_t* data() const { return p; }
    size_t size() const { return n; }
};
std::atomic<uint64_t> g_gpu_inflate_redone(0);    // members the GPU decoder refused or got wrong (then decoded on the host)
std::atomic<uint64_t> g_inflate_fallbacks(0);     // blocks the fast decoder refused or got wrong (then decoded by zlib)

// BGZF members from file offset `off` on: up to the end of the file, or up to and including the member that starts at
// `last_member_off`. Output offsets continue from `total`.
void scan_members(const MappedFile& f, const std::string& path, size_t off, size_t last_member_off, std::vector<Block>& blocks, size_t& total) {
    while (off < f.size && off <= last_member_off) {
        if (off + 18 > f.size) throw std::runtime_error(path + " is not a valid bam file");
        const uint8_t* h = f.data + off;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) throw std::runtime_error(path + " is not a valid bam file");
        uint32_t xlen = rd16(h + 10);
        if (off + 12 + (size_t)xlen + 8 > f.size) throw std::runtime_error(path + " is truncated");     // extra field + deflate trailer must be in the file
        uint32_t bsize = 0; bool found = false;
        size_t x = 12;
        while (x + 4 <= 12 + xlen) {
            uint8_t si1 = h[x], si2 = h[x + 1]; uint32_t slen = rd16(h + x + 2);
            if (x + 4 + slen > 12 + (size_t)xlen) throw std::runtime_error(path + " is not a valid bam file");   // subfield runs past the extra field
            if (si1 == 'B' && si2 == 'C' && slen == 2) { bsize = rd16(h + x + 4); found = true; }
            x += 4 + slen;
        }
        if (!found) throw std::runtime_error(path + " is not a valid bam file");
        size_t blen = (size_t)bsize + 1;
        if (off + blen > f.size || blen < 12 + xlen + 8) throw std::runtime_error(path + " is truncated");
        Block b;
        b.file_off = off;
        b.in_off = off + 12 + xlen;
        b.in_len = (uint32_t)(blen - 12 - xlen - 8);
        b.out_len = rd32(f.data + off + blen - 4);
        b.out_off = total;
        total += b.out_len;
        blocks.push_back(b);
        off += blen;
    }
}

// Inflate `blocks` into `out` (already sized; the blocks' out_off are offsets into it) with `threads` workers.
void inflate_members(const MappedFile& f, const std::string& path, int threads, const std::vector<Block>& blocks, RawBuf& out) {
    size_t total = 0;
    for (auto const& b : blocks) total += b.out_len;
    std::atomic<bool> bad(false);
    // Every block goes through the table-driven decoder of fast_inflate.hpp first; its result is accepted only if the block's
    // CRC32 (BGZF footer, checked by carry-less multiplication) matches, otherwise zlib decodes the block. 1.4x zlib's inflate
    // on real BAM blocks, about even on very compressible files (the synthetic BAMs of the tests). BDK_FAST_INFLATE=0: zlib only.
    static const bool use_fast = !(getenv("BDK_FAST_INFLATE") && atoi(getenv("BDK_FAST_INFLATE")) == 0) && !getenv("BDK_ZLIB_ONLY");
    // BDK_GPU_INFLATE=1: all members are inflated by the GPU first (csrc/bgzf_inflate_warp.cuh through bdk_bgzf_inflate); the workers
    // below then only check the CRC32 of every member and re-inflate on the host whatever the device refused or got wrong.
    std::vector<int32_t> gpu_status;
    if (getenv("BDK_GPU_INFLATE") && atoi(getenv("BDK_GPU_INFLATE")) > 0 && !blocks.empty()) {
        std::vector<bdk_bgzf_member> mem(blocks.size());
        for (size_t i = 0; i < blocks.size(); ++i) { mem[i].in_off = blocks[i].in_off; mem[i].out_off = blocks[i].out_off; mem[i].in_len = blocks[i].in_len; mem[i].out_len = blocks[i].out_len; }
        gpu_status.assign(blocks.size(), -1);
        float kms = 0.f;
        const double t0 = now_s();
        const int dev = getenv("BDK_GPU_INFLATE_DEVICE") ? atoi(getenv("BDK_GPU_INFLATE_DEVICE")) : 0;
        const int rc = bdk_bgzf_inflate(dev, f.data, f.size, mem.data(), mem.size(), out.data(), out.size(), gpu_status.data(), &kms);
        if (rc != 0) throw std::runtime_error(path + ": BDK_GPU_INFLATE is set but the device inflate failed (error " + std::to_string(rc) + "); there is no silent fallback");
        if (getenv("BDK_DECODE_TRACE")) {
            size_t refused = 0;
            for (int32_t v : gpu_status) refused += v != 0;
            fprintf(stderr, "[decode] %s: %zu BGZF members, %.1f MB -> %.1f MB on the GPU: kernel %.3f ms (%.1f GB/s of output), with copies %.3f s, %zu members refused\n",
                    path.c_str(), blocks.size(), f.size / 1e6, total / 1e6, kms, kms > 0 ? total / 1e6 / kms : 0.0, now_s() - t0, refused);
        }
    }
    std::atomic<uint64_t> gpu_wrong(0);
    parallel_for(blocks.size(), 64, threads, [&](uint64_t b0, uint64_t b1) {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) { bad = true; return; }
        std::unique_ptr<finf::Tables> tables(use_fast ? new finf::Tables : nullptr);
        uint64_t fell_back = 0;
        for (uint64_t i = b0; i < b1; ++i) {
            Block const& b = blocks[i];
            if (b.out_len == 0) continue;
            uint8_t* dst = out.data() + b.out_off;
            if (!gpu_status.empty()) {
                if (gpu_status[i] == 0) {
                    const uint32_t want = rd32(f.data + b.in_off + b.in_len);
                    if (finf::crc32_block(dst, b.out_len, [](uint32_t c, const uint8_t* p, size_t n) { return (uint32_t)crc32(c, p, (uInt)n); }) == want) continue;
                }
                ++gpu_wrong;
            }
            if (use_fast && finf::inflate_raw(f.data + b.in_off, b.in_len, dst, b.out_len, *tables)) {
                const uint32_t want = rd32(f.data + b.in_off + b.in_len);
                if (finf::crc32_block(dst, b.out_len, [](uint32_t c, const uint8_t* p, size_t n) { return (uint32_t)crc32(c, p, (uInt)n); }) == want) continue;
            }
            ++fell_back;
            inflateReset(&zs);
            zs.next_in = (Bytef*)(f.data + b.in_off); zs.avail_in = b.in_len;
            zs.next_out = dst; zs.avail_out = b.out_len;
            int rc = inflate(&zs, Z_FINISH);
            if (rc != Z_STREAM_END || zs.avail_out != 0) bad = true;
            // raw inflate checks nothing: a damaged member that still is a well-formed stream of the right length must not pass
            else if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, b.out_len) != rd32(f.data + b.in_off + b.in_len)) bad = true;
        }
        inflateEnd(&zs);
        if (use_fast && fell_back) g_inflate_fallbacks += fell_back;
    });
    if (gpu_wrong) g_gpu_inflate_redone += gpu_wrong;
    if (bad) throw std::runtime_error(path + ": BGZF inflate failed");
}

// Inflate a whole BGZF file into `out`.
void bgzf_inflate_all(const MappedFile& f, const std::string& path, int threads, RawBuf& out) {
    std::vector<Block> blocks;
    size_t total = 0;
    scan_members(f, path, 0, (size_t)-1, blocks, total);
    out.resize(total);
    inflate_members(f, path, threads, blocks, out);
}

struct BamData {
    std::string path;
    RawBuf raw;                           // whole inflated file
    std::string text;
    std::vector<std::string> tid_names;
    std::vector<uint32_t> tid_lens;
    std::vector<uint64_t> rec_off;        // offset of each record's core (after block_size)
    size_t first_rec = 0;                 // offset of the first record's block_size field
    size_t rec_end = 0;                   // where the records to look at end (raw.size(), or the end of the indexed range)
    // Is the BAM header (text, reference names) complete in p[0 .. n)? (the index-driven reader inflates members until it is)
    static bool header_complete(const uint8_t* p, size_t n) {
        if (n < 12) return false;
        size_t o = 8 + (size_t)rd32(p + 4);
        if (o + 4 > n) return false;
        const uint32_t n_ref = rd32(p + o); o += 4;
        for (uint32_t i = 0; i < n_ref; ++i) {
            if (o + 4 > n) return false;
            o += 4 + (size_t)rd32(p + o) + 4;
            if (o > n) return false;
        }
        return true;
    }
    void parse_header() {
        const uint8_t* p = raw.data();
        size_t n = raw.size();
        tid_names.clear(); tid_lens.clear();
        if (n < 12 || memcmp(p, "BAM\1", 4) != 0) throw std::runtime_error(path + " is not a valid bam file");
        uint32_t l_text = rd32(p + 4);
        if (8 + (size_t)l_text + 4 > n) throw std::runtime_error(path + " is not a valid bam file");
        text.assign((const char*)p + 8, l_text);
        size_t o = 8 + l_text;
        uint32_t n_ref = rd32(p + o); o += 4;
        for (uint32_t i = 0; i < n_ref; ++i) {
            if (o + 4 > n) throw std::runtime_error(path + " is not a valid bam file");
            uint32_t l_name = rd32(p + o); o += 4;
            if (o + l_name + 4 > n) throw std::runtime_error(path + " is not a valid bam file");
            tid_names.push_back(std::string((const char*)p + o, l_name ? l_name - 1 : 0)); o += l_name;
            tid_lens.push_back(rd32(p + o)); o += 4;
        }
        first_rec = o;
        rec_end = n;
    }

    // Does a record that could be real start at o (its block_size field)? Only a guess: see find_records.
    bool plausible(size_t o) const { return brec::plausible(raw.data(), rec_end, o, (int32_t)tid_names.size()); }

    // Record boundaries are a chain of block_size prefixes, serial by nature (0.2 s for 4 M records on one core: every step is a
    // cache miss). Here the buffer is cut into segments; every segment but the first GUESSES its first record (the first offset
    // from which three records in a row look real) and follows the chain from there, all segments in parallel. The guesses are
    // then checked, in order: the chain that really arrives in a segment must land on an offset the segment's own chain visited;
    // until it does it is followed one record at a time (and it is the only chain allowed to report a broken file). The result
    // is exactly the serial chain whatever the guesses were.
    // partial_ok (windowed decode): the data may end inside a record; *tail = where that record starts (or rec_end, or the
    // offset of fewer than 4 stray bytes) and the bytes from there on are carried into the next window.
    void find_records(int threads, bool partial_ok = false, size_t* tail = nullptr) {
        const uint8_t* p = raw.data();
        const size_t n = rec_end;
        auto broken = [&]() { return std::runtime_error(path + ": truncated BAM record"); };
        size_t nseg = 1;
        if (threads > 1) nseg = std::min<size_t>((size_t)threads * 4, std::max<size_t>(1, (n - first_rec) >> 20));
        if (const char* e = getenv("BDK_CHAIN_SEGMENTS")) if (atoi(e) > 0) nseg = (size_t)atoi(e);      // tests force small segments
        std::vector<size_t> cut(nseg + 1);
        for (size_t k = 0; k <= nseg; ++k) cut[k] = first_rec + (size_t)((__uint128_t)(n - first_rec) * k / nseg);
        struct Seg { std::vector<uint64_t> own, extra; size_t end = 0, from = 0; };
        std::vector<Seg> segs(nseg);
        parallel_for(nseg, 1, threads, [&](uint64_t k0, uint64_t k1) {
            for (uint64_t k = k0; k < k1; ++k) {
                Seg& sg = segs[k];
                size_t o = cut[k];
                if (k) {
                    for (; o < cut[k + 1]; ++o) {
                        size_t q = o; int depth = 0;
                        while (depth < 3 && plausible(q)) { q += 4 + rd32(p + q); ++depth; }
                        if (depth == 3 || (depth > 0 && q + 4 > n)) break;
                    }
                    if (o >= cut[k + 1]) { sg.end = cut[k + 1]; continue; }         // no guess: the checking pass walks this segment
                }
                sg.own.reserve((cut[k + 1] - o) / 96 + 16);
                while (o + 4 <= n && o < cut[k + 1]) {
                    const uint32_t bs = rd32(p + o);
                    if (bs < 32 || o + 4 + bs > n) break;                            // real only if the true chain gets here
                    sg.own.push_back(o + 4);
                    o += 4 + (size_t)bs;
                }
                sg.end = o;
            }
        });
        size_t cur = first_rec;
        bool done = false;                                                           // partial_ok: the chain reached the cut-off record
        for (size_t k = 0; k < nseg; ++k) {
            Seg& sg = segs[k];
            if (done) { sg.from = sg.own.size(); continue; }
            size_t j = 0;
            bool joined = false;
            while (cur + 4 <= n && cur < cut[k + 1]) {
                while (j < sg.own.size() && sg.own[j] - 4 < cur) ++j;
                if (j < sg.own.size() && sg.own[j] - 4 == cur) { joined = true; break; }
                const uint32_t bs = rd32(p + cur);
                if (bs < 32 || cur + 4 + bs > n) {
                    if (partial_ok && bs >= 32) { done = true; break; }
                    throw broken();
                }
                sg.extra.push_back(cur + 4);
                cur += 4 + (size_t)bs;
            }
            if (joined) {
                sg.from = j;
                cur = sg.end;
                if (cur + 4 <= n && cur < cut[k + 1]) {                              // the segment's chain stopped on a record that cannot be
                    if (partial_ok && rd32(p + cur) >= 32) done = true;              // ... or that the window cuts off
                    else throw broken();
                }
            } else sg.from = sg.own.size();
        }
        if (tail) *tail = cur;
        if (getenv("BDK_DECODE_TRACE")) {
            size_t wrong = 0, walked = 0;
            for (auto const& sg : segs) { wrong += sg.from > 0 && sg.from <= sg.own.size() && !sg.own.empty(); walked += sg.extra.size(); }
            fprintf(stderr, "[decode] record chain: %zu segments, %zu with a wrong or unused guess, %zu records followed serially\n", nseg, wrong, walked);
        }
        std::vector<size_t> base(nseg + 1, 0);
        for (size_t k = 0; k < nseg; ++k) base[k + 1] = base[k] + segs[k].extra.size() + (segs[k].own.size() - segs[k].from);
        rec_off.resize(base[nseg]);
        parallel_for(nseg, 1, threads, [&](uint64_t k0, uint64_t k1) {
            for (uint64_t k = k0; k < k1; ++k) {
                Seg const& sg = segs[k];
                uint64_t* dst = rec_off.data() + base[k];
                if (!sg.extra.empty()) memcpy(dst, sg.extra.data(), sg.extra.size() * 8);
                if (sg.own.size() > sg.from) memcpy(dst + sg.extra.size(), sg.own.data() + sg.from, (sg.own.size() - sg.from) * 8);
            }
        });
    }
};

using brec::Core;
using brec::read_core;
using brec::aux2i;
inline const uint8_t* aux_get(const uint8_t* s, const uint8_t* e, const char tag[2]) { return brec::aux_get(s, e, tag[0], tag[1]); }

struct Region { bool on = false; int tid = -1; int beg = 0, end = 1 << 29; };

// bam_parse_region: "name", "name:beg", "name:beg-end" (1-based inclusive, commas ignored)
Region parse_region(const std::string& s, const std::vector<std::string>& names, const std::string& path) {
    Region r; r.on = true;
    std::string name = s; int beg = 0, end = 1 << 29;
    auto find_tid = [&](const std::string& nm) { for (size_t i = 0; i < names.size(); ++i) if (names[i] == nm) return (int)i; return -1; };
    int tid = find_tid(s);
    if (tid < 0) {
        size_t colon = s.rfind(':');
        if (colon != std::string::npos) {
            name = s.substr(0, colon);
            std::string rest;
            for (char c : s.substr(colon + 1)) if (c != ',') rest += c;
            size_t dash = rest.find('-');
            beg = atoi(rest.substr(0, dash).c_str());
            if (dash != std::string::npos) end = atoi(rest.substr(dash + 1).c_str());
            if (beg > 0) --beg;
            tid = find_tid(name);
        }
    }
    if (tid < 0 || beg > end)
        throw std::runtime_error("Failed to parse bam region '" + s + "' in file " + path + ". ");
    r.tid = tid; r.beg = beg; r.end = end;
    return r;
}

// Column storage that is not zero-filled (every element is written by the extraction pass; std::vector::resize would fill
// 45 bytes per record on one thread first) and that can be handed over to the stream as it is: with one bam there is nothing
// to merge, so the fields are extracted straight into the (pinned, if asked for) arrays the GPU side reads.
inline void* col_alloc(size_t bytes, int pinned) {
    if (bytes == 0) bytes = 16;
    void* p = 0;
    if (pinned) {
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess)
            throw std::runtime_error("cudaHostAlloc failed for the record columns");
    } else {
        if (posix_memalign(&p, 256, bytes) != 0) throw std::bad_alloc();
    }
    return p;
}
inline void col_free(void* p, int pinned) { if (!p) return; if (pinned) cudaFreeHost(p); else free(p); }

template <class T> struct ColBuf {
    T* p = nullptr;
    size_t n = 0;
    int pinned = 0;
    ColBuf() {}
    ColBuf(const ColBuf&) = delete;
    ColBuf& operator=(const ColBuf&) = delete;
    ColBuf(ColBuf&& o) noexcept : p(o.p), n(o.n), pinned(o.pinned) { o.p = nullptr; o.n = 0; }
    ~ColBuf() { col_free(p, pinned); }
    void resize(size_t m) { col_free(p, pinned); p = nullptr; n = 0; p = (T*)col_alloc(m * sizeof(T), pinned); n = m; }
    void grow(size_t m) {                                   // keeps the first min(n, m) elements (windowed decode appends)
        T* q = (T*)col_alloc(m * sizeof(T), pinned);
        if (p && n) memcpy(q, p, std::min(n, m) * sizeof(T));
        col_free(p, pinned);
        p = q; n = m;
    }
    T* take() { T* q = p; p = nullptr; n = 0; return q; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
};

struct Columns {
    ColBuf<int32_t> pos, mpos, tid, mtid, isize, qlen;
    ColBuf<uint16_t> flag, rgid;
    ColBuf<uint8_t> mapq;
    ColBuf<uint64_t> qid, rec;  // rec = offset of the raw record in its BamData
    bool want_rec = true;
    void set_pinned(int on) { pos.pinned = mpos.pinned = tid.pinned = mtid.pinned = isize.pinned = qlen.pinned = flag.pinned = rgid.pinned = mapq.pinned = qid.pinned = on; }
    size_t cap = 0;             // allocated elements (windowed decode grows geometrically); size() of the members is the used count
    void resize(size_t n) {
        pos.resize(n); mpos.resize(n); tid.resize(n); mtid.resize(n); isize.resize(n); qlen.resize(n);
        flag.resize(n); rgid.resize(n); mapq.resize(n); qid.resize(n);
        if (want_rec) rec.resize(n);
        cap = n;
    }
    void reserve(size_t used, size_t want) {                // room for `want` elements, the first `used` kept
        if (want <= cap) return;
        cap = want;
        pos.grow(cap); mpos.grow(cap); tid.grow(cap); mtid.grow(cap); isize.grow(cap); qlen.grow(cap);
        flag.grow(cap); rgid.grow(cap); mapq.grow(cap); qid.grow(cap);
        if (want_rec) rec.grow(cap);
        pos.n = mpos.n = tid.n = mtid.n = isize.n = qlen.n = flag.n = rgid.n = mapq.n = qid.n = used;
        if (want_rec) rec.n = used;
    }
    void append_room(size_t used, size_t more) {            // afterwards every column holds used + more elements, the first `used` kept
        const size_t need = used + more;
        if (need > cap) {
            cap = std::max(need, cap + cap / 2);
            pos.grow(cap); mpos.grow(cap); tid.grow(cap); mtid.grow(cap); isize.grow(cap); qlen.grow(cap);
            flag.grow(cap); rgid.grow(cap); mapq.grow(cap); qid.grow(cap);
            if (want_rec) rec.grow(cap);
        }
        pos.n = mpos.n = tid.n = mtid.n = isize.n = qlen.n = flag.n = rgid.n = mapq.n = qid.n = need;
        if (want_rec) rec.n = need;
    }
};

struct RgTable {
    std::mutex mu;
    std::map<std::pair<int, std::string>, int> ids;
    std::vector<int32_t> rg_lib, rg_bam;
    const Config* cfg;
    int get(int bam, const std::string& rg) {
        std::lock_guard<std::mutex> g(mu);
        auto key = std::make_pair(bam, rg);
        auto it = ids.find(key);
        if (it != ids.end()) return it->second;
        int id = (int)rg_lib.size();
        ids[key] = id;
        rg_lib.push_back(cfg->rg_lib(rg));
        rg_bam.push_back(bam);
        return id;
    }
};

}  // namespace

}  // namespace bdh

// ------------------------------------------------------------------------------------------------
struct bdh_stream {
    std::vector<bdh::BamData> bams;
    uint64_t n = 0;
    // final merged columns (malloc'ed or cudaHostAlloc'ed)
    int pinned = 0;
    int32_t *pos = 0, *mpos = 0, *tid = 0, *mtid = 0, *isize = 0, *qlen = 0;
    uint16_t *flag = 0, *rgid = 0;
    uint8_t* mapq = 0;
    uint64_t* qid = 0;
    int sorted = 1;                   // the merged stream is ordered by (tid, pos): what the summary statistics and the region builder assume
    std::vector<uint8_t> rec_bam;     // per merged record: source bam (keep_records)
    std::vector<uint64_t> rec_off;    // per merged record: raw offset (keep_records)
    std::vector<int32_t> rg_lib, rg_bam;
    std::vector<std::string> tid_names;
    double t_inflate = 0, t_extract = 0, t_merge = 0;
    std::string tmp_name;

    void* alloc(size_t bytes) { return bdh::col_alloc(bytes, pinned); }
    void release(void* p) { bdh::col_free(p, pinned); }
    ~bdh_stream() {
        release(pos); release(mpos); release(tid); release(mtid); release(isize); release(qlen);
        release(flag); release(rgid); release(mapq); release(qid);
    }
};

namespace bdh {
namespace {

// append = false: `out` receives exactly this file's kept records; append = true (windowed decode): they go behind what it holds.
void extract_bam(BamData& bd, int bam_idx, const Region& region, RgTable& rgt, int threads, Columns& out, bool append = false) {
    size_t nrec = bd.rec_off.size();
    const uint8_t* raw = bd.raw.data();
    const double tp0 = now_s();
    // pass 1: filter flags (primary && tid >= 0 [&& region overlap]) -> keep mask + prefix
    std::vector<uint8_t> keep(nrec);
    const brec::RegionSel sel{region.on ? 1 : 0, region.tid, region.beg, region.end};
    const uint64_t G = 1 << 16;
    size_t ng = (nrec + G - 1) / G;
    std::vector<uint64_t> goff(ng + 1, 0);
    parallel_for(ng, 1, threads, [&](uint64_t g0, uint64_t g1) {
      for (uint64_t g = g0; g < g1; ++g) {
        uint64_t kept = 0;
        for (uint64_t i = g * G; i < std::min<uint64_t>(nrec, (g + 1) * G); ++i) {
            if (i + 16 < nrec) __builtin_prefetch(raw + bd.rec_off[i + 16]);
            const bool ok = brec::keep_record(raw + bd.rec_off[i], sel);
            keep[i] = ok;
            kept += ok;
        }
        goff[g + 1] = kept;
      }
    });
    for (size_t g = 0; g < ng; ++g) goff[g + 1] += goff[g];
    if (append) {
        const size_t used = out.pos.size();
        out.append_room(used, goff[ng]);
        for (size_t g = 0; g <= ng; ++g) goff[g] += used;
    } else out.resize(goff[ng]);
    const double tp1 = now_s();
    // pass 2: field extraction
    parallel_for(ng, 1, threads, [&](uint64_t g0, uint64_t g1) {
        std::string last_rg; int last_id = -1; bool have_last = false;
        std::unordered_map<std::string, int> seen;  // per-thread cache in front of the shared table
        for (uint64_t g = g0; g < g1; ++g) {
            uint64_t o = goff[g];
            for (uint64_t i = g * G; i < std::min<uint64_t>(nrec, (g + 1) * G); ++i) {
                if (!keep[i]) continue;
                if (i + 13 < nrec) {                           // a record is two or three cache lines: the core, and the aux fields at its end
                    const uint8_t* nx = raw + bd.rec_off[i + 12];
                    __builtin_prefetch(nx);
                    __builtin_prefetch(raw + bd.rec_off[i + 13] - 68);
                    __builtin_prefetch(nx + 64);
                }
                const brec::Fields f = brec::record_fields(raw + bd.rec_off[i]);
                out.pos[o] = f.pos; out.mpos[o] = f.mpos; out.tid[o] = f.tid; out.mtid[o] = f.mtid;
                out.isize[o] = f.isize; out.qlen[o] = f.qlen; out.flag[o] = f.flag;
                out.mapq[o] = f.mapq;
                out.qid[o] = f.qid;
                if (out.want_rec) out.rec[o] = bd.rec_off[i];
                const char* rgs = f.rg ? (const char*)f.rg : ""; const size_t rgl = f.rg_len;
                if (!have_last || last_rg.size() != rgl || memcmp(last_rg.data(), rgs, rgl) != 0) {
                    last_rg.assign(rgs, rgl);
                    auto it = seen.find(last_rg);
                    if (it != seen.end()) last_id = it->second;
                    else { last_id = rgt.get(bam_idx, last_rg); seen[last_rg] = last_id; }
                    have_last = true;
                }
                out.rgid[o] = (uint16_t)last_id;
                ++o;
            }
        }
    });
    if (getenv("BDK_DECODE_TRACE")) fprintf(stderr, "[decode] extract: filter pass %.3f s, field pass %.3f s (%zu records)\n", tp1 - tp0, now_s() - tp1, nrec);
}

// ---- -o with a .bai: only the members that hold the reference sequence's records are inflated -----------------------------
// The reference reads `-o` regions through samtools' index (RegionLimitedBamReader.hpp:40-66: bam_index_load, bam_parse_region,
// bam_iter_query); without this every per-chromosome process (the reference's way to use many cores, and our per-GPU shards of
// a whole-genome bam) would inflate the whole file. Written from the SAM specification, section 5.2: per reference sequence a
// list of bins, each a list of chunks (begin, end) of virtual offsets (file offset of a member << 16 | offset inside its
// output). All records with this tid lie between the smallest chunk begin and the largest chunk end (the file is sorted);
// whatever else the range holds is removed by the same overlap test as without an index.
struct BaiRange { bool found = false, any = false; uint64_t beg = ~0ull, end = 0; };

BaiRange bai_reference_range(const std::string& bam_path, int tid) {
    BaiRange r;
    std::vector<std::string> cand{bam_path + ".bai"};
    if (bam_path.size() > 4 && bam_path.compare(bam_path.size() - 4, 4, ".bam") == 0) cand.push_back(bam_path.substr(0, bam_path.size() - 4) + ".bai");
    std::vector<uint8_t> d;
    for (auto const& c : cand) {
        std::ifstream in(c, std::ios::binary);
        if (!in) continue;
        d.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
        break;
    }
    if (d.size() < 8 || memcmp(d.data(), "BAI\1", 4) != 0) return r;
    size_t o = 4;
    auto need = [&](size_t k) { if (o + k > d.size()) throw std::runtime_error(bam_path + ": truncated bam index"); };
    need(4);
    const int32_t n_ref = rdi32(&d[o]); o += 4;
    if (tid < 0 || tid >= n_ref) return r;
    for (int t = 0; t <= tid; ++t) {
        need(4);
        const int32_t n_bin = rdi32(&d[o]); o += 4;
        for (int32_t b = 0; b < n_bin; ++b) {
            need(8);
            const uint32_t bin = rd32(&d[o]);
            const int32_t n_chunk = rdi32(&d[o + 4]); o += 8;
            if (n_chunk < 0) throw std::runtime_error(bam_path + ": truncated bam index");
            need(16 * (size_t)n_chunk);
            if (t == tid && bin < 37450)                                  // 37450: samtools' pseudo-bin with counts, not chunks
                for (int32_t c = 0; c < n_chunk; ++c) {
                    r.beg = std::min(r.beg, rd64(&d[o + 16 * (size_t)c]));
                    r.end = std::max(r.end, rd64(&d[o + 16 * (size_t)c + 8]));
                    r.any = true;
                }
            o += 16 * (size_t)n_chunk;
        }
        need(4);
        const int32_t n_intv = rdi32(&d[o]); o += 4;
        if (n_intv < 0) throw std::runtime_error(bam_path + ": truncated bam index");
        need(8 * (size_t)n_intv);
        o += 8 * (size_t)n_intv;
    }
    r.found = true;
    return r;
}

// Per reference sequence: record counts from samtools' pseudo-bin 37450 (second chunk = mapped, unmapped) where the index has
// it, and the compressed bytes its chunks span; what a planner needs to spread chromosomes over GPUs without decoding anything.
// Returns the number of reference sequences in the index, or -1 without a usable index.
int bai_reference_stats(const std::string& bam_path, std::vector<int64_t>& records, std::vector<int64_t>& bytes) {
    std::vector<std::string> cand{bam_path + ".bai"};
    if (bam_path.size() > 4 && bam_path.compare(bam_path.size() - 4, 4, ".bam") == 0) cand.push_back(bam_path.substr(0, bam_path.size() - 4) + ".bai");
    std::vector<uint8_t> d;
    for (auto const& c : cand) {
        std::ifstream in(c, std::ios::binary);
        if (!in) continue;
        d.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
        break;
    }
    if (d.size() < 8 || memcmp(d.data(), "BAI\1", 4) != 0) return -1;
    size_t o = 4;
    auto need = [&](size_t k) { if (o + k > d.size()) throw std::runtime_error(bam_path + ": truncated bam index"); };
    need(4);
    const int32_t n_ref = rdi32(&d[o]); o += 4;
    if (n_ref < 0) return -1;
    records.assign(n_ref, -1); bytes.assign(n_ref, 0);
    for (int t = 0; t < n_ref; ++t) {
        need(4);
        const int32_t n_bin = rdi32(&d[o]); o += 4;
        uint64_t lo = ~0ull, hi = 0;
        for (int32_t b = 0; b < n_bin; ++b) {
            need(8);
            const uint32_t bin = rd32(&d[o]);
            const int32_t n_chunk = rdi32(&d[o + 4]); o += 8;
            if (n_chunk < 0) throw std::runtime_error(bam_path + ": truncated bam index");
            need(16 * (size_t)n_chunk);
            if (bin == 37450) {
                if (n_chunk >= 2) records[t] = (int64_t)(rd64(&d[o + 16]) + rd64(&d[o + 24]));
            } else {
                for (int32_t c = 0; c < n_chunk; ++c) { lo = std::min(lo, rd64(&d[o + 16 * (size_t)c])); hi = std::max(hi, rd64(&d[o + 16 * (size_t)c + 8])); }
            }
            o += 16 * (size_t)n_chunk;
        }
        if (hi > 0 && lo != ~0ull) bytes[t] = (int64_t)((hi >> 16) - (lo >> 16)) + 1;
        else if (records[t] < 0) records[t] = 0;
        need(4);
        const int32_t n_intv = rdi32(&d[o]); o += 4;
        if (n_intv < 0) throw std::runtime_error(bam_path + ": truncated bam index");
        need(8 * (size_t)n_intv);
        o += 8 * (size_t)n_intv;
    }
    return n_ref;
}

// Returns false when there is no index to use (the caller then reads the whole file).
bool inflate_region_with_index(const MappedFile& mf, const std::string& path, const char* region, int threads, BamData& bd, Region& rg) {
    {
        std::ifstream probe(path + ".bai", std::ios::binary);
        std::ifstream probe2(path.size() > 4 ? path.substr(0, path.size() - 4) + ".bai" : std::string(), std::ios::binary);
        if (!probe && !probe2) return false;
    }
    // the header: members from the start of the file, one at a time, until it is complete
    std::vector<Block> blocks;
    size_t total = 0, off = 0;
    std::vector<uint8_t> hraw;
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error(path + ": BGZF inflate failed");
    for (;;) {
        const size_t before = blocks.size();
        scan_members(mf, path, off, off, blocks, total);
        if (blocks.size() == before) { inflateEnd(&zs); throw std::runtime_error(path + " is not a valid bam file"); }
        Block const& b = blocks.back();
        hraw.resize(total);
        if (b.out_len) {
            inflateReset(&zs);
            zs.next_in = (Bytef*)(mf.data + b.in_off); zs.avail_in = b.in_len;
            zs.next_out = hraw.data() + b.out_off; zs.avail_out = b.out_len;
            if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.avail_out != 0) { inflateEnd(&zs); throw std::runtime_error(path + ": BGZF inflate failed"); }
        }
        off = b.in_off + b.in_len + 8;
        if (hraw.size() >= 4 && memcmp(hraw.data(), "BAM\1", 4) != 0) { inflateEnd(&zs); throw std::runtime_error(path + " is not a valid bam file"); }
        if (BamData::header_complete(hraw.data(), hraw.size())) break;
    }
    inflateEnd(&zs);
    const size_t hdr_end_off = off;
    bd.raw.resize(hraw.size());
    if (!hraw.empty()) memcpy(bd.raw.data(), hraw.data(), hraw.size());
    bd.parse_header();
    rg = parse_region(region, bd.tid_names, path);
    const BaiRange br = bai_reference_range(path, rg.tid);
    if (!br.found) return false;
    if (!br.any) {                                                        // no record of this reference sequence
        bd.rec_end = bd.first_rec;
        if (getenv("BDK_DECODE_TRACE")) fprintf(stderr, "[decode] %s: region %s through the index: no records, header members only\n", path.c_str(), region);
        return true;
    }
    const uint64_t cb = br.beg >> 16, ub = br.beg & 0xffff, ce = br.end >> 16, ue = br.end & 0xffff;
    auto mismatch = [&]() { return std::runtime_error(path + ": the bam index does not match the file"); };
    if (br.beg > br.end || cb >= mf.size || ce > mf.size) throw mismatch();
    if (ue > 0 || ce > 0) scan_members(mf, path, std::max<size_t>(cb, hdr_end_off), ue ? (size_t)ce : (size_t)ce - 1, blocks, total);
    bd.raw.resize(total);
    inflate_members(mf, path, threads, blocks, bd.raw);
    bd.parse_header();
    auto locate = [&](uint64_t coff, uint64_t uoff) -> size_t {
        auto it = std::lower_bound(blocks.begin(), blocks.end(), coff, [](Block const& b, uint64_t c) { return b.file_off < c; });
        if (it == blocks.end() || it->file_off != coff || uoff > it->out_len) throw mismatch();
        return it->out_off + uoff;
    };
    bd.first_rec = std::max(bd.first_rec, locate(cb, ub));
    bd.rec_end = ue ? locate(ce, ue) : total;
    if (bd.first_rec > bd.rec_end) throw mismatch();
    if (getenv("BDK_DECODE_TRACE"))
        fprintf(stderr, "[decode] %s: region %s through the index: %zu of the file's members inflated (%.1f MB)\n", path.c_str(), region, blocks.size(), total / 1e6);
    return true;
}

// ---- decode in windows: files whose inflated bytes do not fit in memory -------------------------------------------------------
// The whole-file path holds all inflated bytes at once (3-4x the file: a 30x genome is some 400 GB). Here the members are
// taken a window at a time: inflate, find the records (the window starts on a record; where it ends inside one, those bytes are
// carried into the next window), extract behind what the columns already hold, drop the window. Same records, same order.
void decode_bam_windowed(const MappedFile& mf, const std::string& path, const char* region, int threads, size_t window_bytes,
                         BamData& bd, int bam_idx, RgTable& rgt, Columns& out, Region& rg) {
    std::vector<Block> blocks;
    size_t total = 0;
    scan_members(mf, path, 0, (size_t)-1, blocks, total);
    std::vector<uint8_t> carry;
    size_t m = 0, windows = 0;
    bool first = true;
    while (first || m < blocks.size()) {
        size_t j = m, bytes = 0;
        while (j < blocks.size() && (j == m || bytes + blocks[j].out_len <= window_bytes)) bytes += blocks[j++].out_len;
        std::vector<Block> wb(blocks.begin() + m, blocks.begin() + j);
        for (auto& b : wb) b.out_off = carry.size() + (b.out_off - blocks[m].out_off);
        BamData wd;
        wd.path = path;
        wd.raw.resize(carry.size() + bytes);
        if (!carry.empty()) memcpy(wd.raw.data(), carry.data(), carry.size());
        inflate_members(mf, path, threads, wb, wd.raw);
        const bool last = j == blocks.size();
        if (first) {
            if (!BamData::header_complete(wd.raw.data(), wd.raw.size()) && !last) { window_bytes *= 2; continue; }     // a header larger than the window
            wd.parse_header();
            bd.text = wd.text; bd.tid_names = wd.tid_names; bd.tid_lens = wd.tid_lens;
            if (region && region[0]) rg = parse_region(region, bd.tid_names, path);
        } else {
            wd.tid_names = bd.tid_names;
            wd.first_rec = 0;
            wd.rec_end = wd.raw.size();
        }
        size_t tail = wd.rec_end;
        wd.find_records(threads, !last, &tail);
        extract_bam(wd, bam_idx, rg, rgt, threads, out, true);
        carry.assign(wd.raw.data() + tail, wd.raw.data() + wd.rec_end);
        if (first && !last && bytes > 0) {
            // the first window tells how many records to expect: one allocation instead of growing window by window
            const size_t used = out.pos.size();
            out.reserve(used, (size_t)((double)used / (double)bytes * (double)total * 1.05) + 4096);
        }
        m = j;
        first = false;
        ++windows;
    }
    if (getenv("BDK_DECODE_TRACE")) fprintf(stderr, "[decode] %s: %.1f MB inflated in %zu windows\n", path.c_str(), total / 1e6, windows);
}

struct Head { int bam; uint64_t i; };

}  // namespace
}  // namespace bdh

static void set_err2(char* err, int cap, const char* msg) {
    if (err && cap > 0) { strncpy(err, msg, cap - 1); err[cap - 1] = 0; }
}

extern "C" {

static void note_sortedness(bdh_stream* s, int threads) {
    std::atomic<bool> ok(true);
    const int32_t* tid = s->tid; const int32_t* pos = s->pos;
    bdh::parallel_for(s->n, 1 << 20, threads, [&](uint64_t lo, uint64_t hi) {
        bool good = true;
        for (uint64_t i = std::max<uint64_t>(lo, 1); i < hi; ++i)
            good &= tid[i - 1] < tid[i] || (tid[i - 1] == tid[i] && pos[i - 1] <= pos[i]);
        if (!good) ok = false;
    });
    s->sorted = ok ? 1 : 0;
}

bdh_stream* bdh_stream_open(const bdh_config* cfgh, const char* const* paths, int npaths, const char* region,
                            int threads, int pinned, int keep_records, char* err, int errcap) {
    using namespace bdh;
    bdh_stream* s = 0;
    try {
        const Config& cfg = cfgh->cfg;
        if (threads <= 0) threads = default_threads();
        std::vector<std::string> files;
        if (paths) for (int i = 0; i < npaths; ++i) files.push_back(paths[i]);
        else files = cfg.bam_files;
        if (files.empty()) throw std::runtime_error("BamMerger created with no input streams!");
        if (files.size() > 255) throw std::runtime_error("more than 255 bam files");
        s = new bdh_stream;
        s->pinned = pinned;
        s->bams.resize(files.size());
        RgTable rgt; rgt.cfg = &cfg;
        std::vector<Columns> cols(files.size());
        // Windowed decode (BDK_DECODE_WINDOW_MB, or by itself when the inflated file would take more than a third of the
        // machine's memory) unless the records themselves are kept for the FASTQ dump.
        bool windowed_any = false;
        auto window_for = [&](size_t file_bytes) -> size_t {
            if (keep_records) return 0;
            if (const char* e = getenv("BDK_DECODE_WINDOW_MB")) return atoll(e) > 0 ? (size_t)atoll(e) << 20 : 0;
            if (const char* e = getenv("BDK_DECODE_WINDOW_KB")) return atoll(e) > 0 ? (size_t)atoll(e) << 10 : 0;      // tests: a member per window
            const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGESIZE);
            if (pages > 0 && psz > 0 && (double)file_bytes * 4.0 > (double)pages * (double)psz / 3.0) return (size_t)4 << 30;
            return 0;
        };
        for (auto& c : cols) c.want_rec = keep_records != 0;
        for (size_t b = 0; b < files.size(); ++b) {
            BamData& bd = s->bams[b];
            bd.path = files[b];
            double t0 = now_s();
            Region rg;
            {
                MappedFile mf; mf.open(files[b]);
                const bool ranged = region && region[0] && !getenv("BDK_NO_BAI") && inflate_region_with_index(mf, files[b], region, threads, bd, rg);
                const size_t window = ranged ? 0 : window_for(mf.size);
                if (window) {
                    windowed_any = true;
                    decode_bam_windowed(mf, files[b], region, threads, window, bd, (int)b, rgt, cols[b], rg);
                    s->t_inflate += now_s() - t0;
                    if (rgt.rg_lib.size() > 65536) throw std::runtime_error("more than 65536 (bam, read group) combinations");
                    continue;
                }
                if (files.size() == 1) cols[b].set_pinned(pinned);
                if (!ranged) {
                    bgzf_inflate_all(mf, files[b], threads, bd.raw);
                    bd.parse_header();
                    if (region && region[0]) rg = parse_region(region, bd.tid_names, files[b]);
                }
            }
            double t1 = now_s();
            const size_t raw_bytes = bd.raw.size();
            bd.find_records(threads);
            double t1b = now_s();
            extract_bam(bd, (int)b, rg, rgt, threads, cols[b]);
            double t2 = now_s();
            if (getenv("BDK_DECODE_TRACE")) fprintf(stderr, "[decode] %s: %.1f MB inflated in %.3f s, record chain %.3f s, extract %.3f s\n", files[b].c_str(), raw_bytes / 1e6, t1 - t0, t1b - t1, t2 - t1b);
            s->t_inflate += t1 - t0; s->t_extract += t2 - t1;
            if (rgt.rg_lib.size() > 65536) throw std::runtime_error("more than 65536 (bam, read group) combinations");
            if (!keep_records) { bd.raw.release(); std::vector<uint64_t>().swap(bd.rec_off); }
        }
        s->tid_names = s->bams[0].tid_names;  // BamMerger uses the first stream's header (BamMerger.cpp:78)
        s->rg_lib = rgt.rg_lib; s->rg_bam = rgt.rg_bam;
        uint64_t n = 0;
        for (auto& c : cols) n += c.pos.size();
        s->n = n;
        double t3 = now_s();
        if (files.size() == 1 && !windowed_any) {
            // nothing to merge: the extraction wrote the final arrays
            Columns& c = cols[0];
            s->pos = c.pos.take(); s->mpos = c.mpos.take(); s->tid = c.tid.take(); s->mtid = c.mtid.take(); s->isize = c.isize.take();
            s->qlen = c.qlen.take(); s->flag = c.flag.take(); s->rgid = c.rgid.take(); s->mapq = c.mapq.take(); s->qid = c.qid.take();
            if (keep_records) { s->rec_bam.assign(n, 0); s->rec_off.assign(c.rec.p, c.rec.p + n); }
            s->t_merge = now_s() - t3;
            note_sortedness(s, threads);
            return s;
        }
        s->pos = (int32_t*)s->alloc(n * 4); s->mpos = (int32_t*)s->alloc(n * 4); s->tid = (int32_t*)s->alloc(n * 4);
        s->mtid = (int32_t*)s->alloc(n * 4); s->isize = (int32_t*)s->alloc(n * 4); s->qlen = (int32_t*)s->alloc(n * 4);
        s->flag = (uint16_t*)s->alloc(n * 2); s->rgid = (uint16_t*)s->alloc(n * 2);
        s->mapq = (uint8_t*)s->alloc(n); s->qid = (uint64_t*)s->alloc(n * 8);
        if (keep_records) { s->rec_bam.resize(n); s->rec_off.resize(n); }
        // merge order: for each output slot o, (bam, i)
        auto put = [&](uint64_t o, int b, uint64_t i) {
            Columns& c = cols[b];
            s->pos[o] = c.pos[i]; s->mpos[o] = c.mpos[i]; s->tid[o] = c.tid[i]; s->mtid[o] = c.mtid[i];
            s->isize[o] = c.isize[i]; s->qlen[o] = c.qlen[i]; s->flag[o] = c.flag[i]; s->rgid[o] = c.rgid[i];
            s->mapq[o] = c.mapq[i]; s->qid[o] = c.qid[i];
            if (keep_records) { s->rec_bam[o] = (uint8_t)b; s->rec_off[o] = c.rec[i]; }
        };
        if (files.size() == 1) {                             // one bam decoded in windows: a plain copy into the final arrays
            parallel_for(n, 1 << 18, threads, [&](uint64_t a, uint64_t e) { for (uint64_t i = a; i < e; ++i) put(i, 0, i); });
        } else {
            // Same container, comparator outcomes and push/pop sequence as the reference's BamMerger, so ties between bams
            // resolve the same way. Only the ORDER is decided serially, on one packed key per
            // record ((tid, pos, strand) lexicographic, the reference's comparison); the columns are moved afterwards by all
            // cores, each output chunk starting from the per-bam cursors noted when the serial pass went by.
            const size_t nb = files.size();
            std::vector<std::vector<uint64_t>> keys(nb);
            std::atomic<bool> sorted(true);            // every bam ordered by (tid, pos)? (the strand bit is not part of a bam's order)
            for (size_t b = 0; b < nb; ++b) {
                Columns const& c = cols[b];
                keys[b].resize(c.pos.size());
                uint64_t* k = keys[b].data();
                parallel_for(c.pos.size(), 1 << 18, threads, [&](uint64_t lo, uint64_t hi) {
                    for (uint64_t i = lo; i < hi; ++i)
                        k[i] = (uint64_t)(uint32_t)c.tid[i] << 33 | (uint64_t)((uint32_t)c.pos[i] ^ 0x80000000u) << 1 | (uint64_t)((c.flag[i] & 0x10) != 0);
                    bool ok = true;
                    for (uint64_t i = std::max<uint64_t>(lo, 1); i < hi; ++i)
                        ok &= c.tid[i - 1] < c.tid[i] || (c.tid[i - 1] == c.tid[i] && c.pos[i - 1] <= c.pos[i]);
                    if (!ok) sorted = false;
                });
            }
            if (nb == 2 && sorted && !getenv("BDK_MERGE_HEAP")) {
                // Two streams: the heap's behaviour has a closed form. After a pop from stream A the heap holds B's head alone;
                // A's next record is pushed below it and sifts up only if it is STRICTLY smaller, so on a tie the stream that
                // did not emit last goes first (at the very start: bam 0, pushed first). A history-free rule, so the merge of
                // two bams that are each ordered by (tid, pos) can be cut at any (tid, pos) that starts in bam 0 and does not
                // occur in bam 1 (everything before it leaves both streams first, and the heads differ after it): found by a few
                // probes after each even cut. Unsorted input takes the priority queue below.
                const uint64_t* K0 = keys[0].data(); const uint64_t* K1 = keys[1].data();
                const uint64_t n0 = keys[0].size(), n1 = keys[1].size();
                std::vector<std::pair<uint64_t, uint64_t>> cuts;
                cuts.push_back({0, 0});
                const uint64_t parts = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)threads * 4, n >> 14));
                for (uint64_t q = 1; q < parts; ++q) {
                    uint64_t i0 = n0 * q / parts;
                    if (i0 <= cuts.back().first) continue;
                    for (int tries = 0; i0 < n0 && tries < 4096; ++tries, ++i0) {
                        if ((K0[i0] >> 1) == (K0[i0 - 1] >> 1)) continue;
                        const uint64_t want = K0[i0] >> 1;
                        const uint64_t i1 = std::lower_bound(K1, K1 + n1, want, [](uint64_t k, uint64_t w) { return (k >> 1) < w; }) - K1;
                        if (i1 < n1 && (K1[i1] >> 1) == want) continue;
                        cuts.push_back({i0, i1});
                        break;
                    }
                }
                cuts.push_back({n0, n1});
                if (getenv("BDK_DECODE_TRACE")) fprintf(stderr, "[decode] two-way merge in %zu independent parts\n", cuts.size() - 1);
                parallel_for(cuts.size() - 1, 1, threads, [&](uint64_t p0, uint64_t p1) {
                    for (uint64_t q = p0; q < p1; ++q) {
                        uint64_t i = cuts[q].first, j = cuts[q].second;
                        const uint64_t e0 = cuts[q + 1].first, e1 = cuts[q + 1].second;
                        uint64_t o = i + j;
                        int last = 1;
                        while (i < e0 && j < e1) {
                            const bool take0 = last == 0 ? K0[i] < K1[j] : K0[i] <= K1[j];
                            if (take0) { put(o++, 0, i++); last = 0; } else { put(o++, 1, j++); last = 1; }
                        }
                        while (i < e0) put(o++, 0, i++);
                        while (j < e1) put(o++, 1, j++);
                    }
                });
                s->t_merge = now_s() - t3;
                note_sortedness(s, threads);
                return s;
            }
            // Three or more bams (or unsorted input): the queue itself. Up to 16 bams of fewer than 2^28 records each go through
            // nway_merge.hpp -- libstdc++'s heap on a fixed array and, for up to five sorted bams, its exact parallel form (parts
            // cut at position boundaries, one run per valid heap layout at the cut); BDK_MERGE_HEAP keeps the plain queue below.
            bool small = nb <= 16 && !getenv("BDK_MERGE_HEAP");
            for (size_t b = 0; b < nb; ++b) small = small && keys[b].size() < (1ull << 28);
            if (small) {
                std::vector<const uint64_t*> kp(nb);
                std::vector<uint64_t> counts(nb);
                for (size_t b = 0; b < nb; ++b) { kp[b] = keys[b].data(); counts[b] = keys[b].size(); }
                std::vector<uint32_t> order(n);
                if (!(sorted && threads >= 8 && bdh::nway_merge_order_parallel(kp.data(), counts.data(), (int)nb, 28, order.data(), std::min(threads, 32))))
                    bdh::nway_merge_order(kp.data(), counts.data(), (int)nb, 28, order.data());
                parallel_for(n, 1 << 16, threads, [&](uint64_t o0, uint64_t o1) {
                    for (uint64_t o = o0; o < o1; ++o) put(o, (int)(order[o] >> 28), order[o] & ((1u << 28) - 1u));
                });
                s->t_merge = now_s() - t3;
                note_sortedness(s, threads);
                return s;
            }
            auto greater = [&](const Head& x, const Head& y) { return keys[x.bam][x.i] > keys[y.bam][y.i]; };
            std::priority_queue<Head, std::vector<Head>, decltype(greater)> pq(greater);
            for (size_t b = 0; b < nb; ++b)
                if (!cols[b].pos.empty()) pq.push(Head{(int)b, 0});
            const uint64_t CH = 1 << 16;
            const uint64_t nch = (n + CH - 1) / CH;
            std::vector<uint8_t> src(n);
            std::vector<uint64_t> cursor((nch + 1) * nb, 0), emitted(nb, 0);
            uint64_t o = 0;
            while (!pq.empty()) {
                Head h = pq.top(); pq.pop();
                if ((o & (CH - 1)) == 0) for (size_t b = 0; b < nb; ++b) cursor[(o / CH) * nb + b] = emitted[b];
                src[o++] = (uint8_t)h.bam;
                ++emitted[h.bam];
                if (h.i + 1 < cols[h.bam].pos.size()) pq.push(Head{h.bam, h.i + 1});
            }
            parallel_for(nch, 1, threads, [&](uint64_t c0, uint64_t c1) {
                std::vector<uint64_t> cur(nb);
                for (uint64_t ch = c0; ch < c1; ++ch) {
                    for (size_t b = 0; b < nb; ++b) cur[b] = cursor[ch * nb + b];
                    for (uint64_t q = ch * CH; q < std::min<uint64_t>(n, (ch + 1) * CH); ++q) { const int b = src[q]; put(q, b, cur[b]++); }
                }
            });
        }
        s->t_merge = now_s() - t3;
        note_sortedness(s, threads);
        return s;
    } catch (std::exception const& e) {
        set_err2(err, errcap, e.what());
        delete s;
        return 0;
    }
}

void bdh_stream_free(bdh_stream* s) { delete s; }
uint64_t bdh_stream_n(const bdh_stream* s) { return s->n; }
void bdh_stream_cols(const bdh_stream* s, bdk_soa* o) {
    o->pos = s->pos; o->mpos = s->mpos; o->tid = s->tid; o->mtid = s->mtid; o->isize = s->isize;
    o->flag = s->flag; o->mapq = s->mapq; o->rgid = s->rgid; o->qlen = s->qlen; o->qid = s->qid;
}
int bdh_stream_nrg(const bdh_stream* s) { return (int)s->rg_lib.size(); }
const int32_t* bdh_stream_rg_lib(const bdh_stream* s) { return s->rg_lib.data(); }
const int32_t* bdh_stream_rg_bam(const bdh_stream* s) { return s->rg_bam.data(); }
int bdh_stream_ntid(const bdh_stream* s) { return (int)s->tid_names.size(); }
const char* bdh_stream_tid_name(const bdh_stream* s, int tid) {
    return tid >= 0 && tid < (int)s->tid_names.size() ? s->tid_names[tid].c_str() : "";
}
int bdh_bai_reference_stats(const char* bam_path, int64_t* records, int64_t* bytes, int cap, char* err, int errcap) {
    try {
        std::vector<int64_t> r, b;
        const int n = bdh::bai_reference_stats(bam_path, r, b);
        for (int i = 0; i < n && i < cap; ++i) { if (records) records[i] = r[i]; if (bytes) bytes[i] = b[i]; }
        return n;
    } catch (std::exception const& e) {
        set_err2(err, errcap, e.what());
        return -2;
    }
}
void bdh_inflate_counters(uint64_t* host_fallbacks, uint64_t* gpu_redone) {
    if (host_fallbacks) *host_fallbacks = bdh::g_inflate_fallbacks.load();
    if (gpu_redone) *gpu_redone = bdh::g_gpu_inflate_redone.load();
}
int bdh_stream_sorted(const bdh_stream* s) { return s ? s->sorted : 1; }
void bdh_stream_timings(const bdh_stream* s, double* a, double* b, double* c) {
    if (a) *a = s->t_inflate;
    if (b) *b = s->t_extract;
    if (c) *c = s->t_merge;
}

const char* bdh_stream_qname(const bdh_stream* s, uint64_t i) {
    if (i >= s->n || s->rec_off.empty()) return 0;
    return (const char*)(s->bams[s->rec_bam[i]].raw.data() + s->rec_off[i] + 32);
}

int bdh_stream_fastq(const bdh_stream* s, uint64_t i, char* buf, int cap) {
    using namespace bdh;
    if (i >= s->n || s->rec_off.empty()) return -1;
    const uint8_t* r = s->bams[s->rec_bam[i]].raw.data() + s->rec_off[i];
    Core c = read_core(r);
    const char* name = (const char*)r + 32;
    const uint8_t* seq = r + 32 + c.l_qname + 4 * (size_t)c.n_cigar;
    const uint8_t* qual = seq + ((size_t)c.l_qseq + 1) / 2;
    std::string out = "@"; out += name; out += "\n";
    static const char* nt16 = "=ACMGRSVTWYHKDBN";
    for (int k = 0; k < c.l_qseq; ++k) out += nt16[(seq[k >> 1] >> ((~k & 1) << 2)) & 0xf];
    out += "\n+\n";
    if (c.l_qseq > 0 && qual[0] != 0xff) for (int k = 0; k < c.l_qseq; ++k) out += char(qual[k] + 33);
    out += "\n";
    if ((int)out.size() + 1 > cap) return -1;
    memcpy(buf, out.c_str(), out.size() + 1);
    return (int)out.size();
}

// ---- device-resident decode: the host's share (bdk_push_bam does the rest on the GPU) ---------------------------------------
// Map the file, list its BGZF members (through the .bai for a region, like inflate_region_with_index), inflate and parse the
// header members here (a few KB), and number the read groups: the config's read groups first, one more id for every other
// string (the reference maps those to the first bam's library, BamConfig.hpp:63-72).
struct bdh_bamdev {
    bdh::MappedFile mf;
    std::string path;
    bdh::BamData hdr;
    bdh::Region rg;
    std::vector<bdk_bgzf_member> members;
    uint64_t first_record = 0, end_offset = 0;
    std::vector<uint64_t> rg_hash;
    std::vector<uint16_t> rg_id;
    std::vector<int32_t> rg_lib, rg_bam;
    double t_open = 0;
    int id_base = 0;                  // first read-group id of this file (second bam of a two-bam run: behind the first bam's ids)
};

static bdh_bamdev* bamdev_open_impl(const bdh_config* cfgh, const char* path, const char* region, int id_base, char* err, int errcap) {
    using namespace bdh;
    bdh_bamdev* d = nullptr;
    try {
        const double t0 = now_s();
        const Config& cfg = cfgh->cfg;
        int bam_index = 0;
        std::string file;
        if (path && path[0]) {
            file = path;
            for (size_t b = 0; b < cfg.bam_files.size(); ++b) if (cfg.bam_files[b] == file) bam_index = (int)b;
        } else {
            if (cfg.bam_files.size() != 1) throw std::runtime_error("the device decode reads one bam file; the config lists " + std::to_string(cfg.bam_files.size()));
            file = cfg.bam_files[0];
        }
        d = new bdh_bamdev;
        d->path = file;
        d->mf.open(file);
        // header members, one at a time, until the header is complete
        std::vector<Block> blocks;
        size_t total = 0, off = 0;
        std::vector<uint8_t> hraw;
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error(file + ": BGZF inflate failed");
        for (;;) {
            const size_t before = blocks.size();
            scan_members(d->mf, file, off, off, blocks, total);
            if (blocks.size() == before) { inflateEnd(&zs); throw std::runtime_error(file + " is not a valid bam file"); }
            Block const& b = blocks.back();
            hraw.resize(total);
            if (b.out_len) {
                inflateReset(&zs);
                zs.next_in = (Bytef*)(d->mf.data + b.in_off); zs.avail_in = b.in_len;
                zs.next_out = hraw.data() + b.out_off; zs.avail_out = b.out_len;
                if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.avail_out != 0) { inflateEnd(&zs); throw std::runtime_error(file + ": BGZF inflate failed"); }
                if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), hraw.data() + b.out_off, b.out_len) != rd32(d->mf.data + b.in_off + b.in_len)) { inflateEnd(&zs); throw std::runtime_error(file + ": BGZF inflate failed"); }
            }
            off = b.in_off + b.in_len + 8;
            if (hraw.size() >= 4 && memcmp(hraw.data(), "BAM\1", 4) != 0) { inflateEnd(&zs); throw std::runtime_error(file + " is not a valid bam file"); }
            if (BamData::header_complete(hraw.data(), hraw.size())) break;
            if (off >= d->mf.size) { inflateEnd(&zs); throw std::runtime_error(file + " is not a valid bam file"); }
        }
        inflateEnd(&zs);
        const size_t hdr_end_off = off;
        d->hdr.path = file;
        d->hdr.raw.resize(hraw.size());
        if (!hraw.empty()) memcpy(d->hdr.raw.data(), hraw.data(), hraw.size());
        d->hdr.parse_header();
        size_t first_rec = d->hdr.first_rec, rec_end = 0;       // offsets in the concatenated output of `blocks`; rec_end 0 = to the end
        bool none = false;
        bool ranged = false;
        if (region && region[0]) {
            d->rg = parse_region(region, d->hdr.tid_names, file);
            if (!getenv("BDK_NO_BAI")) {
                const BaiRange br = bai_reference_range(file, d->rg.tid);
                if (br.found) {
                    ranged = true;
                    if (!br.any) none = true;
                    else {
                        const uint64_t cb = br.beg >> 16, ub = br.beg & 0xffff, ce = br.end >> 16, ue = br.end & 0xffff;
                        auto mismatch = [&]() { return std::runtime_error(file + ": the bam index does not match the file"); };
                        if (br.beg > br.end || cb >= d->mf.size || ce > d->mf.size) throw mismatch();
                        if (ue > 0 || ce > 0) scan_members(d->mf, file, std::max<size_t>(cb, hdr_end_off), ue ? (size_t)ce : (size_t)ce - 1, blocks, total);
                        auto locate = [&](uint64_t coff, uint64_t uoff) -> size_t {
                            auto it = std::lower_bound(blocks.begin(), blocks.end(), coff, [](Block const& b, uint64_t c) { return b.file_off < c; });
                            if (it == blocks.end() || it->file_off != coff || uoff > it->out_len) throw mismatch();
                            return it->out_off + uoff;
                        };
                        first_rec = std::max(first_rec, locate(cb, ub));
                        rec_end = ue ? locate(ce, ue) : total;
                        if (first_rec > rec_end) throw mismatch();
                    }
                }
            }
        }
        if (!ranged) scan_members(d->mf, file, hdr_end_off, (size_t)-1, blocks, total);
        if (!none) {
            // from the member that holds the first record; members without output are left out
            size_t i0 = 0;
            while (i0 < blocks.size() && blocks[i0].out_off + blocks[i0].out_len <= first_rec) ++i0;
            if (i0 < blocks.size()) {
                const size_t shift = blocks[i0].out_off;
                uint64_t o = 0;
                for (size_t i = i0; i < blocks.size(); ++i) {
                    if (!blocks[i].out_len) continue;
                    if (blocks[i].out_off - shift != o) throw std::runtime_error(file + ": internal error in the member list");
                    d->members.push_back(bdk_bgzf_member{(uint64_t)blocks[i].in_off, o, blocks[i].in_len, blocks[i].out_len});
                    o += blocks[i].out_len;
                }
                d->first_record = first_rec - shift;
                d->end_offset = rec_end ? rec_end - shift : 0;
                if (rec_end && rec_end - shift == d->first_record) d->members.clear();
            }
        }
        // read groups: the config's, in its map order, then "any other"
        for (auto const& kv : cfg.readgroup_library) {
            d->rg_hash.push_back(brec::hash_bytes((const uint8_t*)kv.first.data(), kv.first.size()));
            d->rg_id.push_back((uint16_t)(id_base + d->rg_lib.size()));
            d->rg_lib.push_back(cfg.rg_lib(kv.first));
            d->rg_bam.push_back(bam_index);
        }
        if (id_base + d->rg_lib.size() >= 65535) throw std::runtime_error("more than 65535 read groups in the config");
        d->id_base = id_base;
        {
            auto li = cfg.lib_index.find(cfg.first_bam_library);
            d->rg_lib.push_back(cfg.first_bam_library.empty() || li == cfg.lib_index.end() ? -1 : li->second);
            d->rg_bam.push_back(bam_index);
        }
        d->t_open = now_s() - t0;
        return d;
    } catch (std::exception const& e) {
        set_err2(err, errcap, e.what());
        delete d;
        return 0;
    }
}
bdh_bamdev* bdh_bamdev_open(const bdh_config* cfgh, const char* path, const char* region, char* err, int errcap) {
    return bamdev_open_impl(cfgh, path, region, 0, err, errcap);
}
bdh_bamdev* bdh_bamdev_open_next(const bdh_config* cfgh, const bdh_bamdev* first, const char* path, const char* region, char* err, int errcap) {
    if (!first || !path || !path[0]) { set_err2(err, errcap, "bdh_bamdev_open_next: the first bam and a path are needed"); return 0; }
    return bamdev_open_impl(cfgh, path, region, first->id_base + (int)first->rg_lib.size(), err, errcap);
}
void bdh_bamdev_free(bdh_bamdev* d) { delete d; }
int bdh_bamdev_nrg(const bdh_bamdev* d) { return (int)d->rg_lib.size(); }
const int32_t* bdh_bamdev_rg_lib(const bdh_bamdev* d) { return d->rg_lib.data(); }
const int32_t* bdh_bamdev_rg_bam(const bdh_bamdev* d) { return d->rg_bam.data(); }
int bdh_bamdev_ntid(const bdh_bamdev* d) { return (int)d->hdr.tid_names.size(); }
const char* bdh_bamdev_tid_name(const bdh_bamdev* d, int tid) { return tid >= 0 && tid < (int)d->hdr.tid_names.size() ? d->hdr.tid_names[tid].c_str() : ""; }
uint64_t bdh_bamdev_members(const bdh_bamdev* d) { return d->members.size(); }
uint64_t bdh_bamdev_file_bytes(const bdh_bamdev* d) { return d->mf.size; }
static void bamdev_source(const bdh_bamdev* d, bdk_bam_source& s) {
    memset(&s, 0, sizeof s);
    s.file = d->mf.data; s.file_bytes = d->mf.size;
    s.members = d->members.data(); s.n_members = d->members.size();
    s.first_record = d->first_record; s.end_offset = d->end_offset;
    s.n_ref = (int32_t)d->hdr.tid_names.size();
    s.region_on = d->rg.on ? 1 : 0; s.region_tid = d->rg.tid; s.region_beg = d->rg.beg; s.region_end = d->rg.end;
    s.n_rg = (uint32_t)d->rg_hash.size(); s.rg_hash = d->rg_hash.data(); s.rg_id = d->rg_id.data();
    s.rg_other = (uint16_t)(d->id_base + d->rg_lib.size() - 1);
}
int bdh_bamdev_push(bdh_bamdev* d, bdk_ctx* ctx, bdk_bam_stats* stats) {
    bdk_bam_source s;
    bamdev_source(d, s);
    return bdk_push_bam(ctx, &s, stats);
}
int bdh_bamdev_push2(bdh_bamdev* first, bdh_bamdev* second, bdk_ctx* ctx, bdk_bam_stats* stats2) {
    bdk_bam_source s[2];
    bamdev_source(first, s[0]); bamdev_source(second, s[1]);
    return bdk_push_bams(ctx, s, 2, stats2);
}
int bdh_bamdev_decode2(bdh_bamdev* first, bdh_bamdev* second, bdk_ctx* ctx, const bdk_soa* host_out, uint64_t cap, bdk_bam_stats* stats2) {
    bdk_bam_source s[2];
    bamdev_source(first, s[0]); bamdev_source(second, s[1]);
    return bdk_decode_bams(ctx, s, 2, host_out, cap, stats2);
}
int bdh_bamdev_pushn(bdh_bamdev* const* devs, int n, bdk_ctx* ctx, bdk_bam_stats* stats) {
    if (!devs || n < 1) return BDK_ERR_ARG;
    std::vector<bdk_bam_source> s((size_t)n);
    for (int b = 0; b < n; ++b) { if (!devs[b]) return BDK_ERR_ARG; bamdev_source(devs[b], s[b]); }
    return bdk_push_bams(ctx, s.data(), n, stats);
}
int bdh_bamdev_decoden(bdh_bamdev* const* devs, int n, bdk_ctx* ctx, const bdk_soa* host_out, uint64_t cap, bdk_bam_stats* stats) {
    if (!devs || n < 1) return BDK_ERR_ARG;
    std::vector<bdk_bam_source> s((size_t)n);
    for (int b = 0; b < n; ++b) { if (!devs[b]) return BDK_ERR_ARG; bamdev_source(devs[b], s[b]); }
    return bdk_decode_bams(ctx, s.data(), n, host_out, cap, stats);
}
int bdh_bamdev_decode(bdh_bamdev* d, bdk_ctx* ctx, const bdk_soa* host_out, uint64_t cap, bdk_bam_stats* stats) {
    bdk_bam_source s;
    bamdev_source(d, s);
    return bdk_decode_bam(ctx, &s, host_out, cap, stats);
}

// ---- writer -------------------------------------------------------------------------------------
int bdh_write_bam(const char* path, int ntid, const char* const* tid_names, const uint32_t* tid_lens,
                  int nrg, const char* const* rg_names, const bdk_soa* cols, uint64_t n,
                  const char* name_prefix, int write_am, int level, int threads, char* err, int errcap) {
    using namespace bdh;
    try {
        if (threads <= 0) threads = default_threads();
        FILE* fp = fopen(path, "wb");
        if (!fp) throw std::runtime_error(std::string("cannot create ") + path);
        std::vector<uint8_t> stream;
        auto put32 = [&](std::vector<uint8_t>& v, uint32_t x) { uint8_t b[4]; memcpy(b, &x, 4); v.insert(v.end(), b, b + 4); };
        // header
        std::string text = "@HD\tVN:1.0\tSO:coordinate\n";
        for (int i = 0; i < ntid; ++i) text += std::string("@SQ\tSN:") + tid_names[i] + "\tLN:" + std::to_string(tid_lens[i]) + "\n";
        for (int i = 0; i < nrg; ++i) text += std::string("@RG\tID:") + rg_names[i] + "\tSM:s\tLB:" + rg_names[i] + "\n";
        stream.insert(stream.end(), {'B', 'A', 'M', 1});
        put32(stream, (uint32_t)text.size());
        stream.insert(stream.end(), text.begin(), text.end());
        put32(stream, (uint32_t)ntid);
        for (int i = 0; i < ntid; ++i) {
            uint32_t l = (uint32_t)strlen(tid_names[i]) + 1;
            put32(stream, l);
            stream.insert(stream.end(), tid_names[i], tid_names[i] + l);
            put32(stream, tid_lens[i]);
        }
        const uint64_t BATCH = 1 << 21;
        const size_t BLK = 0xff00;
        auto flush_blocks = [&](std::vector<uint8_t>& data, bool final) {
            size_t nblk = final ? (data.size() + BLK - 1) / BLK : data.size() / BLK;
            std::vector<std::vector<uint8_t>> comp(nblk);
            std::atomic<bool> bad(false);
            parallel_for(nblk, 16, threads, [&](uint64_t a, uint64_t b) {
                z_stream zs; memset(&zs, 0, sizeof(zs));
                if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad = true; return; }
                for (uint64_t k = a; k < b; ++k) {
                    size_t off = k * BLK, len = std::min(BLK, data.size() - off);
                    std::vector<uint8_t>& o = comp[k];
                    o.resize(18 + compressBound(len) + 64 + 8);
                    deflateReset(&zs);
                    zs.next_in = data.data() + off; zs.avail_in = (uInt)len;
                    zs.next_out = o.data() + 18; zs.avail_out = (uInt)(o.size() - 18 - 8);
                    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { bad = true; break; }
                    size_t clen = zs.total_out;
                    size_t total = 18 + clen + 8;
                    if (total > 65536) { bad = true; break; }
                    const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0};
                    memcpy(o.data(), hdr, 16);
                    uint16_t bs = (uint16_t)(total - 1); memcpy(o.data() + 16, &bs, 2);
                    uint32_t crc = crc32(crc32(0, 0, 0), data.data() + off, (uInt)len), isz = (uint32_t)len;
                    memcpy(o.data() + 18 + clen, &crc, 4); memcpy(o.data() + 18 + clen + 4, &isz, 4);
                    o.resize(total);
                }
                deflateEnd(&zs);
            });
            if (bad) throw std::runtime_error("BGZF deflate failed");
            for (auto& o : comp) if (fwrite(o.data(), 1, o.size(), fp) != o.size()) throw std::runtime_error("short write");
            data.erase(data.begin(), data.begin() + std::min(data.size(), nblk * BLK));
        };
        std::string prefix = name_prefix ? name_prefix : "r";
        for (uint64_t b0 = 0; b0 < n; b0 += BATCH) {
            uint64_t b1 = std::min(n, b0 + BATCH);
            // build records of the batch in parallel pieces, then append in order
            const uint64_t PIECE = 1 << 15;
            size_t np = (b1 - b0 + PIECE - 1) / PIECE;
            std::vector<std::vector<uint8_t>> pieces(np);
            parallel_for(np, 1, threads, [&](uint64_t p0, uint64_t p1) {
                for (uint64_t p = p0; p < p1; ++p) {
                    std::vector<uint8_t>& v = pieces[p];
                    v.reserve(PIECE * 160);
                    for (uint64_t i = b0 + p * PIECE; i < std::min(b1, b0 + (p + 1) * PIECE); ++i) {
                        char nm[64];
                        int nl = snprintf(nm, sizeof nm, "%s%llu", prefix.c_str(), (unsigned long long)cols->qid[i]) + 1;
                        int32_t ql = cols->qlen[i];
                        uint16_t fl = cols->flag[i];
                        int ncig = (ql > 0 && !(fl & 4)) ? 1 : 0;
                        const char* rg = rg_names[cols->rgid[i]];
                        size_t rgl = strlen(rg) + 1;
                        size_t auxl = 3 + rgl + (write_am ? 4 : 0);
                        uint32_t bs = 32 + nl + 4 * ncig + (ql + 1) / 2 + ql + (uint32_t)auxl;
                        size_t at = v.size();
                        v.resize(at + 4 + bs);
                        uint8_t* r = v.data() + at;
                        memcpy(r, &bs, 4); r += 4;
                        int32_t pos = cols->pos[i], tid = cols->tid[i];
                        uint32_t endp = ncig ? (uint32_t)pos + ql : (uint32_t)pos + 1, beg = (uint32_t)pos, e = endp - 1;
                        uint16_t bin;
                        if (beg >> 14 == e >> 14) bin = 4681 + (beg >> 14);
                        else if (beg >> 17 == e >> 17) bin = 585 + (beg >> 17);
                        else if (beg >> 20 == e >> 20) bin = 73 + (beg >> 20);
                        else if (beg >> 23 == e >> 23) bin = 9 + (beg >> 23);
                        else if (beg >> 26 == e >> 26) bin = 1 + (beg >> 26);
                        else bin = 0;
                        memcpy(r, &tid, 4); memcpy(r + 4, &pos, 4);
                        r[8] = (uint8_t)nl; r[9] = cols->mapq[i]; memcpy(r + 10, &bin, 2);
                        uint16_t nc = (uint16_t)ncig; memcpy(r + 12, &nc, 2); memcpy(r + 14, &fl, 2);
                        memcpy(r + 16, &ql, 4); memcpy(r + 20, &cols->mtid[i], 4); memcpy(r + 24, &cols->mpos[i], 4);
                        memcpy(r + 28, &cols->isize[i], 4);
                        uint8_t* q = r + 32;
                        memcpy(q, nm, nl); q += nl;
                        if (ncig) { uint32_t cg = ((uint32_t)ql << 4) | 0; memcpy(q, &cg, 4); q += 4; }
                        uint64_t h = cols->qid[i] * 0x9E3779B97F4A7C15ull + fl;
                        for (int k = 0; k < (ql + 1) / 2; ++k) {
                            static const uint8_t code[4] = {1, 2, 4, 8};
                            h = h * 6364136223846793005ull + 1442695040888963407ull;
                            uint8_t hi = code[(h >> 60) & 3], lo = (2 * k + 1 < ql) ? code[(h >> 58) & 3] : 0;
                            *q++ = (uint8_t)(hi << 4 | lo);
                        }
                        for (int k = 0; k < ql; ++k) *q++ = (uint8_t)(20 + ((h >> (k & 31)) & 15));
                        q[0] = 'R'; q[1] = 'G'; q[2] = 'Z'; memcpy(q + 3, rg, rgl); q += 3 + rgl;
                        if (write_am) { q[0] = 'A'; q[1] = 'M'; q[2] = 'C'; q[3] = cols->mapq[i]; q += 4; }
                    }
                }
            });
            for (auto& p : pieces) stream.insert(stream.end(), p.begin(), p.end());
            flush_blocks(stream, false);
        }
        flush_blocks(stream, true);
        static const uint8_t eof[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        fwrite(eof, 1, 28, fp);
        if (fclose(fp) != 0) throw std::runtime_error("close failed");
        return 0;
    } catch (std::exception const& e) {
        set_err2(err, errcap, e.what());
        return -1;
    }
}

}  // extern "C"

#include "bam2cfg_impl.hpp"
