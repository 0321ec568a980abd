// ref_nodecode.cpp -- the UNMODIFIED reference algorithm timed WITHOUT BAM decoding.
//
// TEST / BENCH INFRASTRUCTURE ONLY: built by oracle/build_ref.sh into oracle/_ref/breakdancer-max-nodecode, executed only by
// bench.py's cpu_baseline_decode_free leg and tests/. Nothing under breakdancer_b200/ uses it.
//
// How: the reference's main() (src/exe/breakdancer-max/BreakDancerMax.cpp, compiled where it lies with -Dmain=reference_main)
// and all of its library objects are linked unchanged EXCEPT io/BamIo.o, whose two factory functions openBam / openBams
// (src/lib/io/BamIo.cpp:6-31) are defined here instead: they hand out readers over records that were decoded into memory
// beforehand -- a BamReaderBase subclass (src/lib/io/BamReaderBase.hpp:9-28) serving in-memory bam1_t, the seam the reference's
// own unit tests use (test/lib/io/TestAlignment.cpp:26-43). The first openBam of a path decodes the file (samtools, as the
// reference would) and that time is accounted separately; both of the reference's passes (BamSummary::_analyze_bams,
// BreakDancer::run) then run on memory. At exit one line goes to stderr:
//   nodecode wall_s=<whole run> load_s=<decoding into memory> algorithm_s=<wall - load> records=<records served per pass>
// The filter is the reference's (primary && tid >= 0, BamIo.cpp:11-18); a region (-o) keeps the records samtools' iterator
// would return (overlap with the region).
#include "io/BamIo.hpp"

#include <sam.h>
#include <bam.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

struct Loaded {
    bam_header_t* header;
    std::vector<bam1_t> recs;          // data pointers own their memory
};

std::map<std::string, Loaded*> g_cache;
double g_load_s = 0.0;
size_t g_served = 0;

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

Loaded* load(std::string const& path, std::string const& region) {
    std::string const key = path + "\t" + region;
    std::map<std::string, Loaded*>::iterator it = g_cache.find(key);
    if (it != g_cache.end()) return it->second;
    double const t0 = now_s();
    samfile_t* in = samopen(path.c_str(), "rb", 0);
    if (!in || !in->header) throw std::runtime_error("Failed to open BAM file " + path);
    Loaded* L = new Loaded;
    L->header = in->header;             // the file stays open: it owns the header
    int tid = -1, beg = 0, end = 0x7fffffff;
    if (!region.empty() && bam_parse_region(in->header, region.c_str(), &tid, &beg, &end) < 0)
        throw std::runtime_error("Failed to parse bam region " + region);
    bam1_t* b = bam_init1();
    while (samread(in, b) > 0) {
        if ((b->core.flag & (BAM_FSECONDARY | 2048)) || b->core.tid < 0) continue;          // AlignmentFilter: IsPrimary && IsAligned
        if (!region.empty()) {
            if (b->core.tid != tid) continue;
            int const e = bam_calend(&b->core, bam1_cigar(b));
            if (b->core.pos >= end || (e > b->core.pos ? e : b->core.pos + 1) <= beg) continue;
        }
        bam1_t c = *b;
        c.data = (uint8_t*)malloc(b->data_len);
        memcpy(c.data, b->data, b->data_len);
        c.m_data = b->data_len;
        L->recs.push_back(c);
    }
    bam_destroy1(b);
    g_cache[key] = L;
    g_load_s += now_s() - t0;
    return L;
}

class MemoryBamReader : public BamReaderBase {
public:
    MemoryBamReader(std::string const& path, std::string const& region) : _path(path), _l(load(path, region)), _i(0) {}
    int next(bam1_t* e) {
        if (_i >= _l->recs.size()) return -1;
        bam1_t const& s = _l->recs[_i++];
        if (e->m_data < s.data_len) {
            e->m_data = s.data_len;
            kroundup32(e->m_data);
            e->data = (uint8_t*)realloc(e->data, e->m_data);
        }
        e->core = s.core; e->l_aux = s.l_aux; e->data_len = s.data_len;
        memcpy(e->data, s.data, s.data_len);
        ++g_served;
        return s.data_len;
    }
    bam_header_t* header() const { return _l->header; }
    std::string const& path() const { return _path; }
private:
    std::string _path;
    Loaded* _l;
    size_t _i;
};

}  // namespace

BamReaderBase* openBam(std::string const& path, std::string const& region) { return new MemoryBamReader(path, region); }

std::vector<boost::shared_ptr<BamReaderBase> > openBams(std::vector<std::string> const& paths, std::string const& region) {
    std::vector<boost::shared_ptr<BamReaderBase> > rv;
    for (size_t i = 0; i < paths.size(); ++i) rv.push_back(boost::shared_ptr<BamReaderBase>(openBam(paths[i], region)));
    return rv;
}

int reference_main(int argc, char* argv[]);

int main(int argc, char* argv[]) {
    double const t0 = now_s();
    int const rc = reference_main(argc, argv);
    double const wall = now_s() - t0;
    fflush(stdout);
    fprintf(stderr, "nodecode wall_s=%.6f load_s=%.6f algorithm_s=%.6f records=%zu\n", wall, g_load_s, wall - g_load_s, g_served / 2);
    return rc;
}
