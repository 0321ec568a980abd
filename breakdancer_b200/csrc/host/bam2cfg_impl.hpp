// bam2cfg: the configuration-file generator of the tool chain (reference: perl/bam2cfg.pl:48-247 with
// perl/AlnParser.pm:31-124 for the record -> Maq-flag translation and perl/bam2cfg.pl:284-744 for the
// Shapiro-Wilk normality figure). The Perl script pipes `samtools view -h` text through regular expressions; here
// the BAM is decoded directly (this file is included at the end of bam_io.cpp and shares its BGZF / record helpers).
// Per library: insert size mean / s.d. from the first `-n` proper FR pairs (Maq flags 18 / 20, non-negative
// distance) after dropping observations above mean + 5 s.d., separate lower / upper spreads around the mean for
// the cut-offs, mean read length, and one output line per read group in the grammar BamConfigEntry.cpp:43-54 parses.
// Not carried over: -h (GD histogram plots). The order of the output lines is the order of the @RG header lines
// (the Perl script prints in hash order, i.e. randomly).
#pragma once
#include <cmath>
#include <map>
#include <sstream>

namespace bdh {
namespace {

// ---- Royston's W test as the script spells it (AS R94 with AS 241 / AS 66) ------------------------------------
double b2c_poly(const double* c, int nord, double x) {
    double v = c[0];
    if (nord == 1) return v;
    double p = x * c[nord - 1];
    if (nord == 2) return v + p;
    for (int i = 1, j = nord - 1; i <= nord - 2; ++i, --j) p = (p + c[j - 1]) * x;
    return v + p;
}

double b2c_ppnd(double p) {
    const double a0 = 3.3871327179, a1 = 5.0434271938 * 10, a2 = 1.5929113202 * 100, a3 = 5.9109374720 * 10;
    const double b1 = 1.7895169469 * 10, b2 = 7.8757757664 * 10, b3 = 6.7187563600 * 10;
    const double c0 = 1.4234372777, c1 = 2.7568153900, c2 = 1.3067284816, c3 = 1.7023821103 * 0.1;
    const double d1 = 7.3700164250 * 0.1, d2 = 1.2021132975 * 0.1;
    const double e0 = 6.6579051150, e1 = 3.0812263860, e2 = 4.2868294337 * 0.1, e3 = 1.7337203997 * 0.01;
    const double f1 = 2.4197894225 * 0.1, f2 = 1.2258202635 * 0.01;
    const double q = p - 0.5;
    if (std::fabs(q) <= 0.425) {
        const double r = 0.180625 - q * q;
        return q * (((a3 * r + a2) * r + a1) * r + a0) / (((b3 * r + b2) * r + b1) * r + 1.0);
    }
    double r = q < 0.0 ? p : 1.0 - p;
    if (r <= 0.0) return 0.0;
    r = std::sqrt(-std::log(r));
    double v;
    if (r <= 5.0) { r -= 1.6; v = (((c3 * r + c2) * r + c1) * r + c0) / ((d2 * r + d1) * r + 1.0); }
    else { r -= 5.0; v = (((e3 * r + e2) * r + e1) * r + e0) / ((f2 * r + f1) * r + 1.0); }
    return q < 0.0 ? -v : v;
}

double b2c_alnorm(double x, bool upper) {
    const double p = 0.398942280444, q = 0.39990348504, r = 0.398942280385;
    const double a1 = 5.75885480458, a2 = 2.62433121679, a3 = 5.92885724438, b1 = -29.8213557807, b2 = 48.6959930692;
    const double c1 = -3.8052e-8, c2 = 3.98064794e-4, c3 = -0.151679116635, c4 = 4.8385912808, c5 = 0.742380924027, c6 = 3.99019417011;
    const double d1 = 1.00000615302, d2 = 1.98615381364, d3 = 5.29330324926, d4 = -15.1508972451, d5 = 30.789933034;
    bool up = upper;
    double z = x, v;
    if (z < 0.0) { up = !up; z = -z; }
    if (z <= 7.0 || (up && z <= 18.66)) {
        const double y = 0.5 * z * z;
        if (z > 1.28) v = r * std::exp(-y) / (z + c1 + d1 / (z + c2 + d2 / (z + c3 + d3 / (z + c4 + d4 / (z + c5 + d5 / (z + c6))))));
        else v = 0.5 - z * (p - q * y / (y + a1 + b1 / (y + a2 + b2 / (y + a3))));
    } else v = 0.0;
    return up ? v : 1.0 - v;
}

// p-value of W for the ascending sample x; the script's special returns: -1, -2.2, -2.3 ("data not qualified"), 0
double b2c_shapiro_wilk(const std::vector<double>& x) {
    static const double c1[] = {0.0, 0.221157, -0.147981, -2.07119, 4.434685, -2.706056};
    static const double c2[] = {0.0, 0.042981, -0.293762, -1.752461, 5.682633, -3.582633};
    static const double c3[] = {0.5440, -0.39978, 0.025054, -0.6714e-3};
    static const double c4[] = {1.3822, -0.77857, 0.062767, -0.0020322};
    static const double c5[] = {-1.5861, -0.31082, -0.083751, 0.0038915};
    static const double c6[] = {-0.4803, -0.082676, 0.0030302};
    static const double g[] = {-2.273, 0.459};
    const double small = 1e-19, pi6 = 1.909859, stqr = 1.047198, sqrth = 0.70711;
    const long n = (long)x.size();
    if (n < 3) return -1;
    const long nn2 = n / 2;
    const double an = (double)n;
    std::vector<double> a(nn2 + 1, 0.0);
    if (n == 3) a[0] = sqrth;
    else {
        const double an25 = an + 0.25;
        double summ2 = 0.0;
        for (long i = 1; i <= nn2; ++i) { a[i - 1] = b2c_ppnd((i - 0.375) / an25); summ2 += a[i - 1] * a[i - 1]; }
        summ2 *= 2.0;
        const double ssumm2 = std::sqrt(summ2), rsn = 1.0 / std::sqrt(an);
        const double a1 = b2c_poly(c1, 6, rsn) - a[0] / ssumm2;
        long i1; double fac;
        if (n > 5) {
            i1 = 3;
            const double a2 = -a[1] / ssumm2 + b2c_poly(c2, 6, rsn);
            fac = std::sqrt((summ2 - 2.0 * a[0] * a[0] - 2.0 * a[1] * a[1]) / (1.0 - 2.0 * a1 * a1 - 2.0 * a2 * a2));
            a[0] = a1; a[1] = a2;
        } else {
            i1 = 2;
            fac = std::sqrt((summ2 - 2.0 * a[0] * a[0]) / (1.0 - 2.0 * a1 * a1));
            a[0] = a1;
        }
        for (long i = i1; i <= nn2; ++i) a[i - 1] = -a[i - 1] / fac;
    }
    const double range = x[n - 1] - x[0];
    if (range < small) return -2.2;
    double xx = x[0] / range, sx = xx, sa = -a[0];
    auto coef = [&](long i, long j) { return (i - j >= 0 ? 1.0 : -1.0) * a[(i <= j ? i : j) - 1]; };
    for (long i = 2, j = n - 1; i <= n; ++i, --j) {
        const double xi = x[i - 1] / range;
        if (xx - xi > small) return -2.3;
        sx += xi;
        if (i != j) sa += coef(i, j);
        xx = xi;
    }
    sa /= n; sx /= n;
    double ssa = 0.0, ssx = 0.0, sax = 0.0;
    for (long i = 1, j = n; i <= n; ++i, --j) {
        const double asa = (i != j) ? coef(i, j) - sa : -sa;
        const double xsx = x[i - 1] / range - sx;
        ssa += asa * asa; ssx += xsx * xsx; sax += asa * xsx;
    }
    const double ssassx = std::sqrt(ssa * ssx);
    const double w1 = (ssassx - sax) * (ssassx + sax) / (ssa * ssx);
    const double w = 1.0 - w1;
    if (n == 3) {
        const double s = std::sqrt(w);
        if (s > 1 || s < -1) return 0.0;
        return pi6 * (std::atan2(s, std::sqrt(1 - s * s)) - stqr);
    }
    double y = std::log(w1), m, s;
    const double lx = std::log(an);
    if (n <= 11) {
        const double gamma = b2c_poly(g, 2, an);
        if (y >= gamma) return small;
        y = -std::log(gamma - y);
        m = b2c_poly(c3, 4, an);
        s = std::exp(b2c_poly(c4, 4, an));
    } else {
        m = b2c_poly(c5, 4, lx);
        s = std::exp(b2c_poly(c6, 3, lx));
    }
    return b2c_alnorm((y - m) / s, true);
}

// Statistics::Descriptive as the script uses it: mean, and s.d. from the running sums (n - 1 in the denominator)
struct B2cStat {
    std::vector<double> data;
    double sum = 0, sumsq = 0;
    void add(double v) { data.push_back(v); sum += v; sumsq += v * v; }
    size_t count() const { return data.size(); }
    double mean() const { return data.empty() ? 0.0 : sum / (double)data.size(); }
    double sd() const {
        const size_t n = data.size();
        if (n < 2) return 0.0;
        const double mu = mean();
        double var = (sumsq - (double)n * mu * mu) / (double)(n - 1);
        return var < 0 ? 0.0 : std::sqrt(var);
    }
};

struct B2cLib {
    bool open = true;            // still in %libs
    bool has_stat = false;       // exists in %insert_stat
    B2cStat insert, readlen;
    uint64_t libpos = 0;
};

// AlnParser::in for format 'sam' (perl/AlnParser.pm:31-124): Maq-style pair flag of a record
int b2c_maq_flag(const Core& c, bool solid) {
    const uint32_t f = c.flag;
    if (f & 0x400) return 0;
    if (!(f & 0x1)) return 0;
    const bool rev = f & 0x10, mrev = f & 0x20;
    if (f & 0x4) return 192;
    if (f & 0x8) return 64;
    if (!(c.mtid >= 0 && c.mtid == c.tid)) return 32;      // samtools prints '=' only for the same reference
    if (f & 0x2) {
        if (solid) return 18;
        return (c.pos < c.mpos) ? (rev ? 20 : 18) : (rev ? 18 : 20);
    }
    if (solid) {
        if (rev != mrev) return mrev ? 8 : 1;
        const bool read1 = f & 0x40;
        if (!rev) return read1 ? (c.pos < c.mpos ? 2 : 4) : (c.pos > c.mpos ? 2 : 4);
        return read1 ? (c.pos > c.mpos ? 2 : 4) : (c.pos < c.mpos ? 2 : 4);
    }
    if (rev == mrev) return mrev ? 8 : 1;
    if ((c.mpos > c.pos && rev) || (c.pos > c.mpos && !rev)) return 4;
    return 2;
}

std::string b2c_header_field(const std::string& line, const char* key) {      // /KEY:(\S+)/ on an @RG line
    size_t p = line.find(key);
    if (p == std::string::npos) return std::string();
    p += strlen(key);
    size_t e = p;
    while (e < line.size() && !isspace((unsigned char)line[e])) ++e;
    return line.substr(p, e - p);
}

void b2c_one_bam(const std::string& path, const bdh_bam2cfg_opts& o, const std::map<std::string, std::string>& rg_lib_file, std::ostringstream& out) {
    MappedFile mf;
    mf.open(path);
    BamData bd;
    bd.path = path;
    bgzf_inflate_all(mf, path, default_threads(), bd.raw);
    bd.parse_header();
    bd.find_records(default_threads());
    // @RG lines (perl/bam2cfg.pl:73-87); -f mappings first, header lines override / add
    std::vector<std::string> rg_order;
    std::map<std::string, std::string> rg_lib(rg_lib_file), rg_platform;
    std::map<std::string, B2cLib> libs;
    for (auto const& kv : rg_lib_file) { libs[kv.second]; if (std::find(rg_order.begin(), rg_order.end(), kv.first) == rg_order.end()) rg_order.push_back(kv.first); }
    {
        std::istringstream hs(bd.text);
        std::string line;
        while (std::getline(hs, line)) {
            if (line.compare(0, 3, "@RG") != 0) continue;
            const std::string id = b2c_header_field(line, "ID:"), lb = b2c_header_field(line, "LB:"), pl = b2c_header_field(line, "PL:");
            if (id.empty()) continue;
            if (std::find(rg_order.begin(), rg_order.end(), id) == rg_order.end()) rg_order.push_back(id);
            rg_lib[id] = lb; rg_platform[id] = pl;
            if (!lb.empty()) libs[lb];
        }
    }
    std::map<std::string, std::map<int, uint64_t>> flag_hist;
    std::map<std::string, uint64_t> flag_all;
    long long expected_max = 0, recordcounter = 0;
    int last_tid = -2; int32_t ppos = 0;
    auto n_open = [&]() { size_t k = 0; for (auto const& kv : libs) k += kv.second.open; return k; };
    auto n_stat = [&]() { size_t k = 0; for (auto const& kv : libs) k += kv.second.has_stat; return k; };
    for (uint64_t off : bd.rec_off) {
        const uint8_t* r = bd.raw.data() + off;
        const uint32_t bs = rd32(r - 4);
        const Core c = read_core(r);
        const size_t open_libs = n_open();
        if (open_libs == 0) {
            if (n_stat() > 0) break;
            libs["NA"]; rg_lib["NA"] = "NA"; rg_platform["NA"] = o.solid ? "solid" : "illumina";
            if (std::find(rg_order.begin(), rg_order.end(), "NA") == rg_order.end()) rg_order.push_back("NA");
        }
        if (expected_max <= 0) expected_max = 3LL * (long long)open_libs * o.n_obs;
        if (recordcounter > expected_max) break;
        if (c.tid != last_tid) ppos = 0;
        last_tid = c.tid;
        if (c.pos + 1 < ppos) throw std::runtime_error("Please sort bam by position");
        ppos = c.pos + 1;
        const uint8_t* aux = r + 32 + c.l_qname + 4 * (size_t)c.n_cigar + (size_t)(c.l_qseq + 1) / 2 + (size_t)c.l_qseq;
        const uint8_t* end = r + bs;
        std::string rg; bool has_rg = false;
        if (const uint8_t* v = aux_get(aux, end, "RG")) if (*v == 'Z') { rg.assign((const char*)v + 1); has_rg = true; }
        std::string platform = o.solid ? "solid" : "illumina";
        if (has_rg) { auto it = rg_platform.find(rg); platform = (it != rg_platform.end() && !it->second.empty()) ? it->second : "illumina"; }
        auto is_int = [](const uint8_t* v) { return v && strchr("cCsSiI", *v) != nullptr; };
        long long qual = c.mapq;
        if (!o.use_mapq) {
            const uint8_t* v = aux_get(aux, end, "Aq");
            if (is_int(v) && aux2i(v) >= 0) qual = aux2i(v);
            else { v = aux_get(aux, end, "AM"); if (is_int(v) && aux2i(v) >= 0) qual = aux2i(v); }
        }
        int flag;
        { const uint8_t* v = aux_get(aux, end, "MF"); flag = (is_int(v) && aux2i(v) >= 0) ? aux2i(v) : b2c_maq_flag(c, strcasestr(platform.c_str(), "solid") != nullptr); }
        const double readlen = c.l_qseq > 0 ? c.l_qseq : 1;      // samtools prints "*" for an absent sequence
        std::string lib;
        if (has_rg) { auto it = rg_lib.find(rg); if (it == rg_lib.end() || it->second.empty()) continue; lib = it->second; }
        else lib = "NA";
        auto lit = libs.find(lib);
        if (lit == libs.end() || !lit->second.open) continue;
        B2cLib& L = lit->second;
        L.readlen.add(readlen);
        if (qual <= o.min_mapq) continue;
        ++recordcounter; ++L.libpos;
        if (has_rg) { ++flag_hist[rg][flag]; ++flag_all[rg]; }
        const double nreads = L.has_stat ? (double)L.insert.count() : 1.0;
        if (nreads / (double)L.libpos < 1e-4) { L.open = false; L.has_stat = false; L.insert = B2cStat(); }   // single-end lane
        if (!((flag == 18 || flag == 20) && c.isize >= 0)) continue;
        L.has_stat = true;
        L.insert.add((double)c.isize);
        if ((long long)L.insert.count() > o.n_obs) L.open = false;
    }
    struct Final { bool ok = false; double mean = 0, sd = 0, stdm = 0, stdp = 0, readlen = 0; size_t num = 0; double sw = 0; };
    std::map<std::string, Final> fin;
    for (auto& kv : libs) {
        B2cLib& L = kv.second;
        if (!L.has_stat) continue;
        const double mean0 = L.insert.mean(), sd0 = L.insert.sd();
        B2cStat kept;
        for (double x : L.insert.data) if (!(x > mean0 + 5 * sd0)) kept.add(x);
        const double mean = kept.mean(), sd = kept.sd();
        if (mean < o.min_mean) continue;
        const double cv = sd / mean;
        if (cv >= o.max_cv) {
            fprintf(stderr, "Coefficient of variation %g in library %s is larger than the cutoff %g, poor quality data, excluding from further analysis.\n", cv, kv.first.c_str(), o.max_cv);
            continue;
        }
        if (kept.count() < 100) continue;
        double stdm = 0, stdp = 0; size_t nm = 0, np = 0;
        for (double x : kept.data) { if (x > mean) { stdp += (x - mean) * (x - mean); ++np; } else { stdm += (x - mean) * (x - mean); ++nm; } }
        Final f;
        f.ok = true; f.mean = mean; f.sd = sd; f.num = kept.count(); f.readlen = L.readlen.mean();
        f.stdm = std::sqrt(stdm / ((double)nm - 1)); f.stdp = std::sqrt(stdp / ((double)np - 1));
        std::vector<double> sorted(kept.data);
        std::sort(sorted.begin(), sorted.end());
        f.sw = b2c_shapiro_wilk(sorted);
        fin[kv.first] = f;
    }
    char buf[512];
    for (auto const& rg : rg_order) {
        auto li = rg_lib.find(rg);
        if (li == rg_lib.end()) continue;
        auto fi = fin.find(li->second);
        if (fi == fin.end() || !fi->second.ok) continue;
        const Final& f = fi->second;
        auto pi = rg_platform.find(rg);
        const std::string platform = (pi != rg_platform.end() && !pi->second.empty()) ? pi->second : "illumina";
        double upper = f.mean + o.cut_sd * f.stdp, lower = f.mean - o.cut_sd * f.stdm;
        if (lower < 0) lower = 0;
        snprintf(buf, sizeof buf, "readlen:%.2f", f.readlen);
        out << "readgroup:" << rg << "\tplatform:" << platform << "\tmap:" << path << "\t" << buf << "\tlib:" << li->second << "\tnum:" << f.num;
        snprintf(buf, sizeof buf, "\tlower:%.2f\tupper:%.2f\tmean:%.2f\tstd:%.2f", lower, upper, f.mean, f.sd);
        out << buf;
        if (f.sw > 0) { snprintf(buf, sizeof buf, "\tSWnormality:%.2f", std::log(f.sw) / std::log(10.0)); out << buf; }
        else if (f.sw == -1) out << "\tSWnormality:data not qualified -1";
        else if (f.sw == -2.1) out << "\tSWnormality:data not qualified -2.1";
        else if (f.sw == -2.2) out << "\tSWnormality:data not qualified -2.2";
        else if (f.sw == -2.3) out << "\tSWnormality:data not qualified -2.3";
        else if (f.sw == 0) out << "\tSWnormality:minus infinity";
        if (o.flag_hist) {
            out << "\tflag:";
            std::vector<std::pair<std::string, uint64_t>> fl;      // `sort keys`: the flags as strings
            for (auto const& kv : flag_hist[rg]) fl.push_back({std::to_string(kv.first), kv.second});
            std::sort(fl.begin(), fl.end());
            for (auto const& kv : fl) { snprintf(buf, sizeof buf, "%s(%.2f%%)", kv.first.c_str(), (double)kv.second * 100 / (double)flag_all[rg]); out << buf; }
            out << flag_all[rg];
        }
        out << "\texe:samtools view\n";
    }
}

}  // namespace
}  // namespace bdh

extern "C" {

void bdh_bam2cfg_defaults(bdh_bam2cfg_opts* o) {
    if (!o) return;
    o->min_mapq = 35; o->n_obs = 10000; o->cut_sd = 4; o->min_mean = 50; o->max_cv = 1; o->use_mapq = 0; o->solid = 0; o->flag_hist = 0; o->rg_lib_file = nullptr;
}

int64_t bdh_bam2cfg(const char* const* bam_paths, int nbam, const bdh_bam2cfg_opts* opts, char* buf, int64_t cap, char* err, int errcap) {
    try {
        bdh_bam2cfg_opts o;
        if (opts) o = *opts; else bdh_bam2cfg_defaults(&o);
        std::map<std::string, std::string> rg_lib_file;
        if (o.rg_lib_file && *o.rg_lib_file) {
            std::ifstream f(o.rg_lib_file);
            if (!f) throw std::runtime_error(std::string("unable to open ") + o.rg_lib_file);
            std::string rg, lib;
            while (f >> rg >> lib) rg_lib_file[rg] = lib;
        }
        std::ostringstream out;
        for (int i = 0; i < nbam; ++i) bdh::b2c_one_bam(bam_paths[i], o, rg_lib_file, out);
        const std::string s = out.str();
        if (buf && (int64_t)s.size() + 1 <= cap) memcpy(buf, s.c_str(), s.size() + 1);
        return (int64_t)s.size();
    } catch (std::exception const& e) {
        if (err && errcap > 0) { strncpy(err, e.what(), errcap - 1); err[errcap - 1] = 0; }
        return -1;
    }
}

}  // extern "C"
