// bam2cfg configuration file -> library table.
// Host-side mirror of BamConfig / BamConfigEntry (reference src/lib/io/BamConfig.cpp:19-122,
// BamConfigEntry.cpp:31-86): same grammar (tab-separated key:value fields, keys matched by the
// legacy case-insensitive suffix patterns, parsing stops at the first empty line), same library
// and bam ordering (sorted names), same initial window rule.
#include "host.hpp"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace bdh {

enum Field { BAM_FILE, LIBRARY_NAME, READ_GROUP, MEAN, STDDEV, READ_LENGTH, UPPER, LOWER, MIN_MAP_QUAL, SAMPLE, UNKNOWN };

namespace {

bool is_word(char c) { return std::isalnum((unsigned char)c) || c == '_'; }

bool ieq(const std::string& s, size_t at, const char* lit) {
    size_t n = strlen(lit);
    if (at + n > s.size()) return false;
    for (size_t i = 0; i < n; ++i)
        if (std::tolower((unsigned char)s[at + i]) != lit[i]) return false;
    return true;
}

bool all_word_from(const std::string& s, size_t at) {
    for (size_t i = at; i < s.size(); ++i) if (!is_word(s[i])) return false;
    return true;
}

// regex_search(key, /<lit>\w*$/i)
bool match_prefix_tail(const std::string& key, const char* lit) {
    size_t n = strlen(lit);
    for (size_t i = 0; i + n <= key.size(); ++i)
        if (ieq(key, i, lit) && all_word_from(key, i + n)) return true;
    return false;
}

// regex_search(key, /map\w*qual\w*$/i)
bool match_mapqual(const std::string& key) {
    for (size_t i = 0; i + 3 <= key.size(); ++i) {
        if (!ieq(key, i, "map")) continue;
        for (size_t j = i + 3; j + 4 <= key.size(); ++j) {
            if (ieq(key, j, "qual") && all_word_from(key, j + 4)) return true;
            if (!is_word(key[j])) break;
        }
    }
    return false;
}

}  // namespace

// The reference iterates a flat_map<regex, Field>, i.e. patterns ordered by their text
// (BamConfigEntry.cpp:43-54); first hit wins.
Field translate_token(const std::string& key) {
    if (key.size() >= 5 && ieq(key, key.size() - 5, "group")) return READ_GROUP;               // group$
    if (match_prefix_tail(key, "lib")) return LIBRARY_NAME;                                   // lib\w*$
    if (match_prefix_tail(key, "low")) return LOWER;                                          // low\w*$
    if (key.size() >= 3 && ieq(key, key.size() - 3, "map")) return BAM_FILE;                   // map$
    if (match_mapqual(key)) return MIN_MAP_QUAL;                                              // map\w*qual\w*$
    if (match_prefix_tail(key, "mean")) return MEAN;                                          // mean\w*$
    if (match_prefix_tail(key, "readlen")) return READ_LENGTH;                                // readlen\w*$
    if (match_prefix_tail(key, "samp")) return SAMPLE;                                        // samp\w*$
    if (match_prefix_tail(key, "std")) return STDDEV;                                         // std\w*$
    if (match_prefix_tail(key, "upp")) return UPPER;                                          // upp\w*$
    return UNKNOWN;
}

namespace {

float to_float(const std::string& s) {  // boost::lexical_cast<float>: whole string must parse
    if (s.empty() || std::isspace((unsigned char)s[0])) throw std::runtime_error("bad lexical cast: source type value could not be interpreted as target");
    char* end = 0;
    float v = strtof(s.c_str(), &end);
    if (*end) throw std::runtime_error("bad lexical cast: source type value could not be interpreted as target");
    return v;
}
int to_int(const std::string& s) {
    if (s.empty() || std::isspace((unsigned char)s[0])) throw std::runtime_error("bad lexical cast: source type value could not be interpreted as target");
    char* end = 0;
    long v = strtol(s.c_str(), &end, 10);
    if (*end) throw std::runtime_error("bad lexical cast: source type value could not be interpreted as target");
    return (int)v;
}

struct TmpLib {
    std::string name, bam;
    float mean = 0, stddev = 0, upper = 0, lower = 0, readlen = 0;
    int mqual = -1;
    bool operator!=(TmpLib const& o) const {
        return name != o.name || bam != o.bam || mean != o.mean || stddev != o.stddev || upper != o.upper
            || lower != o.lower || readlen != o.readlen || mqual != o.mqual;
    }
};

}  // namespace

Config Config::parse(std::istream& in, int cut_sd) {
    Config cfg;
    int window = 100000000;  // DEFAULT_MAX_READ_WINDOW_SIZE (BamConfig.cpp:12)
    std::map<std::string, TmpLib> tmp;
    std::map<std::string, std::string> bam_library;
    std::string line;
    size_t line_num = 0;
    while (std::getline(in, line)) {
        ++line_num;
        if (line.empty()) break;
        std::map<Field, std::string> d;
        size_t start = 0;
        while (true) {
            size_t tab = line.find('\t', start);
            std::string f = line.substr(start, tab == std::string::npos ? std::string::npos : tab - start);
            size_t colon = f.find(':');
            if (colon != std::string::npos) {
                Field fn = translate_token(f.substr(0, colon));
                if (fn != UNKNOWN) d[fn] = f.substr(colon + 1);
            }
            if (tab == std::string::npos) break;
            start = tab + 1;
        }
        TmpLib L;
        std::string readgroup;
        if (d.count(LIBRARY_NAME)) L.name = d[LIBRARY_NAME];
        else if (d.count(SAMPLE)) L.name = d[SAMPLE];
        if (!d.count(BAM_FILE)) {
            std::ostringstream m;
            m << "Required field 'map' not found in config at line " << line_num << "!";
            throw std::runtime_error(m.str());
        }
        L.bam = d[BAM_FILE];
        readgroup = d.count(READ_GROUP) ? d[READ_GROUP] : L.name;
        cfg.readgroup_library[readgroup] = L.name;
        bam_library[L.bam] = L.name;
        if (d.count(READ_LENGTH)) L.readlen = to_float(d[READ_LENGTH]);
        if (d.count(MIN_MAP_QUAL)) L.mqual = to_int(d[MIN_MAP_QUAL]);
        bool have_mean = d.count(MEAN), have_std = d.count(STDDEV), have_lower = d.count(LOWER), have_upper = d.count(UPPER);
        if (have_mean) L.mean = to_float(d[MEAN]);
        if (have_std) L.stddev = to_float(d[STDDEV]);
        if (have_lower) L.lower = to_float(d[LOWER]);
        if (have_upper) L.upper = to_float(d[UPPER]);
        if (have_mean && have_std && (!have_upper || !have_lower)) {
            L.upper = L.mean + L.stddev * cut_sd;
            L.lower = L.mean - L.stddev * cut_sd;
            L.lower = L.lower > 0 ? L.lower : 0;
        }
        auto ins = tmp.insert(std::make_pair(L.name, L));
        if (!ins.second && ins.first->second != L) {
            fprintf(stderr, "WARNING: at line %zu, library %s overwritten!\n", line_num, L.name.c_str());
            ins.first->second = L;
        }
        int t = L.mean - L.readlen * 2;
        window = std::min(window, t);
    }
    for (auto const& kv : bam_library) cfg.bam_files.push_back(kv.first);
    cfg.first_bam_library = bam_library.empty() ? std::string() : bam_library.begin()->second;
    for (auto const& kv : tmp) {
        TmpLib const& L = kv.second;
        bdk_lib o;
        o.mean_insertsize = L.mean; o.std_insertsize = L.stddev; o.uppercutoff = L.upper; o.lowercutoff = L.lower;
        o.readlens = L.readlen; o.min_mapping_quality = L.mqual;
        auto it = std::find(cfg.bam_files.begin(), cfg.bam_files.end(), L.bam);
        if (it == cfg.bam_files.end())
            throw std::runtime_error("Bam file '" + L.bam + "' referenced by library '" + L.name + "' but not found in bam list!");
        o.bam_index = int(it - cfg.bam_files.begin());
        cfg.lib_index[L.name] = (int)cfg.libs.size();
        cfg.libs.push_back(o);
        cfg.lib_names.push_back(L.name);
    }
    cfg.window = std::max(window, 50);
    return cfg;
}

int Config::rg_lib(const std::string& rg) const {
    auto it = readgroup_library.find(rg);
    const std::string& lib = it != readgroup_library.end() ? it->second : first_bam_library;
    if (lib.empty()) return -1;  // AlignmentSource.hpp:59: flag stays NA, lib index unset
    auto li = lib_index.find(lib);
    return li == lib_index.end() ? -1 : li->second;
}

}  // namespace bdh

// ---- C ABI ------------------------------------------------------------------------------------
static void set_err(char* err, int cap, const char* msg) {
    if (err && cap > 0) { strncpy(err, msg, cap - 1); err[cap - 1] = 0; }
}

extern "C" {

bdh_config* bdh_config_parse(const char* text, int cut_sd, char* err, int errcap) {
    try {
        std::istringstream in(text);
        bdh_config* c = new bdh_config;
        c->cfg = bdh::Config::parse(in, cut_sd);
        return c;
    } catch (std::exception const& e) { set_err(err, errcap, e.what()); return 0; }
}

bdh_config* bdh_config_load(const char* path, int cut_sd, char* err, int errcap) {
    try {
        std::ifstream in(path);
        // the reference does not check the stream: an unreadable config yields an empty BamConfig
        bdh_config* c = new bdh_config;
        c->cfg = bdh::Config::parse(in, cut_sd);
        return c;
    } catch (std::exception const& e) { set_err(err, errcap, e.what()); return 0; }
}

void bdh_config_free(bdh_config* c) { delete c; }
int bdh_config_nlib(const bdh_config* c) { return (int)c->cfg.libs.size(); }
int bdh_config_nbam(const bdh_config* c) { return (int)c->cfg.bam_files.size(); }
int bdh_config_window(const bdh_config* c) { return c->cfg.window; }
const bdk_lib* bdh_config_libs(const bdh_config* c) { return c->cfg.libs.data(); }
const char* bdh_config_lib_name(const bdh_config* c, int i) { return c->cfg.lib_names[i].c_str(); }
const char* bdh_config_bam_name(const bdh_config* c, int i) { return c->cfg.bam_files[i].c_str(); }
int bdh_config_rg_lib(const bdh_config* c, const char* rg) { return c->cfg.rg_lib(rg); }
/* BamConfigEntry::translate_token (BamConfigEntry.cpp:31-59); returns the Field ordinal, 10 = UNKNOWN */
int bdh_config_translate_token(const char* key) { return (int)bdh::translate_token(key); }

}  // extern "C"
