// -g BED and -d FASTQ dumps of the SV-supporting reads.
// Mirrors BedWriter::write (reference src/lib/breakdancer/BedWriter.cpp:21-56),
// BreakDancer::dump_fastq (BreakDancer.cpp:514-534), FastqWriter (FastqWriter.cpp:22-46) and
// Alignment::to_fastq (Alignment.cpp:66-84). The GPU reports, for every anomalous read, which SV
// call consumed it; names, bases and qualities come from the raw records kept by the decoder.
#include "host.hpp"

#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <unordered_map>

namespace bdh {

void write_support_reads(bdk_ctx* ctx, const bdh_stream* stream, const bdk_params& p, const bdk_result& res,
                         const std::vector<std::string>& lib_names, const std::vector<std::string>& tid_names,
                         std::ostream* bed, const std::string& fastq_prefix) {
    const bdk_aread* ar = nullptr; const int32_t* rr = nullptr; const int32_t* sv_of = nullptr;
    uint64_t A = 0, A2 = 0;
    if (bdk_get_areads(ctx, &ar, &rr, &A) != 0 || bdk_get_support(ctx, &sv_of, &A2) != 0)
        throw std::runtime_error(bdk_last_error(ctx));
    // FASTQ streams: every library gets its two files up front, opened in append mode
    std::map<std::string, std::unique_ptr<std::ofstream>> fq;
    auto fq_open = [&](const std::string& lib, bool read1) -> std::ofstream& {
        std::string path = fastq_prefix + "." + lib + "." + (read1 ? "1" : "2") + ".fastq";
        auto it = fq.find(path);
        if (it == fq.end()) it = fq.emplace(path, std::unique_ptr<std::ofstream>(new std::ofstream(path.c_str(), std::ofstream::app))).first;
        if (!*it->second) throw std::runtime_error("Failed to open fastq file '" + path + "' for writing");
        return *it->second;
    };
    if (!fastq_prefix.empty())
        for (int l = 0; l < p.nlib; ++l) { fq_open(lib_names[l], true); fq_open(lib_names[l], false); }
    // reads per SV, in stream order
    std::vector<std::vector<uint32_t>> per_sv(res.n_sv);
    for (uint64_t j = 0; j < A; ++j)
        if (sv_of[j] >= 0 && (uint64_t)sv_of[j] < res.n_sv) per_sv[sv_of[j]].push_back((uint32_t)j);
    std::vector<char> buf(1 << 16);
    for (uint64_t i = 0; i < res.n_sv; ++i) {
        const bdk_sv& sv = res.sv[i];
        // support_reads: for each pair in order of its second-seen mate: (second, first)
        std::vector<uint32_t> support;
        std::unordered_map<uint64_t, uint32_t> seen;
        for (uint32_t j : per_sv[i]) {
            auto it = seen.find(ar[j].qid);
            if (it == seen.end()) seen[ar[j].qid] = j;
            else { support.push_back(j); support.push_back(it->second); seen.erase(it); }
        }
        const std::string& seq_name = tid_names[sv.chr[0]];
        const char* type = sv_type_name(sv.flag, p.illumina_long_insert != 0);
        if (bed) {
            *bed << "track name=" << seq_name << "_" << sv.pos[0] << "_" << type << "_" << sv.diffspan
                 << "\tdescription=\"BreakDancer" << " " << seq_name << " " << sv.pos[0] << " " << type << " " << sv.diffspan
                 << "\"\tuseScore=0\n";
        }
        std::map<uint64_t, int> pairing;
        for (uint32_t j : support) {
            const bdk_aread& y = ar[j];
            const int flag = (int)(y.meta & 0xF), rev = (int)((y.meta >> 4) & 1), lib = (int)((y.meta >> 8) & 0xFF), q = (int)((y.meta >> 16) & 0xFF);
            if (y.qlen <= 0 || flag != sv.flag) continue;   // has_sequence() && bdflag == flag
            if (bed) {
                const int aln_end = y.pos + y.qlen;
                if (strncmp("chr", seq_name.c_str(), 3) != 0) *bed << "chr";
                *bed << tid_names[y.tid] << "\t" << y.pos << "\t" << aln_end << "\t" << bdh_stream_qname(stream, y.record) << "|" << lib_names[lib]
                     << "\t" << q * 10 << "\t" << rev << "\t" << y.pos << "\t" << aln_end << "\t" << (rev ? "255,0,0" : "0,0,255") << "\n";
            }
            if (!fastq_prefix.empty()) {
                // the first read seen of a pair goes to file 2, the second to file 1 (BreakDancer.cpp:526-530)
                const bool is_read1 = pairing.count(y.qid) != 0;
                int n = bdh_stream_fastq(stream, y.record, buf.data(), (int)buf.size());
                if (n < 0) throw std::runtime_error("raw record not available for FASTQ dump");
                fq_open(lib_names[lib], is_read1).write(buf.data(), n);
                pairing[y.qid] = 1;
            }
        }
    }
}

}  // namespace bdh
