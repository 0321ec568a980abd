// hostsim.cpp -- runs the DEVICE logic of the hot path (breakdancer_b200/csrc/bdk_logic.h,
// bdk_finalize.h: the __host__ __device__ functions the CUDA kernels call) on the host, stage by
// stage in the same decomposition the GPU pipeline uses:
//   K1 classify + compaction + proper-pair prefix counts    K2 break flags / candidates / regions
//   K3 mate join + link sort + run-length                   K4 components + connection walk + score
// with the parallel mechanics (scans, radix sort, hash join, atomics) replaced by trivial
// sequential loops.  TEST HARNESS ONLY: built by tests/conftest.py into tests/_build/, never
// linked into libbdk.so -- it lets `pytest -m "not gpu"` check the decomposition (telescoped
// counts, time-stamp rules, per-component independence) against the oracle without a GPU.
#include "../../breakdancer_b200/csrc/bdk_logic.h"
#include "../../breakdancer_b200/csrc/bdk_finalize.h"
#include "../../oracle/bdo_api.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <unordered_map>
#include <vector>

using namespace bdk;

extern "C" int hostsim_run(const bdk_params* pp, const bdk_soa* c, uint64_t n, bdo_output* out) {
    const bdk_params& p = *pp;
    int nkey = nkey_of(p), nlib = p.nlib;
    std::vector<LibDev> libs = make_libdev(p);
    std::vector<float> lib_mean = make_lib_mean(p);
    ClassifyOpts co{p.max_sd, p.transchr_rearrange, p.illumina_long_insert};

    // ---- K1 -----------------------------------------------------------------------------------
    SummaryAcc acc;
    acc.rg_sproper.assign(p.nrg, 0);
    acc.hist.assign((size_t)nlib * BDK_NUM_FLAGS, 0);
    acc.first.assign((size_t)p.nbam * p.ntid, ~0ull);
    acc.last.assign((size_t)p.nbam * p.ntid, 0);
    std::vector<bdk_aread> ar;
    std::vector<std::vector<uint32_t>> P(nkey);
    std::vector<uint32_t> run(nkey, 0);
    std::vector<uint8_t> rec_class(n, 255);
    for (uint64_t i = 0; i < n; ++i) {
        int rg = c->rgid[i];
        if (rg >= p.nrg || p.rg_lib[rg] < 0) return BDK_ERR_DATA;
        int lib = p.rg_lib[rg], bam = p.rg_bam[rg];
        uint32_t cr = classify_record(c->pos[i], c->mpos[i], c->tid[i], c->mtid[i], c->isize[i], c->flag[i], c->mapq[i], libs[lib], co);
        uint64_t key = (i << 32) | (uint32_t)c->pos[i];
        size_t bt = (size_t)bam * p.ntid + c->tid[i];
        if (acc.first[bt] == ~0ull) acc.first[bt] = key;   // device: atomicMin / atomicMax of the same key
        acc.last[bt] = key;
        if (cr & CR_SPROPER) ++acc.rg_sproper[rg];
        int hf = (cr >> CR_HIST_SHIFT) & 0xF;
        if (hf) ++acc.hist[lib * BDK_NUM_FLAGS + hf];
        if (cr & CR_KEPT) rec_class[i] = cr & CR_FLAG_MASK;
        if (cr & CR_MPROPER) ++run[libs[lib].key];
        if (cr & CR_ANOM) {
            bdk_aread a;
            a.pos = c->pos[i]; a.tid = c->tid[i]; a.qlen = c->qlen[i];
            a.abs_isize = c->isize[i] < 0 ? -c->isize[i] : c->isize[i];
            a.meta = make_meta(cr, lib, c->mapq[i]); a.record = (uint32_t)i; a.qid = c->qid[i];
            ar.push_back(a);
            for (int k = 0; k < nkey; ++k) P[k].push_back(run[k]);
        }
    }
    int64_t A = (int64_t)ar.size();
    bdk_summary_t S;
    std::vector<float> density;
    finalize_summary(p, acc, n, A, &S, &density);
    int window = S.window;

    // ---- K2 -----------------------------------------------------------------------------------
    std::vector<int32_t> read_cand(A), read_region(A, -1);
    std::vector<int64_t> cand_first;
    for (int64_t j = 0; j < A; ++j) {
        bool brk = j == 0 || k2_is_break(ar[j - 1].tid, ar[j - 1].pos, ar[j].tid, ar[j].pos, window);
        if (brk) cand_first.push_back(j);
        read_cand[j] = (int32_t)cand_first.size() - 1;
    }
    int ncand = (int)cand_first.size();
    std::vector<int32_t> cand_maxlen(ncand), cand_regs(ncand);
    std::vector<RegionRec> reg;
    int dummy = (A > 0) ? dummy_region_of(p) : 0;
    if (dummy) { RegionRec d; d.tid = -1; d.start = -1; d.end = -1; d.fwd = d.rev = 0; d.first_read = 0; d.n_reads = 0; d.stored = 0; d.cand = -1; reg.push_back(d); }
    std::vector<uint8_t> alive(A, 0);
    for (int cidx = 0; cidx < ncand; ++cidx) {
        int64_t s = cand_first[cidx], e = (cidx + 1 < ncand ? cand_first[cidx + 1] : A) - 1;
        CandAgg g = k2_cand_aggregate(ar.data(), s, e, A);
        cand_maxlen[cidx] = g.maxlen;
        if (k2_accept(ar[s].pos, ar[e].pos, g, p.min_len, p.seq_coverage_lim)) {
            RegionRec R;
            R.tid = ar[s].tid; R.start = ar[s].pos; R.end = ar[e].pos; R.fwd = g.fwd; R.rev = g.rev;
            R.first_read = (int32_t)s; R.n_reads = (int32_t)(e - s + 1);
            int valid = p.chr_restricted ? g.nonctx : R.n_reads;
            R.stored = valid >= p.min_read_pair; R.cand = cidx;
            for (int64_t j = s; j <= e; ++j) { read_region[j] = (int32_t)reg.size(); alive[j] = R.stored; }
            reg.push_back(R);
        }
        cand_regs[cidx] = (int32_t)reg.size();
    }
    int nreg = (int)reg.size();
    int period = period_of(p);

    // ---- K3: mate join, links, sort + run-length ------------------------------------------------
    std::vector<int32_t> mate(A, -1);
    {
        std::unordered_map<uint64_t, int32_t> seen;
        for (int64_t j = 0; j < A; ++j) {
            auto it = seen.find(ar[j].qid);
            if (it == seen.end()) seen[ar[j].qid] = (int32_t)j;
            else { mate[j] = it->second; mate[it->second] = (int32_t)j; }
        }
    }
    std::vector<uint64_t> links;
    for (int64_t y = 0; y < A; ++y) {
        int x = mate[y];
        if (x >= 0 && x < y && read_region[x] >= 0 && read_region[y] >= 0)
            links.push_back(((uint64_t)(uint32_t)read_region[x] << 32) | (uint32_t)read_region[y]);
    }
    std::sort(links.begin(), links.end());
    struct UEdge { int r0, r1, w; };
    std::vector<UEdge> ue;
    for (size_t i = 0; i < links.size();) {
        size_t j = i;
        while (j < links.size() && links[j] == links[i]) ++j;
        ue.push_back({(int)(links[i] >> 32), (int)(links[i] & 0xffffffffu), (int)(j - i)});
        i = j;
    }
    // ---- K4 (closed form, bdk_logic.h "second formulation") ------------------------------------------------------
    // strong bit per read, per-read info
    std::unordered_map<uint64_t, int> weight;
    for (auto const& e : ue) weight[((uint64_t)(uint32_t)e.r0 << 32) | (uint32_t)e.r1] = e.w;
    std::vector<ReadInfo2> ri((size_t)std::max<int64_t>(A, 1));
    for (int64_t j = 0; j < A; ++j) {
        bool strong = false;
        const int m = mate[j];
        if (m >= 0 && read_region[j] >= 0 && read_region[m] >= 0) {
            const int a = std::min(read_region[j], read_region[m]), b2 = std::max(read_region[j], read_region[m]);
            strong = weight[((uint64_t)(uint32_t)a << 32) | (uint32_t)b2] >= p.min_read_pair;
        }
        ri[j] = k4n_make_read_info(ar.data(), mate.data(), read_region.data(), read_cand.data(), reg.data(), cand_regs.data(), period, (int)j, strong);
    }
    K4N KS;
    KS.ri = ri.data(); KS.ar = ar.data(); KS.reg = reg.data(); KS.nreg = nreg; KS.period = period;
    KS.chr_restricted = p.chr_restricted; KS.min_read_pair = p.min_read_pair;
    SoloTeam T;
    // fixed point of the deletion table. Starting table: "cleared in the last window it is active in" (or an empty table),
    // sweeps in Jacobi order (every region against the previous table) or in place in DESCENDING order -- nothing may depend
    // on the order.
    std::vector<int32_t> del(nreg + 1, K4_NEVER), del_next(nreg + 1, K4_NEVER);
    std::vector<uint8_t> dirty(nreg + 1, 1), dirty_next(nreg + 1, 0);
    if (!getenv("HOSTSIM_NO_GUESS"))
        for (int v = 0; v < nreg; ++v) {
            int wl = -1;
            for (int j = reg[v].first_read; j < reg[v].first_read + reg[v].n_reads; ++j)
                if (ri[j].mate_region >= 0) wl = std::max(wl, std::max(v, ri[j].mate_region) / period);
            if (wl >= 0 && v != nreg - 1) del[v] = wl;
        }
    const bool in_place = getenv("HOSTSIM_INPLACE") != nullptr;
    int sweeps = 0;
    for (;; ++sweeps) {
        if (sweeps > 1000000) return -101;
        int nchanged = 0;
        if (!in_place) del_next = del;
        for (int v = nreg - 1; v >= 0; --v) {
            if (!dirty[v]) continue;
            const int d = k4n_region_deletion(T, KS, del.data(), v);
            const int old = del[v];
            (in_place ? del : del_next)[v] = d;
            if (d != old) {
                ++nchanged;
                for (int j = reg[v].first_read; j < reg[v].first_read + reg[v].n_reads; ++j)
                    if (ri[j].mate_region >= 0 && ri[j].mate_region != v) dirty_next[ri[j].mate_region] = 1;
            }
        }
        if (!in_place) del.swap(del_next);
        dirty.swap(dirty_next);
        std::fill(dirty_next.begin(), dirty_next.end(), 0);
        if (getenv("HOSTSIM_VERBOSE")) fprintf(stderr, "hostsim: sweep %d -> %d regions changed\n", sweeps, nchanged);
        if (!nchanged) break;
    }
    std::vector<uint8_t> deleted(nreg + 1, 0);
    for (int v = 0; v < nreg; ++v) deleted[v] = del[v] != K4_NEVER;
    // followed edges per window, directed copies sorted by (window, src, dst); slots in window order
    struct WE { int win, src, dst; };
    std::vector<WE> we;
    for (auto const& e : ue) {
        if (e.w < p.min_read_pair) continue;
        const int win = e.r1 / period;
        we.push_back({win, e.r0, e.r1});
        if (e.r0 != e.r1) we.push_back({win, e.r1, e.r0});
    }
    std::sort(we.begin(), we.end(), [](const WE& a, const WE& b) { return a.win != b.win ? a.win < b.win : (a.src != b.src ? a.src < b.src : a.dst < b.dst); });
    int nrow_cap = 0;
    for (auto const& x : we) nrow_cap += x.src <= x.dst;
    std::vector<int32_t> c1(nreg + 1, K4_NEVER);
    for (int v = 0; v < nreg; ++v) c1[v] = k4n_first_call(T, KS, del.data(), v);
    std::vector<uint32_t> Pflat((size_t)nkey * std::max<int64_t>(A, 1));
    for (int k = 0; k < nkey; ++k) for (int64_t d = 0; d < A; ++d) Pflat[(size_t)d * nkey + k] = P[k][d];   // [A][nkey] like the device
    std::vector<uint8_t> row_emit(nrow_cap + 1, 0);
    std::vector<int32_t> sv_of_read(A, -1), row_lib_count((size_t)(nrow_cap + 1) * nlib), row_lib_span((size_t)(nrow_cap + 1) * nlib);
    std::vector<uint32_t> row_cn_count((size_t)(nrow_cap + 1) * nkey);
    std::vector<float> row_cn((size_t)(nrow_cap + 1) * nkey);
    std::vector<bdk_sv> rows(nrow_cap + 1);
    {
        std::vector<SEdge> se; std::vector<uint8_t> fl; std::vector<int32_t> queue;
        int slot0 = 0;
        // windows in DESCENDING order on purpose (they are independent); slot bases are the counts of the earlier windows
        std::vector<std::pair<size_t, size_t>> ranges;
        for (size_t i = 0; i < we.size();) { size_t j = i; while (j < we.size() && we[j].win == we[i].win) ++j; ranges.push_back({i, j}); i = j; }
        std::vector<int> base(ranges.size() + 1, 0);
        for (size_t k = 0; k < ranges.size(); ++k) { int u = 0; for (size_t t = ranges[k].first; t < ranges[k].second; ++t) u += we[t].src <= we[t].dst; base[k + 1] = base[k] + u; }
        for (size_t k = ranges.size(); k-- > 0;) {
            const size_t i = ranges[k].first, j = ranges[k].second;
            se.clear(); for (size_t t = i; t < j; ++t) { SEdge x; x.src = we[t].src; x.dst = we[t].dst; se.push_back(x); }
            fl.assign(j - i, 0); queue.assign(j - i + 2, 0);
            slot0 = base[k];
            for (int k = (int)(j - i) - 1; k >= 0; --k) k4n_window_prepare(del.data(), c1.data(), se.data(), (int)(j - i), fl.data(), we[i].win, k);
            const int used = k4n_window_calls(se.data(), (int)(j - i), fl.data(), queue.data(), we[i].win, slot0, rows.data(), row_emit.data());
            if (used > base[k + 1] - base[k]) return -100;
        }
    }
    K4NOut KO{sv_of_read.data(), rows.data(), row_lib_count.data(), row_lib_span.data(), row_emit.data(), nlib};
    for (int r = nrow_cap - 1; r >= 0; --r) if (row_emit[r] & K4_ROW_CALL) k4n_call(T, KS, KO, r);
    K4Static SS;
    memset(&SS, 0, sizeof SS);
    SS.ar = ar.data(); SS.reg = reg.data(); SS.P = Pflat.data(); SS.cand_maxlen = cand_maxlen.data(); SS.lib_mean = lib_mean.data();
    SS.hist = acc.hist.data(); SS.density = density.data(); SS.A = (uint64_t)A; SS.nreg = nreg; SS.ncand = ncand;
    SS.period = period; SS.nkey = nkey; SS.nlib = nlib; SS.chr_restricted = p.chr_restricted;
    SS.min_read_pair = p.min_read_pair; SS.score_threshold = p.score_threshold; SS.fisher = p.fisher;
    SS.covered_ref_len = S.covered_ref_len;
    K4Mut KM;
    memset(&KM, 0, sizeof KM);
    KM.rows = rows.data(); KM.row_lib_count = row_lib_count.data(); KM.row_lib_span = row_lib_span.data();
    KM.row_cn_count = row_cn_count.data(); KM.row_cn = row_cn.data(); KM.row_emit = row_emit.data();
    if (getenv("HOSTSIM_VERBOSE")) fprintf(stderr, "hostsim: %d regions, %zu edges, %d call slots, %d sweeps\n", nreg, ue.size(), nrow_cap, sweeps + 1);
    for (int r = 0; r < nrow_cap; ++r) if (row_emit[r] == K4_ROW_PENDING) k4_score_row(SS, KM, r);
    // output order = slot order: slots are numbered window by window, and inside a window in the order of the calls
    std::vector<int> order;
    for (int r = 0; r < nrow_cap; ++r) if (row_emit[r] == K4_ROW_EMIT) order.push_back(r);

    // ---- pack outputs ---------------------------------------------------------------------------
    memset(out, 0, sizeof(*out));
    out->summary = S; out->nkey = nkey;
    size_t ns = order.size();
    out->n_sv = ns;
    out->sv = (bdk_sv*)calloc(ns + 1, sizeof(bdk_sv));
    out->lib_count = (int32_t*)calloc(ns * nlib + 1, 4);
    out->cn_count = (uint32_t*)calloc(ns * nkey + 1, 4);
    out->copy_number = (float*)calloc(ns * nkey + 1, 4);
    std::vector<int> slot_to_order(nrow_cap + 1, -1);
    for (size_t i = 0; i < ns; ++i) {
        int r = order[i];
        slot_to_order[r] = (int)i;
        out->sv[i] = rows[r]; out->sv[i].order = (int)i;
        memcpy(out->lib_count + i * nlib, &row_lib_count[(size_t)r * nlib], nlib * 4);
        memcpy(out->cn_count + i * nkey, &row_cn_count[(size_t)r * nkey], nkey * 4);
        memcpy(out->copy_number + i * nkey, &row_cn[(size_t)r * nkey], nkey * 4);
    }
    out->n_regions = nreg;
    out->regions = (bdk_region*)calloc(nreg + 1, sizeof(bdk_region));
    out->region_alive = (uint8_t*)calloc(nreg + 1, 1);
    for (int r = 0; r < nreg; ++r) {
        bdk_region& o = out->regions[r];
        o.tid = reg[r].tid; o.start = reg[r].start; o.end = reg[r].end; o.fwd = reg[r].fwd; o.rev = reg[r].rev;
        o.first_read = reg[r].first_read; o.n_reads = reg[r].n_reads; o.stored = reg[r].stored; o.window = r / period;
        out->region_alive[r] = !deleted[r];
    }
    out->n_areads = A;
    out->areads = (bdk_aread*)calloc(A + 1, sizeof(bdk_aread));
    if (A) memcpy(out->areads, ar.data(), A * sizeof(bdk_aread));
    out->aread_region = (int32_t*)calloc(A + 1, 4);
    if (A) memcpy(out->aread_region, read_region.data(), A * 4);
    out->sv_of_read = (int32_t*)calloc(A + 1, 4);
    for (int64_t j = 0; j < A; ++j) out->sv_of_read[j] = sv_of_read[j] >= 0 ? slot_to_order[sv_of_read[j]] : -1;
    out->rec_class = (uint8_t*)malloc(n + 1);
    memcpy(out->rec_class, rec_class.data(), n);
    out->support_off = (uint64_t*)calloc(ns + 1, 8);
    out->support = (uint32_t*)calloc(1, 4);
    out->n_flush = nreg / period + 1;
    return 0;
}

extern "C" double hostsim_poisson_logsf(double lambda, int k) { return poisson_log_sf(lambda, k); }
extern "C" double hostsim_gamma_q(double a, double x) { return gamma_q_d(a, x); }
extern "C" uint32_t hostsim_classify(int32_t pos, int32_t mpos, int32_t tid, int32_t mtid, int32_t isize, uint32_t flag,
                                     uint32_t bdqual, float upper, float lower, int32_t min_mapq, int32_t max_sd,
                                     int32_t transchr, int32_t long_insert) {
    LibDev L{upper, lower, min_mapq, 0};
    ClassifyOpts o{max_sd, transchr, long_insert};
    return classify_record(pos, mpos, tid, mtid, isize, flag, bdqual, L, o);
}

// Exhaustive equivalence of the streaming pass's classify_hot() with classify_record() over every
// flag word, tid/pos/isize relation, mapq relation and option combination. Returns the number of
// mismatches (0 expected) and the number of cases through *ncases.
extern "C" long hostsim_classify_hot_check(long* ncases) {
    long bad = 0, n = 0;
    const float upper = 400.5f, lower = 200.25f;
    const int32_t isz[] = {0, 1, 100, 200, 201, 300, 400, 401, -1, -200, -201, -400, -401, 999, 1000, 1001, -1001, 2147483647, -2147483647};
    const int32_t posp[][2] = {{100, 500}, {500, 100}, {100, 100}, {0, 0}, {-1, 5}};
    const int32_t tidp[][2] = {{1, 1}, {1, 2}, {0, 0}, {3, -1}};
    for (int long_insert = 0; long_insert < 2; ++long_insert)
        for (int transchr = 0; transchr < 2; ++transchr)
            for (int max_sd : {1000, 1000000000, 0})
                for (uint32_t flag = 0; flag < 0x1000; ++flag)
                    for (auto& tp : tidp)
                        for (auto& pp : posp)
                            for (int32_t is : isz)
                                for (uint32_t mq : {0u, 35u, 36u, 255u}) {
                                    ClassifyOpts o{max_sd, transchr, long_insert};
                                    uint32_t cr = classify_record(pp[0], pp[1], tp[0], tp[1], is, flag, mq, upper, lower, 35, o);
                                    uint32_t ch = classify_hot(pp[0], pp[1], tp[0], tp[1], is, flag, mq, upper, lower, 35, o);
                                    uint32_t want = ((cr & CR_ANOM) ? CH_ANOM : 0u) | ((cr & CR_MPROPER) ? CH_MPROPER : 0u) |
                                                    (((cr >> CR_HIST_SHIFT) & 0xFu) ? CH_HIST : 0u) | ((cr & CR_SPROPER) ? CH_SPROPER : 0u);
                                    ++n;
                                    if (ch != want) ++bad;
                                }
    if (ncases) *ncases = n;
    return bad;
}
