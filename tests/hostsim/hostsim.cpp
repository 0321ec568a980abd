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

// Host rendering of the piece-parallel walk the CUDA path uses for very large components (k4_component_cta): same
// decomposition, pieces of a window walked in DESCENDING order (nothing may depend on their order), finality pass evaluated
// against the earlier state and resolved afterwards.
static int component_by_pieces(const K4Static& S, K4Mut& M, DEdge* e, int ne, int row0, int nrows) {
    SoloTeam T;
    if (S.rerun) {
        for (int q = 0; q < ne; ++q) k4_reset_slot(S, M, e, q);
        for (int r = 0; r < nrows; ++r) M.row_emit[row0 + r] = 0;
    }
    int row_base = row0, i = 0;
    std::vector<int32_t> queue;
    while (i < ne) {
        const int w = e[i].win;
        int j = i;
        while (j < ne && e[j].win == w) ++j;
        WindowInfo wi = k4_window_info(S, w);
        std::vector<int> vtx, rs;
        for (int t = i; t < j; ++t) if (t == i || e[t].src != e[t - 1].src) { vtx.push_back(e[t].src); rs.push_back(t); }
        rs.push_back(j);
        const int R = (int)vtx.size();
        auto run_of = [&](int v) { return (int)(std::lower_bound(vtx.begin(), vtx.end(), v) - vtx.begin()); };
        std::vector<int> label(R);
        std::iota(label.begin(), label.end(), 0);
        auto followable = [&](const DEdge& x) { return x.w >= S.min_read_pair && !M.deleted[x.dst] && !M.deleted[x.src]; };
        for (bool changed = true; changed;) {
            changed = false;
            for (int r = 0; r < R; ++r)
                for (int t = rs[r]; t < rs[r + 1]; ++t) {
                    if (!followable(e[t])) continue;
                    int r2 = run_of(e[t].dst);
                    int m = std::min(label[r], label[r2]);
                    if (label[r] != m || label[r2] != m) { label[r] = label[r2] = m; changed = true; }
                }
        }
        std::vector<int> prow(R, 0), pq(R, 0);
        for (int r = 0; r < R; ++r)
            for (int t = rs[r]; t < rs[r + 1]; ++t)
                if (followable(e[t])) { ++pq[label[r]]; if (e[t].src <= e[t].dst) ++prow[label[r]]; }
        std::vector<int> pieces, rowoff;
        int rows = 0;
        for (int r = 0; r < R; ++r) if (label[r] == r && prow[r] > 0) { pieces.push_back(r); rowoff.push_back(rows); rows += prow[r]; }
        for (int p = (int)pieces.size() - 1; p >= 0; --p) {
            int row = row_base + rowoff[p];
            queue.assign(pq[pieces[p]] + 2, 0);
            for (int r = 0; r < R; ++r)
                if (label[r] == pieces[p]) row = k4_bfs_from(T, S, M, e, i, j, w, wi, rs[r], rs[r + 1], queue.data(), row);
            if (row - (row_base + rowoff[p]) != prow[pieces[p]]) return -1;      // every followable edge of a piece is followed exactly once
        }
        row_base += rows;
        // finality pass
        std::vector<int32_t> cand;
        for (int r = 0; r < R; ++r) if (!S.never_final[vtx[r]] && !M.deleted[vtx[r]] && vtx[r] != wi.last_region) cand.push_back(vtx[r]);
        std::vector<uint8_t> state(cand.size() + 1, K4_FIN_NOT);
        for (size_t c = 0; c < cand.size(); ++c) state[c] = k4_region_final(T, S, M, cand[c], wi) ? K4_FIN_UNDECIDED : K4_FIN_NOT;
        for (bool pending = true; pending;) {
            pending = false;
            for (int c = (int)cand.size() - 1; c >= 0; --c) {                    // descending on purpose
                if (state[c] != K4_FIN_UNDECIDED) continue;
                int d = k4_final_deps(T, S, M, cand[c], cand.data(), state.data(), (int)cand.size());
                if (d & 1) state[c] = K4_FIN_NOT;
                else if (!(d & 2)) state[c] = K4_FIN_CLEARED;
                else pending = true;
            }
        }
        for (size_t c = 0; c < cand.size(); ++c) if (state[c] == K4_FIN_CLEARED) { M.deleted[cand[c]] = 1; M.del_cur[cand[c]] = w; }
        i = j;
    }
    return row_base - row0;
}

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
    std::vector<int32_t> cand_maxlen(ncand);
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
    // ---- components (union-find, smaller root wins) ----------------------------------------------
    std::vector<int> parent(nreg);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    // components over the edges the walk can follow (weight >= -r); a weaker edge is kept as a directed copy in each end's component
    for (auto const& e : ue) { if (e.w < p.min_read_pair) continue; int a = find(e.r0), b = find(e.r1); if (a != b) { if (a < b) parent[b] = a; else parent[a] = b; } }
    std::vector<int32_t> root_of(nreg);
    for (int r = 0; r < nreg; ++r) root_of[r] = find(r);
    std::vector<int> comp_ne(nreg, 0), comp_strong(nreg, 0);
    for (auto const& e : ue) {
        ++comp_ne[root_of[e.r0]];
        if (e.r0 != e.r1) ++comp_ne[root_of[e.r1]];
        if (e.w >= p.min_read_pair) ++comp_strong[root_of[e.r0]];
    }
    std::vector<int> de_off(nreg + 1, 0), row_off(nreg + 1, 0);
    for (int r = 0; r < nreg; ++r) { de_off[r + 1] = de_off[r] + comp_ne[r]; row_off[r + 1] = row_off[r] + comp_strong[r]; }
    std::vector<DEdge> de(de_off[nreg] + 1);
    std::vector<int> fill(nreg, 0);
    std::vector<int> win_first(nreg + 1, 0x7fffffff), win_last(nreg + 1, -1);
    for (auto const& e : ue) {
        int win = e.r1 / period;  // r0 <= r1: the edge is counted when r1 is registered
        for (int v : {e.r0, e.r1}) { win_first[v] = std::min(win_first[v], win); win_last[v] = std::max(win_last[v], win); }
        int r = root_of[e.r0];
        de[de_off[r] + fill[r]++] = DEdge{win, e.r0, e.r1, e.w, 0};
        if (e.r0 != e.r1) { r = root_of[e.r1]; de[de_off[r] + fill[r]++] = DEdge{win, e.r1, e.r0, e.w, 0}; }
    }
    int nrow_cap = row_off[nreg];

    // ---- K4 -----------------------------------------------------------------------------------
    std::vector<uint32_t> Pflat((size_t)nkey * std::max<int64_t>(A, 1));
    for (int k = 0; k < nkey; ++k) for (int64_t d = 0; d < A; ++d) Pflat[(size_t)d * nkey + k] = P[k][d];   // [A][nkey] like the device
    std::vector<uint8_t> freed(A, 0), deleted(nreg, 0), row_emit(nrow_cap + 1, 0);
    std::vector<int32_t> sv_of_read(A, -1), row_lib_count((size_t)(nrow_cap + 1) * nlib), row_lib_span((size_t)(nrow_cap + 1) * nlib);
    std::vector<uint32_t> row_cn_count((size_t)(nrow_cap + 1) * nkey);
    std::vector<float> row_cn((size_t)(nrow_cap + 1) * nkey);
    std::vector<bdk_sv> rows(nrow_cap + 1);
    std::vector<uint64_t> row_key(nrow_cap + 1, 0);
    std::vector<int32_t> del_prev(nreg + 1, K4_NEVER), del_cur(nreg + 1, K4_NEVER);
    std::vector<uint8_t> dirty(nreg + 1, 0);
    K4Static KS;
    std::vector<ReadInfo> ri((size_t)std::max<int64_t>(A, 1));
    for (int64_t j = 0; j < A; ++j) ri[j] = make_read_info(ar.data(), mate.data(), read_region.data(), read_cand.data(), (int)j);
    KS.ri = ri.data();
    KS.ar = ar.data(); KS.read_region = read_region.data(); KS.read_cand = read_cand.data(); KS.mate = mate.data();
    KS.reg = reg.data(); KS.P = Pflat.data(); KS.cand_maxlen = cand_maxlen.data(); KS.lib_mean = lib_mean.data();
    KS.hist = acc.hist.data(); KS.density = density.data(); KS.A = (uint64_t)A; KS.nreg = nreg; KS.ncand = ncand;
    KS.period = period; KS.nkey = nkey; KS.nlib = nlib; KS.chr_restricted = p.chr_restricted;
    KS.min_read_pair = p.min_read_pair; KS.score_threshold = p.score_threshold; KS.fisher = p.fisher;
    KS.covered_ref_len = S.covered_ref_len;
    KS.root_of = root_of.data(); KS.del_prev = del_prev.data(); KS.rerun = 0;
    K4Mut KM;
    KM.alive = alive.data(); KM.freed = freed.data(); KM.deleted = deleted.data(); KM.sv_of_read = sv_of_read.data();
    KM.rows = rows.data(); KM.row_lib_count = row_lib_count.data(); KM.row_lib_span = row_lib_span.data();
    KM.row_cn_count = row_cn_count.data(); KM.row_cn = row_cn.data(); KM.row_emit = row_emit.data(); KM.row_key = row_key.data();
    KM.del_cur = del_cur.data();
    std::vector<uint8_t> never_final(nreg + 1, 0);
    KS.never_final = never_final.data();
    for (int v = 0; v < nreg; ++v) {
        del_prev[v] = getenv("HOSTSIM_NO_GUESS") ? K4_NEVER : k4_guess_deletion(KS, alive.data(), v, win_last[v]);
        never_final[v] = k4_never_final(KS, alive.data(), v) ? 1 : 0;
    }
    const int big = getenv("HOSTSIM_BIG") ? atoi(getenv("HOSTSIM_BIG")) : 4096;   // tests lower it to exercise the deferral of big components
    int nsmall_prev = 1;
    std::vector<int32_t> queue;
    // sweeps over the components until the table of deletion times is stable (same driver as bdk_finish). The components are
    // walked in DESCENDING root order on purpose: nothing may depend on the order inside a sweep.
    int sweeps = 0;
    for (;; ++sweeps) {
        if (sweeps > 100000) return -101;
        KS.rerun = sweeps ? 1 : 0;
        std::vector<int> deferred;
        for (int r = nreg - 1; r >= 0; --r) {
            if (!comp_ne[r] || (sweeps && !dirty[r])) continue;
            if (nsmall_prev && comp_ne[r] > big) { deferred.push_back(r); continue; }
            queue.assign(comp_ne[r] + 2, 0);
            DEdge* es = de.data() + de_off[r];
            de_sort(es, comp_ne[r]);
            const int pieces_min = getenv("HOSTSIM_PIECES") ? atoi(getenv("HOSTSIM_PIECES")) : 4096;   // tests lower it
            int used = comp_ne[r] > pieces_min ? component_by_pieces(KS, KM, es, comp_ne[r], row_off[r], comp_strong[r])
                                               : k4_component(SoloTeam(), KS, KM, es, comp_ne[r], queue.data(), row_off[r], comp_strong[r]);
            if (used < 0) return -102;
            if (used > comp_strong[r]) return -100;
        }
        std::fill(dirty.begin(), dirty.end(), 0);
        int ndirty = 0, nsmall = 0;
        for (int r : deferred) { dirty[r] = 1; ++ndirty; }
        for (int r = 0; r < nreg; ++r)
            for (int t = de_off[r]; t < de_off[r + 1]; ++t) {
                const DEdge& x = de[t];
                if (root_of[x.dst] == r) continue;
                if (never_final[x.src]) continue;
                if (!dirty[r] && k4_change_matters(x.src, x.dst, win_first[x.src], std::min(win_last[x.src], del_cur[x.src]), del_prev[x.dst], del_cur[x.dst])) {
                    dirty[r] = 1; ++ndirty;
                    if (comp_ne[r] <= big) ++nsmall;
                }
            }
        for (int v = 0; v < nreg; ++v) del_prev[v] = del_cur[v];      // (the walk resets the regions of its own component)
        nsmall_prev = nsmall;
        if (getenv("HOSTSIM_VERBOSE")) fprintf(stderr, "hostsim: sweep %d -> %d components to walk again\n", sweeps, ndirty);
        if (!ndirty) break;
    }
    if (getenv("HOSTSIM_VERBOSE")) {
        int mx = 0, nf = 0; for (int r = 0; r < nreg; ++r) { mx = std::max(mx, comp_ne[r]); nf += never_final[r]; }
        fprintf(stderr, "hostsim: %d regions (%d never final), %zu edges, largest component %d directed edges, %d sweeps\n", nreg, nf, ue.size(), mx, sweeps + 1);
    }
    for (int r = 0; r < nrow_cap; ++r) if (row_emit[r] == K4_ROW_PENDING) k4_score_row(KS, KM, r);
    // final order: stable by (window, BFS start vertex), slot order inside
    std::vector<int> order;
    for (int r = 0; r < nrow_cap; ++r) if (row_emit[r] == K4_ROW_EMIT) order.push_back(r);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return row_key[a] < row_key[b]; });

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

// k4_change_matters against its definition: the answers "rm cleared before (w, v)" for deletion windows a and b differ for some
// window w in [wf, wl]. Exhaustive over a small range (K4_NEVER included). Returns the number of mismatches.
extern "C" long hostsim_change_matters_check(long* ncases) {
    long bad = 0, n = 0;
    const int vals[] = {0, 1, 2, 3, 4, 5, 6, K4_NEVER};
    auto before = [](int d, int rm, int w, int v) { return d < w || (d == w && rm < v); };
    for (int v = 0; v < 3; ++v)
        for (int rm = 0; rm < 3; ++rm) {
            if (rm == v) continue;
            for (int wf = 0; wf <= 6; ++wf)
                for (int wl : {-1, 0, 1, 2, 3, 4, 5, 6, K4_NEVER})
                    for (int a : vals)
                        for (int b : vals) {
                            bool want = false;
                            for (int w = wf; w <= 6 && w <= wl; ++w) want = want || (before(a, rm, w, v) != before(b, rm, w, v));
                            // windows beyond 6 (wl = K4_NEVER): both finite deletions are "before" there; K4_NEVER never is
                            if (wl == K4_NEVER && ((a == K4_NEVER) != (b == K4_NEVER))) want = true;
                            ++n;
                            if (k4_change_matters(v, rm, wf, wl, a, b) != want) ++bad;
                        }
        }
    if (ncases) *ncases = n;
    return bad;
}
