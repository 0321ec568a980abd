"""Full-size synthetic workloads generated directly in GPU memory with torch (bench.py only).

Same distributions as ``synth.config2`` (BASELINE.json configs[1]: single library, 30x, chr1-sized,
DEL-only: 0.5 % uniform large-insert pairs + planted deletions of Poisson(15) supporting pairs), but
100 M records are produced in about a second on the device instead of minutes on the host.
torch is plumbing here (random numbers, sort, gather); nothing of this is on the measured path.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import api

MEAN, STD, READLEN = 315.09, 43.92, 75
CHR1_LEN = 248956422


def _mapq(n, gen, device):
    u = torch.rand(n, generator=gen, device=device)
    m = torch.full((n,), 60, dtype=torch.int32, device=device)
    mid = u < 0.05
    lo = (u >= 0.05) & (u < 0.10)
    m = torch.where(mid, (36 + torch.rand(n, generator=gen, device=device) * 24).to(torch.int32), m)
    m = torch.where(lo, (torch.rand(n, generator=gen, device=device) * 36).to(torch.int32), m)
    return m


# GRCh38 primary assembly, chr1..22, X, Y (3 088 269 832 bp: BASELINE configs[3] / configs[4])
GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622, 133275309,
          114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]


def genome_plan(total_pairs: int, ctx_frac: float, lens=GRCH38):
    """Read pairs per chromosome (proportional to length) and the symmetric matrix of inter-chromosomal pair counts
    (ctx_frac of the pairs, proportional to the product of the lengths) of a whole-genome workload."""
    tot = float(sum(lens))
    pairs = [int(round(total_pairs * l / tot)) for l in lens]
    denom = tot * tot - sum(float(l) * l for l in lens)
    n = len(lens)
    ctx = [[0] * n for _ in range(n)]
    for i in range(n):
        for j in range(i + 1, n):
            ctx[i][j] = ctx[j][i] = int(round(2.0 * ctx_frac * total_pairs * lens[i] * lens[j] / denom))
    return pairs, ctx


def config2_device(n_pairs: int, seed: int, device, chrom_len: int = CHR1_LEN, tid: int = 0, fixed_length: bool = False) -> Dict[str, torch.Tensor]:
    """Position-sorted record columns (2 records per pair) on `device`.
    Columns use torch dtypes with the bit patterns bdk_soa expects: flag/rgid int16, qid int64.
    fixed_length: the chromosome keeps its length whatever the number of pairs (else it shrinks with the pairs below 50 M
    so that the coverage stays 30x)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    scale = n_pairs / 50_000_000
    L = chrom_len if (scale >= 1 or fixed_length) else max(200000, int(chrom_len * scale))
    n_noise = int(round(n_pairs * 0.005))
    n_normal = n_pairs - n_noise
    ncl = max(1, int(5000 * scale))
    sizes = torch.poisson(torch.full((ncl,), 15.0, device=dev), generator=gen).to(torch.int64)
    n_del = int(sizes.sum().item())

    def uni(n, lo, hi):
        return (lo + torch.rand(n, generator=gen, device=dev, dtype=torch.float64) * (hi - lo)).to(torch.int64)

    def insert(n):
        x = torch.round(MEAN + STD * torch.randn(n, generator=gen, device=dev)).to(torch.int64)
        return torch.clamp(x, min=READLEN + 1)

    # normal FR proper pairs
    p1n = uni(n_normal, 1000, L - 59000)
    insn = insert(n_normal)
    # uniform large-insert noise
    p1z = uni(n_noise, 1000, L - 59000)
    insz = uni(n_noise, 600, 50000)
    # planted deletions
    cpos = uni(ncl, 1000, L - 59000)
    clen = uni(ncl, 500, 20000)
    rep = torch.repeat_interleave(torch.arange(ncl, device=dev), sizes)
    p1d = cpos[rep] - uni(n_del, 0, int(MEAN) - READLEN)
    insd = insert(n_del) + clen[rep]

    p1 = torch.cat([p1n, p1z, p1d]).clamp_(min=0)
    isz = torch.cat([insn, insz, insd])
    proper = torch.cat([torch.ones(n_normal, dtype=torch.bool, device=dev),
                        torch.zeros(n_noise + n_del, dtype=torch.bool, device=dev)])
    npair = p1.numel()
    p2 = p1 + isz - READLEN
    base = 0x1 | torch.where(proper, 0x2, 0)
    flag1 = (base | 0x20 | 0x40).to(torch.int16)
    flag2 = (base | 0x10 | 0x80).to(torch.int16)
    mq1, mq2 = _mapq(npair, gen, dev), _mapq(npair, gen, dev)
    qid = torch.arange(1, npair + 1, device=dev, dtype=torch.int64)

    pos = torch.cat([p1, p2])
    key = pos * 2 + torch.cat([torch.zeros(npair, dtype=torch.int64, device=dev), torch.ones(npair, dtype=torch.int64, device=dev)])
    order = torch.sort(key, stable=True).indices
    del key

    def both(a, b, dtype):
        return torch.cat([a, b]).to(dtype)[order].contiguous()

    cols = {
        "pos": both(p1, p2, torch.int32), "mpos": both(p2, p1, torch.int32),
        "isize": both(isz, -isz, torch.int32),
        "flag": both(flag1, flag2, torch.int16), "mapq": both(mq1, mq2, torch.uint8),
        "qid": both(qid, qid, torch.int64),
    }
    n = 2 * npair
    cols["tid"] = torch.full((n,), tid, dtype=torch.int32, device=dev)
    cols["mtid"] = torch.full((n,), tid, dtype=torch.int32, device=dev)
    cols["rgid"] = torch.zeros(n, dtype=torch.int16, device=dev)
    cols["qlen"] = torch.full((n,), READLEN, dtype=torch.int32, device=dev)
    return cols


def genome_shard_device(n_pairs: int, seed: int, device, tid: int, ntid: int, ctx_frac: float, clustered: float = 0.3,
                        chrom_len: int = CHR1_LEN, ctx_counts=None, lens=None) -> Dict[str, torch.Tensor]:
    """Chromosome `tid` of an `ntid`-chromosome genome (BASELINE.json configs[3]/[4] shape): the config-2 records of
    the chromosome plus its side of the inter-chromosomal pairs (ctx_frac of the pairs; `clustered` of them in planted
    translocation clusters of Poisson(10) pairs, the rest uniform). The pairs between chromosomes i < j come from a
    generator seeded by (seed, i, j), so the ranks that own i and j produce matching mates without talking.
    With `lens` (all chromosome lengths) and `ctx_counts` (pairs with every other chromosome, genome_plan) the
    chromosomes keep their real lengths and may differ in size."""
    dev = torch.device(device)
    cols = config2_device(n_pairs, seed * 1000 + tid, dev, chrom_len if lens is None else lens[tid], tid, fixed_length=lens is not None)
    cols["qid"] += (tid + 1) << 40                                 # read names unique across chromosomes
    L = chrom_len if n_pairs >= 50_000_000 else max(200000, int(chrom_len * n_pairs / 50_000_000))
    n_ctx = int(round(n_pairs * ctx_frac)) if ctx_counts is None else sum(ctx_counts)
    if ntid < 2 or n_ctx == 0:
        return cols
    per = max(1, n_ctx // (ntid - 1))
    extra = {k: [] for k in ("pos", "mpos", "mtid", "flag", "qid")}
    for other in range(ntid):
        if other == tid:
            continue
        i, j = min(tid, other), max(tid, other)
        gen = torch.Generator(device=dev)
        gen.manual_seed((seed * 1_000_003 + i) * 1009 + j)
        Li, Lj = (L, L) if lens is None else (lens[i], lens[j])
        if ctx_counts is not None:
            per = ctx_counts[other]
            if per <= 0:
                continue
        ncl = max(1, int(per * clustered / 10))
        sizes = torch.poisson(torch.full((ncl,), 10.0, device=dev), generator=gen).to(torch.int64)
        n_cl = int(sizes.sum().item())
        n_un = max(0, per - n_cl)

        def uni(n, lo, hi):
            return (lo + torch.rand(n, generator=gen, device=dev, dtype=torch.float64) * (hi - lo)).to(torch.int64)

        ci, cj = uni(ncl, 1000, Li - 1000), uni(ncl, 1000, Lj - 1000)
        rep = torch.repeat_interleave(torch.arange(ncl, device=dev), sizes)
        pi = torch.cat([ci[rep] - uni(n_cl, 0, 240), uni(n_un, 1000, Li - 1000)])
        pj = torch.cat([cj[rep] + uni(n_cl, 0, 240), uni(n_un, 1000, Lj - 1000)])
        ri = torch.cat([torch.zeros(n_cl, dtype=torch.int64, device=dev), uni(n_un, 0, 2)])      # clusters: + on i, - on j
        rj = torch.cat([torch.ones(n_cl, dtype=torch.int64, device=dev), uni(n_un, 0, 2)])
        k = torch.arange(n_cl + n_un, device=dev, dtype=torch.int64)
        qid = (1 << 62) | ((i * ntid + j) << 36) | k
        if tid == i:
            extra["pos"].append(pi); extra["mpos"].append(pj)
            extra["flag"].append(0x1 | 0x40 | (ri << 4) | (rj << 5))
        else:
            extra["pos"].append(pj); extra["mpos"].append(pi)
            extra["flag"].append(0x1 | 0x80 | (rj << 4) | (ri << 5))
        extra["mtid"].append(torch.full_like(k, other))
        extra["qid"].append(qid)
    ne = sum(int(x.numel()) for x in extra["pos"])
    if ne == 0:
        return cols
    add = {
        "pos": torch.cat(extra["pos"]).to(torch.int32), "mpos": torch.cat(extra["mpos"]).to(torch.int32),
        "mtid": torch.cat(extra["mtid"]).to(torch.int32), "flag": torch.cat(extra["flag"]).to(torch.int16),
        "qid": torch.cat(extra["qid"]),
        "tid": torch.full((ne,), tid, dtype=torch.int32, device=dev), "isize": torch.zeros(ne, dtype=torch.int32, device=dev),
        "mapq": torch.full((ne,), 60, dtype=torch.uint8, device=dev), "rgid": torch.zeros(ne, dtype=torch.int16, device=dev),
        "qlen": torch.full((ne,), READLEN, dtype=torch.int32, device=dev),
    }
    order = torch.sort(torch.cat([cols["pos"], add["pos"]]).to(torch.int64), stable=True).indices
    out = {}
    for name in list(cols):
        out[name] = torch.cat([cols[name], add[name]])[order].contiguous()
        del cols[name]
    return out


CONFIG3_GENOME = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559)]
CONFIG3_LIBS = [("normal_a", "normal.bam", 315.0, 44.0), ("normal_b", "normal.bam", 312.0, 43.0),
                ("tumor_a", "tumor.bam", 467.0, 32.0), ("tumor_b", "tumor.bam", 476.0, 29.0)]     # 2 read groups each


def config3_device(n_pairs: int, seed: int, device) -> Dict[str, torch.Tensor]:
    """BASELINE.json configs[2] on the device: chr1-3, 2 BAMs x 2 libraries x 2 read groups, 2 % anomalous pairs
    (DEL 50 / INS 15 / INV 15 / ITX 10 / CTX 10 %), half of them in planted clusters of Poisson(12) pairs, half uniform
    noise. Same construction as synth.generate (host), vectorised with torch; sorted by (tid, pos, strand).
    Read group id = 2 * library + {0, 1}; library index = rank of the library name (as the config parser assigns it)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    scale = min(1.0, n_pairs / 300_000_000)
    glen = torch.tensor([max(300000, int(l * scale)) for _, l in CONFIG3_GENOME], device=dev, dtype=torch.int64)
    gcum = torch.cumsum(glen.to(torch.float64) / float(glen.sum()), 0)
    means = torch.tensor([l[2] for l in CONFIG3_LIBS], device=dev, dtype=torch.float64)
    stds = torch.tensor([l[3] for l in CONFIG3_LIBS], device=dev, dtype=torch.float64)
    RL = READLEN

    def rnd(n):
        return torch.rand(n, generator=gen, device=dev, dtype=torch.float64)

    def randint(n, lo, hi):
        return (lo + rnd(n) * (hi - lo)).to(torch.int64)

    def place(n):
        tid = torch.bucketize(rnd(n), gcum).clamp_(max=2)
        pos = (rnd(n) * (glen[tid] - 60000).to(torch.float64)).to(torch.int64) + 1000
        return tid, pos

    def insert(lib):
        x = torch.round(means[lib] + stds[lib] * torch.randn(lib.numel(), generator=gen, device=dev, dtype=torch.float64)).to(torch.int64)
        return torch.clamp(x, min=RL + 1)

    parts = []
    next_id = [1]

    def emit(tid1, p1, tid2, p2, rev1, rev2, proper, isz, lib):
        n = p1.numel()
        if n == 0:
            return
        same = tid1 == tid2
        isz1 = torch.where(same, torch.where(p1 <= p2, isz, -isz), torch.zeros_like(isz))
        fc = 0x1 | torch.where(proper, 0x2, 0)
        r1, r2 = rev1.to(torch.int64), rev2.to(torch.int64)
        flag1 = fc | (r1 << 4) | (r2 << 5) | 0x40
        flag2 = fc | (r2 << 4) | (r1 << 5) | 0x80
        rg = (2 * lib + randint(n, 0, 2)).to(torch.int16)
        qid = torch.arange(next_id[0], next_id[0] + n, device=dev, dtype=torch.int64)
        next_id[0] += n
        for t, p, mt, mp, i_s, fl in ((tid1, p1, tid2, p2, isz1, flag1), (tid2, p2, tid1, p1, -isz1, flag2)):
            parts.append(dict(pos=p.clamp(min=0).to(torch.int32), mpos=mp.clamp(min=0).to(torch.int32), tid=t.to(torch.int32), mtid=mt.to(torch.int32),
                              isize=i_s.to(torch.int32), flag=fl.to(torch.int16), mapq=_mapq(n, gen, dev).to(torch.uint8), rgid=rg, qid=qid))

    n_anom = int(round(n_pairs * 0.02))
    n_normal = n_pairs - n_anom
    # normal FR pairs, in slabs to bound the temporaries
    done = 0
    while done < n_normal:
        m = min(100_000_000, n_normal - done)
        lib = randint(m, 0, 4)
        tid, pos = place(m)
        ins = insert(lib)
        f, t = torch.zeros(m, dtype=torch.bool, device=dev), torch.ones(m, dtype=torch.bool, device=dev)
        emit(tid, pos, tid, pos + ins - RL, f, t, t, ins, lib)
        done += m
        del lib, tid, pos, ins

    kinds = ["DEL", "INS", "INV", "ITX", "CTX"]
    shares = torch.tensor([0.5, 0.15, 0.15, 0.10, 0.10], device=dev, dtype=torch.float64)
    kcum = torch.cumsum(shares, 0)

    def anomalous(kind, tid, pos, lib, span):
        n = pos.numel()
        if n == 0:
            return
        ins = insert(lib)
        t2 = tid.clone()
        rev1 = torch.zeros(n, dtype=torch.bool, device=dev)
        rev2 = torch.ones(n, dtype=torch.bool, device=dev)
        if kind == "DEL":
            p1 = pos - randint(n, 0, 240); isz = ins + span; p2 = p1 + isz - RL
        elif kind == "INS":
            isz = torch.clamp((means[lib] - 4 * stds[lib]).to(torch.int64) - randint(n, 20, 120), min=RL + 1)
            p1 = pos + randint(n, 0, 100); p2 = p1 + isz - RL
        elif kind == "INV":
            ff = rnd(n) < 0.5
            rev1 = ~ff; rev2 = ~ff
            p1 = pos + randint(n, 0, 150); p2 = p1 + (span % 4700 + 300) + randint(n, 0, 150); isz = p2 - p1 + RL
        elif kind == "ITX":
            rev1 = torch.ones(n, dtype=torch.bool, device=dev); rev2 = torch.zeros(n, dtype=torch.bool, device=dev)
            p1 = pos + randint(n, 0, 150); p2 = p1 + (span % 4700 + 300) + randint(n, 0, 150); isz = p2 - p1 + RL
        else:   # CTX: mate on another chromosome; clusters share the partner locus (span carries it), noise is uniform
            t2 = (tid + 1 + randint(n, 0, 2)) % 3
            p2 = (span.to(torch.float64) / 20000.0 * (glen[t2] - 60000).to(torch.float64)).to(torch.int64) + 1000 + randint(n, 0, 150)
            p1 = pos + randint(n, 0, 150); isz = torch.zeros(n, dtype=torch.int64, device=dev)
        emit(tid, p1, t2, p2, rev1, rev2, torch.zeros(n, dtype=torch.bool, device=dev), isz, lib)

    n_cl_pairs = n_anom // 2
    ncl = max(1, n_cl_pairs // 12)
    sizes = torch.poisson(torch.full((ncl,), 12.0, device=dev), generator=gen).to(torch.int64)
    ck = torch.bucketize(rnd(ncl), kcum).clamp_(max=4)
    ctid, cpos = place(ncl)
    cspan = randint(ncl, 500, 20000)
    clib_t = randint(ncl, 0, 4)
    rep = torch.repeat_interleave(torch.arange(ncl, device=dev), sizes)
    n_noise = n_anom - int(rep.numel())
    nk = torch.bucketize(rnd(max(n_noise, 0)), kcum).clamp_(max=4)
    for ki, kn in enumerate(kinds):
        r = rep[ck[rep] == ki]
        som = rnd(r.numel()) < 0.3                                   # somatic clusters: tumor libraries only
        lib = torch.where(som, 2 + randint(r.numel(), 0, 2), randint(r.numel(), 0, 4))
        if kn == "CTX":                                              # partner chromosome fixed per cluster
            n = r.numel()
            if n:
                ins = insert(lib)
                t2 = (ctid[r] + 1 + (cspan[r] % 2)) % 3
                p2 = (cspan[r].to(torch.float64) / 20000.0 * (glen[t2] - 60000).to(torch.float64)).to(torch.int64) + 1000 + randint(n, 0, 150)
                emit(ctid[r], cpos[r] + randint(n, 0, 150), t2, p2, torch.zeros(n, dtype=torch.bool, device=dev), torch.ones(n, dtype=torch.bool, device=dev),
                     torch.zeros(n, dtype=torch.bool, device=dev), torch.zeros(n, dtype=torch.int64, device=dev), lib)
        else:
            anomalous(kn, ctid[r], cpos[r], lib, cspan[r])
        m = int((nk == ki).sum().item())
        t, p = place(m)
        anomalous(kn, t, p, randint(m, 0, 4), randint(m, 500, 20000))
    del clib_t

    cols = {}
    key = torch.cat([(x["tid"].to(torch.int64) << 33) | (x["pos"].to(torch.int64) << 1) | ((x["flag"].to(torch.int64) >> 4) & 1) for x in parts])
    order = torch.sort(key, stable=True).indices
    del key
    for name in ("pos", "mpos", "tid", "mtid", "isize", "flag", "mapq", "rgid", "qid"):
        cols[name] = torch.cat([x[name] for x in parts])[order].contiguous()
        for x in parts:
            del x[name]
    cols["qlen"] = torch.full((cols["pos"].numel(),), RL, dtype=torch.int32, device=dev)
    return cols


def config3_bundle():
    """(ParamBundle with `-c 3 -q 35`, config) matching config3_device's read-group numbering."""
    from . import synth
    libs = [synth.LibSpec(n, b, m, s, READLEN, [f"{n}.1", f"{n}.2"], tumor=n.startswith("tumor")) for n, b, m, s in CONFIG3_LIBS]
    rg_names = [r for l in libs for r in l.read_groups]
    rg_lib = [l.name for l in libs for _ in l.read_groups]
    rg_bam = [l.bam for l in libs for _ in l.read_groups]
    wl = synth.Workload({}, list(CONFIG3_GENOME), libs, rg_names, rg_lib, rg_bam)
    cfg = api.BamConfig(text=wl.config_text())
    bams = sorted(set(rg_bam))
    rgl = np.array([cfg.rg_lib(r) for r in rg_names], np.int32)
    rgb = np.array([bams.index(b) for b in rg_bam], np.int32)
    opts = api.Options(min_map_qual=35, cut_sd=3)
    return api.ParamBundle(opts, cfg.libs, cfg.nbam, rgl, rgb, cfg.window, len(CONFIG3_GENOME)), cfg


def soa_of(cols: Dict[str, torch.Tensor]) -> api.Soa:
    return api.soa_from_pointers({k: cols[k].data_ptr() for k in api.COLUMN_DTYPES})


def to_pinned(cols: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, t in cols.items():
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        out[k] = h
    return out


def to_numpy(cols: Dict[str, torch.Tensor]) -> Dict[str, np.ndarray]:
    out = {}
    for k, dt in api.COLUMN_DTYPES.items():
        a = cols[k].cpu().numpy()
        out[k] = np.ascontiguousarray(a.view(dt) if a.dtype != dt else a)
    return out
