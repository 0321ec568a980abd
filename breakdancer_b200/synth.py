"""Synthetic position-sorted read-pair records shaped like the BASELINE.json configs.

Not part of the hot path: this only manufactures inputs (struct-of-arrays columns, the layout
bdk_soa describes) for tests and bench.py.  Generation is numpy on the host for the parity tests
(deterministic from the seed) and torch on the GPU for the full-size bench workloads.

Record conventions follow SAM: each pair yields two records; the leftmost read carries
+isize, the rightmost -isize; flag bits 0x1 paired, 0x2 proper, 0x10 reverse, 0x20 mate reverse,
0x40/0x80 first/second.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

GRCH38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555),
          ("chr5", 181538259), ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636),
          ("chr9", 138394717), ("chr10", 133797422), ("chr11", 135086622), ("chr12", 133275309),
          ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189), ("chr16", 90338345),
          ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
          ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]


@dataclasses.dataclass
class LibSpec:
    name: str
    bam: str
    mean: float
    std: float
    readlen: int = 75
    read_groups: Sequence[str] = ()
    tumor: bool = False


@dataclasses.dataclass
class Workload:
    """Columns plus everything needed to write BAMs / a bam2cfg config for them."""
    cols: Dict[str, np.ndarray]
    genome: List[Tuple[str, int]]
    libs: List[LibSpec]
    rg_names: List[str]       # rgid -> read group string
    rg_libname: List[str]     # rgid -> library name
    rg_bamname: List[str]     # rgid -> bam file name

    @property
    def n(self) -> int:
        return len(self.cols["pos"])

    def config_text(self, with_cutoffs: bool = False, cut_sd: int = 3) -> str:
        lines = []
        for rg, lib in zip(self.rg_names, self.rg_libname):
            L = next(l for l in self.libs if l.name == lib)
            f = [f"readgroup:{rg}", "platform:illumina", f"map:{L.bam}", f"readlen:{L.readlen:.2f}",
                 f"lib:{L.name}", "num:10001"]
            if with_cutoffs:
                f += [f"lower:{max(0.0, L.mean - cut_sd * L.std):.2f}", f"upper:{L.mean + cut_sd * L.std:.2f}"]
            f += [f"mean:{L.mean:.2f}", f"std:{L.std:.2f}", "exe:samtools view"]
            lines.append("\t".join(f))
        return "\n".join(lines) + "\n"


def _emit(parts, tid1, pos1, tid2, pos2, rev1, rev2, proper, isz, lib_rg, mapq1, mapq2, qlen, pair_id, extra_flag=0):
    """Append both records of each pair to parts. isz = |isize| reported for same-chromosome pairs."""
    same = tid1 == tid2
    left1 = pos1 <= pos2
    isz1 = np.where(same, np.where(left1, isz, -isz), 0).astype(np.int32)
    f_common = 0x1 | np.where(proper, 0x2, 0) | extra_flag
    flag1 = (f_common | np.where(rev1, 0x10, 0) | np.where(rev2, 0x20, 0) | 0x40).astype(np.uint16)
    flag2 = (f_common | np.where(rev2, 0x10, 0) | np.where(rev1, 0x20, 0) | 0x80).astype(np.uint16)
    n = len(pos1)
    q = np.full(n, qlen, np.int32) if np.isscalar(qlen) else qlen.astype(np.int32)
    for (t, p, mt, mp, isz_s, fl, mq) in ((tid1, pos1, tid2, pos2, isz1, flag1, mapq1),
                                          (tid2, pos2, tid1, pos1, -isz1, flag2, mapq2)):
        parts.append(dict(pos=p.astype(np.int32), mpos=mp.astype(np.int32), tid=t.astype(np.int32),
                          mtid=mt.astype(np.int32), isize=isz_s.astype(np.int32), flag=fl, mapq=mq.astype(np.uint8),
                          rgid=lib_rg.astype(np.uint16), qlen=q, qid=pair_id.astype(np.uint64)))


def _mapq(rng, n, lo_frac=0.05, mid_frac=0.05):
    u = rng.random(n)
    m = np.full(n, 60, np.int64)
    mid = u < mid_frac
    lo = (u >= mid_frac) & (u < mid_frac + lo_frac)
    m[mid] = rng.integers(36, 60, mid.sum())
    m[lo] = rng.integers(0, 36, lo.sum())
    return m


def generate(genome: Sequence[Tuple[str, int]], libs: Sequence[LibSpec], n_pairs: int, seed: int,
             anomaly_frac: float = 0.02, mix: Optional[Dict[str, float]] = None, cluster_frac: float = 0.5,
             cluster_mean_pairs: float = 12.0, noise_large_frac: float = 0.0, n_del_clusters: int = 0,
             del_cluster_pairs: float = 15.0, somatic_frac: float = 0.0, odd_frac: float = 0.002,
             first_tid: int = 0, sort: bool = True) -> Workload:
    """Generate n_pairs read pairs over `genome` (pairs placed proportionally to chromosome length).

    mix: shares of anomalous pairs by type among {"DEL","INS","INV","ITX","CTX"}; a fraction
    cluster_frac of them comes in planted clusters of Poisson(cluster_mean_pairs) pairs, the rest is
    uniform noise. n_del_clusters adds planted deletions (config 2 style) with
    Poisson(del_cluster_pairs) supporting pairs each; noise_large_frac adds uniform large-insert
    pairs. odd_frac of the normal pairs get a filter-exercising oddity (duplicate flag, unmapped
    mate, unpaired, low mapq on one mate).
    """
    rng = np.random.default_rng(seed)
    mix = mix or {"DEL": 0.5, "INS": 0.15, "INV": 0.15, "ITX": 0.10, "CTX": 0.10}
    glen = np.array([g[1] for g in genome], np.int64)
    gprob = glen / glen.sum()
    nlib = len(libs)
    rg_names, rg_libname, rg_bamname, lib_rgs = [], [], [], []
    for L in libs:
        rgs = list(L.read_groups) or [L.name + ".rg"]
        ids = []
        for r in rgs:
            ids.append(len(rg_names)); rg_names.append(r); rg_libname.append(L.name); rg_bamname.append(L.bam)
        lib_rgs.append(np.array(ids))
    means = np.array([L.mean for L in libs]); stds = np.array([L.std for L in libs])
    rls = np.array([L.readlen for L in libs])
    tumor = np.array([L.tumor for L in libs])
    parts: List[dict] = []
    next_id = [1]

    def ids(n):
        a = np.arange(next_id[0], next_id[0] + n, dtype=np.uint64); next_id[0] += n
        return a

    def pick_lib(n, only_tumor=None):
        lib = rng.integers(0, nlib, n)
        if only_tumor is not None and tumor.any():
            tl = np.flatnonzero(tumor)
            lib = np.where(only_tumor, tl[rng.integers(0, len(tl), n)], lib)
        return lib

    def pick_rg(lib):
        out = np.empty(len(lib), np.int64)
        for l in range(nlib):
            m = lib == l
            out[m] = lib_rgs[l][rng.integers(0, len(lib_rgs[l]), m.sum())]
        return out

    def insert(lib):
        return np.maximum(rls[lib] + 1, np.rint(rng.normal(means[lib], stds[lib]))).astype(np.int64)

    def place(n):
        tid = rng.choice(len(genome), n, p=gprob)
        pos = (rng.random(n) * (glen[tid] - 60000)).astype(np.int64) + 1000
        return tid, pos

    n_anom = int(round(n_pairs * anomaly_frac))
    n_noise_large = int(round(n_pairs * noise_large_frac))
    n_normal = max(0, n_pairs - n_anom - n_noise_large)

    # ---- normal FR pairs ----------------------------------------------------------------------
    lib = pick_lib(n_normal)
    tid, pos = place(n_normal)
    ins = insert(lib)
    rl = rls[lib]
    mq1, mq2 = _mapq(rng, n_normal), _mapq(rng, n_normal)
    extra = np.zeros(n_normal, np.int64)
    odd = rng.random(n_normal) < odd_frac
    kind = rng.integers(0, 4, n_normal)
    extra[odd & (kind == 0)] = 0x400                     # duplicate
    proper = np.ones(n_normal, bool)
    _emit(parts, tid + first_tid, pos, tid + first_tid, pos + ins - rl, np.zeros(n_normal, bool), np.ones(n_normal, bool),
          proper, ins, pick_rg(lib), mq1, mq2, rl, ids(n_normal), extra)
    # mate-unmapped / unpaired oddities are patched in afterwards on the emitted records
    odd_unmapped = odd & (kind == 1)
    odd_unpaired = odd & (kind == 2)
    for rec, other in ((parts[-2], parts[-1]), (parts[-1], parts[-2])):
        rec["flag"] = rec["flag"].copy()
    parts[-2]["flag"][odd_unmapped] |= 0x8
    parts[-2]["flag"][odd_unmapped] &= ~np.uint16(0x2)
    parts[-1]["flag"][odd_unmapped] |= 0x4
    parts[-1]["flag"][odd_unmapped] &= ~np.uint16(0x2)
    parts[-2]["flag"][odd_unpaired] &= ~np.uint16(0x1)
    parts[-1]["flag"][odd_unpaired] &= ~np.uint16(0x1)

    # ---- anomalous pairs ------------------------------------------------------------------------
    def anomalous(kind_name, tid, pos, lib, somatic=None):
        n = len(pos)
        if n == 0:
            return
        ins = insert(lib); rl = rls[lib]
        t2 = tid.copy(); rev1 = np.zeros(n, bool); rev2 = np.ones(n, bool)
        if kind_name == "DEL":
            span = rng.integers(500, 20000, n) if somatic is None else somatic
            p1 = pos - rng.integers(0, np.maximum(1, means[lib].astype(np.int64) - rl), n); p2 = p1 + ins + span - rl
            isz = ins + span
        elif kind_name == "INS":
            isz = np.maximum(rl + 1, (means[lib] - 4 * stds[lib] - rng.integers(20, 120, n)).astype(np.int64))
            p1 = pos + rng.integers(0, 100, n); p2 = p1 + isz - rl
        elif kind_name == "INV":
            ff = rng.random(n) < 0.5
            rev1 = ~ff; rev2 = ~ff
            span = rng.integers(300, 5000, n) if somatic is None else somatic
            p1 = pos + rng.integers(0, 150, n); p2 = p1 + span + rng.integers(0, 150, n); isz = p2 - p1 + rl
        elif kind_name == "ITX":
            rev1 = np.ones(n, bool); rev2 = np.zeros(n, bool)
            span = rng.integers(300, 5000, n) if somatic is None else somatic
            p1 = pos + rng.integers(0, 150, n); p2 = p1 + span + rng.integers(0, 150, n); isz = p2 - p1 + rl
        elif kind_name == "CTX":
            if len(genome) > 1:
                t2 = (tid + 1 + rng.integers(0, len(genome) - 1, n)) % len(genome)
            if somatic is not None:
                p2 = somatic + rng.integers(0, 150, n)
            else:
                p2 = (rng.random(n) * (glen[t2] - 60000)).astype(np.int64) + 1000
            p1 = pos + rng.integers(0, 150, n); isz = np.zeros(n, np.int64)
        else:
            raise ValueError(kind_name)
        _emit(parts, tid + first_tid, p1, t2 + first_tid, p2, rev1, rev2, np.zeros(n, bool), isz, pick_rg(lib),
              _mapq(rng, n), _mapq(rng, n), rl, ids(n))

    kinds = list(mix)
    shares = np.array([mix[k] for k in kinds], float); shares /= shares.sum()
    n_cluster_pairs = int(n_anom * cluster_frac)
    n_noise = n_anom - n_cluster_pairs
    # planted clusters
    if n_cluster_pairs > 0:
        ncl = max(1, int(n_cluster_pairs / cluster_mean_pairs))
        sizes = rng.poisson(cluster_mean_pairs, ncl)
        ck = rng.choice(len(kinds), ncl, p=shares)
        ctid, cpos = place(ncl)
        csom = rng.random(ncl) < somatic_frac
        cspan = rng.integers(500, 20000, ncl)
        cpos2 = (rng.random(ncl) * 1e7).astype(np.int64) + 1000
        rep = np.repeat(np.arange(ncl), sizes)
        for ki, kn in enumerate(kinds):
            m = ck[rep] == ki
            r = rep[m]
            lib = pick_lib(len(r), only_tumor=csom[r] if somatic_frac > 0 else None)
            som = cspan[r] if kn != "CTX" else cpos2[r]
            anomalous(kn, ctid[r], cpos[r], lib, somatic=som)
    if n_noise > 0:
        nk = rng.choice(len(kinds), n_noise, p=shares)
        for ki, kn in enumerate(kinds):
            m = nk == ki
            t, p = place(int(m.sum()))
            anomalous(kn, t, p, pick_lib(len(p)))
    # config-2 style planted deletions and large-insert noise
    if n_del_clusters > 0:
        sizes = rng.poisson(del_cluster_pairs, n_del_clusters)
        ctid, cpos = place(n_del_clusters)
        clen = rng.integers(500, 20000, n_del_clusters)
        rep = np.repeat(np.arange(n_del_clusters), sizes)
        anomalous("DEL", ctid[rep], cpos[rep], pick_lib(len(rep)), somatic=clen[rep])
    if n_noise_large > 0:
        t, p = place(n_noise_large)
        lib = pick_lib(n_noise_large)
        ins = rng.integers(600, 50000, n_noise_large); rl = rls[lib]
        _emit(parts, t + first_tid, p, t + first_tid, p + ins - rl, np.zeros(n_noise_large, bool), np.ones(n_noise_large, bool),
              np.zeros(n_noise_large, bool), ins, pick_rg(lib), _mapq(rng, n_noise_large), _mapq(rng, n_noise_large), rl,
              ids(n_noise_large))

    cols = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    np.clip(cols["pos"], 0, None, out=cols["pos"])
    np.clip(cols["mpos"], 0, None, out=cols["mpos"])
    if sort:
        cols = sort_columns(cols)
    return Workload(cols, list(genome), list(libs), rg_names, rg_libname, rg_bamname)


def sort_columns(cols: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Stable sort by (tid, pos, strand): the order samtools sort / BamMerger deliver."""
    strand = (cols["flag"] & 0x10) != 0
    order = np.lexsort((strand, cols["pos"], cols["tid"]))
    return {k: np.ascontiguousarray(v[order]) for k, v in cols.items()}


def split_by_bam(w: Workload) -> Dict[str, Dict[str, np.ndarray]]:
    """Per-bam record subsets (in stream order) for writing one BAM file per bam name."""
    out = {}
    bam_of_rg = np.array([sorted(set(w.rg_bamname)).index(b) for b in w.rg_bamname])
    names = sorted(set(w.rg_bamname))
    b = bam_of_rg[w.cols["rgid"]]
    for i, name in enumerate(names):
        m = b == i
        out[name] = {k: np.ascontiguousarray(v[m]) for k, v in w.cols.items()}
    return out


def config2(n_pairs: int = 50_000_000, seed: int = 20260101, chrom_len: int = 248956422) -> Workload:
    """BASELINE.json configs[1]: single-library 30x chr1, DEL-only."""
    scale = n_pairs / 50_000_000
    lib = LibSpec("lib1", "syn_chr1.bam", 315.09, 43.92, 75, ["rg1"])
    L = max(200000, int(chrom_len * min(1.0, scale))) if scale < 1 else chrom_len
    return generate([("chr1", L)], [lib], n_pairs, seed, anomaly_frac=0.0, noise_large_frac=0.005,
                    n_del_clusters=max(1, int(5000 * scale)), del_cluster_pairs=15.0)


def config3(n_pairs: int = 300_000_000, seed: int = 20260102) -> Workload:
    """BASELINE.json configs[2]: 4-library tumor/normal chr1-3, all five SV types."""
    scale = n_pairs / 300_000_000
    genome = [(n, max(300000, int(l * min(1.0, scale)))) for n, l in GRCH38[:3]]
    libs = [LibSpec("normal_a", "normal.bam", 315, 44, 75, ["n_a1", "n_a2"]),
            LibSpec("normal_b", "normal.bam", 312, 43, 75, ["n_b1", "n_b2"]),
            LibSpec("tumor_a", "tumor.bam", 467, 32, 75, ["t_a1", "t_a2"], tumor=True),
            LibSpec("tumor_b", "tumor.bam", 476, 29, 75, ["t_b1", "t_b2"], tumor=True)]
    return generate(genome, libs, n_pairs, seed, anomaly_frac=0.02, somatic_frac=0.3)
