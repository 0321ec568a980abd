"""Per-chromosome sharding across the GPUs of one node (BASELINE.json configs[3]).

The reference scales by running one `breakdancer-max -o <chr>` process per chromosome (README:31);
each such run is self-contained (own summary statistics, own window, own region indices). The same
decomposition is used here: chromosomes are packed onto ranks (longest-processing-time first on the
record counts), every rank runs its chromosomes one after the other on its own GPU through its own
bdk context with -o semantics, and the per-chromosome SV tables are gathered to rank 0 in tid order.
There is NO collective on the data path; torch.distributed only carries the final (small) tables.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import numpy as np


def lpt_pack(weights: Sequence[int], nranks: int) -> List[List[int]]:
    """Longest-processing-time-first packing of items (chromosomes) onto nranks bins.
    Returns, per rank, the ascending list of item indices. Deterministic (ties -> lower index / rank)."""
    order = sorted(range(len(weights)), key=lambda i: (-int(weights[i]), i))
    load = [0] * nranks
    bins: List[List[int]] = [[] for _ in range(nranks)]
    for i in order:
        if weights[i] == 0:
            continue
        r = min(range(nranks), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += int(weights[i])
    return [sorted(b) for b in bins]


def chromosome_counts(cols: Dict[str, np.ndarray], ntid: int) -> np.ndarray:
    return np.bincount(cols["tid"], minlength=ntid)[:ntid]


def chromosome_slices(cols: Dict[str, np.ndarray], ntid: int) -> List[slice]:
    """Record ranges of each chromosome in a (tid, pos)-sorted stream."""
    edges = np.searchsorted(cols["tid"], np.arange(ntid + 1))
    return [slice(int(edges[t]), int(edges[t + 1])) for t in range(ntid)]


def run_sharded(cols: Dict[str, np.ndarray], ntid: int, rank: int, world: int,
                run_chromosome: Callable[[int, Dict[str, np.ndarray]], object], gather: bool = True):
    """Run `run_chromosome(tid, columns_of_tid)` for this rank's chromosomes; with gather=True return on
    rank 0 the list [(tid, result)] of ALL ranks ordered by tid (None elsewhere)."""
    counts = chromosome_counts(cols, ntid)
    mine = lpt_pack(counts.tolist(), world)[rank]
    sl = chromosome_slices(cols, ntid)
    local = []
    for t in mine:
        sub = {k: np.ascontiguousarray(v[sl[t]]) for k, v in cols.items()}
        local.append((t, run_chromosome(t, sub)))
    if not gather or world == 1:
        return sorted(local, key=lambda x: x[0]) if rank == 0 or not gather else None
    import torch.distributed as dist
    out = [None] * world if rank == 0 else None
    dist.gather_object(local, out, dst=0)
    if rank != 0:
        return None
    merged = [x for part in out for x in part]
    return sorted(merged, key=lambda x: x[0])


# ---- per-chromosome shards straight from indexed bam files ---------------------------------------------------------------
def plan_from_index(bam_paths: Sequence[str], nranks: int):
    """Chromosomes -> ranks from the bams' .bai files alone (no decoding): weights are the record counts the index holds
    per reference sequence (samtools' pseudo-bin), or the compressed bytes the sequence's chunks span where an index has no
    counts. Returns (weights, per-rank ascending tid lists), or None if a bam has no index."""
    total = None
    for path in bam_paths:
        from . import api
        st = api.bai_reference_stats(path)
        if st is None:
            return None
        rec, byt = st
        # ~64 compressed bytes per record when only spans are known; a sequence that has chunks at all weighs at least 1, so that
        # lpt_pack (which skips weight 0 = no records) never drops a small contig whose chunks lie inside one BGZF member
        w = np.where(rec >= 0, rec, np.where(byt > 0, np.maximum(1, byt // 64), 0))
        total = w.copy() if total is None else total[:min(len(total), len(w))] + w[:min(len(total), len(w))]
    return total, lpt_pack(total.tolist(), nranks)


def run_sharded_bams(cfg, rank: int, world: int, run_chromosome: Callable[[int, str, object], object], gather: bool = True,
                     threads: int = 0, pinned: bool = False):
    """Like run_sharded, from the config's bam files: this rank opens only its chromosomes, each through the index (only that
    sequence's BGZF members are inflated), and calls run_chromosome(tid, name, BamStream). Needs a .bai next to every bam."""
    from . import api
    plan = plan_from_index(cfg.bam_files, world)
    if plan is None:
        raise RuntimeError("per-chromosome shards need a .bai next to every bam file")
    _, bins = plan
    local = []
    names = api.bam_reference_names(cfg.bam_files[0])      # BamMerger uses the first stream's header (BamMerger.cpp:78)
    for t in bins[rank]:
        st = api.BamStream(cfg, region=names[t], threads=threads, pinned=pinned)
        local.append((t, run_chromosome(t, names[t], st)))
        st.close()
    if not gather or world == 1:
        return sorted(local, key=lambda x: x[0]) if rank == 0 or not gather else None
    import torch.distributed as dist
    out = [None] * world if rank == 0 else None
    dist.gather_object(local, out, dst=0)
    if rank != 0:
        return None
    return sorted([x for part in out for x in part], key=lambda x: x[0])


def run_sharded_bams_device(cfg, rank: int, world: int, opts, device: int = 0, gather: bool = True):
    """The per-chromosome shards of this rank from the config's bam file(s), DECODED ON THE GPU: for every chromosome of the rank
    only that sequence's BGZF members (through the .bai) cross PCIe, are inflated, parsed and classified on `device`
    (api.BamDevice + Context.push_bam / push_bams; one or two bams), with `-o <chromosome>` semantics like the reference's own
    one-process-per-chromosome runs (README:31). Returns the (tid, name, summary, SvTable) of all ranks on rank 0 (or of this
    rank with gather=False). No collective on the data path."""
    import dataclasses
    from . import api
    if not 1 <= len(cfg.bam_files) <= 2:
        raise RuntimeError("the device decode takes one or two bams; use run_sharded_bams for more")
    plan = plan_from_index(cfg.bam_files, world)
    if plan is None:
        raise RuntimeError("per-chromosome shards need a .bai next to every bam file")
    _, bins = plan
    names = api.bam_reference_names(cfg.bam_files[0])
    local = []
    for t in bins[rank]:
        o = dataclasses.replace(opts, chr=names[t])
        d0 = api.BamDevice(cfg, path=cfg.bam_files[0], region=names[t])
        d1 = api.BamDevice(cfg, path=cfg.bam_files[1], region=names[t], after=d0) if len(cfg.bam_files) == 2 else None
        rg_lib = d0.rg_lib if d1 is None else np.concatenate([d0.rg_lib, d1.rg_lib])
        rg_bam = d0.rg_bam if d1 is None else np.concatenate([d0.rg_bam, d1.rg_bam])
        ctx = api.Context(api.ParamBundle(o, cfg.libs, cfg.nbam, rg_lib, rg_bam, cfg.window, max(1, len(d0.tid_names))), device)
        if d1 is None:
            ctx.push_bam(d0)
        else:
            ctx.push_bams(d0, d1)
        local.append((t, names[t], ctx.summary(), ctx.finish()))
        ctx.close()
        d0.close()
        if d1 is not None:
            d1.close()
    if not gather or world == 1:
        return sorted(local, key=lambda x: x[0])
    import torch.distributed as dist
    out = [None] * world if rank == 0 else None
    dist.gather_object([(t, n, bytes(s), tb) for t, n, s, tb in local], out, dst=0)
    if rank != 0:
        return None
    return sorted([x for part in out for x in part], key=lambda x: x[0])


# ---- one job over several GPUs (whole-genome / -t semantics, include/bdk.h "Multi-GPU") -----------------------------
def stream_slices(n_records: int, world: int) -> List[slice]:
    """Contiguous, near-equal slices of the globally (tid, pos)-sorted record stream, one per rank in rank order.
    Cuts may fall anywhere (also inside a chromosome): pass 1 is per record plus prefix counts, everything that needs
    neighbours runs on the gathered anomalous reads."""
    edges = [(n_records * r) // world for r in range(world + 1)]
    return [slice(edges[r], edges[r + 1]) for r in range(world)]


def run_one_job(cols: Dict[str, np.ndarray], rank: int, world: int, make_context: Callable[[], object], unique_id: bytes):
    """Rank `rank` of a `world`-GPU job: attach the communicator, push this rank's slice, finish collectively.
    make_context() returns a breakdancer_b200.api.Context on this rank's GPU; unique_id comes from rank 0
    (api.comm_unique_id(), distributed by the caller). Returns (summary, table): identical on every rank."""
    ctx = make_context()
    ctx.comm_init(unique_id, rank, world)
    sl = stream_slices(len(cols["pos"]), world)[rank]
    ctx.push({k: np.ascontiguousarray(v[sl]) for k, v in cols.items()})
    summary = ctx.summary()
    table = ctx.finish()
    return summary, table
