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


def config2_device(n_pairs: int, seed: int, device, chrom_len: int = CHR1_LEN, tid: int = 0) -> Dict[str, torch.Tensor]:
    """Position-sorted record columns (2 records per pair) on `device`.
    Columns use torch dtypes with the bit patterns bdk_soa expects: flag/rgid int16, qid int64."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    scale = n_pairs / 50_000_000
    L = chrom_len if scale >= 1 else max(200000, int(chrom_len * scale))
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
