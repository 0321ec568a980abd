"""Shared helpers of the test-suite (not product code)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from breakdancer_b200 import api, synth
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CHR21 = os.path.join(GOLDEN, "chr21")
CLI = os.path.join(ROOT, "breakdancer_b200", "bin", "breakdancer_max")

GENOME3 = [("chrA", 3000000), ("chrB", 2000000), ("chrC", 1500000)]
LIBS4 = [synth.LibSpec("normal_a", "normal.bam", 315, 44, 75, ["n_a1", "n_a2"]),
         synth.LibSpec("normal_b", "normal.bam", 312, 43, 75, ["n_b1"]),
         synth.LibSpec("tumor_a", "tumor.bam", 467, 32, 75, ["t_a1", "t_a2"], tumor=True),
         synth.LibSpec("tumor_b", "tumor.bam", 476, 29, 100, ["t_b1"], tumor=True)]

OPTION_SETS = [dict(), dict(CN_lib=True), dict(transchr_rearrange=True), dict(Illumina_long_insert=True), dict(fisher=True),
               dict(min_read_pair=1, score_threshold=0), dict(min_map_qual=10, cut_sd=2), dict(buffer_size=3), dict(chr="chrB"),
               dict(max_sd=10000), dict(min_len=50, seq_coverage_lim=2), dict(min_len=-1), dict(buffer_size=0),
               dict(min_read_pair=3), dict(CN_lib=True, print_AF=True, chr="chrA")]


def strip_header(text: str) -> str:
    return "".join(l for l in text.splitlines(True) if not (l.startswith("#Command") or l.startswith("#Software")))


def workload_bundle(w: synth.Workload, opts: api.Options):
    """(ParamBundle, columns, lib_names, bam_names, tid_names) for a synthetic workload, going through
    the real config parser (text written the way bam2cfg writes it)."""
    cfg = api.BamConfig(text=w.config_text(), cut_sd=opts.cut_sd)
    bams = sorted(set(w.rg_bamname))
    rg_lib = np.array([cfg.rg_lib(r) for r in w.rg_names], np.int32)
    rg_bam = np.array([bams.index(b) for b in w.rg_bamname], np.int32)
    cols = w.cols
    names = [g[0] for g in w.genome]
    if opts.chr:
        m = cols["tid"] == names.index(opts.chr)
        cols = {k: np.ascontiguousarray(v[m]) for k, v in cols.items()}
    b = api.ParamBundle(opts, cfg.libs, cfg.nbam, rg_lib, rg_bam, cfg.window, len(names))
    return b, cols, cfg.lib_names, cfg.bam_files, names


_hostsim = None


def hostsim_lib():
    global _hostsim
    if _hostsim is None:
        so = os.path.join(ROOT, "tests", "_build", "libhostsim.so")
        src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
        deps = [src] + [os.path.join(ROOT, "breakdancer_b200", "csrc", f) for f in ("bdk_logic.h", "bdk_finalize.h")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            os.makedirs(os.path.dirname(so), exist_ok=True)
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-Wall", "-shared", src, "-o", so])
        L = C.CDLL(so)
        L.hostsim_run.restype = C.c_int
        L.hostsim_run.argtypes = [C.POINTER(api.Params), C.POINTER(api.Soa), C.c_uint64, C.POINTER(oracle.Output)]
        L.hostsim_poisson_logsf.restype = C.c_double
        L.hostsim_poisson_logsf.argtypes = [C.c_double, C.c_int]
        L.hostsim_gamma_q.restype = C.c_double
        L.hostsim_gamma_q.argtypes = [C.c_double, C.c_double]
        L.hostsim_classify.restype = C.c_uint32
        L.hostsim_classify.argtypes = [C.c_int32] * 5 + [C.c_uint32] * 2 + [C.c_float] * 2 + [C.c_int32] * 4
        _hostsim = L
    return _hostsim


def run_hostsim(bundle: api.ParamBundle, cols: Dict[str, np.ndarray]) -> oracle.OracleResult:
    L = hostsim_lib()
    soa = api.make_soa(cols)
    out = oracle.Output()
    n = len(cols["pos"])
    rc = L.hostsim_run(C.byref(bundle.params), C.byref(soa), n, C.byref(out))
    assert rc == 0, rc
    try:
        return oracle.OracleResult(out, n, bundle.params.nlib)
    finally:
        oracle.load().bdo_free(C.byref(out))


def assert_logp_close(ref: np.ndarray, got: np.ndarray, atol: float = 1e-6):
    """Poisson log-probabilities within 1e-6 wherever the tail probability is a normal double
    (log p > -700); below that both must agree that it is below (every such score is 99)."""
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    assert ref.shape == got.shape
    fin = np.isfinite(ref) & (ref > -700)
    if fin.any():
        assert np.max(np.abs(ref[fin] - got[fin])) <= atol, np.max(np.abs(ref[fin] - got[fin]))
    assert np.all((got[~fin] < -700) | ~np.isfinite(got[~fin]))


def assert_tables_equal(ref: api.SvTable, got: api.SvTable, what: str = ""):
    assert len(ref.sv) == len(got.sv), f"{what}: {len(ref.sv)} vs {len(got.sv)} SV rows"
    for f in ("chr", "pos", "fwd", "rev", "flag", "diffspan", "score", "num_pairs", "region", "window", "order"):
        assert np.array_equal(ref.sv[f], got.sv[f]), f"{what}: sv.{f} differs"
    assert_logp_close(ref.sv["logp"], got.sv["logp"])
    a, b = ref.sv["allele_frequency"], got.sv["allele_frequency"]
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: allele_frequency NaN pattern"
    assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), f"{what}: allele_frequency"
    assert np.array_equal(ref.lib_count, got.lib_count), f"{what}: lib_count"
    assert np.array_equal(ref.cn_count, got.cn_count), f"{what}: cn_count"
    assert np.array_equal(ref.copy_number, got.copy_number), f"{what}: copy_number"


def assert_summary_equal(a: api.SummaryT, b: api.SummaryT, what: str = ""):
    """Bit-exact, except that NaNs (0/0 densities of an empty input) may differ in sign/payload:
    x86 produces the negative default NaN, the GPU the positive canonical one."""
    for name, _ in api.SummaryT._fields_:
        x = np.frombuffer(bytes(getattr(a, name)) if not isinstance(getattr(a, name), int) else np.array([getattr(a, name)]).tobytes(), np.uint8)
        y = np.frombuffer(bytes(getattr(b, name)) if not isinstance(getattr(b, name), int) else np.array([getattr(b, name)]).tobytes(), np.uint8)
        if name in ("seq_coverage", "read_density"):
            fx, fy = x.view(np.float32), y.view(np.float32)
            assert np.array_equal(np.isnan(fx), np.isnan(fy)), f"{what}: summary.{name} NaN pattern"
            assert np.array_equal(fx[~np.isnan(fx)], fy[~np.isnan(fy)]), f"{what}: summary.{name}"
        else:
            assert np.array_equal(x, y), f"{what}: summary.{name} differs (reflen {a.covered_ref_len} vs {b.covered_ref_len}, window {a.window} vs {b.window})"


def assert_result_matches_oracle(ro: oracle.OracleResult, table: api.SvTable, summary: api.SummaryT, regions=None,
                                 areads=None, read_region=None, support=None, what: str = ""):
    assert_summary_equal(ro.summary, summary, what)
    if areads is not None:
        assert np.array_equal(ro.areads, areads), f"{what}: anomalous read stream differs"
    if read_region is not None:
        assert np.array_equal(ro.aread_region, read_region), f"{what}: read -> region map differs"
    if regions is not None:
        assert np.array_equal(ro.regions, regions), f"{what}: region table differs"
    assert_tables_equal(ro.table, table, what)
    if support is not None:
        assert np.array_equal(ro.sv_of_read, support), f"{what}: supporting-read assignment differs"


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def run_gpu(bundle: api.ParamBundle, cols: Dict[str, np.ndarray], chunks: Optional[int] = None, device_push: bool = False,
            pinned: bool = False):
    """Run the CUDA path through the C ABI; returns (table, summary, regions, areads, read_region, support, ctx)."""
    ctx = api.Context(bundle, 0)
    n = len(cols["pos"])
    if pinned:   # pinned host columns: qlen / qid are read in place (zero copy), the rest goes through the copy engine
        import torch
        pin = {k: torch.from_numpy(v.view(np.int64) if v.dtype == np.uint64 else (v.view(np.int16) if v.dtype == np.uint16 else v)).pin_memory()
               for k, v in cols.items()}
        soa = api.soa_from_pointers({k: t.data_ptr() for k, t in pin.items()})
        ctx.push_soa(soa, n, device=False)
        assert n == 0 or ctx.h2d_bytes() == 25 * n, (ctx.h2d_bytes(), n)
    elif device_push:
        import torch
        dev = {k: torch.from_numpy(v.view(np.int64) if v.dtype == np.uint64 else (v.view(np.int16) if v.dtype == np.uint16 else v)).cuda()
               for k, v in cols.items()}
        soa = api.soa_from_pointers({k: t.data_ptr() for k, t in dev.items()})
        ctx.push_soa(soa, n, device=True)
    elif chunks and n:
        edges = np.linspace(0, n, chunks + 1).astype(np.int64)
        for a, b in zip(edges[:-1], edges[1:]):
            if b > a:
                ctx.push({k: np.ascontiguousarray(v[a:b]) for k, v in cols.items()})
    else:
        ctx.push(cols)
    summary = ctx.summary()
    table = ctx.finish()
    regions = ctx.regions()
    areads, rr = ctx.areads()
    support = ctx.support()
    return table, summary, regions, areads, rr, support, ctx
