"""ctypes bridge to the CPU restatement oracle (oracle/bd_oracle.cpp -> oracle/_build/libbdoracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Sequence

import numpy as np

from breakdancer_b200 import api

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libbdoracle.so")
REF_DIR = os.path.join(_HERE, "_ref")
REF_BIN = os.path.join(REF_DIR, "breakdancer-max")
REF_NODECODE = os.path.join(REF_DIR, "breakdancer-max-nodecode")   # the reference on records decoded into memory beforehand (ref_nodecode.cpp)
REF_SAMTOOLS = os.path.join(REF_DIR, "samtools")
REF_SCORE = os.path.join(REF_DIR, "score_ref")


class Output(C.Structure):
    _fields_ = [("summary", api.SummaryT), ("n_sv", C.c_uint64), ("sv", C.c_void_p), ("lib_count", C.c_void_p),
                ("cn_count", C.c_void_p), ("copy_number", C.c_void_p), ("nkey", C.c_int32),
                ("n_regions", C.c_uint64), ("regions", C.c_void_p), ("region_alive", C.c_void_p),
                ("n_areads", C.c_uint64), ("areads", C.c_void_p), ("aread_region", C.c_void_p),
                ("sv_of_read", C.c_void_p), ("rec_class", C.c_void_p), ("n_support", C.c_uint64),
                ("support_off", C.c_void_p), ("support", C.c_void_p), ("n_flush", C.c_int32)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(_HERE), "oracle/_build/libbdoracle.so"])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.bdo_run.restype = C.c_int
        L.bdo_run.argtypes = [C.POINTER(api.Params), C.POINTER(api.Soa), C.c_uint64, C.POINTER(Output)]
        L.bdo_free.argtypes = [C.POINTER(Output)]
        L.bdo_poisson_logsf.restype = C.c_double
        L.bdo_poisson_logsf.argtypes = [C.c_double, C.c_int]
        L.bdo_gamma_q.restype = C.c_double
        L.bdo_gamma_q.argtypes = [C.c_double, C.c_double]
        L.bdo_format.restype = C.c_int64
        L.bdo_format.argtypes = [C.POINTER(api.Params), C.POINTER(Output), C.c_char_p, C.c_char_p, C.c_char_p,
                                 C.c_int, C.c_char_p, C.c_int64]
        _lib = L
    return _lib


class OracleResult:
    """Everything the oracle computed for one run, copied into numpy arrays."""

    def __init__(self, out: Output, n: int, nlib: int):
        cp = api._copy_array
        self.summary = api.SummaryT.from_buffer_copy(out.summary)
        ns = out.n_sv
        self.nkey = out.nkey
        self.table = api.SvTable(
            cp(out.sv, ns, api.SV_DTYPE),
            cp(out.lib_count, ns * nlib, np.dtype(np.int32)).reshape(ns, nlib),
            cp(out.cn_count, ns * out.nkey, np.dtype(np.uint32)).reshape(ns, out.nkey),
            cp(out.copy_number, ns * out.nkey, np.dtype(np.float32)).reshape(ns, out.nkey), out.nkey)
        self.regions = cp(out.regions, out.n_regions, api.REGION_DTYPE)
        self.region_alive = cp(out.region_alive, out.n_regions, np.dtype(np.uint8))
        self.areads = cp(out.areads, out.n_areads, api.AREAD_DTYPE)
        self.aread_region = cp(out.aread_region, out.n_areads, np.dtype(np.int32))
        self.sv_of_read = cp(out.sv_of_read, out.n_areads, np.dtype(np.int32))
        self.rec_class = cp(out.rec_class, n, np.dtype(np.uint8))
        self.support_off = cp(out.support_off, ns + 1, np.dtype(np.uint64))
        self.support = cp(out.support, out.n_support, np.dtype(np.uint32))
        self.n_flush = out.n_flush


def run(bundle: api.ParamBundle, cols: Dict[str, np.ndarray]) -> OracleResult:
    L = load()
    soa = api.make_soa(cols)
    out = Output()
    n = len(cols["pos"])
    rc = L.bdo_run(C.byref(bundle.params), C.byref(soa), n, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle failed: {rc}")
    try:
        return OracleResult(out, n, bundle.params.nlib)
    finally:
        L.bdo_free(C.byref(out))


def run_text(bundle: api.ParamBundle, cols: Dict[str, np.ndarray], lib_names: Sequence[str],
             bam_names: Sequence[str], tid_names: Sequence[str]):
    """(OracleResult, stdout text from '#Library Statistics:' on) using the oracle's own formatter."""
    L = load()
    soa = api.make_soa(cols)
    out = Output()
    n = len(cols["pos"])
    rc = L.bdo_run(C.byref(bundle.params), C.byref(soa), n, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle failed: {rc}")
    try:
        j = lambda v: ("\n".join(v) + "\n").encode()
        cap = 1 << 20
        while True:
            buf = C.create_string_buffer(cap)
            k = L.bdo_format(C.byref(bundle.params), C.byref(out), j(lib_names), j(bam_names), j(tid_names),
                             int(bundle.opts.print_AF), buf, cap)
            if k < cap:
                break
            cap = k + 1
        return OracleResult(out, n, bundle.params.nlib), buf.value.decode()
    finally:
        L.bdo_free(C.byref(out))


def poisson_logsf(lam: float, k: int) -> float:
    return load().bdo_poisson_logsf(lam, k)


def have_reference() -> bool:
    return os.access(REF_BIN, os.X_OK)


def run_reference(args: Sequence[str], cwd: str, timeout: float = 3600.0) -> str:
    """stdout of the unmodified reference binary, minus the #Software/#Command lines."""
    p = subprocess.run([REF_BIN] + list(args), cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       timeout=timeout, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"reference failed rc={p.returncode}: {p.stderr[-2000:]}")
    return "".join(l for l in p.stdout.splitlines(True) if not (l.startswith("#Command") or l.startswith("#Software")))
