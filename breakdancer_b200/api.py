"""ctypes binding of the bdk C ABI (include/bdk.h, include/bdk_host.h).

The product is the shared library ``breakdancer_b200/libbdk.so`` (CUDA kernels for sm_100a behind a
C ABI, plus the C++ host side: config parser, BAM decoder, formatter) and the drop-in executable
``breakdancer_b200/bin/breakdancer_max``.  This module is only the thin Python face used by the
tests and by bench.py; it contains no algorithmic code and has NO fallback: if the library is
missing, importing anything that needs it raises.

Names mirror the reference's classes for this path: ``Options`` (src/lib/common/Options.hpp:12-71),
``BamConfig`` (src/lib/io/BamConfig.hpp:15-48), ``LibraryConfig`` (LibraryConfig.hpp:11-27),
``BamSummary`` (BamSummary.hpp:16-57).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbdk.so")

NUM_FLAGS = 11
MAX_LIBS = 255
MAX_BAMS = 64
FLAG_NAMES = ["NA", "ARP_FF", "ARP_LARGE_INSERT", "ARP_SMALL_INSERT", "ARP_RF", "ARP_RR",
              "NORMAL_FR", "NORMAL_RF", "ARP_CTX", "MATE_UNMAPPED", "UNMAPPED"]


class Lib(C.Structure):
    _fields_ = [("mean_insertsize", C.c_float), ("std_insertsize", C.c_float), ("uppercutoff", C.c_float),
                ("lowercutoff", C.c_float), ("readlens", C.c_float), ("min_mapping_quality", C.c_int32),
                ("bam_index", C.c_int32)]


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "min_len", "max_sd", "min_map_qual", "min_read_pair", "seq_coverage_lim", "buffer_size",
        "score_threshold", "transchr_rearrange", "fisher", "illumina_long_insert", "cn_lib",
        "chr_restricted", "initial_window", "nlib", "nbam", "nrg", "ntid")] + [
        ("libs", C.POINTER(Lib)), ("rg_lib", C.POINTER(C.c_int32)), ("rg_bam", C.POINTER(C.c_int32))]


class Soa(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("mpos", C.c_void_p), ("tid", C.c_void_p), ("mtid", C.c_void_p),
                ("isize", C.c_void_p), ("flag", C.c_void_p), ("mapq", C.c_void_p), ("rgid", C.c_void_p),
                ("qlen", C.c_void_p), ("qid", C.c_void_p)]


COLUMN_DTYPES = {"pos": np.int32, "mpos": np.int32, "tid": np.int32, "mtid": np.int32, "isize": np.int32,
                 "flag": np.uint16, "mapq": np.uint8, "rgid": np.uint16, "qlen": np.int32, "qid": np.uint64}
HOT_COLUMNS = ("pos", "mpos", "tid", "mtid", "isize", "flag", "mapq", "rgid")  # 25 B / record


class SummaryT(C.Structure):
    _fields_ = [("covered_ref_len", C.c_uint32), ("window", C.c_int32), ("n_records", C.c_uint64),
                ("n_anomalous", C.c_uint64), ("read_count_per_bam", C.c_uint32 * MAX_BAMS),
                ("ref_len_per_bam", C.c_uint64 * MAX_BAMS), ("lib_read_count", C.c_uint32 * MAX_LIBS),
                ("read_counts_by_flag", (C.c_uint32 * NUM_FLAGS) * MAX_LIBS),
                ("seq_coverage", C.c_float * MAX_LIBS), ("read_density", C.c_float * MAX_LIBS)]


class Sv(C.Structure):
    _fields_ = [("chr", C.c_int32 * 2), ("pos", C.c_int32 * 2), ("fwd", C.c_int32 * 2), ("rev", C.c_int32 * 2),
                ("flag", C.c_int32), ("diffspan", C.c_int32), ("score", C.c_int32), ("num_pairs", C.c_int32),
                ("logp", C.c_double), ("allele_frequency", C.c_float), ("cn_present", C.c_uint32),
                ("region", C.c_int32 * 2), ("window", C.c_int32), ("order", C.c_int32)]


SV_DTYPE = np.dtype([("chr", np.int32, 2), ("pos", np.int32, 2), ("fwd", np.int32, 2), ("rev", np.int32, 2),
                     ("flag", np.int32), ("diffspan", np.int32), ("score", np.int32), ("num_pairs", np.int32),
                     ("logp", np.float64), ("allele_frequency", np.float32), ("cn_present", np.uint32),
                     ("region", np.int32, 2), ("window", np.int32), ("order", np.int32)])
assert SV_DTYPE.itemsize == C.sizeof(Sv)

REGION_DTYPE = np.dtype([("tid", np.int32), ("start", np.int32), ("end", np.int32), ("fwd", np.int32),
                         ("rev", np.int32), ("first_read", np.int32), ("n_reads", np.int32),
                         ("stored", np.int32), ("window", np.int32)])
AREAD_DTYPE = np.dtype([("pos", np.int32), ("tid", np.int32), ("qlen", np.int32), ("abs_isize", np.int32),
                        ("meta", np.uint32), ("record", np.uint32), ("qid", np.uint64)])


class Result(C.Structure):
    _fields_ = [("n_sv", C.c_uint64), ("sv", C.POINTER(Sv)), ("lib_count", C.POINTER(C.c_int32)),
                ("cn_count", C.POINTER(C.c_uint32)), ("copy_number", C.POINTER(C.c_float)), ("nkey", C.c_int32)]


@dataclasses.dataclass
class Options:
    """Command-line options with the reference's defaults (Options.cpp:27-40; note -y 30)."""
    chr: str = ""                 # -o
    min_len: int = 7              # -s
    cut_sd: int = 3               # -c
    max_sd: int = 1000000000      # -m
    min_map_qual: int = 35        # -q
    min_read_pair: int = 2        # -r
    seq_coverage_lim: int = 1000  # -x
    buffer_size: int = 100        # -b
    transchr_rearrange: bool = False  # -t
    fisher: bool = False          # -f
    Illumina_long_insert: bool = False  # -l
    CN_lib: bool = False          # -a
    print_AF: bool = False        # -h
    score_threshold: int = 30     # -y


class Packed(C.Structure):
    """bdk_packed (include/bdk.h): one run of records in the 12-byte wire format."""
    _fields_ = [("pos", C.c_void_p), ("meta", C.c_void_p), ("rel", C.c_void_p), ("qlen", C.c_void_p), ("qid", C.c_void_p),
                ("tid", C.c_int32), ("reserved", C.c_uint32), ("nx", C.c_uint64),
                ("x_index", C.c_void_p), ("x_mpos", C.c_void_p), ("x_mtid", C.c_void_p), ("x_isize", C.c_void_p),
                ("x_flag", C.c_void_p), ("x_rgid", C.c_void_p)]


class BamSource(C.Structure):
    """bdk_bam_source (include/bdk.h): one BAM file for the device-resident decode."""
    _fields_ = [("file", C.c_void_p), ("file_bytes", C.c_uint64), ("members", C.c_void_p), ("n_members", C.c_uint64),
                ("first_record", C.c_uint64), ("end_offset", C.c_uint64), ("n_ref", C.c_int32),
                ("region_on", C.c_int32), ("region_tid", C.c_int32), ("region_beg", C.c_int32), ("region_end", C.c_int32),
                ("n_rg", C.c_uint32), ("rg_hash", C.c_void_p), ("rg_id", C.c_void_p), ("rg_other", C.c_uint16), ("reserved", C.c_uint16),
                ("window_bytes", C.c_uint64)]


class BamStats(C.Structure):
    """bdk_bam_stats (include/bdk.h)."""
    _fields_ = [("records", C.c_uint64), ("kept", C.c_uint64), ("h2d_bytes", C.c_uint64), ("inflated_bytes", C.c_uint64),
                ("windows", C.c_uint32), ("guess_misses", C.c_uint32), ("sorted", C.c_int32),
                ("inflate_ms", C.c_float), ("chain_ms", C.c_float), ("extract_ms", C.c_float), ("stage_ms", C.c_float), ("wall_ms", C.c_float),
                ("merge_parts", C.c_uint32), ("merge_longest_part", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class PackedRun:
    """A packed run (bdk_pack) and the buffers behind it; keeps the source columns alive (pos / qlen / qid are used in place)."""

    def __init__(self, soa: "Soa", n: int, keep=None, threads: int = 0):
        L = load_library()
        self.view, self._buf, self.n, self._keep = Packed(), C.c_void_p(), n, keep
        rc = L.bdk_pack(C.byref(soa), n, threads, C.byref(self._buf), C.byref(self.view))
        if rc != 0:
            raise BdkError(f"bdk_pack failed ({rc}): {L.bdk_last_error(None).decode()}")

    def close(self):
        if self._buf:
            load_library().bdk_pack_free(self._buf)
            self._buf = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_runs(cols: Dict[str, np.ndarray]) -> List["PackedRun"]:
    """Packed runs of (tid, pos)-sorted columns, one per reference sequence, in stream order."""
    tid = cols["tid"]
    n = len(tid)
    if n == 0:
        return []
    edges = [0] + (np.flatnonzero(np.diff(tid)) + 1).tolist() + [n]
    runs = []
    for a, b in zip(edges[:-1], edges[1:]):
        sub = {k: np.ascontiguousarray(v[a:b]) for k, v in cols.items()}
        runs.append(PackedRun(make_soa(sub), b - a, keep=sub))
    return runs


_lib: Optional[C.CDLL] = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Load libbdk.so and declare every prototype. Raises if the library was not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} not found: build it with `make` (or __graft_entry__.build()); "
                           "there is no CPU fallback for the bdk hot path")
    L = C.CDLL(p)
    vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
    sig = {
        "bdk_create": (C.c_int, [C.POINTER(vp), C.c_int, C.POINTER(Params)]),
        "bdk_destroy": (None, [vp]),
        "bdk_last_error": (C.c_char_p, [vp]),
        "bdk_set_stream": (C.c_int, [vp, vp]),
        "bdk_reset": (C.c_int, [vp]),
        "bdk_push": (C.c_int, [vp, C.POINTER(Soa), u64]),
        "bdk_push_device": (C.c_int, [vp, C.POINTER(Soa), u64]),
        "bdk_push_packed": (C.c_int, [vp, C.POINTER(Packed), u64]),
        "bdk_pack": (C.c_int, [C.POINTER(Soa), u64, C.c_int, C.POINTER(vp), C.POINTER(Packed)]),
        "bdk_pack_free": (None, [vp]),
        "bdk_summary": (C.c_int, [vp, C.POINTER(SummaryT)]),
        "bdk_finish": (C.c_int, [vp, C.POINTER(Result)]),
        "bdk_get_regions": (C.c_int, [vp, C.POINTER(vp), C.POINTER(u64)]),
        "bdk_get_areads": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]),
        "bdk_get_support": (C.c_int, [vp, C.POINTER(vp), C.POINTER(u64)]),
        "bdk_kernel_times": (C.c_int, [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]),
        "bdk_kernel_launches": (u64, [vp]),
        "bdk_h2d_bytes": (u64, [vp]),
        "bdk_d2h_bytes": (u64, [vp]),
        "bdk_host_alloc": (vp, [u64]),
        "bdk_host_free": (None, [vp]),
        "bdk_set_comm": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "bdk_comm_unique_id": (C.c_int, [vp, C.c_int]),
        "bdk_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "bdk_comm_bytes": (u64, [vp]),
        "bdk_k4_sweeps": (C.c_uint32, [vp]),
        "bdk_duplicate_names": (C.c_uint32, [vp]),
        "bdk_poisson_logsf": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(i32), C.POINTER(C.c_double), u64]),
        "bdk_version": (C.c_char_p, []),
        "bdk_push_bam": (C.c_int, [vp, C.POINTER(BamSource), C.POINTER(BamStats)]),
        "bdk_hash_bytes": (u64, [C.c_void_p, u64]),
        "bdh_bamdev_open": (vp, [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
        "bdh_bamdev_free": (None, [vp]),
        "bdh_bamdev_nrg": (C.c_int, [vp]),
        "bdh_bamdev_rg_lib": (C.POINTER(i32), [vp]),
        "bdh_bamdev_rg_bam": (C.POINTER(i32), [vp]),
        "bdh_bamdev_ntid": (C.c_int, [vp]),
        "bdh_bamdev_tid_name": (C.c_char_p, [vp, C.c_int]),
        "bdh_bamdev_members": (u64, [vp]),
        "bdh_bamdev_file_bytes": (u64, [vp]),
        "bdh_bamdev_push": (C.c_int, [vp, vp, C.POINTER(BamStats)]),
        "bdh_bamdev_decode": (C.c_int, [vp, vp, C.POINTER(Soa), u64, C.POINTER(BamStats)]),
        "bdh_bamdev_open_next": (vp, [vp, vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
        "bdh_bamdev_push2": (C.c_int, [vp, vp, vp, C.POINTER(BamStats)]),
        "bdh_bamdev_pushn": (C.c_int, [C.POINTER(vp), C.c_int, vp, C.POINTER(BamStats)]),
        "bdh_bamdev_decoden": (C.c_int, [C.POINTER(vp), C.c_int, vp, C.POINTER(Soa), u64, C.POINTER(BamStats)]),
        "bdh_bamdev_decode2": (C.c_int, [vp, vp, vp, C.POINTER(Soa), u64, C.POINTER(BamStats)]),
        "bdk_push_bams": (C.c_int, [vp, C.POINTER(BamSource), C.c_int, C.POINTER(BamStats)]),
        "bdk_decode_bams": (C.c_int, [vp, C.POINTER(BamSource), C.c_int, C.POINTER(Soa), u64, C.POINTER(BamStats)]),
        "bdk_decode_bam": (C.c_int, [vp, C.POINTER(BamSource), C.POINTER(Soa), u64, C.POINTER(BamStats)]),
        "bdk_bgzf_inflate": (C.c_int, [C.c_int, C.c_void_p, u64, C.c_void_p, u64, C.c_void_p, u64, C.POINTER(i32), C.POINTER(C.c_float)]),
        "bdh_config_parse": (vp, [C.c_char_p, C.c_int, C.c_char_p, C.c_int]),
        "bdh_config_load": (vp, [C.c_char_p, C.c_int, C.c_char_p, C.c_int]),
        "bdh_config_free": (None, [vp]),
        "bdh_config_nlib": (C.c_int, [vp]),
        "bdh_config_nbam": (C.c_int, [vp]),
        "bdh_config_window": (C.c_int, [vp]),
        "bdh_config_libs": (C.POINTER(Lib), [vp]),
        "bdh_config_lib_name": (C.c_char_p, [vp, C.c_int]),
        "bdh_config_bam_name": (C.c_char_p, [vp, C.c_int]),
        "bdh_config_rg_lib": (C.c_int, [vp, C.c_char_p]),
        "bdh_config_translate_token": (C.c_int, [C.c_char_p]),
        "bdh_stream_open": (vp, [vp, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]),
        "bdh_stream_free": (None, [vp]),
        "bdh_stream_n": (u64, [vp]),
        "bdh_stream_cols": (None, [vp, C.POINTER(Soa)]),
        "bdh_stream_nrg": (C.c_int, [vp]),
        "bdh_stream_rg_lib": (C.POINTER(i32), [vp]),
        "bdh_stream_rg_bam": (C.POINTER(i32), [vp]),
        "bdh_stream_ntid": (C.c_int, [vp]),
        "bdh_stream_tid_name": (C.c_char_p, [vp, C.c_int]),
        "bdh_stream_qname": (C.c_char_p, [vp, u64]),
        "bdh_stream_fastq": (C.c_int, [vp, u64, C.c_char_p, C.c_int]),
        "bdh_bai_reference_stats": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int, C.c_char_p, C.c_int]),
        "bdh_inflate_counters": (None, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "bdh_stream_sorted": (C.c_int, [vp]),
        "bdh_stream_timings": (None, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "bdh_write_bam": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.c_int,
                                    C.POINTER(C.c_char_p), C.POINTER(Soa), u64, C.c_char_p, C.c_int, C.c_int,
                                    C.c_int, C.c_char_p, C.c_int]),
        "bdh_bam2cfg_defaults": (None, [vp]),
        "bdh_bam2cfg": (C.c_int64, [C.POINTER(C.c_char_p), C.c_int, vp, C.c_char_p, C.c_int64, C.c_char_p, C.c_int]),
        "bdh_format_header": (C.c_int64, [C.POINTER(Params), C.POINTER(SummaryT), vp, C.c_int, C.c_char_p, C.c_int64]),
        "bdh_format_rows": (C.c_int64, [C.POINTER(Params), C.POINTER(Result), vp, vp, C.c_int, C.POINTER(C.c_int),
                                         C.c_char_p, C.c_int64]),
    }
    missing = []
    for name, (res, args) in sig.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    L._bdk_missing = missing
    L._bdk_symbols = list(sig)
    if path is None:
        _lib = L
    return L


# ------------------------------------------------------------------------------------------------
# columns helpers
# ------------------------------------------------------------------------------------------------
def make_soa(cols: Dict[str, np.ndarray]) -> Soa:
    """Soa struct pointing at contiguous numpy columns of the right dtype (no copy)."""
    s = Soa()
    n = None
    for name, dt in COLUMN_DTYPES.items():
        a = cols[name]
        if a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
            raise TypeError(f"column {name} must be contiguous {np.dtype(dt)}")
        if n is None:
            n = len(a)
        elif len(a) != n:
            raise ValueError("columns differ in length")
        setattr(s, name, a.ctypes.data)
    return s


def soa_from_pointers(ptrs: Dict[str, int]) -> Soa:
    s = Soa()
    for name in COLUMN_DTYPES:
        setattr(s, name, ptrs[name])
    return s


class BamConfig:
    """Parsed bam2cfg configuration (BamConfig, src/lib/io/BamConfig.cpp:19-122)."""

    def __init__(self, text: Optional[str] = None, path: Optional[str] = None, cut_sd: int = 3):
        L = load_library()
        err = C.create_string_buffer(512)
        if text is not None:
            self._h = L.bdh_config_parse(text.encode(), cut_sd, err, 512)
        else:
            self._h = L.bdh_config_load(path.encode(), cut_sd, err, 512)
        if not self._h:
            raise RuntimeError(err.value.decode())
        self._L = L
        self.nlib = L.bdh_config_nlib(self._h)
        self.nbam = L.bdh_config_nbam(self._h)
        self.window = L.bdh_config_window(self._h)
        libs = L.bdh_config_libs(self._h)
        self.libs = [Lib.from_buffer_copy(libs[i]) for i in range(self.nlib)]
        self.lib_names = [L.bdh_config_lib_name(self._h, i).decode() for i in range(self.nlib)]
        self.bam_files = [L.bdh_config_bam_name(self._h, i).decode() for i in range(self.nbam)]

    def rg_lib(self, rg: str) -> int:
        return self._L.bdh_config_rg_lib(self._h, rg.encode())

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.bdh_config_free(self._h)
            self._h = None


class BamStream:
    """Decoded, merged, position-sorted record stream of the config's bams
    (openBams + BamMerger + AlignmentSource)."""

    def __init__(self, cfg: BamConfig, paths: Optional[Sequence[str]] = None, region: str = "",
                 threads: int = 0, pinned: bool = False, keep_records: bool = False):
        L = load_library()
        err = C.create_string_buffer(512)
        if paths is not None:
            arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
            h = L.bdh_stream_open(cfg._h, arr, len(paths), region.encode(), threads, int(pinned), int(keep_records), err, 512)
        else:
            h = L.bdh_stream_open(cfg._h, None, 0, region.encode(), threads, int(pinned), int(keep_records), err, 512)
        if not h:
            raise RuntimeError(err.value.decode())
        self._h, self._L, self.cfg = h, L, cfg
        self.n = L.bdh_stream_n(h)
        soa = Soa()
        L.bdh_stream_cols(h, C.byref(soa))
        self.soa = soa
        self.cols = {}
        for name, dt in COLUMN_DTYPES.items():
            ptr = getattr(soa, name)
            buf = (C.c_char * (self.n * np.dtype(dt).itemsize)).from_address(ptr) if self.n else b""
            self.cols[name] = np.frombuffer(buf, dtype=dt, count=self.n)
        nrg = L.bdh_stream_nrg(h)
        self.rg_lib = np.ctypeslib.as_array(L.bdh_stream_rg_lib(h), (nrg,)).copy() if nrg else np.zeros(0, np.int32)
        self.rg_bam = np.ctypeslib.as_array(L.bdh_stream_rg_bam(h), (nrg,)).copy() if nrg else np.zeros(0, np.int32)
        self.tid_names = [L.bdh_stream_tid_name(h, i).decode() for i in range(L.bdh_stream_ntid(h))]

    def qname(self, i: int) -> str:
        return self._L.bdh_stream_qname(self._h, i).decode()

    def fastq(self, i: int) -> str:
        buf = C.create_string_buffer(1 << 16)
        n = self._L.bdh_stream_fastq(self._h, i, buf, 1 << 16)
        if n < 0:
            raise RuntimeError("record not kept")
        return buf.value.decode()

    def timings(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._L.bdh_stream_timings(self._h, C.byref(a), C.byref(b), C.byref(c))
        return {"inflate_s": a.value, "extract_s": b.value, "merge_s": c.value}

    def close(self):
        if getattr(self, "_h", None):
            self.cols = {}
            self._L.bdh_stream_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


class BamDevice:
    """One BAM file opened for the device-resident decode (bdh_bamdev_*, include/bdk_host.h): the host maps the file, lists
    its BGZF members and parses the header; Context.push_bam() then inflates, parses and classifies it on the GPU."""

    def __init__(self, cfg: BamConfig, path: str = "", region: str = "", after: Optional["BamDevice"] = None):
        L = load_library()
        err = C.create_string_buffer(512)
        if after is not None:       # the second bam of a two-bam run: read-group ids behind the first bam's
            h = L.bdh_bamdev_open_next(cfg._h, after._h, path.encode(), region.encode(), err, 512)
        else:
            h = L.bdh_bamdev_open(cfg._h, path.encode(), region.encode(), err, 512)
        if not h:
            raise RuntimeError(err.value.decode())
        self._h, self._L, self.cfg = h, L, cfg
        nrg = L.bdh_bamdev_nrg(h)
        self.rg_lib = np.ctypeslib.as_array(L.bdh_bamdev_rg_lib(h), (nrg,)).copy()
        self.rg_bam = np.ctypeslib.as_array(L.bdh_bamdev_rg_bam(h), (nrg,)).copy()
        self.tid_names = [L.bdh_bamdev_tid_name(h, i).decode() for i in range(L.bdh_bamdev_ntid(h))]
        self.n_members = int(L.bdh_bamdev_members(h))
        self.file_bytes = int(L.bdh_bamdev_file_bytes(h))

    def bundle(self, opts: "Options") -> "ParamBundle":
        return ParamBundle(opts, self.cfg.libs, self.cfg.nbam, self.rg_lib, self.rg_bam, self.cfg.window, max(1, len(self.tid_names)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.bdh_bamdev_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


def hash_bytes(b: bytes) -> int:
    """The decoders' 64-bit key of a byte string (read names, read-group strings)."""
    return int(load_library().bdk_hash_bytes(b, len(b)))


def bam_reference_names(bam_path: str) -> List[str]:
    """Reference sequence names of a bam's header; only the BGZF members that hold the header are read."""
    import struct
    import zlib
    raw = b""
    with open(bam_path, "rb") as f:
        while True:
            head = f.read(18)
            if len(head) < 18 or head[:4] != b"\x1f\x8b\x08\x04":
                raise RuntimeError(bam_path + " is not a valid bam file")
            xlen, = struct.unpack_from("<H", head, 10)
            bsize = struct.unpack_from("<H", head, 16)[0] + 1          # BC is the first extra field in every writer we know
            body = f.read(bsize - 18)
            raw += zlib.decompress(body[xlen - 6:len(body) - 8], -15)
            if len(raw) >= 12:
                if raw[:4] != b"BAM\x01":
                    raise RuntimeError(bam_path + " is not a valid bam file")
                o = 8 + struct.unpack_from("<I", raw, 4)[0]
                if o + 4 <= len(raw):
                    n_ref, = struct.unpack_from("<I", raw, o)
                    o += 4
                    names = []
                    while len(names) < n_ref and o + 4 <= len(raw):
                        l_name, = struct.unpack_from("<I", raw, o)
                        if o + 4 + l_name + 4 > len(raw):
                            break
                        names.append(raw[o + 4:o + 4 + l_name - 1].decode())
                        o += 4 + l_name + 4
                    if len(names) == n_ref:
                        return names


def bai_reference_stats(bam_path: str):
    """(records, bytes) per reference sequence from the bam's index, or None without one (include/bdk_host.h)."""
    L = load_library()
    err = C.create_string_buffer(512)
    n = L.bdh_bai_reference_stats(bam_path.encode(), None, None, 0, err, 512)
    if n == -1:
        return None
    if n < 0:
        raise RuntimeError(err.value.decode())
    rec = np.zeros(max(n, 1), dtype=np.int64)
    byt = np.zeros(max(n, 1), dtype=np.int64)
    L.bdh_bai_reference_stats(bam_path.encode(), rec.ctypes.data_as(C.POINTER(C.c_int64)), byt.ctypes.data_as(C.POINTER(C.c_int64)), n, err, 512)
    return rec[:n], byt[:n]


def inflate_counters():
    """(members the host's fast decoder handed to zlib, members the GPU decoder got wrong and the host redid), process-wide."""
    a, b = C.c_uint64(), C.c_uint64()
    load_library().bdh_inflate_counters(C.byref(a), C.byref(b))
    return a.value, b.value


class ParamBundle:
    """bdk_params plus the Python objects that own its arrays."""

    def __init__(self, opts: Options, libs: Sequence[Lib], nbam: int, rg_lib, rg_bam, initial_window: int,
                 ntid: int):
        self.opts = opts
        self.libs = (Lib * max(1, len(libs)))(*libs)
        self.rg_lib = np.ascontiguousarray(rg_lib, dtype=np.int32)
        self.rg_bam = np.ascontiguousarray(rg_bam, dtype=np.int32)
        p = Params()
        p.min_len, p.max_sd, p.min_map_qual = opts.min_len, opts.max_sd, opts.min_map_qual
        p.min_read_pair, p.seq_coverage_lim, p.buffer_size = opts.min_read_pair, opts.seq_coverage_lim, opts.buffer_size
        p.score_threshold = opts.score_threshold
        p.transchr_rearrange, p.fisher = int(opts.transchr_rearrange), int(opts.fisher)
        p.illumina_long_insert, p.cn_lib = int(opts.Illumina_long_insert), int(opts.CN_lib)
        p.chr_restricted = int(bool(opts.chr))
        p.initial_window = initial_window
        p.nlib, p.nbam, p.nrg, p.ntid = len(libs), nbam, len(self.rg_lib), ntid
        p.libs = C.cast(self.libs, C.POINTER(Lib))
        p.rg_lib = self.rg_lib.ctypes.data_as(C.POINTER(C.c_int32))
        p.rg_bam = self.rg_bam.ctypes.data_as(C.POINTER(C.c_int32))
        self.params = p
        self.nkey = len(libs) if opts.CN_lib else nbam

    @classmethod
    def from_stream(cls, opts: Options, cfg: BamConfig, stream: BamStream) -> "ParamBundle":
        return cls(opts, cfg.libs, cfg.nbam, stream.rg_lib, stream.rg_bam, cfg.window, len(stream.tid_names))


class BdkError(RuntimeError):
    pass


class Context:
    """One per-GPU bdk context (bdk_create .. bdk_destroy)."""

    def __init__(self, bundle: ParamBundle, device: int = 0):
        self._L = load_library()
        self.bundle = bundle
        h = C.c_void_p()
        rc = self._L.bdk_create(C.byref(h), device, C.byref(bundle.params))
        if rc != 0:
            raise BdkError(f"bdk_create failed ({rc}): {self._L.bdk_last_error(None).decode()}")
        self._h = h

    def _check(self, rc, what):
        if rc != 0:
            raise BdkError(f"{what} failed ({rc}): {self._L.bdk_last_error(self._h).decode()}")

    def set_stream(self, cuda_stream: int):
        self._check(self._L.bdk_set_stream(self._h, C.c_void_p(cuda_stream)), "bdk_set_stream")

    def reset(self):
        self._check(self._L.bdk_reset(self._h), "bdk_reset")

    def push(self, cols: Dict[str, np.ndarray]):
        n = len(cols["pos"])
        soa = make_soa(cols)
        self._check(self._L.bdk_push(self._h, C.byref(soa), n), "bdk_push")

    def push_packed(self, run: "PackedRun"):
        self._check(self._L.bdk_push_packed(self._h, C.byref(run.view), run.n), "bdk_push_packed")

    def push_bam(self, dev: "BamDevice") -> Dict[str, float]:
        """Decode the file on this GPU and classify its records (bdk_push_bam); returns bdk_bam_stats as a dict."""
        st = BamStats()
        self._check(self._L.bdh_bamdev_push(dev._h, self._h, C.byref(st)), "bdk_push_bam")
        return st.as_dict()

    def push_bams(self, first: "BamDevice", second: "BamDevice"):
        """Two bams decoded on this GPU, merged there in BamMerger's order and classified (bdk_push_bams); the two stats dicts."""
        st = (BamStats * 2)()
        self._check(self._L.bdh_bamdev_push2(first._h, second._h, self._h, st), "bdk_push_bams")
        return st[0].as_dict(), st[1].as_dict()

    def decode_bams(self, first: "BamDevice", second: "BamDevice", cap: int):
        """The merged records of two bams as host columns (bdk_decode_bams): (columns, stats of the two files)."""
        cols = {k: np.empty(cap, dt) for k, dt in COLUMN_DTYPES.items()}
        soa = make_soa(cols)
        st = (BamStats * 2)()
        self._check(self._L.bdh_bamdev_decode2(first._h, second._h, self._h, C.byref(soa), cap, st), "bdk_decode_bams")
        n = st[0].kept + st[1].kept
        return {k: v[:n] for k, v in cols.items()}, (st[0].as_dict(), st[1].as_dict())

    def push_bams_n(self, devs):
        """Any number of bams (config order) decoded on this GPU, merged in BamMerger's order and classified (bdk_push_bams):
        one stats dict per file. Three or more: the merge order is the priority queue's, computed on the host from device keys."""
        n = len(devs)
        hs = (C.c_void_p * n)(*[d._h for d in devs])
        st = (BamStats * n)()
        self._check(self._L.bdh_bamdev_pushn(hs, n, self._h, st), "bdk_push_bams")
        return [st[i].as_dict() for i in range(n)]

    def decode_bams_n(self, devs, cap: int):
        """The merged records of any number of bams as host columns (bdk_decode_bams): (columns, stats per file)."""
        n = len(devs)
        cols = {k: np.empty(cap, dt) for k, dt in COLUMN_DTYPES.items()}
        soa = make_soa(cols)
        hs = (C.c_void_p * n)(*[d._h for d in devs])
        st = (BamStats * n)()
        self._check(self._L.bdh_bamdev_decoden(hs, n, self._h, C.byref(soa), cap, st), "bdk_decode_bams")
        kept = sum(st[i].kept for i in range(n))
        return {k: v[:kept] for k, v in cols.items()}, [st[i].as_dict() for i in range(n)]

    def decode_bam(self, dev: "BamDevice", cap: int):
        """The file's records decoded on this GPU, as host columns (bdk_decode_bam): (columns, stats)."""
        cols = {k: np.empty(cap, dt) for k, dt in COLUMN_DTYPES.items()}
        soa = make_soa(cols)
        st = BamStats()
        self._check(self._L.bdh_bamdev_decode(dev._h, self._h, C.byref(soa), cap, C.byref(st)), "bdk_decode_bam")
        return {k: v[:st.kept] for k, v in cols.items()}, st.as_dict()

    def push_soa(self, soa: Soa, n: int, device: bool):
        fn = self._L.bdk_push_device if device else self._L.bdk_push
        self._check(fn(self._h, C.byref(soa), n), "bdk_push_device" if device else "bdk_push")

    def summary(self) -> SummaryT:
        s = SummaryT()
        self._check(self._L.bdk_summary(self._h, C.byref(s)), "bdk_summary")
        return s

    def finish(self) -> "SvTable":
        r = Result()
        self._check(self._L.bdk_finish(self._h, C.byref(r)), "bdk_finish")
        return SvTable.from_result(r, self.bundle.params.nlib)

    def finish_raw(self) -> Result:
        r = Result()
        self._check(self._L.bdk_finish(self._h, C.byref(r)), "bdk_finish")
        return r

    def regions(self) -> np.ndarray:
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._L.bdk_get_regions(self._h, C.byref(p), C.byref(n)), "bdk_get_regions")
        return _copy_array(p.value, n.value, REGION_DTYPE)

    def areads(self):
        p, q, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._check(self._L.bdk_get_areads(self._h, C.byref(p), C.byref(q), C.byref(n)), "bdk_get_areads")
        return _copy_array(p.value, n.value, AREAD_DTYPE), _copy_array(q.value, n.value, np.dtype(np.int32))

    def support(self) -> np.ndarray:
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._L.bdk_get_support(self._h, C.byref(p), C.byref(n)), "bdk_get_support")
        return _copy_array(p.value, n.value, np.dtype(np.int32))

    def kernel_times(self) -> Dict[str, Dict[str, float]]:
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        cnt = (C.c_int * cap)()
        k = self._L.bdk_kernel_times(self._h, names, ms, cnt, cap)
        return {names[i].decode(): {"ms": ms[i], "launches": cnt[i]} for i in range(k)}

    def kernel_launches(self) -> int:
        return int(self._L.bdk_kernel_launches(self._h))

    def h2d_bytes(self) -> int:
        """Bytes the last push() copied host -> device."""
        return int(self._L.bdk_h2d_bytes(self._h))

    def d2h_bytes(self) -> int:
        """Bytes the last finish() copied device -> host."""
        return int(self._L.bdk_d2h_bytes(self._h))

    # ---- multi-GPU whole-genome mode (include/bdk.h: bdk_comm_*) ---------------------------------------
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        """Create this context's NCCL communicator from the 128-byte id rank 0 obtained with comm_unique_id()."""
        _prefer_bundled_nccl()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self._L.bdk_comm_init(self._h, C.cast(buf, C.c_void_p), rank, nranks), "bdk_comm_init")

    def comm_init_from_dist(self):
        """Same, with torch.distributed (any backend) carrying the id from rank 0 to the other ranks."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.comm_init(box[0], rank, world)

    def comm_bytes(self) -> int:
        return int(self._L.bdk_comm_bytes(self._h))

    def k4_sweeps(self) -> int:
        return int(self._L.bdk_k4_sweeps(self._h))

    def duplicate_names(self) -> int:
        return int(self._L.bdk_duplicate_names(self._h))

    def poisson_logsf(self, lam: np.ndarray, k: np.ndarray) -> np.ndarray:
        lam = np.ascontiguousarray(lam, np.float64)
        k = np.ascontiguousarray(k, np.int32)
        out = np.empty(len(lam), np.float64)
        self._check(self._L.bdk_poisson_logsf(self._h, lam.ctypes.data_as(C.POINTER(C.c_double)),
                                              k.ctypes.data_as(C.POINTER(C.c_int32)),
                                              out.ctypes.data_as(C.POINTER(C.c_double)), len(lam)), "bdk_poisson_logsf")
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._L.bdk_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def _prefer_bundled_nccl():
    """libbdk dlopens NCCL by soname at the first communicator call. In a process that will also import torch the
    library must be the one torch was linked against (an older system libnccl.so.2 loaded first would be picked up by
    torch's own DT_NEEDED lookup and miss symbols), so point BDK_NCCL_LIB at the wheel's copy when there is one."""
    if os.environ.get("BDK_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            p = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(p):
                os.environ["BDK_NCCL_LIB"] = p
                return
    except Exception:
        pass


def comm_unique_id() -> bytes:
    """128-byte NCCL unique id (rank 0 calls this and distributes it)."""
    _prefer_bundled_nccl()
    L = load_library()
    buf = C.create_string_buffer(128)
    rc = L.bdk_comm_unique_id(C.cast(buf, C.c_void_p), 128)
    if rc != 0:
        raise BdkError(f"bdk_comm_unique_id failed ({rc}): {L.bdk_last_error(None).decode()}")
    return buf.raw


def _copy_array(ptr: Optional[int], n: int, dtype: np.dtype) -> np.ndarray:
    if not n or not ptr:
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


@dataclasses.dataclass
class SvTable:
    """SV calls in the reference's output order (copied out of the context)."""
    sv: np.ndarray            # SV_DTYPE
    lib_count: np.ndarray     # [n, nlib]
    cn_count: np.ndarray      # [n, nkey]
    copy_number: np.ndarray   # [n, nkey]
    nkey: int

    @classmethod
    def from_result(cls, r: Result, nlib: int) -> "SvTable":
        n = r.n_sv
        sv = _copy_array(C.cast(r.sv, C.c_void_p).value, n, SV_DTYPE)
        lc = _copy_array(C.cast(r.lib_count, C.c_void_p).value, n * nlib, np.dtype(np.int32)).reshape(n, nlib)
        cc = _copy_array(C.cast(r.cn_count, C.c_void_p).value, n * r.nkey, np.dtype(np.uint32)).reshape(n, r.nkey)
        cn = _copy_array(C.cast(r.copy_number, C.c_void_p).value, n * r.nkey, np.dtype(np.float32)).reshape(n, r.nkey)
        return cls(sv, lc, cc, cn, r.nkey)

    def as_result(self) -> Result:
        """Result struct over this table's arrays (keeps them alive through self)."""
        r = Result()
        r.n_sv = len(self.sv)
        self._keep = [np.ascontiguousarray(self.sv), np.ascontiguousarray(self.lib_count, np.int32),
                      np.ascontiguousarray(self.cn_count, np.uint32), np.ascontiguousarray(self.copy_number, np.float32)]
        r.sv = C.cast(self._keep[0].ctypes.data, C.POINTER(Sv))
        r.lib_count = C.cast(self._keep[1].ctypes.data, C.POINTER(C.c_int32))
        r.cn_count = C.cast(self._keep[2].ctypes.data, C.POINTER(C.c_uint32))
        r.copy_number = C.cast(self._keep[3].ctypes.data, C.POINTER(C.c_float))
        r.nkey = self.nkey
        return r


def _name_array(names: Sequence[str]):
    return (C.c_char_p * max(1, len(names)))(*[s.encode() for s in names])


def format_output(bundle: ParamBundle, summary: SummaryT, table: SvTable, lib_names: Sequence[str],
                  bam_names: Sequence[str], tid_names: Sequence[str]) -> str:
    """The reference's stdout from '#Library Statistics:' on (C++ formatter in libbdk.so;
    BreakDancerMax.cpp:82-153 + BreakDancer.cpp:465-497)."""
    L = load_library()
    ln, bn, tn = _name_array(lib_names), _name_array(bam_names), _name_array(tid_names)
    names = (C.c_void_p * 2)(C.cast(ln, C.c_void_p), C.cast(bn, C.c_void_p))
    cap = 1 << 16
    while True:
        buf = C.create_string_buffer(cap)
        n = L.bdh_format_header(C.byref(bundle.params), C.byref(summary), C.cast(names, C.c_void_p),
                                int(bundle.opts.print_AF), buf, cap)
        if n < cap:
            head = buf.value.decode()
            break
        cap = n + 1
    res = table.as_result()
    cap = max(1 << 16, 256 * (len(table.sv) + 1) * (1 + len(lib_names) // 4))
    sticky = C.c_int(0)
    while True:
        buf = C.create_string_buffer(cap)
        sticky.value = 0
        n = L.bdh_format_rows(C.byref(bundle.params), C.byref(res), C.cast(names, C.c_void_p), C.cast(tn, C.c_void_p),
                              int(bundle.opts.print_AF), C.byref(sticky), buf, cap)
        if n < cap:
            return head + buf.value.decode()
        cap = n + 1


class Bam2cfgOpts(C.Structure):
    _fields_ = [("min_mapq", C.c_int32), ("n_obs", C.c_int32), ("cut_sd", C.c_double), ("min_mean", C.c_double), ("max_cv", C.c_double),
                ("use_mapq", C.c_int32), ("solid", C.c_int32), ("flag_hist", C.c_int32), ("rg_lib_file", C.c_char_p)]


def bam2cfg(bams: Sequence[str], **kw) -> str:
    """perl/bam2cfg.pl in C++ (bdh_bam2cfg): the configuration text for position-sorted BAM files.
    Keyword options: min_mapq, n_obs, cut_sd, min_mean, max_cv, use_mapq, solid, flag_hist, rg_lib_file."""
    L = load_library()
    o = Bam2cfgOpts()
    L.bdh_bam2cfg_defaults(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v.encode() if isinstance(v, str) else v)
    arr = (C.c_char_p * len(bams))(*[b.encode() for b in bams])
    err = C.create_string_buffer(512)
    n = L.bdh_bam2cfg(arr, len(bams), C.byref(o), None, 0, err, 512)
    if n < 0:
        raise BdkError(err.value.decode())
    buf = C.create_string_buffer(n + 1)
    L.bdh_bam2cfg(arr, len(bams), C.byref(o), buf, n + 1, err, 512)
    return buf.value.decode()


def write_bam(path: str, tid_names: Sequence[str], tid_lens: Sequence[int], rg_names: Sequence[str],
              cols: Dict[str, np.ndarray], name_prefix: str = "r", write_am: bool = True, level: int = 1,
              threads: int = 0):
    L = load_library()
    soa = make_soa(cols)
    err = C.create_string_buffer(512)
    tl = (C.c_uint32 * len(tid_lens))(*tid_lens)
    rc = L.bdh_write_bam(path.encode(), len(tid_names), _name_array(tid_names), tl, len(rg_names),
                         _name_array(rg_names), C.byref(soa), len(cols["pos"]), name_prefix.encode(),
                         int(write_am), level, threads, err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
