"""BGZF members inflated on the GPU (csrc/bgzf_inflate_warp.cuh: the warp-per-member decoder of bdk_push_bam, also behind the C ABI
bdk_bgzf_inflate and BDK_GPU_INFLATE=1 in the host BAM reader).

CPU: the member decoder every warp runs, compiled for the host, differential- and corruption-fuzzed against zlib under
AddressSanitizer / UBSan (tests/hostsim/gpu_inflate_warp_host.cpp). GPU: the kernel (with its CRC-32 check) against zlib on the
members of generated BAM files, on streams made of overlapping matches of every distance, and on damaged members; then the host
reader with the device stage switched on: same columns as the host path and not one member redone by the host."""
import ctypes as C
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from breakdancer_b200 import api, synth
from tests import util


def test_warp_member_decoder_and_sliced_crc_fuzz_on_the_host():
    """bgzf_inflate_warp.cuh (the decoder of the device-resident BAM decode): the 32 lanes of a phase run one after the other on
    the host; once with the checked byte-wise bit reader and once with the device's word-window reader (member at every alignment
    inside a buffer shaped like the device's, 512 bytes of output slack)."""
    src = os.path.join(util.ROOT, "tests", "hostsim", "gpu_inflate_warp_host.cpp")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    os.makedirs(os.path.join(util.ROOT, "tests", "_build"), exist_ok=True)
    for flag, name in (([], "gpu_inflate_warp_host"), (["-DBGZW_WINDOW_READER_ON_HOST"], "gpu_inflate_warp_host_w")):
        exe = os.path.join(util.ROOT, "tests", "_build", name)
        subprocess.check_call([cxx, "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer",
                               *flag, src, "-o", exe, "-lz"])
        p = subprocess.run([exe, "400"], capture_output=True, text=True)
        assert p.returncode == 0, p.stdout + p.stderr
        assert "refused_good=0 mismatched=0 crc_bad=0" in p.stdout, p.stdout


def _members(data: bytes):
    """(in_off, in_len, out_len) of every BGZF member of a file image."""
    out, off = [], 0
    while off + 18 <= len(data):
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        bsize = struct.unpack_from("<H", data, off + 16)[0] + 1
        isize = struct.unpack_from("<I", data, off + bsize - 4)[0]
        out.append((off + 12 + xlen, bsize - 12 - xlen - 8, isize))
        off += bsize
    return out


def _gpu_inflate(data: bytes, members):
    L = api.load_library()
    arr = np.zeros(len(members), dtype=np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("in_len", "<u4"), ("out_len", "<u4")]))
    o = 0
    for i, (in_off, in_len, out_len) in enumerate(members):
        arr[i] = (in_off, o, in_len, out_len)
        o += out_len
    out = np.full(o + 64, 0xEE, dtype=np.uint8)
    status = np.full(len(members), -1, dtype=np.int32)
    ms = C.c_float()
    buf = np.frombuffer(data, dtype=np.uint8)
    rc = L.bdk_bgzf_inflate(0, buf.ctypes.data, len(data), arr.ctypes.data, len(members), out.ctypes.data, o,
                            status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(ms))
    assert rc == 0, rc
    assert np.all(out[o:] == 0xEE)
    return out[:o], status, arr, ms.value


def _synthetic_bam(tmp_path, n_pairs, level):
    w = synth.generate(util.GENOME3, util.LIBS4, n_pairs, seed=21, anomaly_frac=0.05)
    d = tmp_path / ("l%d" % level)
    d.mkdir()
    for bam, cols in synth.split_by_bam(w).items():
        api.write_bam(str(d / bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=level)
    (d / "cfg").write_text(w.config_text())
    return w, d


@pytest.mark.gpu
def test_kernel_against_zlib_on_bam_members(tmp_path):
    for level in (1, 6, 9, 0):
        w, d = _synthetic_bam(tmp_path, 60000, level)
        for bam in sorted(synth.split_by_bam(w)):
            data = (d / bam).read_bytes()
            members = _members(data)
            got, status, arr, ms = _gpu_inflate(data, members)
            assert np.all(status == 0), (level, bam, status[status != 0][:10])
            want = b"".join(zlib.decompress(data[a:a + n], -15) for a, n, _ in members)
            assert got.tobytes() == want, (level, bam)
    # the bundled real BAMs (samtools-written members)
    for name in ("NA19238_chr21_del_inv.bam", "NA19240_chr21_del_inv.bam"):
        data = open(os.path.join(util.CHR21, name), "rb").read()
        members = _members(data)
        got, status, _, _ = _gpu_inflate(data, members)
        assert np.all(status == 0)
        assert got.tobytes() == b"".join(zlib.decompress(data[a:a + n], -15) for a, n, _ in members)


def _member_file(payloads, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    """A buffer shaped like a BGZF file (18 bytes of header, the raw DEFLATE stream, CRC-32 and ISIZE) for each payload, and
    its member list."""
    data, members = bytearray(), []
    for raw in payloads:
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        comp = co.compress(raw) + co.flush()
        data += b"\x1f\x8b\x08\x04" + bytes(14)
        members.append((len(data), len(comp), len(raw)))
        data += comp + struct.pack("<II", zlib.crc32(raw) & 0xffffffff, len(raw))
    return bytes(data) + bytes(64), members


@pytest.mark.gpu
def test_kernel_on_overlapping_matches_of_every_distance_and_other_shapes():
    """Matches whose source overlaps their destination (byte j = byte j mod distance of the bytes before: the warp copy does that
    modulo without an integer division), for every distance 1..300 and lengths up to 258; stored blocks, fixed codes, Huffman-only
    and incompressible members, empty and one-byte members."""
    rng = np.random.default_rng(11)
    payloads = []
    for d in range(1, 301):
        unit = rng.integers(0, 256, size=d, dtype=np.uint8).tobytes()
        payloads.append((unit * (900 // d + 3))[:900 + d] + rng.integers(0, 4, size=40, dtype=np.uint8).tobytes())
    payloads += [b"", b"A", bytes(65000), rng.integers(0, 256, size=60000, dtype=np.uint8).tobytes(),
                 (b"ACGT" * 9 + b"FFFFFFFFFFFFFFFFFFFFFFFF#######" * 3) * 300]
    for level, strategy in ((6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (0, zlib.Z_DEFAULT_STRATEGY),
                            (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
        data, members = _member_file(payloads, level, strategy)
        got, status, arr, _ = _gpu_inflate(data, members)
        assert np.all(status == 0), (level, strategy, np.nonzero(status)[0][:10], status[status != 0][:10])
        assert got.tobytes() == b"".join(payloads), (level, strategy)


@pytest.mark.gpu
def test_kernel_refuses_damaged_members_without_touching_their_neighbours(tmp_path):
    w, d = _synthetic_bam(tmp_path, 30000, 6)
    bam = sorted(synth.split_by_bam(w))[0]
    data = bytearray((d / bam).read_bytes())
    members = _members(bytes(data))
    good = [zlib.decompress(bytes(data[a:a + n]), -15) for a, n, _ in members]
    rng = np.random.default_rng(3)
    damaged = sorted(rng.choice(len(members) - 1, size=min(12, len(members) - 1), replace=False).tolist())
    for k, m in enumerate(damaged):
        a, n, _ = members[m]
        if k % 3 == 0:
            data[a + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))        # a flipped bit
        elif k % 3 == 1:
            members[m] = (a, max(1, n // 2), members[m][2])                          # truncated input
        else:
            members[m] = (a, n, members[m][2] + 1)                                   # wrong output length
    got, status, arr, _ = _gpu_inflate(bytes(data), members)
    for i, (a, n, olen) in enumerate(members):
        o = int(arr[i]["out_off"])
        if i not in damaged:
            assert status[i] == 0 and got[o:o + olen].tobytes() == good[i], i
        elif status[i] == 0:                        # a flipped bit may still give a well-formed stream of the right length: the CRC catches it
            assert olen == len(good[i])
    assert any(status[m] != 0 for m in damaged)


@pytest.mark.gpu
def test_reader_with_the_device_inflate_gives_the_same_columns(tmp_path, monkeypatch):
    w, d = _synthetic_bam(tmp_path, 80000, 6)
    cwd = os.getcwd()
    os.chdir(d)
    try:
        cfg = api.BamConfig(text=w.config_text())
        monkeypatch.delenv("BDK_GPU_INFLATE", raising=False)
        host = api.BamStream(cfg, threads=4)
        want = {k: v.copy() for k, v in host.cols.items()}
        host.close()
        redone_before = api.inflate_counters()[1]
        monkeypatch.setenv("BDK_GPU_INFLATE", "1")
        dev = api.BamStream(cfg, threads=4)
        assert dev.n == w.n
        for k, v in want.items():
            assert np.array_equal(v, dev.cols[k]), k
        dev.close()
        assert api.inflate_counters()[1] == redone_before          # the GPU decoded every member itself
    finally:
        os.chdir(cwd)
