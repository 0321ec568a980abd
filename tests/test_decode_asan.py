"""The host BAM decoder under AddressSanitizer / UBSan (tests/hostsim/decode_asan.cpp, built from the host sources): every
reading mode -- whole file, -o through the index and without it, windows, forced chain segments, both merges, records kept
for the FASTQ dump -- on good, decoy-ridden, truncated and bit-flipped files. Modes must agree on the checksum of the columns
and nothing may touch memory it does not own."""
import os
import subprocess

import numpy as np
import pytest

from breakdancer_b200 import api, synth
from tests import util
from tests.test_config_and_bam import _decoy_records, _handmade_bam

CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
SAN = ["-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer"]


@pytest.fixture(scope="module")
def exe():
    core = os.path.join(util.ROOT, "build", "bdk_core.o")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(core) or not os.path.isdir(os.path.join(cuda, "lib64")):
        pytest.skip("build/bdk_core.o or the CUDA runtime library is not here (run make first)")
    out = os.path.join(util.ROOT, "tests", "_build", "asan")
    os.makedirs(out, exist_ok=True)
    objs, procs = [], []
    for name in ("config", "bam_io", "format", "options", "support"):
        o = os.path.join(out, name + ".o")
        objs.append(o)
        procs.append(subprocess.Popen([CXX] + SAN + ["-I" + os.path.join(cuda, "include"), "-c",
                                                      os.path.join(util.ROOT, "breakdancer_b200", "csrc", "host", name + ".cpp"), "-o", o]))
    assert all(p.wait() == 0 for p in procs)
    path = os.path.join(util.ROOT, "tests", "_build", "decode_asan")
    subprocess.check_call([CXX] + SAN + [os.path.join(util.ROOT, "tests", "hostsim", "decode_asan.cpp")] + objs + [core, "-o", path,
                          "-L" + os.path.join(cuda, "lib64"), "-lcudart_static", "-lz", "-lpthread", "-ldl", "-lrt"])
    return path


def _run(exe, cwd, cfg, region="", threads=4, keep=0, **env):
    e = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", **{k: str(v) for k, v in env.items()})
    for k in ("BDK_NO_BAI", "BDK_DECODE_WINDOW_KB", "BDK_CHAIN_SEGMENTS", "BDK_MERGE_HEAP", "BDK_FAST_INFLATE"):
        if k not in env:
            e.pop(k, None)
    p = subprocess.run([exe, cfg, region, str(threads), str(keep)], cwd=cwd, env=e, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "ERROR: AddressSanitizer" not in p.stderr and "runtime error" not in p.stderr, p.stdout + p.stderr[-3000:]
    return p.stdout.strip()


def test_every_reading_mode_under_sanitizers(exe, tmp_path):
    # bundled real bams (two files, indexed)
    ref = _run(exe, util.CHR21, "inv_del_bam_config")
    assert ref.startswith("n=5917 ")
    for env in ({"BDK_MERGE_HEAP": 1}, {"BDK_DECODE_WINDOW_KB": 1}, {"BDK_CHAIN_SEGMENTS": 9}, {"BDK_FAST_INFLATE": 0}, {"BDK_DECODE_WINDOW_KB": 1, "BDK_CHAIN_SEGMENTS": 3}):
        assert _run(exe, util.CHR21, "inv_del_bam_config", **env) == ref, env
    reg = _run(exe, util.CHR21, "inv_del_bam_config", region="21:29185000-29186000")
    assert _run(exe, util.CHR21, "inv_del_bam_config", region="21:29185000-29186000", BDK_NO_BAI=1) == reg
    assert _run(exe, util.CHR21, "inv_del_bam_config", region="21", keep=1) == _run(exe, util.CHR21, "inv_del_bam_config", region="21", keep=1, BDK_NO_BAI=1)
    assert "Failed to parse bam region" in _run(exe, util.CHR21, "inv_del_bam_config", region="nope")
    # synthetic two-bam data set, three sequences, indexed where the reference's samtools is built
    w = synth.generate(util.GENOME3, util.LIBS4, 30000, seed=77, anomaly_frac=0.05)
    for bam, cols in synth.split_by_bam(w).items():
        api.write_bam(str(tmp_path / bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=6)
        samtools = os.path.join(util.ROOT, "oracle", "_ref", "samtools")
        if os.path.exists(samtools):
            subprocess.check_call([samtools, "index", str(tmp_path / bam)])
    (tmp_path / "cfg").write_text(w.config_text())
    ref = _run(exe, tmp_path, "cfg")
    assert ref.startswith("n=%d " % w.n)
    for env in ({"BDK_MERGE_HEAP": 1}, {"BDK_DECODE_WINDOW_KB": 1}, {"BDK_DECODE_WINDOW_KB": 700, "BDK_CHAIN_SEGMENTS": 4}, {"BDK_CHAIN_SEGMENTS": 1000}):
        assert _run(exe, tmp_path, "cfg", threads=8, **env) == ref, env
    assert _run(exe, tmp_path, "cfg", threads=1) == ref
    for region in (w.genome[0][0], w.genome[2][0] + ":5000-90000"):
        a = _run(exe, tmp_path, "cfg", region=region, keep=1)
        assert a == _run(exe, tmp_path, "cfg", region=region, keep=1, BDK_NO_BAI=1) and not a.startswith("n=0 ")
        assert _run(exe, tmp_path, "cfg", region=region) == _run(exe, tmp_path, "cfg", region=region, BDK_NO_BAI=1, BDK_DECODE_WINDOW_KB=1)
    # decoys, truncation, damaged bytes
    recs, _ = _decoy_records(n=1500)
    (tmp_path / "decoy.bam").write_bytes(_handmade_bam(recs))
    (tmp_path / "dcfg").write_text("map:decoy.bam\tlib:L\tmean:300\tstd:30\treadlen:36\n")
    ref = _run(exe, tmp_path, "dcfg", BDK_CHAIN_SEGMENTS=1)
    assert ref.startswith("n=1500 ")
    for env in ({"BDK_CHAIN_SEGMENTS": 50}, {"BDK_CHAIN_SEGMENTS": 5000}, {"BDK_DECODE_WINDOW_KB": 1, "BDK_CHAIN_SEGMENTS": 40}):
        assert _run(exe, tmp_path, "dcfg", **env) == ref, env
    whole = _handmade_bam(recs)
    rng = np.random.default_rng(1)
    for k in range(12):
        bad = bytearray(whole)
        if k % 3 == 0:
            bad = bad[:int(rng.integers(30, len(bad)))]                                   # cut anywhere
        else:
            for _ in range(int(rng.integers(1, 4))):
                bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))      # flipped bits, headers and payload alike
        (tmp_path / "bad.bam").write_bytes(bytes(bad))
        (tmp_path / "bcfg").write_text("map:bad.bam\tlib:L\tmean:300\tstd:30\treadlen:36\n")
        for env in ({}, {"BDK_DECODE_WINDOW_KB": 1}):
            out = _run(exe, tmp_path, "bcfg", **env)
            assert out.startswith("error: ") or out == ref, out       # refused, or the flip hit bytes that do not matter (e.g. gzip MTIME)
