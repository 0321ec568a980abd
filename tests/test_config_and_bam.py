"""Host side: bam2cfg config grammar, BAM decode / merge / write (CPU only)."""
import os

import numpy as np
import pytest

from breakdancer_b200 import api, synth
from tests import util

# reference test/lib/io/TestBamConfigEntry.cpp:36-100 (field ordinals of BamConfigEntry::Field)
BAM_FILE, LIBRARY_NAME, READ_GROUP, MEAN, STDDEV, READ_LENGTH, UPPER, LOWER, MIN_MAP_QUAL, SAMPLE, UNKNOWN = range(11)
ALIASES = {
    "map": BAM_FILE, "lib": LIBRARY_NAME, "libname": LIBRARY_NAME, "library_name": LIBRARY_NAME,
    "groUp": READ_GROUP, "ReadgroUp": READ_GROUP, "Read_groUp": READ_GROUP,
    "mean": MEAN, "mean_insert": MEAN, "mean_insert_size": MEAN,
    "std": STDDEV, "stddev": STDDEV, "insert_stddev": STDDEV, "insert_size_stddev": STDDEV, "stddev_insert": STDDEV,
    "stddev_insert_size": STDDEV, "readlen": READ_LENGTH, "rEaDlEnGtH": READ_LENGTH, "average_readlen": READ_LENGTH,
    "average_readlength": READ_LENGTH, "upp": UPPER, "upper": UPPER, "uppEr_cutOff": UPPER, "inseRt_size_uPper_cutoff": UPPER,
    "low": LOWER, "lower": LOWER, "lower_cuToff": LOWER, "insert_size_lower_cutoff": LOWER,
    "mapqual": MIN_MAP_QUAL, "mapPing_quAlity": MIN_MAP_QUAL, "samp": SAMPLE, "sample": SAMPLE, "samplename": SAMPLE,
    "sample_name": SAMPLE,
}


def test_translate_token_alias_table():
    L = api.load_library()
    assert L.bdh_config_translate_token(b"ZIOJFksfjlaiaowinfd") == UNKNOWN
    for key, field in ALIASES.items():
        for k in (key, key.upper(), key.lower()):
            assert L.bdh_config_translate_token(k.encode()) == field, k


def test_chr21_config():
    cfg = api.BamConfig(path=os.path.join(util.CHR21, "inv_del_bam_config"))
    assert cfg.lib_names == ["H_IJ-NA19238-NA19238-extlibs", "H_IJ-NA19240-NA19240-extlibs"]   # sorted by name
    assert cfg.bam_files == ["NA19238_chr21_del_inv.bam", "NA19240_chr21_del_inv.bam"]
    assert cfg.window == 287
    l0, l1 = cfg.libs
    assert abs(l0.uppercutoff - 532.53) < 1e-3 and abs(l0.lowercutoff - 311.36) < 1e-3 and l0.bam_index == 0
    assert abs(l1.mean_insertsize - 467.59) < 1e-3 and l1.min_mapping_quality == -1 and l1.bam_index == 1
    assert cfg.rg_lib("2880590781-110718_I806_FCD0E3VABXX_L4_HUMxqmRADDIABPEI-142") == 0
    assert cfg.rg_lib("no-such-read-group") == 0   # falls back to the first bam's library


def test_config_rules():
    text = ("map:b.bam\tlib:L2\tmean:300\tstd:10\treadlen:50\tmapqual:10\n"
            "map:a.bam\tsample:S1\tmean:400\tstd:20\treadlen:100\tupper:500\tlower:100\n"
            "\n"
            "map:ignored.bam\tlib:ZZ\n")
    cfg = api.BamConfig(text=text, cut_sd=4)
    assert cfg.bam_files == ["a.bam", "b.bam"] and cfg.lib_names == ["L2", "S1"]   # parse stops at the empty line
    L2, S1 = cfg.libs
    assert L2.uppercutoff == 340.0 and L2.lowercutoff == 260.0 and L2.min_mapping_quality == 10 and L2.bam_index == 1
    assert S1.uppercutoff == 500.0 and S1.lowercutoff == 100.0 and S1.bam_index == 0
    assert cfg.window == 200                       # min over lines of int(mean - 2 * readlen)
    with pytest.raises(RuntimeError, match="Required field 'map'"):
        api.BamConfig(text="lib:x\tmean:1\n")


def test_chr21_decode_counts_and_order(tmp_path):
    cwd = os.getcwd()
    os.chdir(util.CHR21)
    try:
        cfg = api.BamConfig(path="inv_del_bam_config")
        one = api.BamStream(cfg, paths=["NA19238_chr21_del_inv.bam"])
        two = api.BamStream(cfg, paths=["NA19240_chr21_del_inv.bam"])
        assert (one.n, two.n) == (3069, 2848)       # reference test-data/TestData.hpp.in:20-23
        both = api.BamStream(cfg, keep_records=True)
        assert both.n == 3069 + 2848
        c = both.cols
        key = c["tid"].astype(np.int64) << 33 | c["pos"].astype(np.int64) << 1 | ((c["flag"] & 16) != 0)
        assert np.all(np.diff(key) >= 0)             # merged by (tid, pos, strand)
        assert len(both.rg_lib) == 14 and set(both.rg_lib) == {0, 1}
        assert both.qname(0).startswith("FC")
        region = api.BamStream(cfg, region="21")
        assert region.n == both.n and np.array_equal(region.cols["pos"], c["pos"])
        sub = api.BamStream(cfg, region="21:29185000-29186000")
        assert 0 < sub.n < both.n and sub.cols["pos"].max() < 29186000
    finally:
        os.chdir(cwd)


def test_bam_write_read_round_trip(tmp_path):
    w = synth.generate(util.GENOME3, util.LIBS4, 20000, seed=3)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for bam, cols in synth.split_by_bam(w).items():
            api.write_bam(bam, [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols)
        cfg = api.BamConfig(text=w.config_text())
        st = api.BamStream(cfg, keep_records=True)
        assert st.n == w.n
        for name in ("pos", "mpos", "tid", "mtid", "isize", "flag", "mapq", "qlen"):
            a = np.sort(st.cols[name].astype(np.int64))
            b = np.sort(w.cols[name].astype(np.int64))
            assert np.array_equal(a, b), name
        # names are "r<pair id>": both mates hash to the same key, different pairs to different keys
        _, counts = np.unique(st.cols["qid"], return_counts=True)
        assert counts.max() == 2
        fq = st.fastq(0).split("\n")
        assert fq[0] == "@" + st.qname(0) and len(fq[1]) == st.cols["qlen"][0] and fq[2] == "+"
    finally:
        os.chdir(cwd)


# ---- record boundaries found in parallel (BamData::find_records) ----------------------------------------------------------
def _bam_record(tid, pos, name, flag, lseq, mtid, mpos, isize, aux=b"", mapq=30):
    import struct
    name_b = name.encode() + b"\0"
    cigar = struct.pack("<I", (lseq << 4) | 0)
    seq = bytes((lseq + 1) // 2)
    qual = bytes([30]) * lseq
    core = struct.pack("<iiBBHHHiiii", tid, pos, len(name_b), mapq, 4680, 1, flag, lseq, mtid, mpos, isize)
    body = core + name_b + cigar + seq + qual + aux
    return struct.pack("<I", len(body)) + body


def _bgzf(raw: bytes, block=20000) -> bytes:
    import struct
    import zlib
    out = b""
    for o in list(range(0, len(raw), block)) + [None]:
        chunk = b"" if o is None else raw[o:o + block]          # the empty last block is the BGZF end-of-file marker
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        data = co.compress(chunk) + co.flush()
        out += struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(data) + 25) + data
        out += struct.pack("<II", zlib.crc32(chunk), len(chunk))
    return out


def _handmade_bam(records) -> bytes:
    import struct
    text = b"@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:c1\tLN:100000000\n@RG\tID:g\tLB:L\n"
    head = b"BAM\1" + struct.pack("<I", len(text)) + text + struct.pack("<I", 1) + struct.pack("<I", 3) + b"c1\0" + struct.pack("<I", 100000000)
    return _bgzf(head + b"".join(records))


def _decoy_records(seed=5, n=3000):
    """Records whose aux bytes hold perfectly formed records: (records, their positions)."""
    rng = np.random.default_rng(seed)
    recs, want_pos = [], []
    pos = 100
    for i in range(n):
        pos += int(rng.integers(1, 50))
        aux = b"RGZg\0"
        kind = i % 4
        if kind == 1:      # a run of decoy records ending exactly where the real record ends (the decoy chain joins the real one)
            decoys = b"".join(_bam_record(0, 7_000_000 + k, "decoy%d" % k, 99, 20, 0, 7_000_100, 120) for k in range(int(rng.integers(3, 8))))
            aux += b"XDBC" + np.uint32(len(decoys)).tobytes() + decoys
        elif kind == 2:    # decoys followed by padding: the decoy chain runs off into garbage
            decoys = b"".join(_bam_record(0, 8_000_000 + k, "d%d" % k, 147, 10, 0, 8_000_100, -90) for k in range(4))
            pad = bytes(rng.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8))
            aux += b"XDBC" + np.uint32(len(decoys) + len(pad)).tobytes() + decoys + pad
        elif kind == 3:    # a decoy that claims a huge record reaching far past the real one
            import struct
            big = bytearray(_bam_record(0, 9_000_000, "big", 99, 30, 0, 9_000_100, 150))
            big[0:4] = struct.pack("<I", 50_000)
            rest = b"".join(_bam_record(0, 9_100_000 + k, "e%d" % k, 99, 30, 0, 9_000_100, 150) for k in range(3))
            aux += b"XDBC" + np.uint32(len(big) + len(rest)).tobytes() + bytes(big) + rest
        recs.append(_bam_record(0, pos, "r%d" % i, 99 if i % 2 == 0 else 147, 36, 0, pos + 200, 236, aux))
        want_pos.append(pos)
    return recs, want_pos


def test_record_chain_segments_with_decoys(tmp_path, monkeypatch):
    """Records whose aux bytes hold perfectly formed records (decoys): the parallel search for record boundaries must still
    return the serial chain, however the buffer is cut (csrc/host/bam_io.cpp find_records)."""
    recs, want_pos = _decoy_records()
    (tmp_path / "h.bam").write_bytes(_handmade_bam(recs))
    cfg = api.BamConfig(text="map:%s\tlib:L\tmean:300\tstd:30\treadlen:36\n" % (tmp_path / "h.bam"))
    results = []
    for nseg in (1, 2, 7, 64, 1000, 20000):
        monkeypatch.setenv("BDK_CHAIN_SEGMENTS", str(nseg))
        st = api.BamStream(cfg, threads=4)
        results.append({k: v.copy() for k, v in st.cols.items()})
        st.close()
    assert results[0]["pos"].tolist() == want_pos
    for r in results[1:]:
        for k in results[0]:
            assert np.array_equal(results[0][k], r[k]), k
    # a file cut in the middle of a record is refused the same way for every cut
    raw_recs = b"".join(recs)
    import struct
    text = b"@SQ\tSN:c1\tLN:100000000\n"
    head = b"BAM\1" + struct.pack("<I", len(text)) + text + struct.pack("<I", 1) + struct.pack("<I", 3) + b"c1\0" + struct.pack("<I", 100000000)
    (tmp_path / "t.bam").write_bytes(_bgzf(head + raw_recs[:len(raw_recs) - 17]))
    cfg_t = api.BamConfig(text="map:%s\tlib:L\tmean:300\tstd:30\treadlen:36\n" % (tmp_path / "t.bam"))
    for nseg in (1, 7, 1000):
        monkeypatch.setenv("BDK_CHAIN_SEGMENTS", str(nseg))
        with pytest.raises(RuntimeError, match="truncated BAM record"):
            api.BamStream(cfg_t, threads=4)


def test_two_bam_merge_in_parallel_equals_the_heap_merge(tmp_path, monkeypatch):
    """Two bams full of equal (tid, pos, strand) keys: the partitioned two-way merge (closed form of the reference's
    priority-queue ties, BamMerger.cpp:40-126) against the priority queue itself; also three bams (heap path with packed keys)
    against the same records split differently, and an empty second bam."""
    rng = np.random.default_rng(9)

    def bam(path, n, tag, lo=100, hi=40000, sort=True):
        pos = rng.integers(lo, hi, n)                                       # 1-2 records per position and bam: ties everywhere, gaps too
        if sort:
            pos = np.sort(pos)
        recs = [_bam_record(0, int(p), "%s%d" % (tag, i), 99 if rng.random() < 0.5 else 147, 36, 0, int(p) + 200, 236, b"RGZg\0")
                for i, p in enumerate(pos)]
        # records with equal positions must come in the order the comparator gives inside one (sorted) bam: any order is legal
        path.write_bytes(_handmade_bam(recs))

    bam(tmp_path / "a.bam", 60000, "a")
    bam(tmp_path / "b.bam", 45000, "b")
    bam(tmp_path / "c.bam", 20000, "c")
    bam(tmp_path / "d.bam", 40000, "d", hi=400)                             # every position of d also occurs in a: no place to cut
    bam(tmp_path / "u.bam", 30000, "u", sort=False)                         # not sorted: the priority queue decides, as in the reference
    (tmp_path / "e.bam").write_bytes(_handmade_bam([]))
    line = "map:%s\tlib:L%d\tmean:300\tstd:30\treadlen:36\n"
    for names in (("a", "b"), ("b", "a"), ("a", "d"), ("d", "a"), ("a", "u"), ("a", "e"), ("e", "a"), ("a", "b", "c")):
        cfg = api.BamConfig(text="".join(line % (tmp_path / (nm + ".bam"), i) for i, nm in enumerate(names)))
        paths = [str(tmp_path / (nm + ".bam")) for nm in names]
        got = {}
        for mode in ("parallel", "heap"):
            if mode == "heap":
                monkeypatch.setenv("BDK_MERGE_HEAP", "1")
            else:
                monkeypatch.delenv("BDK_MERGE_HEAP", raising=False)
            st = api.BamStream(cfg, paths=paths, threads=8, keep_records=True)
            got[mode] = {k: v.copy() for k, v in st.cols.items()}
            got[mode]["names"] = [st.qname(i) for i in range(0, st.n, 97)]
            st.close()
        for k in got["heap"]:
            assert np.array_equal(got["heap"][k], got["parallel"][k]) if k != "names" else got["heap"][k] == got["parallel"][k], (names, k)
        p = got["heap"]["pos"]
        assert "u" in names or np.all(p[1:] >= p[:-1])


def test_region_through_the_bam_index_equals_the_full_scan(tmp_path, monkeypatch, capfd):
    """-o with a .bai next to the bam: only the members holding that reference sequence are inflated (RegionLimitedBamReader.hpp:40-66
    reads through samtools' index); same records as the scan of the whole file. The index is written by the reference's own
    vendored samtools (oracle/_ref, test infrastructure)."""
    samtools = os.path.join(util.ROOT, "oracle", "_ref", "samtools")
    if not os.path.exists(samtools):
        pytest.skip("oracle/_ref/samtools not built")
    import subprocess
    genome = [("chrA", 400000), ("chrB", 300000), ("chrEmpty", 50000), ("chrC", 350000)]
    w = synth.generate([g for g in genome if g[0] != "chrEmpty"], util.LIBS4, 60000, seed=17, anomaly_frac=0.05)
    # the header names a sequence without any record between chrB and chrC
    names = [g[0] for g in genome]
    remap = np.array([0, 1, 3], dtype=np.int32)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for bam, cols in synth.split_by_bam(w).items():
            cols = dict(cols)
            cols["tid"] = remap[cols["tid"]]
            cols["mtid"] = np.where(cols["mtid"] >= 0, remap[np.maximum(cols["mtid"], 0)], cols["mtid"]).astype(np.int32)
            api.write_bam(bam, names, [g[1] for g in genome], w.rg_names, cols, level=6)
            subprocess.check_call([samtools, "index", bam])
            assert os.path.exists(bam + ".bai")
        cfg = api.BamConfig(text=w.config_text())
        monkeypatch.setenv("BDK_DECODE_TRACE", "1")
        for region in ("chrA", "chrB", "chrC", "chrEmpty", "chrB:1000-90000", "chrC:200,000", "chrA:399000-400000"):
            monkeypatch.setenv("BDK_NO_BAI", "1")
            full = api.BamStream(cfg, region=region, threads=4, keep_records=True)
            capfd.readouterr()
            monkeypatch.delenv("BDK_NO_BAI")
            idx = api.BamStream(cfg, region=region, threads=4, keep_records=True)
            assert "through the index" in capfd.readouterr().err, region
            assert idx.n == full.n, region
            assert idx.n == 0 if region == "chrEmpty" else (idx.n > 0 or ":" in region)
            for k in full.cols:
                assert np.array_equal(full.cols[k], idx.cols[k]), (region, k)
            for i in range(0, idx.n, 501):
                assert idx.qname(i) == full.qname(i) and idx.fastq(i) == full.fastq(i)
            full.close(); idx.close()
        with pytest.raises(RuntimeError, match="Failed to parse bam region"):
            api.BamStream(cfg, region="chrZ", threads=2)
    finally:
        os.chdir(cwd)


def test_segment_form_of_the_record_chain_is_exact_or_refuses(tmp_path):
    """csrc/bam_records.h segment_guess / segment_consistent (the bodies of the kernels in csrc/bam_decode.cuh), on the host:
    consistent segments reproduce the serial chain; with decoy records the check refuses instead of returning another chain."""
    import subprocess
    src = os.path.join(util.ROOT, "tests", "hostsim", "bam_chain_host.cpp")
    exe = os.path.join(util.ROOT, "tests", "_build", "bam_chain_host")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", src, "-o", exe, "-lz"])
    w = synth.generate(util.GENOME3, util.LIBS4, 30000, seed=12, anomaly_frac=0.05)
    for bam, cols in synth.split_by_bam(w).items():
        api.write_bam(str(tmp_path / bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols)
    normal = [str(tmp_path / b) for b in synth.split_by_bam(w)] + [os.path.join(util.CHR21, "NA19238_chr21_del_inv.bam")]
    for path in normal:
        p = subprocess.run([exe, path], capture_output=True, text=True)
        assert p.returncode == 0, p.stdout
        lines = p.stdout.strip().split("\n")
        assert len(lines) == 4 and all("consistent=1 equal=1" in l for l in lines), p.stdout
    recs, _ = _decoy_records()
    (tmp_path / "decoy.bam").write_bytes(_handmade_bam(recs))
    p = subprocess.run([exe, str(tmp_path / "decoy.bam")], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout                                   # never consistent with another chain
    assert "consistent=0" in p.stdout, p.stdout                          # and the decoys do derail some segment size


def test_windowed_decode_equals_the_whole_file_decode(tmp_path, monkeypatch):
    """Files too large to hold inflated are decoded a window of BGZF members at a time (records cut by a window's end are carried
    over): same columns as the whole-file path, for one and two bams, with decoys, with -o without an index, and the same
    refusal of a truncated file."""
    w = synth.generate(util.GENOME3, util.LIBS4, 40000, seed=31, anomaly_frac=0.05)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for bam, cols in synth.split_by_bam(w).items():
            api.write_bam(bam, [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=1)
        recs, want_pos = _decoy_records(n=2000)
        open("decoy.bam", "wb").write(_handmade_bam(recs))
        raw = b"".join(recs)
        import struct
        text = b"@SQ\tSN:c1\tLN:100000000\n"
        head = b"BAM\1" + struct.pack("<I", len(text)) + text + struct.pack("<I", 1) + struct.pack("<I", 3) + b"c1\0" + struct.pack("<I", 100000000)
        open("cut.bam", "wb").write(_bgzf(head + raw[:len(raw) - 23]))
        cfg2 = api.BamConfig(text=w.config_text())
        one = sorted(synth.split_by_bam(w))[0]
        cases = [(cfg2, None, ""), (cfg2, [one], ""), (cfg2, None, w.genome[1][0]), (cfg2, None, w.genome[2][0] + ":1000-200000"),
                 (api.BamConfig(text="map:decoy.bam\tlib:L\tmean:300\tstd:30\treadlen:36\n"), None, "")]
        monkeypatch.setenv("BDK_NO_BAI", "1")
        for cfg, paths, region in cases:
            monkeypatch.delenv("BDK_DECODE_WINDOW_KB", raising=False)
            monkeypatch.delenv("BDK_CHAIN_SEGMENTS", raising=False)
            whole = api.BamStream(cfg, paths=paths, region=region, threads=4)
            want = {k: v.copy() for k, v in whole.cols.items()}
            want_lib = whole.rg_lib[whole.cols["rgid"]]
            names = whole.tid_names
            whole.close()
            assert len(want["pos"]) > 0
            for kb, nseg in ((1, None), (200, None), (1, 5), (1500, 3)):
                monkeypatch.setenv("BDK_DECODE_WINDOW_KB", str(kb))
                if nseg:
                    monkeypatch.setenv("BDK_CHAIN_SEGMENTS", str(nseg))
                else:
                    monkeypatch.delenv("BDK_CHAIN_SEGMENTS", raising=False)
                st = api.BamStream(cfg, paths=paths, region=region, threads=4)
                assert st.tid_names == names
                for k, v in want.items():
                    if k == "rgid":           # ids are handed out in order of first sight, which depends on the threads
                        continue
                    assert np.array_equal(v, st.cols[k]), (region, kb, nseg, k)
                assert np.array_equal(want_lib, st.rg_lib[st.cols["rgid"]])
                st.close()
        cfg_cut = api.BamConfig(text="map:cut.bam\tlib:L\tmean:300\tstd:30\treadlen:36\n")
        for kb in (1, 100000):
            monkeypatch.setenv("BDK_DECODE_WINDOW_KB", str(kb))
            with pytest.raises(RuntimeError, match="truncated BAM record"):
                api.BamStream(cfg_cut, threads=4)
    finally:
        os.chdir(cwd)


def test_repair_rounds_of_the_device_record_chain_reach_the_serial_chain(tmp_path):
    """The rule chain_resolve_kernel (csrc/bam_decode.cuh) uses on the GPU -- 8 KiB segments guess their first record, a segment
    out of place is re-entered from its predecessor's end once the predecessor is in place -- run on the host with the same
    bam_records.h functions, on generated records, on records that contain decoy records, with a cut-off last record and with
    spoiled guesses (singly, in runs, at the window's end): always the serial chain, in a few rounds."""
    import struct
    import subprocess
    import zlib
    exe = os.path.join(util.ROOT, "tests", "_build", "bam_chain_rounds")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", os.path.join(util.ROOT, "tests", "hostsim", "bam_chain_rounds.cpp"), "-o", exe])

    def record_bytes(path):
        data = open(path, "rb").read()
        out, off = bytearray(), 0
        while off + 18 <= len(data):
            xlen = struct.unpack_from("<H", data, off + 10)[0]
            bsize = struct.unpack_from("<H", data, off + 16)[0] + 1
            out += zlib.decompress(data[off + 12 + xlen:off + bsize - 8], -15)
            off += bsize
        o = 8 + struct.unpack_from("<I", out, 4)[0]
        nref = struct.unpack_from("<I", out, o)[0]
        o += 4
        for _ in range(nref):
            o += 4 + struct.unpack_from("<I", out, o)[0] + 4
        return bytes(out[o:]), nref

    w = synth.generate(util.GENOME3, util.LIBS4, 60000, seed=8, anomaly_frac=0.05)
    bam, cols = sorted(synth.split_by_bam(w).items())[0]
    api.write_bam(str(tmp_path / bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=1)
    recs, _ = _decoy_records(n=3000)
    (tmp_path / "decoy.bam").write_bytes(_handmade_bam(recs))
    for name in (bam, "decoy.bam"):
        raw, nref = record_bytes(str(tmp_path / name))
        (tmp_path / "raw.bin").write_bytes(raw)
        nseg = (len(raw) + 8191) // 8192
        for cut, spoiled in ((0, []), (777, []), (0, [1, 2, 3, 4, 5]), (50, [nseg // 2]), (3, [nseg - 1, nseg - 2]), (0, list(range(7, nseg, 11)))):
            p = subprocess.run([exe, str(tmp_path / "raw.bin"), str(nref), str(cut)] + [str(k) for k in spoiled], capture_output=True, text=True)
            assert p.returncode == 0 and " OK" in p.stdout, (name, cut, spoiled[:5], p.stdout)
            rounds = int(p.stdout.split("rounds=")[1].split()[0])
            assert rounds <= len(spoiled) + 40, p.stdout
