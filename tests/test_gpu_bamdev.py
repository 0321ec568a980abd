"""Device-resident BAM decode (bdk_push_bam / bdk_decode_bam: csrc/bdk_bam.inl, bgzf_inflate_warp.cuh, bam_decode.cuh).

The GPU inflates the BGZF members (one warp per member, CRC32 checked on the device), finds the record boundaries, filters and
extracts the columns and classifies them, window by window. Checked here against the host decoder (bdh_stream_open, itself
pinned by the CPU tests): every column of every record, then the whole job (summary, anomalous reads, regions, SV table)
against a job fed from the host-decoded columns -- for every deflate level, with a member per window (every record boundary
case of the carry-over), on the bundled real BAMs whole / by region through the index / by region without it, on records that
contain decoy records, and on damaged files, which must be refused. The CLI tests of test_gpu_parity.py run through this path
too (one bam, no read dump)."""
import os
import subprocess

import numpy as np
import pytest

from breakdancer_b200 import api, synth
from tests import util
from tests.test_config_and_bam import _bam_record, _bgzf, _decoy_records, _handmade_bam

pytestmark = pytest.mark.gpu


def _write(tmp_path, n_pairs, level, seed=21, libs=None):
    w = synth.generate(util.GENOME3, libs or util.LIBS4, n_pairs, seed=seed, anomaly_frac=0.05)
    d = tmp_path / ("l%d_%d" % (level, seed))
    d.mkdir()
    for bam, cols in synth.split_by_bam(w).items():
        api.write_bam(str(d / bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=level)
    return w, d


def _compare_columns(cfg, path, region=""):
    """Host decoder vs device decoder on one file: all columns, read groups through their library."""
    host = api.BamStream(cfg, paths=[path], region=region, threads=4)
    dev = api.BamDevice(cfg, path=path, region=region)
    ctx = api.Context(dev.bundle(api.Options(chr=region)))
    cols, st = ctx.decode_bam(dev, host.n + 64)
    assert st["kept"] == host.n, (st, host.n)
    for k in api.COLUMN_DTYPES:
        if k == "rgid":
            assert np.array_equal(host.rg_lib[host.cols[k]], dev.rg_lib[cols[k]]), k
        else:
            assert np.array_equal(host.cols[k], cols[k]), (k, path, region)
    n = host.n
    ctx.close(); dev.close(); host.close()
    return st, n


def _job_from_host(cfg, path, opts):
    host = api.BamStream(cfg, paths=[path], region=opts.chr, threads=4)
    b = api.ParamBundle(opts, cfg.libs, cfg.nbam, host.rg_lib, host.rg_bam, cfg.window, max(1, len(host.tid_names)))
    ctx = api.Context(b)
    ctx.push({k: np.ascontiguousarray(v) for k, v in host.cols.items()})
    out = (ctx.summary(), ctx.finish(), ctx.regions(), ctx.areads())
    ctx.close(); host.close()
    return out


def _job_from_device(cfg, path, opts):
    dev = api.BamDevice(cfg, path=path, region=opts.chr)
    ctx = api.Context(dev.bundle(opts))
    st = ctx.push_bam(dev)
    out = (ctx.summary(), ctx.finish(), ctx.regions(), ctx.areads())
    times = ctx.kernel_times()
    ctx.close(); dev.close()
    return out, st, times


def _assert_same_job(a, b):
    (sa, ta, ra, (aa, rra)), (sb, tb, rb, (ab, rrb)) = a, b
    assert bytes(sa) == bytes(sb)
    assert ta.sv.tobytes() == tb.sv.tobytes() and np.array_equal(ta.lib_count, tb.lib_count) and np.array_equal(ta.cn_count, tb.cn_count)
    assert ta.copy_number.tobytes() == tb.copy_number.tobytes()
    assert np.array_equal(ra, rb) and np.array_equal(aa, ab) and np.array_equal(rra, rrb)


@pytest.mark.parametrize("level", [6, 1, 9, 0])
def test_device_decode_equals_host_decode_on_every_deflate_level(tmp_path, level):
    w, d = _write(tmp_path, 60000, level)
    cfg = api.BamConfig(text=w.config_text())
    for bam in sorted(synth.split_by_bam(w)):
        st, n = _compare_columns(cfg, str(d / bam))
        assert st["records"] >= n and st["sorted"] == 1 and st["windows"] >= 1


def test_a_member_per_window_carries_every_cut_record(tmp_path, monkeypatch):
    w, d = _write(tmp_path, 40000, 6, seed=5)
    cfg = api.BamConfig(text=w.config_text())
    bam = str(d / sorted(synth.split_by_bam(w))[0])
    monkeypatch.setenv("BDK_BAMDEV_WINDOW_KB", "1")          # every window is one member
    st, n = _compare_columns(cfg, bam)
    assert st["windows"] > 50
    monkeypatch.setenv("BDK_BAMDEV_WINDOW_KB", "100")
    st2, _ = _compare_columns(cfg, bam)
    assert 1 < st2["windows"] < st["windows"]
    for region in ("chrB", "chrB:100000-900000", "chrC"):     # the reader's overlap filter, no index
        _compare_columns(cfg, bam, region)


def test_whole_job_from_the_device_decode_equals_the_job_from_host_columns(tmp_path, monkeypatch):
    libs = [synth.LibSpec("lib_a", "one.bam", 315, 44, 75, ["a1", "a2"]), synth.LibSpec("lib_b", "one.bam", 467, 32, 100, ["b1"])]
    w, d = _write(tmp_path, 150000, 6, seed=9, libs=libs)
    cfg = api.BamConfig(text=w.config_text())
    bam = str(d / "one.bam")
    for opts in (api.Options(), api.Options(CN_lib=True, min_read_pair=1), api.Options(chr="chrB")):
        want = _job_from_host(cfg, bam, opts)
        got, st, times = _job_from_device(cfg, bam, opts)
        _assert_same_job(want, got)
        assert len(got[1].sv) > 0 and times["bam_inflate"]["launches"] >= 2 and times["bam_extract"]["ms"] > 0
    monkeypatch.setenv("BDK_BAMDEV_WINDOW_KB", "64")
    got, st, _ = _job_from_device(cfg, bam, api.Options())
    _assert_same_job(_job_from_host(cfg, bam, api.Options()), got)
    assert st["windows"] > 3


def test_bundled_real_bams_whole_by_index_and_by_scan(monkeypatch):
    cfg = api.BamConfig(path=os.path.join(util.CHR21, "inv_del_bam_config"))
    cwd = os.getcwd()
    os.chdir(util.CHR21)
    try:
        for name in cfg.bam_files:
            _compare_columns(cfg, name)
            _compare_columns(cfg, name, "21")
            _compare_columns(cfg, name, "21:15000000-30000000")
            monkeypatch.setenv("BDK_NO_BAI", "1")
            _compare_columns(cfg, name, "21:15000000-30000000")
            monkeypatch.delenv("BDK_NO_BAI")
    finally:
        os.chdir(cwd)


def test_records_that_contain_decoy_records(tmp_path, monkeypatch):
    recs, want_pos = _decoy_records(n=3000)
    bam = str(tmp_path / "decoy.bam")
    open(bam, "wb").write(_handmade_bam(recs))
    cfg = api.BamConfig(text="map:%s\tlib:L\tmean:300\tstd:30\treadlen:36\n" % bam)
    for kb in ("", "1", "40"):
        if kb:
            monkeypatch.setenv("BDK_BAMDEV_WINDOW_KB", kb)
        st, n = _compare_columns(cfg, bam)
        assert n == len(want_pos)
    assert st["guess_misses"] >= 0


def test_damaged_files_are_refused(tmp_path):
    w, d = _write(tmp_path, 30000, 6, seed=3)
    cfg = api.BamConfig(text=w.config_text())
    name = sorted(synth.split_by_bam(w))[0]
    data = bytearray((d / name).read_bytes())
    # a flipped bit in the middle of some member's DEFLATE stream: refused by the decoder or by the CRC
    bad = bytearray(data)
    bad[len(bad) // 2] ^= 0x10
    p = tmp_path / "flipped.bam"
    p.write_bytes(bytes(bad))
    dev = api.BamDevice(cfg, path=str(p))
    ctx = api.Context(dev.bundle(api.Options()))
    with pytest.raises(api.BdkError) as e:
        ctx.push_bam(dev)
    assert "did not inflate" in str(e.value) or "truncated" in str(e.value)
    ctx.close(); dev.close()
    # the data ends inside a record: the last members are dropped at a member boundary (the EOF marker stays)
    from tests.test_zz_gpu_inflate import _members
    mem = _members(bytes(data))
    cut = mem[len(mem) // 2][0] - 18                    # start of a member in the middle of the file
    eof = bytes(data[-28:])
    p2 = tmp_path / "cut.bam"
    p2.write_bytes(bytes(data[:cut]) + eof)
    dev = api.BamDevice(cfg, path=str(p2))
    ctx = api.Context(dev.bundle(api.Options()))
    with pytest.raises(api.BdkError) as e:
        ctx.push_bam(dev)
    assert "truncated" in str(e.value)
    ctx.close(); dev.close()


def test_cli_takes_the_device_path_and_falls_back_for_what_it_cannot_do(tmp_path):
    libs = [synth.LibSpec("lib_a", "one.bam", 315, 44, 75, ["a1"])]
    w, d = _write(tmp_path, 50000, 6, seed=4, libs=libs)
    (d / "cfg").write_text(w.config_text())
    outs = {}
    for mode in ("1", "0"):
        env = dict(os.environ, BDK_GPU_DECODE=mode)
        stats = d / ("stats%s.json" % mode)
        p = subprocess.run([util.CLI, "--stats-json", str(stats), "cfg"], cwd=d, env=env, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        outs[mode] = util.strip_header(p.stdout)
        import json
        js = json.loads(stats.read_text())
        assert js["device_decode"] == int(mode)
    assert outs["1"] == outs["0"] and outs["1"].count("\n") > 5


def _two_bam_columns(cfg, paths, region=""):
    host = api.BamStream(cfg, paths=paths, region=region, threads=4)
    d0 = api.BamDevice(cfg, path=paths[0], region=region)
    d1 = api.BamDevice(cfg, path=paths[1], region=region, after=d0)
    rg_lib = np.concatenate([d0.rg_lib, d1.rg_lib])
    rg_bam = np.concatenate([d0.rg_bam, d1.rg_bam])
    b = api.ParamBundle(api.Options(chr=region), cfg.libs, cfg.nbam, rg_lib, rg_bam, cfg.window, max(1, len(d0.tid_names)))
    ctx = api.Context(b)
    cols, (s0, s1) = ctx.decode_bams(d0, d1, host.n + 64)
    assert len(cols["pos"]) == host.n, (len(cols["pos"]), host.n)
    for k in api.COLUMN_DTYPES:
        if k == "rgid":
            assert np.array_equal(host.rg_lib[host.cols[k]], rg_lib[cols[k]]) and np.array_equal(host.rg_bam[host.cols[k]], rg_bam[cols[k]]), k
        else:
            assert np.array_equal(host.cols[k], cols[k]), (k, region)
    n = host.n
    ctx.close(); d0.close(); d1.close(); host.close()
    return s0, s1, n


def test_two_bams_merged_on_the_device_in_the_reference_order(tmp_path, monkeypatch):
    """Tumor / normal: both bams decoded on the GPU, merged there with BamMerger's tie rule -- every column of the merged stream
    against the host decoder's merge (itself tested against the priority queue and the reference binary), also with a region,
    with tiny decode windows, and on tie-heavy input (the same reads in both bams)."""
    w, d = _write(tmp_path, 120000, 6, seed=31)
    cfg = api.BamConfig(text=w.config_text())
    cwd = os.getcwd()
    os.chdir(d)
    try:
        assert cfg.bam_files == ["normal.bam", "tumor.bam"]
        s0, s1, n = _two_bam_columns(cfg, cfg.bam_files)
        assert s0["merge_parts"] > 10 and s0["merge_longest_part"] < n
        _two_bam_columns(cfg, cfg.bam_files, "chrB")
        monkeypatch.setenv("BDK_BAMDEV_WINDOW_KB", "64")
        _two_bam_columns(cfg, cfg.bam_files)
        monkeypatch.delenv("BDK_BAMDEV_WINDOW_KB")
        # ties everywhere: the second bam holds the first bam's records again (other read names do not matter for the order)
        import shutil
        shutil.copy("normal.bam", "tumor.bam")
        s0, s1, n = _two_bam_columns(cfg, cfg.bam_files)
        assert s0["merge_parts"] == 1            # no position occurs in one bam only: one part, merged by one thread
    finally:
        os.chdir(cwd)


def test_two_bam_job_and_cli_on_the_device_equal_the_host_decoder(tmp_path):
    w, d = _write(tmp_path, 150000, 6, seed=32)
    (d / "cfg").write_text(w.config_text())
    cfg = api.BamConfig(text=w.config_text())
    cwd = os.getcwd()
    os.chdir(d)
    try:
        for opts in (api.Options(), api.Options(CN_lib=True)):
            host = api.BamStream(cfg, threads=4)
            b = api.ParamBundle.from_stream(opts, cfg, host)
            ctx = api.Context(b)
            ctx.push({k: np.ascontiguousarray(v) for k, v in host.cols.items()})
            want = (ctx.summary(), ctx.finish(), ctx.regions(), ctx.areads())
            ctx.close(); host.close()
            d0 = api.BamDevice(cfg, path=cfg.bam_files[0])
            d1 = api.BamDevice(cfg, path=cfg.bam_files[1], after=d0)
            b = api.ParamBundle(opts, cfg.libs, cfg.nbam, np.concatenate([d0.rg_lib, d1.rg_lib]), np.concatenate([d0.rg_bam, d1.rg_bam]), cfg.window, len(d0.tid_names))
            ctx = api.Context(b)
            ctx.push_bams(d0, d1)
            got = (ctx.summary(), ctx.finish(), ctx.regions(), ctx.areads())
            ctx.close(); d0.close(); d1.close()
            _assert_same_job(want, got)
            assert len(got[1].sv) > 0
    finally:
        os.chdir(cwd)
    outs = {}
    for mode in ("1", "0"):
        stats = d / ("stats%s.json" % mode)
        p = subprocess.run([util.CLI, "--stats-json", str(stats), "cfg"], cwd=d, env=dict(os.environ, BDK_GPU_DECODE=mode), capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        outs[mode] = util.strip_header(p.stdout)
        import json
        assert json.loads(stats.read_text())["device_decode"] == int(mode)
    assert outs["1"] == outs["0"] and outs["1"].count("\n") > 5


LIBS_4BAMS = [synth.LibSpec("lane1", "lane1.bam", 315, 44, 75, ["l1a", "l1b"]),
              synth.LibSpec("lane2", "lane2.bam", 312, 43, 75, ["l2"]),
              synth.LibSpec("lane3", "lane3.bam", 467, 32, 75, ["l3"], tumor=True),
              synth.LibSpec("lane4", "lane4.bam", 476, 29, 100, ["l4"], tumor=True)]


def _n_bam_columns(cfg, paths, region=""):
    host = api.BamStream(cfg, paths=paths, region=region, threads=4)
    devs = []
    for p in paths:
        devs.append(api.BamDevice(cfg, path=p, region=region, after=devs[-1] if devs else None))
    rg_lib = np.concatenate([d.rg_lib for d in devs])
    rg_bam = np.concatenate([d.rg_bam for d in devs])
    b = api.ParamBundle(api.Options(chr=region), cfg.libs, cfg.nbam, rg_lib, rg_bam, cfg.window, max(1, len(devs[0].tid_names)))
    ctx = api.Context(b)
    cols, stats = ctx.decode_bams_n(devs, host.n + 64)
    assert len(cols["pos"]) == host.n, (len(cols["pos"]), host.n)
    for k in api.COLUMN_DTYPES:
        if k == "rgid":
            assert np.array_equal(host.rg_lib[host.cols[k]], rg_lib[cols[k]]) and np.array_equal(host.rg_bam[host.cols[k]], rg_bam[cols[k]]), k
        else:
            assert np.array_equal(host.cols[k], cols[k]), (k, region, len(paths))
    n = host.n
    ctx.close(); host.close()
    for d in devs:
        d.close()
    return n


@pytest.mark.gpu
def test_three_and_four_bams_on_the_device_in_the_reference_order(tmp_path):
    """One bam per lane: every bam decoded on the GPU, the merge order from the priority queue itself (host, on keys the device
    hands back), the columns gathered on the GPU -- every column of the merged stream against the host decoder's priority-queue
    merge (itself tested against the reference binary), for three and four bams, with a region, and on tie-heavy input (the same
    records in three bams: the order among equal keys is the heap's, history and all). Then the whole job and the CLI."""
    import json
    import shutil
    w, d = _write(tmp_path, 120000, 6, seed=41, libs=LIBS_4BAMS)
    (d / "cfg").write_text(w.config_text())
    cfg = api.BamConfig(text=w.config_text())
    cwd = os.getcwd()
    os.chdir(d)
    try:
        assert cfg.bam_files == ["lane1.bam", "lane2.bam", "lane3.bam", "lane4.bam"]
        n4 = _n_bam_columns(cfg, cfg.bam_files)
        assert n4 == w.n
        _n_bam_columns(cfg, cfg.bam_files, "chrB")
        # whole job: device decode of four bams against the job from host-decoded columns
        opts = api.Options()
        host = api.BamStream(cfg, threads=4)
        b = api.ParamBundle.from_stream(opts, cfg, host)
        ctx = api.Context(b)
        ctx.push({k: np.ascontiguousarray(v) for k, v in host.cols.items()})
        want = (ctx.summary(), ctx.finish(), ctx.regions(), ctx.areads())
        ctx.close(); host.close()
        devs = []
        for p in cfg.bam_files:
            devs.append(api.BamDevice(cfg, path=p, after=devs[-1] if devs else None))
        b = api.ParamBundle(opts, cfg.libs, cfg.nbam, np.concatenate([x.rg_lib for x in devs]), np.concatenate([x.rg_bam for x in devs]), cfg.window, len(devs[0].tid_names))
        ctx = api.Context(b)
        ctx.push_bams_n(devs)
        got = (ctx.summary(), ctx.finish(), ctx.regions(), ctx.areads())
        ctx.close()
        for x in devs:
            x.close()
        _assert_same_job(want, got)
        assert len(got[1].sv) > 0
    finally:
        os.chdir(cwd)
    outs = {}
    for mode in ("1", "0"):
        stats = d / ("stats%s.json" % mode)
        p = subprocess.run([util.CLI, "--stats-json", str(stats), "cfg"], cwd=d, env=dict(os.environ, BDK_GPU_DECODE=mode), capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        outs[mode] = util.strip_header(p.stdout)
        assert json.loads(stats.read_text())["device_decode"] == int(mode)
    assert outs["1"] == outs["0"] and outs["1"].count("\n") > 5
    # ties everywhere: three bams with the same records
    os.chdir(d)
    try:
        shutil.copy("lane1.bam", "lane2.bam")
        shutil.copy("lane1.bam", "lane3.bam")
        _n_bam_columns(cfg, cfg.bam_files[:3])
    finally:
        os.chdir(cwd)


def test_per_chromosome_shards_decoded_on_the_device(tmp_path):
    """shard.run_sharded_bams_device: every chromosome through the index and the device decode (-o semantics), against the same
    chromosome through the host decoder; the plan comes from the .bai alone. The index is written by the reference's samtools
    when oracle/_ref has it (else the test is the no-index error)."""
    from breakdancer_b200 import shard
    from oracle import oracle
    libs = [synth.LibSpec("lib_a", "one.bam", 315, 44, 75, ["a1", "a2"])]
    w, d = _write(tmp_path, 90000, 6, seed=12, libs=libs)
    cfg_text = w.config_text()
    cwd = os.getcwd()
    os.chdir(d)
    try:
        cfg = api.BamConfig(text=cfg_text)
        samtools = oracle.REF_SAMTOOLS
        if not os.access(samtools, os.X_OK):
            with pytest.raises(RuntimeError):
                shard.run_sharded_bams_device(cfg, 0, 1, api.Options())
            return
        subprocess.check_call([samtools, "index", "one.bam"])
        got = shard.run_sharded_bams_device(cfg, 0, 1, api.Options())
        assert [t for t, *_ in got] == [0, 1, 2]
        for t, name, summ, table in got:
            want = _job_from_host(cfg, "one.bam", api.Options(chr=name))
            assert bytes(summ) == bytes(want[0]) and table.sv.tobytes() == want[1].sv.tobytes(), name
    finally:
        os.chdir(cwd)


def test_read_groups_the_config_does_not_know_and_records_without_am_tag(tmp_path):
    """A read group missing from the config falls back to the first bam's library (BamConfig.hpp:63-72) on both decoders; without
    AM tags bdqual is MAPQ (Alignment.cpp:12-23)."""
    libs = [synth.LibSpec("lib_a", "one.bam", 315, 44, 75, ["a1", "a2"]), synth.LibSpec("lib_b", "one.bam", 467, 32, 100, ["b1"])]
    w = synth.generate(util.GENOME3, libs, 50000, seed=14, anomaly_frac=0.05)
    bam = str(tmp_path / "one.bam")
    api.write_bam(bam, [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, synth.split_by_bam(w)["one.bam"], write_am=False, level=6)
    full = w.config_text()
    partial = "".join(l for l in full.splitlines(True) if "readgroup:a2" not in l)
    assert partial != full
    for text in (full, partial):
        cfg = api.BamConfig(text=text)
        _compare_columns(cfg, bam)
        want = _job_from_host(cfg, bam, api.Options())
        got, _, _ = _job_from_device(cfg, bam, api.Options())
        _assert_same_job(want, got)
