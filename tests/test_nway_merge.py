"""The n-way merge order used for three or more bams decoded on the device (csrc/host/nway_merge.hpp) against the reference's
formulation of BamMerger's priority queue (tests/hostsim/nway_merge_check.cpp), under AddressSanitizer / UBSan."""
import os
import subprocess

from tests import util


def test_nway_merge_order_equals_the_priority_queue_over_streams():
    src = os.path.join(util.ROOT, "tests", "hostsim", "nway_merge_check.cpp")
    exe = os.path.join(util.ROOT, "tests", "_build", "nway_merge_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", src, "-o", exe])
    p = subprocess.run([exe, "800"], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.startswith("ok"), p.stdout + p.stderr


def test_host_decoder_merges_four_bams_like_the_plain_priority_queue(tmp_path, monkeypatch):
    """The host decoder's merge of four bams (nway_merge.hpp: the fixed-array heap and, with eight threads, the parallel form) against
    the same decoder with the plain std::priority_queue (BDK_MERGE_HEAP=1): every column, also when three bams hold the same records
    (ties everywhere: the order among equal keys is the heap's)."""
    import shutil
    import numpy as np
    from breakdancer_b200 import api, synth
    libs = [synth.LibSpec("lane1", "lane1.bam", 315, 44, 75, ["l1a", "l1b"]), synth.LibSpec("lane2", "lane2.bam", 312, 43, 75, ["l2"]),
            synth.LibSpec("lane3", "lane3.bam", 467, 32, 75, ["l3"], tumor=True), synth.LibSpec("lane4", "lane4.bam", 476, 29, 100, ["l4"], tumor=True)]
    w = synth.generate(util.GENOME3, libs, 600000, seed=43, anomaly_frac=0.05)
    for bam, cols in synth.split_by_bam(w).items():
        api.write_bam(str(tmp_path / bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=1)
    cfg = api.BamConfig(text=w.config_text())
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for ties in (False, True):
            if ties:
                shutil.copy("lane1.bam", "lane2.bam")
                shutil.copy("lane1.bam", "lane3.bam")
            got = {}
            for mode in ("new", "plain"):
                if mode == "plain":
                    monkeypatch.setenv("BDK_MERGE_HEAP", "1")
                else:
                    monkeypatch.delenv("BDK_MERGE_HEAP", raising=False)
                for threads in ((8, 2) if mode == "new" else (2,)):
                    s = api.BamStream(cfg, threads=threads)
                    got[(mode, threads)] = {k: v.copy() for k, v in s.cols.items()}
                    s.close()
            want = got[("plain", 2)]
            assert len(want["pos"]) >= 1_000_000
            for key, cols in got.items():
                for k, v in want.items():
                    assert np.array_equal(v, cols[k]), (key, k, ties)
    finally:
        os.chdir(cwd)
