"""Parity of the CUDA path (through the C ABI) with the oracle and with the reference's goldens.
Every test here needs a B200: run with  pytest -m gpu."""
import os
import subprocess

import numpy as np
import pytest

from breakdancer_b200 import api, synth
from oracle import oracle
from tests import util
from tests.test_oracle_golden import CASES, _golden_poisson

pytestmark = pytest.mark.gpu


def _check(b, cols, what, **kw):
    ro = oracle.run(b, cols)
    table, summary, regions, areads, rr, support, ctx = util.run_gpu(b, cols, **kw)
    util.assert_result_matches_oracle(ro, table, summary, regions, areads, rr, support, what)
    ctx.close()
    return ro


# ---- the reference's own integration test (integration-test/breakdancer_test.py) through the drop-in CLI
@pytest.mark.parametrize("chr_,cn_lib,af,golden", CASES)
def test_cli_reproduces_chr21_goldens(chr_, cn_lib, af, golden):
    args = (["-a"] if cn_lib else []) + (["-h"] if af else []) + (["-o", chr_] if chr_ else []) + ["inv_del_bam_config"]
    p = subprocess.run([util.CLI] + args, cwd=util.CHR21, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert util.strip_header(p.stdout) == util.strip_header(open(os.path.join(util.CHR21, golden)).read())


def test_cli_bed_dump(tmp_path):
    bed = tmp_path / "out.bed"
    p = subprocess.run([util.CLI, "-g", str(bed), "inv_del_bam_config"], cwd=util.CHR21, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert bed.read_text() == open(os.path.join(util.CHR21, "expected.bed")).read()


def test_cli_fastq_dump(tmp_path):
    prefix = tmp_path / "actual"
    p = subprocess.run([util.CLI, "-o", "21", "-d", str(prefix), "inv_del_bam_config"], cwd=util.CHR21, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert util.strip_header(p.stdout) == util.strip_header(open(os.path.join(util.CHR21, "expected_output")).read())
    for lib in ("H_IJ-NA19238-NA19238-extlibs", "H_IJ-NA19240-NA19240-extlibs"):
        for k in (1, 2):
            got = open(f"{prefix}.{lib}.{k}.fastq").read()
            assert got == open(os.path.join(util.CHR21, f"expected.{lib}.{k}.fastq")).read(), (lib, k)


def test_cli_errors_like_the_reference(tmp_path):
    p = subprocess.run([util.CLI, str(tmp_path / "missing.cfg")], capture_output=True, text=True)
    assert p.returncode == 1 and "Error: no bams files in config file!" in p.stdout
    (tmp_path / "c.cfg").write_text("map:nope.bam\tlib:l\tmean:300\tstd:30\treadlen:75\n")
    p = subprocess.run([util.CLI, "c.cfg"], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 1 and p.stderr.startswith("ERROR: ")


# ---- stage-by-stage parity with the oracle on synthetic data ---------------------------------------
@pytest.mark.parametrize("od", util.OPTION_SETS, ids=lambda d: ",".join(f"{k}={v}" for k, v in d.items()) or "default")
def test_gpu_matches_oracle_option_sweep(od):
    w = synth.generate(util.GENOME3, util.LIBS4, 100000, seed=1, anomaly_frac=0.04, somatic_frac=0.3)
    b, cols, *_ = util.workload_bundle(w, api.Options(**od))
    ro = _check(b, cols, str(od))
    assert len(ro.table.sv) > 5


@pytest.mark.parametrize("seed", [2, 3, 4])
def test_gpu_matches_oracle_seeds(seed):
    w = synth.generate(util.GENOME3, util.LIBS4, 200000, seed=seed, anomaly_frac=0.05, somatic_frac=0.3)
    b, cols, *_ = util.workload_bundle(w, api.Options(min_read_pair=1, score_threshold=-100))
    _check(b, cols, f"seed {seed}")


def test_gpu_chunked_pushes_and_device_push():
    w = synth.generate(util.GENOME3, util.LIBS4, 150000, seed=6, anomaly_frac=0.04)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    _check(b, cols, "3 chunks", chunks=3)
    _check(b, cols, "17 chunks", chunks=17)
    _check(b, cols, "device push", device_push=True)
    _check(b, cols, "pinned host columns, qlen/qid zero-copy", pinned=True)


def test_gpu_many_read_groups_and_library_bam_pairs():
    """More (library, bam) pairs than private counter columns (warp-vote fallback) and more read groups than fit
    the shared-memory table: same results as the oracle."""
    libs = [synth.LibSpec(f"lib{i:02d}", f"b{i % 3}.bam", 300 + i, 30, 75, [f"rg{i:02d}_{j}" for j in range(2)]) for i in range(40)]
    w = synth.generate(util.GENOME3, libs, 120000, seed=11, anomaly_frac=0.05)
    for od in (dict(), dict(CN_lib=True)):
        b, cols, *_ = util.workload_bundle(w, api.Options(**od))
        _check(b, cols, f"40 libraries / 80 read groups {od}")
    libs = [synth.LibSpec(f"L{i}", "one.bam", 300 + 10 * i, 30, 75, [f"g{i}_{j}" for j in range(600)]) for i in range(2)]
    w = synth.generate(util.GENOME3, libs, 60000, seed=12, anomaly_frac=0.05)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    assert len(b.rg_lib) > 1023
    _check(b, cols, "1200 read groups")


def test_gpu_single_bam_single_key_path():
    w = synth.config2(300000, seed=20260101)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    ro = _check(b, cols, "config2-shaped")
    assert len(ro.table.sv) > 10


def test_gpu_edge_cases_empty_ragged_tiny():
    w = synth.generate(util.GENOME3, util.LIBS4, 6000, seed=5, anomaly_frac=0.05)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    for n in (0, 1, 2, 3, 5, 127, 128, 129, 511, 512, 513, 4095, 4096, 4097, 8191, 12000):
        sub = {k: np.ascontiguousarray(v[:n]) for k, v in cols.items()}
        _check(b, sub, f"n={n}")


def test_gpu_all_anomalous_staging_overflow_and_giant_region():
    w = synth.generate([("c", 400000)], [synth.LibSpec("l", "b.bam", 300, 30, 75, ["g"])], 40000, seed=9,
                       anomaly_frac=1.0, cluster_frac=1.0, cluster_mean_pairs=400, odd_frac=0.0)
    for od in (dict(), dict(seq_coverage_lim=1), dict(min_read_pair=1, score_threshold=-1)):
        b, cols, *_ = util.workload_bundle(w, api.Options(**od))
        _check(b, cols, str(od))


def test_gpu_segment_overflow_retry(monkeypatch):
    """Per-CTA output segments far too small (forced): the push detects it, grows them and runs again."""
    monkeypatch.setenv("BDK_SEG_CAP_MIN", "16")
    w = synth.generate(util.GENOME3, util.LIBS4, 150000, seed=21, anomaly_frac=0.2)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    _check(b, cols, "tiny segments, one push")
    _check(b, cols, "tiny segments, 5 pushes", chunks=5)
    w = synth.config2(400000, seed=3)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    _check(b, cols, "tiny segments, single-key path")


@pytest.mark.parametrize("od", [dict(), dict(CN_lib=True), dict(transchr_rearrange=True), dict(CN_lib=True, min_map_qual=10, cut_sd=2)],
                         ids=lambda d: ",".join(f"{k}={v}" for k, v in d.items()) or "default")
def test_gpu_general_multi_key_classifier_matches_oracle(monkeypatch, od):
    """Two bams / four libraries normally take the four-key variant of the classify kernel (key bit planes); the general
    variant (up to 64 keys, warp votes) is forced here on the same inputs. Both must give the oracle's stream and counts."""
    monkeypatch.setenv("BDK_K1_GENERAL", "1")
    w = synth.generate(util.GENOME3, util.LIBS4, 150000, seed=77, anomaly_frac=0.05, somatic_frac=0.3)
    b, cols, *_ = util.workload_bundle(w, api.Options(**od))
    _check(b, cols, f"general K1 {od}")
    _check(b, cols, f"general K1 {od}, 3 pushes", chunks=3)


def _check_packed(b, cols, what, pinned_side=False):
    """The same job fed as packed runs (12-byte wire format, include/bdk.h: bdk_packed), one per reference sequence."""
    ro = oracle.run(b, cols)
    runs = api.pack_runs(cols)
    ctx = api.Context(b, 0)
    nexc = 0
    for r in runs:
        ctx.push_packed(r)
        nexc += int(r.view.nx)
    summary = ctx.summary()
    table = ctx.finish()
    ar, rr = ctx.areads()
    util.assert_result_matches_oracle(ro, table, summary, ctx.regions(), ar, rr, ctx.support(), what)
    ctx.close()
    return ro, nexc


@pytest.mark.parametrize("od", [dict(), dict(CN_lib=True), dict(transchr_rearrange=True), dict(max_sd=10000)],
                         ids=lambda d: ",".join(f"{k}={v}" for k, v in d.items()) or "default")
def test_gpu_packed_runs_match_oracle(od):
    """Records pushed in the packed wire format give the oracle's result: three chromosomes (three runs), inter-chromosomal
    mates and inserts beyond 16 bits travel as exceptions."""
    w = synth.generate(util.GENOME3, util.LIBS4, 120000, seed=91, anomaly_frac=0.06, somatic_frac=0.3)
    b, cols, *_ = util.workload_bundle(w, api.Options(**od))
    ro, nexc = _check_packed(b, cols, f"packed {od}")
    assert nexc > 100 and len(ro.table.sv) > 5


def test_gpu_packed_chunked_run_with_exceptions_in_every_chunk(monkeypatch):
    """A run longer than one 8 Mi-record chunk (config-2 shape, 9 M pairs): exceptions are cut per chunk."""
    w = synth.config2(9_000_000, seed=5)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    ro, nexc = _check_packed(b, cols, "packed 18 M records")
    assert nexc > 1000 and len(cols["pos"]) > 2 * 8 * 1024 * 1024


def test_gpu_packed_rejects_a_run_over_two_chromosomes():
    w = synth.generate(util.GENOME3, util.LIBS4, 5000, seed=1)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    with pytest.raises(api.BdkError, match="one reference sequence"):
        api.PackedRun(api.make_soa(cols), len(cols["pos"]))


def test_gpu_read_name_with_three_reads_is_left_unpaired():
    """A read name that occurs three times among the anomalous reads (bams with overlapping names, a key collision) no longer
    aborts the job: its reads stay unpaired. Expected result = the oracle on the same records with those reads renamed apart."""
    w = synth.generate(util.GENOME3, util.LIBS4, 60000, seed=15, anomaly_frac=0.05, somatic_frac=0.3)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    ro = oracle.run(b, cols)
    q = ro.areads["qid"]
    u, cnt = np.unique(q, return_counts=True)
    pairs = u[cnt == 2]
    rec = ro.areads["record"]
    cols3 = {k: v.copy() for k, v in cols.items()}
    want = {k: v.copy() for k, v in cols.items()}
    fresh = np.uint64(1) << np.uint64(63)
    for t in range(5):                                  # five names get a third read (the later read of another anomalous pair)
        name, other = pairs[10 + 2 * t], pairs[11 + 2 * t]
        z = rec[q == other][1]
        cols3["qid"][z] = name
        for i, r in enumerate(list(rec[q == name]) + [z]):
            want["qid"][r] = fresh + np.uint64(10 * t + i)
    ctx = api.Context(b, 0)
    ctx.push(cols3)
    summary, table = ctx.summary(), ctx.finish()
    assert ctx.duplicate_names() == 5
    rw = oracle.run(b, want)
    ar, rr = ctx.areads()
    assert np.array_equal(ar["record"], rw.areads["record"]) and np.array_equal(rr, rw.aread_region)
    util.assert_tables_equal(rw.table, table, "three reads of one name")
    assert np.array_equal(ctx.support(), rw.sv_of_read)
    ctx.close()


def test_gpu_more_rows_than_the_first_result_copy(monkeypatch):
    """More SV rows than the first device-to-host copy was sized for (forced: 16 rows): the rest comes with a second copy."""
    monkeypatch.setenv("BDK_ROWS_GUESS", "16")
    w = synth.generate(util.GENOME3, util.LIBS4, 200000, seed=31, anomaly_frac=0.05, somatic_frac=0.3)
    b, cols, *_ = util.workload_bundle(w, api.Options(min_read_pair=1, score_threshold=-100))
    ro = _check(b, cols, "many rows")
    assert len(ro.table.sv) > 500


def test_gpu_reset_and_reuse_is_idempotent():
    w = synth.generate(util.GENOME3, util.LIBS4, 80000, seed=8, anomaly_frac=0.04)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    ctx = api.Context(b, 0)
    tables = []
    for _ in range(3):
        ctx.reset()
        ctx.push(cols)
        tables.append(ctx.finish())
    for t in tables[1:]:
        assert t.sv.tobytes() == tables[0].sv.tobytes() and np.array_equal(t.lib_count, tables[0].lib_count)
    assert ctx.kernel_launches() > 10
    ctx.close()


def test_gpu_poisson_tail_matches_boost_known_answers():
    w = synth.generate(util.GENOME3, util.LIBS4, 100, seed=1)
    b, *_ = util.workload_bundle(w, api.Options())
    ctx = api.Context(b, 0)
    lam, k, lp = _golden_poisson()
    got = ctx.poisson_logsf(lam, k)
    util.assert_logp_close(lp, got)
    ctx.close()


def test_gpu_rejects_read_group_without_library():
    w = synth.generate(util.GENOME3, util.LIBS4, 5000, seed=1)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    b.rg_lib[0] = -1
    ctx = api.Context(b, 0)
    with pytest.raises(api.BdkError, match="library index out of range"):
        ctx.push(cols)
    ctx.close()


def test_gpu_config2_two_million_pairs_matches_oracle():
    """BASELINE config 2 shape (single library chr1-like, DEL-only) at 2 M pairs, bit-exact vs the oracle."""
    w = synth.config2(2_000_000, seed=20260101)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    ro = _check(b, cols, "config2 2M pairs")
    assert len(ro.table.sv) > 100 and (ro.table.sv["flag"] == 2).sum() > 50   # the planted deletions


def test_gpu_config3_shape_matches_oracle():
    """BASELINE config 3 shape (2 bams x 2 libraries, all five SV types, -c 3 -q 35) at 1.5 M pairs."""
    w = synth.config3(1_500_000, seed=20260102)
    b, cols, *_ = util.workload_bundle(w, api.Options(cut_sd=3, min_map_qual=35))
    ro = _check(b, cols, "config3 1.5M pairs")
    flags = set(ro.table.sv["flag"].tolist())
    assert {1, 2, 3, 4, 8} <= flags


def test_config3_device_generator_matches_oracle():
    """The generator of the full-size configs[2] workload (torch, on the device) at a size the oracle handles."""
    import torch
    from breakdancer_b200 import synth_torch
    cols = synth_torch.config3_device(1_500_000, seed=11, device=torch.device("cuda", 0))
    b, _ = synth_torch.config3_bundle()
    ro = _check(b, synth_torch.to_numpy(cols), "config3_device 1.5M pairs")
    assert len(ro.table.sv) > 300 and len(set(ro.table.sv["flag"].tolist())) >= 5


# ---- alternative mechanics of K3 / K4, forced -------------------------------------------------------------------------
K4_FORCED = {
    "one_launch_per_sweep_instead_of_the_cooperative_kernel": dict(BDK_K4_HOST_LOOP="1"),
    "every_window_through_the_big_window_kernel": dict(BDK_K4W_CAP="0"),
    "windows_over_8_edges_through_the_big_window_kernel": dict(BDK_K4_HOST_LOOP="1", BDK_K4W_CAP="8"),
}


@pytest.mark.parametrize("name", list(K4_FORCED))
def test_gpu_k4_forced_paths_match_oracle(monkeypatch, name):
    """The sweeps over the table of deletion windows as one launch each (what a shared GPU falls back to) and the kernel
    that orders the calls of windows with too many followed edges for one warp's shared memory (a CTA each, sorted in
    global memory): all must give the oracle's result."""
    for k, v in K4_FORCED[name].items():
        monkeypatch.setenv(k, v)
    w = synth.generate(util.GENOME3, util.LIBS4, 120000, seed=41, anomaly_frac=0.08, somatic_frac=0.3)
    for od in (dict(), dict(min_read_pair=1, score_threshold=-100), dict(buffer_size=3), dict(transchr_rearrange=True), dict(CN_lib=True, chr="chrA")):
        b, cols, *_ = util.workload_bundle(w, api.Options(**od))
        ro = _check(b, cols, f"{name} {od}")
        assert len(ro.table.sv) > 3
    import torch
    from breakdancer_b200 import synth_torch
    cols = synth_torch.config3_device(1_000_000, seed=13, device=torch.device("cuda", 0))     # dense noise: long chains of followed edges
    b, _ = synth_torch.config3_bundle()
    _check(b, synth_torch.to_numpy(cols), f"{name} config3-shaped")


def test_gpu_config3_shape_thirty_million_pairs_matches_oracle():
    """configs[2]'s generator at a tenth of its size (30 M pairs, 1.3 M anomalous reads, ~900 flush windows), bit-exact vs the
    oracle; the full 300 M pairs are compared in test_gpu_fullsize.py (BDK_FULLSIZE_ORACLE=1)."""
    import torch
    from breakdancer_b200 import synth_torch
    cols = synth_torch.config3_device(30_000_000, seed=20260102, device=torch.device("cuda", 0))
    b, _ = synth_torch.config3_bundle()
    ro = _check(b, synth_torch.to_numpy(cols), "config3-shaped 30M pairs")
    assert len(ro.table.sv) > 10000


def test_gpu_config3_shape_eight_million_pairs_matches_oracle():
    """Dense config-3-shaped data at 8 M pairs (tens of thousands of followed edges per chromosome), bit-exact vs the oracle."""
    import torch
    from breakdancer_b200 import synth_torch
    cols = synth_torch.config3_device(8_000_000, seed=17, device=torch.device("cuda", 0))
    b, _ = synth_torch.config3_bundle()
    ro = _check(b, synth_torch.to_numpy(cols), "config3-shaped 8M pairs")
    assert len(ro.table.sv) > 1000
