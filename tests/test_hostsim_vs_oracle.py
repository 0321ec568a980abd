"""The device logic (bdk_logic.h / bdk_finalize.h compiled for the host) run through the GPU
pipeline's stage decomposition must reproduce the oracle exactly: anomalous stream, regions, SV rows,
supporting reads. This is the CPU-side check of the parallel formulation; the CUDA mechanics are
checked on the GPU (tests marked gpu)."""
import numpy as np
import pytest

from breakdancer_b200 import api, synth
from oracle import oracle
from tests import util


@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("od", util.OPTION_SETS, ids=lambda d: ",".join(f"{k}={v}" for k, v in d.items()) or "default")
def test_stage_decomposition_matches_oracle(seed, od):
    w = synth.generate(util.GENOME3, util.LIBS4, 100000, seed=seed, anomaly_frac=0.04, somatic_frac=0.3)
    b, cols, *_ = util.workload_bundle(w, api.Options(**od))
    ro = oracle.run(b, cols)
    rh = util.run_hostsim(b, cols)
    assert np.array_equal(ro.rec_class, rh.rec_class)
    assert np.array_equal(ro.region_alive, rh.region_alive)
    util.assert_result_matches_oracle(ro, rh.table, rh.summary, rh.regions, rh.areads, rh.aread_region, rh.sv_of_read, str(od))
    assert len(ro.table.sv) > 5


WALK_VARIANTS = {
    "jacobi_sweeps_from_the_guess": dict(),                                       # every region against the previous table
    "in_place_sweeps_descending": dict(HOSTSIM_INPLACE="1"),                      # table rewritten while it is read, regions descending
    "no_starting_guess": dict(HOSTSIM_NO_GUESS="1"),                              # any starting table must converge to the same result
    "in_place_no_guess": dict(HOSTSIM_INPLACE="1", HOSTSIM_NO_GUESS="1"),
}


@pytest.mark.parametrize("variant", list(WALK_VARIANTS))
def test_connection_walk_variants_match_oracle(monkeypatch, variant):
    """The closed form of the connection walk (bdk_logic.h, K4 second formulation): the table of deletion windows as a fixed
    point reached in any sweep order and from any starting table, windows evaluated independently (in descending order), calls
    evaluated independently. Dense config-3-shaped data (long followed-edge chains) and sparse data."""
    import torch
    from breakdancer_b200 import synth_torch
    for k, v in WALK_VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    cols = synth_torch.to_numpy(synth_torch.config3_device(1_500_000, 23, torch.device("cpu")))
    b, _ = synth_torch.config3_bundle()
    ro, rh = oracle.run(b, cols), util.run_hostsim(b, cols)
    util.assert_result_matches_oracle(ro, rh.table, rh.summary, rh.regions, rh.areads, rh.aread_region, rh.sv_of_read, variant + " config3")
    assert len(ro.table.sv) > 300
    w = synth.generate(util.GENOME3, util.LIBS4, 100000, seed=3, anomaly_frac=0.06, somatic_frac=0.3)
    for od in (dict(), dict(min_read_pair=1, score_threshold=-100), dict(buffer_size=3), dict(transchr_rearrange=True), dict(chr="chrB")):
        b, cols, *_ = util.workload_bundle(w, api.Options(**od))
        ro, rh = oracle.run(b, cols), util.run_hostsim(b, cols)
        util.assert_result_matches_oracle(ro, rh.table, rh.summary, rh.regions, rh.areads, rh.aread_region, rh.sv_of_read, f"{variant} {od}")


def test_edge_cases_empty_and_tiny():
    w = synth.generate(util.GENOME3, util.LIBS4, 50, seed=5, anomaly_frac=0.0, odd_frac=0.0)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    for n in (0, 1, 2, 7):
        sub = {k: np.ascontiguousarray(v[:n]) for k, v in cols.items()}
        ro, rh = oracle.run(b, sub), util.run_hostsim(b, sub)
        util.assert_result_matches_oracle(ro, rh.table, rh.summary, rh.regions, rh.areads, rh.aread_region, rh.sv_of_read, f"n={n}")


def test_all_anomalous_dense_cluster():
    """Every read anomalous and one giant region: exercises -x coverage rejection and big components."""
    w = synth.generate([("c", 400000)], [synth.LibSpec("l", "b.bam", 300, 30, 75, ["g"])], 20000, seed=9,
                       anomaly_frac=1.0, cluster_frac=1.0, cluster_mean_pairs=400, odd_frac=0.0)
    for od in (dict(), dict(seq_coverage_lim=1), dict(min_read_pair=1, score_threshold=-1)):
        b, cols, *_ = util.workload_bundle(w, api.Options(**od))
        ro, rh = oracle.run(b, cols), util.run_hostsim(b, cols)
        util.assert_result_matches_oracle(ro, rh.table, rh.summary, rh.regions, rh.areads, rh.aread_region, rh.sv_of_read, str(od))


def test_classify_hot_equals_classify_record():
    """The streaming kernel's 4-decision classifier equals the full classifier on every flag word x
    tid/pos/isize/mapq relation x option combination (about 1.1e8 cases, in C)."""
    import ctypes as C
    hs = util.hostsim_lib()
    hs.hostsim_classify_hot_check.restype = C.c_long
    hs.hostsim_classify_hot_check.argtypes = [C.POINTER(C.c_long)]
    n = C.c_long(0)
    assert hs.hostsim_classify_hot_check(C.byref(n)) == 0
    assert n.value > 10 ** 7


def test_classifier_truth_table():
    """pe_classify truth table (reference TestIlluminaPEReadClassifier.cpp only prints it; asserted here
    against IlluminaPEReadClassifier.cpp:13-101 by enumeration through the oracle's classifier)."""
    hs = util.hostsim_lib()
    upper, lower = 400.0, 200.0
    for flag in range(0, 0x800):
        if flag & 0x100:
            continue
        for (tid, mtid) in ((1, 1), (1, 2)):
            for (pos, mpos) in ((100, 500), (500, 100), (100, 100)):
                for isz in (100, 200, 300, 400, 401, -1000):
                    cr = hs.hostsim_classify(pos, mpos, tid, mtid, isz, flag, 60, upper, lower, 35, 10 ** 9, 0, 0)
                    a = abs(isz)
                    dup, paired, unm, munm = flag & 0x400, flag & 1, flag & 4, flag & 8
                    if dup or not paired:
                        exp = "NA"
                    elif unm:
                        exp = "UNMAPPED"
                    elif munm:
                        exp = "MATE_UNMAPPED"
                    elif tid != mtid:
                        exp = "ARP_CTX"
                    else:
                        rr, mr = bool(flag & 0x10), bool(flag & 0x20)
                        if rr == mr:
                            exp = "ARP_RR" if rr else "ARP_FF"
                        elif (pos < mpos) == rr:
                            exp = "ARP_RF"
                        elif a > upper:
                            exp = "ARP_LARGE_INSERT"
                        elif a < lower:
                            exp = "ARP_SMALL_INSERT"
                        else:
                            exp = "NORMAL_FR"
                    kept = exp not in ("NA", "UNMAPPED", "MATE_UNMAPPED")
                    assert bool(cr & 16) == kept, (flag, exp)
                    if kept:
                        final = "ARP_FF" if exp == "ARP_RR" else exp     # main pass merges RR into FF
                        assert api.FLAG_NAMES[cr & 0xF] == final
                        assert bool(cr & 64) == (final != "NORMAL_FR")
                        hist = (cr >> 8) & 0xF
                        assert hist == (0 if exp == "NORMAL_FR" else api.FLAG_NAMES.index(exp))
