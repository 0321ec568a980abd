"""Pins the oracle row-for-row against the UNMODIFIED reference binary (oracle/_ref/breakdancer-max) on
seeded synthetic BAMs that exercise what the chr21 goldens do not: many flush windows, CTX/ITX,
-t, -l, -f, -m, several libraries and bams, scores below the cap."""
import os

import pytest

from breakdancer_b200 import api, synth
from oracle import oracle
from tests import util

pytestmark = pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not built (run oracle/build_ref.sh)")

CASES = [[], ["-a", "-h"], ["-t"], ["-l"], ["-f"], ["-r", "1", "-y", "0"], ["-q", "10", "-c", "2"], ["-b", "3"],
         ["-o", "chrB", "-a", "-h"], ["-m", "10000"], ["-s", "50", "-x", "2"], ["-t", "-a", "-h", "-y", "10"]]


def _opts(args):
    o = api.Options()
    it = iter(args)
    names = {"-s": "min_len", "-c": "cut_sd", "-m": "max_sd", "-q": "min_map_qual", "-r": "min_read_pair",
             "-x": "seq_coverage_lim", "-b": "buffer_size", "-y": "score_threshold"}
    flags = {"-t": "transchr_rearrange", "-f": "fisher", "-l": "Illumina_long_insert", "-a": "CN_lib", "-h": "print_AF"}
    for a in it:
        if a == "-o":
            o.chr = next(it)
        elif a in names:
            setattr(o, names[a], int(next(it)))
        else:
            setattr(o, flags[a], True)
    return o


@pytest.fixture(scope="module")
def bam_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("synbam")
    w = synth.generate(util.GENOME3, util.LIBS4, 120000, seed=11, anomaly_frac=0.04, somatic_frac=0.3)
    cwd = os.getcwd()
    os.chdir(d)
    try:
        for bam, cols in synth.split_by_bam(w).items():
            api.write_bam(bam, [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols)
            os.system(f"{oracle.REF_SAMTOOLS} index {bam}")
        open("cfg", "w").write(w.config_text())
    finally:
        os.chdir(cwd)
    return str(d)


@pytest.mark.parametrize("args", CASES, ids=lambda a: " ".join(a) or "default")
def test_oracle_matches_reference_binary(bam_dir, args):
    opts = _opts(args)
    ref = oracle.run_reference(list(args) + ["cfg"], bam_dir)
    cwd = os.getcwd()
    os.chdir(bam_dir)
    try:
        cfg = api.BamConfig(path="cfg", cut_sd=opts.cut_sd)
        st = api.BamStream(cfg, region=opts.chr)
        b = api.ParamBundle.from_stream(opts, cfg, st)
        res, text = oracle.run_text(b, {k: v.copy() for k, v in st.cols.items()}, cfg.lib_names, cfg.bam_files, st.tid_names)
    finally:
        os.chdir(cwd)
    assert text == ref
    assert len(res.table.sv) > 20


def test_oracle_scores_match_traced_reference(bam_dir):
    """Raw log-probabilities: the tracing build of the reference prints every Poisson tail it evaluates;
    with -y -1000 every scored call is emitted, so rows and trace lines align."""
    import numpy as np
    import subprocess
    trace = os.path.join(oracle.REF_DIR, "breakdancer-max-trace")
    args = ["-y", "-1000", "-r", "1"]
    p = subprocess.run([trace] + args + ["cfg"], cwd=bam_dir, capture_output=True, text=True, check=True)
    tr = [l.split("\t") for l in p.stderr.splitlines() if l.startswith("POISSON")]
    lam = np.array([float(t[1]) for t in tr]); lp = np.array([float(t[3]) for t in tr])
    opts = _opts(args)
    cwd = os.getcwd()
    os.chdir(bam_dir)
    try:
        cfg = api.BamConfig(path="cfg")
        st = api.BamStream(cfg)
        b = api.ParamBundle.from_stream(opts, cfg, st)
        res = oracle.run(b, {k: v.copy() for k, v in st.cols.items()})
    finally:
        os.chdir(cwd)
    nl = (res.table.lib_count > 0).sum(axis=1)
    assert nl.sum() == len(lp) and len(res.table.sv) > 100
    off = np.concatenate([[0], np.cumsum(nl)])
    want = np.array([lp[off[i]:off[i + 1]].sum() for i in range(len(nl))])
    fin = np.isfinite(want)
    assert np.max(np.abs(want[fin] - res.table.sv["logp"][fin])) < 1e-9
    assert lam.min() > 0
