"""bam2cfg in C++ (bdh_bam2cfg, bin/bam2cfg) against the goldens the UNMODIFIED reference script perl/bam2cfg.pl produced
(tests/golden/bam2cfg/make_golden.sh). The script prints in Perl hash order, so lines are compared as sorted lists."""
import os
import subprocess

import pytest

from breakdancer_b200 import api, synth
from tests import util

GOLD = os.path.join(util.GOLDEN, "bam2cfg")
B2C = os.path.join(util.ROOT, "breakdancer_b200", "bin", "bam2cfg")


def _gold(name):
    return open(os.path.join(GOLD, name)).read().splitlines()


def _lines(text):
    return sorted(text.splitlines())


def test_chr21_flag_histogram_matches_reference_script():
    cwd = os.getcwd()
    os.chdir(util.CHR21)
    try:
        got = api.bam2cfg(["NA19238_chr21_del_inv.bam", "NA19240_chr21_del_inv.bam"], flag_hist=1)
    finally:
        os.chdir(cwd)
    assert _lines(got) == _gold("chr21_g.cfg")


def test_cli_options_match_reference_script():
    p = subprocess.run([B2C, "-q", "10", "-n", "500", "-c", "3", "NA19238_chr21_del_inv.bam"], cwd=util.CHR21, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert _lines(p.stdout) == _gold("chr21_q10_n500_c3.cfg")
    p = subprocess.run([B2C, "-m", "-s", "100", "-v", "0.5", "NA19240_chr21_del_inv.bam"], cwd=util.CHR21, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert _lines(p.stdout) == _gold("chr21_m.cfg")
    assert subprocess.run([B2C], capture_output=True).returncode == 1
    p = subprocess.run([B2C, "missing.bam"], cwd=util.CHR21, capture_output=True, text=True)
    assert p.returncode == 1 and p.stderr.strip()


def test_synthetic_multi_library_bams_match_reference_script(tmp_path):
    w = synth.generate(util.GENOME3, util.LIBS4, 60000, seed=77, anomaly_frac=0.05)
    for bam, cols in synth.split_by_bam(w).items():
        api.write_bam(str(tmp_path / ("syn_" + bam)), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=1)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        got = api.bam2cfg(["syn_normal.bam", "syn_tumor.bam"], flag_hist=1, n_obs=3000)
    finally:
        os.chdir(cwd)
    assert _lines(got) == _gold("syn_g_n3000.cfg")
    # the generated configuration is what the caller's config parser reads: libraries, cut-offs from lower / upper
    cfg = api.BamConfig(text=got)
    assert sorted(cfg.lib_names) == sorted(w.rg_names)          # write_bam's header names every read group's library after the group
    for name, lib in zip(cfg.lib_names, cfg.libs):
        spec = next(l for l in util.LIBS4 if name in l.read_groups)
        assert abs(lib.mean_insertsize - spec.mean) < 3 and abs(lib.std_insertsize - spec.std) < 3
        assert lib.lowercutoff < spec.mean < lib.uppercutoff


def test_unsorted_bam_is_refused(tmp_path):
    w = synth.generate(util.GENOME3, util.LIBS4[:1], 2000, seed=5, anomaly_frac=0.0)
    cols = {k: v[::-1].copy() for k, v in w.cols.items()}
    api.write_bam(str(tmp_path / "rev.bam"), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=1)
    with pytest.raises(api.BdkError, match="sort bam by position"):
        api.bam2cfg([str(tmp_path / "rev.bam")])
