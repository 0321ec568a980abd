"""The product's C++ formatter against the goldens (fed with the oracle's rows, CPU only)."""
import os

import pytest

from breakdancer_b200 import api
from oracle import oracle
from tests import util
from tests.test_oracle_golden import CASES, _chr21


@pytest.mark.parametrize("chr_,cn_lib,af,golden", CASES)
def test_product_formatter_reproduces_goldens(chr_, cn_lib, af, golden):
    b, cols, cfg, tids = _chr21(chr_, cn_lib, af)
    res = oracle.run(b, cols)
    text = api.format_output(b, res.summary, res.table, cfg.lib_names, cfg.bam_files, tids)
    assert text == util.strip_header(open(os.path.join(util.CHR21, golden)).read())


def test_bench_bam_decode_leg_runs_on_the_host():
    """bench.py's bam_decode object (host BAM decode rate of a bounded sample): CPU only, so it is checked here."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(util.ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    r = bench.bam_decode_sample(20000)
    assert r["unit"] == "read-pairs/s" and r["value"] > 0 and r["records_per_s"] == 2 * r["value"]
    assert set(r["stages_s"]) == {"inflate_s", "extract_s", "merge_s"} and r["bam_bytes"] > 0
