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
