"""The C-ABI library loads and exports every symbol the headers declare; no compute without a GPU."""
import ctypes as C
import os
import re

import pytest

from breakdancer_b200 import api
from tests import util


def _declared(header):
    text = open(os.path.join(util.ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(bd[kh]_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    L = api.load_library()
    assert L._bdk_missing == []
    for header in ("bdk.h", "bdk_host.h"):
        for sym in _declared(header):
            assert hasattr(L, sym), f"{sym} declared in include/{header} but not exported"


def test_python_binding_covers_every_declared_symbol():
    L = api.load_library()
    declared = _declared("bdk.h") | _declared("bdk_host.h")
    assert declared <= set(L._bdk_symbols), declared - set(L._bdk_symbols)


def test_struct_layouts_match_header_sizes():
    assert C.sizeof(api.Sv) == 80
    assert C.sizeof(api.Lib) == 28
    assert api.REGION_DTYPE.itemsize == 36 and api.AREAD_DTYPE.itemsize == 32
    assert C.sizeof(api.Packed) == 5 * 8 + 8 + 8 + 6 * 8
    assert C.sizeof(api.SummaryT) == 8 + 16 + 4 * 64 + 8 * 64 + 4 * 255 + 4 * 11 * 255 + 4 * 255 + 4 * 255


def test_version_string():
    assert b"sm_100a" in api.load_library().bdk_version()


@pytest.mark.skipif(util.have_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    w = util.synth.generate(util.GENOME3, util.LIBS4, 1000, seed=1)
    b, cols, *_ = util.workload_bundle(w, api.Options())
    with pytest.raises(api.BdkError) as e:
        api.Context(b, 0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_create_rejects_bad_arguments():
    w = util.synth.generate(util.GENOME3, util.LIBS4, 1000, seed=1)
    b, *_ = util.workload_bundle(w, api.Options(min_read_pair=0))
    with pytest.raises(api.BdkError) as e:
        api.Context(b, 0)
    assert "min_read_pair" in str(e.value)
