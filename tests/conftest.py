import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build libbdk.so / the oracle / the host simulation if they are missing (CPU-only compile)."""
    need = [os.path.join(ROOT, "breakdancer_b200", "libbdk.so"), os.path.join(ROOT, "oracle", "_build", "libbdoracle.so"),
            os.path.join(ROOT, "breakdancer_b200", "bin", "breakdancer_max")]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call(["make", "-s", "-j8", "-C", ROOT])
    yield
