"""breakdancer_b200 -- B200-native BreakDancerMax hot path.

The product is native: ``libbdk.so`` (hand-written sm_100a CUDA kernels behind the C ABI of
``include/bdk.h`` plus the C++ host side of ``include/bdk_host.h``) and the drop-in executable
``bin/breakdancer_max``.  ``api`` is the ctypes face used by tests and bench.py; ``synth`` makes
synthetic inputs.  Nothing here falls back to a CPU implementation.
"""
from . import api  # noqa: F401

__all__ = ["api"]
