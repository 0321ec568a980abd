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
