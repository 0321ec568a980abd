"""The table-driven DEFLATE decoder of the BAM reader (csrc/host/fast_inflate.hpp): differential test against zlib on generated
streams of every block type, and corrupted streams under AddressSanitizer / UBSan (tests/hostsim/inflate_fuzz.cpp); then the
reader itself: same columns with the fast decoder and with zlib only."""
import os
import subprocess

import numpy as np

from breakdancer_b200 import api, synth
from tests import util


def test_differential_and_corruption_fuzz_under_sanitizers():
    src = os.path.join(util.ROOT, "tests", "hostsim", "inflate_fuzz.cpp")
    exe = os.path.join(util.ROOT, "tests", "_build", "inflate_fuzz")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer",
                           src, "-o", exe, "-lz"])
    p = subprocess.run([exe, "400"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "refused_good=0 mismatched=0" in p.stdout, p.stdout


def test_reader_gives_the_same_columns_with_and_without_the_fast_decoder(tmp_path):
    w = synth.generate(util.GENOME3, util.LIBS4, 30000, seed=8, anomaly_frac=0.05)
    for level in (1, 6, 9):
        d = tmp_path / f"l{level}"
        d.mkdir()
        for bam, cols in synth.split_by_bam(w).items():
            api.write_bam(str(d / bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=level)
        (d / "cfg").write_text(w.config_text())
    out = {}
    for mode in ("fast", "zlib"):
        code = ("import os, sys, numpy as np\n"
                "sys.path.insert(0, %r)\n"
                "from breakdancer_b200 import api\n"
                "res = []\n"
                "for level in (1, 6, 9):\n"
                "    os.chdir(os.path.join(%r, 'l%%d' %% level))\n"
                "    cfg = api.BamConfig(path='cfg'); st = api.BamStream(cfg, threads=2)\n"
                "    res.append(np.concatenate([v.view(np.uint8).ravel() for k, v in sorted(st.cols.items())]))\n"
                "np.save(sys.argv[1], np.concatenate(res))\n") % (util.ROOT, str(tmp_path))
        env = dict(os.environ)
        env["BDK_FAST_INFLATE"] = "1" if mode == "fast" else "0"
        f = str(tmp_path / (mode + ".npy"))
        subprocess.check_call(["python", "-c", code, f], env=env)
        out[mode] = np.load(f)
    assert len(out["fast"]) > 100000 and np.array_equal(out["fast"], out["zlib"])
