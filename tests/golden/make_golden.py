"""Regenerates the committed golden vectors that come from the reference itself.

Run here (where /root/reference and oracle/_ref exist):  python tests/golden/make_golden.py

  poisson_logp.tsv   lambda, k, log(cdf(complement(poisson(lambda), k))) from the reference's vendored
                     boost 1.54 (oracle/_ref/score_ref, source oracle/score_ref.cpp) -- pins
                     ComputeProbScore's Poisson tail (BreakDancer.cpp:62-68), which the chr21 goldens do
                     not pin because all their scores are capped at 99.
  chisq_sf.tsv       ndf, x, cdf(complement(chi_squared(ndf), x)) -- the Fisher branch (BreakDancer.cpp:73-76)
  chr21_poisson_trace.tsv   lambda, k, log p of every Poisson tail the reference evaluates on the bundled
                     chr21 BAMs (oracle/_ref/breakdancer-max-trace, source oracle/ref_trace.cpp)
  chr21/*            copied verbatim from /root/reference/test-data (BAMs, config, expected outputs)
  synth_ref_*.txt    stdout of the reference binary on seeded synthetic BAMs (see SYNTH_CASES)
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    score = os.path.join(ROOT, "oracle", "_ref", "score_ref")
    rng = np.random.default_rng(20260101)
    rows = [(0.00099607400381935022, 1), (0.00079451299019851805, 1), (0.0073383139403436276, 21),
            (0.00031993811685846363, 2), (421.89173009736072, 15), (1e-10, 1), (1e-10, 5), (1.0, 1), (2.5, 2)]
    for _ in range(400):
        lam = float(10 ** rng.uniform(-10, 3))
        k = int(rng.integers(1, 80))
        rows.append((lam, k))
    for k in (1, 2, 3, 5, 10, 20, 50, 100, 200):
        for lam in (k * 0.5, k * 0.9, k + 0.5, k + 1.0, k + 1.5, k * 2.0, k * 10.0):
            rows.append((float(lam), k))
    inp = "".join(f"{lam!r} {k}\n" for lam, k in rows)
    out = subprocess.run([score], input=inp, capture_output=True, text=True, check=True).stdout
    open(os.path.join(HERE, "poisson_logp.tsv"), "w").write(out)
    rows = []
    for n in (1, 2, 3, 4, 8):
        for x in (0.1, 1.0, 5.0, 20.0, 50.0, 100.0, 190.0, 250.0):
            rows.append((2 * n, 2 * x))
    inp = "".join(f"{a} {b!r}\n" for a, b in rows)
    out = subprocess.run([score, "f"], input=inp, capture_output=True, text=True, check=True).stdout
    open(os.path.join(HERE, "chisq_sf.tsv"), "w").write(out)
    # every Poisson tail the reference evaluates on the bundled chr21 data (tracing build of the reference)
    trace = os.path.join(ROOT, "oracle", "_ref", "breakdancer-max-trace")
    p = subprocess.run([trace, "-o", "21", "inv_del_bam_config"], cwd=os.path.join(HERE, "chr21"), capture_output=True, text=True, check=True)
    lines = [l.split("\t", 1)[1] for l in p.stderr.splitlines() if l.startswith("POISSON")]
    open(os.path.join(HERE, "chr21_poisson_trace.tsv"), "w").write("\n".join(lines) + "\n")
    print("wrote poisson_logp.tsv, chisq_sf.tsv, chr21_poisson_trace.tsv")


if __name__ == "__main__":
    main()
