#!/usr/bin/env python
"""Host BAM decode throughput (BGZF inflate + field extraction into the pinned SoA columns + merge): the bound of the drop-in
executable on real BAM files.  CPU only, no GPU needed.

    python scripts/bam_decode_bench.py [--pairs 2000000] [--threads 0] [--reps 3] [--level 6] [--dir /tmp/bdk_decode_bench]

Writes a synthetic one-library BAM once (re-used on later runs), then opens it `reps` times and prints the best timing as one
JSON line.  BDK_FAST_INFLATE=0 selects the zlib path for the comparison."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from breakdancer_b200 import api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=2_000_000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--dir", default="/tmp/bdk_decode_bench")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3"], help="config3: two bams (tumor / normal), merged")
    a = ap.parse_args()
    a.dir = os.path.join(a.dir, a.workload)
    os.makedirs(a.dir, exist_ok=True)
    w = synth.config2(a.pairs, seed=11, chrom_len=50_000_000) if a.workload == "config2" else synth.config3(a.pairs, seed=11)
    cwd = os.getcwd()
    os.chdir(a.dir)
    try:
        for bam, cols in synth.split_by_bam(w).items():
            tag = "%s.%d.%d" % (bam, a.pairs, a.level)
            if not os.path.exists(tag):
                api.write_bam(bam, [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=a.level)
                open(tag, "w").close()
        size = sum(os.path.getsize(b) for b in synth.split_by_bam(w))
        cfg = api.BamConfig(text=w.config_text())
        best = None
        for _ in range(a.reps):
            t0 = time.perf_counter()
            st = api.BamStream(cfg, threads=a.threads)
            dt = time.perf_counter() - t0
            t = st.timings()
            n = st.n
            st.close()
            if best is None or dt < best[0]:
                best = (dt, t)
        dt, t = best
        print(json.dumps({"records": n, "bam_bytes": size, "open_s": round(dt, 4), "records_per_s": round(n / dt),
                          "inflate_s": round(t["inflate_s"], 4), "extract_s": round(t["extract_s"], 4),
                          "merge_s": round(t["merge_s"], 4), "threads": a.threads or os.cpu_count(),
                          "fast_inflate": os.environ.get("BDK_FAST_INFLATE", "1")}))
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    main()
