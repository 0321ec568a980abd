"""Sweep of the device-decode pipeline parameters on one BAM (in-process bdk_push_bam, best of 3 per setting)."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from breakdancer_b200 import api, synth
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
tmp = tempfile.mkdtemp(prefix="bdk_sweep_")
w = synth.config2(pairs, seed=20260106, chrom_len=max(1_000_000, 5 * pairs))
for bam, cols in synth.split_by_bam(w).items():
    api.write_bam(os.path.join(tmp, bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=6)
open(os.path.join(tmp, "cfg"), "w").write(w.config_text())
os.chdir(tmp)
cfg = api.BamConfig(path="cfg")
settings = [dict(BDK_BAMDEV_NOPRIO="1"), dict(), dict(BDK_BAMDEV_WSLOTS="4"), dict(BDK_BAMDEV_WSLOTS="6"),
            dict(BDK_BAMDEV_CHUNK_KB="32768", BDK_BAMDEV_STREAMS="4", BDK_BAMDEV_WSLOTS="4"),
            dict(BDK_BAMDEV_CHUNK_KB="32768", BDK_BAMDEV_STREAMS="8", BDK_BAMDEV_WSLOTS="4", BDK_BAMDEV_WINDOW_KB="131072"),
            dict(BDK_BAMDEV_CHUNK_KB="65536", BDK_BAMDEV_STREAMS="4", BDK_BAMDEV_WSLOTS="4", BDK_BAMDEV_WINDOW_KB="131072"),
            dict(BDK_BAMDEV_CHUNK_KB="8192", BDK_BAMDEV_STREAMS="16", BDK_BAMDEV_WSLOTS="4"),
            dict(BDK_BAMDEV_WSLOTS="4", BDK_BAMDEV_WINDOW_KB="32768")]
if len(sys.argv) > 2:
    settings = json.loads(sys.argv[2])
KNOBS = ("BDK_BAMDEV_STREAMS", "BDK_BAMDEV_CHUNK_KB", "BDK_BAMDEV_WINDOW_KB", "BDK_BAMDEV_WSLOTS", "BDK_BAMDEV_NOPRIO", "BDK_BAMDEV_INFLATE_ONLY", "BDK_BAMDEV_COPY_THREADS")
for s in settings:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(s)
    dev = api.BamDevice(cfg)
    ctx = api.Context(dev.bundle(api.Options()))
    best = None
    for it in range(4):
        ctx.reset()
        t0 = time.perf_counter()
        st = ctx.push_bam(dev)
        dt = time.perf_counter() - t0
        if it and (best is None or dt < best[0]):
            best = (dt, st)
    dt, st = best
    print(json.dumps({"setting": s, "push_bam_ms": round(dt * 1e3, 1), "inflate_ms": round(st["inflate_ms"], 1), "chain_ms": round(st["chain_ms"], 1),
                      "extract_ms": round(st["extract_ms"], 1), "stage_ms": round(st["stage_ms"], 1), "windows": st["windows"], "GBps": round(st["inflated_bytes"] / 1e6 / st["inflate_ms"], 1)}), flush=True)
    ctx.close(); dev.close()
