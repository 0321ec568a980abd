"""Four bams (one per lane / library) through the drop-in executable: device decode + the merge order from the priority queue on
the host + device gather, against the host decoder (BDK_GPU_DECODE=0), whole process, best of 3; and in-process bdk_push_bams.
usage: bamdev_benchn.py [pairs]"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from breakdancer_b200 import api, synth
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 12_000_000
tmp = tempfile.mkdtemp(prefix="bdk_bamn_")
genome = [("chr1", 120_000_000), ("chr2", 90_000_000), ("chr3", 60_000_000)]
libs = [synth.LibSpec("lane1", "lane1.bam", 315, 44, 75, ["l1a", "l1b"]), synth.LibSpec("lane2", "lane2.bam", 312, 43, 75, ["l2"]),
        synth.LibSpec("lane3", "lane3.bam", 467, 32, 75, ["l3"], tumor=True), synth.LibSpec("lane4", "lane4.bam", 476, 29, 100, ["l4"], tumor=True)]
w = synth.generate(genome, libs, pairs, seed=78, anomaly_frac=0.02, somatic_frac=0.3)
sizes = {}
for bam, cols in synth.split_by_bam(w).items():
    api.write_bam(os.path.join(tmp, bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=6)
    sizes[bam] = os.path.getsize(os.path.join(tmp, bam))
open(os.path.join(tmp, "cfg"), "w").write(w.config_text())
out = {"pairs": w.n // 2, "bam_bytes": sizes}
os.chdir(tmp)
cfg = api.BamConfig(path="cfg")
devs = []
for p in cfg.bam_files:
    devs.append(api.BamDevice(cfg, path=p, after=devs[-1] if devs else None))
b = api.ParamBundle(api.Options(), cfg.libs, cfg.nbam, np.concatenate([d.rg_lib for d in devs]), np.concatenate([d.rg_bam for d in devs]), cfg.window, len(devs[0].tid_names))
ctx = api.Context(b)
runs = []
for it in range(3):
    ctx.reset()
    t0 = time.perf_counter()
    st = ctx.push_bams_n(devs)
    ctx.summary()
    dt = time.perf_counter() - t0
    runs.append({"push_bams_s": round(dt, 4), "decode_ms": [round(s["wall_ms"], 1) for s in st], "pairs_per_s": round(sum(s["kept"] for s in st) / 2 / dt)})
out["in_process"] = runs
nsv = len(ctx.finish().sv)
ctx.close()
for d in devs:
    d.close()
cli = os.path.join(ROOT, "breakdancer_b200", "bin", "breakdancer_max")
for mode in ("1", "0"):
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        rc = subprocess.run([cli, "--stats-json", "stats.json", "cfg"], env=dict(os.environ, BDK_GPU_DECODE=mode), stdout=open("out%s.tsv" % mode, "w"), stderr=subprocess.PIPE, text=True)
        dt = time.perf_counter() - t0
        assert rc.returncode == 0, rc.stderr[-400:]
        if best is None or dt < best[0]:
            best = (dt, json.load(open("stats.json")))
    out["cli_device" if mode == "1" else "cli_host"] = {"wall_s": round(best[0], 3), "pairs_per_s": round((w.n // 2) / best[0]), "stats": best[1]}
out["cli_outputs_identical"] = open("out1.tsv").read().split("\n", 2)[2] == open("out0.tsv").read().split("\n", 2)[2]
out["sv_calls"] = nsv
print(json.dumps(out))
