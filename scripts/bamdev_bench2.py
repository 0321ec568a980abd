"""Two bams (tumor / normal, four libraries) through the drop-in executable: device decode + device merge against the host decoder
(BDK_GPU_DECODE=0), whole process, best of 3; and in-process bdk_push_bams.   usage: bamdev_bench2.py [pairs]"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from breakdancer_b200 import api, synth
from tests import util
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 12_000_000
tmp = tempfile.mkdtemp(prefix="bdk_bam2_")
genome = [("chr1", 120_000_000), ("chr2", 90_000_000), ("chr3", 60_000_000)]
w = synth.generate(genome, util.LIBS4, pairs, seed=77, anomaly_frac=0.02, somatic_frac=0.3)
t0 = time.perf_counter()
sizes = {}
for bam, cols in synth.split_by_bam(w).items():
    api.write_bam(os.path.join(tmp, bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=6)
    sizes[bam] = os.path.getsize(os.path.join(tmp, bam))
open(os.path.join(tmp, "cfg"), "w").write(w.config_text())
out = {"pairs": w.n // 2, "bam_bytes": sizes, "write_s": round(time.perf_counter() - t0, 1)}
os.chdir(tmp)
cfg = api.BamConfig(path="cfg")
d0 = api.BamDevice(cfg, path=cfg.bam_files[0]); d1 = api.BamDevice(cfg, path=cfg.bam_files[1], after=d0)
b = api.ParamBundle(api.Options(), cfg.libs, cfg.nbam, np.concatenate([d0.rg_lib, d1.rg_lib]), np.concatenate([d0.rg_bam, d1.rg_bam]), cfg.window, len(d0.tid_names))
ctx = api.Context(b)
runs = []
for it in range(3):
    ctx.reset()
    t0 = time.perf_counter()
    s0, s1 = ctx.push_bams(d0, d1)
    ctx.summary()
    dt = time.perf_counter() - t0
    runs.append({"push_bams_s": round(dt, 4), "decode_ms": [round(s0["wall_ms"], 1), round(s1["wall_ms"], 1)], "merge_parts": s0["merge_parts"], "longest_part": s0["merge_longest_part"],
                 "kernel_ms": {k: round(v["ms"], 2) for k, v in ctx.kernel_times().items() if v["ms"] > 0}, "pairs_per_s": round((s0["kept"] + s1["kept"]) / 2 / dt)})
out["in_process"] = runs
nsv = len(ctx.finish().sv)
ctx.close(); d0.close(); d1.close()
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
