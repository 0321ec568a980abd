"""Device-resident BAM decode measured on the GPU box: one BAM of the bench workload (config-2 records, deflate level 6)
 (a) in-process: bdk_push_bam (wall, stage times from bdk_bam_stats / the kernel timers) against the host decoder + bdk_push,
 (b) the drop-in executable with the device decode and with the host decoder (BDK_GPU_DECODE=0), whole process, best of 3.
usage: python scripts/bamdev_bench.py [pairs] [level]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from breakdancer_b200 import api, synth  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
tmp = tempfile.mkdtemp(prefix="bdk_bamdev_")
w = synth.config2(pairs, seed=20260106, chrom_len=max(1_000_000, 5 * pairs))
t0 = time.perf_counter()
for bam, cols in synth.split_by_bam(w).items():
    api.write_bam(os.path.join(tmp, bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=level)
    path = os.path.join(tmp, bam)
open(os.path.join(tmp, "cfg"), "w").write(w.config_text())
out = {"pairs": w.n // 2, "bam_bytes": os.path.getsize(path), "level": level, "write_s": round(time.perf_counter() - t0, 2)}
os.chdir(tmp)
cfg = api.BamConfig(path="cfg")

# (a) in-process
dev = api.BamDevice(cfg)
ctx = api.Context(dev.bundle(api.Options()))
runs = []
for it in range(4):
    ctx.reset()
    t0 = time.perf_counter()
    st = ctx.push_bam(dev)
    summ = ctx.summary()
    dt = time.perf_counter() - t0
    kt = ctx.kernel_times()
    runs.append({"push_bam_s": round(dt, 4), "inflate_ms": round(st["inflate_ms"], 2), "chain_ms": round(st["chain_ms"], 2), "extract_ms": round(st["extract_ms"], 2),
                 "stage_ms": round(st["stage_ms"], 2), "wall_ms": round(st["wall_ms"], 2), "k1_ms": round(kt["k1_classify"]["ms"], 3), "windows": st["windows"], "guess_misses": st["guess_misses"], "records": st["records"], "kept": st["kept"],
                 "inflate_GBps_out": round(st["inflated_bytes"] / 1e6 / max(st["inflate_ms"], 1e-6), 2), "pairs_per_s": round(st["kept"] / 2 / dt)})
table = ctx.finish()
out["in_process_device"] = runs
out["sv_calls_device"] = len(table.sv)
n_dev = int(summ.n_records)
ctx.close(); dev.close()

t0 = time.perf_counter()
host = api.BamStream(cfg)
t1 = time.perf_counter()
b = api.ParamBundle.from_stream(api.Options(), cfg, host)
ctx = api.Context(b)
t2 = time.perf_counter()
ctx.push({k: v for k, v in host.cols.items()})
s2 = ctx.summary()
t3 = time.perf_counter()
table2 = ctx.finish()
out["in_process_host"] = {"decode_s": round(t1 - t0, 4), "stages": host.timings(), "push_s": round(t3 - t2, 4), "records": host.n, "sv_calls": len(table2.sv)}
assert host.n == n_dev and len(table2.sv) == len(table.sv), (host.n, n_dev, len(table2.sv), len(table.sv))
ctx.close(); host.close()

# (b) the executable
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
same = open("out1.tsv").read().split("\n", 2)[2] == open("out0.tsv").read().split("\n", 2)[2]
out["cli_outputs_identical"] = same
print(json.dumps(out))
