"""In-process bdk_push_bam rate of several builds of libbdk.so (BDK_LIB=path) on one BAM; one subprocess per build.
usage: bamdev_variants.py pairs lib1.so lib2.so ..."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from breakdancer_b200 import api
if sys.argv[1] == "--child":
    api.LIB_PATH = sys.argv[2]
    os.chdir(sys.argv[3])
    cfg = api.BamConfig(path="cfg")
    dev = api.BamDevice(cfg)
    ctx = api.Context(dev.bundle(api.Options()))
    best = None
    for it in range(5):
        ctx.reset()
        t0 = time.perf_counter()
        st = ctx.push_bam(dev)
        dt = time.perf_counter() - t0
        if it and (best is None or dt < best[0]):
            best = (dt, st)
    dt, st = best
    print(json.dumps({"lib": os.path.basename(sys.argv[2]), "push_bam_ms": round(dt * 1e3, 1), "inflate_ms": round(st["inflate_ms"], 1), "GBps": round(st["inflated_bytes"] / 1e6 / st["inflate_ms"], 1),
                      "pairs_per_s": round(st["kept"] / 2 / dt), "kept": st["kept"]}))
    sys.exit(0)
from breakdancer_b200 import synth
pairs = int(sys.argv[1])
tmp = tempfile.mkdtemp(prefix="bdk_var_")
w = synth.config2(pairs, seed=20260106, chrom_len=max(1_000_000, 5 * pairs))
for bam, cols in synth.split_by_bam(w).items():
    api.write_bam(os.path.join(tmp, bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=6)
open(os.path.join(tmp, "cfg"), "w").write(w.config_text())
for lib in sys.argv[2:]:
    p = subprocess.run([sys.executable, __file__, "--child", os.path.abspath(lib), tmp], capture_output=True, text=True)
    print(p.stdout.strip() or p.stderr[-500:], flush=True)
    if os.environ.get("BDK_DECODE_TRACE"):
        print("\n".join(p.stderr.strip().split("\n")[-14:]), flush=True)
