"""Why is the first bdk_push_bam of a process slower than the following ones? One fresh process per variant:
 plain / GPU made busy first (clocks up) / CUDA_MODULE_LOADING=EAGER / both.   usage: coldstart_probe.py [pairs]"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 2 and sys.argv[1] == "--child":
    mode, tmp = sys.argv[2], sys.argv[3]
    os.chdir(tmp)
    from breakdancer_b200 import api
    t0 = time.perf_counter()
    if "busy" in mode:
        import torch
        x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
        for _ in range(60):
            y = x @ x
        torch.cuda.synchronize()
    cfg = api.BamConfig(path="cfg")
    dev = api.BamDevice(cfg)
    ctx = api.Context(dev.bundle(api.Options()))
    t1 = time.perf_counter()
    out = []
    for it in range(3):
        ctx.reset()
        ta = time.perf_counter()
        st = ctx.push_bam(dev)
        out.append({"push_ms": round((time.perf_counter() - ta) * 1e3, 1), "inflate_ms": round(st["inflate_ms"], 1), "stage_ms": round(st["stage_ms"], 1)})
    print(json.dumps({"mode": mode, "setup_s": round(t1 - t0, 3), "calls": out}))
    sys.exit(0)
from breakdancer_b200 import api, synth
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 12_000_000
tmp = tempfile.mkdtemp(prefix="bdk_cold_")
w = synth.config2(pairs, seed=20260106, chrom_len=max(1_000_000, 5 * pairs))
for bam, cols in synth.split_by_bam(w).items():
    api.write_bam(os.path.join(tmp, bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=6)
open(os.path.join(tmp, "cfg"), "w").write(w.config_text())
for mode, env in (("plain", {}), ("busy", {}), ("eager", {"CUDA_MODULE_LOADING": "EAGER"}), ("busy+eager", {"CUDA_MODULE_LOADING": "EAGER"}), ("plain", {})):
    p = subprocess.run([sys.executable, __file__, "--child", mode, tmp], env=dict(os.environ, **env), capture_output=True, text=True)
    print(p.stdout.strip() or p.stderr[-400:], flush=True)
