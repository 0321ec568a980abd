"""Per-sweep phase times of the persistent K4 kernel (BDK_K4_TRACE=1) on a full-size workload: python scripts/k4_trace.py [2|3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BDK_K4_TRACE"] = "1"
import numpy as np, torch
from breakdancer_b200 import api, synth, synth_torch
config = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else (300_000_000 if config == 3 else 50_000_000)
dev = torch.device("cuda", 0)
if config == 3:
    cols = synth_torch.config3_device(pairs, 20260102, dev)
    bundle, cfg = synth_torch.config3_bundle()
else:
    cols = synth_torch.config2_device(pairs, 20260101, dev, tid=0)
    lib = synth.LibSpec("lib1", "syn_chr1.bam", synth_torch.MEAN, synth_torch.STD, synth_torch.READLEN, ["rg1"])
    wl = synth.Workload({}, [("chr1", synth_torch.CHR1_LEN)], [lib], ["rg1"], ["lib1"], ["syn_chr1.bam"])
    cfg = api.BamConfig(text=wl.config_text())
    bundle = api.ParamBundle(api.Options(), cfg.libs, cfg.nbam, np.zeros(1, np.int32), np.zeros(1, np.int32), cfg.window, 1)
n = cols["pos"].numel()
ctx = api.Context(bundle, 0)
dsoa = synth_torch.soa_of(cols)
for i in range(2):
    ctx.reset(); ctx.push_soa(dsoa, n, device=True); r = ctx.finish_raw()
print({k: round(v["ms"], 3) for k, v in ctx.kernel_times().items()}, "sv", r.n_sv)
