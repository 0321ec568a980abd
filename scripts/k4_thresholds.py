"""K4 time of a config-3-shaped job for several settings of the walker thresholds (env BDK_K4_CTA_MIN / BDK_K4_BIG).
python scripts/k4_thresholds.py [pairs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from breakdancer_b200 import api, synth_torch
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
dev = torch.device("cuda", 0)
cols = synth_torch.config3_device(pairs, 20260102, dev)
bundle, cfg = synth_torch.config3_bundle()
n = cols["pos"].numel()
dsoa = synth_torch.soa_of(cols)
for env in ({}, {"BDK_K4_DEFER_FIRST": "0"}, {"BDK_K4_DEFER_FIRST": "0", "BDK_K4_BIG": "1024"}) if len(sys.argv) > 2 else ({}, {"BDK_K4_CTA_MIN": "128"}, {"BDK_K4_BIG": "1024"}, {"BDK_K4_CTA_MIN": "128", "BDK_K4_BIG": "1024"}, {"BDK_K4_CTA_MIN": "2048"}, {"BDK_K4_BIG": "16384"},
            {"BDK_K4_CTA_MIN": "128", "BDK_K4_BIG": "512"}):
    for k in ("BDK_K4_CTA_MIN", "BDK_K4_BIG", "BDK_K4_DEFER_FIRST"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = api.Context(bundle, 0)
    for job in range(2):
        ctx.reset(); ctx.push_soa(dsoa, n, device=True); r = ctx.finish_raw()
    print(env or "default", "sv", r.n_sv, "sweeps", ctx.k4_sweeps(), "K4 ms", round(ctx.kernel_times()["k4_sv_score"]["ms"], 1), flush=True)
    ctx.close()
