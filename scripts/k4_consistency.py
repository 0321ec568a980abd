"""Same config-3-shaped job through the warp-only walker, the CTA walker (twice) and with tracing: the SV tables must be identical.
python scripts/k4_consistency.py [pairs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from breakdancer_b200 import api, synth_torch
from tests import util
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
dev = torch.device("cuda", 0)
cols = synth_torch.config3_device(pairs, 20260102, dev)
bundle, cfg = synth_torch.config3_bundle()
n = cols["pos"].numel()
dsoa = synth_torch.soa_of(cols)
ref = None
for name, env in (("warp walker only", {"BDK_K4_CTA_MIN": str(1 << 30)}), ("cta walker", {}), ("cta walker again", {}), ("cta walker, traced", {"BDK_K4_TRACE": "1"}),
                  ("cta for all", {"BDK_K4_CTA_MIN": "0"})):
    for k in ("BDK_K4_CTA_MIN", "BDK_K4_TRACE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = api.Context(bundle, 0)
    for job in range(1 if ref is None else 3):      # the same context again and again (reset + re-run)
        ctx.reset()
        ctx.push_soa(dsoa, n, device=True)
        t = ctx.finish()
        sup = ctx.support()
        print(name, "job", job, len(t.sv), "sv,", ctx.k4_sweeps(), "sweeps, K4 ms", round(ctx.kernel_times()["k4_sv_score"]["ms"], 1), flush=True)
        if ref is None:
            ref = (t, sup)
        else:
            try:
                util.assert_tables_equal(ref[0], t, name)
                assert np.array_equal(ref[1], sup), "support differs"
                print("   identical to the warp walker's table")
            except AssertionError as ex:
                print("   DIFFERS:", ex)
    ctx.close()
