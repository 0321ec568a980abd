import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from breakdancer_b200 import api, synth_torch
cols = synth_torch.config3_device(400_000, 13, torch.device("cuda", 0))
bundle, cfg = synth_torch.config3_bundle()
n = cols["pos"].numel()
ctx = api.Context(bundle, 0)
ctx.push_soa(synth_torch.soa_of(cols), n, device=True)
t = ctx.finish()
print("sv", len(t.sv), "sweeps", ctx.k4_sweeps())
