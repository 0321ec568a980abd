"""One-off: component size distribution of the bench workload (config 2, 50 M pairs) from the GPU run's own tables."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from breakdancer_b200 import api, synth, synth_torch
import scipy.sparse as sp, scipy.sparse.csgraph as cg
dev = torch.device("cuda", 0)
cols = synth_torch.config2_device(50_000_000, seed=20260101, device=dev, tid=0)
n = cols["pos"].numel()
lib = synth.LibSpec("lib1", "syn_chr1.bam", synth_torch.MEAN, synth_torch.STD, synth_torch.READLEN, ["rg1"])
wl = synth.Workload({}, [("chr1", synth_torch.CHR1_LEN)], [lib], ["rg1"], ["lib1"], ["syn_chr1.bam"])
cfg = api.BamConfig(text=wl.config_text())
bundle = api.ParamBundle(api.Options(), cfg.libs, cfg.nbam, np.zeros(1, np.int32), np.zeros(1, np.int32), cfg.window, 1)
ctx = api.Context(bundle, 0)
ctx.push_soa(synth_torch.soa_of(cols), n, device=True)
t = ctx.finish()
regions = ctx.regions(); ar, rr = ctx.areads()
print("A", len(ar), "regions", len(regions), "sv", len(t.sv))
order = np.argsort(ar['qid'], kind='stable'); q = ar['qid'][order]
same = np.nonzero(q[1:] == q[:-1])[0]
x, y = order[same], order[same + 1]
rx, ry = rr[x], rr[y]
ok = (rx >= 0) & (ry >= 0)
ex = np.stack([np.minimum(rx[ok], ry[ok]), np.maximum(rx[ok], ry[ok])], 1)
ue, w = np.unique(ex, axis=0, return_counts=True)
print("links", ok.sum(), "edges", len(ue), "strong(>=2)", (w >= 2).sum())
g = sp.coo_matrix((np.ones(len(ue)), (ue[:, 0], ue[:, 1])), shape=(len(regions),) * 2)
nc, lab = cg.connected_components(g, directed=False)
ec = np.bincount(lab[ue[:, 0]], minlength=nc)
print("components with edges", (ec > 0).sum(), "max edges", ec.max(), "top", np.sort(ec)[-10:])
print("edge-count histogram (capped 40):", np.bincount(np.minimum(ec, 40)))
sc = np.bincount(lab[ue[:, 0]], weights=(w >= 2), minlength=nc)
print("strong edges per component: max", sc.max(), "top", np.sort(sc)[-10:])
print("reads per region mean", regions['n_reads'].mean(), "max", regions['n_reads'].max())
print(ctx.kernel_times())
