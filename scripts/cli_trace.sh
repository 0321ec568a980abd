cd /tmp && python - <<'PY'
import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from breakdancer_b200 import api, synth
pairs = 24_000_000
w = synth.config2(pairs, seed=20260106, chrom_len=5 * pairs)
os.makedirs("/tmp/tr", exist_ok=True)
for bam, cols in synth.split_by_bam(w).items():
    api.write_bam(os.path.join("/tmp/tr", bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=6)
open("/tmp/tr/cfg", "w").write(w.config_text())
PY
cd /tmp/tr
for i in 1 2; do env BDK_DECODE_TRACE=1 $GRAFT_REPO_ROOT/breakdancer_b200/bin/breakdancer_max --stats-json s.json cfg > out.tsv; cat s.json | cut -c1-400; done
