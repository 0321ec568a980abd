#!/bin/bash
# Goldens for the C++ bam2cfg: the UNMODIFIED reference script (perl/bam2cfg.pl) run on the bundled chr21 BAMs and on a
# synthetic multi-library BAM, with the stand-in CPAN modules of perl_stubs/ and the samtools built by oracle/build_ref.sh.
# Lines are sorted (the script prints in Perl hash order). Needs /root/reference; the outputs are committed.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../../.." && pwd)
export PATH="$ROOT/oracle/_ref:$PATH"
B2C="perl -I $HERE/perl_stubs -I /root/reference/perl /root/reference/perl/bam2cfg.pl"
cd "$ROOT/tests/golden/chr21"
$B2C -g NA19238_chr21_del_inv.bam NA19240_chr21_del_inv.bam 2>/dev/null | sort > "$HERE/chr21_g.cfg"
$B2C -q 10 -n 500 -c 3 NA19238_chr21_del_inv.bam 2>/dev/null | sort > "$HERE/chr21_q10_n500_c3.cfg"
$B2C -m -s 100 -v 0.5 NA19240_chr21_del_inv.bam 2>/dev/null | sort > "$HERE/chr21_m.cfg"
python - <<PY
import sys
sys.path.insert(0, "$ROOT")
from breakdancer_b200 import api, synth
from tests import util
w = synth.generate(util.GENOME3, util.LIBS4, 60000, seed=77, anomaly_frac=0.05)
for bam, cols in synth.split_by_bam(w).items():
    api.write_bam("$HERE/syn_" + bam, [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=6)
PY
cd "$HERE"
$B2C -g -n 3000 syn_normal.bam syn_tumor.bam 2>/dev/null | sort > "$HERE/syn_g_n3000.cfg"
wc -l "$HERE"/*.cfg
