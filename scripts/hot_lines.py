#!/usr/bin/env python
"""Join ncu's per-instruction counters (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line info of the
same cubin, and print the hottest source lines of a kernel.   usage: hot_lines.py REPORT.ncu-rep KERNEL_SUBSTRING [N]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "build", "bdk_core.o")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l][0]
end = next(i for i in range(start + 1, len(dis)) if dis[i].startswith(".text.") or dis[i].startswith(".section"))
cur, locs = None, []
for l in dis[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        locs.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
print(f"{len(locs)} instructions in the cubin, {len(data)} in the report")
a, s = collections.Counter(), collections.Counter()
for loc, r in zip(locs, data):
    a[loc] += int(r[ia]); s[loc] += int(r[isamp])
ta, ts = sum(a.values()), sum(s.values())
for loc, v in a.most_common(topn):
    print(f"{loc[0]}:{loc[1]:<5d} inst {100 * v / ta:5.1f}%  samples {100 * s[loc] / ts:5.1f}%")
