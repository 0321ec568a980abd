#!/usr/bin/env python
"""Turn the ncu artefacts a GPU round leaves in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/summarize_profiles.py TAG [kernel-report.ncu-rep ...]

* gpurun_out/launches_TAG.csv  (ncu --metrics gpu__time_duration.sum --clock-control none of `bench.py`)
    -> profiles/launches_TAG.md : per-kernel launch count, total / mean duration and share of all GPU time,
       split into "bdk" kernels (ours) and everything else (torch workload generator, copies).
* every .ncu-rep given (ncu --set full of one kernel) -> profiles/<name>_TAG.md : the metrics the roofline
  and the next optimisation step are read from (duration, DRAM bytes, throughput, occupancy, stalls).
  For the K1 report it also writes profiles/k1_traffic.json (DRAM bytes per record) which bench.py reports
  as roofline.traffic.
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        e = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"]])
        e[0] += 1
        e[1] += float(r["Metric Value"]) / 1e3
    ours = {k: v for k, v in agg.items() if k.startswith("bdk::")}
    tot_ours = sum(v[1] for v in ours.values())
    tot_all = sum(v[1] for v in agg.values())
    out = [f"# ncu launch list `{tag}` (`ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --no-cpu`)",
           "", f"{len(rows)} launches captured; bdk kernels {tot_ours:.0f} us of {tot_all:.0f} us total GPU time "
           "(the rest is the torch workload generator and memsets/copies, outside the timed region).",
           "Per-launch times under ncu are serialised and cold-cache: use the SHARE, not the absolute.", "",
           "| kernel | launches | total us | mean us | share of bdk time | grid | block |", "|---|---:|---:|---:|---:|---|---|"]
    for k, v in sorted(ours.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / max(tot_ours, 1e-9):.1f} % | {v[2]} | {v[3]} |")
    out += ["", "Largest non-bdk kernels (workload generation, not timed):", ""]
    for k, v in sorted(((k, v) for k, v in agg.items() if k not in ours), key=lambda x: -x[1][1])[:6]:
        out.append(f"* `{k[:90]}` x{v[0]}: {v[1]:.0f} us")
    open(os.path.join(ROOT, "profiles", f"launches_{tag}.md"), "w").write("\n".join(out) + "\n")


WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (elapsed)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__shared_mem_per_block_static", "static smem / block"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction"),
]


def report(tag, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print("cannot read", rep)
        return
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    kname = m.get("Kernel Name", ("?", ""))[0]
    out = [f"# ncu --set full: `{kname[:100]}` ({tag})", "", f"source report: `{os.path.basename(rep)}` (gpurun_out/, not tracked)", "",
           "| metric | value |", "|---|---|"]
    for key, label in WANT:
        if key in m:
            out.append(f"| {label} (`{key}`) | {m[key][0]} {m[key][1]} |")
    stalls = sorted(((float(v[0]), k) for k, v in m.items() if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", k) and v[0]), reverse=True)
    out += ["", "Top warp stall reasons (warps stalled per issue-active cycle):", ""]
    for v, k in stalls[:6]:
        out.append(f"* {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
    base = re.sub(r"\.ncu-rep$", "", os.path.basename(rep))
    open(os.path.join(ROOT, "profiles", f"{base}.md"), "w").write("\n".join(out) + "\n")

    def num(key):
        v, u = m.get(key, ("0", ""))
        x = float(v)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        return x * scale.get(u, 1)
    if "k1_classify" in kname:
        # DRAM bytes of ONE launch of the classify kernel, with the workload it was captured on (records of the launch from the
        # bench line of the same command) and a digest of the kernel's sources: bench.py reports it as roofline.traffic only for
        # the same workload and the same sources (profiles/k1_traffic.json = configs[1], k1_traffic_config3.json = configs[2]).
        total = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        c3 = "config3" in os.path.basename(rep)
        n = None
        try:
            n = json.load(open(os.path.join(ROOT, "gpurun_out", f"bench_c3_{tag}.json" if c3 else f"bench_{tag}.json")))["config"]["records_per_gpu"]
        except Exception:
            pass
        if n:
            sys.path.insert(0, ROOT)
            from bench import k1_source_digest
            json.dump({"tag": tag, "dram_bytes": total, "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_written": num("dram__bytes_write.sum"),
                       "records": n, "dram_bytes_per_record": total / n, "kernel": kname[:80], "k1_source_digest": k1_source_digest(),
                       "source": os.path.basename(rep)}, open(os.path.join(ROOT, "profiles", "k1_traffic_config3.json" if c3 else "k1_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launches(tag)
    for rep in sys.argv[2:]:
        report(tag, rep)
