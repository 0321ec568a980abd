#!/usr/bin/env python
"""bench.py -- read-pairs/sec to SV calls (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3] [--pairs P] [--impl ours|reference]

A "step" is one whole job of the hot path over one synthetic batch: BASELINE.json configs[1]
(single-library 30x chr1, DEL-only, 50 M read pairs = 100 M position-sorted records per GPU):
classify -> regions -> link graph -> scored SV table.  `value` is measured with the record columns
already resident in HBM (bdk_push_device); `e2e` is the same job through the C ABI with HOST (pinned)
columns, so host->device copies and the device->host read of the SV table are inside the timed
region.  `roofline` is the classify kernel (25 algorithmic bytes per record) against the measured HBM
peak; `cpu_baseline` is the unmodified reference binary on the host cores on a bounded sample;
`cpu_baseline_decode_free` is the reference's algorithm alone on records decoded into memory beforehand (the comparator of
`e2e`, which also starts from decoded records); `file_e2e` is the drop-in executable from a BAM file to the SV table (whole
process, wall clock: the comparator of `cpu_baseline`); `bam_decode` (N = 1) is the host's BAM decode of a bounded sample on
all host cores -- what bounds the drop-in executable on real files. `comparisons` puts the like-for-like pairs side by side.
For N > 1 each rank runs its own chromosome-shaped shard (the path shards by chromosome with
no data-path collective: weak scaling); torch.distributed/NCCL is used only for the barrier and the
max-over-ranks of the device time.  One JSON line is printed by rank 0.  For N > 1 the line also carries
`one_job_all_gpus`: ONE job spread over all ranks with whole-genome semantics (and a second one with -t),
i.e. the NCCL exchanges of csrc/comm.cuh inside the timed region, each with a `parity` block (all ranks return the same
table, and it equals a single-GPU run of the concatenated stream), plus BASELINE configs[3] (24 GRCh38 chromosomes as -o
shards, LPT-packed) and configs[4] (-t, 1 B pairs, 24 chromosomes over the N GPUs).
--config 3 runs BASELINE configs[2] instead (300 M pairs, 4 libraries in 2 BAMs, all five SV types).

--impl reference times the unmodified reference executable (oracle/_ref/breakdancer-max; the oracle
port if that is missing) on the host cores: every step runs one single-threaded process per core, each
on its own bounded sample of the same workload written as BAM (the reference's documented way to use
several cores is one process per chromosome).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "read_pairs_per_sec_to_sv_calls"
UNIT = "read-pairs/s"
BYTES_PER_RECORD = 25          # pos,mpos,tid,mtid,isize int32 + flag u16 + mapq u8 + rgid u16
HOST_BYTES_PER_RECORD = 37     # + qlen int32 + qid u64 side columns copied by bdk_push


def k1_source_digest():
    """Digest of the sources the classify kernel is built from: a DRAM-traffic capture is reported only for the kernel it was taken on."""
    import hashlib
    h = hashlib.sha256()
    for f in ("k1_classify.cuh", "bdk_logic.h", "common.cuh"):
        h.update(open(os.path.join(ROOT, "breakdancer_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def k1_traffic(n_records, config):
    """(DRAM bytes of one K1 launch from the tracked ncu --set full capture, where it comes from) -- None unless the capture was
    taken on this workload (same records per launch) with the kernel sources as they are now."""
    name = "k1_traffic_config3.json" if config == 3 else "k1_traffic.json"
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None, f"no capture (profiles/{name})"
    if int(t.get("records", -1)) != int(n_records):
        return None, f"profiles/{name} ({t.get('tag')}) was captured on {t.get('records')} records, this launch has {n_records}"
    if t.get("k1_source_digest") != k1_source_digest():
        return None, f"profiles/{name} ({t.get('tag')}) was captured on other kernel sources; re-run scripts/gpu_round.sh"
    return float(t["dram_bytes"]), f"profiles/{name}: ncu --set full {t.get('source')} ({t.get('tag')}), dram__bytes_read.sum + dram__bytes_write.sum of one launch"


def workload_name(pairs, config=2):
    if config == 3:
        return f"synthetic 4-library tumor/normal chr1-3, all 5 SV types (-c 3 -q 35), {pairs / 1e6:g}M read pairs per GPU (BASELINE configs[2])"
    return f"synthetic single-library 30x chr1-shaped, DEL-only, {pairs / 1e6:g}M read pairs per GPU (BASELINE configs[1])"


# ---------------------------------------------------------------------------------------------------
# CPU reference arm
# ---------------------------------------------------------------------------------------------------
def _write_sample_bams(tmp, nproc, pairs_each, seed0):
    """nproc independent BAMs (+ configs) of `pairs_each` read pairs each, config-2 distribution."""
    from breakdancer_b200 import api, synth
    jobs = []
    for i in range(nproc):
        w = synth.config2(pairs_each, seed=seed0 + i)
        d = os.path.join(tmp, f"s{i}")
        os.makedirs(d, exist_ok=True)
        for bam, cols in synth.split_by_bam(w).items():
            api.write_bam(os.path.join(d, bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=1)
        open(os.path.join(d, "cfg"), "w").write(w.config_text())
        jobs.append((d, w.n // 2))
    return jobs


def _run_reference_once(jobs):
    """One process per job, all concurrently; returns (wall seconds, total pairs, kind)."""
    from oracle import oracle
    if oracle.have_reference():
        t0 = time.perf_counter()
        procs = [subprocess.Popen([oracle.REF_BIN, "cfg"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for d, _ in jobs]
        rcs = [p.wait() for p in procs]
        dt = time.perf_counter() - t0
        if any(rcs):
            raise RuntimeError(f"reference exited with {rcs}")
        return dt, sum(n for _, n in jobs), "reference"
    # oracle port (single thread, decode included through our host decoder)
    from breakdancer_b200 import api
    t0 = time.perf_counter()
    for d, _ in jobs:
        cwd = os.getcwd()
        os.chdir(d)
        try:
            cfg = api.BamConfig(path="cfg")
            st = api.BamStream(cfg, threads=1)
            b = api.ParamBundle.from_stream(api.Options(), cfg, st)
            oracle.run(b, {k: v.copy() for k, v in st.cols.items()})
        finally:
            os.chdir(cwd)
    return time.perf_counter() - t0, sum(n for _, n in jobs), "port"


def _run_reference_decode_free(jobs):
    """The unmodified reference algorithm on records decoded into memory beforehand (oracle/_ref/breakdancer-max-nodecode:
    its BAM readers are replaced by in-memory ones), one process per job, all concurrently. Returns the slowest process's
    algorithm time (wall minus decoding) and the pairs."""
    import re
    from oracle import oracle
    procs = [subprocess.Popen([oracle.REF_NODECODE, "cfg"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True) for d, _ in jobs]
    algo = []
    for p in procs:
        err = p.communicate()[1]
        m = re.search(r"nodecode wall_s=([0-9.]+) load_s=([0-9.]+) algorithm_s=([0-9.]+)", err)
        if p.returncode != 0 or not m:
            raise RuntimeError(f"decode-free reference failed rc={p.returncode}: {err[-300:]}")
        algo.append(float(m.group(3)))
    return max(algo), sum(n for _, n in jobs)


def cpu_baseline(pairs_total, nproc, seed0=20260101):
    """(cpu_baseline, cpu_baseline_decode_free): the reference executable from BAM files, and the reference's algorithm alone
    (records already in memory: the comparator of a job that starts from decoded records), on the same samples."""
    from oracle import oracle
    tmp = tempfile.mkdtemp(prefix="bdk_cpu_", dir=os.environ.get("TMPDIR", "/tmp"))
    try:
        jobs = _write_sample_bams(tmp, nproc, max(1000, pairs_total // nproc), seed0)
        dt, pairs, kind = _run_reference_once(jobs)
        base = {"value": pairs / dt, "unit": UNIT, "cores": nproc if kind == "reference" else 1, "kind": kind,
                "sample": f"{nproc} x {pairs // nproc} read pairs of the same workload as BAM, one single-threaded process each, {dt:.1f} s wall"}
        free = None
        if os.access(oracle.REF_NODECODE, os.X_OK):
            try:
                adt, apairs = _run_reference_decode_free(jobs)
                free = {"value": apairs / adt, "unit": UNIT, "cores": nproc, "kind": "reference",
                        "sample": f"the same {nproc} samples, BAM decoded into memory before the clock starts (the reference's readers replaced by "
                                  f"in-memory ones, both of its passes timed), slowest process {adt:.2f} s"}
            except Exception as ex:
                free = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)[:200]}
        return base, free
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def file_e2e_sample(pairs, level=6):
    """From a BAM FILE to the SV table with the drop-in executable (breakdancer_b200/bin/breakdancer_max), wall clock of the
    whole process (start, CUDA context, decode, classification, SV calls, output), on one BAM of the bench workload:
    `value` is the default path (the file is decoded ON THE GPU: bdk_push_bam, only compressed bytes cross PCIe);
    `host_decoder` the same executable with BDK_GPU_DECODE=0 (BGZF inflate and record parsing on all host cores, columns pushed
    to the GPU); `in_process` the device path without the process around it (bdk_push_bam + bdk_finish on an open context, best
    of 3: what a long-running caller or a file much larger than the start-up cost sees). Stage times from --stats-json."""
    from breakdancer_b200 import api, synth
    cli = os.path.join(ROOT, "breakdancer_b200", "bin", "breakdancer_max")
    tmp = tempfile.mkdtemp(prefix="bdk_file_", dir=os.environ.get("TMPDIR", "/tmp"))
    cwd = os.getcwd()
    try:
        w = synth.config2(pairs, seed=20260106, chrom_len=max(1_000_000, 5 * pairs))
        size = 0
        for bam, cols in synth.split_by_bam(w).items():
            api.write_bam(os.path.join(tmp, bam), [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=level)
            size += os.path.getsize(os.path.join(tmp, bam))
        open(os.path.join(tmp, "cfg"), "w").write(w.config_text())
        npairs = w.n // 2
        del w

        def run_cli(mode):
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                rc = subprocess.run([cli, "--stats-json", "stats.json", "cfg"], cwd=tmp, env=dict(os.environ, BDK_GPU_DECODE=mode),
                                    stdout=open(os.path.join(tmp, f"out{mode}.tsv"), "w"), stderr=subprocess.PIPE, text=True)
                dt = time.perf_counter() - t0
                if rc.returncode != 0:
                    raise RuntimeError(f"breakdancer_max failed: {rc.stderr[-300:]}")
                if best is None or dt < best[0]:
                    best = (dt, json.load(open(os.path.join(tmp, "stats.json"))))
            dt, st = best
            return {"value": npairs / dt, "unit": UNIT, "wall_s": round(dt, 3), "device_decode": st.get("device_decode"),
                    "stages_s": {k: round(v, 4) for k, v in st.items() if k.endswith("_s") and k != "read_pairs_per_s"},
                    "device_ms": {k: round(v, 2) for k, v in st.items() if k.startswith("device_") and k.endswith("_ms")},
                    "h2d_bytes": st.get("h2d_bytes"), "sv_calls": st.get("sv_calls")}
        dev, host = run_cli("1"), run_cli("0")
        same = open(os.path.join(tmp, "out1.tsv")).read().split("\n", 2)[2] == open(os.path.join(tmp, "out0.tsv")).read().split("\n", 2)[2]
        # the device path inside a running process
        os.chdir(tmp)
        cfg = api.BamConfig(path="cfg")
        bd = api.BamDevice(cfg)
        ctx = api.Context(bd.bundle(api.Options()))
        best = None
        for _ in range(4):
            ctx.reset()
            t0 = time.perf_counter()
            st = ctx.push_bam(bd)
            table = ctx.finish()
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, st, len(table.sv))
        ctx.close(); bd.close()
        dt, st, nsv = best
        inproc = {"value": npairs / dt, "unit": UNIT, "s": round(dt, 4), "sv_calls": nsv, "windows": st["windows"],
                  "inflate_ms": round(st["inflate_ms"], 2), "inflate_GBps_of_output": round(st["inflated_bytes"] / 1e6 / max(st["inflate_ms"], 1e-6), 2),
                  "record_chain_ms": round(st["chain_ms"], 2), "extract_ms": round(st["extract_ms"], 2), "host_staging_ms": round(st["stage_ms"], 2),
                  "h2d_bytes": st["h2d_bytes"], "inflated_bytes": st["inflated_bytes"]}
        out = dict(dev)
        out.update({"cores": os.cpu_count() or 1, "bam_bytes": size, "host_decoder": host, "in_process": inproc, "same_output_both_decoders": same,
                    "sample": f"one BAM of {npairs} read pairs of the same workload (deflate level {level}); whole process (start, CUDA context, decode, GPU, "
                              "output), best of 3, for `value` (file decoded on the GPU) and `host_decoder` (BDK_GPU_DECODE=0); `in_process`: "
                              "bdk_push_bam + bdk_finish on an open context, best of 4"})
        return out
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


def bam_decode_sample(pairs, level=6):
    """What bounds the drop-in executable on real input: the host's BAM decode (BGZF inflate, record boundaries, field extraction
    into the columns the GPU reads), all host cores, on a bounded sample of the bench workload written as a BAM file. No GPU work
    in here; the device side of a job of this size is microseconds. Reported next to e2e, not part of it."""
    from breakdancer_b200 import api, synth
    tmp = tempfile.mkdtemp(prefix="bdk_bam_", dir=os.environ.get("TMPDIR", "/tmp"))
    cwd = os.getcwd()
    try:
        w = synth.config2(pairs, seed=20260105, chrom_len=max(1_000_000, 5 * pairs))
        os.chdir(tmp)
        size = 0
        for bam, cols in synth.split_by_bam(w).items():
            api.write_bam(bam, [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=level)
            size += os.path.getsize(bam)
        cfg = api.BamConfig(text=w.config_text())
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            st = api.BamStream(cfg)
            dt = time.perf_counter() - t0
            stages, n = st.timings(), st.n
            st.close()
            if best is None or dt < best[0]:
                best = (dt, stages, n)
        dt, stages, n = best
        return {"value": n / 2 / dt, "unit": UNIT, "records_per_s": n / dt, "cores": os.cpu_count() or 1, "bam_bytes": size,
                "stages_s": {k: round(v, 4) for k, v in stages.items()}, "open_s": round(dt, 4),
                "sample": f"{n // 2} read pairs of the same workload as one BAM (deflate level {level}), best of 3 decodes, host only"}
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nproc = max(1, min(cores, 32))
    pairs_each = args.ref_pairs
    tmp = tempfile.mkdtemp(prefix="bdk_ref_", dir=os.environ.get("TMPDIR", "/tmp"))
    try:
        jobs = _write_sample_bams(tmp, nproc, pairs_each, 20260101)
        for _ in range(args.warmup):
            _run_reference_once(jobs)
        t = []
        kind = "reference"
        pairs = 0
        for _ in range(args.steps):
            dt, pairs, kind = _run_reference_once(jobs)
            t.append(dt)
        ms = 1e3 * sum(t) / len(t)
        value = pairs / (ms / 1e3)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int32", "data": "synthetic",
                "config": {"workload": workload_name(args.pairs), "step_sample": f"{nproc} x {pairs_each} read pairs as BAM, one process per host core"},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc if kind == "reference" else 1, "kind": kind,
                                 "sample": f"{nproc} BAMs x {pairs_each} read pairs per step"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every ~2 ms from a thread (the timed
    region of the default run is tens of milliseconds, too short for `nvidia-smi -lms`); nvidia-smi is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device):
        self.samples = []          # (t, sm_mhz, reasons bitmask)
        self.max_mhz = None
        self.stop_flag = False
        self.nvml = None
        self.proc = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        self.bits = bits
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                r = int(get_reasons(self.h))
                self.samples.append((time.perf_counter(), mhz, r))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for l in self.proc.stdout:
            f = [x.strip() for x in l.strip().split(",")]
            if len(f) < 8:
                continue
            try:
                r = sum(1 << i for i, v in enumerate(f[4:8]) if v.lower().startswith("active"))
                self.samples.append((time.perf_counter(), float(f[1]), r))
                self.max_mhz = float(f[2])
            except ValueError:
                continue

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
        elif self.nvml:
            self.t.join(timeout=1.0)
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        inside = [(mhz, r) for ts, mhz, r in self.samples if t0 <= ts <= t1]
        sm = sorted(m for m, _ in inside)
        reasons = set()
        for _, r in inside:
            if self.nvml:
                reasons |= {name for name, bit in self.bits.items() if r & bit}
            else:
                reasons |= {name for i, name in enumerate(self.NAMES) if r & (1 << i)}
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self.nvml else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from breakdancer_b200 import api, synth, synth_torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the bdk hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    pairs = args.pairs
    if args.config == 3:
        cols = synth_torch.config3_device(pairs, seed=20260102 + rank, device=dev)
        bundle, cfg = synth_torch.config3_bundle()
    else:
        cols = synth_torch.config2_device(pairs, seed=20260101 + rank, device=dev, tid=0)
        lib = synth.LibSpec("lib1", "syn_chr1.bam", synth_torch.MEAN, synth_torch.STD, synth_torch.READLEN, ["rg1"])
        wl = synth.Workload({}, [("chr1", synth_torch.CHR1_LEN)], [lib], ["rg1"], ["lib1"], ["syn_chr1.bam"])
        cfg = api.BamConfig(text=wl.config_text())
        bundle = api.ParamBundle(api.Options(), cfg.libs, cfg.nbam, np.zeros(1, np.int32), np.zeros(1, np.int32), cfg.window, 1)
    n = cols["pos"].numel()
    npairs = n // 2
    ctx = api.Context(bundle, local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    dsoa = synth_torch.soa_of(cols)

    ktimes = {}
    launches = [0]

    def add_times():
        for k, v in ctx.kernel_times().items():
            e = ktimes.setdefault(k, {"ms": 0.0, "launches": 0})
            e["ms"] += v["ms"]; e["launches"] += v["launches"]
        launches[0] += ctx.kernel_launches()

    host_ms = {"reset": 0.0, "push": 0.0, "finish": 0.0}

    def step_device():
        t0 = time.perf_counter()
        ctx.reset()
        t1 = time.perf_counter()
        ctx.push_soa(dsoa, n, device=True)
        t2 = time.perf_counter()
        r = ctx.finish_raw()
        t3 = time.perf_counter()
        host_ms["reset"] += 1e3 * (t1 - t0); host_ms["push"] += 1e3 * (t2 - t1); host_ms["finish"] += 1e3 * (t3 - t2)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res = step_device()
    n_sv = int(res.n_sv) if args.warmup else 0
    sampler = ClockSampler(local) if rank == 0 else None
    for k in host_ms:
        host_ms[k] = 0.0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        res = step_device()
        add_times()
    e1.record()
    barrier()
    t1 = time.perf_counter()
    n_sv = int(res.n_sv)
    k4_sweeps = ctx.k4_sweeps()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop(t0, t1) if sampler else None
    summ = ctx.summary()
    n_anom = int(summ.n_anomalous)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tot = torch.tensor([npairs], device=dev, dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_pairs = float(tot.item())
    else:
        total_pairs = float(npairs)
    value = total_pairs / (ms / 1e3)

    # ---- end to end: host (pinned) columns through bdk_push, H2D inside the timed region -------------
    k1 = dict(ktimes.get("k1_classify", {"ms": 0.0, "launches": 0}))
    gpu_launches = launches[0]
    hcols = synth_torch.to_pinned(cols)
    hsoa = api.soa_from_pointers({k: hcols[k].data_ptr() for k in api.COLUMN_DTYPES})

    def step_host():
        ctx.reset()
        ctx.push_soa(hsoa, n, device=False)
        return ctx.finish_raw()

    for _ in range(min(2, args.warmup)):
        res = step_host()
    barrier()
    e0.record()
    w0 = time.perf_counter()
    esteps = max(1, min(args.steps, 5))
    for _ in range(esteps):
        res = step_host()
    h2d = ctx.h2d_bytes()
    e1.record()
    barrier()
    e_ms_wall = 1e3 * (time.perf_counter() - w0) / esteps
    e_ms = max(e0.elapsed_time(e1) / esteps, 0.0)
    e_ms = max(e_ms, e_ms_wall) if world == 1 else e_ms
    d2h = ctx.d2h_bytes()
    if world > 1:
        t = torch.tensor([e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
    e2e_value = total_pairs / (e_ms / 1e3)

    # ---- the same with the records in the decoder's wire format (12 bytes per record, include/bdk.h: bdk_packed). Packing is the
    # decoder's job (bdk_pack here, outside the timed region like the decoding itself); the device expands each chunk.
    packed = None
    if args.config == 2:
        tp0 = time.perf_counter()
        run = api.PackedRun(hsoa, n, keep=hcols)
        pack_s = time.perf_counter() - tp0

        def step_packed():
            ctx.reset()
            ctx.push_packed(run)
            return ctx.finish_raw()

        for _ in range(min(2, args.warmup)):
            res = step_packed()
        packed_sv = int(res.n_sv)
        barrier()
        e0.record()
        w0 = time.perf_counter()
        for _ in range(esteps):
            res = step_packed()
        p_h2d = ctx.h2d_bytes()
        e1.record()
        barrier()
        p_wall = 1e3 * (time.perf_counter() - w0) / esteps
        p_ms = max(e0.elapsed_time(e1) / esteps, 0.0)
        p_ms = max(p_ms, p_wall) if world == 1 else p_ms
        if world > 1:
            t = torch.tensor([p_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            p_ms = float(t.item())
        packed = {"value": total_pairs / (p_ms / 1e3), "ms_per_step": p_ms, "h2d_bytes_per_step": p_h2d, "d2h_bytes_per_step": ctx.d2h_bytes(),
                  "exceptions": int(run.view.nx), "same_sv_calls_as_columns": packed_sv == n_sv,
                  "pack_host_s": round(pack_s, 3), "pack_host_threads": os.cpu_count() or 1}
        run.close()

    genome = None
    if world > 1 and not args.no_genome and args.config == 2:
        del cols, hcols, dsoa, hsoa
        torch.cuda.empty_cache()
        genome = {"whole_genome": one_job_mode(args, rank, world, local, dev, False),
                  "ctx_only_t": one_job_mode(args, rank, world, local, dev, True)}
        if not args.no_configs45:
            genome["config4_lpt_shards"] = lpt_shards_mode(args, rank, world, local, dev, args.genome_pairs)
            genome["config5_ctx_t"] = one_job_mode(args, rank, world, local, dev, True, plan="grch38", total_pairs=args.ctx_pairs,
                                                   name=f"one job, -t, 24 GRCh38 chromosomes, {args.ctx_pairs / 1e6:g}M read pairs + 2% inter-chromosomal pairs over {world} GPUs (BASELINE configs[4])")

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        k1_ms = k1["ms"] / max(1, k1["launches"])
        achieved = n * BYTES_PER_RECORD / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
        traffic, traffic_src = k1_traffic(n, args.config)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(pairs, args.config), "records_per_gpu": n, "anomalous_reads_per_gpu": n_anom, "sv_calls_per_gpu": n_sv,
                       "sharding": "one chromosome-shaped shard per GPU, no data-path collective" if world > 1 else "single GPU",
                       "l2": f"inputs ({n * HOST_BYTES_PER_RECORD / 1e9:.1f} GB) are larger than L2, no flush needed", "options": "defaults (-c 3 -q 35 -r 2 -y 30)"},
            "roofline": {"bound": "hbm", "kernel": "k1_classify_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": n * BYTES_PER_RECORD, "kernel_ms": k1_ms},
            "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in ktimes.items()},
            "host_call_ms_per_step": {k: v / args.steps for k, v in host_ms.items()},
            "e2e": ({"value": packed["value"], "unit": UNIT, "ms_per_step": packed["ms_per_step"], "h2d_bytes_per_step": packed["h2d_bytes_per_step"],
                     "d2h_bytes_per_step": packed["d2h_bytes_per_step"], "zero_copy_side_columns": True,
                     "input": "pinned host records in the 12-byte wire format (bdk_push_packed), packed outside the timed region like the decoding; "
                              "e2e_columns is the same job from 25-byte columns (bdk_push)",
                     "exceptions": packed["exceptions"], "same_sv_calls_as_columns": packed["same_sv_calls_as_columns"],
                     "pack_host_s": packed["pack_host_s"], "pack_host_threads": packed["pack_host_threads"]} if packed else
                    {"value": e2e_value, "unit": UNIT, "ms_per_step": e_ms, "h2d_bytes_per_step": h2d,
                     "zero_copy_side_columns": h2d < n * HOST_BYTES_PER_RECORD, "d2h_bytes_per_step": d2h}),
            "e2e_columns": {"value": e2e_value, "unit": UNIT, "ms_per_step": e_ms, "h2d_bytes_per_step": h2d,
                            "zero_copy_side_columns": h2d < n * HOST_BYTES_PER_RECORD, "d2h_bytes_per_step": d2h},
            "gpu_launches": gpu_launches, "k4_sweeps": k4_sweeps,
            "clocks": clocks,
        }
        if genome:
            line["one_job_all_gpus"] = genome
        if world == 1 and not args.no_cpu:
            try:
                line["cpu_baseline"], free = cpu_baseline(args.cpu_pairs, max(1, min(os.cpu_count() or 1, 32)))
                if free:
                    line["cpu_baseline_decode_free"] = free
            except Exception as ex:   # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)[:200]}
            try:
                line["file_e2e"] = file_e2e_sample(args.file_pairs)
            except Exception as ex:   # reported, never required
                line["file_e2e"] = {"value": None, "unit": UNIT, "sample": str(ex)[:200]}
            # what the headline ratios compare (the driver computes e2e / reference arm itself)
            cb, cf, fe = line.get("cpu_baseline", {}), line.get("cpu_baseline_decode_free", {}), line.get("file_e2e", {})
            line["comparisons"] = {
                "records_in_host_memory_to_sv": {"ours": line["e2e"]["value"], "reference_decode_free": cf.get("value"),
                                                 "ratio": (line["e2e"]["value"] / cf["value"]) if cf.get("value") else None},
                "bam_file_to_sv": {"ours": fe.get("value"), "reference": cb.get("value"),
                                   "ratio": (fe["value"] / cb["value"]) if fe.get("value") and cb.get("value") else None,
                                   "ours_in_process": (fe.get("in_process") or {}).get("value")}}
            try:
                line["bam_decode"] = bam_decode_sample(args.bam_pairs)
            except Exception as ex:   # reported, never required
                line["bam_decode"] = {"value": None, "unit": UNIT, "sample": str(ex)[:200]}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _digest(summ, table):
    """One hash over everything a job returns (summary statistics, SV rows in output order, per-library counts, copy numbers)."""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    h.update(bytes(summ))
    for a in (table.sv, table.lib_count, table.cn_count, table.copy_number):
        h.update(a.tobytes())
    return h.hexdigest()


def _cat_cols(parts):
    import torch
    return {k: torch.cat([p[k] for p in parts]).contiguous() for k in parts[0]}


def one_job_mode(args, rank, world, local, dev, transchr, plan=None, total_pairs=None, name=""):
    """ONE job over all N GPUs (csrc/comm.cuh): rank r holds a contiguous slice of the globally (tid, pos)-sorted stream in
    HBM; per step: reset, K1 on the local slice, the NCCL exchange of the anomalous reads, K2-K4 on the global stream, the
    complete table on every rank. plan=None: N chromosomes of chr1's length with args.pairs each (whole-genome semantics, or -t);
    plan="grch38": 24 GRCh38 chromosomes, `total_pairs` in total, contiguous chromosome ranges per rank (BASELINE configs[4] with -t).
    `parity`: the digest of (summary, SV table) is the same on every rank AND equals a single-GPU run of the concatenated
    stream on rank 0 (when that fits one GPU). Returns the block rank 0 prints (max over ranks of the device time)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from breakdancer_b200 import api, synth, synth_torch
    seed = 20260104
    if plan == "grch38":
        lens = synth_torch.GRCH38
        ctx_frac = 0.02
        pairs_c, ctxm = synth_torch.genome_plan(total_pairs, ctx_frac, lens)
        ntid = len(lens)
        cuts, acc, tot = [0], 0, sum(pairs_c)               # contiguous chromosome ranges with about equal pairs
        for t in range(ntid):
            acc += pairs_c[t]
            while len(cuts) < world and acc >= tot * len(cuts) / world:
                cuts.append(t + 1)
        while len(cuts) < world:
            cuts.append(ntid)
        cuts.append(ntid)
        shard = lambda t: synth_torch.genome_shard_device(pairs_c[t], seed, dev, t, ntid, ctx_frac, ctx_counts=ctxm[t], lens=lens)
        mine = list(range(cuts[rank], cuts[rank + 1]))
        genome = [(f"chr{i + 1}", l) for i, l in enumerate(lens)]
    else:
        ctx_frac = 0.02 if transchr else 0.002
        ntid = world
        shard = lambda t: synth_torch.genome_shard_device(args.pairs, seed, dev, t, world, ctx_frac)
        mine = [rank]
        genome = [(f"chr{i + 1}", synth_torch.CHR1_LEN) for i in range(world)]
    cols = _cat_cols([shard(t) for t in mine]) if mine else None
    n = cols["pos"].numel() if cols is not None else 0
    lib = synth.LibSpec("lib1", "syn_genome.bam", synth_torch.MEAN, synth_torch.STD, synth_torch.READLEN, ["rg1"])
    wl = synth.Workload({}, genome, [lib], ["rg1"], ["lib1"], ["syn_genome.bam"])
    cfg = api.BamConfig(text=wl.config_text())
    opts = api.Options(transchr_rearrange=bool(transchr))
    bundle = api.ParamBundle(opts, cfg.libs, cfg.nbam, np.zeros(1, np.int32), np.zeros(1, np.int32), cfg.window, ntid)
    ctx = api.Context(bundle, local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.comm_init_from_dist()
    dsoa = synth_torch.soa_of(cols) if n else None

    def step():
        ctx.reset()
        if n:
            ctx.push_soa(dsoa, n, device=True)
        return ctx.finish_raw()

    for _ in range(3):
        res = step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(1, min(args.steps, 5))
    kt = {}
    e0.record()
    for _ in range(steps):
        res = step()
        for k, v in ctx.kernel_times().items():
            kt[k] = kt.get(k, 0.0) + v["ms"] / steps
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    names = ["k1_classify", "comm_gather_reads", "k2_regions", "k3_links_graph", "k4_sv_score", "d2h_results"]
    t = torch.tensor([ms] + [kt.get(k, 0.0) for k in names], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = torch.tensor([n // 2, ctx.comm_bytes()], device=dev, dtype=torch.float64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    # ---- parity: every rank holds the same result, and it is the single-GPU result of the concatenated stream
    ctx.reset()
    if n:
        ctx.push_soa(dsoa, n, device=True)
    summ = ctx.summary()
    table = ctx.finish()
    mine_digest = _digest(summ, table)
    digests = [None] * world
    dist.all_gather_object(digests, mine_digest)
    parity = {"ranks_agree": len(set(digests)) == 1, "digest": digests[0]}
    total_records = int(summ.n_records)
    ctx.close()
    del cols, dsoa
    torch.cuda.empty_cache()
    fits = total_records * HOST_BYTES_PER_RECORD < 45e9          # the stream, the generator's temporaries and the job's work space on one GPU
    if not fits and plan == "grch38" and total_pairs > 300_000_000:
        # too large for one GPU: the same plan (24 chromosomes, same options, same cuts) at 300 M pairs, all ranks + one GPU
        small = one_job_mode(args, rank, world, local, dev, transchr, plan=plan, total_pairs=300_000_000, name="parity run")
        parity["at_300M_pairs"] = small["parity"]
        parity["parity_ok"] = small["parity"].get("parity_ok")
        parity["single_gpu_digest"] = "full size does not fit one GPU: see at_300M_pairs (same plan, 300 M pairs)"
    elif fits:
        if rank == 0:
            try:
                allc = _cat_cols([shard(t) for t in range(ntid)])
                c1 = api.Context(bundle, local)
                c1.set_stream(torch.cuda.current_stream().cuda_stream)
                c1.push_soa(synth_torch.soa_of(allc), allc["pos"].numel(), device=True)
                s1 = c1.summary()
                t1 = c1.finish()
                parity["single_gpu_digest"] = _digest(s1, t1)
                parity["single_gpu_sv_calls"] = int(len(t1.sv))
                c1.close()
                del allc
            except Exception as ex:   # reported, never fatal
                parity["single_gpu_digest"] = None
                parity["single_gpu_error"] = str(ex)[:200]
            torch.cuda.empty_cache()
            parity["parity_ok"] = bool(parity["ranks_agree"] and parity.get("single_gpu_digest") == parity["digest"])
        dist.barrier()
    else:
        parity["parity_ok"] = None
        parity["single_gpu_digest"] = "skipped: the concatenated stream does not fit one GPU beside the generator's temporaries"
    out = {"value": float(tot[0].item()) / (float(t[0].item()) / 1e3), "unit": UNIT, "ms_per_step": float(t[0].item()),
           "workload": name or (f"one job, {world} chromosomes x {args.pairs / 1e6:g}M pairs (config-2 records + {ctx_frac * 100:g}% inter-chromosomal pairs)"
                                + (", -t" if transchr else ", whole-genome semantics")),
           "records_total": total_records, "anomalous_reads_total": int(summ.n_anomalous), "sv_calls": int(len(table.sv)),
           "nvlink_bytes_received_per_step_all_ranks": float(tot[1].item()), "parity": parity,
           "max_over_ranks_ms": dict(zip(names, [float(x) for x in t[1:].tolist()]))}
    return out


def lpt_shards_mode(args, rank, world, local, dev, total_pairs):
    """BASELINE configs[3]: a 30x whole genome (24 GRCh38 chromosomes) as per-chromosome shards, each an independent job with -o
    semantics (own summary statistics, own window, own region indices), chromosomes LPT-packed onto the GPUs by their pair
    counts (breakdancer_b200/shard.py). No data-path collective; total work is fixed, so this is strong scaling."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from breakdancer_b200 import api, shard, synth, synth_torch
    lens = synth_torch.GRCH38
    ctx_frac = 0.002
    pairs_c, ctxm = synth_torch.genome_plan(total_pairs, ctx_frac, lens)
    ntid = len(lens)
    mine = shard.lpt_pack(pairs_c, world)[rank]
    genome = [(f"chr{i + 1}", l) for i, l in enumerate(lens)]
    lib = synth.LibSpec("lib1", "syn_genome.bam", synth_torch.MEAN, synth_torch.STD, synth_torch.READLEN, ["rg1"])
    wl = synth.Workload({}, genome, [lib], ["rg1"], ["lib1"], ["syn_genome.bam"])
    cfg = api.BamConfig(text=wl.config_text())
    shards = []
    for t in mine:
        cols = synth_torch.genome_shard_device(pairs_c[t], 20260103, dev, t, ntid, ctx_frac, ctx_counts=ctxm[t], lens=lens)
        bundle = api.ParamBundle(api.Options(chr=genome[t][0]), cfg.libs, cfg.nbam, np.zeros(1, np.int32), np.zeros(1, np.int32), cfg.window, ntid)
        shards.append((t, cols, synth_torch.soa_of(cols), cols["pos"].numel(), bundle))
    ctx = api.Context(shards[0][4], local) if shards else None      # the options of all shards are the same (-o: chr_restricted)
    if ctx:
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    def step():
        calls = 0
        for t, cols, soa, n, bundle in shards:
            ctx.reset()
            ctx.push_soa(soa, n, device=True)
            calls += int(ctx.finish_raw().n_sv)
        return calls

    for _ in range(3):
        calls = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(1, min(args.steps, 5))
    e0.record()
    for _ in range(steps):
        calls = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    my_pairs = sum(s[3] for s in shards) // 2
    t = torch.tensor([ms, -ms, float(my_pairs), -float(my_pairs)], device=dev, dtype=torch.float64)
    tot = torch.tensor([my_pairs, calls], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if ctx:
        ctx.close()
    del shards
    torch.cuda.empty_cache()
    return {"value": float(tot[0].item()) / (float(t[0].item()) / 1e3), "unit": UNIT, "ms_per_step": float(t[0].item()), "scaling": "strong",
            "workload": f"24 GRCh38 chromosomes, {tot[0].item() / 1e6:.1f}M read pairs in total, one -o shard per chromosome, LPT-packed onto {world} GPU(s) (BASELINE configs[3])",
            "sv_calls": int(tot[1].item()), "slowest_rank_ms": float(t[0].item()), "fastest_rank_ms": -float(t[1].item()),
            "pairs_on_fullest_rank": float(t[2].item()), "pairs_on_emptiest_rank": -float(t[3].item())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=None, help="read pairs per GPU (BASELINE configs[1]: 50 M; configs[2]: 300 M)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3], help="2: BASELINE configs[1] (the bench line); 3: configs[2], 4 libraries / all SV types")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-pairs", type=int, default=6_000_000, help="total read pairs of the CPU baseline sample")
    ap.add_argument("--ref-pairs", type=int, default=400_000, help="read pairs per process and step of --impl reference")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--bam-pairs", type=int, default=2_000_000, help="size of the BAM sample of the bam_decode leg")
    ap.add_argument("--file-pairs", type=int, default=24_000_000, help="read pairs of the BAM the file_e2e leg runs the drop-in executable on")
    ap.add_argument("--no-genome", action="store_true", help="N > 1: skip the one-job-over-all-GPUs (NCCL exchange) measurements")
    ap.add_argument("--no-configs45", action="store_true", help="N > 1: skip the BASELINE configs[3] (LPT shards) and configs[4] (-t, 1 B pairs) blocks")
    ap.add_argument("--genome-pairs", type=int, default=617_700_000, help="read pairs of the configs[3] whole genome")
    ap.add_argument("--ctx-pairs", type=int, default=1_000_000_000, help="read pairs of the configs[4] -t job")
    args = ap.parse_args()
    if args.pairs is None:
        args.pairs = 300_000_000 if args.config == 3 else 50_000_000
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
