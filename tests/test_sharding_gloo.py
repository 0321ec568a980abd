"""N > 1 host logic on CPU: chromosome packing and the gather of per-chromosome SV tables over
torch.distributed (gloo, world_size 2). The per-chromosome engine here is the oracle; on the GPU box the
same driver (shard.run_sharded) is given a bdk context per rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from breakdancer_b200 import api, shard, synth
from oracle import oracle
from tests import util


def test_lpt_pack_balances_and_is_deterministic():
    w = [g[1] for g in synth.GRCH38]
    for n in (1, 2, 4, 8):
        bins = shard.lpt_pack(w, n)
        assert sorted(i for b in bins for i in b) == list(range(len(w)))
        loads = [sum(w[i] for i in b) for b in bins]
        assert max(loads) <= sum(w) / n * 1.08 + max(w) * (n == 8) * 0.0 or n == 1
        assert bins == shard.lpt_pack(w, n)
    assert shard.lpt_pack([5, 0, 3], 2) == [[0], [2]]


def _engine(w, opts_kw):
    def run(tid, cols):
        o = api.Options(chr=w.genome[tid][0], **opts_kw)
        b, _, *_ = util.workload_bundle(w, o)
        r = oracle.run(b, cols)
        return r.table.sv.tobytes(), r.table.lib_count.tobytes(), len(r.regions)
    return run


def _worker(rank, world, port, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.generate(util.GENOME3, util.LIBS4, 60000, seed=seed, anomaly_frac=0.04)
    res = shard.run_sharded(w.cols, len(w.genome), rank, world, _engine(w, {}))
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_run_equals_per_chromosome_runs():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 21, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = synth.generate(util.GENOME3, util.LIBS4, 60000, seed=21, anomaly_frac=0.04)
    want = shard.run_sharded(w.cols, len(w.genome), 0, 1, _engine(w, {}))
    assert [t for t, _ in got] == [0, 1, 2] and got == want
    assert sum(len(r[0]) for _, r in got) > 0


def _slice_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # what a rank does before the GPU part of a one-job run: rank 0 obtains the NCCL id (no GPU needed), everyone gets it,
    # and every rank takes its slice of the stream
    box = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    sl = shard.stream_slices(n, world)[rank]
    t = torch.tensor([sl.start, sl.stop, sl.stop - sl.start], dtype=torch.int64)
    out = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(out, t)
    if rank == 0:
        q.put((len(box[0]), [o.tolist() for o in out]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stream_slices_and_communicator_id():
    assert [(s.start, s.stop) for s in shard.stream_slices(10, 3)] == [(0, 3), (3, 6), (6, 10)]
    assert [(s.start, s.stop) for s in shard.stream_slices(2, 4)] == [(0, 0), (0, 1), (1, 1), (1, 2)]
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_slice_worker, args=(r, 2, port, 1001, q)) for r in range(2)]
    for p in procs:
        p.start()
    idlen, slices = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert idlen == 128
    assert slices == [[0, 500, 500], [500, 1001, 501]]


def test_plan_and_shards_from_indexed_bams(tmp_path):
    """Per-chromosome shards straight from bam files: weights from the .bai alone, every rank decodes only its chromosomes
    through the index; together the ranks see exactly the records of a whole-file decode, chromosome by chromosome."""
    import subprocess
    samtools = os.path.join(util.ROOT, "oracle", "_ref", "samtools")
    if not os.path.exists(samtools):
        pytest.skip("oracle/_ref/samtools not built")
    w = synth.generate(util.GENOME3, util.LIBS4, 50000, seed=4, anomaly_frac=0.05)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        assert shard.plan_from_index(["missing.bam"], 2) is None
        for bam, cols in synth.split_by_bam(w).items():
            api.write_bam(bam, [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=1)
            subprocess.check_call([samtools, "index", bam])
        cfg = api.BamConfig(text=w.config_text())
        assert api.bam_reference_names(cfg.bam_files[0]) == [g[0] for g in w.genome]
        whole = api.BamStream(cfg, threads=2)
        true_counts = np.bincount(whole.cols["tid"], minlength=len(w.genome))
        weights, bins = shard.plan_from_index(cfg.bam_files, 2)
        assert weights.tolist() == true_counts.tolist()              # the synthetic bams hold primary, placed records only
        assert bins == shard.lpt_pack(true_counts.tolist(), 2)
        seen = {}
        for rank in range(2):
            def run(tid, name, st):
                assert name == w.genome[tid][0]
                return {k: v.copy() for k, v in st.cols.items()}
            for tid, cols in shard.run_sharded_bams(cfg, rank, 2, run, gather=False):
                seen[tid] = cols
        assert sorted(seen) == [t for t in range(len(w.genome)) if true_counts[t]]
        sl = shard.chromosome_slices(whole.cols, len(w.genome))
        for tid, cols in seen.items():
            for k in ("pos", "mpos", "tid", "mtid", "isize", "flag", "mapq", "qlen", "qid"):
                assert np.array_equal(cols[k], whole.cols[k][sl[tid]]), (tid, k)
    finally:
        os.chdir(cwd)


def _bam_engine(w, cfg):
    """run_chromosome for run_sharded_bams: the oracle on the columns the rank decoded for that chromosome (-o semantics)."""
    def run(tid, name, st):
        o = api.Options(chr=name)
        b = api.ParamBundle.from_stream(o, cfg, st)
        r = oracle.run(b, {k: v.copy() for k, v in st.cols.items()})
        return st.n, r.table.sv.tobytes(), len(r.regions)
    return run


def _bam_worker(rank, world, port, d, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    os.chdir(d)
    w = synth.generate(util.GENOME3, util.LIBS4, 40000, seed=6, anomaly_frac=0.05)
    cfg = api.BamConfig(text=w.config_text())
    res = shard.run_sharded_bams(cfg, rank, world, _bam_engine(w, cfg), threads=2)
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_from_indexed_bams_equal_the_single_rank_run(tmp_path):
    import subprocess
    samtools = os.path.join(util.ROOT, "oracle", "_ref", "samtools")
    if not os.path.exists(samtools):
        pytest.skip("oracle/_ref/samtools not built")
    w = synth.generate(util.GENOME3, util.LIBS4, 40000, seed=6, anomaly_frac=0.05)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for bam, cols in synth.split_by_bam(w).items():
            api.write_bam(bam, [g[0] for g in w.genome], [g[1] for g in w.genome], w.rg_names, cols, level=1)
            subprocess.check_call([samtools, "index", bam])
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_bam_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
        for p in procs:
            p.start()
        got = q.get(timeout=300)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        cfg = api.BamConfig(text=w.config_text())
        want = shard.run_sharded_bams(cfg, 0, 1, _bam_engine(w, cfg), threads=2)
        assert [t for t, _ in got] == [0, 1, 2] and got == want
        assert sum(r[0] for _, r in got) == w.n and sum(len(r[1]) for _, r in got) > 0
    finally:
        os.chdir(cwd)
