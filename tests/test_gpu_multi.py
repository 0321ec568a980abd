"""Multi-GPU whole-genome mode (include/bdk.h: bdk_comm_*, csrc/comm.cuh): one job whose record stream is cut
into contiguous slices, one per GPU; the result on EVERY rank must equal the oracle's result on the whole
stream bit for bit (log p within 1e-6) -- summary, anomalous-read stream, regions, SV table, supporting reads.

* nranks = 1 runs the whole exchange machinery on the single GPU of the standard `-m gpu` run;
* the world-size-2/4 tests need that many GPUs (gpurun --gpus N) and are skipped otherwise."""
import os
import socket

import numpy as np
import pytest

from breakdancer_b200 import api, synth
from oracle import oracle
from tests import util

pytestmark = pytest.mark.gpu

CASES = {
    "default": dict(opts=dict(), pairs=60000, seed=5, anomaly_frac=0.04),
    "transchr": dict(opts=dict(transchr_rearrange=True), pairs=60000, seed=6, anomaly_frac=0.06),
    "cn_lib_small_buffer": dict(opts=dict(CN_lib=True, buffer_size=3, min_read_pair=1, score_threshold=0), pairs=30000, seed=7, anomaly_frac=0.05),
    "fisher_long_insert": dict(opts=dict(fisher=True, Illumina_long_insert=True), pairs=40000, seed=8, anomaly_frac=0.04),
    "empty_tail_rank": dict(opts=dict(), pairs=20000, seed=9, anomaly_frac=0.03, cuts="lopsided"),
}


def _workload(case):
    w = synth.generate(util.GENOME3, util.LIBS4, case["pairs"], seed=case["seed"], anomaly_frac=case["anomaly_frac"], somatic_frac=0.3)
    b, cols, *_ = util.workload_bundle(w, api.Options(**case["opts"]))
    return b, cols


def _cuts(n, world, how):
    if how == "lopsided":            # the last rank gets nothing, the first almost everything
        e = [0] + [max(0, n - 7 * (world - 1 - r)) for r in range(1, world)] + [n]
        return e
    # uneven on purpose, never on a chromosome boundary
    f = np.cumsum([0.0] + [1.0 + 0.37 * r for r in range(world)])
    return [int(n * x / f[-1]) for x in f]


def _run_rank(case_name, rank, world, device, unique_id):
    case = CASES[case_name]
    b, cols = _workload(case)
    n = len(cols["pos"])
    e = _cuts(n, world, case.get("cuts"))
    mine = {k: np.ascontiguousarray(v[e[rank]:e[rank + 1]]) for k, v in cols.items()}
    ctx = api.Context(b, device)
    ctx.comm_init(unique_id, rank, world)
    for rep in range(2):             # the second job reuses the context (buffers were swapped by the first)
        ctx.reset()
        if rep == 0 and len(mine["pos"]) > 10:
            h = len(mine["pos"]) // 3
            ctx.push({k: np.ascontiguousarray(v[:h]) for k, v in mine.items()})
            ctx.push({k: np.ascontiguousarray(v[h:]) for k, v in mine.items()})
        else:
            ctx.push(mine)
        summary = ctx.summary()
        table = ctx.finish()
        regions = ctx.regions()
        areads, rr = ctx.areads()
        support = ctx.support()
        ro = oracle.run(b, cols)
        util.assert_result_matches_oracle(ro, table, summary, regions, areads, rr, support, f"{case_name} rank {rank}/{world} job {rep}")
        peers_have_records = e[rank] > 0 or e[rank + 1] < n
        assert world == 1 or ctx.comm_bytes() > 0 or len(areads) == 0 or not peers_have_records
    nsv = len(table.sv)
    ctx.close()
    return nsv


def _worker(rank, world, port, case_name, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    box = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    try:
        nsv = _run_rank(case_name, rank, world, rank, box[0])
        q.put((rank, "ok", nsv))
    except BaseException as ex:   # report instead of hanging the peers' collectives silently
        q.put((rank, f"{type(ex).__name__}: {ex}", -1))
        q.close(); q.join_thread()          # the feeder thread must get the message out before the process goes
        os._exit(1)
    dist.barrier()
    dist.destroy_process_group()


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("case_name", list(CASES))
def test_single_rank_communicator_equals_oracle(case_name):
    nsv = _run_rank(case_name, 0, 1, 0, api.comm_unique_id())
    assert nsv > 0


def _spawn(world, case_name):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, port, case_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = []
    try:
        for _ in range(world):
            got.append(q.get(timeout=240))
            if got[-1][1] != "ok":
                break
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    assert all(g[1] == "ok" for g in got) and len(got) == world, got
    assert len({g[2] for g in got}) == 1 and got[0][2] > 0, got


@pytest.mark.parametrize("case_name", list(CASES))
def test_two_ranks_equal_oracle(case_name):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _spawn(2, case_name)


@pytest.mark.parametrize("case_name", ["default", "transchr"])
def test_four_ranks_equal_oracle(case_name):
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs (gpurun --gpus 4)")
    _spawn(4, case_name)
