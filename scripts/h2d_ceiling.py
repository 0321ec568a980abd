"""What the box gives N processes that copy pinned host memory to their GPUs at the same time (no kernels): the ceiling of the
`e2e` line of bench.py at N GPUs. Launch with torchrun like bench.py.  usage: h2d_ceiling.py [MB per copy] [copies]"""
import json, os, sys, time
import torch
import torch.distributed as dist

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
host = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
host.fill_(1)
dev = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")


def run(active):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if active:
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt if active else 0.0], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


run(True)                                           # warm-up
alone = run(rank == 0)                              # rank 0 alone
together = run(True)                                # all ranks at once
if rank == 0:
    per = reps * (mb << 20)
    print(json.dumps({"n_gpus": world, "mb_per_copy": mb, "copies": reps, "one_rank_alone_GBps": round(per / alone / 1e9, 1),
                      "all_ranks_aggregate_GBps": round(world * per / together / 1e9, 1), "all_ranks_per_gpu_GBps": round(per / together / 1e9, 1)}))
if world > 1:
    dist.destroy_process_group()
