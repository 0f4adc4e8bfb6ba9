"""Concurrent host->device bandwidth per rank (pinned memory), to explain how the e2e line scales with N:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/h2d_probe.py
Every rank copies a pinned 400 MB buffer to its GPU 20 times, alone (ranks take turns) and then all at once."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 100 * 1024 * 1024
host = torch.empty(n, dtype=torch.float32).pin_memory()
host.normal_()
dev = torch.empty(n, dtype=torch.float32, device="cuda")


def run(reps=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n * 4 / (time.perf_counter() - t0) / 1e9


run(3)
alone = torch.zeros(world, device="cuda")
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        alone[r] = run()
if world > 1:
    dist.barrier()
together = torch.zeros(world, device="cuda")
together[rank] = run()
if world > 1:
    dist.all_reduce(alone)
    dist.all_reduce(together)
if rank == 0:
    print("H2D GB/s per rank, alone   :", [round(float(x), 1) for x in alone.tolist()])
    print("H2D GB/s per rank, together:", [round(float(x), 1) for x in together.tolist()], "sum", round(float(together.sum()), 1))
if world > 1:
    dist.destroy_process_group()
