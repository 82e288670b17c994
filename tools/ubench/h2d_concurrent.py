"""Concurrent host -> device bandwidth of one process per GPU (run under torchrun): every rank copies 64 MiB from its own
pinned buffer to its own GPU 40 times, all ranks start together; whole transfers vs 8 MiB pieces.  Rank 0 prints one line
per variant: the slowest rank's GB/s and the aggregate.  Evidence for the e2e scaling of bench.py (not shipped)."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, REP = 64 << 20, 40
h = torch.empty(N, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty(N, dtype=torch.uint8, device="cuda")
for name, piece in (("64 MiB per copy", N), ("8 MiB pieces", 8 << 20), ("1 MiB pieces", 1 << 20)):
    for rep in range(2):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(REP):
            for off in range(0, N, piece):
                d[off:off + piece].copy_(h[off:off + piece], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        gbs = REP * N / float(t.item()) / 1e9
        print("ranks %d  %-16s  slowest rank %.1f GB/s  aggregate %.1f GB/s" % (world, name, gbs, gbs * world), flush=True)
if world > 1:
    dist.destroy_process_group()
