"""Pinned-memory H2D bandwidth of ONE GPU by stream count and piece size (is the e2e path of bench.py at the link's limit?).
    python tools/ubench/h2d_streams.py   ->  profiles/r02_h2d_streams.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
try:
    import bench
    print("numa:", bench.bind_to_gpu_numa_node(0))
except Exception as e:
    print("bind failed", e)
torch.cuda.set_device(0)
N = 64 << 20
host = torch.empty(N, dtype=torch.uint8).pin_memory()
host.fill_(1)
dev = torch.empty(N, dtype=torch.uint8, device='cuda')
def run(nstreams, piece, reps=20):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_event(e0)
    for r in range(reps):
        off = 0; i = 0
        while off < N:
            with torch.cuda.stream(streams[i % nstreams]):
                dev[off:off+piece].copy_(host[off:off+piece], non_blocking=True)
            off += piece; i += 1
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return reps * N / ms / 1e6
for ns in (1, 2, 4):
    for piece in (64 << 20, 8 << 20, 2 << 20):
        print("streams", ns, "piece MiB", piece >> 20, "GB/s %.1f" % run(ns, piece))
