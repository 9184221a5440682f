#!/usr/bin/env python
"""Copy-only ceiling of the box (not a pytest): every rank moves 32 MiB pinned buffers host->device and device->host
concurrently on two streams, no kernels, for a few seconds; rank 0 prints the aggregate GB/s per direction.  This is
the upper bound of the end-to-end (host-buffer) BWT throughput at N GPUs: the engine moves n bytes in and n bytes out
per block.   torchrun --nproc-per-node N tests/gpu_copyceiling.py      (or plain python for one GPU)"""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 32 << 20
nbuf = 8
h_in = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
h_out = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
d_a = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
d_b = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for mode in ("h2d", "d2h", "both"):
    for rep in range(2):  # first repetition warms up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        iters = 40
        for it in range(iters):
            k = it % nbuf
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s_in):
                    d_a[k].copy_(h_in[k], non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s_out):
                    h_out[k].copy_(d_b[k], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[mode] = world * iters * n / 1e9 / float(t.item())
if rank == 0:
    print(json.dumps({"n_gpus": world, "buffer_mib": 32,
                      "h2d_only_GBps": res["h2d"], "d2h_only_GBps": res["d2h"],
                      "both_directions_GBps_each": res["both"],
                      "note": "aggregate over all ranks, pinned host memory, max-over-ranks time; 'both' = h2d and d2h "
                              "running concurrently, figure is per direction"}))
if world > 1:
    dist.destroy_process_group()
