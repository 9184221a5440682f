#!/usr/bin/env python
"""Host-path diagnostic (not a pytest): the same 32 MiB blocks through every host path of the engine — pageable / pinned /
device-resident buffers, single context and pipeline — compared with each other, plus per-path latency."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bwtc_b200 as bw

n = int(os.environ.get("MIB", "32")) << 20
nb = int(os.environ.get("NB", "6"))
kind = os.environ.get("KIND", "markov")
blocks = [bw.generate(kind, n, seed=1000 + i) for i in range(nb)]
ctx = bw.CudaContext(n)
ref = []
t_page = []
for x in blocks:
    b = x.copy(); LF = np.zeros(8, np.uint32); fr = np.zeros(256, np.uint32)
    t0 = time.perf_counter(); ctx.bwt_block(b, LF, fr); t_page.append(time.perf_counter() - t0)
    ref.append((b, LF.copy(), fr.copy()))
print("pageable ctx: ms/block", [round(t * 1e3, 2) for t in t_page], "gpu_ms", round(ctx.stats()["gpu_ms"], 3))
# pinned ctx
t_pin = []
for i, x in enumerate(blocks):
    t = torch.empty(n, dtype=torch.uint8).pin_memory(); b = t.numpy(); b[:] = x
    LF = np.zeros(8, np.uint32); fr = np.zeros(256, np.uint32)
    t0 = time.perf_counter(); ctx.bwt_block(b, LF, fr); t_pin.append(time.perf_counter() - t0)
    print("pinned ctx block", i, "bytes", bool(np.array_equal(b, ref[i][0])), "LF", bool((LF == ref[i][1]).all()), "fr", bool((fr == ref[i][2]).all()),
          "ndiff", int((b != ref[i][0]).sum()))
print("pinned ctx: ms/block", [round(t * 1e3, 2) for t in t_pin], "gpu_ms", round(ctx.stats()["gpu_ms"], 3))
# device ctx
for i, x in enumerate(blocks):
    d_in = torch.from_numpy(x).cuda(); d_out = torch.empty_like(d_in)
    LF = np.zeros(8, np.uint32); fr = np.zeros(256, np.uint32)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ctx.bwt_block_device(d_in.data_ptr(), d_out.data_ptr(), n, LF, fr); dt = time.perf_counter() - t0
    o = d_out.cpu().numpy()
    print("device ctx block", i, "bytes", bool(np.array_equal(o, ref[i][0])), "LF", bool((LF == ref[i][1]).all()), "ndiff", int((o != ref[i][0]).sum()),
          "ms", round(dt * 1e3, 2), "gpu_ms", round(ctx.stats()["gpu_ms"], 3))
ctx.close()
for depth in (1, 3, 6):
    pipe = bw.Pipeline(n, depth=depth)
    d_in = [torch.from_numpy(x).cuda() for x in blocks]; d_out = [torch.empty_like(t) for t in d_in]
    h_in = [torch.from_numpy(x).pin_memory() for x in blocks]; h_out = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in blocks]
    for rep in range(2):
        LF, nLF, fr, st = pipe.run_ptrs([t.data_ptr() for t in d_in], [t.data_ptr() for t in d_out], [n] * nb, 8, on_device=True)
        okd = [bool(np.array_equal(d_out[i].cpu().numpy(), ref[i][0])) and bool((LF[i, :8] == ref[i][1]).all()) for i in range(nb)]
        t0 = time.perf_counter()
        LFh, nLFh, frh, _ = pipe.run_ptrs([t.data_ptr() for t in h_in], [t.data_ptr() for t in h_out], [n] * nb, 8, on_device=False)
        dt = time.perf_counter() - t0
        okh = [bool(np.array_equal(h_out[i].numpy(), ref[i][0])) and bool((LFh[i, :8] == ref[i][1]).all()) for i in range(nb)]
        work = [x.copy() for x in blocks]
        t1 = time.perf_counter()
        LFp, nLFp, frp, _ = pipe.run(work, 8)
        dtp = time.perf_counter() - t1
        okp = [bool(np.array_equal(work[i], ref[i][0])) for i in range(nb)]
        print(f"pipeline depth {depth} rep {rep}: device {okd} pinned {okh} ({nb * n / 1e6 / dt:.0f} MB/s) pageable {okp} ({nb * n / 1e6 / dtp:.0f} MB/s)", flush=True)
    pipe.close()
