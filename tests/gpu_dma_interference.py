#!/usr/bin/env python
"""Do PCIe copies slow the kernels down?  (not a pytest)  Device-resident pipeline throughput (16 x 32 MiB Markov blocks, depth 6)
while a side thread keeps an unrelated copy stream busy: nothing / D2H / H2D / both, 32 MiB pinned buffers, paced to about the
rate the end-to-end path needs (one copy per ~2.7 ms) or unpaced.  python tests/gpu_dma_interference.py"""
import os, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bwtc_b200 as bw

n, nb = 32 << 20, 16
d_in = [torch.from_numpy(bw.generate("markov", n, seed=1000 + i)).cuda() for i in range(nb)]
d_out = [torch.empty_like(t) for t in d_in]
pipe = bw.Pipeline(n, depth=6)
a, b = [t.data_ptr() for t in d_in], [t.data_ptr() for t in d_out]
hp = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
dv = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
stop = False

def side(mode, pace):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    i = 0
    while not stop:
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s1):
                hp[i % 2].copy_(dv[i % 2], non_blocking=True)
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s2):
                dv[2 + i % 2].copy_(hp[2 + i % 2], non_blocking=True)
        i += 1
        if pace:
            time.sleep(pace)
        if i % 8 == 0:
            s1.synchronize(); s2.synchronize()
    s1.synchronize(); s2.synchronize()

for mode, pace in (("none", 0), ("d2h", 0.0027), ("h2d", 0.0027), ("both", 0.0027), ("d2h", 0), ("both", 0)):
    stop = False
    th = None
    if mode != "none":
        th = threading.Thread(target=side, args=(mode, pace)); th.start()
    for _ in range(2):
        pipe.run_ptrs(a, b, [n] * nb, 8, on_device=True, want_stats=False)
    pipe.timing_begin()
    for _ in range(6):
        pipe.run_ptrs(a, b, [n] * nb, 8, on_device=True, want_stats=False)
    ms = pipe.timing_end()
    stop = True
    if th: th.join()
    print(f"side copies {mode:5s} pace {pace}: device-resident value {6 * nb * n / 1e6 / (ms / 1e3):.0f} MB/s", flush=True)
pipe.close()
