#!/usr/bin/env python
"""Throughput of the batched pipeline on many small blocks (BASELINE config 1 shape: 1 MiB blocks)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bwtc_b200 as bw
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = mib << 20
blocks = [bw.generate("markov", n, seed=100 + i) for i in range(nb)]
for depth in [int(d) for d in os.environ.get('DEPTHS', '1,2,4,8').split(',')]:
    pipe = bw.Pipeline(n, depth=depth)
    work = [b.copy() for b in blocks]
    pipe.run(work, 8)
    work = [b.copy() for b in blocks]
    t0 = time.perf_counter()
    pipe.run(work, 8)
    dt = time.perf_counter() - t0
    print(f"blocks {nb} x {mib} MiB depth {depth:2d}: {nb*n/1e6/dt:8.0f} MB/s  ({dt*1e3/nb:.3f} ms/block)", flush=True)
    pipe.close()
