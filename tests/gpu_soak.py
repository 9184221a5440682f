#!/usr/bin/env python
"""Soak run (not a pytest): for SOAK_SECONDS, random mixes of block kinds and sizes go through a pipeline (several
streams in flight, batches of small blocks) and, one by one, through a single context; every block's bytes, LFpowers
and freqs must agree between the two paths and between repetitions.  Catches rare ordering problems of the
look-back kernels that a fixed test set would miss.  python tests/gpu_soak.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bwtc_b200 as bw

seconds = float(os.environ.get("SOAK_SECONDS", "60"))
rng = np.random.default_rng(int(os.environ.get("SOAK_SEED", "1")))
kinds = ["markov", "dna", "repetitive", "random"]
cap = 8 << 20
single = bw.CudaContext(cap)
pipes = {}
t_end = time.time() + seconds
rounds = blocks_done = bytes_done = 0
flags_seen = 0
while time.time() < t_end:
    nb = int(rng.integers(4, 40))
    if rng.random() < 0.5:   # a run of equal-sized small blocks (gets batched), last one shorter
        n0 = int(rng.integers(1, 1 << 19))
        sizes = [n0] * (nb - 1) + [int(rng.integers(1, n0 + 1))]
    else:
        sizes = [int(2 ** rng.uniform(0, 23)) for _ in range(nb)]
    kind = kinds[int(rng.integers(0, 4))]
    blocks = [bw.generate(kind, max(1, s), seed=int(rng.integers(0, 1 << 30))) for s in sizes]
    if rng.random() < 0.3:
        for b in blocks:
            b[rng.integers(0, b.size, max(1, b.size // 50))] = 0
    depth = int(rng.integers(1, 6))
    key = (depth,)
    if key not in pipes:
        pipes[key] = bw.Pipeline(cap, depth=depth)
    work = [b.copy() for b in blocks]
    LF, nLF, fr, stats = pipes[key].run(work, 8)
    for i, b in enumerate(blocks):
        w = b.copy()
        k = bw.num_starting_points(b.size, 8)
        lf1 = np.zeros(k, np.uint32)
        fr1 = np.zeros(256, np.uint32)
        single.bwt_block(w, lf1, fr1)
        flags_seen |= single.stats()["flags"] | stats[i]["flags"]
        if not ((w == work[i]).all() and nLF[i] == k and (LF[i, :k] == lf1).all() and (fr[i] == fr1).all()):
            print("MISMATCH", kind, sizes, i, flush=True)
            sys.exit(1)
        blocks_done += 1
        bytes_done += b.size
    rounds += 1
print(f"soak ok: {rounds} pipeline runs, {blocks_done} blocks, {bytes_done / 1e9:.2f} GB, flags seen 0x{flags_seen:x} "
      f"(bit 0 = ticket fallback: {'YES' if flags_seen & 1 else 'never'})")
