#!/usr/bin/env python
"""Small workload that launches every kernel of the engine on a few hundred KB (single blocks of all four input
families, raw contract, a batch, forced scatter modes) — meant to be run under a memory checker where one is
available; on its own it is a 2-second smoke run (not a pytest)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bwtc_b200 as bw

n = int(os.environ.get("SAN_BYTES", str(300_000)))
ctx = bw.CudaContext(4 * (n + 1))
LF = np.zeros(8, np.uint32)
fr = np.zeros(256, np.uint32)
for kind in ("markov", "dna", "repetitive", "random"):
    x = bw.generate(kind, n, seed=3)
    ctx.bwt_block(x, LF, fr)
    st = ctx.stats()
    print(kind, "rounds", st["rounds"], "launches", st["kernel_launches"], flush=True)
# raw contract + an exact multiple of the radix tile
T = np.concatenate([bw.generate("markov", 8 * 4096 - 1, seed=4), np.zeros(1, np.uint8)])
U = np.empty_like(T)
ctx.divbwtf(T, U, LF, fr)
# batch of 4 blocks (last shorter)
blocks = [bw.generate("markov", n, seed=10 + k) for k in range(3)] + [bw.generate("markov", n // 3, seed=9)]
ctx.bwt_blocks(blocks, 8)
print("batch", ctx.stats()["n_suffixes"], flush=True)
ctx.close()
for env in ({"BWTC_RERANK_WINDOW_MB": "1"}, {"BWTC_SEG": "0", "BWTC_RERANK_WINDOW_MB": "1", "BWTC_BUCKET_MIN_WINDOWS": "0"}):
    os.environ.update(env)
    c2 = bw.CudaContext(n)
    for kind in ("markov", "repetitive"):
        x = bw.generate(kind, n, seed=5)
        c2.bwt_block(x, LF, fr)
    c2.close()
    print("knobs", env, "ok", flush=True)
