#!/usr/bin/env python
"""Largest block the engine accepts (default 0x3FFFFFF0 bytes = 1 GiB; BWTC_CUDA_MAX_BLOCK is 0x7FFFFFFD): property check
(not a pytest: ~40 GB of device scratch).  python tests/gpu_maxblock.py [KIND] [BYTES]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bwtc_b200 as bw

kind = sys.argv[1] if len(sys.argv) > 1 else "random"
n = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x3FFFFFF0
t0 = time.time()
x = bw.generate(kind, n, seed=77)
print("generated", n, "bytes in %.1f s" % (time.time() - t0), flush=True)
ctx = bw.CudaContext(n)
blk = x.copy()
LF = np.zeros(8, np.uint32)
fr = np.zeros(256, np.uint32)
t0 = time.time()
pidx = ctx.bwt_block(blk, LF, fr)
st = ctx.stats()
print("transform %.2f s wall, gpu_ms=%.1f rounds=%d live=%s passes=%s launches=%d" %
      (time.time() - t0, st["gpu_ms"], st["rounds"], st["live"][:st["rounds"]], st["passes"][:st["rounds"]], st["kernel_launches"]), flush=True)
ctx.close()
hist = np.bincount(x, minlength=256)
assert (fr == hist).all(), "freqs"
assert (np.bincount(blk, minlength=256) == hist).all(), "BWT is not a permutation of the block"
N = n + 1
T = np.concatenate([x[::-1], np.zeros(1, np.uint8)])
xs = N // 8
pos = [0] + [N - j * xs for j in range(1, 8)]
keys = [bytes(T[p: p + 256]) for p in pos]
assert list(np.argsort(LF)) == sorted(range(8), key=lambda i: keys[i]), "LFpowers order"
# spot-check L[r] = T[SA[r]-1] through the sampled suffixes: out[LF[j]] is the character before suffix pos[j]
for j in range(1, 8):
    r = int(LF[j])
    want = T[pos[j] - 1]
    got = blk[r] if r != pidx else None
    if r < pidx or r > pidx:
        # ranks above N-1 do not exist; rank N-1's byte sits in the hole at pidx
        if r == N - 1:
            got = blk[pidx]
        assert got == want, (j, r, got, want)
print("OK", kind, n, "pidx", pidx, "MB/s %.0f" % (n / 1e6 / (st["gpu_ms"] / 1e3)))
