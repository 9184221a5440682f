#!/usr/bin/env python
"""Per-stage SM-clock stamps of the LAST radix pass executed (profiling build libbwtc_cuda_prof.so)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bwtc_b200 as bw
n = 32 << 20
lp = sys.argv[1]
ctx = bw.CudaContext(n, lib_path=lp)
x = bw.generate("random", n, seed=5)
ctx.set_round0(8, 8)   # one round, 8 passes of u64 keys, no doubling round afterwards
for rep in range(2):
    blk = x.copy(); LF = np.zeros(8, np.uint32); ctx.bwt_block(blk, LF, None)
st = ctx.stats(); print(st["rounds"], st["passes"], st["gpu_ms"])
tiles = (n + 1 + 4095) // 4096
p = ctx.debug_read(8, np.uint64, tiles * 16).reshape(tiles, 16).astype(np.int64)
names = ["ticket+zero", "load wait", "count+publish", "rank", "scan+stage", "lookback", "write keys", "write vals"]
d = p[:, 1:9] - p[:, 0:8]
ok = (p[:, 8] > 0)
print("tiles stamped", int(ok.sum()), "of", tiles)
for i, nm in enumerate(names[:8]):
    v = d[ok, i]
    print(f"{nm:14s} mean {v.mean():9.0f}  p50 {np.median(v):9.0f}  p90 {np.percentile(v,90):9.0f} cycles")
tot = (p[ok, 8] - p[ok, 0]); print("CTA lifetime mean", tot.mean(), "p50", np.median(tot))

o = ok & (p[:, 11] > 0)
print("bin0 lookback: iters mean %.2f  empty-spins mean %.2f  depth(tiles summed) mean %.1f  own time mean %.0f cycles" % (
    p[o, 9].mean(), p[o, 10].mean(), p[o, 12].mean(), (p[o, 11] - p[o, 5]).mean()))
