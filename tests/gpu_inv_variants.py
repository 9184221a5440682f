#!/usr/bin/env python
"""Times the inverse transform of experiment builds bwtc_b200/libbwtc_cuda_V*.so (not a pytest)."""
import glob, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bwtc_b200 as bw
libs = sorted(glob.glob(os.path.join(ROOT, "bwtc_b200", "libbwtc_cuda_V*.so")))
for kind, mib in (("markov", 32), ("dna", 64), ("repetitive", 16), ("random", 128)):
    n = mib << 20
    x = bw.generate(kind, n, seed=73)
    for lp in libs:
        ctx = bw.CudaContext(n, lib_path=lp)
        b = x.copy(); LF = np.zeros(8, np.uint32)
        ctx.bwt_block(b, LF, None)
        fwd = b.copy()
        best = 1e9
        for _ in range(3):
            b[:] = fwd
            ctx.inverse_block(b, LF)
            best = min(best, ctx.stats()["gpu_ms"])
        ok = bool(np.array_equal(b, x))
        print(f"INV {os.path.basename(lp):28s} {kind:10s} {mib:4d} MiB: {best:.3f} ms = {n / 1e6 / (best / 1e3):.0f} MB/s ok={ok}", flush=True)
        ctx.close()
