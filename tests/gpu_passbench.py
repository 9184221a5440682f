#!/usr/bin/env python
"""Times the radix passes only (results may be WRONG for experimental builds): random 32 MiB, forced 8x u64 passes."""
import glob, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bwtc_b200 as bw
n = 32 << 20
x = bw.generate("random", n, seed=5)
for lp in sorted(glob.glob(os.path.join(ROOT, "bwtc_b200", "libbwtc_cuda_B*.so"))):
    for kb, c in ((8, 8), (4, 4)):
        ctx = bw.CudaContext(n, lib_path=lp)
        ctx.set_timing(1); ctx.set_round0(c, kb); ctx.set_debug(1 if "a" in os.path.basename(lp).split("_")[-1][:2] and "nolb" in lp else 0)
        best = None
        for _ in range(3):
            blk = x.copy(); LF = np.zeros(8, np.uint32)
            try:
                ctx.bwt_block(blk, LF, None)
            except Exception as e:
                pass
            st = ctx.stats()
            if st["sort_ms"] > 0 and (best is None or st["sort_ms"] < best["sort_ms"]):
                best = st
        print(f"PASS {os.path.basename(lp):36s} key{kb*8} sort_ms={best['sort_ms']:.3f} per-pass_us={1e3*best['sort_ms']/best['sort_launches']:.1f} GB/s={best['sort_bytes']/1e6/best['sort_ms']:.0f}", flush=True)
        ctx.close()
