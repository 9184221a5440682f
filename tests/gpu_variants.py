#!/usr/bin/env python
"""tests/gpu_variants.py — parity-checks and times experiment builds of the engine (bwtc_b200/libbwtc_cuda_V*.so,
built with __graft_entry__.build_cuda(extra_defs=[...], out_name=...)) on a GPU box."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bwtc_b200 as bw  # noqa: E402

n = int(os.environ.get("MIB", "32")) << 20
inputs = {"markov": (0, 0), "dna": (0, 0), "random": (0, 0)}
data = {k: bw.generate(k, n, seed=5) for k in inputs}
libs = sorted(glob.glob(os.path.join(ROOT, "bwtc_b200", "libbwtc_cuda_V*.so")))
import ctypes
orc = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
orc.oracle_bwt_block.restype = ctypes.c_int64
orc.oracle_bwt_block.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
xs = bw.generate("markov", 300001, seed=9)
wbuf = np.concatenate([xs, np.zeros(1, np.uint8)]); wLF = np.zeros(256, np.uint32); wk = ctypes.c_uint32(0)
orc.oracle_bwt_block(wbuf.ctypes.data, xs.size, 8, wLF.ctypes.data, ctypes.byref(wk), None)
for lp in libs:
    ctx = bw.CudaContext(1 << 20, lib_path=lp)
    blk = xs.copy(); LF = np.zeros(8, np.uint32)
    ctx.bwt_block(blk, LF, None)
    print("CHECK", os.path.basename(lp), "parity", bool((blk == wbuf[:-1]).all() and (LF == wLF[:8]).all()), flush=True)
    ctx.close()
    for kind, (c, kb) in inputs.items():
        ctx = bw.CudaContext(n, lib_path=lp)
        ctx.set_timing(1)
        ctx.set_round0(c, kb)
        best = None
        for _ in range(3):
            blk = data[kind].copy()
            LF = np.zeros(8, np.uint32)
            ctx.bwt_block(blk, LF, None)
            st = ctx.stats()
            if best is None or st["gpu_ms"] < best["gpu_ms"]:
                best = st
        print(f"VAR {os.path.basename(lp):34s} {kind:8s} gpu_ms={best['gpu_ms']:.3f} sort_ms={best['sort_ms']:.3f} "
              f"sortGB/s={best['sort_bytes']/1e6/best['sort_ms']:.0f} launches={best['sort_launches']}", flush=True)
        ctx.close()
