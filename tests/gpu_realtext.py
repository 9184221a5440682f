#!/usr/bin/env python
"""Real (non-synthetic) data probe, not a pytest: concatenates source files found in the image (Python standard
library and site-packages .py files) into blocks, transforms them on the GPU, inverts them with the REFERENCE's
inverse BWT and prints the round profile.  python tests/gpu_realtext.py [MIB] [NBLOCKS]"""
import glob, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bwtc_b200 as bw
from conftest import Reference

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nblocks = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = mib << 20
need = n * nblocks
chunks, have = [], 0
for root in (os.path.dirname(os.__file__), os.path.dirname(os.path.dirname(np.__file__))):
    for path in sorted(glob.glob(os.path.join(root, "**", "*.py"), recursive=True)):
        try:
            b = open(path, "rb").read()
        except OSError:
            continue
        chunks.append(np.frombuffer(b, np.uint8))
        have += len(b)
        if have >= need:
            break
    if have >= need:
        break
data = np.concatenate(chunks)
print("collected", data.size, "bytes of source text from", len(chunks), "files")
ref = Reference(os.path.join(ROOT, "oracle", "_ref", "libbwtc_ref.so"))
ctx = bw.CudaContext(n)
for k in range(min(nblocks, data.size // n)):
    x = data[k * n:(k + 1) * n].copy()
    best = None
    for rep in range(2):
        blk = x.copy()
        LF = np.zeros(8, np.uint32)
        fr = np.zeros(256, np.uint32)
        ctx.bwt_block(blk, LF, fr)
        st = ctx.stats()
        if best is None or st["gpu_ms"] < best["gpu_ms"]:
            best = st
    back = ref.inverse_block(blk, LF)
    ok = bool((back == x).all()) and bool((fr == np.bincount(x, minlength=256)).all())
    r = best["rounds"]
    print(f"block {k}: {mib} MiB source text sigma={best['sigma']} c={best['chars_round0']} keyB={best['key_bytes_round0']} rounds={r} "
          f"live/N={[round(v / best['n_suffixes'], 3) for v in best['live'][:r]]} passes={best['passes'][:r]} "
          f"gpu_ms={best['gpu_ms']:.3f} MB/s={n / 1e6 / (best['gpu_ms'] / 1e3):.0f} reference-inverse-roundtrip={'OK' if ok else 'FAILED'}", flush=True)
    if not ok:
        sys.exit(1)
