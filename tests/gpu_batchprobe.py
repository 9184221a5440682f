#!/usr/bin/env python
"""Device time of one batch of small blocks vs the same blocks one by one (not a pytest).
  python tests/gpu_batchprobe.py [KIND] [BLOCK_KIB] [COUNT]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bwtc_b200 as bw

kind = sys.argv[1] if len(sys.argv) > 1 else "markov"
kib = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cnt = int(sys.argv[3]) if len(sys.argv) > 3 else 32
n = kib << 10
blocks = [bw.generate(kind, n, seed=300 + i) for i in range(cnt)]
ctx = bw.CudaContext(cnt * (n + 1))
for rep in range(3):
    work = [b.copy() for b in blocks]
    ctx.bwt_blocks(work, 8)
    st = ctx.stats()
print(f"batch  {cnt} x {kib} KiB {kind}: gpu_ms={st['gpu_ms']:.3f} ({cnt * n / 1e6 / st['gpu_ms'] * 1e3:.0f} MB/s) c={st['chars_round0']} "
      f"keyB={st['key_bytes_round0']} bits={st['bits_per_char']} rounds={st['rounds']} live={st['live'][:st['rounds']]} "
      f"launches={st['kernel_launches']}")
tot = 0.0
LF = np.zeros(8, np.uint32)
for rep in range(2):
    tot = 0.0
    for b in blocks:
        w = b.copy()
        ctx.bwt_block(w, LF[: bw.num_starting_points(n, 8)], None)
        tot += ctx.stats()["gpu_ms"]
print(f"single {cnt} x {kib} KiB {kind}: sum gpu_ms={tot:.3f} ({cnt * n / 1e6 / tot * 1e3:.0f} MB/s)")
