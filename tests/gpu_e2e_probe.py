#!/usr/bin/env python
"""Where does the gap between device-resident and host-buffer throughput come from?  (not a pytest)  16 x 32 MiB Markov blocks
through the pipeline from pinned host buffers, with the H2D and / or D2H copies suppressed (BWTC_DEBUG_SKIP_COPIES: the results
are then wrong on purpose, only the timing counts).  python tests/gpu_e2e_probe.py"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import numpy as np, torch
    import bwtc_b200 as bw
    n, nb = 32 << 20, 16
    h_in = [torch.from_numpy(bw.generate("markov", n, seed=1000 + i)).pin_memory() for i in range(nb)]
    h_out = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(nb)]
    d_in = [t.cuda() for t in h_in]; d_out = [torch.empty_like(t) for t in d_in]
    pipe = bw.Pipeline(n, depth=6)
    # prime the contexts' d_in with real data so that a skipped H2D still sorts text
    pipe.run_ptrs([t.data_ptr() for t in h_in], [t.data_ptr() for t in h_out], [n] * nb, 8, on_device=False, want_stats=False)
    for name, ip, op, dev in (("host", h_in, h_out, False), ("device", d_in, d_out, True)):
        a, b = [t.data_ptr() for t in ip], [t.data_ptr() for t in op]
        for _ in range(2):
            pipe.run_ptrs(a, b, [n] * nb, 8, on_device=dev, want_stats=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(6):
            pipe.run_ptrs(a, b, [n] * nb, 8, on_device=dev, want_stats=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"SKIP={os.environ.get('BWTC_DEBUG_SKIP_COPIES', '0')} {name:6s}: {6 * nb * n / 1e6 / dt:.0f} MB/s", flush=True)
    pipe.close()
else:
    for skip in ("0", "1", "2", "3"):
        env = dict(os.environ, BWTC_DEBUG_SKIP_COPIES=skip)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], capture_output=True, text=True, env=env)
        print(r.stdout.strip() or r.stderr[-400:], flush=True)
