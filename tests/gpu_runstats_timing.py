#!/usr/bin/env python
"""A/B timing of the GPU run statistics (not a pytest): pipelined compress of 512 MiB with 32 MiB blocks, coder 'H', with
BWTC_RUN_STATS=1 / 0, for a repetitive and a text-like input.  python tests/gpu_runstats_timing.py"""
import json, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bwtc_b200 as bw
TOOL = os.path.join(ROOT, "tests", "integration_tool.py")
with tempfile.TemporaryDirectory() as d:
    for kind in ("repetitive", "markov"):
        x = np.concatenate([bw.generate(kind, 32 << 20, seed=90 + i) for i in range(16)])
        src = os.path.join(d, kind + ".bin"); x.tofile(src)
        for rs in ("1", "0"):
            env = dict(os.environ, BWTC_RUN_STATS=rs, GLIBC_TUNABLES="glibc.malloc.hugetlb=1")
            best = None
            for rep in range(2):
                r = subprocess.run([sys.executable, TOOL, "pipe_compress", src, os.path.join(d, "o.bwtc"), "181375309", "H", "c", "8",
                                    str(os.cpu_count()), "0", "0", "4"], capture_output=True, text=True, env=env)
                o = json.loads(r.stdout.strip().splitlines()[-1])
                if best is None or o["timings"]["total"] < best["timings"]["total"]:
                    best = o
            t = best["timings"]
            print(f"{kind:10s} run_stats={rs}: total {t['total']:.3f} s ({x.size / 1e6 / t['total']:.0f} MB/s) encoder busy {t['encoder_busy_sum']:.2f} core-s "
                  f"served {best.get('run_statistics_served')} size {best['rc']}", flush=True)
