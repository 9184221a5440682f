#!/usr/bin/env python
"""tests/gpu_sweep.py — perf probes on a GPU box (not a pytest).
  python tests/gpu_sweep.py one  KIND MIB [chars keybytes]      one block (for ncu launch lists)
  python tests/gpu_sweep.py sweep                                round-0 key-shape sweep per input family
  python tests/gpu_sweep.py configs                              one block of each BASELINE config (REPS=1 under ncu)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bwtc_b200 as bw  # noqa: E402


def run(ctx, x, reps=3):
    best = None
    for _ in range(reps):
        blk = x.copy()
        LF = np.zeros(8, np.uint32)
        ctx.bwt_block(blk, LF, None)
        st = ctx.stats()
        if best is None or st["gpu_ms"] < best["gpu_ms"]:
            best = st
    return best


def main():
    mode = sys.argv[1]
    if mode == "one":
        kind, mib = sys.argv[2], int(sys.argv[3])
        n = mib << 20
        ctx = bw.CudaContext(n)
        if len(sys.argv) > 5:
            ctx.set_round0(int(sys.argv[4]), int(sys.argv[5]))
        x = bw.generate(kind, n, seed=5)
        ctx.set_timing(1)
        st = run(ctx, x, reps=int(os.environ.get("REPS", "1")))
        print(st)
    elif mode == "configs":
        reps = int(os.environ.get("REPS", "3"))
        cfgs = (("markov", 1), ("dna", 64), ("repetitive", 16), ("random", 256), ("markov", 32))
        if os.environ.get("CONFIGS"):
            cfgs = tuple((c.split(":")[0], int(c.split(":")[1])) for c in os.environ["CONFIGS"].split(","))
        for kind, mib in cfgs:
            n = mib << 20
            ctx = bw.CudaContext(n)
            ctx.set_timing(1)
            x = bw.generate(kind, n, seed=5)
            st = run(ctx, x, reps=reps)
            r = st["rounds"]
            print(f"CONFIG {kind:10s} {mib:4d}MiB c={st['chars_round0']} keyB={st['key_bytes_round0']} rounds={r} "
                  f"live/N={[round(v / st['n_suffixes'], 4) for v in st['live'][:r]]} passes={st['passes'][:r]} "
                  f"launches={st['kernel_launches']} B_alg={st['algorithmic_bytes']} gpu_ms={st['gpu_ms']:.3f} "
                  f"MB/s={n / 1e6 / (st['gpu_ms'] / 1e3):.0f} B_alg/t={st['algorithmic_bytes'] / 1e9 / (st['gpu_ms'] / 1e3):.0f}GB/s",
                  flush=True)
            ctx.close()
    else:
        sizes = [int(s) for s in os.environ.get("MIBS", "32").split(",")]
        for kind in os.environ.get("KINDS", "markov,dna,repetitive,random").split(","):
            for mib in sizes:
                n = mib << 20
                ctx = bw.CudaContext(n)
                ctx.set_timing(1)
                x = bw.generate(kind, n, seed=5)
                cfgs = [(0, 0)]
                sigma = len(np.unique(x[: 1 << 20])) + (0 if (x[: 1 << 20] == 0).any() else 1)
                b = max(1, int(np.ceil(np.log2(sigma))))
                for kb in (4, 8):
                    cmax = (8 * kb) // b
                    for np_ in range(2, kb + 1):
                        c = (8 * np_) // b
                        if 1 <= c <= cmax and (c, kb) not in cfgs:
                            cfgs.append((c, kb))
                for c, kb in cfgs:
                    ctx.set_round0(c, kb)
                    st = run(ctx, x, reps=2)
                    print(f"SWEEP {kind:10s} {mib:4d}MiB force=({c},{kb}) -> c={st['chars_round0']} kb={st['key_bytes_round0']} "
                          f"rounds={st['rounds']} live1/N={(st['live'][1] if st['rounds'] > 1 else 0)/st['n_suffixes']:.3f} "
                          f"passes={st['passes'][:4]} gpu_ms={st['gpu_ms']:.3f} sort_ms={st['sort_ms']:.3f} "
                          f"MB/s={n/1e6/(st['gpu_ms']/1e3):.0f}", flush=True)
                ctx.close()


if __name__ == "__main__":
    main()
