#!/usr/bin/env python
"""tests/gpu_diag.py — stage-by-stage diagnosis of the CUDA engine on a GPU box (not a pytest).

Runs the block transform on a ladder of inputs and compares with the C oracle (oracle/liboracle.so).  On
the first mismatch it re-runs with the engine stopped after round 0 and checks every stage (text reversal,
key packing, sort order + stability, re-ranking) against numpy restatements, printing the first
discrepancy of each stage.  Usage:  python tests/gpu_diag.py [--quick] [--perf]
"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bwtc_b200 as bw  # noqa: E402

orc = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
orc.oracle_bwt_block.restype = ctypes.c_int64
orc.oracle_bwt_block.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p]


def oracle_block(x, starts):
    buf = np.concatenate([x, np.zeros(1, np.uint8)])
    LF = np.zeros(256, np.uint32)
    n = ctypes.c_uint32(0)
    fr = np.zeros(256, np.uint32)
    orc.oracle_bwt_block(buf.ctypes.data, x.size, starts, LF.ctypes.data, ctypes.byref(n), fr.ctypes.data)
    return buf[:-1].copy(), LF[: n.value].copy(), fr


def stage_check(ctx, x, starts):
    n = x.size
    N = n + 1
    ctx.set_debug(1)
    blk = x.copy()
    LF = np.zeros(bw.num_starting_points(n, starts), np.uint32)
    fr = np.zeros(256, np.uint32)
    try:
        ctx.bwt_block(blk, LF, fr)
    except Exception as e:  # noqa: BLE001
        print("   stage run raised:", e)
    st = ctx.stats()
    ctx.set_debug(0)
    print("   stats:", {k: st[k] for k in ("sigma", "bits_per_char", "chars_round0", "key_bytes_round0", "rounds", "live", "passes")})
    T = np.concatenate([x[::-1], np.zeros(1, np.uint8)])
    text = ctx.debug_read(0, np.uint8, N)
    bad = np.nonzero(text != T)[0]
    print("   text reversal:", "OK" if bad.size == 0 else f"MISMATCH at {bad[:5]} got {text[bad[:5]]} want {T[bad[:5]]}")
    hist = np.bincount(x, minlength=256)
    if not (hist == fr).all():
        print("   freqs MISMATCH", np.nonzero(hist != fr)[0][:8])
    present = hist > 0
    present[0] = True
    lut = np.cumsum(present) - 1
    b, c, kb = st["bits_per_char"], st["chars_round0"], st["key_bytes_round0"]
    codes = np.concatenate([lut[T].astype(np.uint64), np.zeros(c + 1, np.uint64)])
    exp_key = np.zeros(N, np.uint64)
    for j in range(c):
        exp_key = (exp_key << np.uint64(b)) | codes[j: j + N]
    kdt = np.uint32 if kb == 4 else np.uint64
    keys = ctx.debug_read(2, kdt, N).astype(np.uint64)
    idx = ctx.debug_read(3, np.uint32, N).astype(np.int64)
    ok_perm = (np.sort(idx) == np.arange(N)).all()
    print("   idx is a permutation:", ok_perm)
    if not ok_perm:
        cnt = np.bincount(np.clip(idx, 0, N - 1), minlength=N)
        print("      missing ids:", np.nonzero(cnt == 0)[0][:8], "dups:", np.nonzero(cnt > 1)[0][:8], "max idx", idx.max())
    srt = (keys[1:] >= keys[:-1]).all()
    print("   keys sorted:", srt)
    if not srt:
        bp = np.nonzero(keys[1:] < keys[:-1])[0]
        print("      first inversions at", bp[:8], "count", bp.size)
    if ok_perm:
        km = keys == exp_key[idx]
        print("   key[j] == pack(idx[j]):", km.all())
        if not km.all():
            bp = np.nonzero(~km)[0]
            print("      first bad j", bp[:5], "idx", idx[bp[:5]], "got", [hex(int(v)) for v in keys[bp[:5]]], "want",
                  [hex(int(v)) for v in exp_key[idx[bp[:5]]]])
        eq = keys[1:] == keys[:-1]
        stab = (~eq | (idx[1:] < idx[:-1])).all()
        print("   stable (equal keys keep descending ids):", stab)
        # expected ranks after round 0
        thresh = 0 if c > N else N - c + 1
        head = np.ones(N, bool)
        head[1:] = (keys[1:] != keys[:-1]) | (idx[:-1] >= thresh)
        pos = np.where(head, np.arange(N), 0)
        hf = np.maximum.accumulate(pos)
        nxt = np.ones(N, bool)
        nxt[:-1] = head[1:]
        single = head & nxt
        exp_rank = np.zeros(N, np.uint32)
        exp_rank[idx] = (hf.astype(np.uint32)) | (single.astype(np.uint32) << np.uint32(31))
        rank = ctx.debug_read(1, np.uint32, N)
        rm = rank == exp_rank
        print("   rank after round 0:", "OK" if rm.all() else f"MISMATCH count {int((~rm).sum())}")
        if not rm.all():
            bp = np.nonzero(~rm)[0]
            print("      suffixes", bp[:6], "got", [hex(int(v)) for v in rank[bp[:6]]], "want", [hex(int(v)) for v in exp_rank[bp[:6]]])
        ctrl = ctx.debug_read(6, np.uint32, 32)
        print("   ctrl live", ctrl[18], "expected", int((~single).sum()), "err", ctrl[19])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--perf", action="store_true")
    args = ap.parse_args()
    lib = bw.load_library()
    print("version:", lib.bwtc_cuda_version().decode(), "devices:", lib.bwtc_cuda_device_count())
    rng = np.random.default_rng(12345)
    cases = []
    for s in [b"mississippi", b"banana", b"abracadabra", b"aaaaaaaa", b"a", b"ab", b"ba"]:
        cases.append((s.decode(), np.frombuffer(s, np.uint8).copy()))
    cases.append(("zeros300", np.zeros(300, np.uint8)))
    cases.append(("ab x150", np.frombuffer(b"ab" * 150, np.uint8).copy()))
    cases.append(("7i mod 251 n1000", ((7 * np.arange(1000)) % 251).astype(np.uint8)))
    for n in [2, 3, 255, 256, 257, 258, 1000, 4095, 4096, 4097, 5000, 70000, 300000]:
        for sigma in [1, 2, 4, 64, 256]:
            cases.append((f"rand n={n} sigma={sigma}", rng.integers(0, sigma, n).astype(np.uint8)))
    for kind in ["markov", "dna", "repetitive", "random"]:
        for n in [1 << 16, 1 << 20]:
            cases.append((f"{kind} n={n}", bw.generate(kind, n, seed=3)))
    if args.quick:
        cases = cases[:40]
    ctx = bw.CudaContext(4 << 20)
    nfail = 0
    diagnosed = 0
    for name, x in cases:
        for starts in (8,):
            want, wLF, wfr = oracle_block(x, starts)
            blk = x.copy()
            LF = np.zeros(bw.num_starting_points(x.size, starts), np.uint32)
            fr = np.zeros(256, np.uint32)
            t0 = time.time()
            try:
                pidx = ctx.bwt_block(blk, LF, fr)
                err = None
            except Exception as e:  # noqa: BLE001
                err = str(e)
            dt = time.time() - t0
            ok = err is None and (blk == want).all() and (LF == wLF).all() and (fr == wfr).all()
            st = ctx.stats()
            print(f"[{'ok' if ok else 'FAIL'}] {name:28s} n={x.size:8d} rounds={st['rounds']} live={st['live'][:4]} "
                  f"c={st['chars_round0']} kb={st['key_bytes_round0']} gpu_ms={st['gpu_ms']:.3f} wall_ms={dt*1e3:.2f}"
                  + (f" ERR {err}" if err else ""))
            if not ok:
                nfail += 1
                if err is None:
                    bp = np.nonzero(blk != want)[0]
                    print("   bwt mismatches:", bp.size, "first", bp[:6], "LF got", LF[:4], "want", wLF[:4],
                          "freqs ok", (fr == wfr).all())
                if diagnosed < 3:
                    diagnosed += 1
                    stage_check(ctx, x, starts)
    print(f"DIAG SUMMARY: {len(cases) - nfail}/{len(cases)} cases pass")
    if args.perf:
        ctx.close()
        for kind in ["markov", "dna", "repetitive", "random"]:
            for n in [1 << 20, 16 << 20, 32 << 20]:
                c2 = bw.CudaContext(n)
                c2.set_timing(1)
                x = bw.generate(kind, n, seed=5)
                for rep in range(2):
                    blk = x.copy()
                    LF = np.zeros(8, np.uint32)
                    t0 = time.time()
                    c2.bwt_block(blk, LF, None)
                    dt = time.time() - t0
                st = c2.stats()
                print(f"PERF {kind:10s} n={n>>20:3d}MiB rounds={st['rounds']} live={st['live']} passes={st['passes']} "
                      f"c={st['chars_round0']} kb={st['key_bytes_round0']} gpu_ms={st['gpu_ms']:.3f} sort_ms={st['sort_ms']:.3f} "
                      f"wall_ms={dt*1e3:.2f} MB/s(gpu)={n/1e6/(st['gpu_ms']/1e3):.0f} Balg={st['algorithmic_bytes']/n:.0f}B/byte "
                      f"GB/s={st['algorithmic_bytes']/1e9/(st['gpu_ms']/1e3):.0f} sortGB/s={st['sort_bytes']/1e9/max(st['sort_ms'],1e-9)*1e3:.0f}")
                c2.close()
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
