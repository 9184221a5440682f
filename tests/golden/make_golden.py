#!/usr/bin/env python
"""tests/golden/make_golden.py — generates tests/golden/forward_bwt_golden.npz from the UNMODIFIED reference
compiled here (oracle/_ref/libbwtc_ref.so, built by oracle/Makefile from /root/reference).

The reference ships no golden vectors for the forward BWT, so these fixtures ARE outputs of the
reference itself: BWTManager('d').doTransform(BWTBlock&, freqs) (bwtransforms/BWTManager.cpp:53-58), checked
to be identical for engine 's' (SA-IS) before being written.  Run in the build container only:
    python tests/golden/make_golden.py
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libbwtc_ref.so"))


def ref_block(x, starts, algo):
    buf = np.concatenate([x, np.array([0xAB], np.uint8)])
    LF = np.zeros(256, np.uint32)
    n = ctypes.c_uint32(0)
    fr = np.zeros(256, np.uint32)
    ref.ref_bwt_block(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(x.size), ctypes.c_uint(starts),
                      ctypes.c_char(algo), LF.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n),
                      fr.ctypes.data_as(ctypes.c_void_p))
    assert buf[-1] == 0xAB, "the byte after the block must be preserved"
    return buf[:-1].copy(), LF[: n.value].copy(), fr


def lcg_bytes(n, seed=12345):
    out = np.empty(n, np.uint8)
    x = seed
    for i in range(n):
        x = (1103515245 * x + 12345) % (1 << 31)
        out[i] = (x >> 16) & 0xFF
    return out


def main():
    rng = np.random.default_rng(20261018)
    cases = {}
    for s in [b"mississippi", b"banana", b"abracadabra", b"aaaaaaaa", b"a", b"ab", b"ba"]:
        cases["str_" + s.decode()] = (np.frombuffer(s, np.uint8).copy(), 8)
    cases["ab_x150"] = (np.frombuffer(b"ab" * 150, np.uint8).copy(), 8)
    cases["zeros_300"] = (np.zeros(300, np.uint8), 8)
    cases["mod251_1000"] = (((7 * np.arange(1000)) % 251).astype(np.uint8), 8)
    cases["lcg_4096"] = (lcg_bytes(4096), 8)
    for n, sigma, starts in [(257, 2, 256), (510, 3, 256), (1000, 1, 7), (1000, 4, 1), (3000, 256, 2), (5000, 16, 30),
                             (20000, 4, 8), (20000, 64, 256), (33333, 2, 8), (65536, 256, 8)]:
        cases[f"rand_n{n}_s{sigma}_k{starts}"] = (rng.integers(0, sigma, n).astype(np.uint8), starts)
    # data with many 0x00 bytes and a long run (sentinel collisions, deep doubling)
    x = rng.integers(0, 3, 8000).astype(np.uint8)
    x[1000:5000] = 0
    cases["zero_run_8000"] = (x, 8)
    tile = rng.integers(0, 256, 512).astype(np.uint8)
    x = np.tile(tile, 40)
    x[rng.integers(0, x.size, 20)] = rng.integers(0, 256, 20).astype(np.uint8)
    cases["tiled_512x40_mut"] = (x, 8)
    out = {}
    for name, (x, starts) in cases.items():
        d = ref_block(x, starts, b"d")
        s = ref_block(x, starts, b"s")
        assert (d[0] == s[0]).all() and (d[1] == s[1]).all() and (d[2] == s[2]).all(), name
        out[name + "/in"] = x
        out[name + "/starts"] = np.array([starts], np.uint32)
        out[name + "/out"] = d[0]
        out[name + "/LF"] = d[1]
        out[name + "/freqs"] = d[2]
    np.savez_compressed(os.path.join(HERE, "forward_bwt_golden.npz"), **out)
    print("wrote", len(cases), "cases")
    for k in ["str_mississippi", "str_banana", "str_abracadabra", "ab_x150", "zeros_300", "mod251_1000", "lcg_4096"]:
        print(k, bytes(out[k + "/out"][:16]), out[k + "/LF"])


if __name__ == "__main__":
    main()
