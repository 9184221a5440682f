#!/usr/bin/env python
"""tests/integration_block_check.py — run in a fresh process by tests/test_gpu_integration.py: drives the block-level entry
points of the INTEGRATED build (the reference's BWTManager / giveTransformer, patched, with choice 'c' =
bwtc::CudaBWTransform) and compares every result with the C oracle.  Exit code 0 = all equal."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bwtc_b200 as bw  # noqa: E402
from conftest import Oracle  # noqa: E402

lib = ctypes.CDLL(bw.INTEGRATION_LIB_PATH)
orc = Oracle()
err = ctypes.create_string_buffer(1024)
rng = np.random.default_rng(41)


def call_block(fn, x, starts, choice=b"c"):
    buf = np.concatenate([x, np.array([0xCD], np.uint8)])
    LF = np.zeros(256, np.uint32)
    k = ctypes.c_uint(0)
    fr = np.zeros(256, np.uint32)
    rc = fn(ctypes.c_void_p(buf.ctypes.data), ctypes.c_uint(x.size), ctypes.c_uint(starts), ctypes.c_char(choice),
            ctypes.c_void_p(LF.ctypes.data), ctypes.byref(k), ctypes.c_void_p(fr.ctypes.data), err, ctypes.c_uint(1024))
    assert rc == 0, err.value
    assert buf[-1] == 0xCD, "byte after the block must be preserved"
    return buf[:-1], LF[: k.value], fr


cases = [(1, 2, 8), (2, 2, 1), (300, 4, 8), (70000, 64, 8), (200000, 256, 256), (5000, 1, 3), (1 << 20, 64, 8), (255, 3, 8), (257, 3, 8)]
for n, sigma, starts in cases:
    x = rng.integers(0, sigma, n).astype(np.uint8)
    want = orc.block(x, starts)
    # (1) patched BWTManager, choice 'c': fused device path (INTEGRATION.md option C)
    # (2) giveTransformer('c') + the reference's NON-virtual base wrapper: host reverse around the raw virtual (option B)
    for name, fn in (("manager", lib.b200_manager_block), ("base wrapper", lib.b200_base_wrapper_block)):
        got = call_block(fn, x, starts)
        assert (got[0] == want[0]).all() and (got[1] == want[1]).all() and (got[2] == want[2]).all(), (name, n, sigma, starts)
    # the reference's own engines through the same patched manager still answer the same
    if n <= 70000:
        got = call_block(lib.b200_manager_block, x, starts, b"d")
        assert (got[0] == want[0]).all() and (got[1] == want[1]).all() and (got[2] == want[2]).all(), ("d", n)

# (3) raw virtual on a caller-prepared buffer, as test/InverseBwtTest.cpp:57-66 calls it
for trial in range(20):
    n = int(rng.integers(2, 50000))
    T = rng.integers(0, int(rng.choice([2, 4, 256])), n).astype(np.uint8)
    T[-1] = 0
    nLF = int(rng.integers(1, min(n, 256) + 1))
    rc, wbuf, wLF, wfr = orc.raw(T, nLF)
    U = T.copy()
    LF = np.zeros(nLF, np.uint32)
    fr = np.zeros(256, np.uint32)
    r = lib.b200_transformer_raw(ctypes.c_void_p(U.ctypes.data), ctypes.c_uint(n), ctypes.c_uint(nLF), ctypes.c_char(b"c"),
                                 ctypes.c_void_p(LF.ctypes.data), ctypes.c_void_p(fr.ctypes.data), err, ctypes.c_uint(1024))
    assert r == 0, err.value
    assert (U == wbuf).all() and (LF == wLF).all() and (fr == wfr).all(), (trial, n, nLF)

# (4) all slices of one precompressor block in one call (batched on the device)
sizes = [1 << 16] * 9 + [4321, 300, 300, 1]
blocks = [bw.generate(["markov", "dna", "random"][i % 3], n, seed=900 + i) for i, n in enumerate(sizes)]
work = [b.copy() for b in blocks]
count = len(work)
ptrs = (ctypes.c_void_p * count)(*[w.ctypes.data for w in work])
sz = np.array(sizes, np.uint32)
LF = np.zeros((count, 256), np.uint32)
nLF = np.zeros(count, np.uint32)
fr = np.zeros((count, 256), np.uint32)
rc = lib.b200_fused_blocks(ptrs, ctypes.c_void_p(sz.ctypes.data), ctypes.c_uint(count), ctypes.c_uint(8), ctypes.c_void_p(LF.ctypes.data),
                           ctypes.c_void_p(nLF.ctypes.data), ctypes.c_void_p(fr.ctypes.data), err, ctypes.c_uint(1024))
assert rc == 0, err.value
for i, x in enumerate(blocks):
    w = orc.block(x, 8)
    assert (work[i] == w[0]).all() and nLF[i] == w[1].size and (LF[i, : nLF[i]] == w[1]).all() and (fr[i] == w[2]).all(), i
print("integration block check ok")
