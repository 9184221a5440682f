"""SURVEY.md §8f row f3, host side (bwtc_b200/host/RunStatistics.cpp), CPU only: the registry + link-time wrapper around
utils::calculateRunFrequenciesAndStoreRuns (Utils.cpp:150-170) must return, for every section the Huffman coder asks for
(HuffmanCoders.cpp:143), exactly what the reference's own scan returns — run symbols, run lengths, run frequencies —
when it is fed the maximal runs of the whole block, the form the GPU emits them in (bwtc_cuda_runs)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import bwtc_b200 as bw
from conftest import ROOT

CODE = r"""
import ctypes, sys
import numpy as np
lib = ctypes.CDLL(%r)
rng = np.random.default_rng(5)
def check(data, sections):
    data = np.ascontiguousarray(data, np.uint8)
    sec = np.array(sections, np.uint32)
    assert sec.sum() == data.size
    rc = lib.b200_test_run_slicing(ctypes.c_void_p(data.ctypes.data), ctypes.c_uint(data.size), ctypes.c_void_p(sec.ctypes.data), ctypes.c_uint(sec.size))
    assert rc == 0, (rc, data.size, sections[:8])
for n in (1, 2, 3, 100, 10000, 200001):
    for sigma in (1, 2, 3, 256):
        x = rng.integers(0, sigma, n).astype(np.uint8)
        if sigma == 3:
            x = np.repeat(x, 7)[:n]  # long runs, many of them cut by section boundaries
        cuts = sorted(set(rng.integers(0, n + 1, min(n, 40)).tolist()) | {0, n})
        lens = [b - a for a, b in zip(cuts[:-1], cuts[1:])]
        check(x, lens)
        check(x, [n])
check(np.zeros(50000, np.uint8), [1, 49998, 1])
check(np.repeat(np.arange(256, dtype=np.uint8), 300), [10000] * 7 + [6800])
print("run slicing ok")
"""


def test_wrapped_run_scan_equals_reference_scan():
    p = os.path.join(ROOT, "bwtc_b200", "libbwtc_integration.so")
    if not os.path.exists(p):
        pytest.skip("bwtc_b200/libbwtc_integration.so not built")
    r = subprocess.run([sys.executable, "-c", CODE % p], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "run slicing ok" in r.stdout, (r.stdout[-1000:], r.stderr[-3000:])
