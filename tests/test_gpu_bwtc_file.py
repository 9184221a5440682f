"""GPU test, BASELINE.json config 1: `bwtc compress` of 16 MiB order-2 Markov text with 1 MiB blocks and the
Huffman coder.  The reference's own Compressor / HuffmanEncoder / BWTManager / Divsufsorter run UNCHANGED; only
divsufsort.c is replaced at link time by bwtc_b200/host/divsufsort_shim.cpp (-> C-ABI -> sm_100a kernels).
The .bwtc must be byte-identical to the CPU-only reference's and the reference Decompressor must round-trip it.
(INTEGRATION.md option A; options B-D — the real bwtc::CudaBWTransform subclass, the patched BWTManager and the
pipelined compressor — are covered by tests/test_gpu_integration.py.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

import bwtc_b200 as bw
from conftest import ROOT

pytestmark = pytest.mark.gpu
TOOL = os.path.join(ROOT, "tests", "bwtc_file_tool.py")


def _run(*args):
    r = subprocess.run([sys.executable, TOOL, *[str(a) for a in args]], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip()


def _need_ref_builds():
    for f in ("libbwtc_ref.so", "libbwtc_ref_cuda.so"):
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", f)):
            pytest.skip(f"oracle/_ref/{f} not built")


@pytest.mark.parametrize("coder,mib,mem", [("H", 16, 5667979), ("H", 3, 90687655), ("B", 2, 5667979)])
def test_bwtc_file_byte_identical_and_roundtrip(tmp_path, coder, mib, mem):
    _need_ref_builds()
    src = tmp_path / "in.bin"
    x = bw.generate("markov", mib << 20, seed=77)
    x.tofile(src)
    _run("compress", "cpu", src, tmp_path / "ref.bwtc", mem, coder, 8)
    _run("compress", "cuda", src, tmp_path / "gpu.bwtc", mem, coder, 8)
    a = (tmp_path / "ref.bwtc").read_bytes()
    b = (tmp_path / "gpu.bwtc").read_bytes()
    assert len(a) > 1000 and a[0:1] == coder.encode()
    assert a == b, "GPU-built .bwtc differs from the reference's"
    _run("uncompress", "cpu", tmp_path / "gpu.bwtc", tmp_path / "back.bin")
    assert (tmp_path / "back.bin").read_bytes() == x.tobytes()
