"""GPU test, BASELINE.json config 1: `bwtc compress` of 16 MiB order-2 Markov text with 1 MiB blocks and the
Huffman coder.  The reference's own Compressor / HuffmanEncoder / BWTManager / Divsufsorter run UNCHANGED; only
divsufsort.c is replaced at link time by bwtc_b200/host/divsufsort_shim.cpp (-> C-ABI -> sm_100a kernels).
The .bwtc must be byte-identical to the CPU-only reference's and the reference Decompressor must round-trip it.
Also drives the C++ mirror classes of bwtc_b200/host/ (libbwtc_host.so)."""
import ctypes
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


def test_cpp_mirror_classes(oracle):
    lib = ctypes.CDLL(bw.HOST_LIB_PATH)
    rng = np.random.default_rng(41)
    for n, sigma, starts in [(1, 2, 8), (300, 4, 8), (70000, 64, 8), (200000, 256, 256), (5000, 1, 3)]:
        x = rng.integers(0, sigma, n).astype(np.uint8)
        want = oracle.block(x, starts)
        for via_base in (0, 1):  # fused device path / reference host-side wrapper around the raw virtual
            buf = np.concatenate([x, np.array([0xCD], np.uint8)])
            LF = np.zeros(256, np.uint32)
            k = ctypes.c_uint(0)
            fr = np.zeros(256, np.uint32)
            err = ctypes.create_string_buffer(512)
            rc = lib.bwtc_host_manager_transform(ctypes.c_void_p(buf.ctypes.data), ctypes.c_uint(n), ctypes.c_uint(starts),
                                                 ctypes.c_int(via_base), ctypes.c_void_p(LF.ctypes.data), ctypes.byref(k),
                                                 ctypes.c_void_p(fr.ctypes.data), err, ctypes.c_uint(512))
            assert rc == 0, err.value
            assert buf[-1] == 0xCD
            assert (buf[:-1] == want[0]).all() and (LF[: k.value] == want[1]).all() and (fr == want[2]).all(), (n, via_base)
    assert lib.bwtc_host_is_valid_choice(ctypes.c_char(b"c")) == 1
    assert lib.bwtc_host_is_valid_choice(ctypes.c_char(b"d")) == 0


def test_cpp_mirror_batched_manager(oracle):
    """bwtc_b200::BWTManager::doTransform(std::vector<BWTBlock*>&, freqs): the slices of one precompressor block in one
    call (batched on the device) give every block what the single-block call gives it."""
    lib = ctypes.CDLL(bw.HOST_LIB_PATH)
    sizes = [1 << 16] * 9 + [4321, 300, 300, 1]
    blocks = [bw.generate(["markov", "dna", "random"][i % 3], n, seed=900 + i) for i, n in enumerate(sizes)]
    work = [b.copy() for b in blocks]
    count = len(work)
    ptrs = (ctypes.c_void_p * count)(*[w.ctypes.data for w in work])
    sz = np.array(sizes, np.uint32)
    LF = np.zeros((count, 256), np.uint32)
    nLF = np.zeros(count, np.uint32)
    fr = np.zeros((count, 256), np.uint32)
    err = ctypes.create_string_buffer(512)
    rc = lib.bwtc_host_manager_transform_batch(ptrs, ctypes.c_void_p(sz.ctypes.data), ctypes.c_uint(count), ctypes.c_uint(8),
                                               ctypes.c_void_p(LF.ctypes.data), ctypes.c_void_p(nLF.ctypes.data),
                                               ctypes.c_void_p(fr.ctypes.data), err, ctypes.c_uint(512))
    assert rc == 0, err.value
    for i, x in enumerate(blocks):
        w = oracle.block(x, 8)
        assert (work[i] == w[0]).all(), i
        assert nLF[i] == w[1].size and (LF[i, : nLF[i]] == w[1]).all(), i
        assert (fr[i] == w[2]).all(), i
