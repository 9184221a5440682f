"""CPU tests of bwtc::PipelinedCompressor (bwtc_b200/host/PipelinedCompressor.cpp): the batched look-ahead replacement
of Compressor::compress (Compressor.cpp:65-120) must write byte-identical .bwtc files.  Here the BWT choice is one of the
reference's own CPU engines ('d', 's') — they live in the linked reference objects — so the reader / parallel encoder /
ordered writer logic, the sharded part files and the merge are checked without a GPU; tests/test_gpu_integration.py
repeats the same checks with choice 'c' (the product path) on the B200."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import bwtc_b200 as bw
from conftest import ROOT

TOOL = os.path.join(ROOT, "tests", "integration_tool.py")
REFTOOL = os.path.join(ROOT, "tests", "bwtc_file_tool.py")


def run_tool(*args, tool=TOOL, timeout=900):
    r = subprocess.run([sys.executable, tool, *[str(a) for a in args]], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    return r.stdout.strip()


def need_libs():
    for p in (os.path.join(ROOT, "bwtc_b200", "libbwtc_integration.so"), os.path.join(ROOT, "oracle", "_ref", "libbwtc_ref.so")):
        if not os.path.exists(p):
            pytest.skip(f"{os.path.relpath(p, ROOT)} not built")


def mixed_input(nbytes, seed):
    """Markov text with stretches of the other families, so blocks differ in alphabet and compressibility."""
    parts, left, i = [], nbytes, 0
    while left > 0:
        k = min(left, (1 << 18) + 12345 * (i % 5))
        parts.append(bw.generate(["markov", "dna", "markov", "repetitive", "random"][i % 5], k, seed=seed + i))
        left -= k
        i += 1
    return np.concatenate(parts)


@pytest.mark.parametrize("coder,choice,mem,threads,prepr", [
    ("H", "d", 5667979, 4, ""),       # BASELINE config 1 shape: 1 MiB blocks, Huffman, 4 encoder threads
    ("H", "s", 5667979, 3, ""),       # SA-IS engine
    ("H", "d", 1000000, 8, ""),       # 185 000-byte blocks: many small blocks, more threads than cores matter
    ("B", "d", 5667979, 4, ""),       # wavelet coder: state leaks across blocks -> one encoder thread, in order
    ("H", "d", 5667979, 4, "pp"),     # preprocessing: several BWT slices per precompressor block (chained on one thread)
    ("m", "d", 3000000, 2, "p"),
])
def test_pipelined_compress_is_byte_identical_to_reference(tmp_path, coder, choice, mem, threads, prepr):
    need_libs()
    x = mixed_input(5 << 20, seed=3)
    src = tmp_path / "in.bin"
    x.tofile(src)
    # the unmodified reference (separate process: same C++ symbols)
    if prepr == "":
        run_tool("compress", "cpu", src, tmp_path / "ref.bwtc", mem, coder, 8, tool=REFTOOL)
    else:  # the reference driver of oracle/_ref has no preprocessing argument: the integrated build's sync Compressor is the same code
        run_tool("sync_compress", src, tmp_path / "ref.bwtc", mem, coder, choice, 8, prepr)
    out = json.loads(run_tool("pipe_compress", src, tmp_path / "pipe.bwtc", mem, coder, choice, 8, threads, 0, "", 1, 0, 1, prepr))
    a = (tmp_path / "ref.bwtc").read_bytes()
    b = (tmp_path / "pipe.bwtc").read_bytes()
    assert out["rc"] == len(b) == len(a)
    assert a == b, "pipelined .bwtc differs from the reference's"
    assert out["timings"]["input_bytes"] == x.size
    run_tool("uncompress", "cpu", tmp_path / "pipe.bwtc", tmp_path / "back.bin", tool=REFTOOL)
    assert (tmp_path / "back.bin").read_bytes() == x.tobytes()


def test_sync_compressor_of_integrated_build_equals_reference(tmp_path):
    """The integrated library contains the reference's Compressor with the PATCHED BWTManager: choices 'd' and 's' must
    behave exactly as before the patch."""
    need_libs()
    x = mixed_input(3 << 20, seed=9)
    src = tmp_path / "in.bin"
    x.tofile(src)
    run_tool("compress", "cpu", src, tmp_path / "ref.bwtc", 5667979, "H", 8, tool=REFTOOL)
    for choice in ("d", "s", "a"):
        run_tool("sync_compress", src, tmp_path / f"{choice}.bwtc", 5667979, "H", choice, 8)
        assert (tmp_path / f"{choice}.bwtc").read_bytes() == (tmp_path / "ref.bwtc").read_bytes(), choice


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_parts_merge_to_the_reference_file(tmp_path, world):
    """One process per GPU: rank r compresses precompressor blocks r, r+world, ... into a part file; merging the parts
    gives the reference's bytes (no data exchange before the final concat)."""
    need_libs()
    x = mixed_input((4 << 20) + 777, seed=5)
    src = tmp_path / "in.bin"
    x.tofile(src)
    run_tool("compress", "cpu", src, tmp_path / "ref.bwtc", 5667979, "H", 8, tool=REFTOOL)
    parts = []
    for r in range(world):
        p = tmp_path / f"part{r}"
        run_tool("pipe_compress", src, p, 5667979, "H", "d", 8, 2, 0, "", 1, r, world)
        parts.append(p)
    out = json.loads(run_tool("merge_parts", tmp_path / "merged.bwtc", "H", *parts))
    a = (tmp_path / "ref.bwtc").read_bytes()
    b = (tmp_path / "merged.bwtc").read_bytes()
    assert out["rc"] == len(b)
    assert a == b


def test_empty_and_tiny_inputs(tmp_path):
    need_libs()
    for n in (0, 1, 255, 257):
        x = np.arange(n, dtype=np.uint8)
        src = tmp_path / f"in{n}.bin"
        x.tofile(src)
        run_tool("compress", "cpu", src, tmp_path / "ref.bwtc", 5667979, "H", 8, tool=REFTOOL)
        run_tool("pipe_compress", src, tmp_path / "pipe.bwtc", 5667979, "H", "d", 8, 2, 0, "", 1)
        assert (tmp_path / "ref.bwtc").read_bytes() == (tmp_path / "pipe.bwtc").read_bytes(), n


def test_choice_c_is_valid_in_the_patched_manager_and_fails_loudly_without_a_gpu(tmp_path):
    """BWTManager::isValidChoice('c') (patched, BWTManager.cpp:70-72); with no CUDA device the transform throws —
    there is no CPU fallback behind choice 'c'."""
    need_libs()
    import ctypes
    code = ("import ctypes,sys; l=ctypes.CDLL(%r); print(l.b200_is_valid_choice(ctypes.c_char(b'c')), "
            "l.b200_is_valid_choice(ctypes.c_char(b'd')), l.b200_is_valid_choice(ctypes.c_char(b'x')))"
            % os.path.join(ROOT, "bwtc_b200", "libbwtc_integration.so"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.stdout.split() == ["1", "1", "0"], (r.stdout, r.stderr)
    from conftest import has_cuda
    if has_cuda():
        pytest.skip("a CUDA device is present: the failure path cannot be shown here")
    x = bw.generate("markov", 300000, seed=1)
    src = tmp_path / "in.bin"
    x.tofile(src)
    r = subprocess.run([sys.executable, TOOL, "pipe_compress", str(src), str(tmp_path / "o.bwtc"), "5667979", "H", "c", "8", "2",
                        "0", "0", "2"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "rc" in r.stdout and json.loads(r.stdout)["rc"] < 0 and json.loads(r.stdout)["err"] != ""
