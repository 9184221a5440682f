"""GPU tests of the integrated build (INTEGRATION.md options B, C, D): bwtc::CudaBWTransform — a real subclass of the
reference's bwtc::BWTransform, compiled against the reference's headers — reached through the reference's BWTManager
(patched: one new choice character 'c'), through giveTransformer('c'), through the reference's synchronous Compressor,
and through bwtc::PipelinedCompressor (batched look-ahead, parallel CPU entropy coding overlapped with the GPU).
Every .bwtc must be byte-identical to the unmodified CPU reference's and round-trip through the reference Decompressor."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import bwtc_b200 as bw
from conftest import ROOT
from test_pipelined_compressor import REFTOOL, TOOL, mixed_input, need_libs, run_tool

pytestmark = pytest.mark.gpu


def test_block_level_entry_points_of_the_integrated_build():
    need_libs()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "integration_block_check.py")], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "integration block check ok" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


@pytest.mark.parametrize("coder,mib,mem", [("H", 16, 5667979), ("B", 2, 5667979), ("H", 40, 90687655)])
def test_reference_compressor_with_choice_c_is_byte_identical(tmp_path, coder, mib, mem):
    """The reference's own Compressor::compress (synchronous), BWT choice 'c': BASELINE config 1 (16 MiB, 1 MiB blocks,
    Huffman), the wavelet coder, and 16 MiB blocks."""
    need_libs()
    x = bw.generate("markov", mib << 20, seed=78)
    src = tmp_path / "in.bin"
    x.tofile(src)
    run_tool("compress", "cpu", src, tmp_path / "ref.bwtc", mem, coder, 8, tool=REFTOOL)
    run_tool("sync_compress", src, tmp_path / "gpu.bwtc", mem, coder, "c", 8)
    assert (tmp_path / "ref.bwtc").read_bytes() == (tmp_path / "gpu.bwtc").read_bytes()
    run_tool("uncompress", "cpu", tmp_path / "gpu.bwtc", tmp_path / "back.bin", tool=REFTOOL)
    assert (tmp_path / "back.bin").read_bytes() == x.tobytes()


@pytest.mark.parametrize("coder,mem,threads,depth,prepr,mib", [
    ("H", 5667979, 8, 3, "", 16),       # BASELINE config 1: 16 x 1 MiB blocks (batched on the device), 8 encoder threads
    ("H", 181375309, 8, 3, "", 200),    # 32 MiB blocks (BASELINE config 5 block size), last block short
    ("B", 5667979, 4, 2, "", 3),        # wavelet coder: one encoder thread, in order, GPU look-ahead still on
    ("H", 20000000, 6, 2, "pp", 24),    # preprocessing: several BWT slices per precompressor block, all prefetched
])
def test_pipelined_compress_on_gpu_is_byte_identical(tmp_path, coder, mem, threads, depth, prepr, mib):
    need_libs()
    x = mixed_input(mib << 20, seed=11) if mib <= 24 else bw.generate("markov", (mib << 20) + 4099, seed=12)
    src = tmp_path / "in.bin"
    x.tofile(src)
    if prepr == "":
        run_tool("compress", "cpu", src, tmp_path / "ref.bwtc", mem, coder, 8, tool=REFTOOL)
    else:
        run_tool("sync_compress", src, tmp_path / "ref.bwtc", mem, coder, "d", 8, prepr)
    out = json.loads(run_tool("pipe_compress", src, tmp_path / "pipe.bwtc", mem, coder, "c", 8, threads, 0, "0", depth, 0, 1, prepr))
    a = (tmp_path / "ref.bwtc").read_bytes()
    b = (tmp_path / "pipe.bwtc").read_bytes()
    assert out["rc"] == len(b)
    assert a == b, "GPU-pipelined .bwtc differs from the reference's"
    run_tool("uncompress", "cpu", tmp_path / "pipe.bwtc", tmp_path / "back.bin", tool=REFTOOL)
    assert (tmp_path / "back.bin").read_bytes() == x.tobytes()


def test_pipelined_compress_1gib_32mib_blocks(tmp_path):
    """BASELINE config 5 at 1 GiB: 32 blocks of 32 MiB through the look-ahead pipeline with all host cores coding;
    the reference (one core, ~1.5 min) writes the file to compare with."""
    need_libs()
    n = 1 << 30
    x = np.empty(n, np.uint8)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(lambda i: bw.generate("markov", 32 << 20, seed=2000 + i, out=x[i << 25:(i + 1) << 25]), range(32)))
    src = tmp_path / "in.bin"
    x.tofile(src)
    mem = 181375309  # floor(mem * 0.185) = 32 MiB (Compressor.cpp:78)
    run_tool("compress", "cpu", src, tmp_path / "ref.bwtc", mem, "H", 8, tool=REFTOOL, timeout=1500)
    out = json.loads(run_tool("pipe_compress", src, tmp_path / "pipe.bwtc", mem, "H", "c", 8, os.cpu_count() or 4, 0, "0", 4))
    assert out["timings"]["bwt_blocks"] == 32 and out["timings"]["input_bytes"] == n
    r = subprocess.run(["cmp", str(tmp_path / "ref.bwtc"), str(tmp_path / "pipe.bwtc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    print("1 GiB pipelined compress: %.2f s (%.0f MB/s), encoder busy %.1f core-s on %d threads" % (
        out["timings"]["total"], n / 1e6 / out["timings"]["total"], out["timings"]["encoder_busy_sum"], out["timings"]["encoder_threads"]))


def test_sharded_gpu_parts_merge_to_the_reference_file(tmp_path):
    """One process per GPU (here: two ranks sharing the one GPU of the test box): each rank compresses its share of the
    blocks into a part file; the merged file equals the reference's."""
    need_libs()
    x = bw.generate("markov", (9 << 20) + 5, seed=14)
    src = tmp_path / "in.bin"
    x.tofile(src)
    run_tool("compress", "cpu", src, tmp_path / "ref.bwtc", 5667979, "H", 8, tool=REFTOOL)
    parts = []
    for r in range(2):
        p = tmp_path / f"part{r}"
        run_tool("pipe_compress", src, p, 5667979, "H", "c", 8, 3, 0, "0", 2, r, 2)
        parts.append(p)
    run_tool("merge_parts", tmp_path / "merged.bwtc", "H", *parts)
    assert (tmp_path / "ref.bwtc").read_bytes() == (tmp_path / "merged.bwtc").read_bytes()


@pytest.mark.parametrize("coder,mem,mib", [("H", 5667979, 6), ("H", 90687655, 40), ("B", 5667979, 2)])
def test_reference_decompressor_with_cuda_inverse(tmp_path, coder, mem, mib, monkeypatch):
    """SURVEY.md §8f row f4 through the reference's own interface: Decompressor::decompress (Decompressor.cpp:58-94) with
    the patched giveInverseTransformer() returning bwtc::CudaInverseBWTransform (BWTC_CUDA_INVERSE=1) restores the input
    byte for byte from a .bwtc written by the UNMODIFIED reference."""
    need_libs()
    x = mixed_input(mib << 20, seed=21)
    src = tmp_path / "in.bin"
    x.tofile(src)
    run_tool("compress", "cpu", src, tmp_path / "ref.bwtc", mem, coder, 8, tool=REFTOOL)
    monkeypatch.setenv("BWTC_CUDA_INVERSE", "1")
    run_tool("uncompress", tmp_path / "ref.bwtc", tmp_path / "back.bin")
    assert (tmp_path / "back.bin").read_bytes() == x.tobytes()
