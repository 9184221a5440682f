"""CPU tests of the drop-in boundary: the C-ABI library loads and exports exactly what include/bwtc_cuda.h
declares; host-side logic (starting-point sizing, argument errors) behaves like the reference.  No compute
calls here — there is no GPU in this container."""
import ctypes
import os
import re

import numpy as np
import pytest

import bwtc_b200 as bw
from conftest import ROOT, has_cuda


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "bwtc_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bwtc_cuda_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(bw.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bwtc_cuda.h but not exported"
    bound = {n for n, _, _ in bw.C_ABI}
    assert bound == set(names), (bound ^ set(names))


def test_version_string():
    lib = bw.load_library()
    assert b"sm_100a" in lib.bwtc_cuda_version()


def test_stats_struct_layout_matches_binding():
    lib = bw.load_library()
    assert lib.bwtc_cuda_stats_sizeof() == ctypes.sizeof(bw.Stats)


def test_num_starting_points_matches_reference_rules(oracle):
    # BWTManager.cpp:60-64 clamp + BWTBlock.cpp:104-108 sizing
    for n in (1, 2, 255, 256, 257, 1000, 1 << 20):
        for starts in (0, 1, 2, 8, 255, 256, 257, 100000):
            assert bw.num_starting_points(n, starts) == oracle.lib.oracle_num_starting_points(n, starts)


def test_bwtblock_and_manager_mirror():
    blk = bw.BWTBlock(np.zeros(1000, np.uint8))
    blk.prepareLFpowers(8)
    assert blk.LFpowers().size == 8 and not blk.isTransformed()
    small = bw.BWTBlock(np.zeros(256, np.uint8))
    small.prepareLFpowers(8)
    assert small.LFpowers().size == 1
    m = bw.BWTManager(1000)
    assert m.getStartingPoints() == 256
    m.setStartingPoints(0)
    assert m.getStartingPoints() == 1
    assert bw.BWTManager.isValidChoice("c") and not bw.BWTManager.isValidChoice("d")
    with pytest.raises(ValueError):
        m.initialize("d")  # the CPU engines are deliberately not carried: no fallback


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(bw.BwtcCudaUnavailable):
        bw.load_library(str(tmp_path / "libnope.so"))


@pytest.mark.skipif(has_cuda(), reason="only meaningful without a GPU")
def test_no_gpu_means_error_not_fallback():
    with pytest.raises(bw.BwtcCudaError) as e:
        bw.CudaContext(1 << 16)
    assert e.value.code == -3


def test_product_never_imports_oracle():
    """The product path must not route through the oracle or any CPU fallback."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bwtc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".c", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                code = "\n".join(l for l in txt.splitlines() if "oracle" in l and not l.lstrip().startswith(("#", "//", "*", '"', "/*"))
                                 and "never" not in l and "ever calls" not in l)
                assert "liboracle" not in code and "oracle_bwt" not in code and "libbwtc_ref" not in code, (f, code)


def test_generators_are_deterministic():
    for kind in ("markov", "dna", "repetitive", "random"):
        a = bw.generate(kind, 10000, seed=3)
        b = bw.generate(kind, 10000, seed=3)
        c = bw.generate(kind, 10000, seed=4)
        assert (a == b).all() and not (a == c).all()
    assert set(np.unique(bw.generate("dna", 5000))) == set(b"ACGT")
    mk = bw.generate("markov", 50000)
    assert mk.min() >= 32 and mk.max() < 96
