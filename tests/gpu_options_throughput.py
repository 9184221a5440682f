#!/usr/bin/env python
"""BWT-stage throughput through every binding a bwtc maintainer can choose, synchronous, one block at
a time, pageable (malloc'ed) 32 MiB Markov blocks, as Compressor::compress would call it (not a pytest):
  A  unmodified BWTManager('d') -> base wrapper (host std::reverse) -> Divsufsorter -> divbwtf = link-time CUDA shim
  B  giveTransformer('c') + base wrapper (host std::reverse) -> CudaBWTransform raw virtual
  C  patched BWTManager('c') -> CudaBWTransform::doTransformFused (reverse / sentinel / hole fill on the device)
  ref the unmodified reference on one core, for scale.
Each mode runs in its own process (the libraries export the same C++ symbols).  python tests/gpu_options_throughput.py"""
import ctypes, json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def child(mode, nblocks):
    import bwtc_b200 as bw
    n = 32 << 20
    blocks = [np.concatenate([bw.generate("markov", n, seed=300 + i), np.zeros(1, np.uint8)]) for i in range(nblocks)]
    LF = np.zeros(256, np.uint32); k = ctypes.c_uint(0); fr = np.zeros(256, np.uint32); err = ctypes.create_string_buffer(512)
    if mode in ("A", "ref"):
        lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libbwtc_ref_cuda.so" if mode == "A" else "libbwtc_ref.so"))
        def run(b):
            lib.ref_bwt_block(ctypes.c_void_p(b.ctypes.data), ctypes.c_uint(n), ctypes.c_uint(8), ctypes.c_char(b"d"),
                              ctypes.c_void_p(LF.ctypes.data), ctypes.byref(k), ctypes.c_void_p(fr.ctypes.data))
    else:
        lib = ctypes.CDLL(bw.INTEGRATION_LIB_PATH)
        lib.b200_manager_new.restype = ctypes.c_void_p
        lib.b200_transformer_new.restype = ctypes.c_void_p
        if mode == "B":  # a transformer kept across blocks, the reference's non-virtual base wrapper around its raw virtual
            h = ctypes.c_void_p(lib.b200_transformer_new(ctypes.c_char(b"c")))
            def run(b):
                rc = lib.b200_transformer_base_wrapper(h, ctypes.c_void_p(b.ctypes.data), ctypes.c_uint(n), ctypes.c_uint(8),
                                                       ctypes.c_void_p(LF.ctypes.data), ctypes.byref(k), ctypes.c_void_p(fr.ctypes.data),
                                                       err, ctypes.c_uint(512))
                assert rc == 0, err.value
        else:            # the patched BWTManager kept across blocks, as Compressor's m_bwtmanager is
            h = ctypes.c_void_p(lib.b200_manager_new(ctypes.c_char(b"c"), ctypes.c_uint(8)))
            def run(b):
                rc = lib.b200_manager_transform(h, ctypes.c_void_p(b.ctypes.data), ctypes.c_uint(n), ctypes.c_void_p(LF.ctypes.data),
                                                ctypes.byref(k), ctypes.c_void_p(fr.ctypes.data), err, ctypes.c_uint(512))
                assert rc == 0, err.value
    run(blocks[0].copy())  # warm-up: context creation
    t0 = time.perf_counter()
    for b in blocks:
        run(b)
    dt = time.perf_counter() - t0
    print(json.dumps({"mode": mode, "blocks": nblocks, "MBps": nblocks * n / 1e6 / dt, "ms_per_block": 1e3 * dt / nblocks}))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1], int(sys.argv[2]))
    else:
        for mode, nb in (("A", 12), ("B", 12), ("C", 12), ("ref", 1)):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), mode, str(nb)], capture_output=True, text=True)
            print(r.stdout.strip() or r.stderr[-500:], flush=True)
