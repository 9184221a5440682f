#!/usr/bin/env python
"""tests/bwtc_file_tool.py — runs the reference's Compressor / Decompressor from one of the two test builds in a
fresh process (the two .so files export the same C++ symbols, so they are never loaded together):
    bwtc_file_tool.py compress   <cpu|cuda> <in> <out> <memLimit> [coder] [starts]
    bwtc_file_tool.py uncompress <cpu|cuda> <in> <out>
cpu  = oracle/_ref/libbwtc_ref.so       (unmodified reference, divsufsort on the CPU)
cuda = oracle/_ref/libbwtc_ref_cuda.so  (same objects, divsufsort.c replaced at link time by the CUDA shim)"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    op, which = sys.argv[1], sys.argv[2]
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libbwtc_ref.so" if which == "cpu" else "libbwtc_ref_cuda.so"))
    lib.ref_compress.restype = ctypes.c_longlong
    lib.ref_uncompress.restype = ctypes.c_longlong
    if op == "compress":
        coder = (sys.argv[6] if len(sys.argv) > 6 else "H").encode()
        starts = int(sys.argv[7]) if len(sys.argv) > 7 else 8
        r = lib.ref_compress(sys.argv[3].encode(), sys.argv[4].encode(), ctypes.c_ulonglong(int(sys.argv[5])),
                             ctypes.c_char(coder), ctypes.c_char(b"d"), ctypes.c_uint(starts))
    else:
        r = lib.ref_uncompress(sys.argv[3].encode(), sys.argv[4].encode())
    print(r)


if __name__ == "__main__":
    main()
