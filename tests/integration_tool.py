#!/usr/bin/env python
"""tests/integration_tool.py — drives bwtc_b200/libbwtc_integration.so (the reference's objects + patched BWTManager +
bwtc::CudaBWTransform + bwtc::PipelinedCompressor) in a fresh process: it exports the same C++ symbols as
oracle/_ref/libbwtc_ref.so, so the two are never loaded together.  Prints one JSON line.

    integration_tool.py sync_compress  IN OUT MEM CODER CHOICE STARTS [PREPR]
    integration_tool.py pipe_compress  IN OUT MEM CODER CHOICE STARTS THREADS LOOKAHEAD DEVICES DEPTH [RANK WORLD [PREPR]]
    integration_tool.py merge_parts    OUT CODER PART...
    integration_tool.py uncompress     IN OUT
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    lib = ctypes.CDLL(os.path.join(ROOT, "bwtc_b200", "libbwtc_integration.so"))
    for f in ("b200_sync_compress_file", "b200_pipelined_compress_file", "b200_merge_parts", "b200_uncompress_file"):
        getattr(lib, f).restype = ctypes.c_longlong
    return lib


def main():
    lib = load()
    op = sys.argv[1]
    err = ctypes.create_string_buffer(1024)
    out = {}
    if op == "sync_compress":
        src, dst, mem, coder, choice, starts = sys.argv[2:8]
        prepr = sys.argv[8] if len(sys.argv) > 8 else ""
        r = lib.b200_sync_compress_file(src.encode(), dst.encode(), ctypes.c_ulonglong(int(mem)), ctypes.c_char(coder.encode()),
                                        ctypes.c_char(choice.encode()), ctypes.c_uint(int(starts)), prepr.encode(), err,
                                        ctypes.c_uint(1024))
    elif op == "pipe_compress":
        src, dst, mem, coder, choice, starts, threads, look, devs, depth = sys.argv[2:12]
        rank = int(sys.argv[12]) if len(sys.argv) > 12 else 0
        world = int(sys.argv[13]) if len(sys.argv) > 13 else 1
        prepr = sys.argv[14] if len(sys.argv) > 14 else ""
        dl = [int(d) for d in devs.split(",") if d != ""]
        arr = (ctypes.c_int * max(1, len(dl)))(*dl)
        tm = (ctypes.c_double * 10)()
        r = lib.b200_pipelined_compress_file(src.encode(), dst.encode(), ctypes.c_ulonglong(int(mem)), ctypes.c_char(coder.encode()),
                                             ctypes.c_char(choice.encode()), ctypes.c_uint(int(starts)), prepr.encode(),
                                             ctypes.c_uint(int(threads)), ctypes.c_uint(int(look)), arr, ctypes.c_uint(len(dl)),
                                             ctypes.c_int(int(depth)), ctypes.c_uint(rank), ctypes.c_uint(world), tm, err,
                                             ctypes.c_uint(1024))
        out["timings"] = dict(zip(("total", "reader_busy", "encoder_busy_sum", "writer_busy", "bwt_wait_sum", "pb_blocks",
                                   "bwt_blocks", "input_bytes", "encoder_threads"), list(tm)[:9]))
    elif op == "merge_parts":
        dst, coder = sys.argv[2:4]
        parts = sys.argv[4:]
        arr = (ctypes.c_char_p * len(parts))(*[p.encode() for p in parts])
        r = lib.b200_merge_parts(arr, ctypes.c_uint(len(parts)), dst.encode(), ctypes.c_char(coder.encode()), err, ctypes.c_uint(1024))
    elif op == "uncompress":
        r = lib.b200_uncompress_file(sys.argv[2].encode(), sys.argv[3].encode())
    else:
        raise SystemExit("unknown op " + op)
    try:
        lib.b200_run_statistics_served.restype = ctypes.c_size_t
        out["run_statistics_served"] = int(lib.b200_run_statistics_served())
    except AttributeError:
        pass
    out["rc"] = int(r)
    out["err"] = err.value.decode(errors="replace")
    print(json.dumps(out))
    return 0 if r >= 0 else 1


if __name__ == "__main__":
    sys.exit(main())
