"""CPU tests of the N > 1 host-side path (block sharding, ordered metadata gather) with world_size 2 over gloo.
The transform is injected; here the C oracle stands in for the GPU (tests may call oracle/, the product may not)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from bwtc_b200 import sharding
from conftest import ROOT


def test_block_assignment_and_slicing():
    assert sharding.blocks_for_rank(10, 0, 4) == [0, 4, 8]
    assert sharding.blocks_for_rank(10, 3, 4) == [3, 7]
    assert sorted(sum((sharding.blocks_for_rank(257, r, 8) for r in range(8)), [])) == list(range(257))
    assert sharding.slice_blocks(10, 4) == [(0, 4), (4, 4), (8, 2)]
    assert sharding.slice_blocks(8 << 30, 32 << 20)[-1] == ((8 << 30) - (32 << 20), 32 << 20)
    assert len(sharding.slice_blocks(8 << 30, 32 << 20)) == 256
    with pytest.raises(ValueError):
        sharding.blocks_for_rank(4, 2, 2)


WORKER = textwrap.dedent("""
    import os, sys, json, ctypes
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import bwtc_b200 as bw
    from bwtc_b200 import sharding
    from conftest import Oracle
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    total = 7
    sizes = [5000, 1, 300, 70000, 4096, 257, 12345]
    def make(i):
        return bw.generate(["markov", "dna", "repetitive", "random"][i %% 4], sizes[i], seed=50 + i)
    def transform(i, blk):          # stand-in for the GPU pipeline: same contract, in place
        out, LF, fr = orc.block(blk.copy(), 8)
        blk[:] = out
        return LF, fr
    mine = {i: make(i) for i in sharding.blocks_for_rank(total, rank, world)}
    merged = sharding.run_sharded(mine, total, rank, world, transform, dist)
    if rank == 0:
        print("RESULT " + json.dumps([[m["index"], m["owner"], m["size"], m["crc32"], m["LF"]] for m in merged]))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_world_size_2_gloo(tmp_path, oracle):
    import json
    import zlib

    import bwtc_b200 as bw

    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0]
    merged = json.loads(line[7:])
    sizes = [5000, 1, 300, 70000, 4096, 257, 12345]
    assert [m[0] for m in merged] == list(range(7))
    assert [m[1] for m in merged] == [i % 2 for i in range(7)]
    for i, (idx, owner, size, crc, LF) in enumerate(merged):
        x = bw.generate(["markov", "dna", "repetitive", "random"][i % 4], sizes[i], seed=50 + i)
        out, wLF, _ = oracle.block(x, 8)
        assert size == sizes[i] and crc == (zlib.crc32(out.tobytes()) & 0xFFFFFFFF) and LF == wLF.tolist()
