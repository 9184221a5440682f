"""bwtc_b200/sharding.py — host-side block sharding for one-process-per-GPU runs.

Independent BWT blocks (Compressor.hpp:59-61, PrecompressorBlock.cpp:123-134) are dealt round-robin: block i goes
to rank i mod G.  There is NO collective on the data path; torch.distributed is only used to return per-block
metadata (LFpowers, byte histograms, checksums, timings) to rank 0 in FILE ORDER so that an entropy coder can
consume blocks strictly in order and the .bwtc stays byte-identical.
The transform itself is injected (`transform(block_index, block) -> (LF, freqs)`); in production it is
bwtc_b200.Pipeline / CudaBWTransform — this module never provides a CPU implementation.
"""
from __future__ import annotations

import zlib
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np


def blocks_for_rank(total_blocks: int, rank: int, world: int) -> List[int]:
    """Global block indices handled by `rank`: i with i mod world == rank."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, total_blocks, world))


def owner_of(block_index: int, world: int) -> int:
    return block_index % world


def slice_blocks(total_bytes: int, block_bytes: int) -> List[Tuple[int, int]]:
    """(offset, size) of every BWT block of a stream, as PrecompressorBlock::sliceIntoBlocks cuts it when no
    preprocessing is used (Compressor.cpp:77-81: one slice per precompressor block of bwtBlockSize bytes)."""
    if block_bytes < 1:
        raise ValueError("block_bytes must be positive")
    return [(off, min(block_bytes, total_bytes - off)) for off in range(0, total_bytes, block_bytes)]


def run_sharded(blocks: Dict[int, np.ndarray], total_blocks: int, rank: int, world: int,
                transform: Callable[[int, np.ndarray], Tuple[np.ndarray, np.ndarray]], dist=None):
    """Transforms this rank's blocks in place and gathers per-block metadata on every rank in file order.
    `blocks` maps global block index -> uint8 array for the indices of blocks_for_rank().  Returns a list of
    total_blocks dicts {index, owner, size, crc32, LF, freqs}."""
    mine = blocks_for_rank(total_blocks, rank, world)
    if sorted(blocks.keys()) != mine:
        raise ValueError("rank %d was handed blocks %s, expected %s" % (rank, sorted(blocks.keys()), mine))
    local = []
    for i in mine:
        LF, freqs = transform(i, blocks[i])
        local.append({"index": i, "owner": rank, "size": int(blocks[i].size),
                      "crc32": zlib.crc32(blocks[i].tobytes()) & 0xFFFFFFFF,
                      "LF": np.asarray(LF, dtype=np.uint32).tolist(),
                      "freqs": np.asarray(freqs, dtype=np.uint32).tolist()})
    if dist is not None and world > 1:
        gathered: List[Sequence[dict]] = [None] * world  # type: ignore[list-item]
        dist.all_gather_object(gathered, local)
    else:
        gathered = [local]
    merged = sorted((m for part in gathered for m in part), key=lambda m: m["index"])
    if [m["index"] for m in merged] != list(range(total_blocks)):
        raise RuntimeError("block metadata incomplete or duplicated after gather")
    for m in merged:
        if m["owner"] != owner_of(m["index"], world):
            raise RuntimeError("block %d transformed by the wrong rank" % m["index"])
    return merged


def bind_host_to_gpu(device: int):
    """Pin this process (and the worker threads it creates later) to the host cores NVML reports as local to
    `device` — pinned staging buffers are then first-touched on the GPU's NUMA node and H2D/D2H copies do not
    cross the socket interconnect.  One process per GPU (torchrun) calls this once, before allocating pinned
    memory.  Returns the number of cores bound to, or 0 if nothing was changed (single socket, NVML missing,
    BWTC_NUMA_BIND=0).  Host-side plumbing only: no effect on results."""
    import os

    if os.environ.get("BWTC_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return 0
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = None
        try:
            import torch

            uuid = str(torch.cuda.get_device_properties(device).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus or len(cpus) >= len(os.sched_getaffinity(0)):
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0
