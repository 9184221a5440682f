"""GPU tests of the host side of the engine: pinned vs pageable caller buffers (pinned staging ring), the streaming
submit / wait pipeline, the device-controlled ladder of sort-free rounds under every speculation depth, the sticky error
word with a REAL out-of-order tile dispatch (reversed tile map), and the raw contract with U != T."""
import ctypes

import numpy as np
import pytest

import bwtc_b200 as bw

pytestmark = pytest.mark.gpu


def _pinned(n):
    import torch

    return torch.empty(n, dtype=torch.uint8).pin_memory()


def test_pinned_and_pageable_buffers_give_identical_results(oracle):
    """Pageable blocks (what a malloc'ed PrecompressorBlock is, PrecompressorBlock.cpp:37-49) go through the pinned
    staging ring on copy streams of their own; pinned blocks are copied directly.  Sizes around the 4 MiB ring chunk."""
    ctx = bw.CudaContext(20 << 20)
    try:
        for n in (1, 4095, (4 << 20) - 1, 4 << 20, (4 << 20) + 1, (17 << 20) + 12345):
            x = bw.generate("markov", n, seed=n % 1000)
            k = bw.num_starting_points(n, 8)
            a = x.copy()  # pageable
            LFa, fra = np.zeros(k, np.uint32), np.zeros(256, np.uint32)
            ctx.bwt_block(a, LFa, fra)
            t = _pinned(n)
            b = t.numpy()
            b[:] = x
            LFb, frb = np.zeros(k, np.uint32), np.zeros(256, np.uint32)
            ctx.bwt_block(b, LFb, frb)
            assert np.array_equal(a, b) and (LFa == LFb).all() and (fra == frb).all(), n
            if n <= (4 << 20) + 1:
                w = oracle.block(x, 8)
                assert np.array_equal(a, w[0]) and (LFa == w[1]).all() and (fra == w[2]).all(), n
    finally:
        ctx.close()


def test_streaming_submit_wait(oracle):
    """bwtc_cuda_pipeline_submit / _wait: blocks are queued one by one as a reader would produce them, waited for in file
    order (the order an entropy coder consumes them), small ones still get batched when they queue up."""
    sizes = [1 << 18, 1 << 18, 77777, 1 << 20, 300, 1 << 16, 1 << 16, 1 << 16, 1 << 16, 5]
    blocks = [bw.generate(["markov", "dna", "random", "repetitive"][i % 4], s, seed=700 + i) for i, s in enumerate(sizes)]
    pipe = bw.Pipeline(1 << 20, depth=3)
    try:
        work = [b.copy() for b in blocks]
        handles = [pipe.submit(w, 8) for w in work]
        for i, h in enumerate(handles):
            LF, fr = pipe.wait(h)
            w = oracle.block(blocks[i], 8)
            assert np.array_equal(work[i], w[0]), i
            assert (LF == w[1]).all() and (fr == w[2]).all(), i
        # a second wave through the same (persistent) workers, waited for in reverse order
        work = [b.copy() for b in blocks[:4]]
        handles = [pipe.submit(w, 1) for w in work]
        for i in (3, 2, 1, 0):
            LF, fr = pipe.wait(handles[i])
            w = oracle.block(blocks[i], 1)
            assert np.array_equal(work[i], w[0]) and (LF == w[1]).all(), i
    finally:
        pipe.close()


@pytest.mark.parametrize("first,more", [(0, 1), (1, 1), (2, 4), (6, 6)])
def test_ladder_speculation_depth_does_not_change_results(oracle, monkeypatch, first, more):
    """The number of segmented rounds enqueued ahead of the host (BWTC_LADDER_FIRST / _MORE) is a scheduling knob: with 0
    the host steps in after every round, with 6 the whole tail of a DNA / Markov block runs without it."""
    monkeypatch.setenv("BWTC_LADDER_FIRST", str(first))
    monkeypatch.setenv("BWTC_LADDER_MORE", str(more))
    n = (1 << 20) + 77
    ctx = bw.CudaContext(n)
    try:
        for kind in ("markov", "dna", "repetitive", "random"):
            x = bw.generate(kind, n if kind != "repetitive" else 1 << 19, seed=51)
            w = oracle.block(x, 8)
            blk = x.copy()
            LF, fr = np.zeros(8, np.uint32), np.zeros(256, np.uint32)
            ctx.bwt_block(blk, LF, fr)
            assert np.array_equal(blk, w[0]) and (LF == w[1]).all() and (fr == w[2]).all(), (kind, first, more)
    finally:
        ctx.close()


def test_spin_wait_mode_gives_identical_results(oracle, monkeypatch):
    monkeypatch.setenv("BWTC_SPIN_WAIT", "1")
    x = bw.generate("markov", 600001, seed=52)
    w = oracle.block(x, 8)
    ctx = bw.CudaContext(x.size)
    try:
        blk = x.copy()
        LF, fr = np.zeros(8, np.uint32), np.zeros(256, np.uint32)
        ctx.bwt_block(blk, LF, fr)
    finally:
        ctx.close()
    assert np.array_equal(blk, w[0]) and (LF == w[1]).all() and (fr == w[2]).all()


def test_real_out_of_order_tiles_trip_the_watchdog_and_recover(oracle, monkeypatch):
    """Static tile ids assume CTAs are dispatched in index order.  BWTC_DEBUG_REVERSE_TILES maps tile = grid-1-blockIdx,
    the worst possible violation: with more tiles than resident CTAs the first wave waits for tiles that cannot start,
    the spin watchdog (shrunk by BWTC_DEBUG_SPIN_LIMIT so it fires in milliseconds) sets the STICKY error word, every
    later kernel of the block returns at entry instead of consuming half-written buffers, and the host repeats the sort
    phase with tickets from the text it still holds on the device — also for an in-place DEVICE buffer."""
    import torch

    monkeypatch.setenv("BWTC_DEBUG_REVERSE_TILES", "1")
    monkeypatch.setenv("BWTC_DEBUG_SPIN_LIMIT", "2000")
    n = 12 << 20  # 3073 tiles of 4096 records: far more than the ~450-600 resident CTAs
    x = bw.generate("markov", n, seed=53)
    w = oracle.block(x[: 1 << 20], 8)
    ctx = bw.CudaContext(n)
    try:
        # (a) host buffer
        blk = x.copy()
        LF, fr = np.zeros(8, np.uint32), np.full(256, 3, np.uint32)
        ctx.bwt_block(blk, LF, fr)
        assert ctx.stats()["flags"] & 1, "the watchdog should have fired and the block been repeated with tickets"
        assert (fr - 3 == np.bincount(x, minlength=256)).all(), "freqs must be incremented exactly once"
        # the context now uses tickets for good: compare with a fresh ticket-mode context on the same block
        monkeypatch.setenv("BWTC_DEBUG_REVERSE_TILES", "0")
        monkeypatch.setenv("BWTC_STATIC_TILES", "0")
        ref_ctx = bw.CudaContext(n)
        blk2 = x.copy()
        LF2 = np.zeros(8, np.uint32)
        ref_ctx.bwt_block(blk2, LF2, None)
        ref_ctx.close()
        assert np.array_equal(blk, blk2) and (LF == LF2).all()
        # small block, all tiles resident: reversed order completes without the watchdog, result exact
        s = x[: 1 << 20].copy()
        LFs, frs = np.zeros(8, np.uint32), np.zeros(256, np.uint32)
        ctx.bwt_block(s, LFs, frs)
        assert np.array_equal(s, w[0]) and (LFs == w[1]).all() and (frs == w[2]).all()
    finally:
        ctx.close()
    # (b) in-place device buffer, fresh context in reversed-tile mode
    monkeypatch.setenv("BWTC_DEBUG_REVERSE_TILES", "1")
    monkeypatch.delenv("BWTC_STATIC_TILES")
    ctx = bw.CudaContext(n)
    try:
        d = torch.from_numpy(x.copy()).cuda()
        LFd = np.zeros(8, np.uint32)
        ctx.bwt_block_device(d.data_ptr(), d.data_ptr(), n, LFd, None)
        torch.cuda.synchronize()
        assert ctx.stats()["flags"] & 1
        assert np.array_equal(d.cpu().numpy(), blk) and (LFd == LF).all(), "in-place device block after a watchdog retry"
    finally:
        ctx.close()


def test_raw_contract_leaves_U_pidx_untouched_when_U_is_not_T(oracle):
    """divbwtf(T, U != T, ...) never writes U[pidx] (divsufsort.c:506-512; test/DivsufsortTest.cpp:53 calls it that way)."""
    rng = np.random.default_rng(17)
    ctx = bw.CudaContext(1 << 20)
    try:
        for n in (2, 3, 1000, 300000):
            T = rng.integers(0, 4, n).astype(np.uint8)
            T[-1] = 0
            rc, wbuf, wLF, wfr = oracle.raw(T, 1)
            U = np.full(n, 0xEE, np.uint8)
            LF = np.zeros(1, np.uint32)
            fr = np.zeros(256, np.uint32)
            pidx = ctx.divbwtf(T.copy(), U, LF, fr)
            assert pidx == rc == LF[0]
            assert U[pidx] == 0xEE, "U[pidx] must stay untouched"
            keep = np.ones(n, bool)
            keep[pidx] = False
            assert np.array_equal(U[keep], wbuf[keep]) and (fr == wfr).all(), n
    finally:
        ctx.close()


def test_scratch_bytes_is_what_a_context_allocates():
    import torch

    lib = bw.load_library()
    for n in (1 << 20, 32 << 20):
        torch.cuda.synchronize()
        free0, _ = torch.cuda.mem_get_info()
        ctx = bw.CudaContext(n)
        free1, _ = torch.cuda.mem_get_info()
        ctx.close()
        used = free0 - free1
        want = int(lib.bwtc_cuda_scratch_bytes(n))
        assert want <= bw.SCRATCH_BYTES_PER_SUFFIX * (n + 1) + bw.SCRATCH_FIXED_BYTES
        assert abs(used - want) <= (64 << 20), (n, used, want)  # allocator granularity (2 MiB pages per allocation)
