"""CPU tests: the C oracle (oracle/oracle_bwt.c) is pinned against
  (a) the known answers / golden fixtures captured from the compiled reference (tests/golden/),
  (b) the unmodified reference itself (oracle/_ref), engines 'd' and 's', on seeded random inputs,
  (c) the independent definitions the reference's own tests use: sorted suffix array (test/SaisTest.cpp:55-70),
      LFpowers by naive LF iteration (test/LFpowersTest.cpp:119-132), forward o inverse = id
      (test/InverseBwtTest.cpp:51-92)."""
import numpy as np
import pytest


def test_oracle_matches_golden(oracle, golden):
    for name, g in golden.items():
        out, LF, fr = oracle.block(g["in"], int(g["starts"][0]))
        assert (out == g["out"]).all(), name
        assert (LF == g["LF"]).all(), name
        assert (fr == g["freqs"]).all(), name


def test_known_answers_from_survey(oracle):
    # captured from the compiled reference via BWTManager, engines d and s agree
    known = {b"mississippi": (b"msispipissi", [2]), b"banana": (b"bnnaaa", [3]),
             b"abracadabra": (b"abdbcarraaa", [5]), b"aaaaaaaa": (b"aaaaaaaa", [8]), b"a": (b"a", [1]),
             b"ab": (b"ab", [2]), b"ba": (b"ba", [1])}
    for s, (want, lf) in known.items():
        out, LF, fr = oracle.block(np.frombuffer(s, np.uint8).copy(), 8)
        assert bytes(out) == want and list(LF) == lf
    out, LF, _ = oracle.block(np.frombuffer(b"ab" * 150, np.uint8).copy(), 8)
    assert list(LF) == [300, 168, 37, 205, 74, 242, 111, 279]
    out, LF, _ = oracle.block(np.zeros(300, np.uint8), 8)
    assert list(LF) == [300, 36, 73, 110, 147, 184, 221, 258]
    out, LF, _ = oracle.block(((7 * np.arange(1000)) % 251).astype(np.uint8), 8)
    assert list(LF) == [868, 433, 920, 406, 894, 379, 867, 352]


def test_oracle_matches_reference_block(oracle, reference):
    rng = np.random.default_rng(1)
    for trial in range(150):
        n = int(rng.integers(1, 3000))
        sigma = int(rng.choice([1, 2, 3, 4, 16, 256]))
        x = rng.integers(0, sigma, n).astype(np.uint8)
        starts = int(rng.integers(0, 300))
        want = oracle.block(x, starts)
        for algo in (b"d", b"s"):
            got = reference.block(x, starts, algo)
            assert (got[0] == want[0]).all() and (got[1] == want[1]).all() and (got[2] == want[2]).all(), \
                (trial, n, sigma, starts, algo)


def test_oracle_matches_reference_raw(oracle, reference):
    """Raw virtual doTransform(byte*, uint32, vector<uint32>&, freqs) with ARBITRARY last byte.
    One documented reference quirk is excluded: with x = N / nLF == 1 and T[N-2] < T[N-1] divsufsort's
    construct_BWT never samples suffix N-1 (divsufsort.c:372-373 stores it as ~T[n-2]), so LFpowers[1]
    keeps its previous value.  It cannot occur at block level (the last byte is the 0x00 sentinel)."""
    rng = np.random.default_rng(2)
    for trial in range(150):
        n = int(rng.integers(2, 2000))
        sigma = int(rng.choice([1, 2, 3, 4, 16, 256]))
        T = rng.integers(0, sigma, n).astype(np.uint8)
        nLF = int(rng.integers(1, min(n, 40) + 1))
        rc, wbuf, wLF, wfr = oracle.raw(T, nLF)
        assert rc == wLF[0]
        quirk = (n // nLF == 1) and T[n - 2] < T[n - 1]
        for algo in (b"d", b"s"):
            buf, LF, fr = reference.raw(T, nLF, algo)
            if quirk and algo == b"d":
                LF[1] = wLF[1]
            assert (buf == wbuf).all() and (LF == wLF).all() and (fr == wfr).all(), (trial, n, sigma, nLF, algo)


def test_raw_trivial_sizes(oracle):
    # divsufsort.c:488-489: n <= 1 returns n, copies the byte, leaves LFpowers alone
    T = np.array([7], np.uint8)
    rc, buf, LF, fr = oracle.raw(T, 1)
    assert rc == 1 and buf[0] == 7 and fr.sum() == 0


def test_suffix_array_is_sorted(oracle):
    # the property test/SaisTest.cpp:55-70 checks for sais: naive suffix comparison, shorter first
    rng = np.random.default_rng(3)
    for sigma in (2, 4, 256):
        T = rng.integers(0, sigma, 400).astype(np.uint8)
        SA, ISA = oracle.suffix_array(T)
        suf = [bytes(T[i:]) for i in SA]
        assert suf == sorted(suf)
        assert (ISA[SA] == np.arange(T.size)).all()


def _naive_lf_powers(out, pidx, k):
    """LFpowers as test/LFpowersTest.cpp:119-132 defines them: walk the LF mapping from the end-of-block row.
    Rows = the N = n+1 suffixes of T' = reverse(X) + 0x00 (shorter suffix first); L[r] = T'[SA[r]-1], undefined
    at r = pidx (suffix 0).  Row 0 is suffix N-1 (the appended 0x00), which no row maps to, so
    LF(r) = 1 + #{L < L[r]} + #{r' < r : L[r'] = L[r]}; LF^s(row 0) = row of suffix N-1-s."""
    n = out.size
    N = n + 1
    L = np.full(N, -1, np.int64)
    L[:n] = out
    if pidx < n:
        L[N - 1] = out[pidx]   # undo the hole fill (BWTransform.cpp:60)
    L[pidx] = -1
    cnt = np.bincount(L[L >= 0], minlength=256)
    less = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    occ = np.zeros(256, np.int64)
    LFmap = np.full(N, -1, np.int64)
    for r in range(N):
        c = L[r]
        if c >= 0:
            LFmap[r] = 1 + less[c] + occ[c]
            occ[c] += 1
    x = N // k
    rank_of = {N - 1: 0}
    r, s = 0, N - 1
    while s > 0:
        r = LFmap[r]
        s -= 1
        rank_of[s] = int(r)
    assert rank_of[0] == pidx
    return [pidx] + [rank_of[N - j * x] for j in range(1, k)]


def test_lfpowers_definition(oracle):
    rng = np.random.default_rng(4)
    for n, sigma, starts in [(300, 2, 8), (1000, 4, 8), (777, 256, 5), (2000, 3, 30)]:
        x = rng.integers(0, sigma, n).astype(np.uint8)
        out, LF, _ = oracle.block(x, starts)
        assert _naive_lf_powers(out, int(LF[0]), LF.size) == [int(v) for v in LF]


def test_forward_inverse_roundtrip(oracle, reference):
    # test/InverseBwtTest.cpp:51-92: forward (here: the oracle) o reference MTL-SA inverse = identity
    rng = np.random.default_rng(5)
    for n, sigma, starts in [(1, 2, 1), (2, 2, 1), (300, 2, 8), (5000, 4, 8), (20000, 256, 17), (3000, 1, 256)]:
        x = rng.integers(0, sigma, n).astype(np.uint8)
        out, LF, _ = oracle.block(x, starts)
        back = reference.inverse_block(out, LF)
        assert (back == x).all(), (n, sigma, starts)


def test_num_starting_points(oracle):
    f = oracle.lib.oracle_num_starting_points
    assert f(256, 8) == 1 and f(257, 8) == 8 and f(1000, 0) == 1 and f(1000, 300) == 256 and f(10, 256) == 1


def test_oracle_inverse_matches_reference_inverse(oracle, reference):
    """the oracle's inverse restatement (plain LF walk, InverseBWT.cpp:58-115) pinned on the
    reference's own InverseBWTransform::doTransform(BWTBlock&) (MtlSaInverseBWT) and on the forward transform."""
    rng = np.random.default_rng(99)
    for n, sigma, starts in [(1, 2, 8), (2, 2, 8), (3, 1, 1), (300, 4, 8), (5000, 256, 256), (70001, 64, 7), (4096, 2, 1)]:
        x = rng.integers(0, sigma, n).astype(np.uint8)
        b, LF, fr = reference.block(x, starts)
        back_ref = reference.inverse_block(b, LF)
        back_orc = oracle.inverse_block(b, LF[0])
        assert (back_ref == x).all() and (back_orc == x).all(), (n, sigma, starts)
