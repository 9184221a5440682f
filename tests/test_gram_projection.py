"""CPU check of the arithmetic behind the gram mode of round 0 (bwt_kernels.cuh GramParams): for 6-bit and
3-bit codes every digit histogram of the round-0 radix sort is a projection of ONE histogram of the keys' low 12 bits,
plus u head terms and minus u tail terms.  The (u, s) choice below is the one phase_sort makes (bwt_engine.cu, "Gram mode");
the kernels themselves are checked bit-exactly on the GPU (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

GRAM_BITS = 12


def _keys(codes, c, b):
    """key(i) = c codes starting at i, most significant first, zero padding past the end (k_pack_round0)."""
    n = codes.size
    padded = np.concatenate([codes.astype(np.uint64), np.zeros(c, np.uint64)])
    key = np.zeros(n, np.uint64)
    for j in range(c):
        key = (key << np.uint64(b)) | padded[j:j + n]
    return key


def _plan(c, b, rb):
    keybits = c * b
    npass = -(-keybits // rb)
    W = GRAM_BITS // b
    plan = []
    for p in range(npass):
        lo = rb * p
        u = min(lo // b, c - W)
        s = lo - b * u
        assert s + rb <= GRAM_BITS or lo + rb > keybits, "digit outside its 12-bit window"
        plan.append((u, s))
    return plan


@pytest.mark.parametrize("b,c", [(6, 2), (6, 3), (6, 5), (6, 9), (6, 10), (3, 4), (3, 5), (3, 10), (3, 12), (3, 21)])
@pytest.mark.parametrize("rb", [8, 9])
def test_digit_histograms_are_projections_of_the_gram_histogram(b, c, rb):
    rng = np.random.default_rng(100 * b + c + rb)
    dmask = (1 << rb) - 1
    for n, tail in ((65, 0), (1000, (1 << b) - 1), (4099, 1)):
        codes = rng.integers(0, 1 << b, n).astype(np.uint64)
        codes[-30:] = tail
        key = _keys(codes, c, b)
        low = (key & np.uint64((1 << GRAM_BITS) - 1)).astype(np.int64)
        G = np.bincount(low, minlength=1 << GRAM_BITS)
        g = np.arange(1 << GRAM_BITS)
        for p, (u, s) in enumerate(_plan(c, b, rb)):
            direct = np.bincount(((key >> np.uint64(rb * p)) & np.uint64(dmask)).astype(np.int64), minlength=dmask + 1)
            proj = np.bincount((g >> s) & dmask, weights=G, minlength=dmask + 1).astype(np.int64)
            for i in range(u):  # head: suffixes 0..u-1 are not covered by a slid window; tail: the last u have no partner
                proj[int((int(key[i]) >> (rb * p)) & dmask)] += 1
                proj[(int(low[n - 1 - i]) >> s) & dmask] -= 1
            assert (proj == direct).all(), (n, p, u, s)
