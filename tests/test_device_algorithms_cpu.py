"""CPU restatements of the arithmetic tricks the kernels rely on, checked against plain definitions.  These do not run the
kernels (tests/test_gpu_parity.py does, against the oracle); they pin the reasoning that the kernels' comments state, so that a
change to a key layout or an index formula is caught without a GPU."""
import numpy as np
import pytest

import oracle_lib as O  # noqa: F401  (makes sure the checker library builds in this environment)


def _chain_direct(w, L, K, M):
    """binning.c:922-988 restated: restart windows take the LEFTMOST arg-max of w over [i, i+K-M]."""
    W, C = L - K + 1, K - M + 1
    segs, i = [], 0
    while i < W:
        best = i
        for p in range(i, i + C):
            if w[p] > w[best]:
                best = p
        nxt = min(best + 1, W)
        segs.append((i, nxt - i, best))
        i = nxt
    return segs


def _hop_packed(wr, i, C):
    """read_pack.cuh warp_signature_hop_strand<true>: one max over keys (score | inverted offset | strand)."""
    if C <= 32:
        keys = [(((wr[i + l] & ~1) << 5) | ((31 - l) << 1) | (wr[i + l] & 1)) if l < C else 0 for l in range(32)]
        mx = max(keys)
        assert mx < 2 ** 32
        return 31 - ((mx >> 1) & 31), ((mx >> 6) << 1) | (mx & 1)
    keys = []
    for l in range(32):
        v0 = wr[i + l]
        v1 = wr[i + l + 32] if l + 32 < C else 0
        keys.append(max(((v0 & ~1) << 6) | ((63 - l) << 1) | (v0 & 1), ((v1 & ~1) << 6) | ((31 - l) << 1) | (v1 & 1)))
    mx = max(keys)
    assert mx < 2 ** 32
    return 63 - ((mx >> 1) & 63), ((mx >> 7) << 1) | (mx & 1)


@pytest.mark.parametrize("seed", range(6))
def test_one_redux_signature_hop_is_the_leftmost_argmax_with_its_strand(seed):
    rng = np.random.default_rng(seed)
    for _ in range(150):
        M = int(rng.integers(2, 13))
        K = int(rng.integers(2 * M, 65))
        C = K - M + 1
        packed_ok = 2 * M + 1 + (5 if C <= 32 else 6) <= 32  # the launcher's condition for the packed form
        if not packed_ok:
            continue
        L = int(rng.integers(K, K + 120))
        alphabet = [0, 1, 2, 3] if rng.random() < 0.7 else [0, 3]  # two-letter reads: many ties
        v = rng.choice(alphabet, size=L)
        FULL = 4 ** M - 1
        s = [sum(int(v[p + t]) * 4 ** (M - 1 - t) for t in range(M)) for p in range(L - M + 1)]
        w = [max(x, FULL - x) for x in s]
        rev = [1 if FULL - x > x else 0 for x in s]
        wr = [(a << 1) | b for a, b in zip(w, rev)]
        want = _chain_direct(w, L, K, M)
        got, i, W = [], 0, L - K + 1
        while i < W:
            off, w_rev = _hop_packed(wr, i, C)
            sig = i + off
            assert w_rev == wr[sig], "the hop returns w(sig) << 1 | is_rev(sig)"
            nxt = min(sig + 1, W)
            got.append((i, nxt - i, sig))
            i = nxt
        assert got == want


def test_staging_places_of_units_are_disjoint():
    """skr_group.cu: a unit stages its ids at its instance coordinate b and its k-mers at b // (cutoff + 1).  With S <= n // (cutoff + 1)
    surviving k-mers per unit of n instances (a survivor has more than `cutoff` instances) the k-mer ranges never overlap."""
    rng = np.random.default_rng(3)
    for cutoff in (0, 1, 2, 5):
        d = cutoff + 1
        sizes = rng.integers(0, 1300, size=4000)
        base = np.concatenate([[0], np.cumsum(sizes)[:-1]])
        kb = base // d
        smax = sizes // d  # the most survivors a unit can have
        assert (kb[1:] >= kb[:-1] + smax[:-1]).all()
        assert kb[-1] + smax[-1] <= sizes.sum() // d + 1
