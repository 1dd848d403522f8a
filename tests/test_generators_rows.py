"""Row-block generators used by the row-sharded runs of BASELINE configs 4 and 5
(sigma_b200/generators.py fem_p1_csr_rows, erdos_renyi_csr_rows): every rank builds its own
rows, and what it builds must be the corresponding slice of the whole matrix entry for entry
(the whole-matrix generators are themselves pinned on the reference-style builders through the
oracle, tests/test_oracle_kat.py)."""
import numpy as np
import pytest

from sigma_b200 import generators as G


def block_of(ptr, node, val, lo, hi):
    sl = slice(ptr[lo] - 1, ptr[hi] - 1)
    return ptr[lo:hi + 1] - ptr[lo] + 1, node[sl], val[sl]


@pytest.mark.parametrize("N", [5, 12, 37])
def test_fem_row_blocks_equal_slices_of_the_whole_matrix(N):
    n = N * N
    whole = G.fem_p1_csr(N)
    for lo, hi in [(0, n), (0, 1), (n - 1, n), (3, n // 2), (n // 3, 2 * n // 3 + 1), (N, 2 * N), (N - 1, N + 1)]:
        got = G.fem_p1_csr_rows(N, lo, hi)
        for a, b in zip(got, block_of(*whole, lo, hi)):
            assert a.dtype == b.dtype and np.array_equal(a, b), (N, lo, hi)


@pytest.mark.parametrize("kw", [dict(), dict(weights="random"), dict(weights="random", skew=True, shift=2.0),
                                dict(shift=0.0)], ids=["unit", "random", "random_skew", "laplacian"])
def test_er_row_blocks_equal_slices_of_the_whole_matrix(kw):
    n = 700
    whole = G.erdos_renyi_csr(n, seed=5, **kw)
    cache = {}
    G.erdos_renyi_csr_rows(n, 0, 10, seed=5, weights="random", skew=True, cache=cache)   # another variant fills the cache
    for lo, hi in [(0, n), (0, 1), (n - 1, n), (100, 350), (349, 351)]:
        *got, counts = G.erdos_renyi_csr_rows(n, lo, hi, seed=5, cache=cache if lo else None, **kw)
        for a, b in zip(got, block_of(*whole, lo, hi)):
            assert a.dtype == b.dtype and np.array_equal(a, b), (lo, hi)
        assert np.array_equal(counts, np.diff(whole[0]))


def test_blocks_tile_the_matrix():
    """Blocks of a partition concatenate to the whole matrix (what the 8 ranks hold together)."""
    N = 23
    n = N * N
    ptr, node, val = G.fem_p1_csr(N)
    cuts = [0, 70, 71, 300, n]
    nodes = np.concatenate([G.fem_p1_csr_rows(N, a, b)[1] for a, b in zip(cuts[:-1], cuts[1:])])
    vals = np.concatenate([G.fem_p1_csr_rows(N, a, b)[2] for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.array_equal(nodes, node) and np.array_equal(vals, val)


def test_row_blocks_hypothesis():
    """Any block [lo, hi) of either generator equals the slice of the whole matrix."""
    from hypothesis import given, settings, strategies as st

    fem = {N: G.fem_p1_csr(N) for N in (4, 9, 16)}
    er = G.erdos_renyi_csr(400, seed=21, weights="random", skew=True)
    cache = {}

    @settings(max_examples=60, deadline=None)
    @given(st.sampled_from([4, 9, 16]), st.integers(0, 10**6), st.integers(0, 10**6))
    def check(N, a, b):
        n = N * N
        lo, hi = sorted((a % (n + 1), b % (n + 1)))
        if lo == hi:
            return
        for x, y in zip(G.fem_p1_csr_rows(N, lo, hi), block_of(*fem[N], lo, hi)):
            assert np.array_equal(x, y)
        lo, hi = sorted((a % 401, b % 401))
        if lo == hi:
            return
        *got, _ = G.erdos_renyi_csr_rows(400, lo, hi, seed=21, weights="random", skew=True, cache=cache)
        for x, y in zip(got, block_of(*er, lo, hi)):
            assert np.array_equal(x, y)

    check()
