"""Host-side row tiling of the streaming CSR kernel (csrc/kernels_spmv.cu build_tiles_host),
checked without a GPU against the plain row-by-row greedy walk it replaces: consecutive rows
while the tile holds <= 2045 stored entries and <= 512 rows (patterns with at least 12 stored
entries per row on average: <= 1021 entries and <= 256 rows, the small tile shape of gather-bound
operators, csrc/kernels_spmv.cu tile_shape_for); a longer row is a tile of its own."""
import ctypes as C

import numpy as np
import pytest

from sigma_b200._capi import check, lib, ptr

CAP, ROWS = 2045, 512
SMALL_CAP, SMALL_ROWS, SMALL_MIN_ROW = 1021, 256, 12


def shape(ptr1):
    """tile_shape_for: the caps the library picks for this pattern"""
    n = ptr1.size - 1
    nnz = int(ptr1[-1]) - 1
    return (SMALL_CAP, SMALL_ROWS) if n > 0 and nnz >= SMALL_MIN_ROW * n else (CAP, ROWS)


def greedy(ptr1):
    n = ptr1.size - 1
    CAP, ROWS = shape(ptr1)
    tiles, s = [], 0
    while s < n:
        e = s + 1
        base = int(ptr1[s])
        while e < n and int(ptr1[e + 1]) - base <= CAP and e - s < ROWS:
            e += 1
        tiles.append((s, e, int(ptr1[s]) - 1, int(ptr1[e]) - 1))
        s = e
    return np.array(tiles, np.int32).reshape(-1, 4)


def library_tiles(ptr1):
    ptr1 = np.ascontiguousarray(ptr1, np.int32)
    n = ptr1.size - 1
    out = np.empty((max(n, 1), 4), np.int32)
    nt = C.c_int32()
    check(lib().sigb_debug_row_tiles(n, ptr(ptr1), ptr(out), C.byref(nt)))
    return out[: nt.value].copy()


def cases():
    rng = np.random.default_rng(0)
    yield "empty", np.array([1], np.int32)
    yield "one_row", np.array([1, 4], np.int32)
    yield "uniform5", 1 + 5 * np.arange(0, 100_001, dtype=np.int64)
    yield "all_empty_rows", np.ones(3000, np.int64)
    yield "row_of_exactly_cap", np.concatenate([[1], 1 + np.cumsum([CAP, 1, CAP - 1, 1, 1] + [0] * 400)])
    yield "row_of_exactly_small_cap", np.concatenate([[1], 1 + np.cumsum([SMALL_CAP, 1, SMALL_CAP - 1, 1, 1, 2 * SMALL_CAP])])
    yield "uniform22", 1 + 22 * np.arange(0, 50_001, dtype=np.int64)
    yield "long_rows", np.concatenate([[1], 1 + np.cumsum(rng.choice([0, 3, 700, 2045, 2046, 9000], 400))])
    deg = rng.integers(0, 40, 50_000)
    deg[rng.integers(0, deg.size, 30)] = rng.integers(2000, 5000, 30)
    yield "random_with_spikes", np.concatenate([[1], 1 + np.cumsum(deg)])
    yield "dense_band", 1 + 1000 * np.arange(0, 2001, dtype=np.int64)


@pytest.mark.parametrize("case", list(cases()), ids=lambda c: c[0])
def test_tiles_equal_the_greedy_walk(case):
    _, p = case
    p = np.asarray(p, np.int32)
    got, want = library_tiles(p), greedy(p)
    assert np.array_equal(got, want)
    n = p.size - 1
    if n:
        # a partition of the rows, each tile within the limits unless it is a single long row
        assert got[0, 0] == 0 and got[-1, 1] == n and np.array_equal(got[1:, 0], got[:-1, 1])
        nrows, nent = got[:, 1] - got[:, 0], got[:, 3] - got[:, 2]
        cap, rows = shape(p)
        assert np.all(nrows >= 1) and np.all(nrows <= rows)
        assert np.all((nent <= cap) | (nrows == 1))


def device_algorithm_replica(ptr1):
    """numpy replica, step for step, of the EXPERIMENTAL device tiling (csrc/tiles_device.cu):
    next(s) by bisection for all rows at once, the orbit of row 0 under next marked by pointer
    doubling, marked rows compacted by an exclusive scan.  Runs on the CPU so the index logic is
    checked without a GPU; the gated GPU test compares the kernels themselves."""
    p = np.asarray(ptr1, np.int64)
    n = p.size - 1
    if n == 0:
        return np.zeros((0, 4), np.int32), 0
    s = np.arange(n)
    CAP, ROWS = shape(np.asarray(ptr1))
    limit = p[:-1] + CAP
    hi = np.minimum(s + ROWS, n)
    # largest e in [s + 1, hi] with p[e] <= limit, s + 1 when the first row alone is too long
    e = np.searchsorted(p, limit, side="right") - 1          # largest index with p[idx] <= limit
    nxt = np.concatenate([np.clip(e, s + 1, hi), [n]])
    mark = np.zeros(n + 1, bool)
    mark[0] = True
    jump = nxt.copy()
    rounds = 1
    while (1 << rounds) <= n:
        rounds += 1
    used = 0
    for _ in range(rounds):
        src = np.flatnonzero(mark[:n])
        dst = jump[src]
        mark[dst[dst < n]] = True
        done = jump[0] == n
        jump = jump[jump]
        used += 1
        if done:
            break
    starts = np.flatnonzero(mark[:n])
    tiles = np.stack([starts, nxt[starts], p[starts] - 1, p[nxt[starts]] - 1], axis=1).astype(np.int32)
    return tiles, used


@pytest.mark.parametrize("case", list(cases()), ids=lambda c: c[0])
def test_device_algorithm_replica_equals_the_greedy_walk(case):
    _, p = case
    p = np.asarray(p, np.int32)
    got, rounds = device_algorithm_replica(p)
    want = greedy(p)
    assert np.array_equal(got, want)
    if want.shape[0] > 1:
        assert rounds <= int(np.ceil(np.log2(want.shape[0]))) + 1   # log-depth, not one round per tile


def test_tilings_on_random_row_lengths_hypothesis():
    """Random row-length sequences (mixtures of empty, short, cap-sized and oversized rows): the
    library's bisection tiling and the replica of the device algorithm both equal the greedy walk."""
    from hypothesis import given, settings, strategies as st

    lengths = st.lists(st.one_of(st.integers(0, 12), st.sampled_from([0, 0, 1, 2044, 2045, 2046, 4000]),
                                 st.integers(0, 700)), min_size=0, max_size=1500)

    @settings(max_examples=80, deadline=None)
    @given(lengths)
    def check(deg):
        p = np.concatenate([[1], 1 + np.cumsum(np.asarray(deg, np.int64))]).astype(np.int32)
        want = greedy(p)
        assert np.array_equal(library_tiles(p), want)
        assert np.array_equal(device_algorithm_replica(p)[0], want)

    check()


def balanced_tiles(ptr1, groups):
    ptr1 = np.ascontiguousarray(ptr1, np.int32)
    n = ptr1.size - 1
    out = np.empty((max(n, 1), 4), np.int32)
    nt = C.c_int32()
    check(lib().sigb_debug_row_tiles_balanced(n, ptr(ptr1), groups, ptr(out), C.byref(nt)))
    return out[: nt.value].copy()


@pytest.mark.parametrize("groups", [1, 7, 440, 444])
@pytest.mark.parametrize("case", list(cases()), ids=lambda c: c[0])
def test_balanced_tiles_are_a_valid_tiling(case, groups):
    """build_tiles_balanced (row-sharded operators): always a partition of the rows within the limits
    of the kernel (or single long rows); when it does not fall back to the greedy walk, the tile count
    is a multiple of `groups` and the tiles differ by at most one row's worth of entries."""
    _, p = case
    p = np.asarray(p, np.int32)
    got = balanced_tiles(p, groups)
    n = p.size - 1
    if n == 0:
        assert got.shape[0] == 0
        return
    assert got[0, 0] == 0 and got[-1, 1] == n and np.array_equal(got[1:, 0], got[:-1, 1])
    assert np.array_equal(got[:, 2], p[got[:, 0]] - 1) and np.array_equal(got[:, 3], p[got[:, 1]] - 1)
    nrows, nent = got[:, 1] - got[:, 0], got[:, 3] - got[:, 2]
    assert np.all(nrows >= 1) and np.all(nrows <= ROWS)
    assert np.all((nent <= CAP) | (nrows == 1))
    if not np.array_equal(got, greedy(p)):
        assert got.shape[0] % groups == 0
        assert nent.max() - nent.min() <= 2 * np.diff(p.astype(np.int64)).max()


def test_balanced_tiles_on_a_poisson_shard():
    """The 8-GPU shard of the headline matrix (2.1 M rows, 5 entries per row) on the 440 compute CTAs of
    the persistent kernel: 12 tiles for every CTA instead of 11 for some and 12 for the others."""
    N, rows = 4096, 4096 * 512
    cnt = np.full(rows, 5, np.int64)
    cnt[::N] -= 1
    cnt[N - 1::N] -= 1
    p = np.concatenate([[1], 1 + np.cumsum(cnt)]).astype(np.int32)
    g, b = greedy(p), balanced_tiles(p, 440)
    assert g.shape[0] % 440 != 0 and b.shape[0] % 440 == 0
    assert b.shape[0] // 440 == -(-g.shape[0] // 440)          # the same number of tiles for the slowest CTA ...
    nent = b[:, 3] - b[:, 2]
    assert nent.max() - nent.min() <= 10                       # ... and nobody holds more than anybody else
