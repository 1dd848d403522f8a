"""Host-side index work of the device ILDU(0) (sigma_b200/csrc/ldu_host.cpp, no GPU):
patterns bit-exact against the oracle's restatement of incomplete_ldu_sparsity_pattern
(src/solver/ldu_solvers.f90:396-441), destinations of A's entries, and level schedules
that really make the rows of a level independent -- shown by replaying the device
algorithm level by level in numpy, rows of a level in REVERSED order, and getting the
oracle's factors and solves bit for bit."""
import numpy as np
import pytest

import sigma_b200 as sb
from sigma_b200 import generators as G


def shuffled(ptr, node, val, seed):
    rng = np.random.default_rng(seed)
    node, val = node.copy(), val.copy()
    for i in range(ptr.size - 1):
        sl = slice(ptr[i] - 1, ptr[i + 1] - 1)
        perm = rng.permutation(sl.stop - sl.start)
        node[sl], val[sl] = node[sl][perm], val[sl][perm]
    return node, val


def cases():
    yield "tridiag", 50, *G.tridiag_csr(50, 2.5, -1.0, -0.5)
    yield "poisson", 15 * 15, *G.poisson2d_csr(15)
    p, n_, v = G.erdos_renyi_csr(300, seed=3, weights="random", skew=True, shift=2.0)
    yield "er_skew_shuffled", 300, p, *shuffled(p, n_, v, 1)
    p, n_, v = G.fem_p1_csr(12)
    yield "fem_shuffled", 144, p, *shuffled(p, n_, v, 2)
    yield "diagonal", 5, np.arange(1, 7, dtype=np.int32), np.arange(1, 6, dtype=np.int32), np.arange(1.0, 6.0)


def get_value(ptr, node, val, i, j):
    z = 0.0
    for k in range(ptr[i - 1] - 1, ptr[i] - 1):
        if node[k] == j:
            z = val[k]
    return z


def replay_factor(n, ptr, node, val, S):
    """ldu_scatter_kernel + ldu_factor_level_kernel of csrc/ldu.cu, in numpy."""
    Lptr, Lnode, Uptr, Unode = S["Lptr"], S["Lnode"], S["Uptr"], S["Unode"]
    nL, nU = Lnode.size, Unode.size
    fac = np.zeros(nL + nU + n)
    fac[S["dest"]] = val
    Lval, Uval, D = fac[:nL], fac[nL:nL + nU], fac[nL + nU:]
    lev = S["forward_lev"]
    for l in range(lev.size - 1):
        for i in S["forward_rows"][lev[l]:lev[l + 1]][::-1]:          # any order inside a level
            lb, dl = Lptr[i - 1] - 1, Lptr[i] - Lptr[i - 1]
            ub, du = Uptr[i - 1] - 1, Uptr[i] - Uptr[i - 1]
            for ind1 in range(dl):
                k = Lnode[lb + ind1]
                Lik = Lval[lb + ind1]
                Uki = get_value(Uptr, Unode, Uval, k, i)
                Dk = D[k - 1]
                Lik = Lik / Dk
                Lval[lb + ind1] = Lik
                LikDk = Lik * Dk
                for ind2 in range(dl):
                    j = Lnode[lb + ind2]
                    if j > k:
                        Lval[lb + ind2] = Lval[lb + ind2] + -(LikDk * get_value(Uptr, Unode, Uval, k, j))
                D[i - 1] = D[i - 1] - LikDk * Uki
                for ind2 in range(du):
                    j = Unode[ub + ind2]
                    Uval[ub + ind2] = Uval[ub + ind2] + -(LikDk * get_value(Uptr, Unode, Uval, k, j))
            Uval[ub:ub + du] = Uval[ub:ub + du] / D[i - 1]
    return Lval, Uval, D


def replay_solve(n, S, Lval, Uval, D, b):
    """copy_kernel, tri_level_kernel (forward), divide_kernel, tri_level_kernel (backward)."""
    x = b.copy()
    for rows, lev, ptr, node, val in ((S["forward_rows"], S["forward_lev"], S["Lptr"], S["Lnode"], Lval),
                                      (None, None, None, None, None),
                                      (S["backward_rows"], S["backward_lev"], S["Uptr"], S["Unode"], Uval)):
        if rows is None:
            x = x / D
            continue
        for l in range(1, lev.size - 1):
            for i in rows[lev[l]:lev[l + 1]][::-1]:
                z = x[i - 1]
                for k in range(ptr[i - 1] - 1, ptr[i] - 1):
                    z = z - val[k] * x[node[k] - 1]
                x[i - 1] = z
    return x


@pytest.mark.parametrize("case", list(cases()), ids=lambda c: c[0])
def test_symbolic_and_level_replay(orc, case):
    _, n, ptr, node, val = case
    S = sb.ldu_symbolic(n, ptr, node)
    F = orc.ldu_setup(orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr))
    # patterns: index work, bit-exact
    assert np.array_equal(S["Lptr"], F.Lptr) and np.array_equal(S["Lnode"], F.Lnode)
    assert np.array_equal(S["Uptr"], F.Uptr) and np.array_equal(S["Unode"], F.Unode)
    # every row once, ascending inside a level, neighbours strictly in earlier levels
    for rows, lev, p_, nd in ((S["forward_rows"], S["forward_lev"], F.Lptr, F.Lnode),
                              (S["backward_rows"], S["backward_lev"], F.Uptr, F.Unode)):
        assert np.array_equal(np.sort(rows), np.arange(1, n + 1))
        assert lev[0] == 0 and lev[-1] == n and np.all(np.diff(lev) > 0)
        level_of = np.empty(n + 1, np.int64)
        for l in range(lev.size - 1):
            seg = rows[lev[l]:lev[l + 1]]
            assert np.all(np.diff(seg) > 0)
            level_of[seg] = l
        for i in range(1, n + 1):
            nb = nd[p_[i - 1] - 1: p_[i] - 1]
            if nb.size:
                assert level_of[nb].max() == level_of[i] - 1        # tight: as early as possible
            else:
                assert level_of[i] == 0
    # the device algorithm replayed level by level == the serial reference loops, bit for bit
    Lval, Uval, D = replay_factor(n, ptr, node, val, S)
    assert np.array_equal(Lval, F.Lval) and np.array_equal(Uval, F.Uval) and np.array_equal(D, F.D)
    b = np.random.default_rng(0).standard_normal(n)
    assert np.array_equal(replay_solve(n, S, Lval, Uval, D, b), orc.ldu_solve(F, b))


def test_poisson_levels_are_antidiagonals():
    """Natural ordering of the 5-point stencil: level = ix + iy, 2N - 1 levels (why this
    preconditioner is latency-bound on a GPU)."""
    N = 20
    ptr, node, _ = G.poisson2d_csr(N)
    S = sb.ldu_symbolic(N * N, ptr, node)
    assert S["forward_lev"].size - 1 == 2 * N - 1 and S["backward_lev"].size - 1 == 2 * N - 1
    assert np.array_equal(np.diff(S["forward_lev"]), np.concatenate([np.arange(1, N + 1), np.arange(N - 1, 0, -1)]))


def test_symbolic_rejects_bad_columns():
    with pytest.raises(sb.SigmaError) as e:
        sb.ldu_symbolic(2, [1, 2, 3], [1, 5])
    assert e.value.status == 1


def chunked_sweep_model(ptr, node, val, start, n, chunk_rows, backward, rng, warp=4, max_delay=3):
    """Model of tri_chunked_kernel (csrc/ldu.cu): thread t owns the rows [t B, (t + 1) B) counted from
    the sweep's start and solves them in order; warps run trips in lockstep, the warps in a random
    order; per trip two convergent rounds resolve the next entry from the lane's own last result, from
    the last result of the owning lane of the same warp (the shuffle), or from the published words,
    which become visible only a random number of scheduling steps after they were written."""
    nchunks = (n + chunk_rows - 1) // chunk_rows
    nwarps = (nchunks + warp - 1) // warp
    x = np.zeros(n)
    visible_at = np.full(n, np.inf)          # global step at which the published words of row i can be read
    clock = 0
    lanes = []
    for t in range(nwarps * warp):
        if t < nchunks:
            lo, hi = t * chunk_rows, min(n, (t + 1) * chunk_rows)
            i, last = (lo + 1, hi) if not backward else (n - lo, n - hi + 1)
            lanes.append({"i": i, "last": last, "done": False, "have": False, "k": 0, "e": 0, "z": 0.0,
                          "prev_row": 0, "prev_z": 0.0})
        else:
            lanes.append({"done": True, "prev_row": 0, "prev_z": 0.0, "have": False})
    step = -1 if backward else 1

    def owner(j, t_base):
        pos = (n - j) if backward else (j - 1)
        return pos // chunk_rows - t_base

    def consume(ln, xj):
        ln["z"] = ln["z"] - val[ln["k"]] * xj
        ln["k"] += 1

    live = list(range(nwarps))
    idle = 0
    trips = 0
    while live:
        progressed = False
        for w in rng.permutation(live):
            clock += 1
            trips += 1
            base = w * warp
            wl = lanes[base:base + warp]
            for ln in wl:
                if not ln["done"] and not ln["have"]:
                    i = ln["i"]
                    ln["k"], ln["e"], ln["z"], ln["have"] = ptr[i - 1] - 1, ptr[i] - 1, start[i - 1], True
            for _ in range(2):                                   # the convergent rounds
                snap = [(l["prev_row"], l["prev_z"]) for l in wl]    # what the shuffles can deliver
                for ln in wl:
                    if ln["done"] or ln["k"] >= ln["e"]:
                        continue
                    j = node[ln["k"]]
                    ol = owner(j, base)
                    if j == ln["prev_row"]:
                        consume(ln, ln["prev_z"]); progressed = True
                    elif 0 <= ol < warp and snap[ol][0] == j:
                        consume(ln, snap[ol][1]); progressed = True
                    elif visible_at[j - 1] <= clock:
                        consume(ln, x[j - 1]); progressed = True
            for ln in wl:
                if ln["done"]:
                    continue
                while ln["k"] < ln["e"]:
                    j = node[ln["k"]]
                    if j == ln["prev_row"]:
                        consume(ln, ln["prev_z"])
                    elif visible_at[j - 1] <= clock:
                        consume(ln, x[j - 1])
                    else:
                        break
                    progressed = True
                if ln["k"] == ln["e"]:
                    i = ln["i"]
                    x[i - 1] = ln["z"]
                    visible_at[i - 1] = clock + int(rng.integers(0, max_delay + 1))
                    ln["prev_row"], ln["prev_z"], ln["have"] = i, ln["z"], False
                    progressed = True
                    if i == ln["last"]:
                        ln["done"] = True
                    else:
                        ln["i"] = i + step
            if all(l["done"] for l in wl):
                live = [v for v in live if v != w]
        idle = 0 if progressed else idle + 1
        assert idle <= max_delay + 2, "no warp can make progress: the sweep would hang"
    return x, trips


@pytest.mark.parametrize("case", list(cases()), ids=lambda c: c[0])
@pytest.mark.parametrize("chunk_rows", [1, 3, 16])
def test_chunked_sweep_model_terminates_and_matches(orc, case, chunk_rows):
    """The chunked sweeps: under random warp scheduling, delayed visibility of the published words and
    any chunk length, every sweep terminates and gives the serial result bit for bit."""
    _, n, ptr, node, val = case
    F = orc.ldu_setup(orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr))
    rng = np.random.default_rng(chunk_rows)
    b = rng.standard_normal(n)
    y, _ = chunked_sweep_model(F.Lptr, F.Lnode, F.Lval, b, n, chunk_rows, False, rng)
    x, _ = chunked_sweep_model(F.Uptr, F.Unode, F.Uval, y / F.D, n, chunk_rows, True, rng)
    assert np.array_equal(x, orc.ldu_solve(F, b))


def test_chunked_sweep_is_a_wavefront_on_the_stencil(orc):
    """Chunk length = bandwidth on the N x N five-point stencil: the chunks run one trip apart, so a
    sweep takes about N (own rows) + N (fill of the wavefront) lockstep trips per warp when published
    words are visible at once -- the dependency through the neighbouring chunk comes from the shuffle."""
    N = 16
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    F = orc.ldu_setup(orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr))
    b = np.random.default_rng(0).standard_normal(n)

    class InOrder:                       # warps scheduled round-robin, no visibility delay
        def permutation(self, live): return list(live)
        def integers(self, a, b): return 0

    y, trips = chunked_sweep_model(F.Lptr, F.Lnode, F.Lval, b, n, N, False, InOrder(), warp=N, max_delay=0)
    x, trips_b = chunked_sweep_model(F.Uptr, F.Unode, F.Uval, y / F.D, n, N, True, InOrder(), warp=N, max_delay=0)
    assert np.array_equal(x, orc.ldu_solve(F, b))
    assert trips <= 2 * N + 2 and trips_b <= 2 * N + 2, (trips, trips_b)


def test_symbolic_on_random_patterns_hypothesis(orc):
    """Random square patterns -- empty rows, missing diagonals, unsorted rows, dense rows -- through
    the host symbolic analysis: patterns equal the oracle's, every row appears once per schedule,
    and a row's level is one more than the highest level among the rows it reads."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 40), st.floats(0.0, 0.6), st.integers(0, 2**31 - 1))
    def check(n, density, seed):
        rng = np.random.default_rng(seed)
        mask = rng.random((n, n)) < density
        if rng.random() < 0.5:
            mask |= np.eye(n, dtype=bool)
        rows = []
        for i in range(n):
            cols = np.flatnonzero(mask[i]) + 1
            rows.append(rng.permutation(cols))                    # stored order is arbitrary
        ptr = np.concatenate([[1], 1 + np.cumsum([r.size for r in rows])]).astype(np.int32)
        node = (np.concatenate(rows) if ptr[-1] > 1 else np.zeros(0)).astype(np.int32)
        S = sb.ldu_symbolic(n, ptr, node)
        for i in range(n):
            row = node[ptr[i] - 1: ptr[i + 1] - 1]
            assert np.array_equal(S["Lnode"][S["Lptr"][i] - 1: S["Lptr"][i + 1] - 1], row[row < i + 1])
            assert np.array_equal(S["Unode"][S["Uptr"][i] - 1: S["Uptr"][i + 1] - 1], row[row > i + 1])
        # dest is a bijection onto [0, nL + nU) plus the diagonal slots of the rows that store one
        nL, nU = S["Lnode"].size, S["Unode"].size
        dest = S["dest"]
        assert dest.size == node.size and np.unique(dest).size == dest.size
        assert np.all((dest >= 0) & (dest < nL + nU + n))
        for rows_, lev, p_, nd in ((S["forward_rows"], S["forward_lev"], S["Lptr"], S["Lnode"]),
                                   (S["backward_rows"], S["backward_lev"], S["Uptr"], S["Unode"])):
            assert np.array_equal(np.sort(rows_), np.arange(1, n + 1))
            level_of = np.empty(n + 1, np.int64)
            for l in range(lev.size - 1):
                level_of[rows_[lev[l]:lev[l + 1]]] = l
            for i in range(1, n + 1):
                nb = nd[p_[i - 1] - 1: p_[i] - 1]
                assert level_of[i] == (level_of[nb].max() + 1 if nb.size else 0)

    check()


def test_chunked_sweep_model_on_random_patterns_hypothesis(orc):
    """The chunked sweep model on random patterns, warp widths, chunk lengths and visibility delays:
    it always terminates (every row depends on lower rows only, i.e. on its own thread's past or on
    lower threads) and reproduces the serial solve bit for bit."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(2, 60), st.floats(0.02, 0.5), st.sampled_from([1, 2, 4, 8]), st.integers(1, 9),
           st.integers(0, 4), st.integers(0, 2**31 - 1))
    def check(n, density, warp, chunk_rows, delay, seed):
        rng = np.random.default_rng(seed)
        mask = (rng.random((n, n)) < density) | np.eye(n, dtype=bool)
        rows = [rng.permutation(np.flatnonzero(mask[i]) + 1) for i in range(n)]
        ptr = np.concatenate([[1], 1 + np.cumsum([r.size for r in rows])]).astype(np.int32)
        node = np.concatenate(rows).astype(np.int32)
        val = rng.uniform(-1, 1, node.size)
        val[node == np.repeat(np.arange(1, n + 1), np.diff(ptr))] += n      # diagonally dominant: no pivot trouble
        F = orc.ldu_setup(orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr))
        b = rng.standard_normal(n)
        y, _ = chunked_sweep_model(F.Lptr, F.Lnode, F.Lval, b, n, chunk_rows, False, rng, warp=warp, max_delay=delay)
        x, _ = chunked_sweep_model(F.Uptr, F.Unode, F.Uval, y / F.D, n, chunk_rows, True, rng, warp=warp, max_delay=delay)
        assert np.array_equal(x, orc.ldu_solve(F, b))

    check()
