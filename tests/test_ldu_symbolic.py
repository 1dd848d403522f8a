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


def syncfree_sweep_model(rows, ptr, node, val, start, n, grid_threads, rng, warp=4):
    """Model of tri_syncfree_kernel's control flow (csrc/ldu.cu, EXPERIMENTAL): positions in
    level order dealt to `grid_threads` resident threads grid-stride, a warp = `warp`
    consecutive lanes running ONE loop in which every unfinished lane polls once and advances
    as far as it can; warps are scheduled in random order, one loop trip at a time.  Returns
    the solution and the number of trips of the busiest warp; raises if no warp can progress."""
    x = np.full(n, np.nan)
    published = np.zeros(n, bool)
    nwarps = grid_threads // warp
    state = []
    for w in range(nwarps):
        state.append({"base": w * warp, "lanes": None, "trips": 0})

    def load(st):
        lanes = []
        for lane in range(warp):
            p_ = st["base"] + lane
            if p_ < n:
                i = rows[p_]
                lanes.append({"i": i, "k": ptr[i - 1] - 1, "e": ptr[i] - 1, "z": start[i - 1], "done": False})
            else:
                lanes.append({"done": True})
        st["lanes"] = lanes

    live = [w for w in range(nwarps) if state[w]["base"] < n]
    for w in live:
        load(state[w])
    idle_rounds = 0
    while live:
        progressed = False
        for w in rng.permutation(live):
            st = state[w]
            st["trips"] += 1
            for ln in st["lanes"]:
                if ln["done"]:
                    continue
                while ln["k"] < ln["e"] and published[node[ln["k"]] - 1]:
                    ln["z"] = ln["z"] - val[ln["k"]] * x[node[ln["k"]] - 1]
                    ln["k"] += 1
                    progressed = True
                if ln["k"] == ln["e"]:
                    x[ln["i"] - 1] = ln["z"]
                    published[ln["i"] - 1] = True
                    ln["done"] = True
                    progressed = True
            if all(ln["done"] for ln in st["lanes"]):
                st["base"] += grid_threads
                if st["base"] < n:
                    load(st)
                else:
                    live = [v for v in live if v != w]
        idle_rounds = 0 if progressed else idle_rounds + 1
        assert idle_rounds < 2, "no warp can make progress: the sweep would hang"
    return x, max(st["trips"] for st in state)


@pytest.mark.parametrize("case", list(cases()), ids=lambda c: c[0])
@pytest.mark.parametrize("grid_threads", [4, 16, 64])
def test_syncfree_sweep_model_terminates_and_matches(orc, case, grid_threads):
    """The opt-in sync-free sweeps: under random warp scheduling, with levels cutting through
    warps and fewer resident threads than rows, every sweep terminates and gives the serial
    result bit for bit."""
    _, n, ptr, node, val = case
    S = sb.ldu_symbolic(n, ptr, node)
    F = orc.ldu_setup(orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr))
    rng = np.random.default_rng(grid_threads)
    b = rng.standard_normal(n)
    y, _ = syncfree_sweep_model(S["forward_rows"], F.Lptr, F.Lnode, F.Lval, b, n, grid_threads, rng)
    x, _ = syncfree_sweep_model(S["backward_rows"], F.Uptr, F.Unode, F.Uval, y / F.D, n, grid_threads, rng)
    assert np.array_equal(x, orc.ldu_solve(F, b))


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


def test_syncfree_sweep_model_on_random_patterns_hypothesis(orc):
    """The sync-free sweep model on random patterns, warp widths and resident-thread counts: it
    always terminates (no schedule can starve the lowest unfinished position) and reproduces the
    serial solve bit for bit."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(2, 60), st.floats(0.02, 0.5), st.sampled_from([1, 2, 4, 8]), st.integers(1, 6),
           st.integers(0, 2**31 - 1))
    def check(n, density, warp, warps, seed):
        rng = np.random.default_rng(seed)
        mask = (rng.random((n, n)) < density) | np.eye(n, dtype=bool)
        rows = [rng.permutation(np.flatnonzero(mask[i]) + 1) for i in range(n)]
        ptr = np.concatenate([[1], 1 + np.cumsum([r.size for r in rows])]).astype(np.int32)
        node = np.concatenate(rows).astype(np.int32)
        val = rng.uniform(-1, 1, node.size)
        val[node == np.repeat(np.arange(1, n + 1), np.diff(ptr))] += n      # diagonally dominant: no pivot trouble
        S = sb.ldu_symbolic(n, ptr, node)
        F = orc.ldu_setup(orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr))
        b = rng.standard_normal(n)
        g = warp * warps
        y, _ = syncfree_sweep_model(S["forward_rows"], F.Lptr, F.Lnode, F.Lval, b, n, g, rng, warp=warp)
        x, _ = syncfree_sweep_model(S["backward_rows"], F.Uptr, F.Unode, F.Uval, y / F.D, n, g, rng, warp=warp)
        assert np.array_equal(x, orc.ldu_solve(F, b))

    check()
