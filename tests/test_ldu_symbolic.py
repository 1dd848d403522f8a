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
