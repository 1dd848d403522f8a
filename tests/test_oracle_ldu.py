"""The oracle's restatement of the ILDU(0) solver (src/solver/ldu_solvers.f90) held to the
reference's own test, test/solver_test_incomplete_cholesky.f90, on CPU: a random weighted
graph Laplacian + I (nn = 128); the factorisation used as a stationary solver reaches
1e-14 in 10 nn sweeps (:182-202) and as a preconditioner of CG(1e-16) reaches 1e-15
(:213-226).  The reference draws its graph from a time-seeded RNG: seeded inputs of the
same shape here."""
import numpy as np
import pytest

from sigma_b200 import generators as G


def reference_case(orc, nn, seed):
    ptr, node, val = G.erdos_renyi_csr(nn, seed=seed, weights="random", shift=1.0)
    A = orc.Matrix(orc.CSR, nn, nn, node, val, ptr=ptr)
    F = orc.ldu_setup(A)
    rng = np.random.default_rng(seed)
    v = rng.random(nn)
    r = v - orc.matvec(A, v)                 # :168-170: smooth v with one preconditioner application
    v = orc.ldu_solve(F, r)
    f = orc.matvec(A, v)
    return ptr, node, val, A, F, v, f


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_incomplete_cholesky_like_the_reference(orc, seed):
    nn = 128
    ptr, node, val, A, F, v, f = reference_case(orc, nn, seed)
    # stationary iteration u += M^-1 (f - A u)
    u, r = np.zeros(nn), f.copy()
    for _ in range(10 * nn):
        u = u + orc.ldu_solve(F, r)
        r = f - orc.matvec(A, u)
    assert np.abs(u - v).max() <= 1e-14
    # preconditioned CG with the reference's tolerance
    x, it, res2, capped = orc.cg_solve_ldu(A, np.zeros(nn), f, F, 1e-16, 100 * nn)
    assert not capped and np.abs(x - v).max() <= 1e-15
    x0, it0, _, _ = orc.cg_solve(A, np.zeros(nn), f, 1e-16, 100 * nn)
    assert it < it0


def test_ilu0_identities(orc):
    """What ILDU(0) means: (I + L) D (I + U) reproduces A on A's own pattern, the factors
    live on the strict triangles of that pattern, and U = L^T for a symmetric A."""
    nn = 200
    ptr, node, val, A, F, _, _ = reference_case(orc, nn, 7)
    Ad = G.dense_from_csr(nn, nn, ptr, node, val)
    L = G.dense_from_csr(nn, nn, F.Lptr, F.Lnode, F.Lval)
    U = G.dense_from_csr(nn, nn, F.Uptr, F.Unode, F.Uval)
    assert np.array_equal(L != 0, np.tril(Ad != 0, -1)) and np.array_equal(U != 0, np.triu(Ad != 0, 1))
    P = (np.eye(nn) + L) @ np.diag(F.D) @ (np.eye(nn) + U)
    assert np.abs(P - Ad)[Ad != 0].max() <= 1e-13
    assert np.abs(U - L.T).max() <= 1e-15
    # rows keep A's stored order (incomplete_ldu_sparsity_pattern walks A's iterator, :421-433)
    for i in (0, 17, nn - 1):
        row = node[ptr[i] - 1: ptr[i + 1] - 1]
        assert np.array_equal(F.Lnode[F.Lptr[i] - 1: F.Lptr[i + 1] - 1], row[row < i + 1])
        assert np.array_equal(F.Unode[F.Uptr[i] - 1: F.Uptr[i + 1] - 1], row[row > i + 1])


def test_nonsymmetric_and_unsorted_rows(orc):
    """A nonsymmetric operator with shuffled rows: the elimination walks each row's lower
    neighbours in STORED order (ldu_solvers.f90:339), which the restatement must follow; on a
    tridiagonal matrix (no fill at all) ILDU(0) is the exact LDU factorisation."""
    nn = 64
    ptr, node, val = G.tridiag_csr(nn, 2.5, -1.0, -0.5)
    A = orc.Matrix(orc.CSR, nn, nn, node, val, ptr=ptr)
    F = orc.ldu_setup(A)
    xs = np.random.default_rng(0).standard_normal(nn)
    assert np.abs(orc.ldu_solve(F, orc.matvec(A, xs)) - xs).max() <= 1e-13
    # shuffle inside the rows: same matrix, different stored order -> same factors up to
    # rounding (for a tridiagonal matrix each row has one lower neighbour: identical)
    rng = np.random.default_rng(1)
    node2, val2 = node.copy(), val.copy()
    for i in range(nn):
        sl = slice(ptr[i] - 1, ptr[i + 1] - 1)
        perm = rng.permutation(sl.stop - sl.start)
        node2[sl], val2[sl] = node[sl][perm], val[sl][perm]
    F2 = orc.ldu_setup(orc.Matrix(orc.CSR, nn, nn, node2, val2, ptr=ptr))
    assert np.array_equal(F2.D, F.D)
    assert np.abs(orc.ldu_solve(F2, orc.matvec(A, xs)) - xs).max() <= 1e-13
    # csc and ellpack sources give the factors of the same matrix (their iterators differ:
    # a csc matrix streams column by column, so L's rows come out in ascending column order)
    cptr, cnode, cval = G.csr_transpose(nn, nn, ptr, node2, val2)
    Fc = orc.ldu_setup(orc.Matrix(orc.CSC, nn, nn, cnode, cval, ptr=cptr))
    assert np.array_equal(Fc.D, F.D)
    enode, edeg, eval_ = G.csr_to_ell(ptr, node2, val2)
    Fe = orc.ldu_setup(orc.Matrix(orc.ELL, nn, nn, enode, eval_, degrees=edeg))
    assert np.array_equal(Fe.D, F.D)


def dense_bicgstab_pc(Ad, Minv, b, tol, cap):
    """bicgstab_solve_pc (bicgstab_solvers.f90:182-237) written independently with dense numpy
    operators and np.dot: same recurrence, different summation order."""
    n = b.size
    x = np.zeros(n)
    r0 = Minv(b - Ad @ x)
    r = r0.copy()
    rho_old = alpha = omega = 1.0
    v, p = np.zeros(n), np.zeros(n)
    it = 0
    while np.sqrt(r @ r) > tol and it < cap:
        rho = r0 @ r
        beta = rho / rho_old * alpha / omega
        p = r + beta * (p - omega * v)
        v = Minv(Ad @ p)
        alpha = rho / (r0 @ v)
        s = r - alpha * v
        t = Minv(Ad @ s)
        omega = (s @ t) / (t @ t)
        x = x + alpha * p + omega * s
        r = s - omega * t
        rho_old = rho
        it += 1
    return x, it


@pytest.mark.parametrize("seed", [4, 5])
def test_bicgstab_with_the_ldu_preconditioner(orc, seed):
    """bicgstab_solve_pc takes any linear_solver as pc; with pc = ldu() on a nonsymmetric
    operator it must (a) solve the system, (b) need fewer iterations than the bare solver and
    (c) follow the same recurrence as an independent dense transcription."""
    nn = 300
    ptr, node, val = G.erdos_renyi_csr(nn, seed=seed, weights="random", skew=True, shift=1.0)
    A = orc.Matrix(orc.CSR, nn, nn, node, val, ptr=ptr)
    F = orc.ldu_setup(A)
    v = np.random.default_rng(seed).random(nn)
    f = orc.matvec(A, v)
    tol = 1e-13
    x, it, res2, capped = orc.bicgstab_solve_ldu(A, np.zeros(nn), f, F, tol, 50 * nn)
    assert not capped and np.sqrt(res2) <= tol
    assert np.abs(x - v).max() <= 1e-11
    x0, it0, _, capped0 = orc.bicgstab_solve(A, np.zeros(nn), f, tol, 50 * nn)
    assert not capped0 and it < it0
    Ad = G.dense_from_csr(nn, nn, ptr, node, val)
    xd, itd = dense_bicgstab_pc(Ad, lambda z: orc.ldu_solve(F, z), f, tol, 50 * nn)
    assert abs(it - itd) <= max(1, int(np.ceil(0.05 * itd)))
    assert np.abs(x - xd).max() <= 1e-10 * np.abs(xd).max()
