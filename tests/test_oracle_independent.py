"""Independent evidence that the oracle computes the right mathematics (the reference
itself cannot be built here: no Fortran compiler).  The restatement follows the reference
line by line; these tests check its RESULTS against implementations that share no code with
it -- scipy.sparse for products, transposes and conversions, a dense textbook IKJ
elimination for ILU(0), numpy's dense solver for the Krylov solutions -- on randomised
inputs (hypothesis), so that a transcription slip in the oracle cannot hide behind
self-consistency."""
import numpy as np
import scipy.sparse as sp
from hypothesis import given, settings, strategies as st

import oracle as orc

orc.build()


def random_csr(n, m, density, seed, ensure_rows=False):
    rng = np.random.default_rng(seed)
    mask = rng.random((n, m)) < density
    if ensure_rows:
        mask[np.arange(n), rng.integers(0, m, n)] = True
    r, c = np.nonzero(mask)
    ptr = np.concatenate([[1], 1 + np.cumsum(mask.sum(1))]).astype(np.int32)
    node = (c + 1).astype(np.int32)
    val = rng.standard_normal(node.size)
    for i in range(n):                                   # unsorted rows, like the reference's
        sl = slice(ptr[i] - 1, ptr[i + 1] - 1)
        perm = rng.permutation(sl.stop - sl.start)
        node[sl], val[sl] = node[sl][perm], val[sl][perm]
    S = sp.csr_matrix((val, node - 1, ptr - 1), shape=(n, m))
    return ptr, node, val, S


def as_orc(S):
    S = sp.csr_matrix(S)
    S.sort_indices()
    n, m = S.shape
    return orc.Matrix(orc.CSR, n, m, (S.indices + 1).astype(np.int32), S.data.astype(float),
                      ptr=(S.indptr + 1).astype(np.int32))


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 40), m=st.integers(1, 40), seed=st.integers(0, 10_000), dens=st.floats(0.05, 0.6))
def test_matvec_all_formats_against_scipy(n, m, seed, dens):
    ptr, node, val, S = random_csr(n, m, dens, seed, ensure_rows=True)
    # every column populated too, so that the transposed ellpack copy has no empty row
    A = orc.Matrix(orc.CSR, n, m, node, val, ptr=ptr)
    rng = np.random.default_rng(seed + 1)
    x, xt, y0 = rng.standard_normal(m), rng.standard_normal(n), rng.standard_normal(n)
    for fmt in (orc.CSR, orc.CSC, orc.ELL):
        B = orc.copy_matrix(A, fmt)
        assert np.allclose(orc.matvec(B, x), S @ x, rtol=0, atol=1e-12)
        assert np.allclose(orc.matvec(B, xt, trans=True), S.T @ xt, rtol=0, atol=1e-12)
        assert np.allclose(orc.matvec_add(B, x, y0), y0 + S @ x, rtol=0, atol=1e-12)
    for fmt in (orc.CSR, orc.CSC):
        T = orc.copy_matrix(A, fmt, trans=True)
        assert (T.nrow, T.ncol) == (m, n)
        assert np.allclose(orc.matvec(T, xt), S.T @ xt, rtol=0, atol=1e-12)


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 30), m=st.integers(1, 30), seed=st.integers(0, 10_000))
def test_copy_to_csc_has_scipys_structure(n, m, seed):
    """The csc copy of a csr source holds, column by column, the row ids in ascending order
    (= source iteration order): exactly scipy's canonical csc form; the values ride along."""
    ptr, node, val, S = random_csr(n, m, 0.3, seed)
    Cm = orc.copy_matrix(orc.Matrix(orc.CSR, n, m, node, val, ptr=ptr), orc.CSC)
    Sc = S.tocsc()
    Sc.sort_indices()
    assert np.array_equal(Cm.ptr - 1, Sc.indptr) and np.array_equal(Cm.node - 1, Sc.indices)
    assert np.array_equal(Cm.val, Sc.data)


def dense_ilu0(A):
    """Textbook IKJ ILU(0) on a dense copy, fill restricted to A's pattern, returned in the
    LDU form: strict L (unit diagonal implied), D, strict U (unit diagonal implied)."""
    n = A.shape[0]
    P = A != 0
    W = A.astype(float).copy()
    for i in range(1, n):
        for k in range(i):
            if P[i, k]:
                W[i, k] = W[i, k] / W[k, k]
                for j in range(k + 1, n):
                    if P[i, j]:
                        W[i, j] -= W[i, k] * W[k, j]
    D = np.diag(W).copy()
    return np.tril(W, -1), D, np.triu(W, 1) / D[:, None]


@settings(max_examples=20, deadline=None)
@given(n=st.integers(2, 30), seed=st.integers(0, 10_000))
def test_ldu_against_dense_textbook_ilu0(n, seed):
    """With each row's lower neighbours in ascending order the reference's elimination is the
    textbook IKJ ILU(0) written as L D U (ldu_solvers.f90:13-19)."""
    rng = np.random.default_rng(seed)
    mask = rng.random((n, n)) < 0.3
    mask |= mask.T
    np.fill_diagonal(mask, True)
    Ad = np.where(mask, rng.standard_normal((n, n)), 0.0)
    Ad += np.diag(np.abs(Ad).sum(1) + 1.0)                   # diagonally dominant: no pivot trouble
    F = orc.ldu_setup(as_orc(Ad))
    L, D, U = dense_ilu0(Ad)
    Lo = sp.csr_matrix((F.Lval, F.Lnode - 1, F.Lptr - 1), shape=(n, n)).toarray()
    Uo = sp.csr_matrix((F.Uval, F.Unode - 1, F.Uptr - 1), shape=(n, n)).toarray()
    assert np.allclose(Lo, L, rtol=1e-11, atol=1e-13) and np.allclose(Uo, U, rtol=1e-11, atol=1e-13)
    assert np.allclose(F.D, D, rtol=1e-11, atol=1e-13)
    b = rng.standard_normal(n)
    want = np.linalg.solve((np.eye(n) + L) @ np.diag(D) @ (np.eye(n) + U), b)
    assert np.allclose(orc.ldu_solve(F, b), want, rtol=1e-9, atol=1e-11)


@settings(max_examples=15, deadline=None)
@given(n=st.integers(3, 60), seed=st.integers(0, 10_000))
def test_krylov_solutions_against_dense_solve(n, seed):
    rng = np.random.default_rng(seed)
    mask = rng.random((n, n)) < 0.2
    W = np.where(mask, rng.standard_normal((n, n)), 0.0)
    Spd = W @ W.T + n * np.eye(n)                            # symmetric positive definite
    A = as_orc(Spd)
    xs = rng.standard_normal(n)
    b = Spd @ xs
    x, it, _, capped = orc.cg_solve(A, np.zeros(n), b, 1e-10, 50 * n)
    assert not capped and np.allclose(x, xs, rtol=0, atol=1e-8)
    xj, itj, _, capped = orc.cg_solve(A, np.zeros(n), b, 1e-10, 50 * n, idiag=orc.jacobi_setup(A))
    assert not capped and np.allclose(xj, xs, rtol=0, atol=1e-8)
    xl, itl, _, capped = orc.cg_solve_ldu(A, np.zeros(n), b, orc.ldu_setup(A), 1e-10, 50 * n)
    assert not capped and np.allclose(xl, xs, rtol=0, atol=1e-8)
    N = Spd + np.triu(np.where(mask, 0.3, 0.0), 1)           # nonsymmetric perturbation
    xb, itb, _, capped = orc.bicgstab_solve(as_orc(N), np.zeros(n), N @ xs, 1e-10, 50 * n)
    assert not capped and np.allclose(xb, xs, rtol=0, atol=1e-7)


@settings(max_examples=15, deadline=None)
@given(n=st.integers(4, 40), seed=st.integers(0, 10_000))
def test_lanczos_ritz_values_against_numpy(n, seed):
    """n full Lanczos steps with re-orthogonalisation reproduce the spectrum."""
    rng = np.random.default_rng(seed)
    W = np.where(rng.random((n, n)) < 0.3, rng.standard_normal((n, n)), 0.0)
    Sy = W + W.T + np.diag(np.arange(1.0, n + 1))
    info, lam, V = orc.eigensolve(as_orc(Sy), n, rng.uniform(-1, 1, n))
    assert info == 0
    assert np.allclose(np.sort(lam), np.linalg.eigvalsh(Sy), rtol=1e-7, atol=1e-7)
