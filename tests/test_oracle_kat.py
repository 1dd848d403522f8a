"""Pins the oracle on the reference's own known-answer tests and checks the
vectorised generators against the oracle's restatement of the graph builders.
CPU only."""
import json
import os

import numpy as np

from helpers import csr_from_calls, ell_from_tridiag_calls
from sigma_b200 import generators as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_kat_diffusion_1d_cg_ellpack(orc):
    """test/solver_test_diffusion_1d.f90: nn=127, CG(1e-16) on the ELLPACK
    Laplacian, u0 = 0, f = 2 dx^2, exact v = x(1-x); bar :111-120 is 1e-14."""
    nn = 127
    dx = 1.0 / (nn + 1)
    node, deg, val = ell_from_tridiag_calls(orc, nn, 2.0, -1.0, -1.0)
    assert node[0].tolist() == [1, 2, 2] and node[-1].tolist() == [nn - 1, nn, nn]
    A = orc.Matrix(orc.ELL, nn, nn, node, val, degrees=deg)
    f = np.full(nn, 2.0 * dx**2)
    v = np.array([i * dx * (1.0 - i * dx) for i in range(1, nn + 1)])
    u, it, res2, capped = orc.cg_solve(A, np.zeros(nn), f, 1e-16)
    assert not capped
    assert np.abs(u - v).max() <= 1e-14
    gold = json.load(open(os.path.join(GOLD, "kat.json")))["diffusion_1d"]
    assert it == gold["iterations"] == 64
    assert res2 == gold["res2"] == 0.0
    assert np.abs(u - v).max() == gold["misfit"] == 0.0


def test_kat_diffusion_1d_cg_csr(orc):
    nn = 127
    dx = 1.0 / (nn + 1)
    ptr, node, val = G.tridiag_csr(nn)
    A = orc.Matrix(orc.CSR, nn, nn, node, val, ptr=ptr)
    v = np.array([i * dx * (1.0 - i * dx) for i in range(1, nn + 1)])
    u, it, res2, _ = orc.cg_solve(A, np.zeros(nn), np.full(nn, 2.0 * dx**2), 1e-16)
    assert it == 64 and res2 == 0.0 and np.abs(u - v).max() == 0.0


def test_kat_advection_diffusion_1d_bicgstab(orc):
    """test/solver_test_advection_diffusion_1d.f90: nn=1024, c=0.5,
    BiCGSTAB(1e-12); bar :118-127 is 1e-8."""
    nn, c = 1024, 0.5
    dx = 1.0 / (nn + 1)
    node, deg, val = ell_from_tridiag_calls(orc, nn, 2.0, -1.0 + c * dx / 2, -1.0 - c * dx / 2)
    A = orc.Matrix(orc.ELL, nn, nn, node, val, degrees=deg)
    x = np.arange(1, nn + 1) * dx
    v = 2.0 * (x - (np.exp(c * x) - 1) / (np.exp(c) - 1)) / c
    u, it, res2, capped = orc.bicgstab_solve(A, np.zeros(nn), np.full(nn, 2.0 * dx**2), 1e-12)
    assert not capped and np.sqrt(res2) <= 1e-12
    misfit = np.abs(u - v).max()
    assert misfit <= 1e-8
    gold = json.load(open(os.path.join(GOLD, "kat.json")))["advection_diffusion_1d"]
    assert it == gold["iterations"] == 1133
    assert abs(misfit - gold["misfit"]) <= 1e-15


def test_tridiag_generators_match_reference_build(orc):
    for nn in (2, 3, 17, 127):
        node, deg, val = ell_from_tridiag_calls(orc, nn, 2.0, -0.75, -1.25)
        gnode, gdeg, gval = G.tridiag_ell(nn, 2.0, -0.75, -1.25)
        assert np.array_equal(node, gnode) and np.array_equal(deg, gdeg) and np.array_equal(val, gval)
        ei, ej = G.tridiag_add_edge_calls(nn)
        ptr, cnode = csr_from_calls(orc, nn, ei, ej)
        gptr, gcnode, _ = G.tridiag_csr(nn)
        assert np.array_equal(ptr, gptr) and np.array_equal(cnode, gcnode)


def test_poisson_generator_matches_reference_build(orc):
    for N in (2, 3, 5, 8):
        ei, ej = G.poisson2d_add_edge_calls(N)
        ptr, node = csr_from_calls(orc, N * N, ei, ej)
        gptr, gnode, gval = G.poisson2d_csr(N)
        assert np.array_equal(ptr, gptr) and np.array_equal(node, gnode)
        rows = np.repeat(np.arange(1, N * N + 1), np.diff(gptr))
        assert np.array_equal(gval, np.where(gnode == rows, 4.0, -1.0))
    # interior row order [k-N, k-1, k, k+1, k+N]
    N = 5
    ptr, node, _ = G.poisson2d_csr(N)
    k = 13
    assert node[ptr[k - 1] - 1: ptr[k] - 1].tolist() == [k - N, k - 1, k, k + 1, k + N]
    # row blocks agree with the full matrix
    lo, hi = 7, 19
    bptr, bnode, bval = G.poisson2d_csr(N, lo, hi)
    assert np.array_equal(bnode, node[ptr[lo] - 1: ptr[hi] - 1])
    assert np.array_equal(bptr - 1, ptr[lo:hi + 1] - ptr[lo])


def test_poisson_rhs_matches_oracle_matvec(orc):
    N = 16
    ptr, node, val = G.poisson2d_csr(N)
    A = orc.Matrix(orc.CSR, N * N, N * N, node, val, ptr=ptr)
    b, xs = G.poisson2d_rhs(N)
    assert np.array_equal(orc.matvec(A, xs), b)


def test_erdos_renyi_generator_matches_reference_build(orc):
    n = 60
    (ptr, node, val, (i, j)) = G.erdos_renyi_csr(n, p=0.1, seed=3, return_pairs=True)
    ei, ej = G.erdos_renyi_add_edge_calls(n, i, j)
    rptr, rnode = csr_from_calls(orc, n, ei, ej)
    assert np.array_equal(ptr, rptr) and np.array_equal(node, rnode)
    # Laplacian + I: row sums are 1
    rows = np.repeat(np.arange(n), np.diff(ptr))
    assert np.allclose(np.bincount(rows, val, n), 1.0)


def test_csr_to_ell_matches_reference_build(orc):
    n = 40
    ptr, node, val, (i, j) = G.erdos_renyi_csr(n, p=0.15, seed=5, weights="random", return_pairs=True)
    ei, ej = G.erdos_renyi_add_edge_calls(n, i, j)
    si, sj, _ = orc.ll_graph_edges(n, ei, ej)
    enode, edeg = orc.ellpack_graph_build(n, si, sj)
    gnode, gdeg, gval = G.csr_to_ell(ptr, node, val)
    assert np.array_equal(enode, gnode) and np.array_equal(edeg, gdeg)
    eval_ = np.zeros(enode.shape)
    rows = np.repeat(np.arange(1, n + 1), np.diff(ptr))
    for r, c, v in zip(rows, node, val):
        assert orc.ell_set_value(enode, edeg, eval_, int(r), int(c), float(v))
    assert np.array_equal(eval_, gval)


def test_transposed_copy_matches_cs_graph_build_trans(orc):
    """copy_matrix(trans) -> cs_graph_build(trans=.true.) (cs_graphs.f90:122-183):
    the arrays the device transpose must reproduce."""
    n = 50
    ptr, node, val = G.erdos_renyi_csr(n, p=0.12, seed=11, weights="random", skew=True)
    rows = np.repeat(np.arange(1, n + 1, dtype=np.int32), np.diff(ptr))
    tptr, tnode, _ = orc.cs_graph_build(n, rows, node, trans=True)
    gptr, gnode, gval = G.csr_transpose(n, n, ptr, node, val)
    assert np.array_equal(tptr, gptr) and np.array_equal(tnode, gnode)
    # and the CSC kernel on those arrays equals the CSR kernel on the original
    x = np.random.default_rng(0).random(n)
    A = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    Ac = orc.Matrix(orc.CSC, n, n, gnode, gval, ptr=gptr)
    assert np.allclose(orc.matvec(A, x), orc.matvec(Ac, x), rtol=1e-14)


def test_matvec_all_formats_vs_dense(orc):
    """test/matrix_test_basics.f90:332-362: matvec and matvec_t vs dense, 1e-15."""
    n = 64
    ptr, node, val = G.erdos_renyi_csr(n, p=0.1, seed=21, weights="random", skew=True)
    B = G.dense_from_csr(n, n, ptr, node, val)
    x = np.random.default_rng(1).random(n)
    en, ed, ev = G.csr_to_ell(ptr, node, val)
    tp, tn, tv = G.csr_transpose(n, n, ptr, node, val)
    mats = [orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr), orc.Matrix(orc.CSC, n, n, tn, tv, ptr=tp),
            orc.Matrix(orc.ELL, n, n, en, ev, degrees=ed)]
    for A in mats:
        y, yt = orc.matvec(A, x), orc.matvec(A, x, trans=True)
        assert np.abs(y - B @ x).max() / np.abs(B @ x).max() < 1e-15 * 8
        assert np.abs(yt - B.T @ x).max() / np.abs(B.T @ x).max() < 1e-15 * 8


def test_fem_generator_is_symmetric_laplacian(orc):
    N = 9
    ptr, node, val = G.fem_p1_csr(N)
    n = N * N
    B = G.dense_from_csr(n, n, ptr, node, val)
    assert np.allclose(B, B.T, atol=1e-13)
    assert np.linalg.eigvalsh(B).min() > 0
    # interior rows annihilate constants up to the Dirichlet columns we zeroed
    A = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    y = orc.matvec(A, np.ones(n))
    interior = np.zeros((N, N), bool)
    interior[2:-2, 2:-2] = True
    assert np.abs(y[interior.reshape(-1)]).max() < 1e-12


def test_jacobi_pcg_manufactured(orc):
    """test/solver_test_jacobi.f90:138-222 shape: Jacobi-PCG on a random ER
    graph Laplacian + I, manufactured solution."""
    n = 200
    ptr, node, val = G.erdos_renyi_csr(n, seed=31, weights="random")
    A = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    v = np.random.default_rng(2).random(n)
    f = orc.matvec(A, v)
    idiag = orc.jacobi_setup(A)
    rows = np.repeat(np.arange(n), np.diff(ptr))
    assert np.array_equal(idiag, 1.0 / val[node - 1 == rows])
    u, it, res2, capped = orc.cg_solve(A, np.zeros(n), f, 1e-14, max_iter=10 * n, idiag=idiag)
    assert not capped and np.abs(u - v).max() < 1e-12


def test_lanczos_identities(orc):
    """test/eigensolver_test_lanczos.f90:130-170: three-term recurrence and
    orthogonality to 1e-14."""
    n, nq = 128, 11
    ptr, node, val = G.erdos_renyi_csr(n, seed=41, shift=0.0)
    A = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    q1 = 2 * np.random.default_rng(3).random(n) - 1
    T, V = orc.lanczos(A, nq, q1)
    for i in range(1, nq - 1):
        x = orc.matvec(A, V[:, i])
        y = T[1, i] * V[:, i] + T[0, i - 1] * V[:, i - 1] + T[2, i] * V[:, i + 1]
        assert np.sqrt(((y - x) ** 2).sum() / (x**2).sum()) < 1e-14
    Qm = V.T @ V - np.eye(nq)
    assert np.sqrt((Qm**2).sum()) / nq < 1e-14
    info, lam, W = orc.eigensolve(A, nq, q1)
    assert info == 0
    ref = np.linalg.eigvalsh(np.diag(T[1]) + np.diag(T[2, :-1], 1) + np.diag(T[2, :-1], -1))
    assert np.allclose(lam, ref, rtol=1e-12, atol=1e-12)


def test_generalized_lanczos_reference_test(orc):
    """test/eigensolver_test_generalized_lanczos.f90 with its own deterministic
    matrices (48 x 32 periodic P1 grid, stiffness A and mass B, B%set_solver(cg(1e-15)),
    nq = 48): three-term recurrence A v_i = alpha_i B v_i + beta_{i-1} B v_{i-1} +
    beta_i B v_{i+1} and B-orthogonality, both 1e-14 (:168-171,197-200)."""
    ptr, node, vA, vB = G.periodic_p1_grid()
    nn, nq = 48 * 32, 48
    assert np.diff(ptr).max() == 7
    A = orc.Matrix(orc.CSR, nn, nn, node, vA, ptr=ptr)
    B = orc.Matrix(orc.CSR, nn, nn, node, vB, ptr=ptr)
    q1 = 2 * np.random.default_rng(0).random(nn) - 1
    T, V, inner = orc.generalized_lanczos(A, B, nq, q1, 1e-15, 10000)
    assert inner > 0
    U = np.stack([orc.matvec(B, V[:, i]) for i in range(nq)], 1)
    for i in range(1, nq - 1):
        w = orc.matvec(A, V[:, i])
        z = T[1, i] * U[:, i] + T[0, i - 1] * U[:, i - 1] + T[2, i] * U[:, i + 1]
        assert np.sqrt(((w - z) ** 2).sum() / (w * w).sum()) <= 1e-14
    Qm = V.T @ U - np.eye(nq)
    assert np.sqrt((Qm**2).sum()) / nq <= 1e-14
