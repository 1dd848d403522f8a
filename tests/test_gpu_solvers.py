"""Parity of the device-resident Krylov solvers against the oracle and the
reference's own test bars.  north_star bars: CG/BiCGSTAB reach the same
tolerance within +-2 % iterations, solution relative error <= 1e-10."""
import json
import os

import numpy as np
import pytest

from sigma_b200 import generators as G

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


def within(it, ref, frac=0.02):
    return abs(it - ref) <= max(1, int(np.ceil(frac * ref)))


@pytest.mark.parametrize("fmt", ["ellpack", "csr", "csc"])
def test_kat_diffusion_1d(sb, fmt):
    """test/solver_test_diffusion_1d.f90 as the reference runs it (ELLPACK) and
    in CSR/CSC: CG(1e-16), bar 1e-14 (:111-120); oracle: 64 iterations."""
    nn = 127
    dx = 1.0 / (nn + 1)
    ptr, node, val = G.tridiag_csr(nn)
    if fmt == "ellpack":
        A = sb.ellpack_matrix(nn, nn, *G.tridiag_ell(nn))
    elif fmt == "csr":
        A = sb.csr_matrix(nn, nn, ptr, node, val)
    else:
        A = sb.csc_matrix(nn, nn, *G.csr_transpose(nn, nn, ptr, node, val))
    u = np.zeros(nn)
    f = np.full(nn, 2.0 * dx**2)
    v = np.array([i * dx * (1.0 - i * dx) for i in range(1, nn + 1)])
    solver = sb.cg(1e-16)
    solver.set_max_iterations(20 * nn)   # safety net only; must not trigger
    solver.setup(A)
    u = solver.solve(A, u, f)
    it, res2, capped = solver.info()
    assert not capped
    assert np.abs(u - v).max() <= 1e-14
    assert within(it, GOLD["diffusion_1d"]["iterations"])


def test_kat_advection_diffusion_1d(sb):
    """test/solver_test_advection_diffusion_1d.f90: BiCGSTAB(1e-12) on the
    nonsymmetric ELLPACK operator, bar 1e-8 (:118-127).  The iteration count of
    this ill-conditioned case moves by ~7 % with the dot-product summation
    order alone, so only the reference's own bar is asserted."""
    nn, c = 1024, 0.5
    dx = 1.0 / (nn + 1)
    A = sb.ellpack_matrix(nn, nn, *G.tridiag_ell(nn, 2.0, -1.0 + c * dx / 2, -1.0 - c * dx / 2))
    x = np.arange(1, nn + 1) * dx
    v = 2.0 * (x - (np.exp(c * x) - 1) / (np.exp(c) - 1)) / c
    solver = sb.bicgstab(1e-12)
    solver.set_max_iterations(20 * nn)
    solver.setup(A)
    u = solver.solve(A, np.zeros(nn), np.full(nn, 2.0 * dx**2))
    it, res2, capped = solver.info()
    assert not capped and np.sqrt(res2) <= 1e-12
    assert np.abs(u - v).max() <= 1e-8
    assert within(it, GOLD["advection_diffusion_1d"]["iterations"], 0.15)


@pytest.mark.parametrize("N", [64, 128, 256])
def test_cg_poisson_iterations_and_solution(sb, orc, N):
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    b, xs = G.poisson2d_rhs(N)
    tol = 1e-10 * np.linalg.norm(b)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    xo, ito, _, _ = orc.cg_solve(O, np.zeros(n), b, tol)
    A = sb.csr_matrix(n, n, ptr, node, val)
    s = sb.cg(tol)
    s.setup(A)
    x = s.solve(A, np.zeros(n), b)
    it, res2, capped = s.info()
    assert not capped and np.sqrt(res2) <= tol
    assert within(it, ito), (it, ito)
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) <= 1e-10
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) <= 1e-8
    # same operator in ELLPACK
    E = sb.ellpack_matrix(n, n, *G.csr_to_ell(ptr, node, val))
    s2 = sb.cg(tol)
    s2.setup(E)
    x2 = s2.solve(E, np.zeros(n), b)
    assert within(s2.iterations, ito)
    assert np.linalg.norm(x2 - xo) / np.linalg.norm(xo) <= 1e-10


def test_jacobi_setup_and_apply_bit_exact(sb, orc):
    n = 3000
    ptr, node, val = G.erdos_renyi_csr(n, seed=6, weights="random")
    b = np.random.default_rng(7).standard_normal(n)
    for fmt in ("csr", "csc", "ellpack"):
        if fmt == "csr":
            A, O = sb.csr_matrix(n, n, ptr, node, val), orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
        elif fmt == "csc":
            cp, cn, cv = G.csr_transpose(n, n, ptr, node, val)
            A, O = sb.csc_matrix(n, n, cp, cn, cv), orc.Matrix(orc.CSC, n, n, cn, cv, ptr=cp)
        else:
            en, ed, ev = G.csr_to_ell(ptr, node, val)
            A, O = sb.ellpack_matrix(n, n, en, ed, ev), orc.Matrix(orc.ELL, n, n, en, ev, degrees=ed)
        pc = sb.jacobi()
        pc.setup(A)
        idiag = orc.jacobi_setup(O)
        assert np.array_equal(pc.vector("idiag"), idiag)
        assert np.array_equal(pc.solve(A, np.zeros(n), b), orc.jacobi_solve(idiag, b))


def test_jacobi_pcg_like_reference_test(sb, orc):
    """test/solver_test_jacobi.f90:138-222: random-weight ER Laplacian + I in
    CSR, manufactured solution, Jacobi-PCG; reference bar 1e-15 on max|u-v|
    with tol 1e-16 -- we run tol 1e-14 with the oracle beside it."""
    n = 4096
    ptr, node, val = G.erdos_renyi_csr(n, seed=8, weights="random")
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    v = np.random.default_rng(9).random(n)
    f = orc.matvec(O, v)
    A = sb.csr_matrix(n, n, ptr, node, val)
    solver, pc = sb.cg(1e-14), sb.jacobi()
    solver.setup(A)
    pc.setup(A)
    solver.set_max_iterations(10 * n)
    u = solver.solve(A, np.zeros(n), f, pc)
    it, res2, capped = solver.info()
    uo, ito, _, _ = orc.cg_solve(O, np.zeros(n), f, 1e-14, idiag=orc.jacobi_setup(O))
    assert not capped and within(it, ito), (it, ito)
    assert np.abs(u - v).max() <= 1e-12
    assert np.linalg.norm(u - uo) / np.linalg.norm(uo) <= 1e-10


def test_jacobi_pcg_on_csc_like_solver_example_1(sb, orc):
    """examples/solvers/solver_example_1.f90:21,114-119: Jacobi-PCG on a CSC matrix."""
    n = 2500
    ptr, node, val = G.erdos_renyi_csr(n, seed=10, weights="unit", shift=1.0)
    cp, cn, cv = G.csr_transpose(n, n, ptr, node, val)
    A = sb.csc_matrix(n, n, cp, cn, cv)
    O = orc.Matrix(orc.CSC, n, n, cn, cv, ptr=cp)
    b = np.ones(n)
    solver, pc = sb.cg(1e-12), sb.jacobi()
    solver.setup(A)
    pc.setup(A)
    x = solver.solve(A, np.zeros(n), b, pc)
    xo, ito, _, _ = orc.cg_solve(O, np.zeros(n), b, 1e-12, idiag=orc.jacobi_setup(O))
    assert within(solver.iterations, ito)
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) <= 1e-10


@pytest.mark.parametrize("pc_on", [False, True])
def test_bicgstab_nonsymmetric(sb, orc, pc_on):
    """test/solver_test_jacobi.f90:240-291: skew-perturbed ER operator,
    BiCGSTAB with and without the Jacobi preconditioner."""
    n = 4096
    ptr, node, val = G.erdos_renyi_csr(n, seed=12, weights="random", skew=True)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    v = np.random.default_rng(13).random(n)
    f = orc.matvec(O, v)
    A = sb.csr_matrix(n, n, ptr, node, val)
    solver = sb.bicgstab(1e-13)
    solver.setup(A)
    solver.set_max_iterations(10 * n)
    pc = None
    idiag = None
    if pc_on:
        pc = sb.jacobi()
        pc.setup(A)
        idiag = orc.jacobi_setup(O)
    u = solver.solve(A, np.zeros(n), f, pc)
    it, res2, capped = solver.info()
    uo, ito, _, _ = orc.bicgstab_solve(O, np.zeros(n), f, 1e-13, idiag=idiag)
    assert not capped and np.sqrt(res2) <= 1e-13
    assert within(it, ito, 0.05), (it, ito)
    assert np.abs(u - v).max() <= 1e-11
    assert np.linalg.norm(u - uo) / np.linalg.norm(uo) <= 1e-10


def test_solver_bookkeeping(sb):
    n = 32 * 32
    ptr, node, val = G.poisson2d_csr(32)
    b, _ = G.poisson2d_rhs(32)
    A = sb.csr_matrix(n, n, ptr, node, val)
    s = sb.cg(1e-8)
    s.setup(A)
    x = s.solve(A, np.zeros(n), b)
    it1 = s.iterations
    assert it1 > 0
    # iterations accumulate across solves (cg_solvers.f90:145) ...
    s.solve(A, np.zeros(n), b)
    assert s.iterations == 2 * it1
    # ... a converged initial guess does no iteration (loop test first, :133) ...
    s.solve(A, x, b)
    assert s.iterations <= 2 * it1 + 2
    # ... and setup resets the counter (:72)
    s.setup(A)
    assert s.iterations == 0
    # the safety cap is reported, not silent
    s.set_max_iterations(3)
    s.solve(A, np.zeros(n), b)
    assert s.info()[0] == 3 and s.info()[2] is True
    # zero right-hand side: res2 = 0, no iteration
    s.set_max_iterations(-1)
    s.setup(A)
    x0 = s.solve(A, np.zeros(n), np.zeros(n))
    assert s.iterations == 0 and not x0.any()
    # default tolerance is 1e-16 (cg_solvers.f90:106): with a cap it is reported as capped or converged
    d = sb.cg()
    d.set_max_iterations(5000)
    d.setup(A)
    d.solve(A, np.zeros(n), b)
    assert d.info()[1] < 1e-20


def test_setup_rejects_non_square(sb):
    ptr = np.array([1, 2, 3], np.int32)
    node = np.array([1, 3], np.int32)
    A = sb.csr_matrix(2, 3, ptr, node, np.ones(2))
    for mk in (sb.cg, sb.bicgstab, sb.jacobi):
        with pytest.raises(sb.SigmaError) as e:
            mk().setup(A)
        assert e.value.status == 4 and "non-square" in e.value.message


def test_lanczos_identities_and_parity(sb, orc):
    """test/eigensolver_test_lanczos.f90:58-170: ER graph Laplacian (nn=128,
    nq=11), three-term recurrence and orthogonality to 1e-14; plus T against
    the oracle with the same start vector."""
    n, nq = 128, 11
    ptr, node, val = G.erdos_renyi_csr(n, seed=41, shift=0.0)
    A = sb.csr_matrix(n, n, ptr, node, val)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    q1 = 2 * np.random.default_rng(3).random(n) - 1
    T, V = sb.lanczos(A, nq, q1)
    for i in range(1, nq - 1):
        x = A.matvec(V[:, i])
        y = T[1, i] * V[:, i] + T[0, i - 1] * V[:, i - 1] + T[2, i] * V[:, i + 1]
        assert np.sqrt(((y - x) ** 2).sum() / (x**2).sum()) <= 1e-14
    Qm = V.T @ V - np.eye(nq)
    assert np.sqrt((Qm**2).sum()) / nq <= 1e-14
    To, Vo = orc.lanczos(O, nq, q1)
    assert np.array_equal(T[0], T[2])
    assert np.allclose(T, To, rtol=1e-9, atol=1e-11)
    assert np.allclose(V[:, :4], Vo[:, :4], rtol=0, atol=1e-12)
    # library-drawn start vector: deterministic in the seed, still orthonormal
    T1, V1 = sb.lanczos(A, nq, None, seed=5)
    T2, V2 = sb.lanczos(A, nq, None, seed=5)
    assert np.array_equal(T1, T2) and np.array_equal(V1, V2)
    assert np.sqrt(((V1.T @ V1 - np.eye(nq)) ** 2).sum()) / nq <= 1e-14


def test_eigensolve_ritz_values(sb, orc):
    n, nq = 900, 24
    ptr, node, val = G.poisson2d_csr(30)
    A = sb.csr_matrix(n, n, ptr, node, val)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    q1 = 2 * np.random.default_rng(4).random(n) - 1
    lam, V = sb.eigensolve(A, nq, q1)
    info, lamo, Vo = orc.eigensolve(O, nq, q1)
    assert info == 0
    assert np.allclose(lam, lamo, rtol=1e-9, atol=1e-10)
    # extremal Ritz pairs are converged eigenpairs: ||A v - lam v|| small, sign fixed by V(1,i) > 0
    lam_max = 4 - 4 * np.cos(np.pi * 30 / 31)
    assert abs(lam[-1] - lam_max) < 5e-2
    assert np.all(V[0, :] > 0)
    r = A.matvec(V[:, -1]) - lam[-1] * V[:, -1]
    assert np.linalg.norm(r) < 0.2
    assert np.allclose(np.abs(V[:, -1]), np.abs(Vo[:, -1]), atol=1e-8)


def test_full_size_cg_residual_consistency(sb):
    """BASELINE config 2 at full size: 25 capped CG iterations on 4096^2; the
    recurrence residual the device tracks must agree with b - A x recomputed by
    an independent matvec, and the residual must have dropped."""
    N = 4096
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    b, xs = G.poisson2d_rhs(N)
    A = sb.csr_matrix(n, n, ptr, node, val)
    del ptr, node, val
    s = sb.cg(1e-10 * np.linalg.norm(b))
    s.set_max_iterations(25)
    s.setup(A)
    x = s.solve(A, np.zeros(n), b)
    it, res2, capped = s.info()
    assert it == 25 and capped
    r = b - A.matvec(x)
    assert abs(np.sqrt(res2) - np.linalg.norm(r)) <= 1e-9 * np.linalg.norm(b)
    assert np.linalg.norm(r) < np.linalg.norm(b)


def test_generalized_lanczos_like_reference_test(sb, orc):
    """test/eigensolver_test_generalized_lanczos.f90:59-200 on the device: the
    reference's own stiffness / mass matrices, B%set_solver(cg(1.0d-15)), nq = 48;
    the identities the test checks (1e-14) and T against the oracle."""
    ptr, node, vA, vB = G.periodic_p1_grid()
    nn, nq = 48 * 32, 48
    A = sb.csr_matrix(nn, nn, ptr, node, vA)
    B = sb.csr_matrix(nn, nn, ptr, node, vB)
    bs = sb.cg(1e-15)
    bs.setup(B)                       # call B%set_solver(cg(1.0d-15))
    bs.set_max_iterations(5000)
    q1 = 2 * np.random.default_rng(0).random(nn) - 1
    T, V = sb.generalized_lanczos(A, B, bs, nq, q1)
    assert not bs.info()[2] and bs.iterations > 0
    U = np.stack([B.matvec(V[:, i]) for i in range(nq)], 1)
    for i in range(1, nq - 1):
        w = A.matvec(V[:, i])
        z = T[1, i] * U[:, i] + T[0, i - 1] * U[:, i - 1] + T[2, i] * U[:, i + 1]
        assert np.sqrt(((w - z) ** 2).sum() / (w * w).sum()) <= 1e-14
    Qm = V.T @ U - np.eye(nq)
    assert np.sqrt((Qm**2).sum()) / nq <= 1e-14
    OA = orc.Matrix(orc.CSR, nn, nn, node, vA, ptr=ptr)
    OB = orc.Matrix(orc.CSR, nn, nn, node, vB, ptr=ptr)
    To, Vo, _ = orc.generalized_lanczos(OA, OB, nq, q1, 1e-15, 5000)
    assert np.allclose(T[:, :8], To[:, :8], rtol=1e-9, atol=1e-11)
    assert np.allclose(V[:, :4], Vo[:, :4], atol=1e-11)
    # generalized_eigensolve: Ritz values of (A, B) from the same T
    lam, W = sb.generalized_eigensolve(A, B, bs, nq, q1)
    ref = np.linalg.eigvalsh(np.diag(T[1]) + np.diag(T[2, :-1], 1) + np.diag(T[2, :-1], -1))
    assert np.allclose(lam, ref, rtol=1e-8, atol=1e-8)
    # Ritz values of a positive semi-definite pencil: non-negative, smallest one near
    # the zero eigenvalue of the periodic stiffness matrix (48 steps, no re-orthogonalisation)
    assert lam[0] > -1e-8 and lam[0] < 0.05


# ---------------------------------------------------------------------------
# strict-order dot products: the whole solve bit for bit
# ---------------------------------------------------------------------------
def _strict(sb, kind, tol, A, cap):
    s = sb.cg(tol) if kind == "cg" else sb.bicgstab(tol)
    s.set_strict_order(True)
    s.set_max_iterations(cap)
    s.setup(A)
    return s


def test_strict_order_kat_programs_equal_the_serial_loops(sb, orc):
    """The reference's two deterministic test programs with the dot products summed left to right
    (sigb_solver_set_strict_order): 64 CG iterations and 1133 BiCGSTAB iterations -- the counts of the
    serial restatement, exactly -- and the same solution vector, bit for bit."""
    nn = 127
    dx = 1.0 / (nn + 1)
    enode, edeg, eval_ = G.tridiag_ell(nn)
    f = np.full(nn, 2.0 * dx**2)
    O = orc.Matrix(orc.ELL, nn, nn, enode, eval_, degrees=edeg)
    xo, ito, _, _ = orc.cg_solve(O, np.zeros(nn), f, 1e-16, 20 * nn)
    A = sb.ellpack_matrix(nn, nn, enode, edeg, eval_)
    s = _strict(sb, "cg", 1e-16, A, 20 * nn)
    x = s.solve(A, np.zeros(nn), f)
    assert s.info()[0] == ito == GOLD["diffusion_1d"]["iterations"] and not s.info()[2]
    assert np.array_equal(x, xo)

    nn, c = 1024, 0.5
    dx = 1.0 / (nn + 1)
    enode, edeg, eval_ = G.tridiag_ell(nn, 2.0, -1.0 + c * dx / 2, -1.0 - c * dx / 2)
    f = np.full(nn, 2.0 * dx**2)
    O = orc.Matrix(orc.ELL, nn, nn, enode, eval_, degrees=edeg)
    xo, ito, _, _ = orc.bicgstab_solve(O, np.zeros(nn), f, 1e-12, 20 * nn)
    A = sb.ellpack_matrix(nn, nn, enode, edeg, eval_)
    s = _strict(sb, "bicgstab", 1e-12, A, 20 * nn)
    x = s.solve(A, np.zeros(nn), f)
    assert s.info()[0] == ito == GOLD["advection_diffusion_1d"]["iterations"] and not s.info()[2]
    assert np.array_equal(x, xo)


@pytest.mark.parametrize("case", ["cg_poisson", "pcg_er", "bicgstab_er", "bicgstab_jacobi_er"])
def test_strict_order_solves_are_bit_identical(sb, orc, case):
    """CG, Jacobi-PCG, BiCGSTAB and Jacobi-BiCGSTAB on the parity-sized operators: with strict-order
    dot products the iteration count EQUALS the oracle's and the solution is the same array -- which
    shows that every other statement of the recurrences (SpMV rows, element-wise updates, scalar
    expressions, stopping test) is the reference's, rounding included.  The +-2 % / +-5 % bars of the
    tests above are therefore entirely the summation order of the default (parallel) dot products."""
    if case == "cg_poisson":
        N = 96
        n = N * N
        ptr, node, val = G.poisson2d_csr(N)
        b, _ = G.poisson2d_rhs(N)
        tol, kind, pc_on = 1e-10 * np.linalg.norm(b), "cg", False
    else:
        n = 3000
        skew = case.startswith("bicgstab")
        ptr, node, val = G.erdos_renyi_csr(n, seed=12 if skew else 8, weights="random", skew=skew)
        b = orc.matvec(orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr), np.random.default_rng(9).random(n))
        tol, kind, pc_on = 1e-13, ("bicgstab" if skew else "cg"), case in ("pcg_er", "bicgstab_jacobi_er")
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    idiag = orc.jacobi_setup(O) if pc_on else None
    fn = orc.cg_solve if kind == "cg" else orc.bicgstab_solve
    xo, ito, _, cappedo = fn(O, np.zeros(n), b, tol, 10 * n, idiag=idiag)
    A = sb.csr_matrix(n, n, ptr, node, val)
    s = _strict(sb, kind, tol, A, 10 * n)
    pc = None
    if pc_on:
        pc = sb.jacobi()
        pc.setup(A)
    x = s.solve(A, np.zeros(n), b, pc)
    it, res2, capped = s.info()
    assert not capped and not cappedo
    assert it == ito, (case, it, ito)
    assert np.array_equal(x, xo), (case, float(np.abs(x - xo).max()))
