"""ILDU(0) on the device against the oracle's restatement of
src/solver/ldu_solvers.f90, through the C-ABI.  Level scheduling only changes WHICH rows
run concurrently; every row does the reference's arithmetic in the reference's order, so
factors, diagonal and solves are compared bit for bit.  The solver-level checks restate
test/solver_test_incomplete_cholesky.f90 with its own bars (1e-14 stationary, 1e-15 PCG)."""
import numpy as np
import pytest

from sigma_b200 import generators as G
from test_ldu_symbolic import cases

pytestmark = pytest.mark.gpu


def within(it, ref, frac=0.02):
    return abs(it - ref) <= max(1, int(np.ceil(frac * ref)))


def in_format(sb, orc, fmt, n, ptr, node, val):
    if fmt == "csr":
        return sb.csr_matrix(n, n, ptr, node, val), orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    if fmt == "csc":
        cptr, cnode, cval = G.csr_transpose(n, n, ptr, node, val)
        return sb.csc_matrix(n, n, cptr, cnode, cval), orc.Matrix(orc.CSC, n, n, cnode, cval, ptr=cptr)
    enode, edeg, eval_ = G.csr_to_ell(ptr, node, val)
    return sb.ellpack_matrix(n, n, enode, edeg, eval_), orc.Matrix(orc.ELL, n, n, enode, eval_, degrees=edeg)


def same_factors(pc, F):
    Lptr, Lnode, Lval, Uptr, Unode, Uval, D, nf, nb = pc.factors()
    assert np.array_equal(Lptr, F.Lptr) and np.array_equal(Lnode, F.Lnode)
    assert np.array_equal(Uptr, F.Uptr) and np.array_equal(Unode, F.Unode)
    assert np.array_equal(Lval, F.Lval) and np.array_equal(Uval, F.Uval) and np.array_equal(D, F.D)
    return nf, nb


@pytest.mark.parametrize("fmt", ["csr", "csc", "ellpack"])
@pytest.mark.parametrize("case", list(cases()), ids=lambda c: c[0])
def test_factors_and_solve_bit_exact(sb, orc, case, fmt):
    _, n, ptr, node, val = case
    A, O = in_format(sb, orc, fmt, n, ptr, node, val)
    F = orc.ldu_setup(O)
    pc = sb.ldu()
    pc.setup(A)
    same_factors(pc, F)
    b = np.random.default_rng(0).standard_normal(n)
    assert np.array_equal(pc.solve(A, np.zeros(n), b), orc.ldu_solve(F, b))
    # a second setup on new values keeps the pattern and redoes the numbers (:113-126)
    A.set_values(2.0 * np.asarray(O.val).reshape(-1))
    pc.setup(A)
    O2 = type(O)(O.format, n, n, O.node, 2.0 * O.val, ptr=O.ptr, degrees=O.degrees)
    F2 = orc.ldu_setup(O2)
    same_factors(pc, F2)
    assert np.array_equal(pc.solve(A, np.zeros(n), b), orc.ldu_solve(F2, b))


def deep_cases():
    """Schedules deep enough (> 64 levels, >= 32 chunks) for the chunked sweeps of csrc/ldu.cu."""
    from test_ldu_symbolic import shuffled

    p, n_, v = G.poisson2d_csr(64)
    yield "poisson64", 64 * 64, p, n_, v
    yield "poisson48_shuffled", 48 * 48, G.poisson2d_csr(48)[0], *shuffled(*G.poisson2d_csr(48), 7)
    p, n_, v = G.fem_p1_csr(70)
    yield "fem70_shuffled", 70 * 70, p, *shuffled(p, n_, v, 3)
    yield "tridiag5000", 5000, *G.tridiag_csr(5000, 2.5, -1.0, -0.5)


@pytest.mark.parametrize("static", [1, 0], ids=["static_schedule", "chunked_polling"])
@pytest.mark.parametrize("case", list(deep_cases()), ids=lambda c: c[0])
def test_deep_sweeps_bit_exact(sb, orc, case, static, monkeypatch):
    """Deep level schedules run the triangular solves as ONE launch per sweep: on a schedule fixed at setup
    where the pattern has a wavefront (one CTA, a thread per chunk of rows, lock-step trips, values handed on
    through a shared-memory ring; csrc/ldu_sweep.h -- 6 launches per pc%solve with the transposes into and
    out of trip order), else as chunked sweeps that wait for exactly the entries they read (2 launches;
    forced here with SIGB_LDU_STATIC=0).  Same arithmetic per row in stored order either way, so pc%solve
    must equal the serial loops bit for bit -- repeatedly, and for csr and csc sources."""
    name, n, ptr, node, val = case
    monkeypatch.setenv("SIGB_LDU_STATIC", str(static))
    for fmt in ("csr", "csc"):
        A, O = in_format(sb, orc, fmt, n, ptr, node, val)
        F = orc.ldu_setup(O)
        pc = sb.ldu()
        pc.setup(A)
        nf, nb = same_factors(pc, F)
        assert nf > 64 and nb > 64
        rng = np.random.default_rng(1)
        x = np.zeros(n)
        for _ in range(3):
            b = rng.standard_normal(n)
            before = sb.launch_count()
            x = pc.solve(A, x, b)
            launches = sb.launch_count() - before
            assert np.array_equal(x, orc.ldu_solve(F, b))
        assert launches == (6 if static else 2), (name, launches)
        # a second numeric factorisation on the same pattern (the values in trip order are refreshed)
        A2, O2 = in_format(sb, orc, fmt, n, ptr, node, val * 1.25 + 0.0)
        pc.setup(A2)
        b = rng.standard_normal(n)
        assert np.array_equal(pc.solve(A2, np.zeros(n), b), orc.ldu_solve(orc.ldu_setup(O2), b))


def test_incomplete_cholesky_like_the_reference(sb, orc):
    """test/solver_test_incomplete_cholesky.f90: nn = 128 random weighted graph Laplacian + I;
    the factorisation as a stationary solver (10 nn sweeps, 1e-14, :182-202) and as the
    preconditioner of cg(1e-16) (1e-15, :213-226)."""
    nn = 128
    ptr, node, val = G.erdos_renyi_csr(nn, seed=1, weights="random", shift=1.0)
    A, O = in_format(sb, orc, "csr", nn, ptr, node, val)
    solver, pc = sb.cg(1e-16), sb.ldu(incomplete=True, level=0)
    solver.setup(A)
    pc.setup(A)
    F = orc.ldu_setup(O)
    rng = np.random.default_rng(1)
    v = rng.random(nn)
    r = v - A.matvec(v)
    v = pc.solve(A, v, r)
    f = A.matvec(v)
    u, r = np.zeros(nn), f.copy()
    q = np.zeros(nn)
    for _ in range(10 * nn):
        q = pc.solve(A, q, r)
        u = u + q
        r = f - A.matvec(u)
    assert np.abs(u - v).max() <= 1e-14
    solver.set_max_iterations(50 * nn)          # safety net only
    u = solver.solve(A, np.zeros(nn), f, pc)
    it, res2, capped = solver.info()
    assert not capped and np.abs(u - v).max() <= 1e-15
    uo, ito, _, _ = orc.cg_solve_ldu(O, np.zeros(nn), f, F, 1e-16, 50 * nn)
    assert within(it, ito)
    solver0 = sb.cg(1e-16)
    solver0.set_max_iterations(50 * nn)
    solver0.setup(A)
    solver0.solve(A, np.zeros(nn), f)
    assert it < solver0.iterations              # it does precondition


@pytest.mark.parametrize("fmt", ["csr", "ellpack"])
def test_pcg_ldu_poisson(sb, orc, fmt):
    """2-D Poisson 64^2 (127 levels per sweep): ILDU(0)-preconditioned CG against the oracle,
    north_star bars: +-2 % iterations, 1e-10 relative solution error."""
    N = 64
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    A, O = in_format(sb, orc, fmt, n, ptr, node, val)
    b, xs = G.poisson2d_rhs(N)
    tol = 1e-12 * np.linalg.norm(b)
    pc = sb.ldu()
    pc.setup(A)
    nf, nb = same_factors(pc, orc.ldu_setup(O))
    assert nf == 2 * N - 1 and nb == 2 * N - 1
    solver = sb.cg(tol)
    solver.set_max_iterations(10 * n)
    solver.setup(A)
    x = solver.solve(A, np.zeros(n), b, pc)
    it, res2, capped = solver.info()
    xo, ito, _, cappedo = orc.cg_solve_ldu(O, np.zeros(n), b, orc.ldu_setup(O), tol, 10 * n)
    assert not capped and not cappedo and within(it, ito)
    assert np.abs(x - xo).max() <= 1e-10 * np.abs(xo).max()
    xj, itj, _, _ = orc.cg_solve(O, np.zeros(n), b, tol, 10 * n)
    assert it < itj


def test_ldu_errors(sb):
    a = sb.csr_matrix(2, 3, [1, 2, 3], [1, 2], [1.0, 1.0])
    with pytest.raises(sb.SigmaError) as e:
        sb.ldu().setup(a)
    assert e.value.status == 4 and "an LDU solver for a non-square matrix" in e.value.message
    sq = sb.csr_matrix(2, 2, [1, 3, 5], [1, 2, 1, 2], [4.0, 1.0, 1.0, 3.0])
    with pytest.raises(sb.SigmaError) as e:      # sparse_ldu_setup selects on sparse matrices (:111-112)
        sb.ldu().setup(sq + sq)
    assert e.value.status == 7
    pc = sb.ldu()
    pc.setup(sq)
    # 2 x 2 by hand: D = [4, 3 - 1/4], L21 = 1/4, U12 = 1/4
    Lptr, Lnode, Lval, Uptr, Unode, Uval, D, nf, nb = pc.factors()
    assert np.array_equal(D, [4.0, 2.75]) and np.array_equal(Lval, [0.25]) and np.array_equal(Uval, [0.25])
    assert (nf, nb) == (2, 2)
    x = pc.solve(sq, np.zeros(2), np.array([5.0, 4.0]))
    assert np.allclose(x, [1.0, 1.0], rtol=0, atol=1e-15)


def test_bicgstab_with_ldu_preconditioner(sb, orc):
    """bicgstab_solve_pc with pc = ldu() (bicgstab_solvers.f90:182-237; linear_solve_pc takes any
    linear_solver): iterations within 5 % of the oracle's restatement (the count of BiCGSTAB on a
    random nonsymmetric operator moves with summation order alone), solution 1e-10."""
    for nn, seed in ((300, 4), (2000, 5)):
        ptr, node, val = G.erdos_renyi_csr(nn, seed=seed, weights="random", skew=True, shift=1.0)
        A = sb.csr_matrix(nn, nn, ptr, node, val)
        O = orc.Matrix(orc.CSR, nn, nn, node, val, ptr=ptr)
        F = orc.ldu_setup(O)
        v = np.random.default_rng(seed).random(nn)
        f = orc.matvec(O, v)
        tol = 1e-13
        s, pc = sb.bicgstab(tol), sb.ldu()
        s.set_max_iterations(50 * nn); s.setup(A); pc.setup(A)
        x = s.solve(A, np.zeros(nn), f, pc)
        it, res2, capped = s.info()
        xo, ito, _, cappedo = orc.bicgstab_solve_ldu(O, np.zeros(nn), f, F, tol, 50 * nn)
        assert not capped and not cappedo and within(it, ito, 0.05), (it, ito)
        assert np.abs(x - xo).max() <= 1e-10 * np.abs(xo).max()
        assert np.abs(x - v).max() <= 1e-10
        # a capped solve stops where it is told to
        s2 = sb.bicgstab(tol); s2.set_max_iterations(3); s2.setup(A)
        s2.solve(A, np.zeros(nn), f, pc)
        assert s2.info()[0] == 3 and s2.info()[2]
