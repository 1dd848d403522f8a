"""Operator expressions on the device against the oracle,
through the C-ABI: operator_sum, operator_product, operator_adjoint
(test/linear_operator_test_algebra.f90) and the block composite sparse_matrix
(test/matrix_test_composite.f90), then the solvers and Lanczos driven by them.

Every leaf runs the same SpMV kernels in the reference's accumulation order,
so matvec results are compared bit for bit (array_equal); the solvers keep the
north_star bars (+-2 % iterations, 1e-10 relative solution error)."""
import numpy as np
import pytest

from sigma_b200 import generators as G
from test_oracle_operators import exact_composite_product, random_pattern

pytestmark = pytest.mark.gpu


def within(it, ref, frac=0.02):
    return abs(it - ref) <= max(1, int(np.ceil(frac * ref)))


def both(sb, orc, fmt, nrow, ncol, ptr, node, val):
    """The same stored arrays as a product matrix and as an oracle matrix."""
    if fmt == "csr":
        return sb.csr_matrix(nrow, ncol, ptr, node, val), orc.Matrix(orc.CSR, nrow, ncol, node, val, ptr=ptr)
    if fmt == "csc":
        return sb.csc_matrix(nrow, ncol, ptr, node, val), orc.Matrix(orc.CSC, nrow, ncol, node, val, ptr=ptr)
    enode, edeg, eval_ = G.csr_to_ell(ptr, node, val)
    return sb.ellpack_matrix(nrow, ncol, enode, edeg, eval_), orc.Matrix(orc.ELL, nrow, ncol, enode, eval_, degrees=edeg)


def all_four(L, O, orc, x, xt, y0, y0t):
    """matvec, matvec_add, matvec_t, matvec_t_add: product vs oracle, bit for bit."""
    assert np.array_equal(L.matvec(x), orc.matvec(O, x))
    assert np.array_equal(L.matvec_add(x, y0), orc.matvec_add(O, x, y0))
    assert np.array_equal(L.matvec_t(xt), orc.matvec(O, xt, trans=True))
    assert np.array_equal(L.matvec_t_add(xt, y0t), orc.matvec_add(O, xt, y0t, trans=True))


@pytest.mark.parametrize("nn", [64, 3000])
def test_linear_operator_algebra(sb, orc, nn):
    """test/linear_operator_test_algebra.f90 (nn = 64 there): A csr, B csc on random graphs,
    entries 2q-1; L = A + B, A * B, adjoint(A), adjoint(A) * A."""
    rng = np.random.default_rng(nn)
    p = np.log2(nn) / nn
    ptr, node = random_pattern(nn, nn, p, rng)
    val = 2 * rng.random(node.size) - 1
    hptr, hnode = random_pattern(nn, nn, p, rng)
    hval = 2 * rng.random(hnode.size) - 1
    A, OA = both(sb, orc, "csr", nn, nn, ptr, node, val)
    B, OB = both(sb, orc, "csc", nn, nn, hptr, hnode, hval)
    ones = np.ones(nn)
    x, y0 = rng.standard_normal(nn), rng.standard_normal(nn)

    L, O = A + B, orc.operator_sum(OA, OB)
    y = L.matvec_add(ones, np.zeros(nn))
    z = B.matvec_add(ones, A.matvec_add(ones, np.zeros(nn)))
    assert np.abs(y).max() > 0 and np.abs(y - z).max() <= 1e-14       # the reference's check (:190-196)
    all_four(L, O, orc, x, x, y0, y0)

    L, O = A * B, orc.operator_product(OA, OB)
    z = A.matvec(B.matvec(ones))
    y = L.matvec(ones)
    assert np.abs(y).max() > 0 and np.abs(y - z).max() <= 1e-14       # (:232-239)
    all_four(L, O, orc, x, x, y0, y0)

    L, O = sb.adjoint(A), orc.adjoint(OA)
    assert np.abs(L.matvec(ones) - A.matvec_t(ones)).max() <= 1e-12   # (:255-261)
    all_four(L, O, orc, x, x, y0, y0)

    L, O = sb.adjoint(A) * A, orc.operator_product(orc.adjoint(OA), OA)
    assert np.abs(L.matvec(ones) - A.matvec_t(A.matvec(ones))).max() <= 1e-12   # (:277-283)
    all_four(L, O, orc, x, x, y0, y0)
    # a second application gives the same result (scratch vectors carry nothing over)
    assert np.array_equal(L.matvec(x), L.matvec(x))


def test_rectangular_and_nested_expressions(sb, orc):
    """Shapes the reference's test does not reach: rectangular factors (temp_vec_size =
    the largest dimension, linear_operator_products.f90:60), ellpack leaves, a three-factor
    chain built by nesting, sums of products."""
    rng = np.random.default_rng(9)
    n, m, k = 900, 1300, 500
    pa, na = random_pattern(n, m, 8.0 / m, rng)
    pb, nb = random_pattern(m, k, 8.0 / k, rng)
    pc, nc = random_pattern(k, k, 6.0 / k, rng)
    A, OA = both(sb, orc, "csr", n, m, pa, na, rng.standard_normal(na.size))
    # B as a csc_matrix: its column graph has k lines over m ids
    pbt, nbt, vbt = G.csr_transpose(m, k, pb, nb, rng.standard_normal(nb.size))
    B, OB = both(sb, orc, "csc", m, k, pbt, nbt, vbt)
    # C in ELLPACK needs a neighbour in every row
    cmask = np.zeros((k, k), bool)
    cmask[np.repeat(np.arange(k), np.diff(pc)), nc - 1] = True
    cmask[np.arange(k), np.arange(k)] = True
    cr, cc = np.nonzero(cmask)
    pc2 = np.concatenate([[1], 1 + np.cumsum(cmask.sum(1))]).astype(np.int32)
    C, OC = both(sb, orc, "ell", k, k, pc2, (cc + 1).astype(np.int32), rng.standard_normal(cc.size))

    AB, OAB = A * B, orc.operator_product(OA, OB)              # n x k
    assert (AB.nrow, AB.ncol) == (n, k)
    ABC, OABC = AB * C, orc.operator_product(OAB, OC)          # (A B) C, n x k
    x, xt = rng.standard_normal(k), rng.standard_normal(n)
    y0, y0t = rng.standard_normal(n), rng.standard_normal(k)
    all_four(AB, OAB, orc, x, xt, y0, y0t)
    all_four(ABC, OABC, orc, x, xt, y0, y0t)
    # (A B C) + (A B): a sum whose summands are products
    S, OS = ABC + AB, orc.operator_sum(OABC, OAB)
    all_four(S, OS, orc, x, xt, y0, y0t)
    # adjoint of the sum, k x n
    T, OT = sb.adjoint(S), orc.adjoint(OS)
    assert (T.nrow, T.ncol) == (k, n)
    all_four(T, OT, orc, xt, x, y0t, y0)
    # normal-equations operator S^T S, k x k, as one expression
    N, ON = T * S, orc.operator_product(OT, OS)
    all_four(N, ON, orc, x, x, y0t, y0t)


def composite_blocks(sb, orc, nn1, nn2, seed):
    c = G.composite_er_blocks(nn1, nn2, seed=seed)
    ptrh, nodeh = c["h"]
    b11, o11 = both(sb, orc, "csr", nn1, nn1, *c["b11"])
    b22, o22 = both(sb, orc, "csr", nn2, nn2, *c["b22"])
    b12, o12 = both(sb, orc, "csr", nn1, nn2, ptrh, nodeh, c["v12"])
    b21, o21 = both(sb, orc, "csc", nn2, nn1, ptrh, nodeh, c["v21"])   # the same graph h, as columns
    return c, [[b11, b12], [b21, b22]], [[o11, o12], [o21, o22]]


def test_composite_matvec(sb, orc):
    """test/matrix_test_composite.f90: 2 x 2 composite, nn1 = 768, nn2 = 512; matvec against
    the product written out from the graphs, RMS bar 1e-14 (:413-487) -- and bit for bit
    against the oracle's block loops."""
    nn1, nn2 = 768, 512
    c, blocks, oblocks = composite_blocks(sb, orc, nn1, nn2, 11)
    A = sb.sparse_matrix([nn1, nn2], [nn1, nn2], blocks)
    O = orc.composite([nn1, nn2], [nn1, nn2], oblocks)
    n = nn1 + nn2
    assert (A.nrow, A.ncol) == (n, n)
    assert A.nnz == sum(b.nnz for row in blocks for b in row)
    rng = np.random.default_rng(5)
    x = rng.random(n)
    y = A.matvec(x)
    z = exact_composite_product(c, x, nn1)
    assert np.sqrt(np.dot(y - z, y - z) / np.dot(x, x)) <= 1e-14
    y0 = rng.standard_normal(n)
    all_four(A, O, orc, x, x, y0, y0)


def test_composite_rectangular_blocks_and_nesting(sb, orc):
    """3 x 2 block grid with ragged block sizes, an ellpack block, a block that is itself an
    expression, and a composite used as a block of another composite."""
    rng = np.random.default_rng(21)
    rows, cols = [300, 1, 450], [520, 231]

    def blk(r, c, fmt):
        mask = rng.random((r, c)) < min(1.0, 5.0 / c)
        mask[np.arange(r), rng.integers(0, c, r)] = True       # no empty row (ellpack)
        rr, cc = np.nonzero(mask)
        ptr = np.concatenate([[1], 1 + np.cumsum(mask.sum(1))]).astype(np.int32)
        node, val = (cc + 1).astype(np.int32), rng.standard_normal(cc.size)
        if fmt == "csc":
            ptr, node, val = G.csr_transpose(r, c, ptr, node, val)
        return both(sb, orc, fmt, r, c, ptr, node, val)

    fmts = [["csr", "ell"], ["csc", "csr"], ["ell", "csc"]]
    pairs = [[blk(rows[i], cols[j], fmts[i][j]) for j in range(2)] for i in range(3)]
    blocks = [[p[0] for p in row] for row in pairs]
    oblocks = [[p[1] for p in row] for row in pairs]
    # block (1,1) := B11 + B11 (an expression as a sub-matrix)
    blocks[0][0], oblocks[0][0] = blocks[0][0] + pairs[0][0][0], orc.operator_sum(oblocks[0][0], pairs[0][0][1])
    A = sb.sparse_matrix(rows, cols, blocks)
    O = orc.composite(rows, cols, oblocks)
    nr, nc = sum(rows), sum(cols)
    assert (A.nrow, A.ncol) == (nr, nc)
    x, xt = rng.standard_normal(nc), rng.standard_normal(nr)
    y0, y0t = rng.standard_normal(nr), rng.standard_normal(nc)
    all_four(A, O, orc, x, xt, y0, y0t)
    # [[A, A], [A, A]] : composites nest (sparse_matrix_composites.f90:17-19)
    AA = sb.sparse_matrix([nr, nr], [nc, nc], [[A, A], [A, A]])
    OO = orc.composite([nr, nr], [nc, nc], [[O, O], [O, O]])
    x2, xt2 = rng.standard_normal(2 * nc), rng.standard_normal(2 * nr)
    all_four(AA, OO, orc, x2, xt2, rng.standard_normal(2 * nr), rng.standard_normal(2 * nc))


def test_composite_with_nearly_empty_blocks_and_signed_zeros(sb, orc):
    """2-D Poisson 256^2 cut into 2 x 2 blocks: the off-diagonal blocks hold 256 entries in
    32 768 rows, so almost all of their row tiles are empty and are skipped by the
    accumulating contributions.  Still bit for bit the reference's block loops -- including
    the sign of zero: csr_matvec_add does y(i) = y(i) + 0.0 for a row without entries, which
    turns a -0.0 the CALLER passed into +0.0 (only the first contribution can meet one)."""
    N = 256
    n, h = N * N, N * N // 2
    ptr, node, val = G.poisson2d_csr(N)
    rows = np.repeat(np.arange(n), np.diff(ptr))
    cols = node.astype(np.int64) - 1
    blocks, oblocks = [], []
    for r0, r1 in ((0, h), (h, n)):
        brow, orow = [], []
        for c0, c1 in ((0, h), (h, n)):
            m = (rows >= r0) & (rows < r1) & (cols >= c0) & (cols < c1)
            bptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(rows[m] - r0, minlength=r1 - r0))]).astype(np.int32)
            b, o = both(sb, orc, "csr", r1 - r0, c1 - c0, bptr, (cols[m] - c0 + 1).astype(np.int32), val[m])
            brow.append(b)
            orow.append(o)
        blocks.append(brow)
        oblocks.append(orow)
    A = sb.sparse_matrix([h, n - h], [h, n - h], blocks)
    O = orc.composite([h, n - h], [h, n - h], oblocks)
    rng = np.random.default_rng(8)
    x, y0 = rng.standard_normal(n), rng.standard_normal(n)
    all_four(A, O, orc, x, x, y0, y0)
    # close to the monolithic operator (rows cut by the block boundary add in another order)
    mono = sb.csr_matrix(n, n, ptr, node, val)
    assert np.abs(A.matvec(x) - mono.matvec(x)).max() <= 1e-13
    # signed zeros: x = 0 makes every row sum +0.0; y0 = -0.0 everywhere
    zero, negz = np.zeros(n), np.full(n, -0.0)
    for trans in (False, True):
        got = A.matvec_add(zero, negz, trans=trans)
        want = orc.matvec_add(O, zero, negz, trans=trans)
        assert np.array_equal(got, want) and np.array_equal(np.signbit(got), np.signbit(want))
    # ... and a lone off-diagonal block (rows without entries) as its own operator
    B12, O12 = blocks[0][1], oblocks[0][1]
    for trans in (False, True):
        nin, nout = (h, n - h) if trans else (n - h, h)
        got = B12.matvec_add(np.zeros(nin), np.full(nout, -0.0), trans=trans)
        want = orc.matvec_add(O12, np.zeros(nin), np.full(nout, -0.0), trans=trans)
        assert np.array_equal(np.signbit(got), np.signbit(want))


def test_expression_outlives_its_operands(sb, orc):
    """The expression holds references on its operands (add_reference,
    linear_operator_sums.f90:66-67): destroying the caller's handles first is safe."""
    n = 500
    ptr, node, val = G.erdos_renyi_csr(n, seed=3, weights="random")
    A, OA = both(sb, orc, "csr", n, n, ptr, node, val)
    B, OB = both(sb, orc, "ell", n, n, ptr, node, 0.5 * val)
    L = A * B + A
    A.destroy()
    B.destroy()
    x = np.random.default_rng(0).standard_normal(n)
    O = orc.operator_sum(orc.operator_product(OA, OB), OA)
    assert np.array_equal(L.matvec(x), orc.matvec(O, x))
    L.destroy()


def test_expression_errors(sb):
    a = sb.csr_matrix(2, 3, [1, 2, 3], [1, 2], [1.0, 1.0])
    b = sb.csr_matrix(2, 2, [1, 2, 3], [1, 2], [1.0, 1.0])
    with pytest.raises(sb.SigmaError) as e:
        a + b
    assert e.value.status == 1 and "summed are not consistent" in e.value.message
    with pytest.raises(sb.SigmaError) as e:
        a * b
    assert e.value.status == 1 and "multiplied are inconsistent" in e.value.message
    with pytest.raises(sb.SigmaError) as e:
        sb.sparse_matrix([2, 2], [2, 3], [[b, a], [a, b]])
    assert e.value.status == 1 and "Inconsistent dimensions for sub-matrix" in e.value.message
    L = b * a
    assert (L.nrow, L.ncol) == (2, 3)
    assert np.array_equal(L.matvec(np.array([1.0, 2.0, 3.0])), np.array([1.0, 2.0]))
    with pytest.raises(sb.SigmaError) as e:
        L.set_values(np.zeros(2))
    assert e.value.status == 7
    sq = b * b
    pc = sb.jacobi()
    with pytest.raises(sb.SigmaError) as e:   # get_value of a product is undefined in the reference
        pc.setup(sq)
    assert e.value.status == 7
    with pytest.raises(sb.SigmaError) as e:   # "Cannot make a CG solver for a non-square matrix"
        sb.cg().setup(L)
    assert e.value.status == 4


def shifted_composite(sb, orc, nn1, nn2, seed, shift):
    """A + shift * I with A the two-field composite: symmetric positive definite."""
    c, blocks, oblocks = composite_blocks(sb, orc, nn1, nn2, seed)
    n = nn1 + nn2
    A = sb.sparse_matrix([nn1, nn2], [nn1, nn2], blocks)
    O = orc.composite([nn1, nn2], [nn1, nn2], oblocks)
    ip, inode, ival = np.arange(1, n + 2, dtype=np.int32), np.arange(1, n + 1, dtype=np.int32), np.full(n, shift)
    I, OI = both(sb, orc, "csr", n, n, ip, inode, ival)
    return n, A + I, orc.operator_sum(O, OI)


@pytest.mark.parametrize("precond", [False, True])
def test_cg_on_composite_expression(sb, orc, precond):
    """cg_solve / cg_solve_pc take any linear_operator (cg_solvers.f90:116-121,155-161):
    the two-field composite plus a diagonal shift, with and without jacobi."""
    n, S, OS = shifted_composite(sb, orc, 768, 512, 13, 0.25)
    xs = np.random.default_rng(1).random(n)
    b = orc.matvec(OS, xs)
    assert np.array_equal(S.matvec(xs), b)
    tol = 1e-13 * np.linalg.norm(b)
    pc = idiag = None
    if precond:
        pc = sb.jacobi()
        pc.setup(S)
        idiag = orc.jacobi_setup(OS)
        assert np.array_equal(pc.vector("idiag"), idiag)      # get_value through sum and composite
    solver = sb.cg(tol)
    solver.set_max_iterations(10 * n)
    solver.setup(S)
    x = solver.solve(S, np.zeros(n), b, pc)
    it, res2, capped = solver.info()
    xo, ito, res2o, cappedo = orc.cg_solve(OS, np.zeros(n), b, tol, 10 * n, idiag=idiag)
    assert not capped and not cappedo
    assert within(it, ito)
    assert np.abs(x - xo).max() <= 1e-10 * np.abs(xo).max()
    assert np.sqrt(res2) <= tol


def test_bicgstab_on_sum_of_csr_and_csc(sb, orc):
    """bicgstab_solve on L = A + B with A a shifted nonsymmetric Laplacian (csr) and B a
    small csc perturbation; same stopping rule, +-5 % iterations (random nonsymmetric
    operator), 1e-10 relative solution error... against the true solution."""
    n = 4000
    ptr, node, val = G.erdos_renyi_csr(n, seed=5, weights="random", skew=True, shift=2.0)
    A, OA = both(sb, orc, "csr", n, n, ptr, node, val)
    rng = np.random.default_rng(3)
    pb, nb = random_pattern(n, n, 3.0 / n, rng)
    B, OB = both(sb, orc, "csc", n, n, pb, nb, 0.05 * rng.standard_normal(nb.size))
    L, OL = A + B, orc.operator_sum(OA, OB)
    xs = rng.random(n)
    b = orc.matvec(OL, xs)
    tol = 1e-11 * np.linalg.norm(b)
    solver = sb.bicgstab(tol)
    solver.set_max_iterations(5000)
    solver.setup(L)
    x = solver.solve(L, np.zeros(n), b)
    it, res2, capped = solver.info()
    xo, ito, _, cappedo = orc.bicgstab_solve(OL, np.zeros(n), b, tol, 5000)
    assert not capped and not cappedo
    assert within(it, ito, 0.05)
    assert np.abs(x - xs).max() <= 1e-8 * np.abs(xs).max()
    assert np.abs(x - xo).max() <= 1e-8 * np.abs(xo).max()


def test_lanczos_on_composite(sb, orc):
    """lanczos(A, T, Q) takes class(linear_operator) (eigensolver.f90:27-31): same start
    vector, tridiagonal entries against the oracle, orthonormal Q."""
    n, S, OS = shifted_composite(sb, orc, 300, 200, 17, 1.0)
    q1 = np.random.default_rng(2).uniform(-1, 1, n)
    nsteps = 12
    T, Q = sb.lanczos(S, nsteps, q1)
    To, Qo = orc.lanczos(OS, nsteps, q1)
    assert np.array_equal(T[0], T[2])
    assert np.allclose(T, To, rtol=1e-9, atol=1e-11)
    assert np.allclose(Q[:, :4], Qo[:, :4], rtol=0, atol=1e-12)
    assert np.sqrt(((Q.T @ Q - np.eye(nsteps)) ** 2).sum()) / nsteps <= 1e-13
    # three-term recurrence, as test/eigensolver_test_lanczos.f90:130-150 checks it
    for i in range(1, nsteps - 1):
        w = orc.matvec(OS, Q[:, i])
        y = T[1, i] * Q[:, i] + T[0, i - 1] * Q[:, i - 1] + T[2, i] * Q[:, i + 1]
        assert np.sqrt(((y - w) ** 2).sum() / (w**2).sum()) <= 1e-13
