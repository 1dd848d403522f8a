"""The oracle's restatement of the operator expressions (operator_sum,
operator_product, operator_adjoint, composite sparse_matrix) held to the
reference's own test bars, on CPU:

  test/linear_operator_test_algebra.f90   sum get_value 1e-14 (:157-165), sum matvec 1e-14
                                          (:190-196), product 1e-14 (:232-239), adjoint 1e-12
                                          (:255-261), adjoint*A 1e-12 (:277-283)
  test/matrix_test_composite.f90          matvec RMS error 1e-14 against the product written
                                          out from the graphs (:413-487), get_value (:229-285)

The reference draws its inputs from a time-seeded RNG (init_seed), so these are
property tests on seeded inputs of the same shape, like the reference's."""
import numpy as np
import pytest

from sigma_b200 import generators as G


def random_pattern(n, m, p, rng):
    """g%add_edge(i, j) for q < p over the full index square (:83-93): rows ascending."""
    mask = rng.random((n, m)) < p
    r, c = np.nonzero(mask)
    ptr = np.concatenate([[1], 1 + np.cumsum(mask.sum(axis=1))]).astype(np.int32)
    return ptr, (c + 1).astype(np.int32)


@pytest.fixture(scope="module")
def algebra_case(orc):
    nn = 64
    rng = np.random.default_rng(2024)
    p = np.log2(nn) / nn
    ptr, node = random_pattern(nn, nn, p, rng)
    val = 2 * rng.random(node.size) - 1
    # h%add_edge(j, i) then B%init(nn, nn, h) on a csc_matrix: the column graph
    hptr, hnode = random_pattern(nn, nn, p, rng)
    hval = 2 * rng.random(hnode.size) - 1
    A = orc.Matrix(orc.CSR, nn, nn, node, val, ptr=ptr)
    B = orc.Matrix(orc.CSC, nn, nn, hnode, hval, ptr=hptr)
    return nn, A, B


def test_operator_sum(orc, algebra_case):
    nn, A, B = algebra_case
    L = orc.operator_sum(A, B)
    for i in range(1, nn + 1):
        for j in range(1, nn + 1):
            assert abs(orc.get_value(L, i, j) - orc.get_value(A, i, j) - orc.get_value(B, i, j)) <= 1e-14
    x = np.ones(nn)
    z = orc.matvec_add(B, x, orc.matvec_add(A, x, np.zeros(nn)))
    assert np.abs(z).max() > 0
    y = orc.matvec_add(L, x, np.zeros(nn))
    assert np.abs(y).max() > 0 and np.abs(y - z).max() <= 1e-14
    # the summands accumulate into the same y in order: identical, not just close
    assert np.array_equal(y, z)
    zt = orc.matvec_add(B, x, orc.matvec_add(A, x, np.zeros(nn), trans=True), trans=True)
    assert np.array_equal(orc.matvec(L, x, trans=True), zt)


def test_operator_product_and_adjoint(orc, algebra_case):
    nn, A, B = algebra_case
    x = np.ones(nn)
    L = orc.operator_product(A, B)
    z = orc.matvec(A, orc.matvec(B, x))
    y = orc.matvec(L, x)
    assert np.abs(y).max() > 0 and np.abs(y - z).max() <= 1e-14
    # (A B)^T x = B^T (A^T x)
    assert np.array_equal(orc.matvec(L, x, trans=True), orc.matvec(B, orc.matvec(A, x, trans=True), trans=True))
    L = orc.adjoint(A)
    assert np.abs(orc.matvec(L, x) - orc.matvec(A, x, trans=True)).max() <= 1e-12
    assert orc.get_value(L, 3, 7) == orc.get_value(A, 7, 3)
    L = orc.operator_product(orc.adjoint(A), A)
    z = orc.matvec(A, orc.matvec(A, x), trans=True)
    assert np.abs(orc.matvec(L, x) - z).max() <= 1e-12
    # the scratch vectors are left zeroed (:111-112) and a second product gives the same
    assert np.array_equal(orc.matvec(L, x), orc.matvec(L, x))
    y0 = np.arange(nn, dtype=float)
    assert np.array_equal(orc.matvec_add(L, x, y0), y0 + z)


def test_dimension_checks(orc):
    a = orc.Matrix(orc.CSR, 2, 3, [1, 2], [1.0, 1.0], ptr=[1, 2, 3])
    b = orc.Matrix(orc.CSR, 2, 2, [1, 2], [1.0, 1.0], ptr=[1, 2, 3])
    with pytest.raises(ValueError):
        orc.operator_sum(a, b)      # "Dimensions of operators to be summed are not consistent"
    with pytest.raises(ValueError):
        orc.operator_product(a, b)  # "Dimensions of operators to be multiplied are inconsistent"
    L = orc.operator_product(b, a)  # 2x2 * 2x3
    assert (L.nrow, L.ncol) == (2, 3)
    assert np.array_equal(orc.matvec(L, np.array([1.0, 2.0, 3.0])), np.array([1.0, 2.0]))


@pytest.fixture(scope="module")
def composite_case(orc):
    nn1, nn2 = 768, 512
    c = G.composite_er_blocks(nn1, nn2, seed=11)
    ptrh, nodeh = c["h"]
    blocks = [[orc.Matrix(orc.CSR, nn1, nn1, c["b11"][1], c["b11"][2], ptr=c["b11"][0]),
               orc.Matrix(orc.CSR, nn1, nn2, nodeh, c["v12"], ptr=ptrh)],
              [orc.Matrix(orc.CSC, nn2, nn1, nodeh, c["v21"], ptr=ptrh),   # the same graph h, as columns
               orc.Matrix(orc.CSR, nn2, nn2, c["b22"][1], c["b22"][2], ptr=c["b22"][0])]]
    return nn1, nn2, c, blocks, orc.composite([nn1, nn2], [nn1, nn2], blocks)


def exact_composite_product(c, x, nn1):
    """z of test/matrix_test_composite.f90:417-472, written out from the graphs."""
    a1 = c["adj1"].astype(float)
    a2 = c["adj2"].astype(float)
    ah = c["adjh"].astype(float)
    x1, x2 = x[:nn1], x[nn1:]
    z1 = a1.sum(1) * x1 - a1 @ x1 + ah.sum(1) * x1 - ah @ x2
    z2 = a2.sum(1) * x2 - a2 @ x2 + ah.sum(0) * x2 - ah.T @ x1
    return np.concatenate([z1, z2])


def test_composite_matvec_and_entries(orc, composite_case):
    nn1, nn2, c, blocks, A = composite_case
    assert (A.nrow, A.ncol) == (nn1 + nn2, nn1 + nn2)
    rng = np.random.default_rng(5)
    x = rng.random(nn1 + nn2)
    y = orc.matvec(A, x)
    z = exact_composite_product(c, x, nn1)
    mse = np.sqrt(np.dot(y - z, y - z) / np.dot(x, x))
    assert mse <= 1e-14
    # symmetric by construction: the transposed loop gives the same operator
    yt = orc.matvec(A, x, trans=True)
    assert np.sqrt(np.dot(yt - z, yt - z) / np.dot(x, x)) <= 1e-14
    # entries through the owning blocks (:229-285), on a sample
    adj1, adj2, adjh = c["adj1"], c["adj2"], c["adjh"]
    for i in rng.integers(1, nn1 + 1, 40):
        for j in rng.integers(1, nn1 + 1, 10):
            want = (adj1[i - 1].sum() + adjh[i - 1].sum() - 1.0) if i == j else (-1.0 if adj1[i - 1, j - 1] else 0.0)
            assert orc.get_value(A, i, j) == want
        for j in rng.integers(1, nn2 + 1, 10):
            want = -1.0 if adjh[i - 1, j - 1] else 0.0
            assert orc.get_value(A, i, nn1 + j) == want
            assert orc.get_value(A, nn1 + j, i) == want
    for i in range(1, nn1 + nn2 + 1, 37):
        assert orc.get_value(A, i, i) == orc.get_value(blocks[0][0] if i <= nn1 else blocks[1][1],
                                                       i if i <= nn1 else i - nn1, i if i <= nn1 else i - nn1)


def test_solvers_accept_expressions(orc, composite_case):
    """The reference solvers take any linear_operator (cg_solvers.f90:116-121): CG and
    Jacobi-PCG on the composite (symmetric, diagonally dominant + 0 row sums -> shifted)."""
    nn1, nn2, c, blocks, A = composite_case
    n = nn1 + nn2
    ident = orc.Matrix(orc.CSR, n, n, np.arange(1, n + 1), np.ones(n), ptr=np.arange(1, n + 2))
    S = orc.operator_sum(A, ident)
    xs = np.random.default_rng(6).random(n)
    b = orc.matvec(S, xs)
    x, it, res2, capped = orc.cg_solve(S, np.zeros(n), b, 1e-10, 5000)
    assert not capped and np.abs(x - xs).max() <= 1e-9
    idiag = orc.jacobi_setup(S)
    assert np.array_equal(idiag, 1.0 / (np.array([orc.get_value(A, i, i) for i in range(1, n + 1)]) + 1.0))
    xp, itp, _, capped = orc.cg_solve(S, np.zeros(n), b, 1e-10, 5000, idiag=idiag)
    assert not capped and itp <= it and np.abs(xp - xs).max() <= 1e-9
