"""The oracle's restatement of A%copy_matrix(B, trans) held to the reference's own
test, test/matrix_test_copy.f90:83-146, on CPU: a random nn x nn/2 matrix in every
format copied to every format, straight and transposed; every entry must agree
(the reference's bar is 1e-14; copies move values, so they agree exactly)."""
import numpy as np
import pytest


def random_rect_csr(nn, rng, shuffle=True):
    """g%init(nn, nn/2); add_edge(i, j) for q < p (:64-71); A%set(i, j, z) (:93-103).
    Columns are shuffled inside each row: nothing in the reference keeps them sorted."""
    m = nn // 2
    mask = rng.random((nn, m)) < np.log2(nn) / nn * 2
    mask[np.arange(nn), rng.integers(0, m, nn)] = True        # no empty row (ellpack targets)
    mask[rng.integers(0, nn, m), np.arange(m)] = True         # no empty column (transposed ellpack)
    r, c = np.nonzero(mask)
    ptr = np.concatenate([[1], 1 + np.cumsum(mask.sum(1))]).astype(np.int32)
    node = (c + 1).astype(np.int32)
    if shuffle:
        for i in range(nn):
            rng.shuffle(node[ptr[i] - 1: ptr[i + 1] - 1])
    return nn, m, ptr, node, rng.random(node.size)


FMTS = ["csr", "csc", "ellpack"]


def fmt_code(orc, f):
    return {"csr": orc.CSR, "csc": orc.CSC, "ellpack": orc.ELL}[f]


def dense(orc, A):
    D = np.zeros((A.nrow, A.ncol))
    i, j, v = orc.matrix_entries(A)
    D[i - 1, j - 1] = v
    return D


@pytest.mark.parametrize("frmt1", FMTS)
@pytest.mark.parametrize("frmt2", FMTS)
def test_copy_matrix_all_formats(orc, frmt1, frmt2):
    rng = np.random.default_rng(17)
    nn, m, ptr, node, val = random_rect_csr(64, rng)
    base = orc.Matrix(orc.CSR, nn, m, node, val, ptr=ptr)
    A = orc.copy_matrix(base, fmt_code(orc, frmt1))            # "A" of the reference test, in format 1
    DA = dense(orc, A)
    assert np.array_equal(DA, dense(orc, base))
    B = orc.copy_matrix(A, fmt_code(orc, frmt2), trans=False)
    assert (B.nrow, B.ncol) == (nn, m)
    assert all(orc.get_value(B, i, j) == DA[i - 1, j - 1] for i in range(1, nn + 1) for j in range(1, m + 1))
    Cm = orc.copy_matrix(A, fmt_code(orc, frmt2), trans=True)
    assert (Cm.nrow, Cm.ncol) == (m, nn)
    assert all(orc.get_value(Cm, j, i) == DA[i - 1, j - 1] for i in range(1, nn + 1) for j in range(1, m + 1))
    # the copies are operators like any other
    x = rng.standard_normal(m)
    assert np.allclose(orc.matvec(B, x), DA @ x, rtol=0, atol=1e-13)
    assert np.allclose(orc.matvec(Cm, x, trans=True), DA @ x, rtol=0, atol=1e-13)


def test_copy_keeps_iteration_order(orc):
    """What pins the index arrays: each target line holds its entries in the order the
    source iterator produced them (first-free-slot insertion, cs_graphs.f90:163-183)."""
    ptr = np.array([1, 4, 6, 7], np.int32)              # 3 x 4, unsorted rows
    node = np.array([3, 1, 4, 4, 2, 1], np.int32)
    val = np.arange(1.0, 7.0)
    A = orc.Matrix(orc.CSR, 3, 4, node, val, ptr=ptr)
    B = orc.copy_matrix(A, orc.CSR)
    assert np.array_equal(B.ptr, ptr) and np.array_equal(B.node, node) and np.array_equal(B.val, val)
    Cc = orc.copy_matrix(A, orc.CSC)                    # columns: 1 <- rows 1,3 ; 2 <- 2 ; 3 <- 1 ; 4 <- 1,2
    assert np.array_equal(Cc.ptr, [1, 3, 4, 5, 7])
    assert np.array_equal(Cc.node, [1, 3, 2, 1, 1, 2])
    assert np.array_equal(Cc.val, [2.0, 6.0, 5.0, 1.0, 3.0, 4.0])
    T = orc.copy_matrix(A, orc.CSR, trans=True)         # rows of A^T = columns of A: same arrays as the csc copy
    assert np.array_equal(T.ptr, Cc.ptr) and np.array_equal(T.node, Cc.node) and np.array_equal(T.val, Cc.val)
    E = orc.copy_matrix(A, orc.ELL)
    assert np.array_equal(E.degrees, [3, 2, 1])
    assert np.array_equal(E.node, [[3, 1, 4], [4, 2, 2], [1, 1, 1]])      # padding = last neighbour
    assert np.array_equal(E.val, [[1.0, 2.0, 3.0], [4.0, 5.0, 0.0], [6.0, 0.0, 0.0]])
    back = orc.copy_matrix(E, orc.CSR)
    assert np.array_equal(back.ptr, ptr) and np.array_equal(back.node, node) and np.array_equal(back.val, val)
