"""On-device matrix copy / format conversion against the oracle's
restatement of A%copy_matrix(B, trans), through the C-ABI.  Index work: every array
the device builds (ptr / node / degrees, padding included) must equal the reference
builders' output BIT FOR BIT, and the values ride along exactly."""
import numpy as np
import pytest

from sigma_b200 import generators as G
from test_oracle_copy import FMTS, dense, fmt_code, random_rect_csr

pytestmark = pytest.mark.gpu


def product_matrix(sb, A, orc):
    """The oracle matrix A as a product matrix on the same stored arrays."""
    if A.format == orc.CSR:
        return sb.csr_matrix(A.nrow, A.ncol, A.ptr, A.node, A.val)
    if A.format == orc.CSC:
        return sb.csc_matrix(A.nrow, A.ncol, A.ptr, A.node, A.val)
    return sb.ellpack_matrix(A.nrow, A.ncol, A.node, A.degrees, A.val)


def same_arrays(orc, dev, O):
    """Device copy `dev` (sb.Matrix) against oracle copy O: format, shapes, arrays."""
    got = dev.arrays()
    assert (dev.nrow, dev.ncol) == (O.nrow, O.ncol)
    if O.format == orc.ELL:
        assert got[0] == "ellpack"
        assert np.array_equal(got[1], O.degrees)
        assert np.array_equal(got[2], O.node)
        assert np.array_equal(got[3], O.val.reshape(O.node.shape))
    else:
        assert got[0] == ("csr" if O.format == orc.CSR else "csc")
        assert np.array_equal(got[1], O.ptr)
        assert np.array_equal(got[2], O.node)
        assert np.array_equal(got[3], O.val)


@pytest.mark.parametrize("nn", [64, 2500])
@pytest.mark.parametrize("frmt1", FMTS)
@pytest.mark.parametrize("frmt2", FMTS)
def test_copy_matrix_all_formats(sb, orc, nn, frmt1, frmt2):
    """test/matrix_test_copy.f90:83-146 (nn = 64 there), every format to every format,
    straight and transposed."""
    rng = np.random.default_rng(17 + nn)
    n, m, ptr, node, val = random_rect_csr(nn, rng)
    base = orc.Matrix(orc.CSR, n, m, node, val, ptr=ptr)
    OA = orc.copy_matrix(base, fmt_code(orc, frmt1))
    A = product_matrix(sb, OA, orc)
    x, xt = rng.standard_normal(m), rng.standard_normal(n)
    for trans in (False, True):
        B = A.copy_matrix(frmt2, trans)
        OB = orc.copy_matrix(OA, fmt_code(orc, frmt2), trans)
        same_arrays(orc, B, OB)
        assert B.nnz == node.size
        # the copy is a working operator: matvec / matvec_t bit for bit against the oracle on
        # the oracle's copy
        xin, xin_t = (xt, x) if trans else (x, xt)
        assert np.array_equal(B.matvec(xin), orc.matvec(OB, xin))
        if frmt2 != "ellpack":     # ellpack matvec_t sums padding too (covered in test_gpu_spmv)
            assert np.array_equal(B.matvec_t(xin_t), orc.matvec(OB, xin_t, trans=True))
        B.destroy()
    A.destroy()


def test_copy_chain_round_trip_at_scale(sb, orc):
    """2-D Poisson 512^2 (262 144 rows): csr -> csc -> ellpack -> csr on the device returns
    the original arrays; the transposed copy of the transposed copy does too (size-
    independent properties; the oracle's O(ne d) builders are only run at small sizes)."""
    N = 512
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    val = val + 1e-3 * np.random.default_rng(1).standard_normal(val.size)   # distinct values
    A = sb.csr_matrix(n, n, ptr, node, val)
    B = A.copy_matrix("csc")
    E = B.copy_matrix("ellpack")
    back = E.copy_matrix("csr")
    fmt, p2, n2, v2 = back.arrays()
    # csr -> csc sorts each column by row, csc -> (rows) sorts each row by column: the
    # round trip returns every row in ascending column order, which poisson2d_csr already is
    assert fmt == "csr" and np.array_equal(p2, ptr) and np.array_equal(n2, node) and np.array_equal(v2, val)
    _, edeg, enode, eval_ = E.arrays()
    gnode, gdeg, gval = G.csr_to_ell(ptr, node, val)
    assert np.array_equal(edeg, gdeg) and np.array_equal(enode, gnode) and np.array_equal(eval_, gval)
    T = A.copy_matrix("csr", trans=True)
    TT = T.copy_matrix("csr", trans=True)
    _, p3, n3, v3 = TT.arrays()
    assert np.array_equal(p3, ptr) and np.array_equal(n3, node) and np.array_equal(v3, val)
    x = np.random.default_rng(2).standard_normal(n)
    y = A.matvec(x)
    assert np.array_equal(B.matvec(x), y) and np.array_equal(E.matvec(x), y) and np.array_equal(back.matvec(x), y)
    assert np.array_equal(T.matvec_t(x), y)


def test_copy_errors_and_edge_cases(sb, orc):
    # an ellpack copy with an empty row is refused (the reference would read x(0))
    A = sb.csr_matrix(3, 3, [1, 2, 2, 3], [1, 3], [1.0, 2.0])
    with pytest.raises(sb.SigmaError) as e:
        A.copy_matrix("ellpack")
    assert e.value.status == 5
    B = A.copy_matrix("csc")                       # empty lines are fine for compressed targets
    fmt, p, nd, v = B.arrays()
    assert fmt == "csc" and np.array_equal(p, [1, 2, 2, 3]) and np.array_equal(nd, [1, 3]) and np.array_equal(v, [1.0, 2.0])
    assert np.array_equal(B.matvec(np.array([1.0, 1.0, 1.0])), [1.0, 0.0, 2.0])
    # expressions have no stored arrays to copy
    with pytest.raises(sb.SigmaError) as e:
        (A + A).copy_matrix("csr")
    assert e.value.status == 7
    # solvers accept a converted matrix: CG on the ellpack copy of a csr Poisson matrix
    N = 48
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    P = sb.csr_matrix(n, n, ptr, node, val)
    E = P.copy_matrix("ellpack")
    xs = np.random.default_rng(3).random(n)
    b = P.matvec(xs)
    s = sb.cg(1e-12 * np.linalg.norm(b))
    s.set_max_iterations(10 * n)
    s.setup(E)
    x = s.solve(E, np.zeros(n), b)
    assert not s.info()[2] and np.abs(x - xs).max() <= 1e-9
