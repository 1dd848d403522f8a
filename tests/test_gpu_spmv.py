"""Parity of the CUDA matvec path (through the C-ABI) against the oracle.
fp64 SpMV bar from north_star: 1e-12 relative per entry; the stream kernel is
built to be BIT-EXACT (rounded products added in stored order), so most checks
here are array_equal."""
import numpy as np
import pytest

from sigma_b200 import generators as G

pytestmark = pytest.mark.gpu


def _cases():
    yield "tridiag127", 127, *G.tridiag_csr(127)
    yield "poisson64", 64 * 64, *G.poisson2d_csr(64)
    yield "poisson300", 300 * 300, *G.poisson2d_csr(300)
    yield "er_rand_skew", 5000, *G.erdos_renyi_csr(5000, seed=1, weights="random", skew=True)
    yield "er_dense_rows", 700, *G.erdos_renyi_csr(700, p=0.2, seed=2, weights="random")
    yield "fem33", 33 * 33, *G.fem_p1_csr(33)
    yield "one", 1, np.array([1, 2], np.int32), np.array([1], np.int32), np.array([3.5])


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_csr_matvec_bit_exact(sb, orc, case):
    _, n, ptr, node, val = case
    rng = np.random.default_rng(0)
    x, y0 = rng.standard_normal(n), rng.standard_normal(n)
    A = sb.csr_matrix(n, n, ptr, node, val)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    assert np.array_equal(A.matvec(x), orc.matvec(O, x))
    assert np.array_equal(A.matvec_add(x, y0), orc.matvec_add(O, x, y0))
    # matvec_t of a csr_matrix runs csc_matvec_add on the same arrays (cs_matrices.f90:149)
    assert np.array_equal(A.matvec_t(x), orc.matvec(O, x, trans=True))
    assert np.array_equal(A.matvec_t_add(x, y0), orc.matvec_add(O, x, y0, trans=True))


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_csc_matvec_bit_exact_and_transpose_indices(sb, orc, case):
    _, n, ptr, node, val = case
    cptr, cnode, cval = G.csr_transpose(n, n, ptr, node, val)   # CSC storage of the same matrix
    rng = np.random.default_rng(1)
    x, y0 = rng.standard_normal(n), rng.standard_normal(n)
    A = sb.csc_matrix(n, n, cptr, cnode, cval)
    O = orc.Matrix(orc.CSC, n, n, cnode, cval, ptr=cptr)
    assert np.array_equal(A.matvec(x), orc.matvec(O, x))
    assert np.array_equal(A.matvec_add(x, y0), orc.matvec_add(O, x, y0))
    assert np.array_equal(A.matvec_t(x), orc.matvec(O, x, trans=True))
    assert np.array_equal(A.matvec_t_add(x, y0), orc.matvec_add(O, x, y0, trans=True))
    # index work is bit-exact: the device transpose equals cs_graph_build(trans=.true.)
    cols = np.repeat(np.arange(1, n + 1, dtype=np.int32), np.diff(cptr))
    optr, onode, _ = orc.cs_graph_build(n, cols, cnode, trans=True)
    dptr, dnode = A.g.transpose_arrays(n, cnode.size)
    assert np.array_equal(dptr, optr) and np.array_equal(dnode, onode)
    # ... and carries the values along
    tptr, tnode, tval = G.csr_transpose(n, n, cptr, cnode, cval)
    assert np.array_equal(A.transpose_values(), tval)


@pytest.mark.parametrize("case", list(_cases())[:6], ids=lambda c: c[0])
def test_ellpack_matvec_bit_exact(sb, orc, case):
    _, n, ptr, node, val = case
    enode, edeg, eval_ = G.csr_to_ell(ptr, node, val)
    rng = np.random.default_rng(2)
    x, y0 = rng.standard_normal(n), rng.standard_normal(n)
    A = sb.ellpack_matrix(n, n, enode, edeg, eval_)
    O = orc.Matrix(orc.ELL, n, n, enode, eval_, degrees=edeg)
    assert np.array_equal(A.matvec(x), orc.matvec(O, x))
    assert np.array_equal(A.matvec_add(x, y0), orc.matvec_add(O, x, y0))
    # ellpack_matvec_t_add scatters every slot, padding included; the last
    # neighbour of many rows can therefore own a transposed row longer than a
    # tile (1021 entries in the small tile shape this pattern gets, 2045 in the large one),
    # which is summed by a fixed CTA tree (order differs): bit-exact for
    # rows within a tile, 1e-12 relative to sum|a_ij x_j| otherwise
    yt, yto = A.matvec_t(x), orc.matvec(O, x, trans=True)
    yta, ytao = A.matvec_t_add(x, y0), orc.matvec_add(O, x, y0, trans=True)
    counts = np.bincount(enode.reshape(-1) - 1, minlength=n)
    short = counts <= 1000
    assert np.array_equal(yt[short], yto[short]) and np.array_equal(yta[short], ytao[short])
    scale = orc.matvec(orc.Matrix(orc.ELL, n, n, enode, np.abs(eval_), degrees=edeg), np.abs(x), trans=True)
    assert np.all(np.abs(yt - yto) <= 1e-12 * scale) and np.all(np.abs(yta - ytao) <= 1e-12 * (scale + np.abs(y0)))


def test_rectangular_and_empty_rows(sb, orc):
    # 6 x 9 with empty rows 2 and 6, unsorted columns
    ptr = np.array([1, 4, 4, 6, 9, 10, 10], np.int32)
    node = np.array([9, 1, 4, 2, 2 + 5, 3, 8, 1, 5], np.int32)
    val = np.arange(1.0, 10.0)
    x, xt = np.linspace(-1, 1, 9), np.linspace(1, 2, 6)
    A = sb.csr_matrix(6, 9, ptr, node, val)
    O = orc.Matrix(orc.CSR, 6, 9, node, val, ptr=ptr)
    assert np.array_equal(A.matvec(x), orc.matvec(O, x))
    assert np.array_equal(A.matvec_t(xt), orc.matvec(O, xt, trans=True))
    # same matrix as csc (9 columns)
    cptr, cnode, cval = G.csr_transpose(6, 9, ptr, node, val)
    B = sb.csc_matrix(6, 9, cptr, cnode, cval)
    OB = orc.Matrix(orc.CSC, 6, 9, cnode, cval, ptr=cptr)
    assert np.array_equal(B.matvec(x), orc.matvec(OB, x))
    assert np.array_equal(B.matvec_t(xt), orc.matvec(OB, xt, trans=True))
    assert np.array_equal(B.matvec(x), A.matvec(x))


def test_long_rows_within_tolerance(sb, orc):
    """Rows longer than a tile are reduced by a fixed CTA tree (sum order differs
    from the serial loop): 1e-12 relative to sum |a_ij x_j|."""
    n = 6000
    rng = np.random.default_rng(3)
    dense_rows = [0, 2999, n - 1]
    ptr = [1]
    node, val = [], []
    for i in range(n):
        if i in dense_rows:
            cols = rng.permutation(n)[:5000] + 1
        else:
            cols = np.unique(np.array([i, (i * 7) % n, (i + 1) % n])) + 1
        node.append(cols)
        val.append(rng.standard_normal(cols.size))
        ptr.append(ptr[-1] + cols.size)
    ptr, node, val = np.array(ptr, np.int32), np.concatenate(node).astype(np.int32), np.concatenate(val)
    x = rng.standard_normal(n)
    A = sb.csr_matrix(n, n, ptr, node, val)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    y, yo = A.matvec(x), orc.matvec(O, x)
    Oabs = orc.Matrix(orc.CSR, n, n, node, np.abs(val), ptr=ptr)
    scale = orc.matvec(Oabs, np.abs(x))
    assert np.all(np.abs(y - yo) <= 1e-12 * scale)
    short = np.setdiff1d(np.arange(n), dense_rows)
    assert np.array_equal(y[short], yo[short])


def test_small_tile_shape_rows_at_the_cap(sb, orc):
    """Patterns with >= 12 stored entries per row get the small tile shape (1024 staged entries,
    <= 1021 per tile, <= 256 rows): rows of exactly 1021 entries are still summed in stored order by one
    thread (bit-exact), rows of 1022 and more take the CTA tree (1e-12); all forms of the kernel
    (set / add / transposed / csc)."""
    n = 20000
    rng = np.random.default_rng(11)
    special = {5: 1021, 6: 1, 7: 1020, 8: 1, 9: 1022, 4000: 3000, n - 1: 1021}
    ptr = [1]
    node, val = [], []
    for i in range(n):
        d = special.get(i, int(rng.integers(8, 40)))
        cols = rng.choice(n, size=d, replace=False) + 1          # unsorted, scattered
        node.append(cols)
        val.append(rng.standard_normal(d))
        ptr.append(ptr[-1] + d)
    ptr, node, val = np.array(ptr, np.int32), np.concatenate(node).astype(np.int32), np.concatenate(val)
    assert node.size >= 12 * n
    x, y0 = rng.standard_normal(n), rng.standard_normal(n)
    A = sb.csr_matrix(n, n, ptr, node, val)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    y, yo = A.matvec(x), orc.matvec(O, x)
    scale = orc.matvec(orc.Matrix(orc.CSR, n, n, node, np.abs(val), ptr=ptr), np.abs(x))
    assert np.all(np.abs(y - yo) <= 1e-12 * scale)
    tree = np.array([9, 4000])
    exact = np.setdiff1d(np.arange(n), tree)
    assert np.array_equal(y[exact], yo[exact])
    ya, yao = A.matvec_add(x, y0), orc.matvec_add(O, x, y0)
    assert np.array_equal(ya[exact], yao[exact]) and np.all(np.abs(ya - yao) <= 1e-12 * (scale + np.abs(y0)))
    # transposed forms: the device-built stable transpose has its own row lengths (all short here
    # except the columns the long rows share with nobody) and its own tiling
    assert np.allclose(A.matvec_t(x), orc.matvec(O, x, trans=True), rtol=0, atol=1e-12 * np.abs(scale).max())
    cptr, cnode, cval = G.csr_transpose(n, n, ptr, node, val)
    Ac = sb.csc_matrix(n, n, cptr, cnode, cval)
    Oc = orc.Matrix(orc.CSC, n, n, cnode, cval, ptr=cptr)
    yc, yco = Ac.matvec(x), orc.matvec(Oc, x)
    assert np.all(np.abs(yc - yco) <= 1e-12 * scale)


def test_ellpack_isolated_vertex_rejected(sb):
    node = np.array([[1, 2], [0, 0], [3, 3]], np.int32)
    deg = np.array([2, 0, 1], np.int32)
    with pytest.raises(sb.SigmaError) as e:
        sb.Graph.ellpack(3, 3, node, deg)
    assert e.value.status == 5


def test_values_refresh_after_host_mutation(sb, orc):
    """H5: a host mutator marks the mirror dirty; set_values re-uploads and the
    cached transpose follows."""
    n = 400
    ptr, node, val = G.erdos_renyi_csr(n, seed=4, weights="random")
    A = sb.csr_matrix(n, n, ptr, node, val)
    x = np.random.default_rng(5).standard_normal(n)
    _ = A.matvec_t(x)                      # builds + caches the transposed values
    val2 = val * 1.5 + 0.25                # A%scalar_multiply / add_value on the host
    A.set_values(val2)
    O = orc.Matrix(orc.CSR, n, n, node, val2, ptr=ptr)
    assert np.array_equal(A.matvec(x), orc.matvec(O, x))
    assert np.array_equal(A.matvec_t(x), orc.matvec(O, x, trans=True))


def test_full_size_poisson_matvec_bit_exact(sb):
    """BASELINE config 2/3 at full size (4096^2): y = A x against the stencil
    evaluated by numpy in the stored entry order (no FMA) -- bit-exact, CSR and
    ELLPACK, plus the linearity property A(ax + by) ~ a Ax + b Ay."""
    N = 4096
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    b, xs = G.poisson2d_rhs(N)
    A = sb.csr_matrix(n, n, ptr, node, val)
    y = A.matvec(xs)
    assert np.array_equal(y, b)
    x2 = np.cos(np.arange(n) * 1e-3)
    y2 = A.matvec(x2)
    lin = A.matvec(0.5 * xs - 2.0 * x2)
    assert np.max(np.abs(lin - (0.5 * y - 2.0 * y2))) <= 1e-12 * 12
    enode, edeg, eval_ = G.csr_to_ell(ptr, node, val)
    del ptr, node, val
    E = sb.ellpack_matrix(n, n, enode, edeg, eval_)
    assert np.array_equal(E.matvec(xs), b)
