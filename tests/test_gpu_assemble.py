"""On-device assembly: a stream of A%add_value(i, j, z) calls applied
in order, against the oracle's serial loop, through the C-ABI.  Floating-point addition
does not commute with reordering, so the bar is BIT-EXACT values: every stored entry
must have received its contributions in ascending call index."""
import numpy as np
import pytest

from sigma_b200 import generators as G

pytestmark = pytest.mark.gpu


def fem_case(N):
    I, J, V, interior = G.fem_p1_add_value_stream(N)
    ptr, node, _ = G.fem_p1_csr(N)          # the pattern the add_edge calls of the same loop build
    return N * N, ptr, node, (I + 1).astype(np.int32), (J + 1).astype(np.int32), V, interior


def in_format(sb, orc, fmt, n, ptr, node, val):
    if fmt == "csr":
        return sb.csr_matrix(n, n, ptr, node, val), orc.Matrix(orc.CSR, n, n, node, val.copy(), ptr=ptr)
    if fmt == "csc":
        cptr, cnode, cval = G.csr_transpose(n, n, ptr, node, val)
        return sb.csc_matrix(n, n, cptr, cnode, cval), orc.Matrix(orc.CSC, n, n, cnode, cval.copy(), ptr=cptr)
    enode, edeg, eval_ = G.csr_to_ell(ptr, node, val)
    return sb.ellpack_matrix(n, n, enode, edeg, eval_), orc.Matrix(orc.ELL, n, n, enode, eval_.copy(), degrees=edeg)


@pytest.mark.parametrize("fmt", ["csr", "csc", "ellpack"])
def test_fem_assembly_stream(sb, orc, fmt):
    """laplacian2d of examples/fem.f90:28-49 on a 65 x 65 vertex grid: 8192 elements, 73 728
    add_value calls, every interior entry touched by 2-8 of them."""
    n, ptr, node, ci, cj, cz, _ = fem_case(65)
    A, O = in_format(sb, orc, fmt, n, ptr, node, np.zeros(node.size))
    assert orc.add_values(O, ci, cj, cz) == 0
    A.add_values(ci, cj, cz)
    got = A.arrays()
    assert np.array_equal(np.asarray(got[3]).reshape(-1), O.val.reshape(-1))
    # the assembled operator works, transposed mirror included (csc keeps one)
    x = np.random.default_rng(0).standard_normal(n)
    assert np.array_equal(A.matvec(x), orc.matvec(O, x))
    # a second batch accumulates on top of the first, again in order
    half = ci.size // 2
    assert orc.add_values(O, ci[:half], cj[:half], 0.5 * cz[:half]) == 0
    A.add_values(ci[:half], cj[:half], 0.5 * cz[:half])
    assert np.array_equal(np.asarray(A.arrays()[3]).reshape(-1), O.val.reshape(-1))
    assert np.array_equal(A.matvec(x), orc.matvec(O, x))


def test_order_sensitive_accumulation(sb, orc):
    """20 000 contributions of wildly different magnitude into 7 entries: any reordering of
    the additions changes the low bits."""
    n = 5
    ptr = np.array([1, 3, 4, 6, 7, 8], np.int32)
    node = np.array([1, 4, 2, 3, 5, 1, 5], np.int32)
    val = np.linspace(-1.0, 1.0, 7)
    A, O = in_format(sb, orc, "csr", n, ptr, node, val)
    rng = np.random.default_rng(12)
    k = rng.integers(0, 7, 20000)
    rows = np.repeat(np.arange(1, n + 1), np.diff(ptr))
    ci, cj = rows[k].astype(np.int32), node[k]
    cz = rng.standard_normal(20000) * 10.0 ** rng.integers(-8, 9, 20000)
    assert orc.add_values(O, ci, cj, cz) == 0
    A.add_values(ci, cj, cz)
    assert np.array_equal(A.arrays()[3], O.val)
    # sanity of the test itself: a different order really gives different bits
    P = orc.Matrix(orc.CSR, n, n, node, val.copy(), ptr=ptr)
    orc.add_values(P, ci[::-1].copy(), cj[::-1].copy(), cz[::-1].copy())
    assert not np.array_equal(P.val, O.val)


def test_entries_outside_the_pattern_are_refused(sb, orc):
    n, ptr, node, ci, cj, cz, _ = fem_case(9)
    A, O = in_format(sb, orc, "csr", n, ptr, node, np.ones(node.size))
    bad_i, bad_j = ci.copy(), cj.copy()
    bad_i[100], bad_j[100] = 1, n            # vertex 1 and vertex n share no element
    bad_i[7], bad_j[7] = n + 5, 1            # out of range altogether
    with pytest.raises(sb.SigmaError) as e:
        A.add_values(bad_i, bad_j, cz)
    assert e.value.status == 1 and "2 of" in e.value.message and "call 8" in e.value.message
    assert np.array_equal(A.arrays()[3], np.ones(node.size))      # nothing was added
    A.add_values(ci[:0], cj[:0], cz[:0])                           # an empty batch is a no-op
    assert np.array_equal(A.arrays()[3], np.ones(node.size))


def test_assembly_at_scale_matches_the_generator(sb):
    """513 x 513 vertices: 524 288 elements, 4.7 M calls.  The input generator accumulates the
    same stream with a sequential numpy add.at; away from the Dirichlet rows the assembled
    values must agree bit for bit, and the operator must be symmetric with zero row sums."""
    N = 513
    n, ptr, node, ci, cj, cz, interior = fem_case(N)
    _, _, want = G.fem_p1_csr(N)
    A = sb.csr_matrix(n, n, ptr, node, np.zeros(node.size))
    A.add_values(ci, cj, cz)
    got = A.arrays()[3]
    rows = np.repeat(np.arange(n), np.diff(ptr))
    bnd = ~interior.reshape(-1)
    keep = ~(bnd[rows] | bnd[node - 1])
    assert keep.sum() > 0.95 * node.size
    assert np.array_equal(got[keep], want[keep])
    ones = np.ones(n)
    y = A.matvec(ones)
    assert np.abs(y).max() <= 1e-11                                # constants are in the kernel
    x = np.random.default_rng(4).standard_normal(n)
    assert np.allclose(A.matvec(x), A.matvec_t(x), rtol=0, atol=1e-11)
