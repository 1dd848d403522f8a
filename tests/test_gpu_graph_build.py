"""The graph builders on the device from an EDGE STREAM (csrc/graph_build.cu) against the oracle's
restatement of cs_graph_build (src/graph/formats/cs_graphs.f90:109-197) and ellpack_graph_build
(src/graph/formats/ellpack_graphs.f90:105-170): index work, compared bit for bit -- the order inside a
line is the order of first appearance in the stream, which is what fixes the summation order of every
later matvec."""
import numpy as np
import pytest

from sigma_b200 import generators as G

pytestmark = pytest.mark.gpu


def streams(orc):
    """(name, n, src_i, src_j): edge streams in the iteration order of the reference's ll_graph."""
    ei, ej = G.tridiag_add_edge_calls(127)                       # test/solver_test_diffusion_1d.f90:60-65
    yield "diffusion_1d", 127, *orc.ll_graph_edges(127, ei, ej)[:2]
    ei, ej = G.poisson2d_add_edge_calls(40)
    yield "poisson40", 1600, *orc.ll_graph_edges(1600, ei, ej)[:2]
    ptr, node, _, (pi, pj) = G.erdos_renyi_csr(3000, seed=5, return_pairs=True)
    ei, ej = G.erdos_renyi_add_edge_calls(3000, pi, pj)
    yield "erdos_renyi", 3000, *orc.ll_graph_edges(3000, ei, ej)[:2]
    rng = np.random.default_rng(3)                                # an arbitrary stream: any order, no duplicates
    cells = rng.choice(500 * 500, 6000, replace=False)
    yield "random_order", 500, (cells // 500 + 1).astype(np.int32), (cells % 500 + 1).astype(np.int32)


def first_appearance(n, e1, e2):
    """What the reference's first-free-slot insertion + duplicate check + pruning leaves (:163-189)."""
    rows = [[] for _ in range(n)]
    for a, b in zip(e1, e2):
        if b != 0 and b not in rows[a - 1]:
            rows[a - 1].append(int(b))
    ptr = np.concatenate([[1], 1 + np.cumsum([len(r) for r in rows])]).astype(np.int32)
    node = np.array([c for r in rows for c in r], np.int32)
    return ptr, node


def arrays_of(sb, g):
    A = sb.Matrix(g)
    out = A.arrays()
    A.destroy()
    return out


@pytest.mark.parametrize("trans", [False, True])
def test_cs_graph_build_equals_the_serial_builder(sb, orc, trans):
    for name, n, si, sj in streams(orc):
        optr, onode, omd = orc.cs_graph_build(n, si, sj, trans)
        for frmt in ("csr", "csc"):
            g = sb.Graph.build(frmt, n, n, si, sj, trans)
            fmt, p, nd, _ = arrays_of(sb, g)
            assert fmt == frmt and np.array_equal(p, optr) and np.array_equal(nd, onode), (name, frmt)


def test_cs_graph_build_duplicates_and_null_edges(sb, orc):
    """Iterators that return an edge twice, and null edges (endpoint 0): the reference keeps the first
    occurrence, skips the nulls and prunes the unused slots."""
    rng = np.random.default_rng(11)
    n, m = 300, 200
    si = rng.integers(1, n + 1, 9000).astype(np.int32)
    sj = rng.integers(0, m + 1, 9000).astype(np.int32)           # plenty of duplicates, some zeros
    for trans, (e1, e2, nn, mm) in ((False, (si, sj, n, m)),):
        g = sb.Graph.build("csr", nn, mm, si, sj, trans)
        _, p, nd, _ = arrays_of(sb, g)
        wp, wn = first_appearance(nn, e1, e2)
        assert np.array_equal(p, wp) and np.array_equal(nd, wn)
    # empty stream, empty graph
    g = sb.Graph.build("csr", 7, 7, np.zeros(0, np.int32), np.zeros(0, np.int32))
    _, p, nd, _ = arrays_of(sb, g)
    assert np.array_equal(p, np.ones(8, np.int32)) and nd.size == 0
    with pytest.raises(sb.SigmaError) as e:                      # an edge that starts outside the graph
        sb.Graph.build("csr", 5, 5, [6], [1])
    assert e.value.status == 1


@pytest.mark.parametrize("trans", [False, True])
def test_ellpack_graph_build_equals_the_serial_builder(sb, orc, trans):
    for name, n, si, sj in streams(orc):
        if name == "random_order":
            continue                                             # has rows without edges: refused, see below
        onode, odeg = orc.ellpack_graph_build(n, si, sj, trans)
        g = sb.Graph.build("ellpack", n, n, si, sj, trans)
        g.max_d = onode.shape[1]
        fmt, deg, nd, _ = arrays_of(sb, g)
        assert fmt == "ellpack" and np.array_equal(deg, odeg) and np.array_equal(nd, onode), name
    with pytest.raises(sb.SigmaError) as e:                      # the reference would read x(0) (README.md:71-73)
        sb.Graph.build("ellpack", 3, 3, [1, 2], [2, 1])
    assert e.value.status == 5


def test_matvec_on_a_device_built_pattern(sb, orc):
    """The built pattern is a working graph: values set through the ordered add_value stream, matvec
    bit-identical to the serial loop on the oracle's arrays."""
    ei, ej = G.poisson2d_add_edge_calls(30)
    si, sj, _ = orc.ll_graph_edges(900, ei, ej)
    optr, onode, _ = orc.cs_graph_build(900, si, sj)
    oval = np.where(onode == np.repeat(np.arange(1, 901), np.diff(optr)), 4.0, -1.0)
    x = np.random.default_rng(0).standard_normal(900)
    want = orc.matvec(orc.Matrix(orc.CSR, 900, 900, onode, oval, ptr=optr), x)
    for frmt in ("csr", "ellpack"):
        A = sb.Matrix(sb.Graph.build(frmt, 900, 900, si, sj))
        A.add_values(si, sj, np.where(si == sj, 4.0, -1.0))
        assert np.array_equal(A.matvec(x), want), frmt
