"""test/matrix_test_strategy.f90 restated against the oracle on CPU: an Erdos-Renyi graph built
with the test's own add_edge order (:74-87), its Laplacian assembled with the test's own
add_value calls (:109-117) into csr / csc / ellpack storage ("the concrete strategy"), then the
checks of the test: get_value against the graph (:127-153, exact), and A%matvec against the
Laplacian written out from the neighbour lists (:225-254, relative error 1e-14).  The reference
draws the graph from a time-seeded RNG: seeded inputs of the same shape here."""
import numpy as np
import pytest

from sigma_b200 import generators as G


def strategy_case(orc, nn, seed, fmt):
    rng = np.random.default_rng(seed)
    c = np.log2(nn) / nn
    upper = np.triu(rng.random((nn, nn)) < c, 1)
    i, j = np.nonzero(upper)                                   # i < j, ascending i then j: the loop order of :74-87
    ei, ej = G.erdos_renyi_add_edge_calls(nn, i, j)
    si, sj, _ = orc.ll_graph_edges(nn, ei, ej)                 # the graph's iteration order
    # neighbour lists in g%get_neighbors order (= insertion order per vertex)
    nbrs = [[] for _ in range(nn)]
    for a, b in zip(ei, ej):
        nbrs[a - 1].append(int(b))
    if fmt == "ellpack":
        node, deg = orc.ellpack_graph_build(nn, si, sj)
        A = orc.Matrix(orc.ELL, nn, nn, node, np.zeros(node.shape), degrees=deg)
    else:
        ptr, node, _ = orc.cs_graph_build(nn, si, sj, trans=(fmt == "csc"))
        A = orc.Matrix(orc.CSR if fmt == "csr" else orc.CSC, nn, nn, node, np.zeros(node.size), ptr=ptr)
    # call A%add_value(i, j, -1) ; call A%add_value(i, i, +1) for every neighbour j of i (:109-117)
    ci, cj, cz = [], [], []
    for v in range(1, nn + 1):
        for w in nbrs[v - 1]:
            ci += [v, v]
            cj += [w, v]
            cz += [-1.0, 1.0]
    assert orc.add_values(A, ci, cj, cz) == 0
    return A, nbrs, upper | upper.T


@pytest.mark.parametrize("fmt", ["csr", "csc", "ellpack"])
def test_matrix_test_strategy(orc, fmt):
    nn = 256
    A, nbrs, connected = strategy_case(orc, nn, 11, fmt)
    # entries (:127-153): the self edge gets -1 + 1 per neighbour incl. itself => degree - 1
    for i in range(1, nn + 1, 7):
        assert orc.get_value(A, i, i) == len(nbrs[i - 1]) - 1
        for j in range(i + 1, nn + 1):
            assert orc.get_value(A, i, j) == (-1.0 if connected[i - 1, j - 1] else 0.0)
    # matvec against the Laplacian written out from the neighbour lists (:225-254)
    x = np.random.default_rng(5).random(nn)
    y = np.empty(nn)
    for i in range(nn):
        z = len(nbrs[i]) * x[i]
        for w in nbrs[i]:
            z = z - x[w - 1]
        y[i] = z
    w_ = orc.matvec(A, x)
    assert np.sqrt(((y - w_) ** 2).sum() / (x @ x)) <= 1e-14


def multiple_entries_stream(nbrs, nn):
    """The add_multiple_values calls of test/matrix_test_set_multiple_entries.f90:103-114: for
    every vertex i and every neighbour j > i (in g%get_neighbors order), is = [i, j] and
    B = [[1, -1], [-1, 1]]."""
    import sigma_b200 as sb

    pairs = [(i, j) for i in range(1, nn + 1) for j in nbrs[i - 1] if j > i]
    is_ = np.array(pairs, np.int32).reshape(-1, 2)
    B = np.broadcast_to(np.array([[1.0, -1.0], [-1.0, 1.0]]), (is_.shape[0], 2, 2))
    return sb.multiple_values_stream(is_, is_, B)


def test_multiple_values_stream_order():
    import sigma_b200 as sb

    I, J, Z = sb.multiple_values_stream([3, 7], [1, 2, 5], [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    assert I.tolist() == [3, 3, 3, 7, 7, 7] and J.tolist() == [1, 2, 5, 1, 2, 5]      # k outer, l inner
    assert Z.tolist() == [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    I, J, Z = sb.multiple_values_stream([[1, 2], [2, 3]], [[1, 2], [2, 3]], np.arange(8.0).reshape(2, 2, 2))
    assert I.tolist() == [1, 1, 2, 2, 2, 2, 3, 3] and J.tolist() == [1, 2, 1, 2, 2, 3, 2, 3]
    assert Z.tolist() == list(np.arange(8.0))


@pytest.mark.parametrize("fmt", ["csr", "csc", "ellpack"])
def test_matrix_test_set_multiple_entries(orc, fmt):
    """test/matrix_test_set_multiple_entries.f90: the graph Laplacian assembled from 2 x 2 element
    blocks with add_multiple_values must have degree - 1 on the diagonal, -1 on every edge and
    nothing else (:120-152, exact comparisons)."""
    nn = 128
    A, nbrs, connected = strategy_case(orc, nn, 23, fmt)
    A.val[...] = 0.0                                        # call A%zero()
    I, J, Z = multiple_entries_stream(nbrs, nn)
    assert orc.add_values(A, I, J, Z) == 0
    for i in range(1, nn + 1):
        for j in range(1, nn + 1):
            want = len(nbrs[i - 1]) - 1 if i == j else (-1.0 if connected[i - 1, j - 1] else 0.0)
            assert orc.get_value(A, i, j) == want
