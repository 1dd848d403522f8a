"""Shared builders for the tests: oracle matrices and product matrices from the
same arrays."""
import numpy as np


def ell_from_tridiag_calls(orc, nn, diag, upper, lower):
    """Build the 1-D operator exactly as the reference tests do: ll_graph
    add_edge calls -> ellpack graph -> set_value calls."""
    from sigma_b200 import generators as G

    ei, ej = G.tridiag_add_edge_calls(nn)
    si, sj, _ = orc.ll_graph_edges(nn, ei, ej)
    node, deg = orc.ellpack_graph_build(nn, si, sj)
    val = np.zeros(node.shape)
    for i in range(1, nn):
        assert orc.ell_set_value(node, deg, val, i, i, diag)
        assert orc.ell_set_value(node, deg, val, i, i + 1, upper)
        assert orc.ell_set_value(node, deg, val, i + 1, i, lower)
    assert orc.ell_set_value(node, deg, val, nn, nn, diag)
    return node, deg, val


def csr_from_calls(orc, n, ei, ej):
    si, sj, _ = orc.ll_graph_edges(n, ei, ej)
    ptr, node, md = orc.cs_graph_build(n, si, sj)
    return ptr, node
