"""Reference test programs restated on the device (first run in round 2, visit r2a):
test/matrix_test_strategy.f90 and test/matrix_test_set_multiple_entries.f90 through the C-ABI
(CPU twins: tests/test_oracle_strategy.py), and the device-built row tiling against the host's."""
import ctypes as C

import numpy as np
import pytest

import test_oracle_strategy as S
import test_row_tiles as T

pytestmark = pytest.mark.gpu


def test_matrix_test_strategy_on_the_device(sb, orc):
    """test/matrix_test_strategy.f90: a sparse_matrix holding one storage strategy; entries against
    the graph (exact), A%matvec against the Laplacian written out from the neighbour lists (1e-14)."""
    nn = 256
    for fmt in ("csr", "csc", "ellpack"):
        O, nbrs, connected = S.strategy_case(orc, nn, 11, fmt)
        # the same storage, zeroed, assembled on the device with the test's add_value stream
        if fmt == "ellpack":
            A = sb.ellpack_matrix(nn, nn, O.node, O.degrees, np.zeros(O.node.shape))
        elif fmt == "csr":
            A = sb.csr_matrix(nn, nn, O.ptr, O.node, np.zeros(O.node.size))
        else:
            A = sb.csc_matrix(nn, nn, O.ptr, O.node, np.zeros(O.node.size))
        ci, cj, cz = [], [], []
        for v in range(1, nn + 1):
            for w in nbrs[v - 1]:
                ci += [v, v]; cj += [w, v]; cz += [-1.0, 1.0]
        A.add_values(ci, cj, cz)
        assert np.array_equal(A.arrays()[-1].reshape(-1), np.asarray(O.val).reshape(-1)), fmt
        # type(sparse_matrix) as the container of one storage strategy: a 1 x 1 composite
        M = sb.sparse_matrix([nn], [nn], [[A]])
        x = np.random.default_rng(5).random(nn)
        y = np.array([len(nbrs[i]) * x[i] - sum(x[w - 1] for w in nbrs[i]) for i in range(nn)])
        for op in (A, M):
            w_ = op.matvec(x)
            assert np.array_equal(w_, orc.matvec(O, x)), fmt
            assert np.sqrt(((y - w_) ** 2).sum() / (x @ x)) <= 1e-14, fmt


def test_matrix_test_set_multiple_entries_on_the_device(sb, orc):
    """test/matrix_test_set_multiple_entries.f90: the Laplacian assembled from 2 x 2 element blocks
    with add_multiple_values (one device call for all blocks); exact comparisons."""
    nn = 128
    for fmt in ("csr", "csc", "ellpack"):
        O, nbrs, connected = S.strategy_case(orc, nn, 23, fmt)
        O.val[...] = 0.0
        if fmt == "ellpack":
            A = sb.ellpack_matrix(nn, nn, O.node, O.degrees, np.zeros(O.node.shape))
        elif fmt == "csr":
            A = sb.csr_matrix(nn, nn, O.ptr, O.node, np.zeros(O.node.size))
        else:
            A = sb.csc_matrix(nn, nn, O.ptr, O.node, np.zeros(O.node.size))
        pairs = np.array([(i, j) for i in range(1, nn + 1) for j in nbrs[i - 1] if j > i], np.int32)
        B = np.broadcast_to(np.array([[1.0, -1.0], [-1.0, 1.0]]), (pairs.shape[0], 2, 2))
        A.add_multiple_values(pairs, pairs, B)               # all element blocks, one device call
        assert orc.add_values(O, *sb.multiple_values_stream(pairs, pairs, B)) == 0
        assert np.array_equal(A.arrays()[-1].reshape(-1), np.asarray(O.val).reshape(-1)), fmt
        # the checks of the test on the device result, through the matrix's own matvec: row sums
        # of a graph Laplacian vanish, and A e_i reads column i
        assert np.array_equal(A.matvec(np.ones(nn)), np.zeros(nn)), fmt
        for i in (1, 17, nn):
            e = np.zeros(nn); e[i - 1] = 1.0
            col = A.matvec(e)
            want = np.where(connected[:, i - 1], -1.0, 0.0); want[i - 1] = len(nbrs[i - 1]) - 1
            assert np.array_equal(col, want), (fmt, i)


def test_device_built_tiles_equal_the_host_tiling(sb):
    """csrc/tiles_device.cu (bisection per row, pointer doubling, compaction) against the greedy
    host walk of build_tiles_host, entry for entry, on every tiling case of test_row_tiles."""
    from sigma_b200._capi import check, lib, ptr

    for name, p in T.cases():
        p = np.ascontiguousarray(p, np.int32)
        n = p.size - 1
        out = np.empty((max(n, 1), 4), np.int32)
        nt = C.c_int32()
        check(lib().sigb_debug_row_tiles_dev(n, ptr(p), ptr(out), C.byref(nt)))
        want = T.library_tiles(p)
        assert nt.value == want.shape[0] and np.array_equal(out[: nt.value], want), name
