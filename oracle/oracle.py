"""ctypes front-end of oracle/sigma_oracle.c (TEST INFRASTRUCTURE, not product).

All index arrays are 1-based int32 numpy arrays laid out exactly as the
Fortran reference holds them (see the C file's header).  Every wrapper names
the C function it calls; the C function cites the reference file:line.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsigma_oracle.so")

CSR, CSC, ELL = 1, 2, 3
SUM, PRODUCT, ADJOINT, COMPOSITE = 4, 5, 6, 7

__all__ = [
    "CSR", "CSC", "ELL", "build", "lib", "Matrix", "ll_graph_edges", "cs_graph_build",
    "ellpack_graph_build", "matvec", "matvec_add", "jacobi_setup", "jacobi_solve",
    "cg_solve", "bicgstab_solve", "lanczos", "generalized_lanczos", "eigensolve", "tridiag_eig",
    "partition_rows", "halo_build", "cs_set_value", "ell_set_value",
    "SUM", "PRODUCT", "ADJOINT", "COMPOSITE", "operator_sum", "operator_product", "adjoint",
    "composite", "get_value", "matrix_entries", "copy_matrix", "add_values", "Ldu", "ldu_setup", "ldu_solve", "cg_solve_ldu", "bicgstab_solve_ldu",
]


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc, strict fp flags)."""
    src = os.path.join(_HERE, "sigma_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


class _OrcMatrix(C.Structure):
    _fields_ = [
        ("format", C.c_int32), ("nrow", C.c_int32), ("ncol", C.c_int32), ("max_d", C.c_int32),
        ("ptr", C.c_void_p), ("node", C.c_void_p), ("degrees", C.c_void_p), ("val", C.c_void_p),
        # operator expressions (orc_matrix in sigma_oracle.c)
        ("nkids", C.c_int32), ("num_row_mats", C.c_int32), ("num_col_mats", C.c_int32),
        ("temp_vec_size", C.c_int32),
        ("kids", C.c_void_p), ("row_ptr", C.c_void_p), ("col_ptr", C.c_void_p),
        ("z1", C.c_void_p), ("z2", C.c_void_p),
    ]


_lib = None
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    i32, i64, f64, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
    mp = C.POINTER(_OrcMatrix)
    sig = {
        "orc_ll_graph_edges": (i64, [i32, i64, _i32p, _i32p, _i32p, _i32p, C.POINTER(i32)]),
        "orc_cs_graph_build": (i32, [i32, i64, _i32p, _i32p, i32, _i32p, _i32p]),
        "orc_ellpack_max_degree": (i32, [i32, i64, _i32p, _i32p, i32]),
        "orc_ellpack_graph_build": (None, [i32, i64, _i32p, _i32p, i32, i32, _i32p, _i32p]),
        "orc_cs_set_value": (i32, [_i32p, _i32p, _f64p, i32, i32, f64, i32]),
        "orc_cs_get_value": (f64, [_i32p, _i32p, _f64p, i32, i32]),
        "orc_ell_set_value": (i32, [i32, _i32p, _i32p, _f64p, i32, i32, f64, i32]),
        "orc_ell_get_value": (f64, [i32, _i32p, _i32p, _f64p, i32, i32]),
        "orc_copy_matrix_values": (i64, [i32, i32, vp, _i32p, vp, _f64p, i64, _i32p, _i32p, _f64p, i32]),
        "orc_add_values": (i64, [i32, i32, vp, _i32p, vp, _f64p, i64, _i32p, _i32p, _f64p]),
        "orc_ldu_factor": (None, [i32, i64, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, _f64p]),
        "orc_ldu_solve": (None, [i32, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p]),
        "orc_cg_solve_ldu": (i64, [mp, _f64p, _f64p, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, _f64p, f64, i64,
                                   _f64p, C.POINTER(f64), C.POINTER(i32)]),
        "orc_bicgstab_solve_ldu": (i64, [mp, _f64p, _f64p, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, _f64p, f64, i64,
                                         _f64p, C.POINTER(f64), C.POINTER(i32)]),
        "orc_matvec_add": (None, [mp, i32, _f64p, _f64p]),
        "orc_matvec": (None, [mp, i32, _f64p, _f64p]),
        "orc_get_value": (f64, [mp, i32, i32]),
        "orc_jacobi_setup": (None, [mp, _f64p]),
        "orc_jacobi_solve": (None, [i32, _f64p, _f64p, _f64p]),
        "orc_cg_solve": (i64, [mp, _f64p, _f64p, f64, i64, _f64p, C.POINTER(f64), C.POINTER(i32)]),
        "orc_cg_solve_jacobi": (i64, [mp, _f64p, _f64p, _f64p, f64, i64, _f64p, C.POINTER(f64), C.POINTER(i32)]),
        "orc_bicgstab_solve": (i64, [mp, _f64p, _f64p, f64, i64, _f64p, C.POINTER(f64), C.POINTER(i32)]),
        "orc_bicgstab_solve_jacobi": (i64, [mp, _f64p, _f64p, _f64p, f64, i64, _f64p, C.POINTER(f64), C.POINTER(i32)]),
        "orc_lanczos": (None, [mp, i32, _f64p, _f64p, _f64p, _f64p]),
        "orc_generalized_lanczos": (i64, [mp, mp, i32, _f64p, f64, i64, _f64p, _f64p]),
        "orc_tridiag_eig": (i32, [i32, _f64p, _f64p, _f64p]),
        "orc_eigensolve": (i32, [mp, i32, _f64p, _f64p, _f64p]),
        "orc_partition_rows": (None, [i32, _i32p, i32, _i32p]),
        "orc_halo_build": (i32, [i32, i32, _i32p, _i32p, _i32p, _i32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _ = vp
    _lib = L
    return L


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Matrix:
    """Host matrix in the reference's own storage.

    CSR/CSC: ``ptr`` (n+1), ``node`` (ne), ``val`` (ne), 1-based.
    ELL: ``node``/``val`` of shape (nrow, max_d) C-order == Fortran
    ``node(max_d, nrow)`` column-major; ``degrees`` (nrow).
    """

    def __init__(self, fmt, nrow, ncol, node, val, ptr=None, degrees=None):
        self.format, self.nrow, self.ncol = fmt, int(nrow), int(ncol)
        self.node, self.val = _i32(node), _f64(val)
        self.ptr = _i32(ptr) if ptr is not None else None
        self.degrees = _i32(degrees) if degrees is not None else None
        self.max_d = int(self.node.shape[1]) if fmt == ELL else 0
        self._c = _OrcMatrix(
            fmt, self.nrow, self.ncol, self.max_d,
            self.ptr.ctypes.data if self.ptr is not None else None,
            self.node.ctypes.data,
            self.degrees.ctypes.data if self.degrees is not None else None,
            self.val.ctypes.data,
        )

    @property
    def c(self):
        return C.byref(self._c)


class Operator:
    """An operator expression over host matrices: operator_sum, operator_product,
    operator_adjoint (src/linear_operator/linear_operator_{sums,products,adjoints}.f90)
    or a block composite sparse_matrix (src/matrix/sparse_matrix_composites.f90).
    Usable wherever a Matrix is (matvec, get_value, jacobi_setup, the solvers)."""

    def __init__(self, fmt, nrow, ncol, kids, row_ptr=None, col_ptr=None):
        self.format, self.nrow, self.ncol = fmt, int(nrow), int(ncol)
        self.kids = list(kids)
        self._kid_arr = (C.c_void_p * len(self.kids))(*[C.addressof(k._c) for k in self.kids])
        self.row_ptr = _i32(row_ptr) if row_ptr is not None else None
        self.col_ptr = _i32(col_ptr) if col_ptr is not None else None
        tvs = 0
        self._z1 = self._z2 = None
        if fmt == PRODUCT:
            tvs = max(max(k.nrow, k.ncol) for k in self.kids)
            self._z1, self._z2 = np.zeros(tvs), np.zeros(tvs)
        self._c = _OrcMatrix(
            fmt, self.nrow, self.ncol, 0, None, None, None, None,
            len(self.kids),
            (self.row_ptr.size - 1) if self.row_ptr is not None else 0,
            (self.col_ptr.size - 1) if self.col_ptr is not None else 0,
            tvs,
            C.cast(self._kid_arr, C.c_void_p),
            self.row_ptr.ctypes.data if self.row_ptr is not None else None,
            self.col_ptr.ctypes.data if self.col_ptr is not None else None,
            self._z1.ctypes.data if self._z1 is not None else None,
            self._z2.ctypes.data if self._z2 is not None else None,
        )

    @property
    def c(self):
        return C.byref(self._c)


def operator_sum(A, B):
    """add_operators, linear_operator_sums.f90:38-72."""
    if A.nrow != B.nrow or A.ncol != B.ncol:
        raise ValueError("Dimensions of operators to be summed are not consistent")
    return Operator(SUM, A.nrow, A.ncol, [A, B])


def operator_product(A, B):
    """multiply_operators, linear_operator_products.f90:39-73."""
    if A.ncol != B.nrow:
        raise ValueError("Dimensions of operators to be multiplied are inconsistent")
    return Operator(PRODUCT, A.nrow, B.ncol, [A, B])


def adjoint(A):
    """adjoint, linear_operator_adjoints.f90:28-44."""
    return Operator(ADJOINT, A.ncol, A.nrow, [A])


def composite(rows, cols, blocks):
    """sparse_matrix with set_block_sizes(rows, cols) (sparse_matrix_composites.f90:226-262)
    and set_submatrix(it, jt, blocks[it][jt]) (:1031-1065)."""
    row_ptr = np.concatenate([[1], 1 + np.cumsum(rows)]).astype(np.int32)
    col_ptr = np.concatenate([[1], 1 + np.cumsum(cols)]).astype(np.int32)
    kids = [blocks[it][jt] for it in range(len(rows)) for jt in range(len(cols))]
    for it in range(len(rows)):
        for jt in range(len(cols)):
            b = blocks[it][jt]
            if b.nrow != rows[it] or b.ncol != cols[jt]:
                raise ValueError("Inconsistent dimensions for sub-matrix")
    return Operator(COMPOSITE, int(sum(rows)), int(sum(cols)), kids, row_ptr, col_ptr)


def get_value(A, i, j):
    """orc_get_value: A%get_value(i, j), 1-based."""
    return float(lib().orc_get_value(A.c, int(i), int(j)))


def ll_graph_edges(n, ei, ej):
    """orc_ll_graph_edges: replay add_edge calls, return iteration-order edges."""
    ei, ej = _i32(ei), _i32(ej)
    oi, oj = np.empty_like(ei), np.empty_like(ej)
    md = C.c_int32(0)
    ne = lib().orc_ll_graph_edges(n, ei.size, ei, ej, oi, oj, C.byref(md))
    return oi[:ne].copy(), oj[:ne].copy(), md.value


def cs_graph_build(n, src_i, src_j, trans=False):
    """orc_cs_graph_build -> (ptr, node, max_d), 1-based."""
    src_i, src_j = _i32(src_i), _i32(src_j)
    ptr = np.empty(n + 1, np.int32)
    node = np.empty(src_i.size, np.int32)
    md = lib().orc_cs_graph_build(n, src_i.size, src_i, src_j, int(trans), ptr, node)
    if md < 0:
        raise ValueError("edge stream held duplicates")
    return ptr, node, md


def ellpack_graph_build(n, src_i, src_j, trans=False):
    """orc_ellpack_max_degree + orc_ellpack_graph_build -> (node[n,max_d], degrees)."""
    src_i, src_j = _i32(src_i), _i32(src_j)
    md = lib().orc_ellpack_max_degree(n, src_i.size, src_i, src_j, int(trans))
    node = np.empty((n, md), np.int32)
    deg = np.empty(n, np.int32)
    lib().orc_ellpack_graph_build(n, src_i.size, src_i, src_j, int(trans), md, node.reshape(-1), deg)
    return node, deg


def cs_set_value(ptr, node, val, i, j, z, add=False):
    return lib().orc_cs_set_value(ptr, node, val, i, j, z, int(add))


def ell_set_value(node, degrees, val, i, j, z, add=False):
    return lib().orc_ell_set_value(node.shape[1], node.reshape(-1), degrees, val.reshape(-1), i, j, z, int(add))


def matrix_entries(A: Matrix):
    """The (i, j, value) stream of A's entry iterator, in iteration order:
    cs_matrix_get_entries (cs_matrices.f90:415-438; a csc_matrix swaps the pair,
    :793-806) walks the stored arrays line by line; ellpack_matrix_get_entries
    (ellpack_matrices.f90:381-434) walks each row's first degrees(i) slots."""
    if A.format == ELL:
        k = np.arange(A.max_d)[None, :] < A.degrees[:, None]
        i = np.nonzero(k)[0].astype(np.int32) + 1
        return i, A.node[k].astype(np.int32), A.val.reshape(A.nrow, A.max_d)[k]
    line = np.repeat(np.arange(1, A.ptr.size, dtype=np.int32), np.diff(A.ptr))
    return (line, A.node.copy(), A.val.copy()) if A.format == CSR else (A.node.copy(), line, A.val.copy())


def copy_matrix(B: Matrix, fmt, trans=False) -> Matrix:
    """A%copy_matrix(B, trans) for a target A of format `fmt`:
    cs_matrix_copy_matrix (cs_matrices.f90:294-322) / ellpack_matrix_copy_matrix
    (ellpack_matrices.f90:169-198) = build_graph_from_matrix + copy_matrix_values
    (sparse_matrix_interfaces.f90:692-772)."""
    si, sj, sv = matrix_entries(B)
    nrow, ncol = (B.ncol, B.nrow) if trans else (B.nrow, B.ncol)
    if fmt == ELL:
        node, deg = ellpack_graph_build(nrow, si, sj, trans=trans)
        val = np.zeros(node.shape)
        miss = lib().orc_copy_matrix_values(ELL, node.shape[1], None, node.reshape(-1), deg.ctypes.data,
                                            val.reshape(-1), si.size, _i32(si), _i32(sj), _f64(sv), int(trans))
        assert miss == 0
        return Matrix(ELL, nrow, ncol, node, val, degrees=deg)
    # the graph of a csc_matrix holds the columns: built with trans flipped (:311)
    t = (fmt == CSC) != bool(trans)
    nlines = ncol if fmt == CSC else nrow
    ptr, node, _ = cs_graph_build(nlines, si, sj, trans=t)
    val = np.zeros(node.size)
    miss = lib().orc_copy_matrix_values(fmt, 0, ptr.ctypes.data, node, None, val, si.size, _i32(si), _i32(sj),
                                        _f64(sv), int(trans))
    assert miss == 0
    return Matrix(fmt, nrow, ncol, node, val, ptr=ptr)


def add_values(A: Matrix, ci, cj, cz):
    """orc_add_values: apply `call A%add_value(ci[c], cj[c], cz[c])` for c = 0, 1, ... in
    order, in place on A.val; returns the number of calls that missed the pattern."""
    if A.format == ELL:
        return int(lib().orc_add_values(ELL, A.max_d, None, A.node.reshape(-1), A.degrees.ctypes.data,
                                        A.val.reshape(-1), len(ci), _i32(ci), _i32(cj), _f64(cz)))
    return int(lib().orc_add_values(A.format, 0, A.ptr.ctypes.data, A.node, None, A.val, len(ci), _i32(ci),
                                    _i32(cj), _f64(cz)))


class Ldu:
    """sparse_ldu_solver after setup (ldu_solvers.f90:35-58): strict-triangular csr factors
    L, U (unit diagonals implied) and the diagonal D."""

    def __init__(self, n, Lptr, Lnode, Lval, Uptr, Unode, Uval, D):
        self.n = n
        self.Lptr, self.Lnode, self.Lval = Lptr, Lnode, Lval
        self.Uptr, self.Unode, self.Uval = Uptr, Unode, Uval
        self.D = D

    @property
    def args(self):
        return (self.Lptr, self.Lnode, self.Lval, self.Uptr, self.Unode, self.Uval, self.D)


def ldu_setup(A: Matrix) -> Ldu:
    """pc => ldu(); call pc%setup(A) (sparse_ldu_setup ldu_solvers.f90:95-130):
    incomplete_ldu_sparsity_pattern (:396-441: ll_graph add_edge in A's iteration order,
    then L%init / U%init copy the ll_graphs to cs graphs) and the numeric factorisation
    (:275-387)."""
    n = A.nrow
    ai, aj, av = matrix_entries(A)
    lo, up = ai > aj, aj > ai

    def pattern(mask):
        si, sj, _ = ll_graph_edges(n, ai[mask], aj[mask])
        ptr, node, _ = cs_graph_build(n, si, sj)
        return ptr, node

    Lptr, Lnode = pattern(lo)
    Uptr, Unode = pattern(up)
    Lval, Uval, D = np.zeros(Lnode.size), np.zeros(Unode.size), np.zeros(n)
    lib().orc_ldu_factor(n, ai.size, _i32(ai), _i32(aj), _f64(av), Lptr, Lnode, Lval, Uptr, Unode, Uval, D)
    return Ldu(n, Lptr, Lnode, Lval, Uptr, Unode, Uval, D)


def ldu_solve(F: Ldu, b):
    """call pc%solve(A, x, b) (ldu_solve ldu_solvers.f90:160-176)."""
    x = np.empty(F.n)
    lib().orc_ldu_solve(F.n, *F.args, x, _f64(b))
    return x


def cg_solve_ldu(A, x0, b, F: Ldu, tol=1e-16, max_iter=-1):
    """orc_cg_solve_ldu: call solver%solve(A, x, b, pc) with pc = ldu() -> (x, iterations, res2, capped)."""
    x = _f64(x0).copy()
    work = np.zeros(4 * A.nrow)
    res2, capped = C.c_double(0.0), C.c_int32(0)
    it = lib().orc_cg_solve_ldu(A.c, x, _f64(b), *F.args, tol, max_iter, work, C.byref(res2), C.byref(capped))
    return x, int(it), res2.value, bool(capped.value)


def bicgstab_solve_ldu(A, x0, b, F: Ldu, tol=1e-16, max_iter=-1):
    """orc_bicgstab_solve_ldu: bicgstab_solve_pc with pc = ldu() -> (x, iterations, res2, capped)."""
    x = _f64(x0).copy()
    work = np.zeros(8 * A.nrow)
    res2, capped = C.c_double(0.0), C.c_int32(0)
    it = lib().orc_bicgstab_solve_ldu(A.c, x, _f64(b), *F.args, tol, max_iter, work, C.byref(res2), C.byref(capped))
    return x, int(it), res2.value, bool(capped.value)


def matvec(A: Matrix, x, trans=False):
    y = np.empty(A.ncol if trans else A.nrow)
    lib().orc_matvec(A.c, int(trans), _f64(x), y)
    return y


def matvec_add(A: Matrix, x, y, trans=False):
    y = _f64(y).copy()
    lib().orc_matvec_add(A.c, int(trans), _f64(x), y)
    return y


def jacobi_setup(A: Matrix):
    idiag = np.empty(A.nrow)
    lib().orc_jacobi_setup(A.c, idiag)
    return idiag


def jacobi_solve(idiag, b):
    x = np.empty_like(idiag)
    lib().orc_jacobi_solve(idiag.size, idiag, x, _f64(b))
    return x


def _solve(fn, A, x0, b, tol, max_iter, nwork, idiag=None):
    x = _f64(x0).copy()
    work = np.zeros(nwork * A.nrow)
    res2, capped = C.c_double(0.0), C.c_int32(0)
    if idiag is None:
        it = fn(A.c, x, _f64(b), tol, max_iter, work, C.byref(res2), C.byref(capped))
    else:
        it = fn(A.c, x, _f64(b), _f64(idiag), tol, max_iter, work, C.byref(res2), C.byref(capped))
    return x, int(it), res2.value, bool(capped.value)


def cg_solve(A, x0, b, tol=1e-16, max_iter=-1, idiag=None):
    """orc_cg_solve / orc_cg_solve_jacobi -> (x, iterations, res2, capped)."""
    L = lib()
    return _solve(L.orc_cg_solve if idiag is None else L.orc_cg_solve_jacobi, A, x0, b, tol, max_iter, 4, idiag)


def bicgstab_solve(A, x0, b, tol=1e-16, max_iter=-1, idiag=None):
    """orc_bicgstab_solve / orc_bicgstab_solve_jacobi -> (x, iterations, res2, capped)."""
    L = lib()
    return _solve(L.orc_bicgstab_solve if idiag is None else L.orc_bicgstab_solve_jacobi, A, x0, b, tol, max_iter, 8, idiag)


def lanczos(A, n, q1):
    """orc_lanczos -> (T[3,n] as in Fortran T(3,n), Q[nrow,n])."""
    T = np.empty(3 * n)
    Q = np.empty(A.nrow * n)
    w = np.empty(A.nrow)
    lib().orc_lanczos(A.c, n, _f64(q1), T, Q, w)
    return T.reshape(n, 3).T.copy(), Q.reshape(n, A.nrow).T.copy()


def generalized_lanczos(A, B, n, q1, cg_tol=1e-15, cg_max_iter=-1):
    """orc_generalized_lanczos -> (T[3,n], Q[nrow,n], inner CG iterations)."""
    T = np.empty(3 * n)
    Q = np.empty(A.nrow * n)
    inner = lib().orc_generalized_lanczos(A.c, B.c, n, _f64(q1), cg_tol, cg_max_iter, T, Q)
    return T.reshape(n, 3).T.copy(), Q.reshape(n, A.nrow).T.copy(), int(inner)


def tridiag_eig(d, e):
    d = _f64(d).copy()
    n = d.size
    ee = np.zeros(n)
    ee[: n - 1] = e
    Z = np.empty(n * n)
    info = lib().orc_tridiag_eig(n, d, ee, Z)
    return info, d, Z.reshape(n, n).T.copy()


def eigensolve(A, n, q1):
    lam = np.empty(n)
    V = np.empty(A.nrow * n)
    info = lib().orc_eigensolve(A.c, n, _f64(q1), lam, V)
    return info, lam, V.reshape(n, A.nrow).T.copy()


def partition_rows(ptr, P):
    part = np.empty(P + 1, np.int32)
    lib().orc_partition_rows(ptr.size - 1, _i32(ptr), P, part)
    return part


def halo_build(lo, hi, ptr, node):
    ptr, node = _i32(ptr), _i32(node)
    cnt = int(ptr[hi] - ptr[lo])
    halo = np.empty(max(cnt, 1), np.int32)
    local = np.empty(max(cnt, 1), np.int32)
    nh = lib().orc_halo_build(lo, hi, ptr, node, halo, local)
    return halo[:nh].copy(), local[:cnt].copy()
