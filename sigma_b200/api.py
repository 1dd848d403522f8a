"""Host-side mirror of the reference's operator / solver interface for the hot
path, as a thin layer over the C-ABI (include/sigma_b200.h).

Names and call shapes follow the Fortran API so the parity tests read like the
reference's own test programs:

    reference (Fortran)                              here
    ----------------------------------------------   -----------------------------------
    call A%matvec(x, y) / matvec_t / matvec_add      y = A.matvec(x) / A.matvec_t(x) / A.matvec_add(x, y)
    solver => cg(1.d-16); call solver%setup(A)       solver = cg(1e-16); solver.setup(A)
    call solver%solve(A, u, f [, pc])                u = solver.solve(A, u, f, pc)
    pc => jacobi(); call pc%setup(A)                 pc = jacobi(); pc.setup(A)
    call lanczos(A, T, Q)                            T, Q = lanczos(A, n, q1)
    call eigensolve(A, lambda, V)                    lam, V = eigensolve(A, n, q1)
    L = A + B ; L = A * B ; L = adjoint(A)           L = A + B ; L = A * B ; L = adjoint(A)
    type(sparse_matrix): set_block_sizes(rows, cols) S = sparse_matrix(rows, cols, blocks)
      + set_submatrix(it, jt, C)

(linear_operator_interface.f90:185-233, cg_solvers.f90:36-194,
bicgstab_solvers.f90:36-237, jacobi_solvers.f90:26-81, eigensolver.f90:27-184,
linear_operator_sums.f90:38-72, linear_operator_products.f90:39-73,
linear_operator_adjoints.f90:28-44, sparse_matrix_composites.f90:226-262,1031-1129.)
Errors raise SigmaError where the reference prints and calls exit(1).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import ROW, COL, SigmaError, as_f64, as_i32, check, lib, ptr

__all__ = ["init", "Graph", "Matrix", "Solver", "cg", "bicgstab", "jacobi", "ldu", "ldu_symbolic", "lanczos", "lanczos_dev", "eigensolve",
           "generalized_lanczos", "generalized_eigensolve",
           "csr_matrix", "csc_matrix", "ellpack_matrix", "SigmaError", "launch_count", "set_stream",
           "synchronize", "Expression", "operator_sum", "operator_product", "adjoint", "sparse_matrix",
           "multiple_values_stream", "mgpu_init", "mgpu_finalize", "mgpu_csr_matrix"]


def init(device: int = -1):
    check(lib().sigb_init(device))


def launch_count() -> int:
    return int(lib().sigb_launch_count())


def set_stream(stream_ptr):
    check(lib().sigb_set_stream(stream_ptr))


def synchronize():
    check(lib().sigb_synchronize())


class Graph:
    """Device mirror of a cs_graph / ellpack_graph (refcounted like the reference)."""

    def __init__(self, handle, kind, n, m):
        self._h, self.kind, self.n, self.m = handle, kind, n, m

    @classmethod
    def cs(cls, n, m, ptr1, node1, order=ROW):
        ptr1, node1 = as_i32(ptr1), as_i32(node1)
        if ptr1.size != n + 1:
            raise SigmaError(_capi.ERR_ARG, "ptr must have n+1 entries")
        h = C.c_void_p()
        check(lib().sigb_cs_graph_create(n, m, ptr(ptr1), ptr(node1), order, C.byref(h)))
        return cls(h, "csr" if order == ROW else "csc", n, m)

    @classmethod
    def ellpack(cls, n, m, node, degrees):
        """node: (n, max_d) C-order == Fortran node(max_d, n)."""
        node, degrees = as_i32(node), as_i32(degrees)
        h = C.c_void_p()
        check(lib().sigb_ell_graph_create(n, m, node.shape[1], ptr(node), ptr(degrees), C.byref(h)))
        g = cls(h, "ellpack", n, m)
        g.max_d = node.shape[1]
        return g

    @classmethod
    def build(cls, frmt, n, m, src_i, src_j, trans=False):
        """g%build(n, m, get_edges, make_cursor, trans) on the device (cs_graphs.f90:109-197,
        ellpack_graphs.f90:105-170): the pattern from the source graph's edge stream (src_i, src_j) in
        its iteration order; frmt "csr" | "csc" | "ellpack"."""
        src_i, src_j = as_i32(src_i), as_i32(src_j)
        if src_i.size != src_j.size:
            raise SigmaError(_capi.ERR_ARG, "edge stream: src_i and src_j differ in length")
        h = C.c_void_p()
        if frmt == "ellpack":
            check(lib().sigb_ell_graph_build(n, m, src_i.size, ptr(src_i), ptr(src_j), int(trans), C.byref(h)))
        else:
            check(lib().sigb_cs_graph_build(n, m, src_i.size, ptr(src_i), ptr(src_j), int(trans),
                                            ROW if frmt == "csr" else COL, C.byref(h)))
        return cls(h, frmt, n, m)

    def transpose_arrays(self, other_dim, ne):
        ptr_t = np.empty(other_dim + 1, np.int32)
        node_t = np.empty(ne, np.int32)
        check(lib().sigb_cs_graph_get_transpose(self._h, ptr(ptr_t), ptr(node_t)))
        return ptr_t, node_t

    def release(self):
        if self._h:
            check(lib().sigb_graph_release(self._h))
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Matrix:
    """Device mirror of csr_matrix / csc_matrix / ellpack_matrix (a linear_operator)."""

    def __init__(self, graph: "Graph | None", handle=None):
        self.g = graph
        if handle is None:
            handle = C.c_void_p()
            check(lib().sigb_matrix_create(graph._h, C.byref(handle)))
        self._h = handle
        nrow, ncol, nnz = C.c_int32(), C.c_int32(), C.c_int64()
        check(lib().sigb_matrix_get_dims(self._h, C.byref(nrow), C.byref(ncol), C.byref(nnz)))
        self.nrow, self.ncol, self.nnz = nrow.value, ncol.value, nnz.value

    def set_values(self, val):
        val = as_f64(val)
        check(lib().sigb_matrix_set_values(self._h, ptr(val), val.size))
        return self

    # -- linear_operator interface ------------------------------------------
    def matvec(self, x, trans=False):
        x = as_f64(x)
        y = np.empty(self.ncol if trans else self.nrow)
        check(lib().sigb_matvec(self._h, int(trans), ptr(x), ptr(y)))
        return y

    def matvec_t(self, x):
        return self.matvec(x, trans=True)

    def matvec_add(self, x, y, trans=False):
        x, y = as_f64(x), as_f64(y).copy()
        check(lib().sigb_matvec_add(self._h, int(trans), ptr(x), ptr(y)))
        return y

    def matvec_t_add(self, x, y):
        return self.matvec_add(x, y, trans=True)

    # -- device-resident variants (torch tensors or raw addresses) ----------
    def matvec_dev(self, x_dev, y_dev, trans=False, add=False):
        check(lib().sigb_matvec_dev(self._h, int(trans), ptr(x_dev), ptr(y_dev), int(add)))

    def matvec_dot_dev(self, x_dev, y_dev, fetch=True):
        if not fetch:
            check(lib().sigb_matvec_dot_dev(self._h, ptr(x_dev), ptr(y_dev), None))
            return None
        d = C.c_double()
        check(lib().sigb_matvec_dot_dev(self._h, ptr(x_dev), ptr(y_dev), C.byref(d)))
        return d.value

    def transpose_values(self):
        out = np.empty(self.g_ne_transposed())
        check(lib().sigb_matrix_get_transpose_values(self._h, ptr(out)))
        return out

    def g_ne_transposed(self):
        if self.g.kind == "ellpack":
            return self.g.n * self.g.max_d
        return self.nnz

    # -- copy / format conversion on the device ------------------------------
    def copy_matrix(self, frmt: str, trans: bool = False) -> "Matrix":
        """B of format `frmt` ("csr" | "csc" | "ellpack") after `call B%copy_matrix(A, trans)`
        (cs_matrices.f90:294-322, ellpack_matrices.f90:169-198), built on the device."""
        code = {"csr": _capi.FMT_CSR, "csc": _capi.FMT_CSC, "ellpack": _capi.FMT_ELLPACK}[frmt]
        h = C.c_void_p()
        check(lib().sigb_matrix_copy(self._h, code, int(trans), C.byref(h)))
        B = Matrix(None, h)
        B.format = frmt
        return B

    def add_values(self, i1, j1, z):
        """`call A%add_value(i1[c], j1[c], z[c])` for c = 0, 1, ... in order, on the device
        (cs_matrices.f90:868-891,924-947, ellpack_matrices.f90:471-493)."""
        i1, j1, z = as_i32(i1), as_i32(j1), as_f64(z)
        if not (i1.size == j1.size == z.size):
            raise SigmaError(_capi.ERR_ARG, "add_values: i, j, z differ in length")
        check(lib().sigb_matrix_add_values(self._h, i1.size, ptr(i1), ptr(j1), ptr(z)))
        return self

    def add_multiple_values(self, is1, js1, B):
        """`call A%add_multiple_values(is, js, B)` (cs_matrices.f90:934-967, ellpack likewise): the
        add_value stream (is(k), js(l), B(k, l)) with k outer, l inner.  is1 / js1 / B may carry a
        leading batch axis -- (nb, ni), (nb, nj), (nb, ni, nj) -- for an assembly loop that issues
        one small block per element: the blocks are applied in batch order by ONE device call."""
        I, J, Z = multiple_values_stream(is1, js1, B)
        return self.add_values(I, J, Z)

    def arrays(self):
        """The stored arrays, read back from the device exactly as the Fortran holds them:
        ("csr"|"csc", ptr, node, val) or ("ellpack", degrees, node[n, max_d], val[n, max_d])."""
        fmt, n, m, ne, md = C.c_int(), C.c_int32(), C.c_int32(), C.c_int64(), C.c_int32()
        check(lib().sigb_matrix_get_format(self._h, C.byref(fmt), C.byref(n), C.byref(m), C.byref(ne), C.byref(md)))
        if fmt.value == _capi.FMT_ELLPACK:
            deg = np.empty(n.value, np.int32)
            node = np.empty((n.value, md.value), np.int32)
            val = np.empty((n.value, md.value))
            check(lib().sigb_matrix_get_arrays(self._h, ptr(deg), ptr(node), ptr(val)))
            return "ellpack", deg, node, val
        p = np.empty(n.value + 1, np.int32)
        node = np.empty(ne.value, np.int32)
        val = np.empty(ne.value)
        check(lib().sigb_matrix_get_arrays(self._h, ptr(p), ptr(node), ptr(val)))
        return ("csr" if fmt.value == _capi.FMT_CSR else "csc"), p, node, val

    # -- operator algebra: interface operator(+) / operator(*) ----------------
    def __add__(self, other):
        return operator_sum(self, other)

    def __mul__(self, other):
        return operator_product(self, other)

    def destroy(self):
        if self._h:
            check(lib().sigb_matrix_destroy(self._h))
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Expression(Matrix):
    """A lazy operator expression (operator_sum / operator_product / operator_adjoint) or a
    block composite sparse_matrix.  It is a linear_operator like any matrix: matvec, the
    solvers and the Lanczos loops take it unchanged; the device library holds references on
    its operands."""

    def __init__(self, handle, kind, operands):
        super().__init__(None, handle)
        self.kind, self.operands = kind, list(operands)

    def set_values(self, val):
        raise SigmaError(_capi.ERR_UNSUPPORTED, "an operator expression has no values of its own")


def operator_sum(A: Matrix, B: Matrix) -> Expression:
    """L = A + B (add_operators, linear_operator_sums.f90:38-72)."""
    h = C.c_void_p()
    check(lib().sigb_operator_sum(A._h, B._h, C.byref(h)))
    return Expression(h, "sum", [A, B])


def operator_product(A: Matrix, B: Matrix) -> Expression:
    """L = A * B (multiply_operators, linear_operator_products.f90:39-73)."""
    h = C.c_void_p()
    check(lib().sigb_operator_product(A._h, B._h, C.byref(h)))
    return Expression(h, "product", [A, B])


def adjoint(A: Matrix) -> Expression:
    """L = adjoint(A) (linear_operator_adjoints.f90:28-44)."""
    h = C.c_void_p()
    check(lib().sigb_operator_adjoint(A._h, C.byref(h)))
    return Expression(h, "adjoint", [A])


def sparse_matrix(rows, cols, blocks) -> Expression:
    """type(sparse_matrix) composite: A%set_block_sizes(rows, cols) then
    A%set_submatrix(it, jt, blocks[it][jt]) (sparse_matrix_composites.f90:226-262,1031-1065)."""
    rows, cols = as_i32(rows), as_i32(cols)
    flat = [blocks[it][jt] for it in range(rows.size) for jt in range(cols.size)]
    arr = (C.c_void_p * len(flat))(*[b._h for b in flat])
    h = C.c_void_p()
    check(lib().sigb_composite_create(rows.size, cols.size, ptr(rows), ptr(cols), arr, C.byref(h)))
    return Expression(h, "composite", flat)


def multiple_values_stream(is1, js1, B):
    """The add_value calls `A%add_multiple_values(is, js, B)` makes, in its own loop order (k over
    is outer, l over js inner, cs_matrices.f90:944-965); with a leading batch axis, block after
    block.  Returns (i, j, z) ready for add_values / the oracle."""
    is1, js1, B = np.asarray(is1, np.int32), np.asarray(js1, np.int32), np.asarray(B, np.float64)
    if is1.ndim == 1:
        is1, js1, B = is1[None], js1[None], B[None]
    nb, ni = is1.shape
    nj = js1.shape[1]
    if js1.shape[0] != nb or B.shape != (nb, ni, nj):
        raise SigmaError(_capi.ERR_ARG, "add_multiple_values: B must be size(is) x size(js)")
    I = np.repeat(is1[:, :, None], nj, axis=2).reshape(-1)
    J = np.repeat(js1[:, None, :], ni, axis=1).reshape(-1)
    return I, J, B.reshape(-1)


def csr_matrix(n, m, ptr1, node1, val):
    """type(csr_matrix): A%init(n, m); A%set_graph(g); values as stored."""
    return Matrix(Graph.cs(n, m, ptr1, node1, ROW)).set_values(val)


def mgpu_init(ndev: int = 0) -> int:
    """Single-process multi-GPU mode (sigb_mgpu_init): one worker thread per GPU inside the library.
    Returns the number of GPUs in use."""
    check(lib().sigb_mgpu_init(int(ndev)))
    n = C.c_int()
    check(lib().sigb_mgpu_device_count(C.byref(n)))
    return n.value


def mgpu_finalize():
    check(lib().sigb_mgpu_finalize())


def mgpu_csr_matrix(n, ptr1, node1, val):
    """type(csr_matrix) spread over the GPUs of mgpu_init: the whole pattern in, row blocks / halo /
    send lists derived in the library; matvec and the solvers take whole host vectors."""
    ptr1, node1 = as_i32(ptr1), as_i32(node1)
    if ptr1.size != n + 1:
        raise SigmaError(_capi.ERR_ARG, "ptr must have n+1 entries")
    h = C.c_void_p()
    check(lib().sigb_mgpu_csr_create(n, ptr(ptr1), ptr(node1), C.byref(h)))
    return Matrix(None, handle=h).set_values(val)


def csc_matrix(nrow, ncol, ptr1, node1, val):
    """type(csc_matrix): ptr over the ncol columns, node = row ids."""
    return Matrix(Graph.cs(ncol, nrow, ptr1, node1, COL)).set_values(val)


def ellpack_matrix(n, m, node, degrees, val):
    """type(ellpack_matrix): node/val (n, max_d) C-order == Fortran (max_d, n)."""
    return Matrix(Graph.ellpack(n, m, node, degrees)).set_values(np.asarray(val).reshape(-1))


class Solver:
    """linear_solver: cg_solver / bicgstab_solver / jacobi_solver."""

    def __init__(self, kind, tolerance=None):
        h = C.c_void_p()
        tol = -1.0 if tolerance is None else float(tolerance)
        if kind == "cg":
            check(lib().sigb_cg_create(tol, C.byref(h)))
        elif kind == "bicgstab":
            check(lib().sigb_bicgstab_create(tol, C.byref(h)))
        elif kind == "jacobi":
            check(lib().sigb_jacobi_create(C.byref(h)))
        elif kind == "ldu":
            check(lib().sigb_ldu_create(C.byref(h)))
        else:
            raise ValueError(kind)
        self.kind, self._h = kind, h

    def setup(self, A: Matrix):
        check(lib().sigb_solver_setup(self._h, A._h))
        self.nn = A.nrow
        return self

    def set_params(self, tolerance=None):
        check(lib().sigb_solver_set_params(self._h, -1.0 if tolerance is None else float(tolerance)))

    def set_max_iterations(self, cap: int):
        check(lib().sigb_solver_set_max_iterations(self._h, int(cap)))

    def set_persistent(self, mode: int):
        """CG loop form: 1 one persistent kernel, 0 three kernels per iteration, -1 the library's choice."""
        check(lib().sigb_solver_set_persistent(self._h, int(mode)))

    def set_strict_order(self, on: bool = True):
        """Parity aid: dot products summed strictly left to right (one thread), so that the solve equals the
        serial reference loops bit for bit."""
        check(lib().sigb_solver_set_strict_order(self._h, int(bool(on))))

    def solve(self, A: Matrix, x, b, pc: "Solver | None" = None):
        """call solver%solve(A, x, b [, pc]); x is the initial guess, returns the solution."""
        x, b = as_f64(x).copy(), as_f64(b)
        check(lib().sigb_solver_solve(self._h, A._h, ptr(x), ptr(b), pc._h if pc else None))
        return x

    def solve_dev(self, A: Matrix, x_dev, b_dev, pc: "Solver | None" = None):
        check(lib().sigb_solver_solve_dev(self._h, A._h, ptr(x_dev), ptr(b_dev), pc._h if pc else None))

    def info(self):
        it, r2, cp = C.c_int64(), C.c_double(), C.c_int()
        check(lib().sigb_solver_get_info(self._h, C.byref(it), C.byref(r2), C.byref(cp)))
        return it.value, r2.value, bool(cp.value)

    @property
    def iterations(self):
        return self.info()[0]

    def vector(self, name):
        out = np.empty(self.nn)
        check(lib().sigb_solver_get_vector(self._h, name.encode(), ptr(out)))
        return out

    def factors(self):
        """ldu only: (Lptr, Lnode, Lval, Uptr, Unode, Uval, D, forward levels, backward levels)."""
        n, nL, nU, nf, nb = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
        check(lib().sigb_ldu_get_sizes(self._h, C.byref(n), C.byref(nL), C.byref(nU), C.byref(nf), C.byref(nb)))
        Lptr, Uptr = np.empty(n.value + 1, np.int32), np.empty(n.value + 1, np.int32)
        Lnode, Unode = np.empty(nL.value, np.int32), np.empty(nU.value, np.int32)
        Lval, Uval, D = np.empty(nL.value), np.empty(nU.value), np.empty(n.value)
        check(lib().sigb_ldu_get_factors(self._h, ptr(Lptr), ptr(Lnode), ptr(Lval), ptr(Uptr), ptr(Unode),
                                         ptr(Uval), ptr(D)))
        return Lptr, Lnode, Lval, Uptr, Unode, Uval, D, nf.value, nb.value

    def destroy(self):
        if self._h:
            check(lib().sigb_solver_destroy(self._h))
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def cg(tolerance=None):
    return Solver("cg", tolerance)


def bicgstab(tolerance=None):
    return Solver("bicgstab", tolerance)


def jacobi():
    return Solver("jacobi")


def ldu(incomplete=True, level=0):
    """pc => ldu(incomplete, level) (ldu_solvers.f90:73-86); like the reference, always the
    incomplete factorisation of level 0 whatever is asked for (:145,151)."""
    return Solver("ldu")


def ldu_symbolic(n, ptr1, node1):
    """Host-only index work of the ldu setup (no GPU): patterns of L and U, destinations of A's
    entries in [Lval | Uval | D], forward / backward level schedules."""
    ptr1, node1 = as_i32(ptr1), as_i32(node1)
    ne = node1.size
    Lptr, Uptr = np.empty(n + 1, np.int32), np.empty(n + 1, np.int32)
    Lnode, Unode = np.empty(max(ne, 1), np.int32), np.empty(max(ne, 1), np.int32)
    dest = np.empty(max(ne, 1), np.int64)
    frows, brows = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.int32)
    flev, blev = np.empty(n + 1, np.int32), np.empty(n + 1, np.int32)
    nf, nb = C.c_int32(), C.c_int32()
    check(lib().sigb_ldu_symbolic(n, ptr(ptr1), ptr(node1), ptr(Lptr), ptr(Lnode), ptr(Uptr), ptr(Unode), ptr(dest),
                                  ptr(frows), ptr(flev), C.byref(nf), ptr(brows), ptr(blev), C.byref(nb)))
    nL, nU = int(Lptr[n]) - 1, int(Uptr[n]) - 1
    return dict(Lptr=Lptr, Lnode=Lnode[:nL].copy(), Uptr=Uptr, Unode=Unode[:nU].copy(), dest=dest[:ne].copy(),
                forward_rows=frows[:n].copy(), forward_lev=flev[: nf.value + 1].copy(),
                backward_rows=brows[:n].copy(), backward_lev=blev[: nb.value + 1].copy())


def lanczos(A: Matrix, n: int, q1=None, seed: int = 0):
    """call lanczos(A, T, Q) -> T[3, n] (T(1,:), T(2,:), T(3,:)), Q[nrow, n].
    Q (and V of the eigensolves below) comes back as the Fortran array it is: column-major,
    i.e. a transposed VIEW of the buffer the library filled -- no host-side copy."""
    T = np.empty(3 * n)
    Q = np.empty(A.nrow * n)
    q1a = as_f64(q1) if q1 is not None else None
    check(lib().sigb_lanczos(A._h, n, ptr(q1a), seed, ptr(T), ptr(Q)))
    return T.reshape(n, 3).T.copy(), Q.reshape(n, A.nrow).T


def lanczos_dev(A: Matrix, n: int, Q_dev, q1_dev=None, seed: int = 0):
    """lanczos with the basis left on the device: Q_dev is a device buffer of nrow * n doubles
    (column-major Q(nrow, n)), q1_dev an optional device start vector.  Returns T[3, n]."""
    T = np.empty(3 * n)
    check(lib().sigb_lanczos_dev(A._h, n, ptr(q1_dev), seed, ptr(T), ptr(Q_dev)))
    return T.reshape(n, 3).T.copy()


def eigensolve(A: Matrix, n: int, q1=None, seed: int = 0):
    """call eigensolve(A, lambda, V) -> lambda[n] ascending, V[nrow, n]."""
    lam = np.empty(n)
    V = np.empty(A.nrow * n)
    q1a = as_f64(q1) if q1 is not None else None
    check(lib().sigb_eigensolve(A._h, n, ptr(q1a), seed, ptr(lam), ptr(V)))
    return lam, V.reshape(n, A.nrow).T


def generalized_lanczos(A: Matrix, B: Matrix, b_solver: Solver, n: int, q1=None, seed: int = 0, b_pc: "Solver | None" = None):
    """call B%set_solver(b_solver); call generalized_lanczos(A, B, T, Q) -> T[3, n], Q[nrow, n]."""
    T = np.empty(3 * n)
    Q = np.empty(A.nrow * n)
    q1a = as_f64(q1) if q1 is not None else None
    check(lib().sigb_generalized_lanczos(A._h, B._h, b_solver._h, b_pc._h if b_pc else None, n, ptr(q1a), seed,
                                         ptr(T), ptr(Q)))
    return T.reshape(n, 3).T.copy(), Q.reshape(n, A.nrow).T


def generalized_eigensolve(A: Matrix, B: Matrix, b_solver: Solver, n: int, q1=None, seed: int = 0,
                           b_pc: "Solver | None" = None):
    """call generalized_eigensolve(A, B, lambda, V) -> lambda[n] ascending, V[nrow, n]."""
    lam = np.empty(n)
    V = np.empty(A.nrow * n)
    q1a = as_f64(q1) if q1 is not None else None
    check(lib().sigb_generalized_eigensolve(A._h, B._h, b_solver._h, b_pc._h if b_pc else None, n, ptr(q1a), seed,
                                            ptr(lam), ptr(V)))
    return lam, V.reshape(n, A.nrow).T
