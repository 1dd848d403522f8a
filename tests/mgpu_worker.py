"""Worker of tests/test_gpu_mgpu.py: the single-process multi-GPU mode of the C-ABI (sigb_mgpu_*),
one Python process, `ndev` GPUs.  Whole host arrays in, the library shards; everything is checked
against the oracle evaluated serially on the whole matrix."""
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle as orc  # noqa: E402
import sigma_b200 as sb  # noqa: E402
from sigma_b200 import _capi  # noqa: E402
from sigma_b200 import generators as G  # noqa: E402


def within(it, ref, frac=0.02):
    return abs(it - ref) <= max(1, int(np.ceil(frac * ref)))


def main():
    ndev = int(sys.argv[1])
    orc.build()
    got = sb.mgpu_init(ndev)
    assert got == ndev, (got, ndev)
    rng = np.random.default_rng(0)
    cases = [("poisson", 200 * 200, *G.poisson2d_csr(200)),
             ("er", 6000, *G.erdos_renyi_csr(6000, seed=3, weights="random", skew=True)),
             ("fem", 41 * 41, *G.fem_p1_csr(41)),
             ("tiny", 5, np.arange(1, 7, dtype=np.int32), np.arange(1, 6, dtype=np.int32), np.arange(1.0, 6.0))]
    for name, n, ptr, node, val in cases:
        A = sb.mgpu_csr_matrix(n, ptr, node, val)
        assert (A.nrow, A.ncol, A.nnz) == (n, n, node.size), name
        O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
        x, y0 = rng.standard_normal(n), rng.standard_normal(n)
        for _ in range(3):                          # three in a row: both halo landing buffers and their reuse
            assert np.array_equal(A.matvec(x), orc.matvec(O, x)), name
            x = np.cos(x)
        assert np.array_equal(A.matvec_add(x, y0), orc.matvec_add(O, x, y0)), name
        # new values on the same pattern (dirty-mirror refresh of the host mirror)
        A.set_values(2.0 * val)
        assert np.array_equal(A.matvec(x), orc.matvec(orc.Matrix(orc.CSR, n, n, node, 2.0 * val, ptr=ptr), x)), name
        try:
            A.matvec_t(x)
            raise SystemExit("matvec_t of a multi-GPU operator was not refused")
        except sb.SigmaError as e:
            assert e.status == _capi.ERR_UNSUPPORTED
        A.destroy()

    # ---- CG on Poisson: iterations within 2 %, solution within 1e-10; both loop forms ----------
    N = 160
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    b, _ = G.poisson2d_rhs(N)
    tol = 1e-10 * np.linalg.norm(b)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    xo, ito, _, _ = orc.cg_solve(O, np.zeros(n), b, tol)
    A = sb.mgpu_csr_matrix(n, ptr, node, val)
    for form in (1, 0):
        s = sb.cg(tol)
        s.set_persistent(form)
        s.setup(A)
        x = s.solve(A, np.zeros(n), b)
        it, res2, capped = s.info()
        assert not capped and within(it, ito), (form, it, ito)
        assert np.linalg.norm(x - xo) / np.linalg.norm(xo) <= 1e-10, form
        # a second solve accumulates the iteration counter (cg_solvers.f90:72,145)
        s.solve(A, np.zeros(n), b)
        assert s.info()[0] == 2 * it
        s.destroy()
    # a capped solve stops at the cap on every GPU
    s = sb.cg(tol)
    s.set_max_iterations(25)
    s.setup(A)
    x25 = s.solve(A, np.zeros(n), b)
    assert s.info()[0] == 25 and s.info()[2]
    x25o, _, _, _ = orc.cg_solve(O, np.zeros(n), b, tol, 25)
    assert np.abs(x25 - x25o).max() <= 1e-10 * np.abs(x25o).max()
    A.destroy()

    # ---- Jacobi-PCG and BiCGSTAB(+Jacobi) on the Erdos-Renyi operators ---------------------------
    n = 5000
    ptr, node, val = G.erdos_renyi_csr(n, seed=8, weights="random")
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    v = np.random.default_rng(9).random(n)
    f = orc.matvec(O, v)
    A = sb.mgpu_csr_matrix(n, ptr, node, val)
    s, pc = sb.cg(1e-13), sb.jacobi()
    s.setup(A)
    pc.setup(A)
    idiag = orc.jacobi_setup(O)
    assert np.array_equal(pc.vector("idiag"), idiag)
    u = s.solve(A, np.zeros(n), f, pc)
    uo, ito, _, _ = orc.cg_solve(O, np.zeros(n), f, 1e-13, idiag=idiag)
    assert within(s.iterations, ito), (s.iterations, ito)
    assert np.linalg.norm(u - uo) / np.linalg.norm(uo) <= 1e-10
    A.destroy()
    ptr, node, val = G.erdos_renyi_csr(n, seed=12, weights="random", skew=True)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    f = orc.matvec(O, v)
    A = sb.mgpu_csr_matrix(n, ptr, node, val)
    for use_pc in (False, True):
        s = sb.bicgstab(1e-13)
        s.setup(A)
        s.set_max_iterations(10 * n)
        pc = None
        if use_pc:
            pc = sb.jacobi()
            pc.setup(A)
        u = s.solve(A, np.zeros(n), f, pc)
        uo, ito, _, _ = orc.bicgstab_solve(O, np.zeros(n), f, 1e-13, idiag=orc.jacobi_setup(O) if use_pc else None)
        assert not s.info()[2] and within(s.iterations, ito, 0.05), (s.iterations, ito)
        assert np.linalg.norm(u - uo) / np.linalg.norm(uo) <= 1e-10
    A.destroy()

    # ---- Lanczos / eigensolve: whole vectors in and out, every GPU runs its row block ----------------
    n, nq = 4096, 12
    ptr, node, val = G.erdos_renyi_csr(n, seed=41, shift=0.0)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    q1 = 2 * np.random.default_rng(3).random(n) - 1
    A = sb.mgpu_csr_matrix(n, ptr, node, val)
    T, V = sb.lanczos(A, nq, q1)
    To, Vo = orc.lanczos(O, nq, q1)
    assert np.allclose(T, To, rtol=1e-9, atol=1e-11)
    assert np.sqrt(((V.T @ V - np.eye(nq)) ** 2).sum()) / nq <= 1e-14
    lam, W = sb.eigensolve(A, nq, q1)
    info, lamo, Wo = orc.eigensolve(O, nq, q1)
    assert info == 0 and np.allclose(lam, lamo, rtol=1e-9, atol=1e-10)
    assert np.all(W[0, :] > 0)
    for j in (0, nq - 1):
        assert np.allclose(W[:, j], Wo[:, j], atol=1e-7), j
    A.destroy()

    # bad input is refused before anything is sharded
    try:
        sb.mgpu_csr_matrix(3, [1, 2, 3, 4], [1, 2, 9], [1.0, 1.0, 1.0])
        raise SystemExit("a column id outside 1..n was accepted")
    except sb.SigmaError as e:
        assert e.status == _capi.ERR_ARG
    print(f"mgpu ok ({ndev} GPU(s), launches {sb.launch_count()})")
    sb.mgpu_finalize()


if __name__ == "__main__":
    main()
