"""Worker of tests/test_gpu_dist.py: one rank per GPU (torchrun, NCCL).
Row-sharded SpMV / CG / PCG / BiCGSTAB / Lanczos through the C-ABI, checked
against the oracle evaluated serially on the whole matrix."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle as orc  # noqa: E402
import sigma_b200 as sb  # noqa: E402
from sigma_b200 import distributed as D  # noqa: E402
from sigma_b200 import generators as G  # noqa: E402


def within(it, ref, frac=0.02):
    return abs(it - ref) <= max(1, int(np.ceil(frac * ref)))


def shard(comm, n, ptr, node, val):
    part = D.partition_rows(ptr, comm.nranks)
    lo, hi = int(part[comm.rank]), int(part[comm.rank + 1])
    sl = slice(ptr[lo] - 1, ptr[hi] - 1)
    A = D.dist_csr_matrix(comm, n, part, ptr[lo:hi + 1], node[sl], val[sl])
    return A, lo, hi


def gather(v_local):
    world = dist.get_world_size()
    out = [None] * world
    dist.all_gather_object(out, v_local)
    return np.concatenate(out)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sb.init(local)
    comm = D.Comm.from_torch()
    rank, world = comm.rank, comm.nranks
    rng = np.random.default_rng(0)

    cases = [("poisson", 200 * 200, *G.poisson2d_csr(200)),
             ("er", 6000, *G.erdos_renyi_csr(6000, seed=3, weights="random", skew=True)),
             ("fem", 41 * 41, *G.fem_p1_csr(41))]
    for name, n, ptr, node, val in cases:
        A, lo, hi = shard(comm, n, ptr, node, val)
        O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
        x, y0 = rng.standard_normal(n), rng.standard_normal(n)
        # halo list is bit-exact index work
        ohalo, _ = orc.halo_build(lo, hi, ptr, node)
        assert np.array_equal(A.plan.halo, ohalo), name
        # SpMV: bit-exact against the serial reference loop, twice (halo buffer reuse)
        for _ in range(2):
            y = A.matvec(x[lo:hi])
            assert np.array_equal(y, orc.matvec(O, x)[lo:hi]), name
            x = np.cos(x)
        ya = A.matvec_add(x[lo:hi], y0[lo:hi])
        assert np.array_equal(ya, orc.matvec_add(O, x, y0)[lo:hi]), name
        # fused dot: every rank gets the same all-reduced value
        xd = torch.from_numpy(x[lo:hi]).cuda()
        yd = torch.empty_like(xd)
        torch.cuda.synchronize()
        d = A.matvec_dot_dev(xd, yd)
        ref = float(x @ orc.matvec(O, x))
        assert abs(d - ref) <= 1e-12 * float(np.abs(x) @ np.abs(orc.matvec(orc.Matrix(orc.CSR, n, n, node, np.abs(val), ptr=ptr), np.abs(x)))), name
        allv = [None] * world
        dist.all_gather_object(allv, d)
        assert len(set(allv)) == 1, allv

    # ---- CG on Poisson: iterations within 2 %, solution within 1e-10 ------
    N = 128
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    b, xs = G.poisson2d_rhs(N)
    tol = 1e-10 * np.linalg.norm(b)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    xo, ito, _, _ = orc.cg_solve(O, np.zeros(n), b, tol)
    A, lo, hi = shard(comm, n, ptr, node, val)
    s = sb.cg(tol)
    s.setup(A)
    xl = s.solve(A, np.zeros(hi - lo), b[lo:hi])
    it, res2, capped = s.info()
    x = gather(xl)
    assert not capped and within(it, ito), (it, ito)
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) <= 1e-10
    its = [None] * world
    dist.all_gather_object(its, it)
    assert len(set(its)) == 1, its          # every rank stopped at the same iteration

    # ---- Jacobi-PCG and BiCGSTAB(+Jacobi) on the skewed ER operator --------
    n = 5000
    ptr, node, val = G.erdos_renyi_csr(n, seed=8, weights="random")
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    v = np.random.default_rng(9).random(n)
    f = orc.matvec(O, v)
    A, lo, hi = shard(comm, n, ptr, node, val)
    s, pc = sb.cg(1e-13), sb.jacobi()
    s.setup(A)
    pc.setup(A)
    idiag = orc.jacobi_setup(O)
    assert np.array_equal(pc.vector("idiag"), idiag[lo:hi])
    ul = s.solve(A, np.zeros(hi - lo), f[lo:hi], pc)
    uo, ito, _, _ = orc.cg_solve(O, np.zeros(n), f, 1e-13, idiag=idiag)
    assert within(s.iterations, ito), (s.iterations, ito)
    assert np.linalg.norm(gather(ul) - uo) / np.linalg.norm(uo) <= 1e-10

    ptr, node, val = G.erdos_renyi_csr(n, seed=12, weights="random", skew=True)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    f = orc.matvec(O, v)
    A, lo, hi = shard(comm, n, ptr, node, val)
    for use_pc in (False, True):
        s = sb.bicgstab(1e-13)
        s.setup(A)
        s.set_max_iterations(10 * n)
        pc = None
        if use_pc:
            pc = sb.jacobi()
            pc.setup(A)
        ul = s.solve(A, np.zeros(hi - lo), f[lo:hi], pc)
        uo, ito, _, _ = orc.bicgstab_solve(O, np.zeros(n), f, 1e-13, idiag=orc.jacobi_setup(O) if use_pc else None)
        assert not s.info()[2] and within(s.iterations, ito, 0.05), (s.iterations, ito)
        assert np.linalg.norm(gather(ul) - uo) / np.linalg.norm(uo) <= 1e-10

    # ---- Lanczos: T against the oracle with the same (sharded) start vector --
    n, nq = 4096, 12
    ptr, node, val = G.erdos_renyi_csr(n, seed=41, shift=0.0)
    O = orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr)
    q1 = 2 * np.random.default_rng(3).random(n) - 1
    A, lo, hi = shard(comm, n, ptr, node, val)
    T, Vl = sb.lanczos(A, nq, q1[lo:hi])
    To, Vo = orc.lanczos(O, nq, q1)
    assert np.allclose(T, To, rtol=1e-9, atol=1e-11)
    V = np.concatenate([None] * 0 + [gather(np.ascontiguousarray(Vl[:, j])) for j in range(nq)]).reshape(nq, n).T
    assert np.sqrt(((V.T @ V - np.eye(nq)) ** 2).sum()) / nq <= 1e-14
    # eigensolve on the sharded operator: Ritz values against the oracle, vectors with the reference's sign
    # convention V(1, i) > 0 -- the first row lives on rank 0, every rank must scale by the same signs
    lam, Wl = sb.eigensolve(A, nq, q1[lo:hi])
    info, lamo, Wo = orc.eigensolve(O, nq, q1)
    assert info == 0 and np.allclose(lam, lamo, rtol=1e-9, atol=1e-10)
    W = np.concatenate([gather(np.ascontiguousarray(Wl[:, j])) for j in range(nq)]).reshape(nq, n).T
    assert np.all(W[0, :] > 0)
    for j in (0, nq - 1):           # extremal pairs: same vector as the serial oracle, sign included
        assert np.allclose(W[:, j], Wo[:, j], atol=1e-7), j
    # library-drawn start vector does not depend on the sharding
    T1, V1l = sb.lanczos(A, nq, None, seed=5)
    V1 = gather(np.ascontiguousarray(V1l[:, 0]))
    if world > 1 and rank == 0:
        np.save("/tmp/sigb_lanczos_q1.npy", V1)
    dist.barrier()
    if rank == 0:
        print(f"dist gpu ok (world {world}, transport {comm.transport}, launches {sb.launch_count()})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
