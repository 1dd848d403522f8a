#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs on one GPU (JSON lines).

  C3  same Poisson matrix in ELLPACK vs CSR, SpMV-only bandwidth sweep
  C4' P1 FEM Laplacian on a jittered triangulation, Jacobi-preconditioned CG
  C5' Erdos-Renyi graph Laplacian: BiCGSTAB on the shifted skew-perturbed operator, Lanczos steps
(primes: reduced sizes -- the generators are host numpy and the full 50 M / 20 M
instances would spend the GPU visit generating input)
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sigma_b200 as sb  # noqa: E402
from sigma_b200 import generators as G  # noqa: E402

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def emit(**kw):
    print(json.dumps(kw), flush=True)


sb.init(0)
stream = torch.cuda.Stream()
sb.set_stream(stream.cuda_stream)

# ---- C3: ELLPACK vs CSR SpMV ------------------------------------------------
for N in (1024, 2048, 4096):
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    nnz = node.size
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    A = sb.csr_matrix(n, n, ptr, node, val)
    for _ in range(5):
        A.matvec_dev(x, y)
    t_csr = timed(lambda: A.matvec_dev(x, y), 50)
    en, ed, ev = G.csr_to_ell(ptr, node, val)
    del ptr, node, val
    E = sb.ellpack_matrix(n, n, en, ed, ev)
    w = en.shape[1]
    del en, ed, ev
    y2 = torch.empty_like(x)
    for _ in range(5):
        E.matvec_dev(x, y2)
    t_ell = timed(lambda: E.matvec_dev(x, y2), 50)
    same = bool(torch.equal(y, y2))
    b_csr, b_ell = 12 * nnz + 20 * n + 4, 12 * w * n + 16 * n
    emit(config="C3", grid=N, n=n, nnz=int(nnz), ell_width=int(w), csr_us=t_csr * 1e6, ell_us=t_ell * 1e6,
         csr_gbs=b_csr / t_csr / 1e9, ell_gbs=b_ell / t_ell / 1e9, csr_frac=b_csr / t_csr / 1e9 / HBM,
         ell_frac=b_ell / t_ell / 1e9 / HBM, results_identical=same)
    A.destroy(); E.destroy()

# ---- C4': FEM P1 Laplacian, Jacobi-PCG ----------------------------------------
N = 1200
t0 = time.time()
ptr, node, val = G.fem_p1_csr(N)
n = N * N
A = sb.csr_matrix(n, n, ptr, node, val)
nnz = node.size
rng = np.random.default_rng(1)
v = rng.random(n)
f = A.matvec(v)
tol = 1e-10 * float(np.linalg.norm(f))
s, pc = sb.cg(tol), sb.jacobi()
s.setup(A); pc.setup(A)
fd = torch.from_numpy(f).cuda(); xd = torch.zeros(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
t1 = time.time()
s.solve_dev(A, xd, fd, pc)
torch.cuda.synchronize()
dt = time.time() - t1
it, res2, capped = s.info()
err = float(np.abs(xd.cpu().numpy() - v).max())
emit(config="C4'", what="P1 FEM Laplacian (jittered grid %dx%d), Jacobi-PCG to 1e-10*|f|" % (N, N), n=n, nnz=int(nnz),
     iterations=it, seconds=dt, it_per_s=it / dt, final_res=float(np.sqrt(res2)), max_err_vs_manufactured=err,
     gen_seconds=t1 - t0)
A.destroy()

# ---- C5': Erdos-Renyi, BiCGSTAB (skewed, shifted) and Lanczos -------------------
n = 2_000_000
t0 = time.time()
ptr, node, val = G.erdos_renyi_csr(n, seed=7, shift=1.0, weights="random", skew=True)
A = sb.csr_matrix(n, n, ptr, node, val)
nnz = node.size
v = np.random.default_rng(2).random(n)
f = A.matvec(v)
s = sb.bicgstab(1e-10 * float(np.linalg.norm(f)))
s.set_max_iterations(5000)
s.setup(A)
fd = torch.from_numpy(f).cuda(); xd = torch.zeros(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
t1 = time.time()
s.solve_dev(A, xd, fd)
torch.cuda.synchronize()
dt = time.time() - t1
it, res2, capped = s.info()
err = float(np.abs(xd.cpu().numpy() - v).max())
x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
t_spmv = timed(lambda: A.matvec_dev(x, y), 30)
emit(config="C5'", what="Erdos-Renyi n=2e6 p=log2(n)/n, A = L_w + I + skew, BiCGSTAB to 1e-10*|f|", n=n, nnz=int(nnz),
     iterations=it, seconds=dt, it_per_s=it / dt, capped=capped, final_res=float(np.sqrt(res2)), max_err=err,
     spmv_us=t_spmv * 1e6, spmv_gbs=(12 * nnz + 20 * n) / t_spmv / 1e9, spmv_frac=(12 * nnz + 20 * n) / t_spmv / 1e9 / HBM,
     gen_seconds=t1 - t0)
A.destroy()
ptr, node, val = G.erdos_renyi_csr(n, seed=7, shift=0.0)
L = sb.csr_matrix(n, n, ptr, node, val)
nq = 32
q1 = 2 * np.random.default_rng(3).random(n) - 1
t1 = time.time()
T, Q = sb.lanczos(L, nq, q1)
dt = time.time() - t1
orth = float(np.sqrt(((Q.T @ Q - np.eye(nq)) ** 2).sum()) / nq)
ritz = np.linalg.eigvalsh(np.diag(T[1]) + np.diag(T[2, :-1], 1) + np.diag(T[2, :-1], -1))
emit(config="C5'", what="Lanczos %d steps on the ER graph Laplacian (host-pointer call incl. copy-back of Q)" % nq, n=n,
     steps=nq, seconds=dt, steps_per_s=nq / dt, orthogonality=orth, ritz_min=float(ritz[0]), ritz_max=float(ritz[-1]))
