#!/usr/bin/env python
"""Measurements for the rows widened per SURVEY.md 8f (ranks 2 and 3), one JSON line each:

  composite / sum matvec and CG driven by an expression   (operators.cu)
  device matrix copies csr -> csc / ellpack / csr          (convert.cu)
  ordered add_value stream (FEM assembly)                  (assemble.cu)

2-D Poisson 2048^2 (n = 4 194 304, nnz = 20 963 328; 250 MB of matrix arrays, larger than
the 126 MB L2) and the P1 FEM stream of examples/fem.f90 on 1025^2 vertices (18.9 M
add_value calls).  Kernel times: CUDA events on the library's stream after warm-up; the
copy / assembly entry points synchronise internally and are timed with the host clock
around the C-ABI call (host pointers in, as a Fortran caller would issue them).

    python scripts/bench_widened.py [--grid 2048] [--fem 1025] > profiles/r1_widened_rows.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def split_blocks(n, ptr, node, val, h):
    """The n x n CSR matrix as 2 x 2 CSR blocks split at row / column h (0-based count)."""
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ptr))
    cols = node.astype(np.int64) - 1
    out = []
    for (r0, r1) in ((0, h), (h, n)):
        row_blocks = []
        for (c0, c1) in ((0, h), (h, n)):
            m = (rows >= r0) & (rows < r1) & (cols >= c0) & (cols < c1)
            cnt = np.bincount(rows[m] - r0, minlength=r1 - r0)
            bptr = np.concatenate([[1], 1 + np.cumsum(cnt)]).astype(np.int32)
            row_blocks.append((r1 - r0, c1 - c0, bptr, (cols[m] - c0 + 1).astype(np.int32), val[m]))
        out.append(row_blocks)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=2048)
    ap.add_argument("--fem", type=int, default=1025)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--cg-steps", type=int, default=100)
    args = ap.parse_args()

    import torch

    import sigma_b200 as sb
    from sigma_b200 import generators as G

    peak = 6457.1
    pf = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pf):
        peak = float(json.load(open(pf))["hbm_gbs"])
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    sb.init(0)
    stream = torch.cuda.Stream(device=dev)
    sb.set_stream(stream.cuda_stream)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3   # seconds per call

    def emit(**kw):
        print(json.dumps(kw), flush=True)

    # ------------------------------------------------------------------ operators
    N = args.grid
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    nnz = int(node.size)
    b_host, _ = G.poisson2d_rhs(N)
    A = sb.csr_matrix(n, n, ptr, node, val)
    with torch.cuda.stream(stream):
        x = torch.from_numpy(b_host).to(dev)
        y = torch.empty(n, dtype=torch.float64, device=dev)
        y2 = torch.empty(n, dtype=torch.float64, device=dev)
    stream.synchronize()

    bytes_mono = 12 * nnz + 20 * n + 4
    t_mono = timed(lambda: A.matvec_dev(x, y), args.reps)
    emit(row="csr matvec (baseline for the expressions)", grid=N, n=n, nnz=nnz, us=t_mono * 1e6,
         algorithmic_bytes=bytes_mono, gbs=bytes_mono / t_mono / 1e9, frac_of_measured_hbm=bytes_mono / t_mono / 1e9 / peak)

    blocks = split_blocks(n, ptr, node, val, n // 2)
    mats = [[sb.csr_matrix(r, c, p, nd, v) for (r, c, p, nd, v) in row] for row in blocks]
    S = sb.sparse_matrix([n // 2, n - n // 2], [n // 2, n - n // 2], mats)
    S.matvec_dev(x, y2)
    A.matvec_dev(x, y)
    stream.synchronize()
    same = bool(torch.equal(y, y2))   # block sums split each row's additions: equal only when no row is split
    maxdiff = float((y - y2).abs().max().item())
    # bytes: every block streams its entries and its ptr slice; x is read once per block column pair,
    # y is written by the first block of a block row and read + written by the second
    bytes_comp = 12 * nnz + 4 * 2 * n + 8 * n + 24 * n
    t_comp = timed(lambda: S.matvec_dev(x, y2), args.reps)
    emit(row="composite 2x2 sparse_matrix matvec (4 leaf launches)", us=t_comp * 1e6, algorithmic_bytes=bytes_comp,
         gbs=bytes_comp / t_comp / 1e9, frac_of_measured_hbm=bytes_comp / t_comp / 1e9 / peak,
         vs_monolithic=t_comp / t_mono, bit_equal_to_monolithic=same, max_abs_diff=maxdiff)

    # operator_sum: strictly-lower + (diagonal and upper) parts of the same matrix
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ptr))
    low = (node.astype(np.int64) - 1) < rows

    def part(mask):
        cnt = np.bincount(rows[mask], minlength=n)
        return np.concatenate([[1], 1 + np.cumsum(cnt)]).astype(np.int32), node[mask], val[mask]

    L = sb.csr_matrix(n, n, *part(low))
    U = sb.csr_matrix(n, n, *part(~low))
    LU = L + U
    bytes_sum = 12 * nnz + 4 * 2 * n + 2 * 8 * n + 24 * n
    t_sum = timed(lambda: LU.matvec_dev(x, y2), args.reps)
    emit(row="operator_sum L + U matvec (2 leaf launches)", us=t_sum * 1e6, algorithmic_bytes=bytes_sum,
         gbs=bytes_sum / t_sum / 1e9, frac_of_measured_hbm=bytes_sum / t_sum / 1e9 / peak, vs_monolithic=t_sum / t_mono)

    # adjoint(A) * A applied as an expression (2 SpMVs through the device scratch vector)
    AtA = sb.adjoint(A) * A
    t_ata = timed(lambda: AtA.matvec_dev(x, y2), max(10, args.reps // 2))
    emit(row="operator_product adjoint(A) * A matvec (csr SpMV + transposed SpMV)", us=t_ata * 1e6,
         algorithmic_bytes=2 * bytes_mono, gbs=2 * bytes_mono / t_ata / 1e9,
         frac_of_measured_hbm=2 * bytes_mono / t_ata / 1e9 / peak, vs_monolithic=t_ata / t_mono)

    # CG driven by the composite vs by the plain matrix (kernel-per-phase path for both)
    K = args.cg_steps
    tol = 1e-10 * float(np.linalg.norm(b_host))
    os.environ["SIGB_CG_PERSISTENT"] = "0"
    rates = {}
    for name, op in (("csr_matrix", A), ("composite 2x2", S)):
        solver = sb.cg(tol)
        solver.set_max_iterations(K)
        solver.setup(op)
        with torch.cuda.stream(stream):
            xs = torch.zeros(n, dtype=torch.float64, device=dev)
        solver.solve_dev(op, xs, x)          # warm-up
        solver.setup(op)
        with torch.cuda.stream(stream):
            xs.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        solver.solve_dev(op, xs, x)
        e1.record(stream)
        e1.synchronize()
        it = solver.info()[0]
        rates[name] = it / (e0.elapsed_time(e1) * 1e-3)
        solver.destroy()
    emit(row="CG iterations/s driven by an expression", grid=N, csr_matrix=rates["csr_matrix"],
         composite=rates["composite 2x2"], ratio=rates["composite 2x2"] / rates["csr_matrix"], steps=K)

    # ------------------------------------------------------------------ copies
    def wall(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0, r

    A.copy_matrix("csr").destroy()            # warm-up (first-use allocations, kernel loads)
    for target, trans in (("csr", False), ("csc", False), ("ellpack", False), ("csr", True)):
        t, B = wall(lambda: A.copy_matrix(target, trans))
        by = 24 * nnz + 8 * n
        emit(row=f"copy_matrix csr -> {target}{' (transposed)' if trans else ''}", ms=t * 1e3, algorithmic_bytes=by,
             gbs=by / t / 1e9, frac_of_measured_hbm=by / t / 1e9 / peak, entries_per_s=nnz / t)
        B.destroy()

    # host reference point: the oracle's restatement of the first-free-slot builder on a bounded sample
    import oracle as orc

    Ns = 512
    sp, sn, sv = G.poisson2d_csr(Ns)
    O = orc.Matrix(orc.CSR, Ns * Ns, Ns * Ns, sn, sv, ptr=sp)
    t0 = time.perf_counter()
    orc.copy_matrix(O, orc.CSC)
    t_cpu = time.perf_counter() - t0
    emit(row="cpu port: copy_matrix csr -> csc (cs_graph_build + copy_matrix_values)", sample=f"Poisson {Ns}^2",
         entries_per_s=sn.size / t_cpu, cores=1, kind="port")

    # ------------------------------------------------------------------ assembly
    Nf = args.fem
    I, J, V, _ = G.fem_p1_add_value_stream(Nf)
    fptr, fnode, _ = G.fem_p1_csr(Nf)
    ci, cj = (I + 1).astype(np.int32), (J + 1).astype(np.int32)
    nv = Nf * Nf
    F = sb.csr_matrix(nv, nv, fptr, fnode, np.zeros(fnode.size))
    F.add_values(ci[:1000], cj[:1000], V[:1000])      # warm-up
    t, _ = wall(lambda: F.add_values(ci, cj, V))
    emit(row="add_value stream (P1 FEM assembly, csr)", vertices=nv, calls=int(ci.size), ms=t * 1e3,
         calls_per_s=ci.size / t, h2d_bytes=16 * int(ci.size),
         note="host pointers in: H2D of (i, j, z) inside the timed region; locate + stable bucket sort + ordered reduce")
    ns = min(2_000_000, ci.size)
    Of = orc.Matrix(orc.CSR, nv, nv, fnode, np.zeros(fnode.size), ptr=fptr)
    t0 = time.perf_counter()
    orc.add_values(Of, ci[:ns], cj[:ns], V[:ns])
    t_cpu = time.perf_counter() - t0
    emit(row="cpu port: add_value loop", sample=f"first {ns} calls", calls_per_s=ns / t_cpu, cores=1, kind="port")


if __name__ == "__main__":
    main()
