#!/usr/bin/env python
"""ILDU(0) on the device: setup, one application and ILDU-preconditioned
CG on the 2-D Poisson matrix, next to plain and Jacobi-preconditioned CG and to the CPU port.
One JSON line per row.   python scripts/bench_ldu.py [--grid 1024]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    import torch

    import sigma_b200 as sb
    from sigma_b200 import generators as G

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    sb.init(0)
    stream = torch.cuda.Stream(device=dev)
    sb.set_stream(stream.cuda_stream)
    N = args.grid
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    b_host, _ = G.poisson2d_rhs(N)
    A = sb.csr_matrix(n, n, ptr, node, val)
    with torch.cuda.stream(stream):
        b = torch.from_numpy(b_host).to(dev)
        x = torch.zeros(n, dtype=torch.float64, device=dev)
    stream.synchronize()

    def emit(**kw):
        print(json.dumps(kw), flush=True)

    def ev_time(fn, reps=1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    pc = sb.ldu()
    t0 = time.perf_counter()
    pc.setup(A)
    sb.synchronize()
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    pc.setup(A)
    sb.synchronize()
    t_again = time.perf_counter() - t0
    nf, nb = pc.factors()[-2:]
    emit(row="ldu setup", grid=N, n=n, levels_forward=nf, levels_backward=nb, first_ms=t_first * 1e3,
         refactor_ms=t_again * 1e3, note="first = read-back + host symbolic + upload + numeric; refactor = numeric only")
    launches0 = sb.launch_count()
    pc.solve_dev(A, x, b)
    per_apply = sb.launch_count() - launches0
    t_apply = ev_time(lambda: pc.solve_dev(A, x, b), 5)
    emit(row="ldu apply (forward sweep, / D, backward sweep)", ms=t_apply * 1e3, launches=int(per_apply),
         us_per_launch=t_apply * 1e6 / per_apply, algorithmic_bytes=12 * (node.size - n) + 40 * n,
         note="latency-bound: one launch per level")
    K = args.steps
    tol = 1e-10 * float(np.linalg.norm(b_host))
    rates = {}
    for name, pcs in (("cg", None), ("cg + jacobi", sb.jacobi()), ("cg + ldu", pc)):
        if pcs is not None and pcs is not pc:
            pcs.setup(A)
        s = sb.cg(tol)
        s.set_max_iterations(K)
        s.setup(A)
        with torch.cuda.stream(stream):
            x.zero_()
        t = ev_time(lambda: s.solve_dev(A, x, b, pcs))
        rates[name] = {"it_per_s": s.info()[0] / t, "res_after": float(np.sqrt(s.info()[1]))}
        s.destroy()
    emit(row=f"{K} CG iterations", **{k: v for k, v in rates.items()},
         note="res_after: stopping quantity after K iterations (r.r for cg, r.z for the preconditioned forms)")
    # CPU port on a bounded sample
    import oracle as orc

    Ns = min(N, 512)
    sp, sn, sv = G.poisson2d_csr(Ns)
    O = orc.Matrix(orc.CSR, Ns * Ns, Ns * Ns, sn, sv, ptr=sp)
    t0 = time.perf_counter()
    F = orc.ldu_setup(O)
    t_setup = time.perf_counter() - t0
    rb = np.ones(Ns * Ns)
    t0 = time.perf_counter()
    for _ in range(5):
        orc.ldu_solve(F, rb)
    t_cpu = (time.perf_counter() - t0) / 5
    emit(row="cpu port: ldu setup / apply", sample=f"Poisson {Ns}^2", setup_ms=t_setup * 1e3, apply_ms=t_cpu * 1e3,
         rows_per_s_apply=Ns * Ns / t_cpu, cores=1, kind="port")


if __name__ == "__main__":
    main()
