#!/usr/bin/env python
"""BASELINE.json configs 4 and 5 at their named sizes, row-sharded over the GPUs of one box
(JSON lines on rank 0).  Not run in round 1 (GPU budget); first thing to run in round 2.

  C4  P1 FEM Laplacian on a jittered 7071 x 7071 vertex grid (n = 49 999 041, ~3.5e8 stored
      entries), Jacobi-preconditioned CG to 1e-10 |f|
  C5  Erdos-Renyi G(n, log2(n)/n), n = 2e7 (~5e8 stored entries):
      BiCGSTAB on A = L_w + I + skew part, to 1e-10 |f|;  64 Lanczos steps on the graph Laplacian

Every rank generates ONLY its own rows (generators.fem_p1_csr_rows / erdos_renyi_csr_rows:
tests/test_generators_rows.py shows they equal slices of the whole matrix), so generation
costs ~30-60 s of host time per rank, in parallel.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29533 scripts/bench_configs_dist.py [--fem-grid 7071] [--er-n 20000000]
  python scripts/bench_configs_dist.py                      # one GPU, same code path
  ... --dry-run                                             # CPU only (gloo): generation, partition, halo plan

Timing: CUDA events on the library's stream around the device-resident solve, max over ranks.
Parity at these sizes is checked through size-independent properties: the manufactured
solution is recovered (|x - x*|_inf), the recurrence residual reaches the tolerance, Lanczos
vectors are orthonormal to 1e-14 per entry.
"""
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
    ap.add_argument("--fem-grid", type=int, default=7071)
    ap.add_argument("--er-n", type=int, default=20_000_000)
    ap.add_argument("--lanczos-steps", type=int, default=64)
    ap.add_argument("--max-iters", type=int, default=100_000)
    ap.add_argument("--skip", default="", help="comma list of c4,c5b,c5l")
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args()
    skip = set(args.skip.split(",")) if args.skip else set()

    import torch
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_RANK", "0")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])

    import sigma_b200 as sb
    from sigma_b200 import distributed as D
    from sigma_b200 import generators as G

    dev = None
    if args.dry_run:
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=dev)
        sb.init(local)
        stream = torch.cuda.Stream(device=dev)
        sb.set_stream(stream.cuda_stream)
        comm = D.Comm.from_torch(dev)
    hbm = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        hbm = json.load(open(pk)).get("hbm_gbs", hbm)

    def emit(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    def allsum(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev if dev is not None else "cpu")
        dist.all_reduce(t)
        return float(t.item())

    def allmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev if dev is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        return allmax(e0.elapsed_time(e1)) * 1e-3

    def cpu_spmv_sample(n, lo, hi, ptr_blk, node, val, max_rows=2_000_000):
        """The serial port (oracle) on a bounded sample: rank 0's first rows, global columns, one host core."""
        if rank != 0 or args.dry_run:
            return None
        import oracle as orc

        m = min(hi - lo, max_rows)
        p0 = (ptr_blk[: m + 1] - ptr_blk[0] + 1).astype(np.int32)
        ne = int(p0[-1] - 1)
        O = orc.Matrix(orc.CSR, m, n, node[:ne], val[:ne], ptr=p0)
        xg = np.random.default_rng(5).random(n)
        orc.matvec(O, xg)
        t0 = time.perf_counter()
        orc.matvec(O, xg)
        dt = time.perf_counter() - t0
        return {"kind": "port", "cores": 1, "sample": f"first {m} rows of rank 0's block ({ne} entries), serial csr_matvec restatement",
                "seconds": dt, "gbs": (12 * ne + 20 * m) / dt / 1e9}

    def equal_rows(n):
        return np.array([(n * r) // world for r in range(world + 1)], np.int32)

    def dry(name, n, part, ptr_blk, node, gen_s):
        plan = D.build_plan(part, rank, ptr_blk, node)
        emit(config=name, dry_run=True, n=n, world=world, rows_rank0=int(part[1] - part[0]), nnz_rank0=int(node.size),
             halo_rank0=int(plan.halo.size), send_rank0=int(plan.send_rows.size), gen_seconds=gen_s)

    # ------------------------------------------------------------------ C4
    if "c4" not in skip:
        N = args.fem_grid
        n = N * N
        part = equal_rows(n)
        lo, hi = int(part[rank]), int(part[rank + 1])
        t0 = time.time()
        ptr_blk, node, val = G.fem_p1_csr_rows(N, lo, hi)
        gen_s = allmax(time.time() - t0)
        nnz = int(allsum(node.size))
        if args.dry_run:
            dry("C4", n, part, ptr_blk, node, gen_s)
        else:
            cpu = cpu_spmv_sample(n, lo, hi, ptr_blk, node, val)
            A = D.dist_csr_matrix(comm, n, part, ptr_blk, node, val)
            del ptr_blk, node, val
            xs = np.random.default_rng(1).random(n)[lo:hi]            # manufactured solution, independent of the sharding
            f = A.matvec(xs)
            tol = 1e-10 * np.sqrt(allsum(float(f @ f)))
            s, pc = sb.cg(tol), sb.jacobi()
            s.set_max_iterations(args.max_iters)
            s.setup(A)
            pc.setup(A)
            fd = torch.from_numpy(f).to(dev)
            xd = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
            dt = timed(lambda: s.solve_dev(A, xd, fd, pc))
            it, res2, capped = s.info()
            err = allmax(float(np.abs(xd.cpu().numpy() - xs).max()))
            x = torch.rand(hi - lo, dtype=torch.float64, device=dev)
            y = torch.empty_like(x)
            for _ in range(3):
                A.matvec_dev(x, y)
            t_spmv = timed(lambda: [A.matvec_dev(x, y) for _ in range(20)]) / 20
            b_spmv = 12 * nnz + 20 * n
            # PCG iteration as implemented: SpMV+dot, r/z update + r.z (r, q, idiag in; r, z out = 40 n), x/p update (40 n)
            b_it = 12 * nnz + 100 * n
            emit(config="C4", what="P1 FEM Laplacian, jittered %d x %d vertex grid, Jacobi-PCG to 1e-10*|f|" % (N, N), n=n,
                 nnz=nnz, n_gpus=world, iterations=int(it), capped=bool(capped), seconds=dt, it_per_s=it / dt,
                 final_res=float(np.sqrt(res2)), tol=tol, max_err_vs_manufactured=err, spmv_us=t_spmv * 1e6,
                 spmv_gbs=b_spmv / t_spmv / 1e9, spmv_frac_of_hbm=b_spmv / t_spmv / 1e9 / (hbm * world),
                 pcg_gbs=b_it * it / dt / 1e9, pcg_frac_of_hbm=b_it * it / dt / 1e9 / (hbm * world), gen_seconds=gen_s,
                 transport=comm.transport if world > 1 else "none", cpu_baseline_spmv=cpu)
            s.destroy(); pc.destroy(); A.destroy()
            del xd, fd, x, y

    # ------------------------------------------------------------------ C5
    cache = {}
    if "c5b" not in skip or "c5l" not in skip:
        n = args.er_n
        t0 = time.time()
        # counts of the whole graph come with any block; the partition balances stored entries
        _, _, _, counts = G.erdos_renyi_csr_rows(n, 0, 0, seed=7, cache=cache)
        gptr = np.concatenate([[1], 1 + np.cumsum(counts)])
        assert gptr[-1] < 2**31, "stored entries must fit the reference's default integer"
        part = D.partition_rows(gptr.astype(np.int32), world)
        lo, hi = int(part[rank]), int(part[rank + 1])
        del gptr
        pairs_s = allmax(time.time() - t0)

    if "c5b" not in skip:
        t0 = time.time()
        ptr_blk, node, val, _ = G.erdos_renyi_csr_rows(n, lo, hi, seed=7, shift=1.0, weights="random", skew=True,
                                                       cache=cache)
        gen_s = allmax(time.time() - t0) + pairs_s
        nnz = int(allsum(node.size))
        if args.dry_run:
            dry("C5 bicgstab", n, part, ptr_blk, node, gen_s)
        else:
            cpu = cpu_spmv_sample(n, lo, hi, ptr_blk, node, val)
            A = D.dist_csr_matrix(comm, n, part, ptr_blk, node, val)
            del ptr_blk, node, val
            xs = np.random.default_rng(2).random(n)[lo:hi]
            f = A.matvec(xs)
            tol = 1e-10 * np.sqrt(allsum(float(f @ f)))
            s = sb.bicgstab(tol)
            s.set_max_iterations(5000)
            s.setup(A)
            fd = torch.from_numpy(f).to(dev)
            xd = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
            dt = timed(lambda: s.solve_dev(A, xd, fd))
            it, res2, capped = s.info()
            err = allmax(float(np.abs(xd.cpu().numpy() - xs).max()))
            x = torch.rand(hi - lo, dtype=torch.float64, device=dev)
            y = torch.empty_like(x)
            for _ in range(3):
                A.matvec_dev(x, y)
            t_spmv = timed(lambda: [A.matvec_dev(x, y) for _ in range(20)]) / 20
            b_spmv = 12 * nnz + 20 * n
            emit(config="C5", what="Erdos-Renyi n=%d p=log2(n)/n, A = L_w + I + skew, BiCGSTAB to 1e-10*|f|" % n, n=n,
                 nnz=nnz, n_gpus=world, iterations=int(it), capped=bool(capped), seconds=dt, it_per_s=it / dt,
                 final_res=float(np.sqrt(res2)), tol=tol, max_err_vs_manufactured=err, spmv_us=t_spmv * 1e6,
                 spmv_gbs=b_spmv / t_spmv / 1e9, spmv_frac_of_hbm=b_spmv / t_spmv / 1e9 / (hbm * world),
                 halo_per_rank=int(A.plan.halo.size), gen_seconds=gen_s,
                 transport=comm.transport if world > 1 else "none", cpu_baseline_spmv=cpu,
                 note="random columns: gather-bound, not stream-bound -- with x (160 MB) beyond the L2 every gathered "
                      "entry costs a 64-byte DRAM access (ncu: 61.8 B per gather) and a B200 serves 96 G such gathers/s "
                      "to a kernel that does nothing else (profiles/r2_gather_probe.jsonl): 12 algorithmic bytes per "
                      "entry against 12 + 64 moved bound the fraction near 0.16 of HBM per GPU")
            s.destroy(); A.destroy()
            del xd, fd, x, y

    if "c5l" not in skip:
        t0 = time.time()
        ptr_blk, node, val, _ = G.erdos_renyi_csr_rows(n, lo, hi, seed=7, shift=0.0, cache=cache)
        gen_s = allmax(time.time() - t0)
        cache.clear()
        if args.dry_run:
            dry("C5 lanczos", n, part, ptr_blk, node, gen_s)
        else:
            L = D.dist_csr_matrix(comm, n, part, ptr_blk, node, val)
            del ptr_blk, node, val
            nq = args.lanczos_steps
            q1 = (2 * np.random.default_rng(3).random(n) - 1)[lo:hi]
            q1d = torch.from_numpy(q1).to(dev)
            Qd = torch.empty((hi - lo) * nq, dtype=torch.float64, device=dev)
            sb.lanczos_dev(L, min(nq, 3), Qd, q1d)                    # warm-up
            dist.barrier()
            torch.cuda.synchronize()
            T = None

            def run():
                nonlocal T
                T = sb.lanczos_dev(L, nq, Qd, q1d)

            dt = timed(run)
            # orthonormality of the sharded basis: Q^T Q summed over ranks (on the device)
            Qm = Qd.view(nq, hi - lo)
            gram = Qm @ Qm.T
            dist.all_reduce(gram)
            gram = gram.cpu().numpy()
            orth = float(np.sqrt(((gram - np.eye(nq)) ** 2).sum()) / nq)
            ritz = np.linalg.eigvalsh(np.diag(T[1]) + np.diag(T[2, :-1], 1) + np.diag(T[2, :-1], -1))
            emit(config="C5", what="Lanczos %d steps on the ER graph Laplacian, basis resident on the devices (sigb_lanczos_dev)" % nq,
                 n=n, n_gpus=world, steps=nq, seconds=dt, steps_per_s=nq / dt, orthogonality=orth, ritz_min=float(ritz[0]),
                 ritz_max=float(ritz[-1]), gen_seconds=gen_s)
            del Qd, q1d
            L.destroy()

    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
