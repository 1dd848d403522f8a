#!/usr/bin/env python
"""SpMV probe for kernel work on irregular matrices (one GPU, JSON lines; also the target of ncu).

    python scripts/spmv_probe.py --kind er --n 2000000 [--surrogate] [--reps 30]

kind er        : Erdos-Renyi graph Laplacian + I (generators.erdos_renyi_csr, BASELINE config 5 style)
     surrogate : same row-length law (1 + Poisson(log2 n)) and uniformly random columns, drawn directly
                 (seconds instead of minutes at 2e7 rows; same gather behaviour, not symmetric)
     poisson   : 2-D five-point stencil on sqrt(n)^2
Every run checks the device result against numpy evaluated in stored order on a sample of rows.
"""
import argparse
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

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="er", choices=["er", "surrogate", "poisson"])
ap.add_argument("--n", type=int, default=2_000_000)
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--dot", action="store_true", help="time the SpMV fused with the dot product instead")
args = ap.parse_args()

peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
HBM = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0

t0 = time.time()
n = args.n
if args.kind == "er":
    ptr, node, val = G.erdos_renyi_csr(n, seed=7, shift=1.0, weights="random", skew=True)
elif args.kind == "surrogate":
    rng = np.random.default_rng(7)
    deg = 1 + rng.poisson(np.log2(n), n).astype(np.int64)
    ptr = np.concatenate([[1], 1 + np.cumsum(deg)]).astype(np.int32)
    nnz = int(ptr[-1] - 1)
    node = rng.integers(1, n + 1, nnz, dtype=np.int32)
    val = rng.random(nnz)
else:
    N = int(round(np.sqrt(n)))
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
gen_s = time.time() - t0
nnz = int(node.size)

sb.init(0)
stream = torch.cuda.Stream()
sb.set_stream(stream.cuda_stream)
A = sb.csr_matrix(n, n, ptr, node, val)
xh = np.random.default_rng(1).random(n)
with torch.cuda.stream(stream):
    x = torch.from_numpy(xh).cuda()
    y = torch.empty_like(x)
stream.synchronize()

fn = (lambda: A.matvec_dot_dev(x, y, fetch=False)) if args.dot else (lambda: A.matvec_dev(x, y))
for _ in range(5):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(args.reps):
    fn()
e1.record(stream)
e1.synchronize()
t = e0.elapsed_time(e1) / args.reps * 1e-3

# stored-order check on a sample of rows (bit-exact: rounded products added left to right)
yh = y.cpu().numpy()
rows = np.random.default_rng(2).integers(0, n, 2000)
ok = True
for r in rows:
    z = 0.0
    for k in range(ptr[r] - 1, ptr[r + 1] - 1):
        z = z + val[k] * xh[node[k] - 1]
    ok = ok and (z == yh[r])
by = 12 * nnz + 20 * n + 4
print(json.dumps({"kind": args.kind, "n": n, "nnz": nnz, "nnz_per_row": nnz / n, "x_mb": 8 * n / 1e6, "dot": args.dot,
                  "variant": os.environ.get("SIGB_LIB_VARIANT", ""), "env": {k: v for k, v in os.environ.items() if k.startswith("SIGB_")},
                  "g_gathers_per_s": nnz / t / 1e9,
                  "us": t * 1e6, "algorithmic_bytes": by, "gbs": by / t / 1e9, "frac_of_measured_hbm": by / t / 1e9 / HBM,
                  "rows_bit_exact_sample": bool(ok), "gen_s": gen_s}), flush=True)
