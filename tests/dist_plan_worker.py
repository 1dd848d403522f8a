"""Worker of tests/test_dist_plan.py: world_size-2 (or more) gloo run on CPU.
Builds the halo plan of a row-sharded operator through the product's host
logic (sigma_b200.distributed, C-ABI index entry points) and checks it
bit-exactly against the oracle evaluated on the whole matrix."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle as orc  # noqa: E402
from sigma_b200 import distributed as D  # noqa: E402
from sigma_b200 import generators as G  # noqa: E402


def check_case(name, n, ptr, node):
    rank, world = dist.get_rank(), dist.get_world_size()
    part = D.partition_rows(ptr, world)
    assert np.array_equal(part, orc.partition_rows(ptr, world)), name
    lo, hi = int(part[rank]), int(part[rank + 1])
    plan = D.build_plan(part, rank, ptr[lo:hi + 1], node[ptr[lo] - 1: ptr[hi] - 1])
    ohalo, olocal = orc.halo_build(lo, hi, ptr, node)
    assert np.array_equal(plan.halo, ohalo) and np.array_equal(plan.local_node, olocal), name
    # send lists = mirror image of every peer's halo restricted to our rows
    exp_rows, exp_counts = [], []
    for q in range(world):
        qh, _ = orc.halo_build(int(part[q]), int(part[q + 1]), ptr, node)
        mine = qh[(qh > lo) & (qh <= hi)] - lo if q != rank else np.zeros(0, np.int32)
        exp_rows.append(mine.astype(np.int32))
        exp_counts.append(mine.size)
    assert np.array_equal(plan.send_counts, np.array(exp_counts, np.int32)), name
    assert np.array_equal(plan.send_rows, np.concatenate(exp_rows)), name
    # receive counts are consistent with the peers' send counts
    allc = [None] * world
    dist.all_gather_object(allc, plan.send_counts.tolist())
    assert [allc[q][rank] for q in range(world)] == plan.recv_counts.tolist(), name
    # a host-side sharded matvec through the plan equals the oracle's serial one
    rng = np.random.default_rng(5)
    val = rng.standard_normal(node.size)
    x = rng.standard_normal(n)
    send_vals = x[lo:hi][plan.send_rows - 1]
    gathered = [None] * world
    dist.all_gather_object(gathered, (plan.send_counts.tolist(), send_vals))
    halo_vals = []
    for q in range(world):
        cnts, vals = gathered[q]
        off = int(np.sum(cnts[:rank]))
        halo_vals.append(vals[off: off + cnts[rank]])
    xext = np.concatenate([x[lo:hi]] + halo_vals)
    assert np.array_equal(xext[hi - lo:], x[plan.halo - 1]), name
    bptr = ptr[lo:hi + 1] - ptr[lo] + 1
    Ablk = orc.Matrix(orc.CSR, hi - lo, xext.size, plan.local_node, val[ptr[lo] - 1: ptr[hi] - 1], ptr=bptr)
    yfull = orc.matvec(orc.Matrix(orc.CSR, n, n, node, val, ptr=ptr), x)
    assert np.array_equal(orc.matvec(Ablk, xext), yfull[lo:hi]), name


def main():
    dist.init_process_group("gloo")
    N = 24
    ptr, node, _ = G.poisson2d_csr(N)
    check_case("poisson", N * N, ptr, node)
    ptr, node, _ = G.erdos_renyi_csr(500, seed=17)
    check_case("erdos-renyi", 500, ptr, node)
    ptr, node, _ = G.fem_p1_csr(15)
    check_case("fem", 225, ptr, node)
    ptr, node, _ = G.tridiag_csr(9)
    check_case("tiny", 9, ptr, node)
    dist.barrier()
    if dist.get_rank() == 0:
        print("dist plan ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
