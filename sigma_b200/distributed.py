"""Row-sharded operators: the host-side plumbing around sigb_dist_csr_create.

One process per GPU (torchrun).  torch.distributed is used ONLY for setup
plumbing -- broadcasting the communicator id and exchanging the index lists of
the halo plan; the data path (halo exchange, dot-product all-reduces) runs
inside libsigma_b200.so.

The plan is pure int32 index work derived from the sparsity pattern
("graph-derived halo lists"):
  part      row offsets balancing stored entries     (sigb_partition_rows)
  halo      sorted unique global columns a rank reads but does not own
            (sigb_halo_build); grouped by owner because owners are contiguous
  send list the mirror image on the owner: which of its rows each peer needs,
            obtained with one all-to-all of the halo lists.
tests/test_dist_plan.py checks all of it bit-exactly against the oracle, on CPU
with the gloo backend at world_size 2.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import as_f64, as_i32, check, lib, ptr
from .api import Graph, Matrix

__all__ = ["Comm", "HaloPlan", "partition_rows", "build_plan", "dist_csr_matrix", "setup_poisson"]


class Comm:
    """sigb_comm_t joined through torch.distributed (any backend)."""

    def __init__(self, handle, rank, nranks):
        self._h, self.rank, self.nranks = handle, rank, nranks

    @classmethod
    def from_torch(cls, device=None):
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        uid = np.zeros(_capi.UNIQUE_ID_BYTES, np.uint8)
        if rank == 0:
            check(lib().sigb_comm_unique_id(ptr(uid)))
        t = torch.from_numpy(uid)
        if dist.get_backend() == "nccl":
            t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.broadcast(t, 0)
        uid = t.cpu().numpy().copy()
        h = C.c_void_p()
        check(lib().sigb_comm_create(ptr(uid), rank, world, C.byref(h)))
        return cls(h, rank, world)

    @property
    def transport(self):
        """'peer-memory' (IPC windows over NVLink) or 'nccl'."""
        r, n, p = C.c_int(), C.c_int(), C.c_int()
        check(lib().sigb_comm_info(self._h, C.byref(r), C.byref(n), C.byref(p)))
        return "peer-memory" if p.value else "nccl"

    def destroy(self):
        if self._h:
            check(lib().sigb_comm_destroy(self._h))
            self._h = None


def partition_rows(ptr1, nparts):
    """Row offsets (0-based, nparts+1) balancing stored entries."""
    ptr1 = as_i32(ptr1)
    part = np.empty(nparts + 1, np.int32)
    check(lib().sigb_partition_rows(ptr1.size - 1, ptr(ptr1), nparts, ptr(part)))
    return part


@dataclass
class HaloPlan:
    lo: int
    hi: int
    halo: np.ndarray          # sorted unique global 1-based columns we need
    local_node: np.ndarray    # our entries' columns in [owned | halo] numbering
    recv_counts: np.ndarray   # per owner rank
    send_counts: np.ndarray   # per destination rank
    send_rows: np.ndarray     # 1-based local rows, grouped by destination


def _halo_build(lo, hi, ptr_blk, node_glob):
    ptr_blk, node_glob = as_i32(ptr_blk), as_i32(node_glob)
    cnt = int(ptr_blk[-1] - ptr_blk[0])
    halo = np.empty(max(cnt, 1), np.int32)
    local = np.empty(max(cnt, 1), np.int32)
    nh = C.c_int32()
    check(lib().sigb_halo_build(lo, hi, ptr(ptr_blk), ptr(node_glob), ptr(halo), C.byref(nh), ptr(local)))
    return halo[: nh.value].copy(), local[:cnt].copy()


def build_plan(part, rank, ptr_blk, node_glob, exchange=None):
    """Halo list of this rank and, through `exchange`, its send lists.

    exchange(list_of_arrays_per_destination) -> list_of_arrays_per_source is an
    all-to-all of int32 arrays; default: torch.distributed.
    """
    part = as_i32(part)
    P = part.size - 1
    lo, hi = int(part[rank]), int(part[rank + 1])
    halo, local = _halo_build(lo, hi, ptr_blk, node_glob)
    # owner q of 1-based column c: part[q] < c <= part[q+1]
    owner = np.searchsorted(part[1:], halo, side="left") if halo.size else np.zeros(0, np.int64)
    recv_counts = np.bincount(owner, minlength=P).astype(np.int32)
    requests = [halo[owner == q] for q in range(P)]          # what we ask each owner for
    if exchange is None:
        exchange = _torch_all_to_all
    wanted = exchange(requests) if P > 1 else [np.zeros(0, np.int32)]
    send_counts = np.array([w.size for w in wanted], np.int32)
    send_rows = (np.concatenate(wanted).astype(np.int64) - lo).astype(np.int32) if P > 1 else np.zeros(0, np.int32)
    if send_rows.size and (send_rows.min() < 1 or send_rows.max() > hi - lo):
        raise _capi.SigmaError(_capi.ERR_ARG, "a peer asked for a row this rank does not own")
    return HaloPlan(lo, hi, halo, local, recv_counts, send_counts, send_rows)


def _torch_all_to_all(requests):
    import torch
    import torch.distributed as dist

    P = dist.get_world_size()
    objs = [None] * P
    # index lists are small next to the matrix; object all-gather works on every backend
    dist.all_gather_object(objs, [np.asarray(r, np.int32) for r in requests])
    me = dist.get_rank()
    return [np.asarray(objs[src][me], np.int32) for src in range(P)]


def dist_csr_matrix(comm: Comm, n_global, part, ptr_blk, node_glob, val, plan: HaloPlan | None = None):
    """Row block of a global CSR matrix as a device operator (a Matrix whose
    matvec / solvers work on the owned slices of x, y, b)."""
    part, ptr_blk, node_glob = as_i32(part), as_i32(ptr_blk), as_i32(node_glob)
    if plan is None:
        plan = build_plan(part, comm.rank, ptr_blk, node_glob)
    h = C.c_void_p()
    sc, sr = as_i32(plan.send_counts), as_i32(plan.send_rows)
    check(lib().sigb_dist_csr_create(comm._h, int(n_global), ptr(part), ptr(ptr_blk), ptr(node_glob), ptr(sc),
                                     ptr(sr) if sr.size else None, C.byref(h)))
    nloc = int(part[comm.rank + 1] - part[comm.rank])
    g = Graph(None, "csr", nloc, nloc)          # the library owns the local graph
    A = Matrix(g, handle=h)
    A.plan, A.comm, A.n_global = plan, comm, int(n_global)
    val = as_f64(val)
    check(lib().sigb_matrix_set_values(A._h, ptr(val), val.size))
    return A


@dataclass
class PoissonShard:
    A: Matrix
    comm: Comm
    part: np.ndarray
    nloc: int
    nnz_loc: int
    nnz_glob: int
    b: np.ndarray
    xs: np.ndarray


def setup_poisson(N, rank, world, device=None):
    """BASELINE config 2 row-sharded: each rank generates only its own block."""
    from . import generators as G

    n = N * N
    # per-row entry counts of the 5-point stencil (no matrix needed for the partition)
    k = np.arange(n, dtype=np.int64)
    ix, iy = k // N, k % N
    cnt = (1 + (ix > 0) + (iy > 0) + (iy < N - 1) + (ix < N - 1)).astype(np.int64)
    gptr = np.concatenate([[1], 1 + np.cumsum(cnt)]).astype(np.int32)
    part = partition_rows(gptr, world)
    lo, hi = int(part[rank]), int(part[rank + 1])
    ptr_blk, node_glob, val = G.poisson2d_csr(N, lo, hi)
    b, xs = G.poisson2d_rhs(N, row_lo=lo, row_hi=hi)
    comm = Comm.from_torch(device)
    A = dist_csr_matrix(comm, n, part, ptr_blk, node_glob, val)
    return PoissonShard(A, comm, part, hi - lo, int(node_glob.size), int(gptr[-1] - 1), b, xs)
