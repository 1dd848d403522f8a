"""Worker of tests/test_gpu_dist.py::test_a_missing_peer_is_an_error: two ranks own a row-sharded
operator; rank 1 never issues its SpMV.  Rank 0's kernel waits for the halo flag of rank 1, gives up
after SIGB_WAIT_TIMEOUT_MS, and the C-ABI call must return SIGB_ERR_COMM -- not a wrong vector, not
a hung GPU -- and stay in that state."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import sigma_b200 as sb  # noqa: E402
from sigma_b200 import _capi  # noqa: E402
from sigma_b200 import distributed as D  # noqa: E402
from sigma_b200 import generators as G  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sb.init(local)
    comm = D.Comm.from_torch()
    assert comm.nranks == 2 and comm.transport == "peer-memory"
    N = 64
    n = N * N
    ptr, node, val = G.poisson2d_csr(N)
    part = D.partition_rows(ptr, 2)
    lo, hi = int(part[comm.rank]), int(part[comm.rank + 1])
    sl = slice(ptr[lo] - 1, ptr[hi] - 1)
    A = D.dist_csr_matrix(comm, n, part, ptr[lo:hi + 1], node[sl], val[sl])
    x = np.random.default_rng(0).random(n)
    # a healthy exchange first: both ranks take part
    y = A.matvec(x[lo:hi])
    assert np.isfinite(y).all()
    dist.barrier()
    if comm.rank == 0:
        t0 = time.time()
        try:
            A.matvec(x[lo:hi])
            raise SystemExit("the SpMV returned although the peer never pushed its halo")
        except sb.SigmaError as e:
            assert e.status == _capi.ERR_COMM, (e.status, e.message)
            assert "timed out" in e.message, e.message
        waited = time.time() - t0
        assert waited < 30.0, waited
        try:                                   # sticky: the sharded state cannot be trusted any more
            A.matvec(x[lo:hi])
            raise SystemExit("a second call succeeded after the fault")
        except sb.SigmaError as e:
            assert e.status == _capi.ERR_COMM
        print(f"fault ok (gave up after {waited:.1f} s)")
    dist.barrier()
    # no orderly teardown of the communicator: its peer is in an undefined protocol state
    os._exit(0)


if __name__ == "__main__":
    main()
