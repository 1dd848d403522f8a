"""Row-sharded path on real GPUs: one rank per GPU under torchrun.  Runs at
world 1 on any GPU box (the sharded code path with a single shard) and at
world 2 when two GPUs are visible."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_row_sharded_parity(world, transport):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    if world == 1 and transport == "nccl":
        pytest.skip("one rank exchanges nothing")
    env = dict(os.environ)
    if transport == "nccl":
        env["SIGB_TRANSPORT"] = "nccl"
    else:
        env.pop("SIGB_TRANSPORT", None)
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert "dist gpu ok" in r.stdout
    if world > 1:
        assert f"transport {'peer-memory' if transport == 'p2p' else 'nccl'}" in r.stdout, r.stdout[-500:]


def test_a_missing_peer_is_an_error():
    """Every device-side wait is bounded (device_utils.cuh spin_wait): a peer that never shows up turns
    into SIGB_ERR_COMM at the next synchronisation instead of stale halo data in the result."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "dist_fault_worker.py")]
    env = dict(os.environ, SIGB_WAIT_TIMEOUT_MS="1500")
    env.pop("SIGB_TRANSPORT", None)
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert "fault ok" in r.stdout
