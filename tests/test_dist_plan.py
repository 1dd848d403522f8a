"""Multi-rank host logic on CPU: gloo backend, world_size 2 and 3."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_halo_plan_gloo(world, orc):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "dist_plan_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist plan ok" in r.stdout


def test_full_size_config_script_dry_run_gloo():
    """scripts/bench_configs_dist.py --dry-run at world 2 (gloo, no GPU): every rank generates only
    its own rows of the FEM / Erdos-Renyi matrices, the partition and the halo plan are built;
    the ranks must agree on sizes and the halo of a 2-way split FEM grid is one grid line."""
    import json

    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "scripts", "bench_configs_dist.py"), "--dry-run", "--fem-grid", "120", "--er-n", "30000"]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert [x["config"] for x in rows] == ["C4", "C5 bicgstab", "C5 lanczos"]
    fem = rows[0]
    assert fem["n"] == 120 * 120 and fem["rows_rank0"] == 7200
    # rank 0 owns vertex rows 0..59 of the grid; its elements reach one more grid line (+ one vertex
    # of the diagonal cells): a halo of N or N + 1 vertices
    assert 120 <= fem["halo_rank0"] <= 121 and fem["send_rank0"] == fem["halo_rank0"]
    assert rows[1]["n"] == 30000 and rows[1]["nnz_rank0"] == rows[2]["nnz_rank0"]      # same graph, two sets of values
