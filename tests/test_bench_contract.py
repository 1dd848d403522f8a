"""bench.py on the CPU: the reference arm (the serial port timed on host cores) prints one JSON
line with the contract's keys, on the same metric / unit / workload string as the GPU arm, and
the diagnostic hooks of the GPU arm are no-ops on the product build."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "3", "--warmup", "1", "--grid", "256"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in out, key
    assert out["impl"] == "reference" and out["unit"] == "iters/s" and out["dtype"] == "f64"
    assert out["value"] > 0 and out["vs_baseline"] is None and out["higher_is_better"] is True
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["cores"] == 1
    assert out["e2e"] == {"value": out["value"], "unit": out["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench

    assert out["config"]["workload"] == bench.workload(256) and out["metric"] == bench.METRIC
    assert "nnz=%d" % (5 * 256 * 256 - 4 * 256) in out["config"]["workload"]


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--grid", "64"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_diagnostic_hooks_are_noops_on_the_product_build():
    sys.path.insert(0, ROOT)
    import bench

    bench.phase_report(0, None)
    bench.phase_report(0, 0.4)
    bench.spmv_tile_report(0, None)
    bench.spmv_tile_report(0, 200.0)
