"""Single-process multi-GPU mode of the C-ABI (sigb_mgpu_init / sigb_mgpu_csr_create): one process,
one caller thread, whole host arrays; partition, halo and send lists derived in the library.  Runs with
one GPU on any box (the same code path with a single row block) and with 2 / 4 / 8 when visible."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ndev", [1, 2, 4, 8])
def test_single_process_multi_gpu(ndev):
    import torch

    if torch.cuda.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "mgpu_worker.py"), str(ndev)], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert f"mgpu ok ({ndev} GPU(s)" in r.stdout
