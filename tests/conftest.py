import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure), compiled on first use."""
    import oracle

    oracle.build()
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def sb():
    """The product: sigma_b200 over its CUDA library; fails loudly without a GPU."""
    import sigma_b200

    sigma_b200.init(-1)
    return sigma_b200


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = np.maximum(np.abs(b), np.finfo(float).tiny)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0
