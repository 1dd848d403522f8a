"""Regenerates tests/golden/kat.json.

The reference (Fortran 2003) cannot be compiled in this image, and its tests
hold no vector files; the only result-pinning fixtures it has are its two
deterministic test programs, whose pass bars are misfit <= 1e-14
(test/solver_test_diffusion_1d.f90:111-120) and <= 1e-8
(test/solver_test_advection_diffusion_1d.f90:118-127).  This script runs the
oracle's restatement of exactly those programs (ll_graph add_edge calls ->
ellpack graph -> set_value calls -> solver) and records iteration counts and
misfits, which tests/test_oracle_kat.py then pins.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle as orc  # noqa: E402
from helpers import ell_from_tridiag_calls  # noqa: E402


def main():
    out = {"_made_by": "tests/golden/make_golden.py (oracle run of the reference's two deterministic tests)"}
    nn = 127
    dx = 1.0 / (nn + 1)
    node, deg, val = ell_from_tridiag_calls(orc, nn, 2.0, -1.0, -1.0)
    A = orc.Matrix(orc.ELL, nn, nn, node, val, degrees=deg)
    v = np.array([i * dx * (1.0 - i * dx) for i in range(1, nn + 1)])
    u, it, res2, _ = orc.cg_solve(A, np.zeros(nn), np.full(nn, 2.0 * dx**2), 1e-16)
    out["diffusion_1d"] = {"iterations": it, "res2": res2, "misfit": float(np.abs(u - v).max())}

    nn, c = 1024, 0.5
    dx = 1.0 / (nn + 1)
    node, deg, val = ell_from_tridiag_calls(orc, nn, 2.0, -1.0 + c * dx / 2, -1.0 - c * dx / 2)
    A = orc.Matrix(orc.ELL, nn, nn, node, val, degrees=deg)
    x = np.arange(1, nn + 1) * dx
    v = 2.0 * (x - (np.exp(c * x) - 1) / (np.exp(c) - 1)) / c
    u, it, res2, _ = orc.bicgstab_solve(A, np.zeros(nn), np.full(nn, 2.0 * dx**2), 1e-12)
    out["advection_diffusion_1d"] = {"iterations": it, "misfit": float(np.abs(u - v).max())}
    json.dump(out, open(os.path.join(os.path.dirname(__file__), "kat.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
