"""The C++ host mirror (sigma_b200/host/sigma.hpp) compiles on CPU; on a GPU its
restatements of the reference's test programs pass with exit code 0, exactly
how the reference's CTest registers them (test/CMakeLists.txt:27-32)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = os.path.join(ROOT, "tests", "cxx")
PROGS = ["solver_test_diffusion_1d", "solver_test_advection_diffusion_1d", "solver_test_jacobi",
         "eigensolver_test_lanczos", "eigensolver_test_generalized_lanczos", "matrix_test_basics",
         "linear_operator_test_algebra", "matrix_test_composite", "matrix_test_copy",
         "solver_test_incomplete_cholesky"]


# (first run on a GPU in round 2, visit r2a)
PROGS += ["matrix_test_strategy", "matrix_test_set_multiple_entries", "matrix_test_set_entry_with_realloc",
          "matrix_test_permute"]
# the multi-GPU path through the host mirror alone: all visible GPUs, one process, no Python
PROGS += ["solver_test_multi_gpu"]


def build():
    subprocess.check_call(["make", "-s", "-C", CXX])


def test_cxx_programs_compile_and_link():
    build()
    for p in PROGS:
        assert os.path.exists(os.path.join(CXX, "_build", p))


@pytest.mark.gpu
@pytest.mark.parametrize("prog", PROGS)
def test_reference_test_program(prog):
    build()
    r = subprocess.run([os.path.join(CXX, "_build", prog), "-v"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_set_entry_with_realloc_host_side():
    """test/matrix_test_set_entry_with_realloc.f90: growing the graph under set_value / add_value /
    add_multiple_values is host-side work in the reference and in the mirror, so the restated
    program runs here without a GPU (--host-only skips its final device matvec)."""
    build()
    r = subprocess.run([os.path.join(CXX, "_build", "matrix_test_set_entry_with_realloc"), "-v", "--host-only"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("setting unallocated entries works") == 3


def test_permutations_host_side():
    """The permutation part of test/matrix_test_basics.f90 (:139-158, :364-392): right_permute and
    left_permute of csr / csc / ellpack matrices are host-side index work in the reference and in
    the mirror; the restated program runs here without a GPU (--host-only skips its device matvec)."""
    build()
    r = subprocess.run([os.path.join(CXX, "_build", "matrix_test_permute"), "-v", "--host-only"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("right and left permutation work") == 3


def test_strategy_host_side():
    """test/matrix_test_strategy.f90 up to its matvec: entries through the 1 x 1 composite, rows and
    columns through the storage strategy (get_row / get_column), all host-side reads -- run here
    without a GPU (--host-only); the matvec part runs with the gated programs."""
    build()
    r = subprocess.run([os.path.join(CXX, "_build", "matrix_test_strategy"), "-v", "--host-only"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("entries, rows and columns work") == 3


def test_multi_gpu_row_forms_host_side():
    """A csc_matrix / ellpack_matrix in multi-GPU mode is sharded through its rows (sigma.hpp
    rows_in_matvec_order / rows_without_padding): the csr row loop over those rows must reproduce the format's
    own reference matvec loop bit for bit -- host-side index work, checked here without a GPU."""
    build()
    r = subprocess.run([os.path.join(CXX, "_build", "solver_test_multi_gpu"), "-v", "--host-only"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "row forms of csc / ellpack matrices reproduce their own matvec loops" in r.stdout
