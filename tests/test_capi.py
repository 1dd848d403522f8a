"""The C-ABI library loads without a GPU and exports exactly the symbols
include/sigma_b200.h declares; host-only index entry points agree with the
oracle.  CPU only -- no compute call is made."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sigma_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"SIGB_API\s+[\w\s\*]+?\b(sigb_\w+)\s*\(", src)))


def test_header_declares_something():
    syms = declared_symbols()
    assert len(syms) >= 40 and "sigb_solver_solve" in syms and "sigb_matvec" in syms


def test_library_exports_every_declared_symbol():
    from sigma_b200 import _capi

    assert os.path.exists(_capi.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _capi.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    extra = [s for s in exported if s.startswith("sigb_") and s not in declared_symbols()]
    assert not extra, extra


def test_python_prototypes_cover_header():
    from sigma_b200 import _capi

    assert sorted(_capi.PROTOTYPES) == declared_symbols()
    L = _capi.lib()  # dlopen works without a GPU
    assert L.sigb_version().decode().startswith("sigma_b200")


def test_compute_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import sigma_b200
    from sigma_b200 import SigmaError

    with pytest.raises(SigmaError) as e:
        sigma_b200.init(0)
    assert e.value.status == 2 and "no CPU path" in e.value.message


@pytest.mark.parametrize("P", [1, 2, 3, 8])
def test_partition_and_halo_match_oracle(orc, P):
    from sigma_b200 import _capi, generators as G

    L = _capi.lib()
    N = 12
    n = N * N
    ptr, node, _ = G.poisson2d_csr(N)
    part = np.empty(P + 1, np.int32)
    _capi.check(L.sigb_partition_rows(n, _capi.ptr(ptr), P, _capi.ptr(part)))
    assert np.array_equal(part, orc.partition_rows(ptr, P))
    assert part[0] == 0 and part[-1] == n and np.all(np.diff(part) >= 0)
    for r in range(P):
        lo, hi = int(part[r]), int(part[r + 1])
        ohalo, olocal = orc.halo_build(lo, hi, ptr, node)
        bptr = np.ascontiguousarray(ptr[lo:hi + 1])
        bnode = np.ascontiguousarray(node[ptr[lo] - 1: ptr[hi] - 1])
        halo = np.empty(max(bnode.size, 1), np.int32)
        local = np.empty(max(bnode.size, 1), np.int32)
        nh = C.c_int32()
        _capi.check(L.sigb_halo_build(lo, hi, _capi.ptr(bptr), _capi.ptr(bnode), _capi.ptr(halo), C.byref(nh),
                                      _capi.ptr(local)))
        assert np.array_equal(halo[: nh.value], ohalo)
        assert np.array_equal(local[: bnode.size], olocal)
        if P > 1 and hi > lo:
            assert nh.value <= 2 * N  # two grid lines at most


def test_halo_irregular_graph(orc):
    from sigma_b200 import _capi, generators as G

    L = _capi.lib()
    n = 300
    ptr, node, _ = G.erdos_renyi_csr(n, seed=9)
    part = orc.partition_rows(ptr, 4)
    for r in range(4):
        lo, hi = int(part[r]), int(part[r + 1])
        ohalo, olocal = orc.halo_build(lo, hi, ptr, node)
        bptr = np.ascontiguousarray(ptr[lo:hi + 1])
        bnode = np.ascontiguousarray(node[ptr[lo] - 1: ptr[hi] - 1])
        halo = np.empty(bnode.size, np.int32)
        local = np.empty(bnode.size, np.int32)
        nh = C.c_int32()
        _capi.check(L.sigb_halo_build(lo, hi, _capi.ptr(bptr), _capi.ptr(bnode), _capi.ptr(halo), C.byref(nh),
                                      _capi.ptr(local)))
        assert np.array_equal(halo[: nh.value], ohalo) and np.array_equal(local, olocal)


def declared_arity():
    """{symbol: number of parameters} parsed from the header."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"SIGB_API\s+[\w\s\*]+?\b(sigb_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def test_python_prototypes_have_the_header_arity():
    """A ctypes prototype with the wrong number of arguments corrupts the call silently: every
    binding must take exactly as many arguments as the header declares."""
    from sigma_b200 import _capi

    arity = declared_arity()
    assert sorted(arity) == declared_symbols()
    wrong = {name: (len(proto[1]), arity[name]) for name, proto in _capi.PROTOTYPES.items() if len(proto[1]) != arity[name]}
    assert not wrong, wrong


def test_environment_switches_are_documented():
    """Every SIGB_* switch the library reads is named in README.md, and README.md names no switch
    that nothing reads."""
    csrc = os.path.join(ROOT, "sigma_b200", "csrc")
    read = set()
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cpp", ".cuh", ".h")):
            read |= set(re.findall(r'(?:getenv|env_int)\("(SIGB_\w+)"', open(os.path.join(csrc, f)).read()))
    read |= set(re.findall(r'environ\.get\("(SIGB_\w+)"', open(os.path.join(ROOT, "sigma_b200", "_capi.py")).read()))
    readme = open(os.path.join(ROOT, "README.md")).read()
    named = set(re.findall(r"`(SIGB_[A-Z0-9_]+)", readme))
    # families written as SIGB_LDU_SF_* and the test-only / build-only names
    named_prefixes = {m[:-1] for m in re.findall(r"`(SIGB_[A-Z0-9_]+\*)", readme.replace("_\\*", "_*"))}
    undocumented = {s for s in read if s not in named and not any(s.startswith(p) for p in named_prefixes)}
    assert not undocumented, undocumented
    build_or_test_only = {"SIGB_PHASE_TIMERS", "SIGB_PERSIST_MINBLOCKS", "SIGB_TILE_NNZ", "SIGB_TILE_ROWS",
                          "SIGB_ERR_COMM"}     # (the last one is a status code, not a switch)
    stale = {s for s in named if s not in read and s not in build_or_test_only and not s.endswith("_")}
    assert not stale, stale


def test_partition_and_halo_on_random_patterns_hypothesis(orc):
    """Random rectangular-free patterns (empty rows, full rows, any number of parts up to the row
    count): partition offsets, halo lists and local renumbering equal the oracle's, entry for entry."""
    from hypothesis import given, settings, strategies as st

    from sigma_b200 import _capi

    L = _capi.lib()

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 60), st.floats(0.0, 0.5), st.integers(1, 8), st.integers(0, 2**31 - 1))
    def check(n, density, P, seed):
        rng = np.random.default_rng(seed)
        mask = rng.random((n, n)) < density
        rows = [rng.permutation(np.flatnonzero(mask[i]) + 1) for i in range(n)]
        ptr = np.concatenate([[1], 1 + np.cumsum([r.size for r in rows])]).astype(np.int32)
        node = (np.concatenate(rows) if ptr[-1] > 1 else np.zeros(0)).astype(np.int32)
        part = np.empty(P + 1, np.int32)
        _capi.check(L.sigb_partition_rows(n, _capi.ptr(ptr), P, _capi.ptr(part)))
        assert np.array_equal(part, orc.partition_rows(ptr, P))
        assert part[0] == 0 and part[-1] == n and np.all(np.diff(part) >= 0)
        for r in range(P):
            lo, hi = int(part[r]), int(part[r + 1])
            bptr = np.ascontiguousarray(ptr[lo:hi + 1])
            bnode = np.ascontiguousarray(node[ptr[lo] - 1: ptr[hi] - 1])
            halo = np.empty(max(bnode.size, 1), np.int32)
            local = np.empty(max(bnode.size, 1), np.int32)
            nh = C.c_int32()
            _capi.check(L.sigb_halo_build(lo, hi, _capi.ptr(bptr), _capi.ptr(bnode) if bnode.size else _capi.ptr(halo),
                                          _capi.ptr(halo), C.byref(nh), _capi.ptr(local)))
            ohalo, olocal = orc.halo_build(lo, hi, ptr, node)
            assert np.array_equal(halo[: nh.value], ohalo)
            assert np.array_equal(local[: bnode.size], olocal)

    check()
