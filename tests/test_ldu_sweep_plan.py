"""Statically scheduled ILDU(0) sweeps (sigma_b200/csrc/ldu_sweep.h, plan built by ldu_host.cpp; no GPU):
the plan's arrays are fed to a numpy replica of the device kernel -- one "thread" per chunk, lock-step
trips, a ring of the last W positions of every chunk, far values from the trip-ordered solution -- in
which every read is CHECKED to hit the value the serial loop would read (ring slots and trip-ordered
entries carry the (row) they hold), and the result must equal the serial triangular solves of
ldu_solvers.f90:226-235 / :254-263 bit for bit."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from sigma_b200 import generators as G
from sigma_b200._capi import check, lib, ptr


def symbolic(n, p, nd):
    p, nd = np.ascontiguousarray(p, np.int32), np.ascontiguousarray(nd, np.int32)
    ne = nd.size
    Lptr, Uptr = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32)
    Lnode, Unode = np.zeros(max(ne, 1), np.int32), np.zeros(max(ne, 1), np.int32)
    dest = np.zeros(max(ne, 1), np.int64)
    fr, br = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
    fl, bl = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32)
    nf, nb = C.c_int32(), C.c_int32()
    check(lib().sigb_ldu_symbolic(n, ptr(p), ptr(nd), ptr(Lptr), ptr(Lnode), ptr(Uptr), ptr(Unode), ptr(dest), ptr(fr),
                                  ptr(fl), C.byref(nf), ptr(br), ptr(bl), C.byref(nb)))
    return (Lptr, Lnode[: Lptr[n] - 1].copy(), nf.value), (Uptr, Unode[: Uptr[n] - 1].copy(), nb.value)


def sweep_plan(n, p, nd, backward, levels):
    p, nd = np.ascontiguousarray(p, np.int32), np.ascontiguousarray(nd if nd.size else np.zeros(1), np.int32)
    info = np.zeros(16, np.int32)
    check(lib().sigb_debug_ldu_sweep_plan(n, ptr(p), ptr(nd), backward, levels, ptr(info), None, None, None, None))
    keys = ["eligible", "R", "sigma", "C", "trips", "W", "S_max", "w16_max", "nstage", "stage_bytes", "threads"]
    P = dict(zip(keys, (int(v) for v in info[:11])))
    P["total"] = int(np.uint32(info[11])) | (int(info[12]) << 32)
    P["total_s"] = int(np.uint32(info[13])) | (int(info[14]) << 32)
    P["has_far"] = int(info[15])
    if not P["eligible"]:
        return P
    trip = np.zeros((P["trips"], 8), np.int32)
    src = np.zeros(max(P["total_s"], 1), np.int32)
    cnt = np.zeros(max(P["total"], 1), np.uint8)
    valmap = np.zeros(max(P["total_s"], 1), np.int64)
    check(lib().sigb_debug_ldu_sweep_plan(n, ptr(p), ptr(nd), backward, levels, ptr(info), ptr(trip), ptr(src), ptr(cnt),
                                          ptr(valmap)))
    P.update(trip=trip, src=src, cnt=cnt, valmap=valmap)
    return P


def serial_sweep(n, p, nd, val, rhs, backward):
    """lower_/upper_triangular_solve: z = x(i); z = z - M%val(k) * x(node(k)) in stored order; x(i) = z"""
    x = rhs.copy()
    rows = range(n - 1, -1, -1) if backward else range(n)
    for i in rows:
        z = x[i]
        for k in range(p[i] - 1, p[i + 1] - 1):
            z = z - val[k] * x[nd[k] - 1]
        x[i] = z
    return x


def replay(n, p, nd, val, rhs, backward, P):
    """the device side: ldu_sched_in_kernel, ldu_sweep_static_kernel, ldu_sched_out_kernel (csrc/ldu.cu)"""
    R, sg, Cn, W = P["R"], P["sigma"], P["C"], P["W"]
    trip, src, cnt, valmap = P["trip"], P["src"], P["cnt"], P["valmap"]
    row_of = (lambda q: n - 1 - q) if backward else (lambda q: q)           # 0-based row at sweep position q
    pos_of = (lambda i0: n - 1 - i0) if backward else (lambda i0: i0)
    total = P["total"]
    rhs_s = np.full(total, np.nan)
    for t in range(P["trips"]):
        vlo, w, w16, S = (int(a) for a in trip[t, :4])
        off = int(np.uint32(trip[t, 4])) | (int(trip[t, 5]) << 32)
        for u in range(w):
            v = vlo + u
            q = v * R + (t - sg * v)
            assert 0 <= t - sg * v < R
            if q < n:
                rhs_s[off + u] = rhs[row_of(q)]
    xs = np.full(total, np.nan)
    xs_row = np.full(total, -1, np.int64)
    xs_trip = np.full(total, -1, np.int64)
    far_seen = False
    ring = np.full(W * Cn, np.nan)
    ring_row = np.full(W * Cn, -1, np.int64)
    seen = np.zeros(n, bool)
    for t in range(P["trips"]):
        vlo, w, w16, S = (int(a) for a in trip[t, :4])
        off = int(np.uint32(trip[t, 4])) | (int(trip[t, 5]) << 32)
        soff = int(np.uint32(trip[t, 6])) | (int(trip[t, 7]) << 32)
        assert w16 % 16 == 0 and w16 >= w and off % 16 == 0 and soff % 16 == 0
        out = []
        for u in range(w):                       # all threads read ...
            c = int(cnt[off + u])
            v = vlo + u
            pp = t - sg * v
            q = v * R + pp
            if c == 0xFF:
                assert q >= n
                continue
            i0 = row_of(q)
            assert c == p[i0 + 1] - p[i0] and c <= S
            z = rhs_s[off + u]
            for s in range(c):
                slot = soff + s * w16 + u
                k = int(valmap[slot])
                assert k == p[i0] - 1 + s                       # stored order
                j0 = nd[k] - 1
                d = int(src[slot])
                if d >= 0:
                    assert ring_row[d] == j0, "ring slot does not hold the entry the serial loop reads"
                    xj = ring[d]
                else:
                    assert xs_row[-d - 1] == j0, "trip-ordered entry does not hold the entry the serial loop reads"
                    assert t - xs_trip[-d - 1] >= W, "far entry within reach of the ring"
                    far_seen = True
                    xj = xs[-d - 1]
                z = z - val[k] * xj
            out.append((u, v, pp, i0, z))
        for u, v, pp, i0, z in out:              # ... then all write (one barrier per trip)
            xs[off + u] = z
            xs_row[off + u] = i0
            xs_trip[off + u] = t
            ring[(pp % W) * Cn + v] = z
            ring_row[(pp % W) * Cn + v] = i0
            assert not seen[i0]
            seen[i0] = True
    assert seen.all() and far_seen == bool(P["has_far"])
    x = np.empty(n)
    for t in range(P["trips"]):
        vlo, w = int(trip[t, 0]), int(trip[t, 1])
        off = int(np.uint32(trip[t, 4])) | (int(trip[t, 5]) << 32)
        for u in range(w):
            v = vlo + u
            q = v * R + (t - sg * v)
            if q < n:
                x[row_of(q)] = xs[off + u]
    return x


def banded(n, band, fill, seed):
    """random pattern with |i - j| <= band, unsorted rows, a diagonal everywhere"""
    rng = np.random.default_rng(seed)
    p, nd = [1], []
    for i in range(n):
        lo, hi = max(0, i - band), min(n - 1, i + band)
        cand = np.setdiff1d(np.arange(lo, hi + 1), [i])
        pick = cand[rng.random(cand.size) < fill]
        cols = rng.permutation(np.concatenate([pick, [i]])) + 1
        nd.append(cols)
        p.append(p[-1] + cols.size)
    return np.array(p, np.int32), np.concatenate(nd).astype(np.int32)


def stencil(N, offsets):
    """N x N grid in natural ordering, row (iy, ix) coupled to (iy + dy, ix + dx) for the given offsets"""
    p, nd = [1], []
    for iy in range(N):
        for ix in range(N):
            cols = [(iy + dy) * N + ix + dx + 1 for dy, dx in offsets if 0 <= iy + dy < N and 0 <= ix + dx < N]
            nd.extend(cols)
            p.append(p[-1] + len(cols))
    return np.array(p, np.int32), np.array(nd, np.int32)


def cases():
    yield "poisson33", 33 * 33, *G.poisson2d_csr(33)[:2], 1
    yield "poisson50", 50 * 50, *G.poisson2d_csr(50)[:2], 1
    yield "fem40", 40 * 40, *G.fem_p1_csr(40)[:2], None
    yield "tridiag", 6000, *G.tridiag_csr(6000)[:2], None
    yield "stencil9", 45 * 45, *stencil(45, [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 0), (0, 1), (1, -1), (1, 0), (1, 1)]), 2
    yield "stencil_skew3", 40 * 40, *stencil(40, [(-1, 0), (-1, 2), (0, -1), (0, 0), (0, 1), (1, -2), (1, 0)]), 3
    # dense-ish random bands are nearly serial (sigma ~ the band): eligible or not, never wrong
    yield "band12", 3000, *banded(3000, 12, 0.3, 1), 0
    yield "band40_sparse", 4000, *banded(4000, 40, 0.05, 2), 0


@pytest.mark.parametrize("case", list(cases()), ids=lambda c: c[0])
def test_static_sweeps_equal_the_serial_solves(case):
    name, n, p, nd, want_sigma = case
    (Lp, Ln, nf), (Up, Un, nb) = symbolic(n, p, nd)
    rng = np.random.default_rng(5)
    for backward, (tp, tn, lev) in ((0, (Lp, Ln, nf)), (1, (Up, Un, nb))):
        P = sweep_plan(n, tp, tn, backward, lev)
        if want_sigma == 0 and not P["eligible"]:
            continue
        assert P["eligible"], (name, backward, P)
        if want_sigma:
            assert P["sigma"] == want_sigma
        assert P["trips"] == P["R"] + P["sigma"] * (P["C"] - 1)
        assert 2 <= P["W"] <= 8 and 2 <= P["nstage"] <= 4
        assert P["nstage"] * P["stage_bytes"] + P["W"] * P["C"] * 8 <= 216 * 1024
        val = rng.standard_normal(tn.size)
        rhs = rng.standard_normal(n)
        x = replay(n, tp, tn, val, rhs, backward, P)
        assert np.array_equal(x, serial_sweep(n, tp, tn, val, rhs, backward))


def test_poisson_is_the_classic_wavefront():
    N = 64
    (Lp, Ln, nf), (Up, Un, nb) = symbolic(N * N, *G.poisson2d_csr(N)[:2])
    for backward, (tp, tn, lev) in ((0, (Lp, Ln, nf)), (1, (Up, Un, nb))):
        P = sweep_plan(N * N, tp, tn, backward, lev)
        assert (P["eligible"], P["R"], P["C"], P["sigma"], P["trips"], P["W"], P["S_max"]) == (1, N, N, 1, 2 * N - 1, 2, 2)
        assert lev == 2 * N - 1


def test_plans_that_are_not_worth_it_are_refused():
    # a random graph has no band: one chunk, no wavefront
    n = 3000
    p, nd, _ = G.erdos_renyi_csr(n, seed=3, weights="random")
    (Lp, Ln, nf), _ = symbolic(n, p, nd)
    assert not sweep_plan(n, Lp, Ln, 0, nf)["eligible"]
    # a shallow schedule (block-diagonal: 30 levels) stays with the level launches
    p, nd = banded(30, 30, 1.0, 0)
    reps = 200
    pp = np.concatenate([[1], 1 + np.cumsum(np.tile(np.diff(p), reps))]).astype(np.int32)
    nn = np.concatenate([nd + 30 * r for r in range(reps)]).astype(np.int32)
    (Lp, Ln, nf), _ = symbolic(30 * reps, pp, nn)
    assert nf == 30 and not sweep_plan(30 * reps, Lp, Ln, 0, nf)["eligible"]
    # too small
    (Lp, Ln, nf), _ = symbolic(20 * 20, *G.poisson2d_csr(20)[:2])
    assert not sweep_plan(400, Lp, Ln, 0, nf)["eligible"]


@settings(max_examples=25, deadline=None)
@given(st.integers(200, 1500), st.integers(1, 30), st.floats(0.02, 0.6), st.integers(0, 10**6))
def test_fuzz_banded_patterns(n, band, fill, seed):
    p, nd = banded(n, band, fill, seed)
    (Lp, Ln, nf), (Up, Un, nb) = symbolic(n, p, nd)
    rng = np.random.default_rng(seed)
    for backward, (tp, tn, lev) in ((0, (Lp, Ln, nf)), (1, (Up, Un, nb))):
        P = sweep_plan(n, tp, tn, backward, 10**6)      # depth test waived: exercise the schedule itself
        if not P["eligible"]:
            continue
        val = rng.standard_normal(tn.size)
        rhs = rng.standard_normal(n)
        assert np.array_equal(replay(n, tp, tn, val, rhs, backward, P), serial_sweep(n, tp, tn, val, rhs, backward))
