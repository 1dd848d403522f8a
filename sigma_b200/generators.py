"""Deterministic synthetic inputs for the hot path, in the reference's storage.

Every generator returns arrays laid out exactly as the reference would hold
them after building the matrix the way its tests do: build an ``ll_graph`` with
``add_edge`` (insertion order kept per row, src/graph/formats/ll_graphs.f90:355-371),
convert to compressed-sparse / ellpack (first-free-slot insertion,
src/graph/formats/cs_graphs.f90:163-183; last-neighbour padding,
src/graph/formats/ellpack_graphs.f90:164), then ``set_value``.  Columns are
therefore NOT sorted in general.  The large cases are produced with vectorised
numpy that yields the same arrays as that build order; tests/test_generators.py
proves the equality against the oracle's restatement of the builders on small
sizes.

Indices are 1-based int32, values fp64.  Nothing here touches the GPU or the
oracle.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "tridiag_add_edge_calls", "tridiag_csr", "tridiag_ell", "poisson2d_add_edge_calls",
    "poisson2d_csr", "poisson2d_rhs", "csr_to_ell", "csr_transpose", "erdos_renyi_csr",
    "erdos_renyi_add_edge_calls", "fem_p1_csr", "dense_from_csr", "periodic_p1_grid",
]


# --------------------------------------------------------------------------
# 1-D three-point operators (test/solver_test_diffusion_1d.f90:55-76,
# test/solver_test_advection_diffusion_1d.f90:58-82)
# --------------------------------------------------------------------------
def tridiag_add_edge_calls(nn):
    """The (i, j) sequence of g%add_edge calls of both 1-D tests (:60-65)."""
    i = np.arange(1, nn, dtype=np.int32)
    ei = np.stack([i, i, i + 1], 1).reshape(-1)
    ej = np.stack([i, i + 1, i], 1).reshape(-1)
    return np.append(ei, np.int32(nn)), np.append(ej, np.int32(nn))


def tridiag_csr(nn, diag=2.0, upper=-1.0, lower=-1.0):
    """CSR arrays of tridiag(lower, diag, upper): row i holds [i-1, i, i+1]."""
    k = np.arange(1, nn + 1, dtype=np.int32)
    cand = np.stack([k - 1, k, k + 1], 1)
    vals = np.broadcast_to(np.array([lower, diag, upper]), (nn, 3))
    valid = np.stack([k > 1, np.ones(nn, bool), k < nn], 1)
    ptr = np.concatenate([[1], 1 + np.cumsum(valid.sum(1))]).astype(np.int32)
    return ptr, cand[valid].astype(np.int32), vals[valid].astype(np.float64)


def tridiag_ell(nn, diag=2.0, upper=-1.0, lower=-1.0):
    """ELLPACK arrays (node[nn,3], degrees, val[nn,3]) of the same operator.
    Row 1 = [1,2,2], row nn = [nn-1,nn,nn]; padding values are 0."""
    ptr, node, val = tridiag_csr(nn, diag, upper, lower)
    return csr_to_ell(ptr, node, val)


# --------------------------------------------------------------------------
# 2-D five-point Poisson on an N x N grid, Dirichlet (BASELINE config 2)
# --------------------------------------------------------------------------
def poisson2d_add_edge_calls(N):
    """add_edge sequence in the reference idiom (cf. apps/regular_graphs.f90:22-34,
    test/solver_test_diffusion_1d.f90:60-65): loop vertices k = N*(ix-1)+iy,
    add (k,k), then (k,k+1),(k+1,k) if iy<N, then (k,k+N),(k+N,k) if ix<N."""
    ei, ej = [], []
    for ix in range(1, N + 1):
        for iy in range(1, N + 1):
            k = N * (ix - 1) + iy
            ei.append(k); ej.append(k)
            if iy < N:
                ei += [k, k + 1]; ej += [k + 1, k]
            if ix < N:
                ei += [k, k + N]; ej += [k + N, k]
    return np.array(ei, np.int32), np.array(ej, np.int32)


def poisson2d_csr(N, row_lo=0, row_hi=None, diag=4.0, off=-1.0):
    """Rows [row_lo, row_hi) (0-based) of the N^2 x N^2 five-point matrix.

    Interior row k (1-based) is stored as [k-N, k-1, k, k+1, k+N] -- the order
    the build above produces.  Returns (ptr, node, val) with ptr 1-based and
    starting at 1 for the block, node holding GLOBAL 1-based column ids.
    """
    n = N * N
    row_hi = n if row_hi is None else row_hi
    k0 = np.arange(row_lo, row_hi, dtype=np.int64)
    ix, iy = k0 // N, k0 % N
    cand = np.stack([k0 - N, k0 - 1, k0, k0 + 1, k0 + N], 1) + 1
    valid = np.stack([ix > 0, iy > 0, np.ones(k0.size, bool), iy < N - 1, ix < N - 1], 1)
    vals = np.broadcast_to(np.array([off, off, diag, off, off]), cand.shape)
    ptr = np.concatenate([[1], 1 + np.cumsum(valid.sum(1))]).astype(np.int32)
    return ptr, cand[valid].astype(np.int32), vals[valid].astype(np.float64)


def poisson2d_rhs(N, seed=12345, row_lo=0, row_hi=None, diag=4.0, off=-1.0):
    """b = A x*, x* ~ U[0,1) from PCG64(seed); returns
    (b[row_lo:row_hi], x*[row_lo:row_hi]).  b is evaluated with the stencil in
    the stored entry order, without FMA (numpy never contracts)."""
    n = N * N
    row_hi = n if row_hi is None else row_hi
    xs = np.random.Generator(np.random.PCG64(seed)).random(n)
    g = xs.reshape(N, N)
    z = np.zeros((N, N))
    z[1:, :] += off * g[:-1, :]       # k - N
    z[:, 1:] += off * g[:, :-1]       # k - 1
    z += diag * g                     # k
    z[:, :-1] += off * g[:, 1:]       # k + 1
    z[:-1, :] += off * g[1:, :]       # k + N
    return z.reshape(-1)[row_lo:row_hi].copy(), xs[row_lo:row_hi].copy()


# --------------------------------------------------------------------------
# format conversions in the reference's own semantics
# --------------------------------------------------------------------------
def csr_to_ell(ptr, node, val):
    """convert_graph_type(g, "ellpack") + values: slots in stored order, the
    rest of each row filled with the row's last neighbour
    (ellpack_graphs.f90:164) and val = 0 (ellpack_matrices.f90:132)."""
    ptr = np.asarray(ptr, np.int64)
    n = ptr.size - 1
    deg = np.diff(ptr).astype(np.int32)
    w = int(deg.max()) if n else 0
    slot = np.arange(w)[None, :]
    idx = (ptr[:-1, None] - 1) + np.minimum(slot, np.maximum(deg[:, None] - 1, 0))
    ell_node = np.asarray(node)[idx].astype(np.int32)
    ell_val = np.where(slot < deg[:, None], np.asarray(val)[idx], 0.0)
    if (deg == 0).any():
        ell_node[deg == 0] = 0
    return ell_node, deg, ell_val.astype(np.float64)


def csr_transpose(n, m, ptr, node, val):
    """Arrays of the transposed copy the reference builds (cs_graph_build with
    trans, cs_graphs.f90:122-183): line j holds the source rows in ascending
    order.  This is also the CSC storage of the same matrix."""
    ptr = np.asarray(ptr, np.int64)
    rows = np.repeat(np.arange(1, n + 1, dtype=np.int32), np.diff(ptr))
    order = np.argsort(np.asarray(node), kind="stable")
    cnt = np.bincount(np.asarray(node) - 1, minlength=m)
    ptr_t = np.concatenate([[1], 1 + np.cumsum(cnt)]).astype(np.int32)
    return ptr_t, rows[order].astype(np.int32), np.asarray(val)[order].astype(np.float64)


def dense_from_csr(n, m, ptr, node, val):
    B = np.zeros((n, m))
    rows = np.repeat(np.arange(n), np.diff(ptr))
    B[rows, np.asarray(node) - 1] = val
    return B


# --------------------------------------------------------------------------
# Erdos-Renyi graph Laplacians (test/solver_test_jacobi.f90:60-128,240-254,
# test/eigensolver_test_lanczos.f90:58-110, apps/random_graphs.f90:33-42)
# --------------------------------------------------------------------------
def _er_pairs(n, p, rng):
    """Upper-triangle pairs (i<j) of G(n, p) by geometric skipping (the
    reference's O(n^2) double loop is infeasible at 2e7 vertices)."""
    total = n * (n - 1) // 2
    expected = total * p
    chunk = int(expected * 1.1 + 1000)
    pos = []
    cur = -1
    while True:
        gaps = rng.geometric(p, size=chunk).astype(np.int64)
        cs = cur + np.cumsum(gaps)
        keep = cs < total
        pos.append(cs[keep])
        if not keep.all():
            break
        cur = int(cs[-1])
    pos = np.concatenate(pos)
    # linear index -> (i, j), rows enumerated i = 0..n-2 with n-1-i entries each
    # i = floor(((2n-1) - sqrt((2n-1)^2 - 8 pos)) / 2)
    b = 2.0 * n - 1.0
    i = np.floor((b - np.sqrt(b * b - 8.0 * pos)) / 2.0).astype(np.int64)
    start = i * (2 * n - i - 1) // 2
    fix = pos < start
    i[fix] -= 1
    start = i * (2 * n - i - 1) // 2
    fix = pos >= start + (n - 1 - i)
    i[fix] += 1
    start = i * (2 * n - i - 1) // 2
    j = pos - start + i + 1
    return i, j


def erdos_renyi_add_edge_calls(n, i, j):
    """add_edge sequence of the reference tests for given upper pairs (0-based
    i<j, sorted by (i, j)): for each i: (i,i), then (i,j),(j,i) per neighbour."""
    ei, ej = [], []
    order = np.lexsort((j, i))
    i, j = i[order], j[order]
    k = 0
    for v in range(n):
        ei.append(v + 1); ej.append(v + 1)
        while k < i.size and i[k] == v:
            ei += [v + 1, j[k] + 1]; ej += [j[k] + 1, v + 1]
            k += 1
    return np.array(ei, np.int32), np.array(ej, np.int32)


def erdos_renyi_csr(n, p=None, seed=7, shift=1.0, skew=False, weights="unit", return_pairs=False):
    """A = L + shift*I on G(n, p) (p defaults to log2(n)/n).

    weights="unit": L = D - Adj (test/eigensolver_test_lanczos.f90:100-110);
    weights="random": edge weight z ~ U[0,1): A(i,j) = -z, A(i,i) += z
    (test/solver_test_jacobi.f90:111-128, diagonal starts at `shift`).
    skew=True adds the skew-symmetric perturbation +-(2z-1)/16 per edge
    (:240-254) -> nonsymmetric operator for BiCGSTAB.
    The build order of the tests makes every row's columns ascending
    (smaller neighbours, the diagonal, larger neighbours).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    if p is None:
        p = np.log2(n) / n
    i, j = _er_pairs(n, p, rng)
    ne = i.size
    w = np.ones(ne) if weights == "unit" else rng.random(ne)
    rows = np.concatenate([i, j, np.arange(n)])
    cols = np.concatenate([j, i, np.arange(n)])
    deg_w = np.bincount(i, w, n) + np.bincount(j, w, n)
    vals = np.concatenate([-w, -w, shift + deg_w])
    if skew:
        s = (2.0 * rng.random(ne) - 1.0) / 16.0
        vals[:ne] += s
        vals[ne:2 * ne] -= s
    order = np.argsort(rows * np.int64(n) + cols)     # (row, col) pairs are distinct: same order as a lexsort, 4x faster
    rows, cols, vals = rows[order], cols[order], vals[order]
    ptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(rows, minlength=n))]).astype(np.int32)
    out = (ptr, (cols + 1).astype(np.int32), vals.astype(np.float64))
    return out + ((i, j),) if return_pairs else out


def erdos_renyi_csr_rows(n, row_lo, row_hi, p=None, seed=7, shift=1.0, skew=False, weights="unit", cache=None):
    """Rows [row_lo, row_hi) (0-based) of erdos_renyi_csr(n, ...): the same arrays as slicing
    the whole matrix (ptr rebased to 1), without sorting the entries of other rows.  Used by the
    row-sharded runs of BASELINE config 5, where every rank draws the same pair list (same
    seed, same draws in the same order) and keeps its own block.  Also returns the number of
    stored entries of every row of the whole matrix (int64), from which the partition is derived.
    cache: a dict that keeps the pair list (and the generator state right after it) between
    calls with the same (n, p, seed) -- the variants of one graph (unit / random weights,
    skew) differ only in what is drawn AFTER the pairs.
    """
    if p is None:
        p = np.log2(n) / n
    key = ("er_pairs", n, float(p), seed)
    if cache is not None and key in cache:
        i, j, state = cache[key]
        bitgen = np.random.PCG64(seed)
        bitgen.state = state
        rng = np.random.Generator(bitgen)
    else:
        rng = np.random.Generator(np.random.PCG64(seed))
        i, j = _er_pairs(n, p, rng)
        if cache is not None:
            cache[key] = (i, j, rng.bit_generator.state)
    ne = i.size
    w = np.ones(ne) if weights == "unit" else rng.random(ne)
    deg_w = np.bincount(i, w, n) + np.bincount(j, w, n)
    s = (2.0 * rng.random(ne) - 1.0) / 16.0 if skew else None
    counts = np.bincount(i, minlength=n) + np.bincount(j, minlength=n) + 1
    up = (i >= row_lo) & (i < row_hi)        # entries (i, j) of owned rows i
    lo_ = (j >= row_lo) & (j < row_hi)       # entries (j, i) of owned rows j
    d = np.arange(row_lo, row_hi)
    rows = np.concatenate([i[up], j[lo_], d])
    cols = np.concatenate([j[up], i[lo_], d])
    vu, vl = -w[up], -w[lo_]
    if skew:
        vu = vu + s[up]
        vl = vl - s[lo_]
    vals = np.concatenate([vu, vl, shift + deg_w[row_lo:row_hi]])
    order = np.argsort(rows * np.int64(n) + cols)
    rows, cols, vals = rows[order], cols[order], vals[order]
    ptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(rows - row_lo, minlength=row_hi - row_lo))]).astype(np.int32)
    return ptr, (cols + 1).astype(np.int32), vals.astype(np.float64), counts


# --------------------------------------------------------------------------
# Two-field block system of test/matrix_test_composite.f90:104-219
# --------------------------------------------------------------------------
def _er_laplacian_blocks(n, p, rng):
    """erdos_renyi_graph(g, n, n, p, symmetric=.true.) (:560-590) followed by
    erdos_renyi_matrix (:595-620): every vertex carries its self-edge; for each
    stored edge (i, j): A(i,j) += -1, A(i,i) += +1.  Rows come out ascending
    (smaller neighbours, the diagonal, larger neighbours)."""
    upper = np.triu(rng.random((n, n)) < p, k=1)
    adj = upper | upper.T | np.eye(n, dtype=bool)
    rows, cols = np.nonzero(adj)                       # row-major => ascending columns
    deg = adj.sum(axis=1)
    val = np.where(rows == cols, deg[rows] - 1.0, -1.0)
    ptr = np.concatenate([[1], 1 + np.cumsum(deg)]).astype(np.int32)
    return ptr, (cols + 1).astype(np.int32), val.astype(np.float64), adj


def composite_er_blocks(nn1=768, nn2=512, seed=11):
    """The 2 x 2 block matrix of test/matrix_test_composite.f90: random weighted
    Laplacians on the diagonal (CSR), the coupling graph h (nn1 x nn2, p = 6/nn1,
    only j > i, :171-174) as the CSR (1,2) block and -- the SAME graph object --
    as the CSC (2,1) block, coupling values -1 and +1 on both diagonals per
    coupling edge (:204-217).

    Returns dict(b11=(ptr, node, val), b22=(...), h=(ptr, node), v12, v21,
    adj1, adj2, adjh) with 1-based int32 index arrays."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p1 = np.log2(nn1) / nn1
    p2 = np.log2(nn2) / nn2
    ptr1, node1, val1, adj1 = _er_laplacian_blocks(nn1, p1, rng)
    ptr2, node2, val2, adj2 = _er_laplacian_blocks(nn2, p2, rng)
    adjh = rng.random((nn1, nn2)) < 6.0 / nn1
    adjh &= np.arange(nn2)[None, :] > np.arange(nn1)[:, None]
    hr, hc = np.nonzero(adjh)
    hdeg, htdeg = adjh.sum(axis=1), adjh.sum(axis=0)
    ptrh = np.concatenate([[1], 1 + np.cumsum(hdeg)]).astype(np.int32)
    nodeh = (hc + 1).astype(np.int32)
    # A%add(1, 1, i, i, +1) and A%add(2, 2, j, j, +1) once per coupling edge
    r1 = np.repeat(np.arange(nn1), np.diff(ptr1))
    val1 = val1 + np.where(r1 == node1 - 1, hdeg[r1].astype(np.float64), 0.0)
    r2 = np.repeat(np.arange(nn2), np.diff(ptr2))
    val2 = val2 + np.where(r2 == node2 - 1, htdeg[r2].astype(np.float64), 0.0)
    ne_h = nodeh.size
    return dict(b11=(ptr1, node1, val1), b22=(ptr2, node2, val2), h=(ptrh, nodeh),
                v12=np.full(ne_h, -1.0), v21=np.full(ne_h, -1.0), adj1=adj1, adj2=adj2, adjh=adjh)


# --------------------------------------------------------------------------
# P1 finite-element Laplacian on a perturbed structured triangulation
# (BASELINE config 4; element matrices per examples/fem.f90:28-49)
# --------------------------------------------------------------------------
def fem_p1_add_value_stream(N, seed=2024, jitter=0.25):
    """The A%add_value(ele(i,n), ele(j,n), AE(i,j)) calls of laplacian2d
    (examples/fem.f90:28-49) on an N x N vertex grid, each cell split into two
    triangles along a seeded random diagonal, interior vertices jittered by
    <= jitter*h.  Returns (I, J, V, interior) -- 0-based int64 vertex ids and the
    element-matrix entries, in call order: element n, j outer, i inner (:43-47).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    h = 1.0 / (N - 1)
    gx, gy = np.meshgrid(np.arange(N) * h, np.arange(N) * h, indexing="ij")
    interior = np.zeros((N, N), bool)
    interior[1:-1, 1:-1] = True
    gx = gx + np.where(interior, (2 * rng.random((N, N)) - 1) * jitter * h, 0.0)
    gy = gy + np.where(interior, (2 * rng.random((N, N)) - 1) * jitter * h, 0.0)
    x = np.stack([gx.reshape(-1), gy.reshape(-1)])          # x(2, nv)
    vid = np.arange(N * N).reshape(N, N)
    a, b, c, d = vid[:-1, :-1].ravel(), vid[1:, :-1].ravel(), vid[1:, 1:].ravel(), vid[:-1, 1:].ravel()
    flip = rng.random(a.size) < 0.5
    t1 = np.where(flip[:, None], np.stack([a, b, d], 1), np.stack([a, b, c], 1))
    t2 = np.where(flip[:, None], np.stack([b, c, d], 1), np.stack([a, c, d], 1))
    ele = np.stack([t1, t2], 1).reshape(-1, 3)               # cell-major, 2 triangles each
    # element matrices (fem.f90:31-41)
    jj = ele[:, [1, 2, 0]]
    kk = ele[:, [2, 0, 1]]
    V1 = x[1, jj] - x[1, kk]
    V2 = x[0, kk] - x[0, jj]
    det = V1[:, 0] * V2[:, 1] - V2[:, 0] * V1[:, 1]
    area = np.abs(det) / 2.0
    AE = (0.25 / area)[:, None, None] * (V1[:, :, None] * V1[:, None, :] + V2[:, :, None] * V2[:, None, :])
    # entry stream in assembly order: element n, j outer, i inner
    I = ele[:, None, :].repeat(3, 1)        # [n, j, i] -> ele(i)
    J = ele[:, :, None].repeat(3, 2)        # [n, j, i] -> ele(j)
    Vv = AE.transpose(0, 2, 1)              # [n, j, i] -> AE(i, j)
    I, J, Vv = I.reshape(-1), J.reshape(-1), Vv.reshape(-1)
    return I, J, Vv, interior


def fem_p1_csr(N, seed=2024, jitter=0.25):
    """Stiffness matrix of -Laplace on an N x N vertex grid (see
    fem_p1_add_value_stream); Dirichlet rows/columns replaced by identity.

    Graph build order: loop elements n, add_edge(ele(i,n), ele(j,n)) for
    j = 1..3, i = 1..3 (the same nesting as the assembly loop fem.f90:43-47);
    values accumulated in that element order.  Returns (ptr, node, val).
    """
    I, J, Vv, interior = fem_p1_add_value_stream(N, seed, jitter)
    nv = N * N
    bnd = ~interior.reshape(-1)
    # first-occurrence order per row == ll_graph insertion order
    key = I.astype(np.int64) * nv + J
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    # rows ascending, then first appearance (one combined key: `first` values are distinct, so
    # this is the order a lexsort on (row, first) gives, several times faster)
    order = np.argsort((uniq // nv) * np.int64(key.size) + first)
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    vals = np.zeros(uniq.size)
    np.add.at(vals, rank[inv], Vv)                           # sequential, element order
    rows = (uniq // nv)[order]
    cols = (uniq % nv)[order]
    # Dirichlet: identity rows/columns on the boundary
    vals = np.where(bnd[rows] | bnd[cols], np.where(rows == cols, 1.0, 0.0), vals)
    ptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(rows, minlength=nv))]).astype(np.int32)
    return ptr, (cols + 1).astype(np.int32), vals.astype(np.float64)


def fem_p1_csr_rows(N, row_lo, row_hi, seed=2024, jitter=0.25):
    """Rows [row_lo, row_hi) (0-based vertex ids) of fem_p1_csr(N, ...): the same arrays as
    slicing the whole matrix (ptr rebased to 1), built from the elements that touch those
    vertices only.  The random draws (jitter, diagonal choice) are made for the whole grid in
    the same order as in fem_p1_add_value_stream, so the result does not depend on the
    sharding; the elements keep their global order, hence each row keeps its insertion order
    and every entry its accumulation order.  Row-sharded runs of BASELINE config 4.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    h = 1.0 / (N - 1)
    # cells (cx, cy) with cx in [c0, c1) touch vertex rows ix in [c0, c1]; owned ix: [row_lo // N, (row_hi - 1) // N]
    c0 = max(row_lo // N - 1, 0)
    c1 = min((row_hi - 1) // N + 1, N - 1)
    v0, v1 = c0, c1 + 1                                      # vertex rows needed: [v0, v1)
    interior = np.zeros((N, N), bool)
    interior[1:-1, 1:-1] = True
    jx = rng.random((N, N))[v0:v1]
    jy = rng.random((N, N))[v0:v1]
    flip = (rng.random((N - 1) * (N - 1)) < 0.5).reshape(N - 1, N - 1)[c0:c1].ravel()
    gx, gy = np.meshgrid(np.arange(v0, v1) * h, np.arange(N) * h, indexing="ij")
    inter = interior[v0:v1]
    gx = gx + np.where(inter, (2 * jx - 1) * jitter * h, 0.0)
    gy = gy + np.where(inter, (2 * jy - 1) * jitter * h, 0.0)
    x = np.stack([gx.reshape(-1), gy.reshape(-1)])          # local vertex numbering: (ix - v0) * N + iy
    vid = np.arange((v1 - v0) * N, dtype=np.int64).reshape(v1 - v0, N)
    a, b, c, d = vid[:-1, :-1].ravel(), vid[1:, :-1].ravel(), vid[1:, 1:].ravel(), vid[:-1, 1:].ravel()
    t1 = np.where(flip[:, None], np.stack([a, b, d], 1), np.stack([a, b, c], 1))
    t2 = np.where(flip[:, None], np.stack([b, c, d], 1), np.stack([a, c, d], 1))
    ele = np.stack([t1, t2], 1).reshape(-1, 3)
    jj = ele[:, [1, 2, 0]]
    kk = ele[:, [2, 0, 1]]
    V1 = x[1, jj] - x[1, kk]
    V2 = x[0, kk] - x[0, jj]
    det = V1[:, 0] * V2[:, 1] - V2[:, 0] * V1[:, 1]
    area = np.abs(det) / 2.0
    AE = (0.25 / area)[:, None, None] * (V1[:, :, None] * V1[:, None, :] + V2[:, :, None] * V2[:, None, :])
    off = v0 * N
    I = (ele[:, None, :].repeat(3, 1)).reshape(-1) + off
    J = (ele[:, :, None].repeat(3, 2)).reshape(-1) + off
    Vv = AE.transpose(0, 2, 1).reshape(-1)
    keep = (I >= row_lo) & (I < row_hi)
    I, J, Vv = I[keep], J[keep], Vv[keep]
    nv = N * N
    bnd = ~interior.reshape(-1)
    key = I * nv + J
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort((uniq // nv) * np.int64(key.size) + first)
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    vals = np.zeros(uniq.size)
    np.add.at(vals, rank[inv], Vv)
    rows = (uniq // nv)[order]
    cols = (uniq % nv)[order]
    vals = np.where(bnd[rows] | bnd[cols], np.where(rows == cols, 1.0, 0.0), vals)
    ptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(rows - row_lo, minlength=row_hi - row_lo))]).astype(np.int32)
    return ptr, (cols + 1).astype(np.int32), vals.astype(np.float64)


# --------------------------------------------------------------------------
# Stiffness and mass matrices of the generalized-Lanczos test
# (test/eigensolver_test_generalized_lanczos.f90:59-133): periodic nx x ny grid of
# right triangles, built with the test's own add_edge / A%add call order.
# --------------------------------------------------------------------------
def periodic_p1_grid(nx=48, ny=32):
    """Returns (ptr, node, valA, valB): both matrices share the CSR pattern."""
    nn = nx * ny

    def indx(i, j):          # 1-based (i, j) -> 1-based vertex, :203-209
        return ny * (j - 1) + i

    lists = [[] for _ in range(nn)]

    def add_edge(a, b):
        if b not in lists[a - 1]:
            lists[a - 1].append(b)

    for i in range(1, ny + 1):
        for j in range(1, nx + 1):
            k = indx(i, j)
            add_edge(k, k)
            for l in (indx(i % ny + 1, j), indx(i, j % nx + 1), indx(i % ny + 1, j % nx + 1)):
                add_edge(k, l)
                add_edge(l, k)
    ptr = np.concatenate([[1], 1 + np.cumsum([len(l) for l in lists])]).astype(np.int32)
    node = np.array([c for l in lists for c in l], np.int32)
    pos = [{c: ptr[r] - 1 + t for t, c in enumerate(l)} for r, l in enumerate(lists)]
    area = 0.5
    BE = np.full((3, 3), area / 12.0)
    BE[np.arange(3), np.arange(3)] = area / 6.0
    AE = np.array([[area, -area, 0.0], [-area, 2 * area, -area], [0.0, -area, area]])
    valA, valB = np.zeros(node.size), np.zeros(node.size)

    def add(elem):           # A%add(elem, elem, AE): k outer, l inner (cs_matrices.f90:934-967)
        for k in range(3):
            for l in range(3):
                p = pos[elem[k] - 1][elem[l]]
                valA[p] += AE[k, l]
                valB[p] += BE[k, l]

    for i in range(1, ny + 1):
        for j in range(1, nx + 1):
            e = [indx(i, j), indx(i, j % nx + 1), indx(i % ny + 1, j % nx + 1)]
            add(e)
            e[1] = indx(i % ny + 1, j)
            add(e)
    return ptr, node, valA, valB
