// convert.cu -- on-device matrix copy / format conversion:
// the step before the hot path.
//
// Replaces the bodies of
//   cs_matrix_copy_matrix        src/matrix/formats/cs_matrices.f90:294-322
//   ellpack_matrix_copy_matrix   src/matrix/formats/ellpack_matrices.f90:169-198
// i.e. build_graph_from_matrix + copy_matrix_values
// (src/matrix/sparse_matrix_interfaces.f90:692-772) with the builders
// cs_graph_build (src/graph/formats/cs_graphs.f90:109-197) and
// ellpack_graph_build (src/graph/formats/ellpack_graphs.f90:105-170).
//
// The reference walks the source's entries in ITS iteration order (stored
// arrays line by line; an ellpack matrix yields each row's first degrees(i)
// slots) and drops every edge into the first free slot of its target line --
// an O(ne * d) host scan.  The result is fully determined: target line l holds
// its entries in ascending source-iteration index.  So
//   * target lines keyed like the source's lines  -> the arrays are copied;
//   * keyed by the other index                    -> the STABLE transpose of
//     transpose.cu (counting sort, ascending source entry index per line),
//     which equals cs_graph_build(trans) bit for bit;
//   * ellpack on either side                      -> one (de)compaction pass:
//     slots in line order, padding = copy of the last neighbour, val pad = 0.
// Sources whose iterator returns the same edge twice (which the reference's
// own builders never produce: add_edge checks connectivity first) are not
// supported: the reference would de-duplicate them (:175-177).
#include <algorithm>
#include <vector>

#include "dist.h"
#include "internal.h"

namespace sigb {

namespace {

inline int grid_for(int64_t n)
{
    int64_t g = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)ctx().num_sms * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ellpack (slot-major on the device) -> compressed lines: row i's first
// degrees(i) slots, in slot order (ellpack_matrix_get_entries
// ellpack_matrices.f90:381-434)
__global__ void __launch_bounds__(kThreads)
ell_compact_kernel(const int32_t *__restrict__ node_sm, const double *__restrict__ val_sm,
                   const int32_t *__restrict__ degrees, const int32_t *__restrict__ ptr1, int32_t n,
                   int32_t n_pad, int32_t *__restrict__ node_out, double *__restrict__ val_out)
{
    for (int32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const int32_t b = ptr1[i] - 1, d = degrees[i];
        for (int32_t k = 0; k < d; k++) {
            node_out[b + k] = node_sm[(size_t)k * n_pad + i];
            val_out[b + k] = val_sm[(size_t)k * n_pad + i];
        }
    }
}

// compressed lines -> ellpack, slot-major: g%node(d+1:, i) = j on every
// insertion leaves the padding slots holding the row's last neighbour
// (ellpack_graphs.f90:164); A%val = 0 there (ellpack_matrices.f90:194).
// Rows [n, n_pad) are device-side padding: node = 1, val = 0.
__global__ void __launch_bounds__(kThreads)
cs_to_ell_kernel(const int32_t *__restrict__ ptr1, const int32_t *__restrict__ node1,
                 const double *__restrict__ val, int32_t n, int32_t n_pad, int32_t max_d,
                 int32_t *__restrict__ node_sm, double *__restrict__ val_sm, int32_t *__restrict__ degrees)
{
    for (int32_t i = blockIdx.x * kThreads + threadIdx.x; i < n_pad; i += gridDim.x * kThreads) {
        if (i < n) {
            const int32_t b = ptr1[i] - 1, d = ptr1[i + 1] - 1 - b;
            degrees[i] = d;
            const int32_t last = d > 0 ? node1[b + d - 1] : 1;
            for (int32_t k = 0; k < max_d; k++) {
                node_sm[(size_t)k * n_pad + i] = k < d ? node1[b + k] : last;
                val_sm[(size_t)k * n_pad + i] = k < d ? val[b + k] : 0.0;
            }
        } else {
            for (int32_t k = 0; k < max_d; k++) {
                node_sm[(size_t)k * n_pad + i] = 1;
                val_sm[(size_t)k * n_pad + i] = 0.0;
            }
        }
    }
}

// compressed-line arrays on the device
struct Lines {
    int32_t nlines = 0, nids = 0;
    int64_t ne = 0;
    int32_t *ptr = nullptr, *node = nullptr;   // 1-based, with kPad slack
    double *val = nullptr;                     // with kPad slack
    bool owned = false;
    void release()
    {
        if (owned) {
            cudaFree(ptr);
            cudaFree(node);
            cudaFree(val);
        }
        ptr = node = nullptr;
        val = nullptr;
        owned = false;
    }
};

int alloc_lines(Lines &L, int32_t nlines, int32_t nids, int64_t ne)
{
    L.nlines = nlines;
    L.nids = nids;
    L.ne = ne;
    L.owned = true;
    SIGB_CUDA(cudaMalloc((void **)&L.ptr, sizeof(int32_t) * ((size_t)nlines + 1 + kPad)));
    SIGB_CUDA(cudaMalloc((void **)&L.node, sizeof(int32_t) * ((size_t)ne + kPad)));
    SIGB_CUDA(cudaMalloc((void **)&L.val, sizeof(double) * ((size_t)ne + kPad)));
    SIGB_CHECK(fill_i32(L.ptr + nlines + 1, kPad, 1));
    SIGB_CHECK(fill_i32(L.node + ne, kPad, 1));
    SIGB_CHECK(fill_f64(L.val + ne, kPad, 0.0));
    return SIGB_OK;
}

// the source's entries as compressed lines in iteration order
int source_lines(sigb_matrix_t A, Lines &S, bool *lines_are_rows)
{
    sigb_graph_t g = A->g;
    cudaStream_t st = ctx().stream;
    if (g->kind != G_ELL) {
        S.nlines = g->n;
        S.nids = g->m;
        S.ne = g->ne;
        S.ptr = g->stored.ptr;
        S.node = g->stored.node;
        S.val = A->val;
        S.owned = false;
        *lines_are_rows = (g->kind == G_CSR);
        return SIGB_OK;
    }
    SIGB_CHECK(alloc_lines(S, g->n, g->m, g->ne));
    SIGB_CHECK(scan_to_ptr1(g->ell_degrees, g->n, S.ptr));
    if (g->n > 0) {
        ell_compact_kernel<<<grid_for(g->n), kThreads, 0, st>>>(g->ell_node, A->val, g->ell_degrees, S.ptr, g->n,
                                                               g->n_pad, S.node, S.val);
        count_launch();
        SIGB_CUDA(cudaGetLastError());
    }
    *lines_are_rows = true;
    return SIGB_OK;
}

}  // namespace

}  // namespace sigb

using namespace sigb;

extern "C" {

int sigb_matrix_copy(sigb_matrix_t A, int format, int trans, sigb_matrix_t *B_out)
{
    if (A && A->mg) { ::sigb::set_error("sigb_matrix_copy: not available for a single-process multi-GPU operator"); return SIGB_ERR_UNSUPPORTED; }
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(A && B_out, SIGB_ERR_ARG, "sigb_matrix_copy: bad argument");
    SIGB_REQUIRE(format == SIGB_FMT_CSR || format == SIGB_FMT_CSC || format == SIGB_FMT_ELLPACK, SIGB_ERR_ARG,
                 "sigb_matrix_copy: unknown target format %d", format);
    SIGB_REQUIRE(!A->op && !A->dist, SIGB_ERR_UNSUPPORTED,
                 "sigb_matrix_copy: the source must be a stored csr / csc / ellpack matrix");
    cudaStream_t st = ctx().stream;
    const bool tr = trans != 0;

    Lines S, T;
    bool src_lines_are_rows = true;
    int rc = source_lines(A, S, &src_lines_are_rows);
    if (rc != SIGB_OK) { S.release(); return rc; }

    // In SOURCE terms, the target's lines are keyed by the source row index when
    //   csr / ellpack target, not transposed   (target row    = source row)
    //   csc target, transposed                 (target column = source row)
    // and by the source column index otherwise.
    const bool tgt_lines_are_src_rows = ((format == SIGB_FMT_CSC) == tr);
    if (tgt_lines_are_src_rows == src_lines_are_rows) {
        rc = alloc_lines(T, S.nlines, S.nids, S.ne);
        if (rc == SIGB_OK) {
            cudaError_t e = cudaMemcpyAsync(T.ptr, S.ptr, sizeof(int32_t) * ((size_t)S.nlines + 1),
                                            cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess && S.ne > 0)
                e = cudaMemcpyAsync(T.node, S.node, sizeof(int32_t) * (size_t)S.ne, cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess && S.ne > 0)
                e = cudaMemcpyAsync(T.val, S.val, sizeof(double) * (size_t)S.ne, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) rc = cuda_fail(e, "matrix copy", __FILE__, __LINE__);
        }
    } else {
        int32_t *perm = nullptr;
        T.nlines = S.nids;
        T.nids = S.nlines;
        T.ne = S.ne;
        T.owned = true;
        rc = device_transpose_cs(S.ptr, S.node, S.nlines, S.nids, S.ne, &T.ptr, &T.node, &perm);
        if (rc == SIGB_OK) {
            cudaError_t e = cudaMalloc((void **)&T.val, sizeof(double) * ((size_t)T.ne + kPad));
            if (e != cudaSuccess) rc = cuda_fail(e, "matrix copy", __FILE__, __LINE__);
        }
        if (rc == SIGB_OK) rc = fill_f64(T.val + T.ne, kPad, 0.0);
        if (rc == SIGB_OK) rc = gather_values(S.val, perm, T.ne, T.val);
        cudaStreamSynchronize(st);
        cudaFree(perm);
    }
    if (rc != SIGB_OK) { S.release(); T.release(); return rc; }

    // tiling, max_d and the empty-line check come from the device-resident ptr (tiles_device.cu):
    // no read-back, no host loop
    int32_t max_d = 0, min_d = INT32_MAX;
    CsrView dev_view;
    rc = build_tiles_device(T.ptr, T.nlines, T.ne, dev_view, &max_d, &min_d);
    if (rc != SIGB_OK) { S.release(); T.release(); return rc; }
    if (T.nlines == 0) min_d = INT32_MAX;
    S.release();

    sigb_graph_t g = new sigb_graph_s();
    sigb_matrix_t B = new sigb_matrix_s();
    B->g = g;   // create returns the graph with one reference: the matrix's
    g->n = T.nlines;
    g->m = T.nids;
    g->ne = T.ne;
    g->max_d = max_d;
    if (format == SIGB_FMT_ELLPACK) {
        cudaFree(dev_view.tiles);   // an ellpack target needs the line lengths only
        if (T.nlines > 0 && min_d < 1) {
            T.release();
            sigb_matrix_destroy(B);
            set_error("sigb_matrix_copy: the ellpack copy would have a row with no edge; the reference would read "
                      "x(0) in its matvec (README.md:71-73)");
            return SIGB_ERR_ISOLATED;
        }
        g->kind = G_ELL;
        g->max_d = std::max(max_d, 1);
        g->n_pad = (g->n + 63) & ~63;
        B->nrow = g->n;
        B->ncol = g->m;
        const size_t len = (size_t)std::max(g->n_pad, 1) * g->max_d;
        cudaError_t e = cudaMalloc((void **)&g->ell_node, sizeof(int32_t) * len);
        if (e == cudaSuccess) e = cudaMalloc((void **)&g->ell_degrees, sizeof(int32_t) * (size_t)std::max(g->n, 1));
        if (e == cudaSuccess) e = cudaMalloc((void **)&B->val, sizeof(double) * len);
        if (e == cudaSuccess && g->n_pad > 0) {
            cs_to_ell_kernel<<<grid_for(g->n_pad), kThreads, 0, st>>>(T.ptr, T.node, T.val, g->n, g->n_pad, g->max_d,
                                                                     g->ell_node, B->val, g->ell_degrees);
            count_launch();
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        T.release();
        if (e != cudaSuccess) {
            sigb_matrix_destroy(B);
            return cuda_fail(e, "ellpack copy", __FILE__, __LINE__);
        }
    } else {
        g->kind = (format == SIGB_FMT_CSR) ? G_CSR : G_CSC;
        if (g->kind == G_CSC) { B->nrow = g->m; B->ncol = g->n; }
        else { B->nrow = g->n; B->ncol = g->m; }
        CsrView &v = g->stored;
        v.nrows = g->n;
        v.ncols = g->m;
        v.nnz = g->ne;
        v.ptr = T.ptr;      // ownership moves to the graph / matrix
        v.node = T.node;
        B->val = T.val;
        v.tiles = dev_view.tiles;
        v.tile_nnz = dev_view.tile_nnz;
        v.ntiles = dev_view.ntiles;
        v.tiles_nonempty = dev_view.tiles_nonempty;
        v.n_nonempty = dev_view.n_nonempty;
    }
    *B_out = B;
    return SIGB_OK;
}

int sigb_matrix_get_format(sigb_matrix_t A, int *format, int32_t *n_lines, int32_t *n_ids, int64_t *ne,
                           int32_t *max_d)
{
    if (A && A->mg) { ::sigb::set_error("sigb_matrix_get_format: not available for a single-process multi-GPU operator"); return SIGB_ERR_UNSUPPORTED; }
    SIGB_REQUIRE(A, SIGB_ERR_ARG, "sigb_matrix_get_format: null matrix");
    SIGB_REQUIRE(!A->op && !A->dist, SIGB_ERR_UNSUPPORTED, "sigb_matrix_get_format: not a stored matrix");
    sigb_graph_t g = A->g;
    if (format) *format = g->kind == G_CSR ? SIGB_FMT_CSR : (g->kind == G_CSC ? SIGB_FMT_CSC : SIGB_FMT_ELLPACK);
    if (n_lines) *n_lines = g->n;
    if (n_ids) *n_ids = g->m;
    if (ne) *ne = g->ne;
    if (max_d) *max_d = g->max_d;
    return SIGB_OK;
}

int sigb_matrix_get_arrays(sigb_matrix_t A, int32_t *ptr_or_degrees, int32_t *node, double *val)
{
    if (A && A->mg) { ::sigb::set_error("sigb_matrix_get_arrays: not available for a single-process multi-GPU operator"); return SIGB_ERR_UNSUPPORTED; }
    SIGB_REQUIRE(A, SIGB_ERR_ARG, "sigb_matrix_get_arrays: null matrix");
    SIGB_REQUIRE(!A->op && !A->dist, SIGB_ERR_UNSUPPORTED, "sigb_matrix_get_arrays: not a stored matrix");
    sigb_graph_t g = A->g;
    SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
    if (g->kind != G_ELL) {
        if (ptr_or_degrees)
            SIGB_CUDA(cudaMemcpy(ptr_or_degrees, g->stored.ptr, sizeof(int32_t) * ((size_t)g->n + 1), cudaMemcpyDeviceToHost));
        if (node && g->ne > 0)
            SIGB_CUDA(cudaMemcpy(node, g->stored.node, sizeof(int32_t) * (size_t)g->ne, cudaMemcpyDeviceToHost));
        if (val && g->ne > 0)
            SIGB_CUDA(cudaMemcpy(val, A->val, sizeof(double) * (size_t)g->ne, cudaMemcpyDeviceToHost));
        return SIGB_OK;
    }
    // ellpack: back from the device's slot-major layout to node(max_d, n) / val(max_d, n)
    const size_t len = (size_t)g->n_pad * g->max_d;
    if (ptr_or_degrees && g->n > 0)
        SIGB_CUDA(cudaMemcpy(ptr_or_degrees, g->ell_degrees, sizeof(int32_t) * (size_t)g->n, cudaMemcpyDeviceToHost));
    if (node && len > 0) {
        std::vector<int32_t> sm(len);
        SIGB_CUDA(cudaMemcpy(sm.data(), g->ell_node, sizeof(int32_t) * len, cudaMemcpyDeviceToHost));
        for (int32_t i = 0; i < g->n; i++)
            for (int32_t k = 0; k < g->max_d; k++) node[(size_t)i * g->max_d + k] = sm[(size_t)k * g->n_pad + i];
    }
    if (val && len > 0) {
        std::vector<double> sm(len);
        SIGB_CUDA(cudaMemcpy(sm.data(), A->val, sizeof(double) * len, cudaMemcpyDeviceToHost));
        for (int32_t i = 0; i < g->n; i++)
            for (int32_t k = 0; k < g->max_d; k++) val[(size_t)i * g->max_d + k] = sm[(size_t)k * g->n_pad + i];
    }
    return SIGB_OK;
}

}  // extern "C"
