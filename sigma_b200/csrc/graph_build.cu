// graph_build.cu -- the graph builders on the device: a pattern from an EDGE STREAM
// (first half).
//
// Replaces the bodies of
//   cs_graph_build        src/graph/formats/cs_graphs.f90:109-197
//   ellpack_graph_build   src/graph/formats/ellpack_graphs.f90:105-170
// as they are driven by copy_graph / convert_graph_type (src/graph/graph_interfaces.f90:276-318):
// the source graph's edges arrive in ITS iteration order, 64 at a time (batch_size, :267 -- the
// batching does not change the order), optionally with the endpoints swapped (trans).
//
// The reference counts the edges per line, prefix-sums the counts into ptr, then drops every edge
// into the first free slot of its line unless the line already holds it (:163-183: "iterators might
// return the same edge more than once"), skips null edges (node 0) and finally prunes the unused
// slots (:186-189).  That O(ne * d) host scan has a fully determined result: line l holds its
// DISTINCT neighbours in the order of their FIRST appearance in the stream.  On the device:
//   1. group  : the stable counting sort of transpose.cu buckets the edges by line, ascending stream
//               index inside a bucket (the same machinery as the transposed copies and the assembly);
//   2. dedup  : one thread per line walks its bucket in stream order and keeps an edge unless its
//               endpoint is 0 or already kept -- the reference's own check, on the line's own data;
//   3. compact: the kept counts are scanned into ptr and the kept lists moved into place.
// Index work only, and it must equal the serial builders bit for bit (tests/test_gpu_graph_build.py
// against orc_cs_graph_build / orc_ellpack_graph_build).  ellpack: the same lists laid out slot-major
// with the padding slots holding the line's last neighbour (ellpack_graphs.f90:164); a line without
// edges is refused like in sigb_ell_graph_create (the reference would read x(0), README.md:71-73).
#include <algorithm>
#include <vector>

#include "internal.h"

namespace sigb {

int ensure_graph_transposed(sigb_graph_t g);   // api.cu

namespace {

inline int grid_for(int64_t n)
{
    int64_t g = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)ctx().num_sms * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// bucket [bptr1(l) - 1, bptr1(l + 1) - 1) of `perm` lists the stream positions of line l's edges in
// ascending order.  kept[b + t] = t-th distinct non-null endpoint; cnt[l] = how many.
__global__ void __launch_bounds__(kThreads)
dedup_lines_kernel(const int32_t *__restrict__ bptr1, const int32_t *__restrict__ perm,
                   const int32_t *__restrict__ e2, int32_t n, int32_t *__restrict__ kept, int32_t *__restrict__ cnt)
{
    for (int32_t l = blockIdx.x * kThreads + threadIdx.x; l < n; l += gridDim.x * kThreads) {
        const int32_t b = bptr1[l] - 1, e = bptr1[l + 1] - 1;
        int32_t d = 0;
        for (int32_t p = b; p < e; p++) {
            const int32_t j = e2[perm[p]];
            bool skip = (j == 0);                          // a null edge of the source
            for (int32_t t = 0; t < d && !skip; t++) skip = (kept[b + t] == j);   // "if (g%node(l) == j) exit"
            if (!skip) kept[b + d++] = j;
        }
        cnt[l] = d;
    }
}

__global__ void __launch_bounds__(kThreads)
compact_lines_kernel(const int32_t *__restrict__ bptr1, const int32_t *__restrict__ kept,
                     const int32_t *__restrict__ ptr1, int32_t n, int32_t *__restrict__ node)
{
    for (int32_t l = blockIdx.x * kThreads + threadIdx.x; l < n; l += gridDim.x * kThreads) {
        const int32_t b = bptr1[l] - 1, o = ptr1[l] - 1, d = ptr1[l + 1] - 1 - o;
        for (int32_t t = 0; t < d; t++) node[o + t] = kept[b + t];
    }
}

// compressed lines -> ellpack node array, slot-major, padding = last neighbour; rows [n, n_pad): 1
__global__ void __launch_bounds__(kThreads)
lines_to_ell_kernel(const int32_t *__restrict__ ptr1, const int32_t *__restrict__ node1, int32_t n, int32_t n_pad,
                    int32_t max_d, int32_t *__restrict__ node_sm, int32_t *__restrict__ degrees)
{
    for (int32_t i = blockIdx.x * kThreads + threadIdx.x; i < n_pad; i += gridDim.x * kThreads) {
        if (i < n) {
            const int32_t b = ptr1[i] - 1, d = ptr1[i + 1] - 1 - b;
            degrees[i] = d;
            const int32_t last = d > 0 ? node1[b + d - 1] : 1;
            for (int32_t k = 0; k < max_d; k++) node_sm[(size_t)k * n_pad + i] = k < d ? node1[b + k] : last;
        } else {
            for (int32_t k = 0; k < max_d; k++) node_sm[(size_t)k * n_pad + i] = 1;
        }
    }
}

// the distinct-neighbour lists of an edge stream as compressed lines on the device (arrays owned by the caller)
struct BuiltLines {
    int32_t *ptr = nullptr, *node = nullptr;
    int64_t ne = 0;
};

int build_lines(int32_t n, int32_t m, int64_t count, const int32_t *src_i, const int32_t *src_j, int trans,
                BuiltLines *out)
{
    SIGB_REQUIRE(count <= INT32_MAX - 16, SIGB_ERR_ARG, "graph build: at most 2^31 edges in a stream");
    // (validated on the host like the *_create entry points: a bad line id would scatter out of range)
    const int32_t *h1 = trans ? src_j : src_i, *h2 = trans ? src_i : src_j;
    for (int64_t k = 0; k < count; k++) {
        SIGB_REQUIRE(h1[k] >= 1 && h1[k] <= n, SIGB_ERR_ARG, "graph build: edge %lld starts at vertex %d outside 1..%d",
                     (long long)k + 1, h1[k], n);
        SIGB_REQUIRE(h2[k] >= 0 && h2[k] <= m, SIGB_ERR_ARG, "graph build: edge %lld ends at vertex %d outside 0..%d",
                     (long long)k + 1, h2[k], m);
    }
    cudaStream_t st = ctx().stream;
    int32_t *e1 = nullptr, *e2 = nullptr, *one_line = nullptr, *kept = nullptr, *cnt = nullptr;
    int32_t *bptr = nullptr, *unused = nullptr, *perm = nullptr;
    int32_t *ptr = nullptr, *node = nullptr;
    auto cleanup = [&]() {
        tmp_free(e1); tmp_free(e2); tmp_free(one_line); tmp_free(kept); tmp_free(cnt);
        cudaFree(bptr); cudaFree(unused); cudaFree(perm);
    };
    int rc = SIGB_OK;
#define GB_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cleanup(); cudaFree(ptr); cudaFree(node); return cuda_fail(e_, #expr, __FILE__, __LINE__); } } while (0)
#define GB_TRY(expr) do { rc = (expr); if (rc != SIGB_OK) { cleanup(); cudaFree(ptr); cudaFree(node); return rc; } } while (0)
    const size_t cn = (size_t)std::max<int64_t>(count, 1);
    GB_CUDA(tmp_alloc(&e1, cn + kPad));
    GB_CUDA(tmp_alloc(&e2, cn));
    GB_CUDA(tmp_alloc(&kept, cn));
    GB_CUDA(tmp_alloc(&cnt, (size_t)std::max(n, 1)));
    GB_CUDA(tmp_alloc(&one_line, (size_t)(2 + kPad)));
    if (count > 0) {
        GB_CUDA(cudaMemcpyAsync(e1, h1, sizeof(int32_t) * (size_t)count, cudaMemcpyHostToDevice, st));
        GB_CUDA(cudaMemcpyAsync(e2, h2, sizeof(int32_t) * (size_t)count, cudaMemcpyHostToDevice, st));
    }
    {
        int32_t h_line[2] = {1, (int32_t)(count + 1)};
        GB_CUDA(cudaMemcpyAsync(one_line, h_line, sizeof(h_line), cudaMemcpyHostToDevice, st));
        GB_CUDA(cudaStreamSynchronize(st));
    }
    // 1. group by line, stream order kept inside a line
    if (count > 0) GB_TRY(device_transpose_cs(one_line, e1, 1, n, count, &bptr, &unused, &perm));
    else GB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)std::max(n, 1), st));
    // 2. distinct neighbours in first-appearance order
    if (n > 0 && count > 0) {
        dedup_lines_kernel<<<grid_for(n), kThreads, 0, st>>>(bptr, perm, e2, n, kept, cnt);
        count_launch();
        GB_CUDA(cudaGetLastError());
    }
    // 3. ptr = 1 + prefix sums of the kept counts ; node = the kept lists, packed
    GB_CUDA(cudaMalloc((void **)&ptr, sizeof(int32_t) * ((size_t)n + 1 + kPad)));
    GB_TRY(scan_to_ptr1(cnt, n, ptr));
    GB_TRY(fill_i32(ptr + n + 1, kPad, 1));
    int32_t last = 1;
    GB_CUDA(cudaMemcpyAsync(&last, ptr + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    const int64_t ne = (int64_t)last - 1;
    GB_CUDA(cudaMalloc((void **)&node, sizeof(int32_t) * ((size_t)ne + kPad)));
    GB_TRY(fill_i32(node + ne, kPad, 1));
    if (n > 0 && ne > 0) {
        compact_lines_kernel<<<grid_for(n), kThreads, 0, st>>>(bptr, kept, ptr, n, node);
        count_launch();
        GB_CUDA(cudaGetLastError());
    }
    GB_CUDA(cudaStreamSynchronize(st));
    cleanup();
#undef GB_CUDA
#undef GB_TRY
    out->ptr = ptr;
    out->node = node;
    out->ne = ne;
    return SIGB_OK;
}

}  // namespace

}  // namespace sigb

using namespace sigb;

extern "C" {

int sigb_cs_graph_build(int32_t n, int32_t m, int64_t count, const int32_t *src_i, const int32_t *src_j, int trans,
                        int order, sigb_graph_t *out)
{
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(out && n >= 0 && m >= 0 && count >= 0 && (count == 0 || (src_i && src_j)), SIGB_ERR_ARG,
                 "sigb_cs_graph_build: bad argument");
    SIGB_REQUIRE(order == SIGB_ROW || order == SIGB_COL, SIGB_ERR_ARG, "sigb_cs_graph_build: bad order");
    BuiltLines L;
    SIGB_CHECK(build_lines(n, m, count, src_i, src_j, trans, &L));
    sigb_graph_t g = new sigb_graph_s();
    g->kind = (order == SIGB_ROW) ? G_CSR : G_CSC;
    g->n = n;
    g->m = m;
    g->ne = L.ne;
    CsrView &v = g->stored;
    v.nrows = n;
    v.ncols = m;
    v.nnz = L.ne;
    v.ptr = L.ptr;      // ownership moves to the graph
    v.node = L.node;
    int32_t max_d = 0;
    int rc = build_tiles_device(v.ptr, n, L.ne, v, &max_d, nullptr);
    if (rc == SIGB_OK) {
        g->max_d = n > 0 ? max_d : 0;
        if (g->kind == G_CSC) rc = ensure_graph_transposed(g);
    }
    if (rc != SIGB_OK) {
        sigb_graph_release(g);
        return rc;
    }
    *out = g;
    return SIGB_OK;
}

int sigb_ell_graph_build(int32_t n, int32_t m, int64_t count, const int32_t *src_i, const int32_t *src_j, int trans,
                         sigb_graph_t *out)
{
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(out && n >= 0 && m >= 0 && count >= 0 && (count == 0 || (src_i && src_j)), SIGB_ERR_ARG,
                 "sigb_ell_graph_build: bad argument");
    BuiltLines L;
    SIGB_CHECK(build_lines(n, m, count, src_i, src_j, trans, &L));
    struct Free {
        BuiltLines *L;
        ~Free() { cudaFree(L->ptr); cudaFree(L->node); }
    } free_lines{&L};
    CsrView probe;     // only the extreme line lengths are needed
    int32_t max_d = 0, min_d = 0;
    SIGB_CHECK(build_tiles_device(L.ptr, n, L.ne, probe, &max_d, &min_d));
    cudaFree(probe.tiles);
    SIGB_REQUIRE(n == 0 || min_d >= 1, SIGB_ERR_ISOLATED,
                 "sigb_ell_graph_build: a row has no edge; the reference would read x(0) in its matvec (README.md:71-73)");
    sigb_graph_t g = new sigb_graph_s();
    g->kind = G_ELL;
    g->n = n;
    g->m = m;
    g->ne = L.ne;
    g->max_d = std::max(max_d, 1);
    g->n_pad = (n + 63) & ~63;
    cudaStream_t st = ctx().stream;
    const size_t len = (size_t)std::max(g->n_pad, 1) * g->max_d;
    cudaError_t e = cudaMalloc((void **)&g->ell_node, sizeof(int32_t) * len);
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->ell_degrees, sizeof(int32_t) * (size_t)std::max(n, 1));
    if (e == cudaSuccess && g->n_pad > 0) {
        lines_to_ell_kernel<<<grid_for(g->n_pad), kThreads, 0, st>>>(L.ptr, L.node, n, g->n_pad, g->max_d, g->ell_node,
                                                                    g->ell_degrees);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        sigb_graph_release(g);
        return cuda_fail(e, "ellpack graph build", __FILE__, __LINE__);
    }
    *out = g;
    return SIGB_OK;
}

}  // extern "C"
