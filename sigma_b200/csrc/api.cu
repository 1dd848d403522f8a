// api.cu -- the C-ABI of include/sigma_b200.h: runtime, graph / matrix
// mirrors, matvec dispatch and the solver / eigensolver entry points.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "device_utils.cuh"
#include "dist.h"
#include "internal.h"
#include "solvers.h"

namespace sigb {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return SIGB_ERR_CUDA;
}

Ctx &ctx()
{
    static thread_local Ctx c;
    return c;
}

int require_init()
{
    if (!ctx().inited) return sigb_init(-1);
    return SIGB_OK;
}

int check_fault(const char *where)
{
    const Ctx &c = ctx();
    if (c.fault == nullptr) return SIGB_OK;
    const unsigned code = *reinterpret_cast<volatile unsigned int *>(&c.fault->code);
    if (code == 0u) return SIGB_OK;
    static const char *what[] = {"", "a grid barrier", "an all-reduce inbox", "a halo acknowledgement", "a halo flag",
                                 "a sync-free sweep"};
    set_error("%s: a device-side wait on %s timed out (a peer GPU is late or gone, or a kernel was descheduled); "
              "results since then are invalid and the row-sharded operators of this process are unusable",
              where, code < sizeof(what) / sizeof(what[0]) ? what[code] : "a flag");
    return SIGB_ERR_COMM;
}

// Short-lived device scratch of the copy / transpose / assembly entry points comes from the
// device's default memory pool, stream-ordered on the library's stream, with the release
// threshold lifted so that freed blocks stay cached: cudaMalloc / cudaFree synchronise the whole
// device per call, which was most of a transposing copy (round 2: 12.7 ms -> 4.9 ms for 21 M
// entries, profiles/r2_visit_a_1gpu_summary.txt).  Falls back to cudaMalloc when the driver has no pool.
static bool async_alloc_enabled()
{
    static thread_local int v = -1;   // per device: every host thread of this library drives one GPU
    if (v < 0) {
        v = 1;
        cudaMemPool_t pool = nullptr;
        unsigned long long keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx().device) != cudaSuccess ||
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) {
            cudaGetLastError();
            v = 0;
        }
    }
    return v != 0;
}

cudaError_t tmp_alloc_bytes(void **p, size_t bytes)
{
    if (async_alloc_enabled()) return cudaMallocAsync(p, bytes, ctx().stream);
    return cudaMalloc(p, bytes);
}

cudaError_t tmp_free(void *p)
{
    if (p == nullptr) return cudaSuccess;
    if (async_alloc_enabled()) return cudaFreeAsync(p, ctx().stream);
    return cudaFree(p);
}

template <typename T>
static int dev_alloc(T **p, size_t count)
{
    SIGB_CUDA(cudaMalloc((void **)p, sizeof(T) * (count > 0 ? count : 1)));
    return SIGB_OK;
}

// upload the tile table of a CSR view
int upload_tiles(CsrView &v, const std::vector<TileDesc> &tiles)
{
    // one allocation: the full table, then the sub-table of tiles that hold entries
    std::vector<TileDesc> both(tiles);
    for (const TileDesc &t : tiles)
        if (t.ke > t.ks) both.push_back(t);
    v.ntiles = (int32_t)tiles.size();
    v.n_nonempty = (int32_t)(both.size() - tiles.size());
    SIGB_CHECK(dev_alloc(&v.tiles, both.size()));
    v.tiles_nonempty = v.tiles + v.ntiles;
    SIGB_CUDA(cudaMemcpyAsync(v.tiles, both.data(), sizeof(TileDesc) * both.size(),
                              cudaMemcpyHostToDevice, ctx().stream));
    SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
    return SIGB_OK;
}

static void free_view(CsrView &v)
{
    cudaFree(v.ptr);
    cudaFree(v.node);
    cudaFree(v.tiles);   // tiles_interior / tiles_boundary are views into it
    v = CsrView();
}

// Early returns out of the create functions (a CUDA failure half way) must not leak what
// has been built so far: the guards release it unless dismissed on success.
struct GraphGuard {
    sigb_graph_t g;
    ~GraphGuard() { if (g) sigb_graph_release(g); }
};
struct MatrixGuard {
    sigb_matrix_t A;
    ~MatrixGuard() { if (A) sigb_matrix_destroy(A); }
};
struct DevBufGuard {
    void *p;
    ~DevBufGuard() { cudaFree(p); }
};

// Build the stable transpose of a cs / ell graph on the device (once).
int ensure_graph_transposed(sigb_graph_t g)
{
    if (g->has_transposed) return SIGB_OK;
    int32_t *ptr_t = nullptr, *node_t = nullptr, *perm = nullptr;
    int32_t ntargets;
    int64_t ne;
    if (g->kind == G_ELL) {
        ntargets = g->m;
        ne = (int64_t)g->n * g->max_d;
        SIGB_CHECK(device_transpose_ell(g->ell_node, g->n, g->n_pad, g->max_d, ntargets, &ptr_t,
                                        &node_t, &perm));
    } else {
        ntargets = g->m;
        ne = g->ne;
        SIGB_CHECK(device_transpose_cs(g->stored.ptr, g->stored.node, g->n, ntargets, ne, &ptr_t,
                                       &node_t, &perm));
    }
    CsrView &t = g->transposed;
    t.nrows = ntargets;
    t.ncols = g->n;
    t.nnz = ne;
    t.ptr = ptr_t;
    t.node = node_t;
    g->perm_t = perm;
    // the tile table is built on the device from the device-resident ptr (tiles_device.cu): no
    // read-back, no host loop
    SIGB_CHECK(build_tiles_device(ptr_t, ntargets, ne, t, nullptr, nullptr));
    g->has_transposed = true;
    return SIGB_OK;
}

int ensure_transposed(sigb_matrix_t A)
{
    sigb_graph_t g = A->g;
    SIGB_CHECK(ensure_graph_transposed(g));
    if (A->val_t_valid) return SIGB_OK;
    const int64_t ne = g->transposed.nnz;
    if (!A->val_t) {
        SIGB_CHECK(dev_alloc(&A->val_t, (size_t)ne + 8));
        SIGB_CHECK(fill_f64(A->val_t + ne, 8, 0.0));
    }
    if (g->kind == G_ELL)
        SIGB_CHECK(gather_values_ell(A->val, g->perm_t, ne, g->n_pad, g->max_d, A->val_t));
    else
        SIGB_CHECK(gather_values(A->val, g->perm_t, ne, A->val_t));
    A->val_t_valid = true;
    return SIGB_OK;
}

// y = op(A) x or y += op(A) x on device vectors, reproducing the reference's
// accumulation order per format (see SpmvMode).
int matvec_dev(sigb_matrix_t A, int trans, const double *x, double *y, SpmvMode, bool add,
               const DotSpec &dot)
{
    if (A->op) return op_matvec(A, trans, x, y, add, dot);
    sigb_graph_t g = A->g;
    if (A->dist) {
        SIGB_REQUIRE(!trans, SIGB_ERR_UNSUPPORTED, "matvec_t on a row-sharded operator");
        return dist_matvec(A, x, y, add, dot, /*x_has_halo=*/false);
    }
    if (g->kind == G_ELL) {
        if (!trans)
            return launch_ell_spmv(g->n, g->n_pad, g->max_d, g->ell_node, A->val, x, y,
                                   add ? MODE_ADD_AFTER : MODE_SET, dot);
        // ellpack_matvec_t_add scatters y(i) += val*z line by line
        SIGB_CHECK(ensure_transposed(A));
        return launch_csr_spmv(g->transposed, A->val_t, x, y, add ? MODE_ACC_INIT : MODE_SET, dot);
    }
    // line-wise kernel on the stored arrays: csr & !trans, csc & trans
    const bool stored_is_rowwise = (g->kind == G_CSR) != (trans != 0);
    if (stored_is_rowwise)
        return launch_csr_spmv(g->stored, A->val, x, y, add ? MODE_ADD_AFTER : MODE_SET, dot);
    SIGB_CHECK(ensure_transposed(A));
    return launch_csr_spmv(g->transposed, A->val_t, x, y, add ? MODE_ACC_INIT : MODE_SET, dot);
}

int solver_matvec(sigb_matrix_t A, const double *x, double *y, const DotSpec &dot, bool x_has_halo)
{
    if (A->dist) return dist_matvec(A, x, y, false, dot, x_has_halo);
    return matvec_dev(A, 0, x, y, MODE_SET, false, dot);
}

static int ensure_xb(sigb_solver_t s, int64_t len)
{
    if (s->xb_len >= len) return SIGB_OK;
    cudaFree(s->xb);
    s->xb = nullptr;
    SIGB_CHECK(dev_alloc(&s->xb, (size_t)len));
    s->xb_len = len;
    return SIGB_OK;
}

}  // namespace sigb

using namespace sigb;

// ===========================================================================
// runtime
// ===========================================================================
extern "C" {

const char *sigb_last_error(void) { return g_err; }
const char *sigb_version(void) { return "sigma_b200 0.1 (sm_100a)"; }

int sigb_init(int device)
{
    Ctx &c = ctx();
    if (c.inited) return SIGB_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("sigb_init: no CUDA device is usable (%s); this library has no CPU path",
                  cudaGetErrorString(e));
        return SIGB_ERR_CUDA;
    }
    if (device < 0) {
        const char *lr = getenv("LOCAL_RANK");
        device = lr ? atoi(lr) % count : 0;
    }
    SIGB_REQUIRE(device < count, SIGB_ERR_ARG, "sigb_init: device %d of %d", device, count);
    SIGB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SIGB_CUDA(cudaGetDeviceProperties(&prop, device));
    SIGB_REQUIRE(prop.major >= 10, SIGB_ERR_CUDA,
                 "sigb_init: device %d is sm_%d%d; this library is built for sm_100a only", device,
                 prop.major, prop.minor);
    c.device = device;
    c.num_sms = prop.multiProcessorCount;
    SIGB_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    SIGB_CUDA(cudaStreamCreateWithFlags(&c.aux_stream, cudaStreamNonBlocking));
    c.stream = c.own_stream;
    SIGB_CUDA(cudaMalloc(&c.partials, sizeof(double) * kMaxGrid * kMaxDots * kNumTickets));
    SIGB_CUDA(cudaMalloc(&c.tickets, sizeof(unsigned) * kNumTickets));
    SIGB_CUDA(cudaMemset(c.tickets, 0, sizeof(unsigned) * kNumTickets));
    c.pinned_bytes = 1 << 16;
    SIGB_CUDA(cudaMallocHost(&c.pinned, c.pinned_bytes));
    // device-side wait timeouts (device_utils.cuh spin_wait): one block in mapped pinned memory
    SIGB_CUDA(cudaHostAlloc((void **)&c.fault, sizeof(FaultBlock), cudaHostAllocMapped));
    c.fault->code = 0u;
    c.fault->pad_ = 0u;
    c.fault->limit_ns = (unsigned long long)std::max(1, env_int("SIGB_WAIT_TIMEOUT_MS", 30000)) * 1000000ull;
    SIGB_CUDA(cudaHostGetDevicePointer((void **)&c.fault_dev, c.fault, 0));
    c.launches = 0;
    c.inited = true;
    return SIGB_OK;
}

int sigb_finalize(void)
{
    Ctx &c = ctx();
    if (!c.inited) return SIGB_OK;
    cudaDeviceSynchronize();
    cudaFree(c.partials);
    cudaFree(c.tickets);
    cudaFreeHost(c.pinned);
    cudaFreeHost(c.fault);
    cudaFree(c.stage[0]);
    cudaFree(c.stage[1]);
    cudaStreamDestroy(c.own_stream);
    cudaStreamDestroy(c.aux_stream);
    c = Ctx();
    return SIGB_OK;
}

int sigb_set_stream(void *cuda_stream)
{
    SIGB_CHECK(require_init());
    ctx().stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx().own_stream;
    return SIGB_OK;
}

int sigb_synchronize(void)
{
    SIGB_CHECK(require_init());
    SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
    return check_fault("sigb_synchronize");
}

int64_t sigb_launch_count(void) { return ctx().launches + mgpu_launch_count(); }

int sigb_dev_alloc(int64_t bytes, void **ptr_dev)
{
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(ptr_dev && bytes >= 0, SIGB_ERR_ARG, "sigb_dev_alloc: bad argument");
    SIGB_CUDA(cudaMalloc(ptr_dev, (size_t)(bytes > 0 ? bytes : 1)));
    return SIGB_OK;
}

int sigb_dev_free(void *ptr_dev)
{
    SIGB_CUDA(cudaFree(ptr_dev));
    return SIGB_OK;
}

int sigb_copy_h2d(void *dst_dev, const void *src, int64_t bytes)
{
    SIGB_CHECK(require_init());
    SIGB_CUDA(cudaMemcpyAsync(dst_dev, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx().stream));
    SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
    return SIGB_OK;
}

int sigb_copy_d2h(void *dst, const void *src_dev, int64_t bytes)
{
    SIGB_CHECK(require_init());
    SIGB_CUDA(cudaMemcpyAsync(dst, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx().stream));
    SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
    return SIGB_OK;
}

// ===========================================================================
// graphs
// ===========================================================================

int sigb_cs_graph_create(int32_t n, int32_t m, const int32_t *ptr1, const int32_t *node1, int order,
                         sigb_graph_t *out)
{
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(out && ptr1 && n >= 0 && m >= 0, SIGB_ERR_ARG, "sigb_cs_graph_create: bad argument");
    SIGB_REQUIRE(order == SIGB_ROW || order == SIGB_COL, SIGB_ERR_ARG,
                 "sigb_cs_graph_create: order must be SIGB_ROW or SIGB_COL");
    SIGB_REQUIRE(ptr1[0] == 1, SIGB_ERR_ARG, "sigb_cs_graph_create: ptr must be 1-based (ptr(1) = %d)",
                 ptr1[0]);
    const int64_t ne = (int64_t)ptr1[n] - 1;
    SIGB_REQUIRE(ne >= 0 && (ne == 0 || node1), SIGB_ERR_ARG, "sigb_cs_graph_create: bad ptr/node");
    int32_t max_d = 0;
    for (int32_t i = 0; i < n; i++) {
        const int32_t d = ptr1[i + 1] - ptr1[i];
        SIGB_REQUIRE(d >= 0, SIGB_ERR_ARG, "sigb_cs_graph_create: ptr not monotone at line %d", i + 1);
        max_d = std::max(max_d, d);
    }
    for (int64_t k = 0; k < ne; k++)
        SIGB_REQUIRE(node1[k] >= 1 && node1[k] <= m, SIGB_ERR_ARG,
                     "sigb_cs_graph_create: node(%lld) = %d outside 1..%d", (long long)k + 1,
                     node1[k], m);

    sigb_graph_t g = new sigb_graph_s();
    GraphGuard guard{g};
    g->kind = (order == SIGB_ROW) ? G_CSR : G_CSC;
    g->n = n;
    g->m = m;
    g->ne = ne;
    g->max_d = max_d;
    CsrView &v = g->stored;
    v.nrows = n;
    v.ncols = m;
    v.nnz = ne;
    cudaStream_t st = ctx().stream;
    SIGB_CHECK(dev_alloc(&v.ptr, (size_t)n + 1 + kPad));
    SIGB_CHECK(fill_i32(v.ptr + n + 1, kPad, 1));
    SIGB_CHECK(dev_alloc(&v.node, (size_t)ne + kPad));
    SIGB_CUDA(cudaMemcpyAsync(v.ptr, ptr1, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
    if (ne > 0)
        SIGB_CUDA(cudaMemcpyAsync(v.node, node1, sizeof(int32_t) * (size_t)ne, cudaMemcpyHostToDevice, st));
    SIGB_CHECK(fill_i32(v.node + ne, kPad, 1));
    std::vector<TileDesc> tiles;
    TileShape shape;
    build_tiles_host(ptr1, n, tiles, &shape);
    v.tile_nnz = shape.nnz;
    SIGB_CHECK(upload_tiles(v, tiles));
    if (g->kind == G_CSC) SIGB_CHECK(ensure_graph_transposed(g));
    guard.g = nullptr;
    *out = g;
    return SIGB_OK;
}

int sigb_ell_graph_create(int32_t n, int32_t m, int32_t max_d, const int32_t *node_cm,
                          const int32_t *degrees, sigb_graph_t *out)
{
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(out && node_cm && degrees && n >= 0 && m >= 0 && max_d >= 1, SIGB_ERR_ARG,
                 "sigb_ell_graph_create: bad argument");
    int64_t ne = 0;
    for (int32_t i = 0; i < n; i++) {
        SIGB_REQUIRE(degrees[i] >= 1, SIGB_ERR_ISOLATED,
                     "sigb_ell_graph_create: row %d has no edge; the reference would read x(0) "
                     "(README.md:71-73)", i + 1);
        SIGB_REQUIRE(degrees[i] <= max_d, SIGB_ERR_ARG, "sigb_ell_graph_create: degrees(%d) > max_d", i + 1);
        ne += degrees[i];
        for (int32_t k = 0; k < max_d; k++) {
            const int32_t c = node_cm[(size_t)i * max_d + k];
            SIGB_REQUIRE(c >= 1 && c <= m, SIGB_ERR_ARG,
                         "sigb_ell_graph_create: node(%d,%d) = %d outside 1..%d", k + 1, i + 1, c, m);
        }
    }
    sigb_graph_t g = new sigb_graph_s();
    GraphGuard guard{g};
    g->kind = G_ELL;
    g->n = n;
    g->m = m;
    g->ne = ne;
    g->max_d = max_d;
    g->n_pad = (n + 63) & ~63;
    cudaStream_t st = ctx().stream;
    int32_t *tmp = nullptr;
    SIGB_CHECK(dev_alloc(&tmp, (size_t)n * max_d));
    DevBufGuard tmp_guard{tmp};
    SIGB_CUDA(cudaMemcpyAsync(tmp, node_cm, sizeof(int32_t) * (size_t)n * max_d, cudaMemcpyHostToDevice, st));
    SIGB_CHECK(dev_alloc(&g->ell_node, (size_t)g->n_pad * max_d));
    SIGB_CHECK(ell_relayout_node(tmp, n, g->n_pad, max_d, g->ell_node));
    SIGB_CHECK(dev_alloc(&g->ell_degrees, (size_t)n));
    SIGB_CUDA(cudaMemcpyAsync(g->ell_degrees, degrees, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    SIGB_CUDA(cudaStreamSynchronize(st));
    guard.g = nullptr;
    *out = g;
    return SIGB_OK;
}

int sigb_graph_retain(sigb_graph_t g)
{
    SIGB_REQUIRE(g, SIGB_ERR_ARG, "sigb_graph_retain: null graph");
    g->refcount++;
    return SIGB_OK;
}

int sigb_graph_release(sigb_graph_t g)
{
    SIGB_REQUIRE(g, SIGB_ERR_ARG, "sigb_graph_release: null graph");
    if (--g->refcount > 0) return SIGB_OK;
    free_view(g->stored);
    free_view(g->transposed);
    cudaFree(g->perm_t);
    cudaFree(g->ell_node);
    cudaFree(g->ell_degrees);
    delete g;
    return SIGB_OK;
}

int sigb_cs_graph_get_transpose(sigb_graph_t g, int32_t *ptr_t1, int32_t *node_t1)
{
    SIGB_REQUIRE(g, SIGB_ERR_ARG, "sigb_cs_graph_get_transpose: null graph");
    SIGB_CHECK(ensure_graph_transposed(g));
    const CsrView &t = g->transposed;
    if (ptr_t1)
        SIGB_CUDA(cudaMemcpy(ptr_t1, t.ptr, sizeof(int32_t) * ((size_t)t.nrows + 1), cudaMemcpyDeviceToHost));
    if (node_t1 && t.nnz > 0)
        SIGB_CUDA(cudaMemcpy(node_t1, t.node, sizeof(int32_t) * (size_t)t.nnz, cudaMemcpyDeviceToHost));
    return SIGB_OK;
}

// ===========================================================================
// matrices
// ===========================================================================

int sigb_matrix_create(sigb_graph_t g, sigb_matrix_t *out)
{
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(g && out, SIGB_ERR_ARG, "sigb_matrix_create: bad argument");
    sigb_matrix_t A = new sigb_matrix_s();
    MatrixGuard guard{A};
    A->g = g;
    g->refcount++;
    if (g->kind == G_CSC) { A->nrow = g->m; A->ncol = g->n; }
    else { A->nrow = g->n; A->ncol = g->m; }
    const size_t len = (g->kind == G_ELL) ? (size_t)g->n_pad * g->max_d : (size_t)g->ne + 8;
    SIGB_CHECK(dev_alloc(&A->val, len));
    SIGB_CUDA(cudaMemsetAsync(A->val, 0, sizeof(double) * len, ctx().stream));  // A%val = 0
    guard.A = nullptr;
    *out = A;
    return SIGB_OK;
}

int sigb_matrix_set_values(sigb_matrix_t A, const double *val, int64_t count)
{
    SIGB_REQUIRE(A && val, SIGB_ERR_ARG, "sigb_matrix_set_values: bad argument");
    if (A->mg) return mgpu_set_values(A, val, count);
    SIGB_REQUIRE(!A->op, SIGB_ERR_UNSUPPORTED, "sigb_matrix_set_values: an operator expression has no values of its own");
    sigb_graph_t g = A->g;
    cudaStream_t st = ctx().stream;
    if (g->kind == G_ELL) {
        const int64_t want = (int64_t)g->n * g->max_d;
        SIGB_REQUIRE(count == want, SIGB_ERR_ARG, "sigb_matrix_set_values: got %lld values, val(max_d,n) has %lld",
                     (long long)count, (long long)want);
        double *tmp = nullptr;
        SIGB_CHECK(dev_alloc(&tmp, (size_t)want));
        DevBufGuard tmp_guard{tmp};
        SIGB_CUDA(cudaMemcpyAsync(tmp, val, sizeof(double) * (size_t)want, cudaMemcpyHostToDevice, st));
        SIGB_CHECK(ell_relayout_val(tmp, g->n, g->n_pad, g->max_d, A->val));
        SIGB_CUDA(cudaStreamSynchronize(st));
    } else {
        SIGB_REQUIRE(count == g->ne, SIGB_ERR_ARG, "sigb_matrix_set_values: got %lld values, graph has %lld edges",
                     (long long)count, (long long)g->ne);
        if (count > 0)
            SIGB_CUDA(cudaMemcpyAsync(A->val, val, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, st));
        SIGB_CUDA(cudaStreamSynchronize(st));
    }
    A->val_t_valid = false;
    if (g->kind == G_CSC) SIGB_CHECK(ensure_transposed(A));
    return SIGB_OK;
}

int sigb_matrix_destroy(sigb_matrix_t A)
{
    if (!A) return SIGB_OK;
    // remove_reference; the mirror goes away with its last owner (an expression
    // that still points at this operator keeps it alive)
    if (--A->refcount > 0) return SIGB_OK;
    if (A->mg) mgpu_matrix_free(A);
    cudaFree(A->val);
    cudaFree(A->val_t);
    if (A->dist) dist_destroy(A);
    if (A->op) op_destroy(A);
    if (A->g) sigb_graph_release(A->g);
    delete A;
    return SIGB_OK;
}

int sigb_matrix_get_dims(sigb_matrix_t A, int32_t *nrow, int32_t *ncol, int64_t *nnz)
{
    SIGB_REQUIRE(A, SIGB_ERR_ARG, "sigb_matrix_get_dims: null matrix");
    if (nrow) *nrow = A->nrow;
    if (ncol) *ncol = A->ncol;
    if (nnz) {
        // composite_mat_get_nnz sums its blocks (sparse_matrix_composites.f90:445-460);
        // the lazy expressions hold no entries of their own
        int64_t total = 0;
        if (A->mg) {
            total = mgpu_nnz(A);
        } else if (A->op) {
            if (A->op->kind == OP_COMPOSITE)
                for (sigb_matrix_t k : A->op->kids) {
                    int64_t sub = 0;
                    SIGB_CHECK(sigb_matrix_get_dims(k, nullptr, nullptr, &sub));
                    total += sub;
                }
        } else {
            total = A->g->ne;
        }
        *nnz = total;
    }
    return SIGB_OK;
}

int sigb_matrix_get_transpose_values(sigb_matrix_t A, double *val_t)
{
    if (A && A->mg) { ::sigb::set_error("sigb_matrix_get_transpose_values: not available for a single-process multi-GPU operator"); return SIGB_ERR_UNSUPPORTED; }
    SIGB_REQUIRE(A && val_t, SIGB_ERR_ARG, "sigb_matrix_get_transpose_values: bad argument");
    SIGB_REQUIRE(!A->op, SIGB_ERR_UNSUPPORTED, "sigb_matrix_get_transpose_values: not a stored matrix");
    SIGB_CHECK(ensure_transposed(A));
    SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
    if (A->g->transposed.nnz > 0)
        SIGB_CUDA(cudaMemcpy(val_t, A->val_t, sizeof(double) * (size_t)A->g->transposed.nnz, cudaMemcpyDeviceToHost));
    return SIGB_OK;
}

// ===========================================================================
// matvec
// ===========================================================================

// grow-only device staging for the host-pointer entry points: A%matvec through the Fortran shim /
// sigma.hpp comes here on every call, so no cudaMalloc / cudaFree pair (each synchronises the
// device) and nothing to leak on an early return
static int ensure_stage(int which, size_t count, double **p)
{
    Ctx &c = ctx();
    if (c.stage_len[which] < count) {
        SIGB_CUDA(cudaStreamSynchronize(c.stream));
        cudaFree(c.stage[which]);
        c.stage[which] = nullptr;
        c.stage_len[which] = 0;
        const size_t len = count + count / 4 + 32;
        SIGB_CUDA(cudaMalloc((void **)&c.stage[which], sizeof(double) * len));
        c.stage_len[which] = len;
    }
    *p = c.stage[which];
    return SIGB_OK;
}

static int host_matvec(sigb_matrix_t A, int trans, const double *x, double *y, bool add)
{
    SIGB_REQUIRE(A && x && y, SIGB_ERR_ARG, "sigb_matvec: bad argument");
    if (A->mg) return mgpu_matvec(A, trans, x, y, add);
    SIGB_CHECK(check_fault("sigb_matvec"));
    const int64_t nx = trans ? A->nrow : A->ncol, ny = trans ? A->ncol : A->nrow;
    cudaStream_t st = ctx().stream;
    double *xd = nullptr, *yd = nullptr;
    SIGB_CHECK(ensure_stage(0, (size_t)nx, &xd));
    SIGB_CHECK(ensure_stage(1, (size_t)ny, &yd));
    SIGB_CUDA(cudaMemcpyAsync(xd, x, sizeof(double) * (size_t)nx, cudaMemcpyHostToDevice, st));
    if (add) SIGB_CUDA(cudaMemcpyAsync(yd, y, sizeof(double) * (size_t)ny, cudaMemcpyHostToDevice, st));
    DotSpec none;
    SIGB_CHECK(matvec_dev(A, trans, xd, yd, MODE_SET, add, none));
    SIGB_CUDA(cudaMemcpyAsync(y, yd, sizeof(double) * (size_t)ny, cudaMemcpyDeviceToHost, st));
    SIGB_CUDA(cudaStreamSynchronize(st));
    return check_fault("sigb_matvec");
}

int sigb_matvec(sigb_matrix_t A, int trans, const double *x, double *y)
{
    return host_matvec(A, trans, x, y, false);
}

int sigb_matvec_add(sigb_matrix_t A, int trans, const double *x, double *y)
{
    return host_matvec(A, trans, x, y, true);
}

int sigb_matvec_dev(sigb_matrix_t A, int trans, const double *x_dev, double *y_dev, int add)
{
    SIGB_REQUIRE(A && x_dev && y_dev, SIGB_ERR_ARG, "sigb_matvec_dev: bad argument");
    SIGB_REQUIRE(!A->mg, SIGB_ERR_UNSUPPORTED, "sigb_matvec_dev: a multi-GPU operator takes host vectors (its rows live on several devices)");
    DotSpec none;
    return matvec_dev(A, trans, x_dev, y_dev, MODE_SET, add != 0, none);
}

int sigb_matvec_dot_dev(sigb_matrix_t A, const double *x_dev, double *y_dev, double *dot)
{
    SIGB_REQUIRE(A && x_dev && y_dev, SIGB_ERR_ARG, "sigb_matvec_dot_dev: bad argument");
    SIGB_REQUIRE(!A->mg, SIGB_ERR_UNSUPPORTED, "sigb_matvec_dot_dev: a multi-GPU operator takes host vectors");
    SIGB_REQUIRE(A->nrow == A->ncol, SIGB_ERR_NONSQUARE, "sigb_matvec_dot_dev: operator is not square");
    double *slot = ctx().partials + (size_t)kMaxGrid * kMaxDots * (kNumTickets - 1);
    DotSpec d;
    d.ndot = 1;
    d.u = x_dev;
    d.out[0] = slot;
    RedFuse rf;
    const bool fused = dist_red_fuse(A, &rf);   // peer-memory transport: the kernel's last CTA finishes the sum across the GPUs
    if (fused) d.red = &rf;
    SIGB_CHECK(solver_matvec(A, x_dev, y_dev, d, false));
    if (!fused) SIGB_CHECK(dist_allreduce(A, slot, 1));
    if (dot) {  // dot == NULL: leave everything asynchronous (timing loops)
        SIGB_CUDA(cudaMemcpyAsync(dot, slot, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
        SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
        return check_fault("sigb_matvec_dot_dev");
    }
    return SIGB_OK;
}

// ===========================================================================
// solvers
// ===========================================================================

static int make_solver(int kind, double tol, sigb_solver_t *out)
{
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(out, SIGB_ERR_ARG, "solver create: null output");
    sigb_solver_t s = new sigb_solver_s();
    s->kind = kind;
    s->tol = (tol < 0.0) ? 1e-16 : tol;
    s->params_set = true;
    *out = s;
    return SIGB_OK;
}

int sigb_cg_create(double tolerance, sigb_solver_t *s) { return make_solver(S_CG, tolerance, s); }
int sigb_bicgstab_create(double tolerance, sigb_solver_t *s) { return make_solver(S_BICGSTAB, tolerance, s); }
int sigb_jacobi_create(sigb_solver_t *s) { return make_solver(S_JACOBI, -1.0, s); }
int sigb_ldu_create(sigb_solver_t *s) { return make_solver(S_LDU, -1.0, s); }

int sigb_solver_set_params(sigb_solver_t s, double tolerance)
{
    SIGB_REQUIRE(s, SIGB_ERR_ARG, "sigb_solver_set_params: null solver");
    s->tol = (tolerance < 0.0) ? 1e-16 : tolerance;
    s->params_set = true;
    return SIGB_OK;
}

int sigb_solver_set_max_iterations(sigb_solver_t s, int64_t cap)
{
    SIGB_REQUIRE(s, SIGB_ERR_ARG, "sigb_solver_set_max_iterations: null solver");
    s->cap = cap;
    return SIGB_OK;
}

int sigb_solver_set_persistent(sigb_solver_t s, int mode)
{
    SIGB_REQUIRE(s && mode >= -1 && mode <= 1, SIGB_ERR_ARG, "sigb_solver_set_persistent: bad argument");
    s->persistent = mode;
    return SIGB_OK;
}

int sigb_solver_set_strict_order(sigb_solver_t s, int on)
{
    SIGB_REQUIRE(s, SIGB_ERR_ARG, "sigb_solver_set_strict_order: null solver");
    s->strict_order = on ? 1 : 0;
    return SIGB_OK;
}

int sigb_solver_setup(sigb_solver_t s, sigb_matrix_t A)
{
    SIGB_REQUIRE(s && A, SIGB_ERR_ARG, "sigb_solver_setup: bad argument");
    if (A->mg) return mgpu_solver_setup(s, A);
    SIGB_REQUIRE(s->sub.empty(), SIGB_ERR_STATE, "sigb_solver_setup: solver was set up on a multi-GPU operator");
    const int64_t nglob_rows = A->dist ? dist_global_n(A) : A->nrow;
    const int64_t nglob_cols = A->dist ? dist_global_n(A) : A->ncol;
    if (nglob_rows != nglob_cols) {
        const char *what = s->kind == S_CG ? "a CG" : (s->kind == S_BICGSTAB ? "a BiCG-Stab" : (s->kind == S_LDU ? "an LDU" : "a Jacobi"));
        set_error("Cannot make %s solver for a non-square matrix", what);
        return SIGB_ERR_NONSQUARE;
    }
    const int nwork = s->kind == S_CG ? 4 : (s->kind == S_BICGSTAB ? 8 : (s->kind == S_LDU ? 0 : 1));
    // work vectors start on 256-byte boundaries (128-bit accesses in the vector phases)
    const int64_t nvec = (((int64_t)A->nrow + dist_halo_len(A)) + 31) & ~31LL;
    if (s->initialized && (s->nvec != nvec || s->nwork != nwork)) {
        cudaFree(s->work);
        s->work = nullptr;
        s->initialized = false;
    }
    s->nn = A->nrow;
    s->nvec = nvec;
    s->nwork = nwork;
    s->iterations = 0;
    s->A = A;
    if (!s->initialized) {
        SIGB_CHECK(dev_alloc(&s->work, (size_t)nvec * nwork));
        if (!s->state) {
            SIGB_CUDA(cudaMalloc((void **)&s->state, kstate_bytes()));
            SIGB_CUDA(cudaMallocHost((void **)&s->state_host, kstate_bytes()));
        }
        s->initialized = true;
    }
    if (s->kind == S_JACOBI) return jacobi_setup_dev(s, A);
    if (s->kind == S_LDU) {
        const int rc = ldu_setup_dev(s, A);
        if (rc != SIGB_OK) {
            cudaFree(s->work);
            s->work = nullptr;
            s->initialized = false;
        }
        return rc;
    }
    SIGB_CUDA(cudaMemsetAsync(s->work, 0, sizeof(double) * (size_t)nvec * nwork, ctx().stream));
    return SIGB_OK;
}

int sigb_solver_solve_dev(sigb_solver_t s, sigb_matrix_t A, double *x_dev, const double *b_dev,
                          sigb_solver_t pc)
{
    SIGB_REQUIRE(s && A && x_dev && b_dev, SIGB_ERR_ARG, "sigb_solver_solve: bad argument");
    SIGB_REQUIRE(!A->mg, SIGB_ERR_UNSUPPORTED, "sigb_solver_solve_dev: a multi-GPU operator takes host vectors");
    SIGB_REQUIRE(s->initialized, SIGB_ERR_STATE, "sigb_solver_solve: solver%%setup(A) has not been called");
    SIGB_REQUIRE(s->nn == A->nrow, SIGB_ERR_ARG, "sigb_solver_solve: solver was set up for nn = %d, operator has %d rows",
                 s->nn, A->nrow);
    if (pc) {
        // linear_solve_pc takes any linear_solver as pc (linear_operator_interface.f90:238-254): jacobi and
        // ldu behind cg and bicgstab (ldu on one GPU only: its sweeps are not row-sharded)
        const bool ldu_ok = pc->kind == S_LDU && !A->dist && (s->kind == S_CG || s->kind == S_BICGSTAB);
        SIGB_REQUIRE(pc->kind == S_JACOBI || ldu_ok, SIGB_ERR_UNSUPPORTED,
                     "sigb_solver_solve: the device preconditioners are jacobi and ldu (ldu on one GPU only)");
        SIGB_REQUIRE(pc->initialized && pc->nn == s->nn, SIGB_ERR_STATE,
                     "sigb_solver_solve: pc%%setup(A) has not been called");
    }
    SIGB_REQUIRE(!s->strict_order || (!A->dist && !(pc && pc->kind == S_LDU)), SIGB_ERR_UNSUPPORTED,
                 "sigb_solver_solve: strict-order dot products are a one-GPU parity aid (cg, bicgstab, jacobi pc)");
    switch (s->kind) {
    case S_CG: return cg_solve_dev(s, A, x_dev, b_dev, pc);
    case S_BICGSTAB: return bicgstab_solve_dev(s, A, x_dev, b_dev, pc);
    case S_LDU: return ldu_apply_dev(s, x_dev, b_dev, nullptr);   // ldu_solve ldu_solvers.f90:160-176
    default:
        // jacobi has no linear_solve_pc override: the default ignores pc
        // (linear_operator_interface.f90:238-254)
        return jacobi_apply_dev(s, x_dev, b_dev);
    }
}

int sigb_solver_solve(sigb_solver_t s, sigb_matrix_t A, double *x, const double *b, sigb_solver_t pc)
{
    SIGB_REQUIRE(s && A && x && b, SIGB_ERR_ARG, "sigb_solver_solve: bad argument");
    if (A->mg) return mgpu_solver_solve(s, A, x, b, pc);
    SIGB_REQUIRE(s->initialized, SIGB_ERR_STATE, "sigb_solver_solve: solver%%setup(A) has not been called");
    const int64_t n = s->nn;
    SIGB_CHECK(ensure_xb(s, 2 * n));
    cudaStream_t st = ctx().stream;
    double *xd = s->xb, *bd = s->xb + n;
    SIGB_CUDA(cudaMemcpyAsync(xd, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    SIGB_CUDA(cudaMemcpyAsync(bd, b, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    SIGB_CHECK(sigb_solver_solve_dev(s, A, xd, bd, pc));
    SIGB_CUDA(cudaMemcpyAsync(x, xd, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    SIGB_CUDA(cudaStreamSynchronize(st));
    return SIGB_OK;
}

int sigb_solver_get_info(sigb_solver_t s, int64_t *iterations, double *res2, int *capped)
{
    SIGB_REQUIRE(s, SIGB_ERR_ARG, "sigb_solver_get_info: null solver");
    if (iterations) *iterations = s->iterations;
    if (res2) *res2 = s->res2;
    if (capped) *capped = s->capped;
    return SIGB_OK;
}

int sigb_solver_get_vector(sigb_solver_t s, const char *name, double *out)
{
    SIGB_REQUIRE(s && name && out && s->initialized, SIGB_ERR_ARG, "sigb_solver_get_vector: bad argument");
    if (!s->sub.empty()) return mgpu_solver_get_vector(s, name, out);
    static const char *cg_names[] = {"p", "q", "r", "z"};
    static const char *bi_names[] = {"p", "q", "r", "r0", "v", "s", "t", "z"};
    int idx = -1;
    if (s->kind == S_JACOBI && !strcmp(name, "idiag")) idx = 0;
    if (s->kind == S_CG)
        for (int k = 0; k < 4; k++) if (!strcmp(name, cg_names[k])) idx = k;
    if (s->kind == S_BICGSTAB)
        for (int k = 0; k < 8; k++) if (!strcmp(name, bi_names[k])) idx = k;
    SIGB_REQUIRE(idx >= 0, SIGB_ERR_ARG, "sigb_solver_get_vector: no work vector '%s'", name);
    SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
    SIGB_CUDA(cudaMemcpy(out, s->work + (size_t)idx * s->nvec, sizeof(double) * (size_t)s->nn, cudaMemcpyDeviceToHost));
    return SIGB_OK;
}

int sigb_solver_destroy(sigb_solver_t s)
{
    if (!s) return SIGB_OK;
    mgpu_solver_free(s);
    cudaFree(s->work);
    cudaFree(s->state);
    cudaFreeHost(s->state_host);
    cudaFree(s->xb);
    cudaFree(s->bar);
    cudaFree(s->pers_partials);
    ldu_destroy_dev(s);
    delete s;
    return SIGB_OK;
}

int sigb_ldu_get_sizes(sigb_solver_t s, int32_t *n, int64_t *nL, int64_t *nU, int32_t *n_forward_levels,
                       int32_t *n_backward_levels)
{
    SIGB_REQUIRE(s && s->kind == S_LDU, SIGB_ERR_ARG, "sigb_ldu_get_sizes: not an ldu solver");
    return ldu_sizes(s, n, nL, nU, n_forward_levels, n_backward_levels);
}

int sigb_ldu_get_factors(sigb_solver_t s, int32_t *Lptr1, int32_t *Lnode1, double *Lval, int32_t *Uptr1,
                         int32_t *Unode1, double *Uval, double *D)
{
    SIGB_REQUIRE(s && s->kind == S_LDU, SIGB_ERR_ARG, "sigb_ldu_get_factors: not an ldu solver");
    return ldu_read(s, Lptr1, Lnode1, Lval, Uptr1, Unode1, Uval, D);
}

// ===========================================================================
// eigensolver
// ===========================================================================

// B == nullptr: lanczos / eigensolve; else generalized_lanczos / generalized_eigensolve
// with `bs` (+ optional `bpc`) as the solver attached to B.
static int lanczos_common(sigb_matrix_t A, int32_t n, const double *q1, uint64_t seed, double *T,
                          double *Q, double *lambda, bool ritz, sigb_matrix_t B = nullptr,
                          sigb_solver_t bs = nullptr, sigb_solver_t bpc = nullptr)
{
    SIGB_REQUIRE(A && n >= 1, SIGB_ERR_ARG, "lanczos: bad argument");
    if (B) {
        SIGB_REQUIRE(bs && bs->initialized && bs->nn == B->nrow, SIGB_ERR_STATE,
                     "generalized_lanczos: B has no solver set up (call B%%set_solver first)");
        SIGB_REQUIRE(B->nrow == A->nrow && B->ncol == A->ncol, SIGB_ERR_ARG, "generalized_lanczos: A and B differ in shape");
    }
    const int64_t nglob = A->dist ? dist_global_n(A) : A->ncol;
    SIGB_REQUIRE((A->dist ? dist_global_n(A) : A->nrow) == nglob, SIGB_ERR_NONSQUARE, "lanczos: operator is not square");
    const int64_t nr = A->nrow;
    cudaStream_t st = ctx().stream;
    double *Qd = nullptr, *Td = nullptr, *w = nullptr, *q1d = nullptr, *V2 = nullptr, *small = nullptr;
    double *gv = nullptr, *gz = nullptr;
    void *kst = nullptr;
    int rc = SIGB_OK;
    auto cleanup = [&]() {
        cudaFree(Qd); cudaFree(Td); cudaFree(w); cudaFree(q1d); cudaFree(V2); cudaFree(small); cudaFree(kst);
        cudaFree(gv); cudaFree(gz);
    };
#define LZ_TRY(expr) do { rc = (expr); if (rc != SIGB_OK) { cleanup(); return rc; } } while (0)
#define LZ_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cleanup(); return cuda_fail(e_, #expr, __FILE__, __LINE__); } } while (0)
    LZ_TRY(dev_alloc(&Qd, (size_t)nr * n));
    LZ_TRY(dev_alloc(&Td, (size_t)3 * n));
    LZ_TRY(dev_alloc(&w, (size_t)nr));
    LZ_CUDA(cudaMalloc(&kst, kstate_bytes()));
    LZ_CUDA(cudaMemsetAsync(kst, 0, kstate_bytes(), st));
    if (q1) {
        LZ_TRY(dev_alloc(&q1d, (size_t)nr));
        LZ_CUDA(cudaMemcpyAsync(q1d, q1, sizeof(double) * (size_t)nr, cudaMemcpyHostToDevice, st));
    }
    if (B) {
        LZ_TRY(dev_alloc(&gv, (size_t)nr));
        LZ_TRY(dev_alloc(&gz, (size_t)3 * nr));
        LZ_TRY(generalized_lanczos_dev(A, B, bs, bpc, n, q1d, seed, A->dist ? dist_row_offset(A) : 0, Td, Qd, w, gv,
                                       gz, (KState *)kst));
    } else {
        LZ_TRY(lanczos_dev(A, n, q1d, seed, A->dist ? dist_row_offset(A) : 0, Td, Qd, w, (KState *)kst));
    }
    std::vector<double> Th((size_t)3 * n);
    LZ_CUDA(cudaMemcpyAsync(Th.data(), Td, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, st));
    LZ_CUDA(cudaStreamSynchronize(st));
    if (T) memcpy(T, Th.data(), sizeof(double) * 3 * (size_t)n);
    if (ritz) {
        std::vector<double> d(n), e(n), Z((size_t)n * n);
        for (int i = 0; i < n; i++) d[i] = Th[3 * (size_t)i + 1];
        for (int i = 0; i < n - 1; i++) e[i] = Th[3 * (size_t)i + 2];
        const int info = tridiag_eig_host(n, d.data(), e.data(), Z.data());
        if (info != 0) {
            cleanup();
            set_error("eigensolve: tridiagonal QL did not converge");
            return SIGB_ERR_STATE;
        }
        LZ_TRY(dev_alloc(&V2, (size_t)nr * n));
        LZ_TRY(dev_alloc(&small, (size_t)n * n + n));
        LZ_CUDA(cudaMemcpyAsync(small, Z.data(), sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice, st));
        LZ_TRY(ritz_vectors_dev(A, Qd, V2, small, nr, n, B ? nullptr : small + (size_t)n * n));
        if (lambda) memcpy(lambda, d.data(), sizeof(double) * (size_t)n);
    }
    if (Q) LZ_CUDA(cudaMemcpyAsync(Q, Qd, sizeof(double) * (size_t)nr * n, cudaMemcpyDeviceToHost, st));
    LZ_CUDA(cudaStreamSynchronize(st));
    cleanup();
#undef LZ_TRY
#undef LZ_CUDA
    return SIGB_OK;
}

int sigb_lanczos(sigb_matrix_t A, int32_t n, const double *q1, uint64_t seed, double *T, double *Q)
{
    SIGB_REQUIRE(T && Q, SIGB_ERR_ARG, "sigb_lanczos: T and Q are required");
    if (A && A->mg) return mgpu_lanczos(A, n, q1, seed, T, Q, nullptr, false);
    return lanczos_common(A, n, q1, seed, T, Q, nullptr, false);
}

// lanczos with everything resident on the device: Q_dev (nrow x n, column-major like the Fortran
// array) and the optional start vector are device pointers, only T (3 n doubles) comes back to the
// host.  What a caller that keeps the Lanczos basis on the device uses -- and what bench.py times
// (the host-pointer form above spends most of its time copying Q back).
int sigb_lanczos_dev(sigb_matrix_t A, int32_t n, const double *q1_dev, uint64_t seed, double *T, double *Q_dev)
{
    if (A && A->mg) { ::sigb::set_error("sigb_lanczos_dev: not available for a single-process multi-GPU operator"); return SIGB_ERR_UNSUPPORTED; }
    SIGB_REQUIRE(A && n >= 1 && T && Q_dev, SIGB_ERR_ARG, "sigb_lanczos_dev: bad argument");
    const int64_t nglob = A->dist ? dist_global_n(A) : A->ncol;
    SIGB_REQUIRE((A->dist ? dist_global_n(A) : A->nrow) == nglob, SIGB_ERR_NONSQUARE, "lanczos: operator is not square");
    const int64_t nr = A->nrow;
    cudaStream_t st = ctx().stream;
    double *Td = nullptr, *w = nullptr;
    void *kst = nullptr;
    auto cleanup = [&]() { tmp_free(Td); tmp_free(w); tmp_free(kst); };
    cudaError_t e = tmp_alloc(&Td, (size_t)3 * n);
    if (e == cudaSuccess) e = tmp_alloc(&w, (size_t)nr);
    if (e == cudaSuccess) e = tmp_alloc_bytes(&kst, kstate_bytes());
    if (e == cudaSuccess) e = cudaMemsetAsync(kst, 0, kstate_bytes(), st);
    if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "lanczos scratch", __FILE__, __LINE__); }
    int rc = lanczos_dev(A, n, q1_dev, seed, A->dist ? dist_row_offset(A) : 0, Td, Q_dev, w, (KState *)kst);
    if (rc == SIGB_OK) {
        e = cudaMemcpyAsync(T, Td, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "lanczos T", __FILE__, __LINE__);
    }
    cleanup();
    if (rc == SIGB_OK) rc = check_fault("sigb_lanczos_dev");
    return rc;
}

int sigb_eigensolve(sigb_matrix_t A, int32_t n, const double *q1, uint64_t seed, double *lambda, double *V)
{
    SIGB_REQUIRE(lambda && V, SIGB_ERR_ARG, "sigb_eigensolve: lambda and V are required");
    if (A && A->mg) return mgpu_lanczos(A, n, q1, seed, nullptr, V, lambda, true);
    return lanczos_common(A, n, q1, seed, nullptr, V, lambda, true);
}

int sigb_generalized_lanczos(sigb_matrix_t A, sigb_matrix_t B, sigb_solver_t b_solver, sigb_solver_t b_pc,
                             int32_t n, const double *q1, uint64_t seed, double *T, double *Q)
{
    if (A && A->mg) { ::sigb::set_error("sigb_generalized_lanczos: not available for a single-process multi-GPU operator"); return SIGB_ERR_UNSUPPORTED; }
    SIGB_REQUIRE(B && T && Q, SIGB_ERR_ARG, "sigb_generalized_lanczos: B, T and Q are required");
    return lanczos_common(A, n, q1, seed, T, Q, nullptr, false, B, b_solver, b_pc);
}

int sigb_generalized_eigensolve(sigb_matrix_t A, sigb_matrix_t B, sigb_solver_t b_solver, sigb_solver_t b_pc,
                                int32_t n, const double *q1, uint64_t seed, double *lambda, double *V)
{
    if (A && A->mg) { ::sigb::set_error("sigb_generalized_eigensolve: not available for a single-process multi-GPU operator"); return SIGB_ERR_UNSUPPORTED; }
    SIGB_REQUIRE(B && lambda && V, SIGB_ERR_ARG, "sigb_generalized_eigensolve: B, lambda and V are required");
    SIGB_REQUIRE(!(A && A->dist), SIGB_ERR_UNSUPPORTED, "sigb_generalized_eigensolve on a row-sharded operator");
    return lanczos_common(A, n, q1, seed, nullptr, V, lambda, true, B, b_solver, b_pc);
}

}  // extern "C"
