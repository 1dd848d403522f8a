// ldu.cu -- ILDU(0) preconditioner on the device.
//
// Replaces the bodies of
//   sparse_ldu_setup                          src/solver/ldu_solvers.f90:95-130
//   incomplete_ldu_sparsity_pattern           :396-441   (host index work, ldu_host.cpp)
//   sparse_static_pattern_ldu_factorization   :275-387
//   ldu_solve, lower_/upper_triangular_solve  :160-176, :208-263
//
// A ~= (I + L) D (I + U) with L, U csr matrices on the strict triangles of A's
// pattern.  Both the elimination and the triangular solves are sequential in
// the reference; their only dependencies are "row i needs rows k < i that are
// lower neighbours of i" (resp. upper neighbours for the backward solve), so
// rows are grouped into levels on the host once per pattern and each level is
// one launch in which every row runs the reference's own arithmetic, entry by
// entry in stored order, with rounded products (no FMA).  Factors, diagonal
// and every solve are therefore bit-identical to the serial loops.
//
// This is latency-bound by construction: a level of the 2-D five-point stencil
// in natural ordering is one anti-diagonal of the grid (2N - 1 levels of at
// most N rows).  The value of the row is the drop-in coverage of the only
// other preconditioner the reference ships, not bandwidth.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "device_utils.cuh"
#include "dist.h"
#include "solvers.h"

namespace sigb {

struct LduInfo {
    int32_t n = 0;
    int64_t nL = 0, nU = 0, ne = 0;
    // patterns (1-based, device) and the combined value array [ Lval | Uval | D ]
    int32_t *Lptr = nullptr, *Lnode = nullptr, *Uptr = nullptr, *Unode = nullptr;
    double *fac = nullptr;
    int64_t *dest = nullptr;          // source entry -> offset in fac
    int32_t *frows = nullptr, *brows = nullptr;
    std::vector<int32_t> flev, blev;  // level pointers (host): launches are driven from here
    sigb_matrix_t rows = nullptr;     // csc / ellpack sources: their row form (device copy), else null
    // EXPERIMENTAL sync-free sweeps (SIGB_LDU_SYNCFREE=1): solution entries in flight as
    // payload+flag words, per-row "factored" flags, and the sequence number of the last sweep
    RedEntry *xs = nullptr;
    unsigned *ready = nullptr;
    unsigned sf_seq = 0;
    double *Lval() const { return fac; }
    double *Uval() const { return fac + nL; }
    double *D() const { return fac + nL + nU; }
};

namespace {

inline int grid_for(int64_t n)
{
    int64_t g = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)ctx().num_sms * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// "Copy A into L, D, U" (:307-324)
__global__ void __launch_bounds__(kThreads)
ldu_scatter_kernel(const double *__restrict__ aval, const int64_t *__restrict__ dest, int64_t ne,
                   double *__restrict__ fac)
{
    for (int64_t e = blockIdx.x * (int64_t)kThreads + threadIdx.x; e < ne; e += (int64_t)gridDim.x * kThreads)
        fac[dest[e]] = aval[e];
}

// M%get_value(i, j) on a csr pattern: last hit wins, 0 when absent (cs_matrices.f90:709-724)
__device__ __forceinline__ double get_value(const int32_t *ptr1, const int32_t *node1, const double *val, int32_t i,
                                            int32_t j)
{
    double z = 0.0;
    for (int32_t k = ptr1[i - 1] - 1; k < ptr1[i] - 1; k++)
        if (node1[k] == j) z = val[k];
    return z;
}

// rows of one level of the elimination (:331-381), one thread per row
__global__ void __launch_bounds__(kThreads)
ldu_factor_level_kernel(const int32_t *__restrict__ rows, int32_t count, const int32_t *__restrict__ Lptr,
                        const int32_t *__restrict__ Lnode, double *Lval, const int32_t *__restrict__ Uptr,
                        const int32_t *__restrict__ Unode, double *Uval, double *D)
{
    for (int32_t t = blockIdx.x * kThreads + threadIdx.x; t < count; t += gridDim.x * kThreads) {
        const int32_t i = rows[t];
        const int32_t lb = Lptr[i - 1] - 1, dl = Lptr[i] - 1 - lb;
        const int32_t ub = Uptr[i - 1] - 1, du = Uptr[i] - 1 - ub;
        for (int32_t ind1 = 0; ind1 < dl; ind1++) {
            const int32_t k = Lnode[lb + ind1];
            double Lik = Lval[lb + ind1];                              // L%get_value(i, k)   :342
            const double Uki = get_value(Uptr, Unode, Uval, k, i);     // :343
            const double Dk = D[k - 1];
            Lik = Lik / Dk;                                            // :345-346
            Lval[lb + ind1] = Lik;
            const double LikDk = mul(Lik, Dk);
            for (int32_t ind2 = 0; ind2 < dl; ind2++) {                // :350-358
                const int32_t j = Lnode[lb + ind2];
                if (j > k) {
                    const double Ukj = get_value(Uptr, Unode, Uval, k, j);
                    Lval[lb + ind2] = add(Lval[lb + ind2], -mul(LikDk, Ukj));
                }
            }
            D[i - 1] = sub(D[i - 1], mul(LikDk, Uki));                 // :361
            for (int32_t ind2 = 0; ind2 < du; ind2++) {                // :364-368
                const int32_t j = Unode[ub + ind2];
                const double Ukj = get_value(Uptr, Unode, Uval, k, j);
                Uval[ub + ind2] = add(Uval[ub + ind2], -mul(LikDk, Ukj));
            }
        }
        const double Di = D[i - 1];
        for (int32_t ind2 = 0; ind2 < du; ind2++) Uval[ub + ind2] = Uval[ub + ind2] / Di;   // :373-377
    }
}

// rows of one level of lower_/upper_triangular_solve (:226-235, :254-263):
// z = x(i) ; z = z - M%val(k) * x(node(k)) in stored order ; x(i) = z
__global__ void __launch_bounds__(kThreads)
tri_level_kernel(const int32_t *__restrict__ rows, int32_t count, const int32_t *__restrict__ ptr1,
                 const int32_t *__restrict__ node1, const double *__restrict__ val, double *x, const int *skip)
{
    if (skip != nullptr && *skip != 0) return;
    for (int32_t t = blockIdx.x * kThreads + threadIdx.x; t < count; t += gridDim.x * kThreads) {
        const int32_t i = rows[t];
        double z = x[i - 1];
        for (int32_t k = ptr1[i - 1] - 1; k < ptr1[i] - 1; k++) z = sub(z, mul(val[k], x[node1[k] - 1]));
        x[i - 1] = z;
    }
}

// x = b, then the rows of forward level 0 need nothing else; x = x / D between the sweeps
__global__ void __launch_bounds__(kThreads)
copy_kernel(const double *__restrict__ b, double *__restrict__ x, int64_t n, const int *skip)
{
    if (skip != nullptr && *skip != 0) return;
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        x[i] = b[i];
}
__global__ void __launch_bounds__(kThreads)
divide_kernel(double *__restrict__ x, const double *__restrict__ D, int64_t n, const int *skip)
{
    if (skip != nullptr && *skip != 0) return;
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        x[i] = x[i] / D[i];
}

// ---------------------------------------------------------------------------
// EXPERIMENTAL, opt-in (SIGB_LDU_SYNCFREE=1; compiled in, not the default path,
// not yet run on a GPU): the same sweeps WITHOUT one launch per level.
//
// One cooperative launch per sweep.  Threads take the rows in level order
// (frows / brows: every row a row depends on sits at an earlier position) and
// wait for exactly the entries they read instead of for a whole level:
//   * triangular solves: a finished x(i) is published as two 8-byte words, each
//     32 payload bits + the 32-bit sequence number of this sweep (the words of
//     the all-reduce, device_utils.cuh).  A word is delivered as a unit, so a
//     reader needs no fence and no second round trip: one L2 hop per
//     dependency, against one launch (4 us) or one grid barrier (1-2 us) per level;
//   * factorisation: a row reads whole rows of U and D(k) of its lower
//     neighbours, so finished rows are announced by a flag behind a fence and
//     read through L2.
// Lanes never spin on their own: a warp runs one loop in which every unfinished
// lane polls once and advances as far as it can, until all 32 are finished --
// a dependency on a lower lane of the same warp resolves on the next trip.
// Progress: all CTAs are resident (cooperative launch) and each warp takes its
// positions in ascending order, so the lowest unfinished position of the sweep
// always belongs to a running warp and has all its inputs.  Spins are bounded:
// a wrong answer the tests catch, never a hung GPU.
// Every row still does the reference's arithmetic in stored order with rounded
// products, so the factors and the solves stay bit-identical to the serial loops.
// ---------------------------------------------------------------------------
constexpr unsigned kSweepSpinLimit = 1u << 22;   // >= 1 s of polling: far beyond a whole sweep

// payload+flag words as in device_utils.cuh, but at GPU scope: these sweeps never leave the device,
// so the accesses need not be ordered against the peers (system scope) like the all-reduce's
__device__ __forceinline__ void st_word_gpu(unsigned int *p, unsigned int payload, unsigned int flag)
{
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(flag) : "memory");
}
__device__ __forceinline__ uint2 ld_word_gpu(const unsigned int *p)
{
    uint2 r;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}

__device__ __forceinline__ bool ll_try(const RedEntry *e, unsigned seq, double *out)
{
    const uint2 lo = ld_word_gpu(&e->lo), hi = ld_word_gpu(&e->hi);
    if (lo.y != seq || hi.y != seq) return false;
    *out = __longlong_as_double((long long)(((unsigned long long)hi.x << 32) | lo.x));
    return true;
}
__device__ __forceinline__ void ll_publish(RedEntry *e, unsigned seq, double v)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    st_word_gpu(&e->lo, (unsigned)bits, seq);
    st_word_gpu(&e->hi, (unsigned)(bits >> 32), seq);
}

// (I + M) x = src, or with D: (I + M) x = src / D   (the x = x / D statement of ldu_solve
// :169 folded into the start of the backward sweep: same division, same operands)
template <bool DIVIDE>
__global__ void __launch_bounds__(kThreads)
tri_syncfree_kernel(const int32_t *__restrict__ rows, int32_t n, const int32_t *__restrict__ ptr1,
                    const int32_t *__restrict__ node1, const double *__restrict__ val, const double *src,
                    const double *__restrict__ D, double *x, RedEntry *xs, unsigned seq, unsigned sleep_ns,
                    const int *skip)
{
    if (skip != nullptr && *skip != 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t base = blockIdx.x * (int64_t)kThreads + (threadIdx.x - lane); base < n; base += stride) {
        const int64_t p = base + lane;
        int32_t i = 0, k = 0, e = 0;
        double z = 0.0;
        bool done = true;
        if (p < n) {
            i = rows[p];
            k = ptr1[i - 1] - 1;
            e = ptr1[i] - 1;
            z = DIVIDE ? src[i - 1] / D[i - 1] : src[i - 1];
            done = false;
        }
        unsigned spins = 0;
        for (;;) {
            if (!done) {
                while (k < e) {                                   // z = z - M%val(k) * x(node(k)), stored order
                    double xj;
                    if (!ll_try(xs + (node1[k] - 1), seq, &xj)) break;
                    z = sub(z, mul(val[k], xj));
                    k++;
                }
                if (k == e) {
                    x[i - 1] = z;
                    ll_publish(xs + (i - 1), seq, z);
                    done = true;
                }
            }
            if (__all_sync(0xffffffffu, done)) break;
            if (++spins > kSweepSpinLimit) break;
            if (sleep_ns) __nanosleep(sleep_ns << (spins < 3u ? spins : 3u));   // back off: base, 2x, 4x, 8x
        }
    }
}

// U%get_value(k, j) on a row another thread of this launch has written: read at L2
__device__ __forceinline__ double get_value_cg(const int32_t *ptr1, const int32_t *node1, const double *val, int32_t i,
                                               int32_t j)
{
    double z = 0.0;
    for (int32_t k = ptr1[i - 1] - 1; k < ptr1[i] - 1; k++)
        if (node1[k] == j) z = __ldcg(val + k);
    return z;
}

// the elimination (:331-381) in one launch: the body of ldu_factor_level_kernel behind a wait
// for the "factored" flags of the row's lower neighbours
__global__ void __launch_bounds__(kThreads, 2)   // (room for registers: at the default bounds ptxas spilled)
ldu_factor_syncfree_kernel(const int32_t *__restrict__ rows, int32_t n, const int32_t *__restrict__ Lptr,
                           const int32_t *__restrict__ Lnode, double *Lval, const int32_t *__restrict__ Uptr,
                           const int32_t *__restrict__ Unode, double *Uval, double *D, unsigned *ready, unsigned seq,
                           unsigned sleep_ns)
{
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t base = blockIdx.x * (int64_t)kThreads + (threadIdx.x - lane); base < n; base += stride) {
        const int64_t p = base + lane;
        int32_t i = 0, lb = 0, dl = 0, w = 0;
        bool done = true;
        if (p < n) {
            i = rows[p];
            lb = Lptr[i - 1] - 1;
            dl = Lptr[i] - 1 - lb;
            done = false;
        }
        unsigned spins = 0;
        for (;;) {
            if (!done) {
                while (w < dl && *reinterpret_cast<volatile unsigned *>(ready + (Lnode[lb + w] - 1)) == seq) w++;
                if (w == dl) {
                    __threadfence();   // the neighbours' rows were written before their flags
                    const int32_t ub = Uptr[i - 1] - 1, du = Uptr[i] - 1 - ub;
                    for (int32_t ind1 = 0; ind1 < dl; ind1++) {
                        const int32_t k = Lnode[lb + ind1];
                        double Lik = Lval[lb + ind1];
                        const double Uki = get_value_cg(Uptr, Unode, Uval, k, i);
                        const double Dk = __ldcg(D + (k - 1));
                        Lik = Lik / Dk;
                        Lval[lb + ind1] = Lik;
                        const double LikDk = mul(Lik, Dk);
                        for (int32_t ind2 = 0; ind2 < dl; ind2++) {
                            const int32_t j = Lnode[lb + ind2];
                            if (j > k) {
                                const double Ukj = get_value_cg(Uptr, Unode, Uval, k, j);
                                Lval[lb + ind2] = add(Lval[lb + ind2], -mul(LikDk, Ukj));
                            }
                        }
                        D[i - 1] = sub(D[i - 1], mul(LikDk, Uki));
                        for (int32_t ind2 = 0; ind2 < du; ind2++) {
                            const int32_t j = Unode[ub + ind2];
                            const double Ukj = get_value_cg(Uptr, Unode, Uval, k, j);
                            Uval[ub + ind2] = add(Uval[ub + ind2], -mul(LikDk, Ukj));
                        }
                    }
                    const double Di = D[i - 1];
                    for (int32_t ind2 = 0; ind2 < du; ind2++) Uval[ub + ind2] = Uval[ub + ind2] / Di;
                    __threadfence();   // row i of U and D(i) before the flag
                    *reinterpret_cast<volatile unsigned *>(ready + (i - 1)) = seq;
                    done = true;
                }
            }
            if (__all_sync(0xffffffffu, done)) break;
            if (++spins > kSweepSpinLimit) break;
            if (sleep_ns) __nanosleep(sleep_ns << (spins < 3u ? spins : 3u));   // back off: base, 2x, 4x, 8x
        }
    }
}

struct SyncFreeCfg {
    bool on = false;
    int ctas_per_sm = 1;     // SIGB_LDU_SF_CTAS_PER_SM
    int ctas = 0;            // SIGB_LDU_SF_CTAS: absolute grid size (wins when > 0); fewer resident threads
                             // = fewer pollers competing with the wavefront for L2
    unsigned sleep_ns = 0;   // SIGB_LDU_SF_SLEEP_NS: base of the polling back-off (0 = poll flat out)
};
const SyncFreeCfg &syncfree_cfg()
{
    static SyncFreeCfg c;
    static bool read = false;
    if (!read) {
        c.on = env_int("SIGB_LDU_SYNCFREE", 0) == 1;
        c.ctas_per_sm = std::max(1, env_int("SIGB_LDU_SF_CTAS_PER_SM", c.ctas_per_sm));
        c.ctas = std::max(0, env_int("SIGB_LDU_SF_CTAS", 0));
        c.sleep_ns = (unsigned)std::max(0, env_int("SIGB_LDU_SF_SLEEP_NS", 0));
        read = true;
    }
    return c;
}

// grid of a sweep kernel: the requested CTAs per SM, never more than can be resident
template <typename K>
int syncfree_grid(K kernel, int64_t n, int *grid)
{
    int per_sm = 0;
    SIGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0));
    if (per_sm < 1) per_sm = 1;
    const int64_t resident = (int64_t)per_sm * ctx().num_sms;   // what a cooperative launch can hold
    if (per_sm > syncfree_cfg().ctas_per_sm) per_sm = syncfree_cfg().ctas_per_sm;
    int64_t g = (int64_t)per_sm * ctx().num_sms;
    if (syncfree_cfg().ctas > 0) g = std::min<int64_t>(syncfree_cfg().ctas, resident);
    const int64_t need = (n + kThreads - 1) / kThreads;
    if (g > need) g = need;
    if (g < 1) g = 1;
    *grid = (int)g;
    return SIGB_OK;
}

// scratch of the sync-free sweeps; hands out the next sequence number (0 = never written)
int syncfree_prepare(LduInfo *F, unsigned *seq)
{
    cudaStream_t st = ctx().stream;
    const size_t n = (size_t)(F->n > 0 ? F->n : 1);
    if (!F->xs) {
        SIGB_CUDA(cudaMalloc((void **)&F->xs, sizeof(RedEntry) * n));
        SIGB_CUDA(cudaMalloc((void **)&F->ready, sizeof(unsigned) * n));
        F->sf_seq = 0;
    }
    if (F->sf_seq == 0 || F->sf_seq == 0xffffffffu) {   // first use, or the counter is about to wrap
        SIGB_CUDA(cudaMemsetAsync(F->xs, 0, sizeof(RedEntry) * n, st));
        SIGB_CUDA(cudaMemsetAsync(F->ready, 0, sizeof(unsigned) * n, st));
        F->sf_seq = 0;
    }
    *seq = ++F->sf_seq;
    return SIGB_OK;
}

template <bool DIVIDE>
int launch_tri_syncfree(LduInfo *F, const int32_t *rows, const int32_t *ptr1, const int32_t *node1, const double *val,
                        const double *src, double *x, const int *skip)
{
    unsigned seq = 0;
    SIGB_CHECK(syncfree_prepare(F, &seq));
    int grid = 0;
    SIGB_CHECK(syncfree_grid(tri_syncfree_kernel<DIVIDE>, F->n, &grid));
    int32_t n = F->n;
    const double *D = F->D();
    RedEntry *xs = F->xs;
    unsigned sleep_ns = syncfree_cfg().sleep_ns;
    void *params[] = {(void *)&rows, (void *)&n, (void *)&ptr1, (void *)&node1, (void *)&val, (void *)&src,
                      (void *)&D, (void *)&x, (void *)&xs, (void *)&seq, (void *)&sleep_ns, (void *)&skip};
    SIGB_CUDA(cudaLaunchCooperativeKernel((const void *)tri_syncfree_kernel<DIVIDE>, dim3(grid), dim3(kThreads), params,
                                          0, ctx().stream));
    count_launch();
    return SIGB_OK;
}

int launch_factor_syncfree(LduInfo *F)
{
    unsigned seq = 0;
    SIGB_CHECK(syncfree_prepare(F, &seq));
    int grid = 0;
    SIGB_CHECK(syncfree_grid(ldu_factor_syncfree_kernel, F->n, &grid));
    const int32_t *rows = F->frows, *Lptr = F->Lptr, *Lnode = F->Lnode, *Uptr = F->Uptr, *Unode = F->Unode;
    int32_t n = F->n;
    double *Lval = F->Lval(), *Uval = F->Uval(), *D = F->D();
    unsigned *ready = F->ready;
    unsigned sleep_ns = syncfree_cfg().sleep_ns;
    void *params[] = {(void *)&rows, (void *)&n, (void *)&Lptr, (void *)&Lnode, (void *)&Lval, (void *)&Uptr,
                      (void *)&Unode, (void *)&Uval, (void *)&D, (void *)&ready, (void *)&seq, (void *)&sleep_ns};
    SIGB_CUDA(cudaLaunchCooperativeKernel((const void *)ldu_factor_syncfree_kernel, dim3(grid), dim3(kThreads), params,
                                          0, ctx().stream));
    count_launch();
    return SIGB_OK;
}

void free_ldu(LduInfo *F)
{
    if (!F) return;
    cudaFree(F->xs); cudaFree(F->ready);
    cudaFree(F->Lptr); cudaFree(F->Lnode); cudaFree(F->Uptr); cudaFree(F->Unode);
    cudaFree(F->fac); cudaFree(F->dest); cudaFree(F->frows); cudaFree(F->brows);
    if (F->rows) sigb_matrix_destroy(F->rows);
    delete F;
}

template <typename T>
int upload(T **dst, const T *src, size_t count)
{
    SIGB_CUDA(cudaMalloc((void **)dst, sizeof(T) * (count > 0 ? count : 1)));
    if (count > 0)
        SIGB_CUDA(cudaMemcpyAsync(*dst, src, sizeof(T) * count, cudaMemcpyHostToDevice, ctx().stream));
    return SIGB_OK;
}

}  // namespace

void ldu_destroy_dev(sigb_solver_t s)
{
    free_ldu(s->ldu);
    s->ldu = nullptr;
}

// pc%setup(A): the pattern and the schedules once (solver%initialized, :113-121),
// the numeric factorisation every time (:124-126)
int ldu_setup_dev(sigb_solver_t s, sigb_matrix_t A)
{
    SIGB_REQUIRE(!A->op && !A->dist, SIGB_ERR_UNSUPPORTED,
                 "ldu setup needs a stored sparse matrix (sparse_ldu_setup selects on sparse_matrix_interface, "
                 "ldu_solvers.f90:111-112)");
    cudaStream_t st = ctx().stream;
    const int32_t n = A->nrow;
    LduInfo *F = s->ldu;
    if (F && (F->n != n || F->ne != A->g->ne)) {   // another operator: start over
        free_ldu(F);
        F = s->ldu = nullptr;
    }
    // the rows of A in iteration order: a csr_matrix as stored, otherwise its csr copy
    sigb_matrix_t R = A;
    if (A->g->kind != G_CSR) {
        if (F && F->rows) { sigb_matrix_destroy(F->rows); F->rows = nullptr; }
        sigb_matrix_t rows = nullptr;
        SIGB_CHECK(sigb_matrix_copy(A, SIGB_FMT_CSR, 0, &rows));
        R = rows;
    }
    int rc = SIGB_OK;
    if (!F) {
        F = new LduInfo();
        F->n = n;
        F->ne = R->g->ne;
        const int64_t ne = F->ne;
        std::vector<int32_t> ptr((size_t)n + 1), node((size_t)ne);
        std::vector<int32_t> Lptr((size_t)n + 1), Uptr((size_t)n + 1), Lnode((size_t)ne), Unode((size_t)ne);
        std::vector<int64_t> dest((size_t)ne);
        std::vector<int32_t> frows((size_t)n), brows((size_t)n), flev((size_t)n + 1), blev((size_t)n + 1);
        int32_t nf = 0, nb = 0;
        cudaError_t e = cudaStreamSynchronize(st);
        if (e == cudaSuccess)
            e = cudaMemcpy(ptr.data(), R->g->stored.ptr, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && ne > 0)
            e = cudaMemcpy(node.data(), R->g->stored.node, sizeof(int32_t) * (size_t)ne, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "ldu setup read-back", __FILE__, __LINE__);
        if (rc == SIGB_OK)
            rc = sigb_ldu_symbolic(n, ptr.data(), node.data(), Lptr.data(), Lnode.data(), Uptr.data(), Unode.data(),
                                   dest.data(), frows.data(), flev.data(), &nf, brows.data(), blev.data(), &nb);
        if (rc == SIGB_OK) {
            F->nL = (int64_t)Lptr[(size_t)n] - 1;
            F->nU = (int64_t)Uptr[(size_t)n] - 1;
            F->flev.assign(flev.begin(), flev.begin() + nf + 1);
            F->blev.assign(blev.begin(), blev.begin() + nb + 1);
            rc = upload(&F->Lptr, Lptr.data(), (size_t)n + 1);
            if (rc == SIGB_OK) rc = upload(&F->Uptr, Uptr.data(), (size_t)n + 1);
            if (rc == SIGB_OK) rc = upload(&F->Lnode, Lnode.data(), (size_t)F->nL);
            if (rc == SIGB_OK) rc = upload(&F->Unode, Unode.data(), (size_t)F->nU);
            if (rc == SIGB_OK) rc = upload(&F->dest, dest.data(), (size_t)ne);
            if (rc == SIGB_OK) rc = upload(&F->frows, frows.data(), (size_t)n);
            if (rc == SIGB_OK) rc = upload(&F->brows, brows.data(), (size_t)n);
            if (rc == SIGB_OK) {
                cudaError_t e2 = cudaMalloc((void **)&F->fac, sizeof(double) * (size_t)(F->nL + F->nU + n + 1));
                if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);   // the host vectors go out of scope below
                if (e2 != cudaSuccess) rc = cuda_fail(e2, "ldu setup", __FILE__, __LINE__);
            }
        }
        if (rc != SIGB_OK) {
            free_ldu(F);
            if (R != A) sigb_matrix_destroy(R);
            return rc;
        }
        s->ldu = F;
    }
    if (R != A) F->rows = R;

    // L%zero(), U%zero(), D = 0, then copy A in (:302-324)
    SIGB_CUDA(cudaMemsetAsync(F->fac, 0, sizeof(double) * (size_t)(F->nL + F->nU + n + 1), st));
    if (F->ne > 0) {
        ldu_scatter_kernel<<<grid_for(F->ne), kThreads, 0, st>>>(R->val, F->dest, F->ne, F->fac);
        count_launch();
    }
    if (syncfree_cfg().on && n > 0) {   // EXPERIMENTAL: the elimination in one launch
        SIGB_CHECK(launch_factor_syncfree(F));
        SIGB_CUDA(cudaGetLastError());
        return SIGB_OK;
    }
    // the elimination, level by level
    const int nlev = (int)F->flev.size() - 1;
    for (int l = 0; l < nlev; l++) {
        const int32_t b = F->flev[(size_t)l], cnt = F->flev[(size_t)l + 1] - b;
        ldu_factor_level_kernel<<<grid_for(cnt), kThreads, 0, st>>>(F->frows + b, cnt, F->Lptr, F->Lnode, F->Lval(),
                                                                   F->Uptr, F->Unode, F->Uval(), F->D());
        count_launch();
    }
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

// call pc%solve(A, x, b): x = b ; (I + L) x = x ; x = x / D ; (I + U) x = x   (:167-171)
int ldu_apply_dev(sigb_solver_t s, double *x, const double *b, const int *skip_flag)
{
    LduInfo *F = s->ldu;
    SIGB_REQUIRE(F, SIGB_ERR_STATE, "ldu solve: pc%%setup(A) has not been called");
    cudaStream_t st = ctx().stream;
    const int32_t n = F->n;
    if (syncfree_cfg().on && n > 0) {
        // EXPERIMENTAL: two launches; x = b and x = x / D are folded into the sweeps' first reads
        SIGB_CHECK(launch_tri_syncfree<false>(F, F->frows, F->Lptr, F->Lnode, F->Lval(), b, x, skip_flag));
        SIGB_CHECK(launch_tri_syncfree<true>(F, F->brows, F->Uptr, F->Unode, F->Uval(), x, x, skip_flag));
        SIGB_CUDA(cudaGetLastError());
        return SIGB_OK;
    }
    if (x != b) {
        copy_kernel<<<grid_for(n), kThreads, 0, st>>>(b, x, n, skip_flag);
        count_launch();
    }
    // level 0 of either sweep has no neighbours to subtract: nothing to do for it
    for (size_t l = 1; l + 1 < F->flev.size(); l++) {
        const int32_t b0 = F->flev[l], cnt = F->flev[l + 1] - b0;
        tri_level_kernel<<<grid_for(cnt), kThreads, 0, st>>>(F->frows + b0, cnt, F->Lptr, F->Lnode, F->Lval(), x,
                                                            skip_flag);
        count_launch();
    }
    divide_kernel<<<grid_for(n), kThreads, 0, st>>>(x, F->D(), n, skip_flag);
    count_launch();
    for (size_t l = 1; l + 1 < F->blev.size(); l++) {
        const int32_t b0 = F->blev[l], cnt = F->blev[l + 1] - b0;
        tri_level_kernel<<<grid_for(cnt), kThreads, 0, st>>>(F->brows + b0, cnt, F->Uptr, F->Unode, F->Uval(), x,
                                                            skip_flag);
        count_launch();
    }
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

int ldu_sizes(sigb_solver_t s, int32_t *n, int64_t *nL, int64_t *nU, int32_t *nflev, int32_t *nblev)
{
    LduInfo *F = s->ldu;
    SIGB_REQUIRE(F, SIGB_ERR_STATE, "ldu: pc%%setup(A) has not been called");
    if (n) *n = F->n;
    if (nL) *nL = F->nL;
    if (nU) *nU = F->nU;
    if (nflev) *nflev = (int32_t)F->flev.size() - 1;
    if (nblev) *nblev = (int32_t)F->blev.size() - 1;
    return SIGB_OK;
}

int ldu_read(sigb_solver_t s, int32_t *Lptr, int32_t *Lnode, double *Lval, int32_t *Uptr, int32_t *Unode,
             double *Uval, double *D)
{
    LduInfo *F = s->ldu;
    SIGB_REQUIRE(F, SIGB_ERR_STATE, "ldu: pc%%setup(A) has not been called");
    SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
    const size_t n = (size_t)F->n, nL = (size_t)F->nL, nU = (size_t)F->nU;
    if (Lptr) SIGB_CUDA(cudaMemcpy(Lptr, F->Lptr, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost));
    if (Uptr) SIGB_CUDA(cudaMemcpy(Uptr, F->Uptr, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost));
    if (Lnode && nL) SIGB_CUDA(cudaMemcpy(Lnode, F->Lnode, sizeof(int32_t) * nL, cudaMemcpyDeviceToHost));
    if (Unode && nU) SIGB_CUDA(cudaMemcpy(Unode, F->Unode, sizeof(int32_t) * nU, cudaMemcpyDeviceToHost));
    if (Lval && nL) SIGB_CUDA(cudaMemcpy(Lval, F->Lval(), sizeof(double) * nL, cudaMemcpyDeviceToHost));
    if (Uval && nU) SIGB_CUDA(cudaMemcpy(Uval, F->Uval(), sizeof(double) * nU, cudaMemcpyDeviceToHost));
    if (D && n) SIGB_CUDA(cudaMemcpy(D, F->D(), sizeof(double) * n, cudaMemcpyDeviceToHost));
    return SIGB_OK;
}

}  // namespace sigb
