// assemble.cu -- on-device matrix assembly (second half):
// a stream of `call A%add_value(i, j, z)` statements applied in order.
//
// Replaces the bodies of
//   csr_matrix_add_value / csc_matrix_add_value   src/matrix/formats/cs_matrices.f90:868-891, :924-947
//   ellpack_matrix_add_value                      src/matrix/formats/ellpack_matrices.f90:471-493
//   csr_matrix_add_multiple_values (the same statement over B(k, l))   cs_matrices.f90:934-967
// for entries that are already in the sparsity pattern -- what a finite-element
// assembly loop issues once its graph is built (examples/fem.f90:43-47).
//
// Floating-point addition is not associative, so "in order" is part of the
// contract: entry (i, j) must receive its contributions in ascending call
// index, exactly like the serial loop.  An atomicAdd scatter would give a run-
// dependent order.  Instead
//   1. locate : every call finds the stored position of (i, j) by scanning its
//               line, like the reference's own search (cs_matrices.f90:881-886);
//               calls that miss the pattern are counted and the batch is refused
//               (the reference would reallocate the graph, :888-890: out of scope);
//   2. group  : the calls are bucketed by stored position with the stable
//               counting sort of transpose.cu (ascending call index per bucket);
//   3. reduce : one thread per stored entry adds its bucket to val, in order.
// The result is bit-identical to the serial loop and independent of the launch
// shape.  Index work + one pass over the calls: HBM-bound, ~40 B per call.
#include <vector>

#include <chrono>
#include <stdio.h>

#include "internal.h"
#include "device_utils.cuh"

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

struct Miss {
    unsigned long long count;
    unsigned long long first;   // smallest call index that missed
};

// compressed lines: line = i (csr) or j (csc), id looked for = the other one;
// slot = 1-based position in the stored val array, 0 when absent
__global__ void __launch_bounds__(kThreads)
locate_cs_kernel(const int32_t *__restrict__ ptr1, const int32_t *__restrict__ node1, int32_t nlines, int32_t nids,
                 const int32_t *__restrict__ line_of, const int32_t *__restrict__ id_of, int64_t count,
                 int32_t *__restrict__ slot1, Miss *miss)
{
    for (int64_t c = blockIdx.x * (int64_t)kThreads + threadIdx.x; c < count;
         c += (int64_t)gridDim.x * kThreads) {
        const int32_t line = line_of[c], want = id_of[c];
        int32_t found = 0;
        if (line >= 1 && line <= nlines && want >= 1 && want <= nids)
            for (int32_t k = ptr1[line - 1]; k <= ptr1[line] - 1; k++)
                if (node1[k - 1] == want) found = k;
        slot1[c] = found;
        if (!found) {
            atomicAdd(&miss->count, 1ull);
            atomicMin(&miss->first, (unsigned long long)c);
        }
    }
}

// ellpack, slot-major on the device: row i's first degrees(i) slots are
// searched (ellpack_matrices.f90:484-489); position = k * n_pad + (i - 1) + 1
__global__ void __launch_bounds__(kThreads)
locate_ell_kernel(const int32_t *__restrict__ node_sm, const int32_t *__restrict__ degrees, int32_t n, int32_t m,
                  int32_t n_pad, const int32_t *__restrict__ ci, const int32_t *__restrict__ cj, int64_t count,
                  int32_t *__restrict__ slot1, Miss *miss)
{
    for (int64_t c = blockIdx.x * (int64_t)kThreads + threadIdx.x; c < count;
         c += (int64_t)gridDim.x * kThreads) {
        const int32_t i = ci[c], j = cj[c];
        int32_t found = 0;
        if (i >= 1 && i <= n && j >= 1 && j <= m) {
            const int32_t d = degrees[i - 1];
            for (int32_t k = 0; k < d; k++)
                if (node_sm[(size_t)k * n_pad + (i - 1)] == j) found = k * n_pad + (i - 1) + 1;
        }
        slot1[c] = found;
        if (!found) {
            atomicAdd(&miss->count, 1ull);
            atomicMin(&miss->first, (unsigned long long)c);
        }
    }
}

// val(s) = val(s) + z(c) for the calls c of bucket s, ascending c
__global__ void __launch_bounds__(kThreads)
reduce_buckets_kernel(const int32_t *__restrict__ bucket_ptr1, const int32_t *__restrict__ perm,
                      const double *__restrict__ cz, int64_t nslots, double *__restrict__ val)
{
    for (int64_t s = blockIdx.x * (int64_t)kThreads + threadIdx.x; s < nslots;
         s += (int64_t)gridDim.x * kThreads) {
        const int32_t b = bucket_ptr1[s] - 1, e = bucket_ptr1[s + 1] - 1;
        if (e > b) {
            double z = val[s];
            for (int32_t p = b; p < e; p++) z = add(z, cz[perm[p]]);
            val[s] = z;
        }
    }
}

}  // namespace

}  // namespace sigb

using namespace sigb;

extern "C" {

int sigb_matrix_add_values(sigb_matrix_t A, int64_t count, const int32_t *i1, const int32_t *j1, const double *z)
{
    if (A && A->mg) { ::sigb::set_error("sigb_matrix_add_values: not available for a single-process multi-GPU operator"); return SIGB_ERR_UNSUPPORTED; }
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(A && count >= 0 && (count == 0 || (i1 && j1 && z)), SIGB_ERR_ARG, "sigb_matrix_add_values: bad argument");
    SIGB_REQUIRE(!A->op && !A->dist, SIGB_ERR_UNSUPPORTED,
                 "sigb_matrix_add_values: the target must be a stored csr / csc / ellpack matrix");
    SIGB_REQUIRE(count <= INT32_MAX - 16, SIGB_ERR_ARG, "sigb_matrix_add_values: at most 2^31 calls per batch");
    if (count == 0) return SIGB_OK;
    sigb_graph_t g = A->g;
    cudaStream_t st = ctx().stream;
    const int64_t nslots = (g->kind == G_ELL) ? (int64_t)g->n_pad * g->max_d : g->ne;
    SIGB_REQUIRE(nslots <= INT32_MAX - 16, SIGB_ERR_ARG, "sigb_matrix_add_values: matrix too large for int32 positions");

    static const bool verbose = env_int("SIGB_VERBOSE", 0) == 1;
    const auto t_begin = std::chrono::steady_clock::now();
    auto stage = [&](const char *what) {
        if (!verbose) return;
        cudaStreamSynchronize(st);
        fprintf(stderr, "sigma_b200: add_values %-28s at %8.3f ms\n", what,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    int32_t *ci = nullptr, *cj = nullptr, *slot1 = nullptr, *one_line = nullptr;
    int32_t *bucket_ptr = nullptr, *unused_node = nullptr, *perm = nullptr;
    double *cz = nullptr;
    Miss *miss = nullptr;
    int rc = SIGB_OK;
    auto cleanup = [&]() {
        tmp_free(ci); tmp_free(cj); tmp_free(slot1); tmp_free(one_line); tmp_free(cz); tmp_free(miss);
        cudaFree(bucket_ptr); cudaFree(unused_node); cudaFree(perm);   // outputs of device_transpose_cs
    };
#define AS_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cleanup(); return cuda_fail(e_, #expr, __FILE__, __LINE__); } } while (0)
#define AS_TRY(expr) do { rc = (expr); if (rc != SIGB_OK) { cleanup(); return rc; } } while (0)
    AS_CUDA(tmp_alloc(&ci, (size_t)count));
    AS_CUDA(tmp_alloc(&cj, (size_t)count));
    AS_CUDA(tmp_alloc(&cz, (size_t)count));
    AS_CUDA(tmp_alloc(&slot1, (size_t)count + kPad));
    AS_CUDA(tmp_alloc(&miss, 1));
    AS_CUDA(cudaMemcpyAsync(ci, i1, sizeof(int32_t) * (size_t)count, cudaMemcpyHostToDevice, st));
    AS_CUDA(cudaMemcpyAsync(cj, j1, sizeof(int32_t) * (size_t)count, cudaMemcpyHostToDevice, st));
    AS_CUDA(cudaMemcpyAsync(cz, z, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, st));
    Miss h_miss;
    h_miss.count = 0;
    h_miss.first = ~0ull;
    AS_CUDA(cudaMemcpyAsync(miss, &h_miss, sizeof(Miss), cudaMemcpyHostToDevice, st));
    stage("scratch + host-to-device");

    // 1. locate
    if (g->kind == G_ELL)
        locate_ell_kernel<<<grid_for(count), kThreads, 0, st>>>(g->ell_node, g->ell_degrees, g->n, g->m, g->n_pad, ci, cj,
                                                               count, slot1, miss);
    else if (g->kind == G_CSR)
        locate_cs_kernel<<<grid_for(count), kThreads, 0, st>>>(g->stored.ptr, g->stored.node, g->n, g->m, ci, cj, count,
                                                              slot1, miss);
    else   // csc: column j holds the row ids
        locate_cs_kernel<<<grid_for(count), kThreads, 0, st>>>(g->stored.ptr, g->stored.node, g->n, g->m, cj, ci, count,
                                                              slot1, miss);
    count_launch();
    AS_CUDA(cudaGetLastError());
    AS_CUDA(cudaMemcpyAsync(&h_miss, miss, sizeof(Miss), cudaMemcpyDeviceToHost, st));
    AS_CUDA(cudaStreamSynchronize(st));
    if (h_miss.count != 0) {
        const long long c = (long long)h_miss.first;
        cleanup();
        set_error("sigb_matrix_add_values: %llu of %lld entries are not in the sparsity pattern (first: call %lld, "
                  "entry (%d, %d)); the reference would reallocate the graph (cs_matrices.f90:888-890), which the "
                  "device mirror does not do -- nothing was added",
                  h_miss.count, (long long)count, c + 1, i1[c], j1[c]);
        return SIGB_ERR_ARG;
    }

    stage("locate");
    // 2. group the calls by stored position: the stable transpose of a one-line
    //    "graph" whose ids are the positions
    AS_CUDA(tmp_alloc(&one_line, (size_t)(2 + kPad)));
    {
        int32_t h_line[2] = {1, (int32_t)(count + 1)};
        AS_CUDA(cudaMemcpyAsync(one_line, h_line, sizeof(h_line), cudaMemcpyHostToDevice, st));
        AS_CUDA(cudaStreamSynchronize(st));
    }
    AS_TRY(device_transpose_cs(one_line, slot1, 1, (int32_t)nslots, count, &bucket_ptr, &unused_node, &perm));

    stage("stable bucket sort");
    // 3. reduce, in call order
    reduce_buckets_kernel<<<grid_for(nslots), kThreads, 0, st>>>(bucket_ptr, perm, cz, nslots, A->val);
    count_launch();
    AS_CUDA(cudaGetLastError());
    AS_CUDA(cudaStreamSynchronize(st));
    A->val_t_valid = false;
    stage("ordered reduce");
    cleanup();
    stage("scratch released");
#undef AS_CUDA
#undef AS_TRY
    return SIGB_OK;
}

}  // extern "C"
