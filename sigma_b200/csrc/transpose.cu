// transpose.cu -- one-time, on-device, STABLE transposes and re-layouts.
//
// csc_matvec_add (src/matrix/formats/cs_matrices.f90:627-647) scatters
// y(node(k)) += val(k) * x(j) column by column; on a GPU that is an atomic
// free-for-all with a run-dependent summation order.  Instead the pattern is
// transposed once into CSR with a counting sort that keeps, inside every
// target row, the entries in ascending SOURCE ENTRY INDEX.  Source entries are
// laid out line by line (column by column for a CSC matrix), so that order is
// exactly the order in which the reference loop adds contributions into each
// y(i); the streaming CSR kernel (kernels_spmv.cu) then reproduces the
// reference result bit for bit.  The same routine provides A^T for
// matvec_t on CSR / ELLPACK matrices (reference: cs_matvec_t_add
// cs_matrices.f90:513-521, ellpack_matvec_t_add ellpack_matrices.f90:670-693).
//
// The transposed pattern equals what the reference itself builds for a
// transposed copy -- cs_graph_build with trans = .true.
// (src/graph/formats/cs_graphs.f90:109-197: first-free-slot insertion while
// walking the source edges in storage order) -- so the index arrays are checked
// bit-exactly against the oracle's restatement of that routine.
//
// Steps: histogram of targets -> exclusive scan (1-based ptr) -> unordered
// placement with atomics -> per-row sort of the placed source indices (rows
// are short; the sort makes the result deterministic; the few rows longer than
// kLongRow are listed and rank-sorted by one CTA each) -> gather of line ids.
#include "device_utils.cuh"

namespace sigb {

namespace {

constexpr int kScanItems = 2048;  // items per CTA in the scan kernels

__global__ void __launch_bounds__(kThreads)
hist_cs_kernel(const int32_t *__restrict__ node1, int64_t ne, int32_t *__restrict__ cnt)
{
    for (int64_t e = blockIdx.x * (int64_t)kThreads + threadIdx.x; e < ne;
         e += (int64_t)gridDim.x * kThreads)
        atomicAdd(&cnt[node1[e] - 1], 1);
}

// ell: entry (line i, slot k) lives at node_sm[k * n_pad + i]
__global__ void __launch_bounds__(kThreads)
hist_ell_kernel(const int32_t *__restrict__ node_sm, int32_t n, int32_t n_pad, int32_t max_d,
                int32_t *__restrict__ cnt)
{
    const int64_t tot = (int64_t)n * max_d;
    for (int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x; t < tot;
         t += (int64_t)gridDim.x * kThreads) {
        const int32_t k = (int32_t)(t / n), i = (int32_t)(t % n);
        atomicAdd(&cnt[node_sm[(size_t)k * n_pad + i] - 1], 1);
    }
}

// --- three-kernel exclusive scan of int32 counts into a 1-based ptr ---------
__global__ void __launch_bounds__(kThreads)
scan_block_sums(const int32_t *__restrict__ cnt, int64_t n, int64_t *__restrict__ block_sum)
{
    __shared__ int64_t sm[kThreads];
    const int64_t base = (int64_t)blockIdx.x * kScanItems;
    int64_t s = 0;
    for (int k = threadIdx.x; k < kScanItems; k += kThreads)
        if (base + k < n) s += cnt[base + k];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int off = kThreads / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sum[blockIdx.x] = sm[0];
}

__global__ void scan_block_offsets(int64_t *block_sum, int nblocks)
{
    // tiny: one thread turns the CTA sums into exclusive offsets
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t run = 0;
        for (int b = 0; b < nblocks; b++) {
            const int64_t v = block_sum[b];
            block_sum[b] = run;
            run += v;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
scan_apply(const int32_t *__restrict__ cnt, int64_t n, const int64_t *__restrict__ block_off,
           int32_t *__restrict__ ptr1)
{
    constexpr int per = kScanItems / kThreads;
    __shared__ int64_t sm[kThreads];
    const int64_t base = (int64_t)blockIdx.x * kScanItems + (int64_t)threadIdx.x * per;
    int64_t loc[per];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < per; k++) {
        loc[k] = s;
        if (base + k < n) s += cnt[base + k];
    }
    sm[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over the thread sums
    for (int off = 1; off < kThreads; off <<= 1) {
        int64_t v = (threadIdx.x >= off) ? sm[threadIdx.x - off] : 0;
        __syncthreads();
        sm[threadIdx.x] += v;
        __syncthreads();
    }
    const int64_t excl = sm[threadIdx.x] - s + block_off[blockIdx.x];
#pragma unroll
    for (int k = 0; k < per; k++)
        if (base + k < n) ptr1[base + k] = (int32_t)(excl + loc[k] + 1);
    if (base <= n && n < base + per) ptr1[n] = (int32_t)(excl + loc[n - base] + 1);
}

__global__ void __launch_bounds__(kThreads)
place_cs_kernel(const int32_t *__restrict__ node1, int64_t ne,
                const int32_t *__restrict__ ptr_t1, int32_t *__restrict__ cursor,
                int32_t *__restrict__ perm)
{
    for (int64_t e = blockIdx.x * (int64_t)kThreads + threadIdx.x; e < ne;
         e += (int64_t)gridDim.x * kThreads) {
        const int32_t tgt = node1[e] - 1;
        const int32_t slot = atomicAdd(&cursor[tgt], 1);
        perm[ptr_t1[tgt] - 1 + slot] = (int32_t)e;
    }
}

// ell source entry index is LINE-major (i * max_d + k): the order the reference
// loop visits entries (ellpack_matrices.f90:682-689).
__global__ void __launch_bounds__(kThreads)
place_ell_kernel(const int32_t *__restrict__ node_sm, int32_t n, int32_t n_pad, int32_t max_d,
                 const int32_t *__restrict__ ptr_t1, int32_t *__restrict__ cursor,
                 int32_t *__restrict__ perm)
{
    const int64_t tot = (int64_t)n * max_d;
    for (int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x; t < tot;
         t += (int64_t)gridDim.x * kThreads) {
        const int32_t k = (int32_t)(t / n), i = (int32_t)(t % n);
        const int32_t tgt = node_sm[(size_t)k * n_pad + i] - 1;
        const int32_t slot = atomicAdd(&cursor[tgt], 1);
        perm[ptr_t1[tgt] - 1 + slot] = i * max_d + k;
    }
}

// Sort each target row's source indices ascending.  Short rows: insertion sort
// by one thread.  Long rows (> kLongRow) are left to sort_long_rows_kernel.
constexpr int kLongRow = 96;

__global__ void __launch_bounds__(kThreads)
sort_rows_kernel(const int32_t *__restrict__ ptr_t1, int32_t nrows, int32_t *__restrict__ perm,
                 int32_t *__restrict__ long_rows, int32_t *__restrict__ n_long)
{
    for (int32_t r = blockIdx.x * kThreads + threadIdx.x; r < nrows; r += gridDim.x * kThreads) {
        const int32_t b = ptr_t1[r] - 1, e = ptr_t1[r + 1] - 1;
        if (e - b > kLongRow) {
            // left to sort_long_rows_kernel, which only visits the rows listed here
            // (their order in the list does not matter: each is sorted on its own)
            long_rows[atomicAdd(n_long, 1)] = r;
            continue;
        }
        for (int32_t a = b + 1; a < e; a++) {
            const int32_t key = perm[a];
            int32_t c = a - 1;
            while (c >= b && perm[c] > key) {
                perm[c + 1] = perm[c];
                c--;
            }
            perm[c + 1] = key;
        }
    }
}

// One CTA per long row: rank sort (keys are distinct) through a scratch copy.
__global__ void __launch_bounds__(kThreads)
sort_long_rows_kernel(const int32_t *__restrict__ ptr_t1, const int32_t *__restrict__ long_rows, int32_t n_long,
                      int32_t *__restrict__ perm, int32_t *__restrict__ scratch)
{
    for (int32_t l = blockIdx.x; l < n_long; l += gridDim.x) {
        const int32_t r = long_rows[l];
        const int32_t b = ptr_t1[r] - 1, e = ptr_t1[r + 1] - 1;
        for (int32_t a = b + threadIdx.x; a < e; a += kThreads) scratch[a] = perm[a];
        __syncthreads();
        for (int32_t a = b + threadIdx.x; a < e; a += kThreads) {
            const int32_t key = scratch[a];
            int32_t rank = 0;
            for (int32_t c = b; c < e; c++) rank += (scratch[c] < key);
            perm[b + rank] = key;
        }
        __syncthreads();
    }
}

// node_t[pos] = 1-based line id of the source entry perm[pos]
__global__ void __launch_bounds__(kThreads)
lines_cs_kernel(const int32_t *__restrict__ ptr1, int32_t nlines, const int32_t *__restrict__ perm,
                int64_t ne, int32_t *__restrict__ node_t1)
{
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < ne;
         p += (int64_t)gridDim.x * kThreads) {
        const int32_t e1 = perm[p] + 1;  // 1-based entry index
        // largest line l (0-based) with ptr1[l] <= e1
        int32_t lo = 0, hi = nlines - 1;
        while (lo < hi) {
            const int32_t mid = (lo + hi + 1) >> 1;
            if (ptr1[mid] <= e1) lo = mid; else hi = mid - 1;
        }
        node_t1[p] = lo + 1;
    }
}

__global__ void __launch_bounds__(kThreads)
lines_ell_kernel(const int32_t *__restrict__ perm, int64_t ne, int32_t max_d,
                 int32_t *__restrict__ node_t1)
{
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < ne;
         p += (int64_t)gridDim.x * kThreads)
        node_t1[p] = perm[p] / max_d + 1;
}

__global__ void __launch_bounds__(kThreads)
gather_kernel(const double *__restrict__ val, const int32_t *__restrict__ perm, int64_t ne,
              double *__restrict__ out)
{
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < ne;
         p += (int64_t)gridDim.x * kThreads)
        out[p] = val[perm[p]];
}

__global__ void __launch_bounds__(kThreads)
gather_ell_kernel(const double *__restrict__ val_sm, const int32_t *__restrict__ perm, int64_t ne,
                  int32_t n_pad, int32_t max_d, double *__restrict__ out)
{
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < ne;
         p += (int64_t)gridDim.x * kThreads) {
        const int32_t s = perm[p];
        out[p] = val_sm[(size_t)(s % max_d) * n_pad + s / max_d];
    }
}

// line-major (as the Fortran holds node(max_d, n)) -> slot-major, padded rows
// [n, n_pad) get node = 1 / val = 0 so the kernels need no bounds test.
template <typename T>
__global__ void __launch_bounds__(kThreads)
relayout_kernel(const T *__restrict__ src_cm, int32_t n, int32_t n_pad, int32_t max_d, T pad,
                T *__restrict__ dst_sm)
{
    const int64_t tot = (int64_t)n_pad * max_d;
    for (int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x; t < tot;
         t += (int64_t)gridDim.x * kThreads) {
        const int32_t k = (int32_t)(t / n_pad), i = (int32_t)(t % n_pad);
        dst_sm[t] = (i < n) ? src_cm[(size_t)i * max_d + k] : pad;
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) fill_kernel(T *p, int64_t n, T v)
{
    for (int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x; t < n;
         t += (int64_t)gridDim.x * kThreads)
        p[t] = v;
}

inline int grid_for(int64_t n)
{
    int64_t g = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)ctx().num_sms * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

// n counts -> n + 1 one-based offsets (also used by convert.cu)
int scan_to_ptr1(const int32_t *cnt, int64_t n, int32_t *ptr1)
{
    // n counts -> n + 1 one-based offsets
    const int nblocks = (int)((n + 1 + kScanItems - 1) / kScanItems);
    int64_t *block_sum = nullptr;
    SIGB_CUDA(tmp_alloc(&block_sum, (size_t)nblocks));
    cudaStream_t st = ctx().stream;
    scan_block_sums<<<nblocks, kThreads, 0, st>>>(cnt, n, block_sum);
    scan_block_offsets<<<1, 32, 0, st>>>(block_sum, nblocks);
    scan_apply<<<nblocks, kThreads, 0, st>>>(cnt, n, block_sum, ptr1);
    count_launch(3);
    SIGB_CUDA(cudaGetLastError());
    SIGB_CUDA(cudaStreamSynchronize(st));
    SIGB_CUDA(tmp_free(block_sum));
    return SIGB_OK;
}

namespace {

int finish_transpose(int32_t ntargets, int64_t ne, int32_t *ptr_t, int32_t *perm)
{
    cudaStream_t st = ctx().stream;
    // rows longer than kLongRow are listed by the first kernel; at most ne / kLongRow of them
    const int64_t max_long = ne / kLongRow + 1;
    int32_t *long_rows = nullptr, *n_long = nullptr;
    SIGB_CUDA(tmp_alloc(&long_rows, (size_t)max_long));
    SIGB_CUDA(tmp_alloc(&n_long, 1));
    SIGB_CUDA(cudaMemsetAsync(n_long, 0, sizeof(int32_t), st));
    sort_rows_kernel<<<grid_for(ntargets), kThreads, 0, st>>>(ptr_t, ntargets, perm, long_rows, n_long);
    count_launch();
    int32_t h_long = 0;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_long, n_long, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    int32_t *scratch = nullptr;
    if (e == cudaSuccess && h_long > 0) {
        e = tmp_alloc(&scratch, (size_t)(ne > 0 ? ne : 1));
        if (e == cudaSuccess) {
            const int grid = h_long < ctx().num_sms * 4 ? h_long : ctx().num_sms * 4;
            sort_long_rows_kernel<<<grid, kThreads, 0, st>>>(ptr_t, long_rows, h_long, perm, scratch);
            count_launch();
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    tmp_free(scratch);
    tmp_free(long_rows);
    tmp_free(n_long);
    if (e != cudaSuccess) return cuda_fail(e, "finish_transpose", __FILE__, __LINE__);
    return SIGB_OK;
}

}  // namespace

int fill_i32(int32_t *p, int64_t n, int32_t v)
{
    if (n <= 0) return SIGB_OK;
    fill_kernel<int32_t><<<grid_for(n), kThreads, 0, ctx().stream>>>(p, n, v);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

int fill_f64(double *p, int64_t n, double v)
{
    if (n <= 0) return SIGB_OK;
    fill_kernel<double><<<grid_for(n), kThreads, 0, ctx().stream>>>(p, n, v);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

int device_transpose_cs(const int32_t *ptr1, const int32_t *node1, int32_t nlines,
                        int32_t ntargets, int64_t ne, int32_t **ptr_t_out,
                        int32_t **node_t_out, int32_t **perm_out)
{
    cudaStream_t st = ctx().stream;
    int32_t *cnt = nullptr, *ptr_t = nullptr, *node_t = nullptr, *perm = nullptr;
    SIGB_CUDA(tmp_alloc(&cnt, (size_t)ntargets + 1));
    SIGB_CUDA(cudaMalloc(&ptr_t, sizeof(int32_t) * ((size_t)ntargets + 1 + kPad)));
    SIGB_CHECK(fill_i32(ptr_t + ntargets + 1, kPad, 1));
    SIGB_CUDA(cudaMalloc(&node_t, sizeof(int32_t) * ((size_t)ne + 8)));
    SIGB_CUDA(cudaMalloc(&perm, sizeof(int32_t) * ((size_t)ne + 8)));
    SIGB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * ((size_t)ntargets + 1), st));
    SIGB_CHECK(fill_i32(node_t + ne, 8, 1));
    SIGB_CHECK(fill_i32(perm + ne, 8, 0));
    if (ne > 0) {
        hist_cs_kernel<<<grid_for(ne), kThreads, 0, st>>>(node1, ne, cnt);
        count_launch();
    }
    SIGB_CHECK(scan_to_ptr1(cnt, ntargets, ptr_t));
    SIGB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * ((size_t)ntargets + 1), st));
    if (ne > 0) {
        place_cs_kernel<<<grid_for(ne), kThreads, 0, st>>>(node1, ne, ptr_t, cnt, perm);
        count_launch();
        SIGB_CHECK(finish_transpose(ntargets, ne, ptr_t, perm));
        lines_cs_kernel<<<grid_for(ne), kThreads, 0, st>>>(ptr1, nlines, perm, ne, node_t);
        count_launch();
    }
    SIGB_CUDA(cudaGetLastError());
    SIGB_CUDA(cudaStreamSynchronize(st));
    SIGB_CUDA(tmp_free(cnt));
    *ptr_t_out = ptr_t;
    *node_t_out = node_t;
    *perm_out = perm;
    return SIGB_OK;
}

int device_transpose_ell(const int32_t *node_sm, int32_t n, int32_t n_pad, int32_t max_d,
                         int32_t ntargets, int32_t **ptr_t_out, int32_t **node_t_out,
                         int32_t **perm_out)
{
    cudaStream_t st = ctx().stream;
    const int64_t ne = (int64_t)n * max_d;
    int32_t *cnt = nullptr, *ptr_t = nullptr, *node_t = nullptr, *perm = nullptr;
    SIGB_CUDA(tmp_alloc(&cnt, (size_t)ntargets + 1));
    SIGB_CUDA(cudaMalloc(&ptr_t, sizeof(int32_t) * ((size_t)ntargets + 1 + kPad)));
    SIGB_CHECK(fill_i32(ptr_t + ntargets + 1, kPad, 1));
    SIGB_CUDA(cudaMalloc(&node_t, sizeof(int32_t) * ((size_t)ne + 8)));
    SIGB_CUDA(cudaMalloc(&perm, sizeof(int32_t) * ((size_t)ne + 8)));
    SIGB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * ((size_t)ntargets + 1), st));
    SIGB_CHECK(fill_i32(node_t + ne, 8, 1));
    SIGB_CHECK(fill_i32(perm + ne, 8, 0));
    if (ne > 0) {
        hist_ell_kernel<<<grid_for(ne), kThreads, 0, st>>>(node_sm, n, n_pad, max_d, cnt);
        count_launch();
    }
    SIGB_CHECK(scan_to_ptr1(cnt, ntargets, ptr_t));
    SIGB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * ((size_t)ntargets + 1), st));
    if (ne > 0) {
        place_ell_kernel<<<grid_for(ne), kThreads, 0, st>>>(node_sm, n, n_pad, max_d, ptr_t, cnt, perm);
        count_launch();
        SIGB_CHECK(finish_transpose(ntargets, ne, ptr_t, perm));
        lines_ell_kernel<<<grid_for(ne), kThreads, 0, st>>>(perm, ne, max_d, node_t);
        count_launch();
    }
    SIGB_CUDA(cudaGetLastError());
    SIGB_CUDA(cudaStreamSynchronize(st));
    SIGB_CUDA(tmp_free(cnt));
    *ptr_t_out = ptr_t;
    *node_t_out = node_t;
    *perm_out = perm;
    return SIGB_OK;
}

int gather_values(const double *val, const int32_t *perm, int64_t ne, double *val_t)
{
    if (ne <= 0) return SIGB_OK;
    gather_kernel<<<grid_for(ne), kThreads, 0, ctx().stream>>>(val, perm, ne, val_t);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

int gather_values_ell(const double *val_sm, const int32_t *perm, int64_t ne, int32_t n_pad,
                      int32_t max_d, double *val_t)
{
    if (ne <= 0) return SIGB_OK;
    gather_ell_kernel<<<grid_for(ne), kThreads, 0, ctx().stream>>>(val_sm, perm, ne, n_pad, max_d, val_t);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

int ell_relayout_node(const int32_t *node_cm_dev, int32_t n, int32_t n_pad, int32_t max_d,
                      int32_t *node_sm)
{
    relayout_kernel<int32_t><<<grid_for((int64_t)n_pad * max_d), kThreads, 0, ctx().stream>>>(
        node_cm_dev, n, n_pad, max_d, 1, node_sm);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

int ell_relayout_val(const double *val_cm_dev, int32_t n, int32_t n_pad, int32_t max_d,
                     double *val_sm)
{
    relayout_kernel<double><<<grid_for((int64_t)n_pad * max_d), kThreads, 0, ctx().stream>>>(
        val_cm_dev, n, n_pad, max_d, 0.0, val_sm);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

}  // namespace sigb
