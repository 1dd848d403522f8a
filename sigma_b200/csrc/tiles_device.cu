// tiles_device.cu -- the row tiling of the streaming CSR kernel built on the device, so
// that device-side transposes and matrix copies need no read-back of `ptr` and no host
// loop (that read-back + loop was most of what a device copy cost in round 1; with the
// stream-ordered scratch of api.cu a transposing copy of 21 M entries went from 12.7 ms
// to 4.9 ms, profiles/r2_visit_a_1gpu_summary.txt).
//
// The tiling is the greedy one of build_tiles_host (kernels_spmv.cu): from row s a tile
// runs to next(s) = the largest e <= s + kTileRows with ptr(e) - ptr(s) <= kTileCap (at
// least s + 1).  The tile starts are the orbit 0, next(0), next(next(0)), ... -- sequential
// as written, but
//   1. next(s) is a bisection per row (ptr is monotone), all rows at once;
//   2. the orbit is marked by pointer doubling: with jump = next^(2^k), one round marks
//      jump(i) for every marked i and squares jump; after round k the first 2^(k+1) orbit
//      members are marked, so ~log2(#tiles) rounds suffice (rounds launched after
//      jump(0) reached the end return immediately);
//   3. the marked rows are compacted with the scan of transpose.cu into the tile table,
//      followed by the sub-table of tiles that hold entries.
// The result equals build_tiles_host bit for bit (tests/test_row_tiles.py: numpy replica of
// these steps against the greedy walk; tests/test_gpu_convert.py: device against host).
#include <algorithm>

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

// next[s] for s < n, next[n] = n; also the extreme line lengths (minmax[0] = max, [1] = min)
__global__ void __launch_bounds__(kThreads)
tile_next_kernel(const int32_t *__restrict__ ptr1, int32_t n, int32_t cap, int32_t rows, int32_t *__restrict__ next,
                 int32_t *minmax)
{
    int32_t dmax = 0, dmin = INT32_MAX;
    for (int32_t s = blockIdx.x * kThreads + threadIdx.x; s <= n; s += gridDim.x * kThreads) {
        if (s == n) { next[n] = n; continue; }
        const int64_t limit = (int64_t)ptr1[s] + cap;
        int32_t a = s + 1, b = (int32_t)min((int64_t)n, (int64_t)s + rows);
        if ((int64_t)ptr1[a] <= limit) {
            while (a < b) {
                const int32_t mid = (int32_t)(((int64_t)a + b + 1) >> 1);
                if ((int64_t)ptr1[mid] <= limit) a = mid; else b = mid - 1;
            }
        }
        next[s] = a;
        const int32_t d = ptr1[s + 1] - ptr1[s];
        dmax = max(dmax, d);
        dmin = min(dmin, d);
    }
    dmax = __reduce_max_sync(0xffffffffu, dmax);
    dmin = __reduce_min_sync(0xffffffffu, dmin);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&minmax[0], dmax);
        atomicMin(&minmax[1], dmin);
    }
}

__global__ void __launch_bounds__(kThreads)
tile_mark_kernel(const int32_t *__restrict__ jump, int32_t *mark, int32_t n, const int *finished)
{
    if (*finished != 0) return;
    for (int32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
        if (mark[i] != 0) {
            const int32_t j = jump[i];
            if (j < n) mark[j] = 1;     // only orbit members are ever marked, so early marks are harmless
        }
}

// jump_out = jump_in o jump_in; notes when jump_in(0) had already reached the end: the round
// that just ran has then marked the whole orbit
__global__ void __launch_bounds__(kThreads)
tile_double_kernel(const int32_t *__restrict__ jump_in, int32_t *__restrict__ jump_out, int32_t n, int *finished)
{
    if (*finished != 0) return;
    for (int32_t i = blockIdx.x * kThreads + threadIdx.x; i <= n; i += gridDim.x * kThreads)
        jump_out[i] = jump_in[jump_in[i]];
    if (blockIdx.x == 0 && threadIdx.x == 0 && jump_in[0] == n) *finished = 1;
}

__global__ void __launch_bounds__(kThreads)
tile_nonempty_kernel(const int32_t *__restrict__ mark, const int32_t *__restrict__ next,
                     const int32_t *__restrict__ ptr1, int32_t n, int32_t *__restrict__ flag)
{
    for (int32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
        flag[i] = (mark[i] != 0 && ptr1[next[i]] > ptr1[i]) ? 1 : 0;
}

__global__ void __launch_bounds__(kThreads)
tile_emit_kernel(const int32_t *__restrict__ mark, const int32_t *__restrict__ nonempty,
                 const int32_t *__restrict__ pos1, const int32_t *__restrict__ pos1_nonempty,
                 const int32_t *__restrict__ next, const int32_t *__restrict__ ptr1, int32_t n,
                 TileDesc *__restrict__ tiles, int32_t ntiles)
{
    for (int32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
        if (mark[i] != 0) {
            TileDesc d;
            d.rs = i;
            d.re = next[i];
            d.ks = ptr1[i] - 1;
            d.ke = ptr1[d.re] - 1;
            tiles[pos1[i] - 1] = d;
            if (nonempty[i] != 0) tiles[ntiles + pos1_nonempty[i] - 1] = d;
        }
}

}  // namespace

// Tile table (+ its sub-table of tiles with entries) of a pattern whose ptr lives on the
// device; nnz (known to every caller) selects the tile shape as on the host.  max_d / min_d: extreme line lengths (either may be null).
int build_tiles_device(const int32_t *ptr1_dev, int32_t nrows, int64_t nnz, CsrView &v, int32_t *max_d, int32_t *min_d)
{
    const TileShape shape = tile_shape_for(nnz, nrows);
    cudaStream_t st = ctx().stream;
    const size_t len = (size_t)nrows + 1;
    int32_t *next = nullptr, *jump_a = nullptr, *jump_b = nullptr, *mark = nullptr, *nonempty = nullptr;
    int32_t *pos = nullptr, *pos_ne = nullptr, *minmax = nullptr;
    int *finished = nullptr;
    TileDesc *tiles = nullptr;
    auto cleanup = [&]() {
        tmp_free(next); tmp_free(jump_a); tmp_free(jump_b); tmp_free(mark); tmp_free(nonempty);
        tmp_free(pos); tmp_free(pos_ne); tmp_free(minmax); tmp_free(finished);
    };
#define TD_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cleanup(); cudaFree(tiles); return cuda_fail(e_, #expr, __FILE__, __LINE__); } } while (0)
#define TD_TRY(expr) do { int rc_ = (expr); if (rc_ != SIGB_OK) { cleanup(); cudaFree(tiles); return rc_; } } while (0)
    TD_CUDA(tmp_alloc(&next, len));
    TD_CUDA(tmp_alloc(&jump_a, len));
    TD_CUDA(tmp_alloc(&jump_b, len));
    TD_CUDA(tmp_alloc(&mark, len));
    TD_CUDA(tmp_alloc(&nonempty, len));
    TD_CUDA(tmp_alloc(&pos, len));
    TD_CUDA(tmp_alloc(&pos_ne, len));
    TD_CUDA(tmp_alloc(&minmax, 2));
    TD_CUDA(tmp_alloc(&finished, 1));
    const int32_t h_init[2] = {0, INT32_MAX};
    TD_CUDA(cudaMemcpyAsync(minmax, h_init, sizeof(h_init), cudaMemcpyHostToDevice, st));
    TD_CUDA(cudaMemsetAsync(finished, 0, sizeof(int), st));
    TD_CUDA(cudaMemsetAsync(mark, 0, sizeof(int32_t) * len, st));
    int32_t ntiles = 0, n_nonempty = 0, h_minmax[2] = {0, 0};
    if (nrows > 0) {
        TD_TRY(fill_i32(mark, 1, 1));                                   // row 0 starts the first tile
        tile_next_kernel<<<grid_for((int64_t)nrows + 1), kThreads, 0, st>>>(ptr1_dev, nrows, shape.cap, shape.rows, next,
                                                                                      minmax);
        count_launch();
        TD_CUDA(cudaMemcpyAsync(jump_a, next, sizeof(int32_t) * len, cudaMemcpyDeviceToDevice, st));
        int rounds = 1;
        while ((1ll << rounds) <= (long long)nrows) rounds++;           // 2^rounds > nrows >= number of tiles
        int32_t *jin = jump_a, *jout = jump_b;
        for (int k = 0; k < rounds; k++) {
            tile_mark_kernel<<<grid_for(nrows), kThreads, 0, st>>>(jin, mark, nrows, finished);
            tile_double_kernel<<<grid_for((int64_t)nrows + 1), kThreads, 0, st>>>(jin, jout, nrows, finished);
            count_launch(2);
            std::swap(jin, jout);
        }
        tile_nonempty_kernel<<<grid_for(nrows), kThreads, 0, st>>>(mark, next, ptr1_dev, nrows, nonempty);
        count_launch();
        TD_CUDA(cudaGetLastError());
        TD_TRY(scan_to_ptr1(mark, nrows, pos));
        TD_TRY(scan_to_ptr1(nonempty, nrows, pos_ne));
        TD_CUDA(cudaMemcpyAsync(&ntiles, pos + nrows, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        TD_CUDA(cudaMemcpyAsync(&n_nonempty, pos_ne + nrows, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        TD_CUDA(cudaMemcpyAsync(h_minmax, minmax, sizeof(h_minmax), cudaMemcpyDeviceToHost, st));
        TD_CUDA(cudaStreamSynchronize(st));
        ntiles -= 1;
        n_nonempty -= 1;
    }
    TD_CUDA(cudaMalloc((void **)&tiles, sizeof(TileDesc) * (size_t)std::max(ntiles + n_nonempty, 1)));
    if (nrows > 0) {
        tile_emit_kernel<<<grid_for(nrows), kThreads, 0, st>>>(mark, nonempty, pos, pos_ne, next, ptr1_dev, nrows,
                                                              tiles, ntiles);
        count_launch();
        TD_CUDA(cudaGetLastError());
        TD_CUDA(cudaStreamSynchronize(st));
    }
#undef TD_CUDA
#undef TD_TRY
    cleanup();
    v.tiles = tiles;
    v.tile_nnz = shape.nnz;
    v.ntiles = ntiles;
    v.tiles_nonempty = tiles + ntiles;
    v.n_nonempty = n_nonempty;
    if (max_d) *max_d = nrows > 0 ? h_minmax[0] : 0;
    if (min_d) *min_d = nrows > 0 ? h_minmax[1] : 0;
    return SIGB_OK;
}

}  // namespace sigb

using namespace sigb;

extern "C" {

// Diagnostic: the device-built tiling of a pattern given by its host ptr (uploaded here),
// in the layout of sigb_debug_row_tiles.
int sigb_debug_row_tiles_dev(int32_t n, const int32_t *ptr1, int32_t *tiles, int32_t *ntiles)
{
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(n >= 0 && ptr1 && tiles && ntiles, SIGB_ERR_ARG, "sigb_debug_row_tiles_dev: bad argument");
    int32_t *pd = nullptr;
    SIGB_CUDA(cudaMalloc((void **)&pd, sizeof(int32_t) * ((size_t)n + 1)));
    cudaError_t e = cudaMemcpy(pd, ptr1, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(pd); return cuda_fail(e, "upload", __FILE__, __LINE__); }
    CsrView v;
    const int rc = build_tiles_device(pd, n, n > 0 ? (int64_t)ptr1[n] - 1 : 0, v, nullptr, nullptr);
    if (rc == SIGB_OK && v.ntiles > 0) {
        e = cudaMemcpy(tiles, v.tiles, sizeof(TileDesc) * (size_t)v.ntiles, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { cudaFree(pd); cudaFree(v.tiles); return cuda_fail(e, "read-back", __FILE__, __LINE__); }
    }
    if (rc == SIGB_OK) *ntiles = v.ntiles;
    cudaFree(pd);
    cudaFree(v.tiles);
    return rc;
}

}  // extern "C"
