// kernels_spmv.cu -- sm_100a sparse matrix-vector kernels.
//
// Replaces the bodies of
//   csr_matvec_add      src/matrix/formats/cs_matrices.f90:600-622
//   csc_matvec_add      src/matrix/formats/cs_matrices.f90:627-647  (run on the
//                       device-built stable transpose, see transpose.cu)
//   ellpack_matvec_add  src/matrix/formats/ellpack_matrices.f90:640-665
// and fuses the zero-fill of linear_operator_matvec
// (src/linear_operator/linear_operator_interface.f90:185-194) and the Krylov
// dot products that directly follow a matvec (cg_solvers.f90:134-135,
// bicgstab_solvers.f90:159-160,163-164, eigensolver.f90:68-69).
//
// CSR kernel ("stream" layout).  Rows are grouped into tiles of at most 2045 stored
// entries and 512 rows -- 1021 and 256 for patterns with >= 12 entries per row, whose
// SpMV is bound by uncoalesced gathers and wants the L1, not deep stages (the two
// tile shapes, spmv_device.cuh TileCfg).  A persistent CTA walks its tiles
// round-robin; per tile (spmv_device.cuh has the pipeline)
//   stage   : the tile's slices of val / node / ptr are brought into shared
//             memory by the TMA engine (cp.async.bulk, 1-D, completion on an
//             mbarrier), double-buffered, so no register or thread is tied up
//             by the HBM stream;
//   gather  : every entry's x is gathered -- issued one tile ahead of the row
//             sums, so the HBM latency of first-touch x lines is overlapped;
//   product : the ROUNDED product replaces the value in shared memory;
//   row sums: one thread per row adds that row's products in STORED order.
// Because the product is rounded before it is added and the adds run in stored
// order, every y(i) is bit-identical to the reference's serial loop
// (z = z + val(k) * x(node(k)), no FMA) -- not merely within 1e-12.
// A row longer than a tile is reduced by the whole CTA with a fixed tree
// (the only case where the sum order differs from the reference).
//
// HBM traffic per SpMV = 12 B/entry + 4 B/row (ptr) + 8 B/row (x, compulsory)
// + 8 B/row (y) -- the algorithmic minimum; x re-reads
// are served by L1/L2.
//
// (Round-1 A/B, profiles/r1_bench_1gpu*.json: a register-staged variant with
// 128-bit ld.global.nc loads reached 234.6 us per SpMV and 278 us fused with the
// dot on the 4096^2 Poisson matrix; this TMA pipeline 229 us / 246 us.  The
// register variant was removed.)
//
// Row-sharded operators (comm.cu, HALO = true) use the same kernel as a fused
// compute + exchange step over NVLink peer memory: a few communication CTAs store
// the owned x entries other ranks need straight into those ranks' landing buffers
// and publish a sequence number while the others stream the matrix; tiles are
// ordered interior first, so the transfer overlaps the bulk of the work, and a CTA
// only waits on its peers' sequence numbers when it reaches its first boundary
// tile, whose gathers take columns beyond the owned range from the landing buffer.
#include <stdlib.h>

#include <algorithm>

#include "spmv_device.cuh"

#ifndef SIGB_SMALL_TILE_CTAS
#define SIGB_SMALL_TILE_CTAS 4
#endif

namespace sigb {

namespace {

// ---------------------------------------------------------------------------
// stand-alone SpMV kernel
// ---------------------------------------------------------------------------
template <int MODE, int NDOT, bool HALO, class CFG>
__global__ void __launch_bounds__(kThreads)
csr_tma_kernel(const __grid_constant__ CsrKernelArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar[2];
    if (a.skip_flag != nullptr && *a.skip_flag != 0) return;

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    double acc[NDOT > 0 ? NDOT : 1];
#pragma unroll
    for (int d = 0; d < (NDOT > 0 ? NDOT : 1); d++) acc[d] = 0.0;

    // halo_seq is only advanced by the last CTA to finish, so every CTA reads
    // the same value here
    unsigned long long hseq = 0;
    if (HALO && a.sync.win != nullptr)
        hseq = *reinterpret_cast<volatile unsigned long long *>(&a.sync.win->halo_seq) + 1;

    TilePipe pipe;
    const SpmvVecs v{a.x1, a.y, a.u};
    spmv_phase<MODE, NDOT, HALO, true, CFG>(a, v, smem, mbar, pipe, acc, hseq, false);
    finish_dots<NDOT>(a, acc);

    // peer-memory transport: the last CTA tells every source rank that this
    // landing buffer has been consumed
    if (HALO && a.sync.win != nullptr) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned t2 = atomicAdd(&a.sync.win->done_ticket, 1u);
            if (t2 == gridDim.x - 1) {
                a.sync.win->done_ticket = 0u;
                *reinterpret_cast<volatile unsigned long long *>(&a.sync.win->halo_seq) = hseq;
                if (a.sync.src_mask) __threadfence_system();
                for (int q = 0; q < kMaxRanks; q++)
                    if (a.sync.src_mask & (1u << q))
                        *reinterpret_cast<volatile unsigned long long *>(&a.sync.peer[q]->ack[a.sync.me]) = hseq;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// ELLPACK, slot-major on the device: node_sm[k * n_pad + i], val_sm likewise.
// One thread owns two adjacent rows; slot k of a warp's 64 rows is one
// 256-byte (node) + 512-byte (val) fully coalesced request.  All max_d slots
// are multiplied, padding included, exactly like the reference loop
// (ellpack_matrices.f90:655-658): the sum order per row is the stored order,
// so results are bit-identical to the serial loop.
// ---------------------------------------------------------------------------
struct EllKernelArgs {
    const int32_t *node;
    const double *val;
    int32_t n, n_pad, max_d;
    const double *x1;
    double *y;
    const double *u;
    double *out0, *out1;
    double *partials;
    unsigned *ticket;
    const int *skip_flag;
    const double *scale;
};

template <int MODE, int NDOT, int W>
__global__ void __launch_bounds__(kThreads)
ell_kernel(const EllKernelArgs a)
{
    if (a.skip_flag != nullptr && *a.skip_flag != 0) return;
    double acc[NDOT > 0 ? NDOT : 1];
#pragma unroll
    for (int d = 0; d < (NDOT > 0 ? NDOT : 1); d++) acc[d] = 0.0;

    const int npairs = a.n_pad >> 1;
    for (int pr = blockIdx.x * kThreads + threadIdx.x; pr < npairs;
         pr += gridDim.x * kThreads) {
        const int i = pr * 2;
        double z0 = 0.0, z1 = 0.0;
        if (W > 0) {
            int2 c[W > 0 ? W : 1];
            double2 v[W > 0 ? W : 1];
#pragma unroll
            for (int k = 0; k < W; k++) {
                c[k] = ld_stream_i2(a.node + (size_t)k * a.n_pad + i);
                v[k] = ld_stream_d2(a.val + (size_t)k * a.n_pad + i);
            }
            double x0[W > 0 ? W : 1], x1v[W > 0 ? W : 1];
#pragma unroll
            for (int k = 0; k < W; k++) {
                x0[k] = __ldg(a.x1 + c[k].x);
                x1v[k] = __ldg(a.x1 + c[k].y);
            }
#pragma unroll
            for (int k = 0; k < W; k++) {
                z0 = add(z0, mul(v[k].x, x0[k]));
                z1 = add(z1, mul(v[k].y, x1v[k]));
            }
        } else {
            for (int k = 0; k < a.max_d; k++) {
                const int2 c = ld_stream_i2(a.node + (size_t)k * a.n_pad + i);
                const double2 v = ld_stream_d2(a.val + (size_t)k * a.n_pad + i);
                z0 = add(z0, mul(v.x, __ldg(a.x1 + c.x)));
                z1 = add(z1, mul(v.y, __ldg(a.x1 + c.y)));
            }
        }
        if (i < a.n) {
            if (MODE != MODE_SET) z0 = add(a.y[i], z0);
            if (MODE == MODE_SET && a.scale) z0 = mul(a.scale[i], z0);
            a.y[i] = z0;
            if (NDOT >= 1) acc[0] = add(acc[0], mul(a.u[i], z0));
            if (NDOT >= 2) acc[NDOT > 1 ? 1 : 0] = add(acc[NDOT > 1 ? 1 : 0], mul(z0, z0));
        }
        if (i + 1 < a.n) {
            if (MODE != MODE_SET) z1 = add(a.y[i + 1], z1);
            if (MODE == MODE_SET && a.scale) z1 = mul(a.scale[i + 1], z1);
            a.y[i + 1] = z1;
            if (NDOT >= 1) acc[0] = add(acc[0], mul(a.u[i + 1], z1));
            if (NDOT >= 2) acc[NDOT > 1 ? 1 : 0] = add(acc[NDOT > 1 ? 1 : 0], mul(z1, z1));
        }
    }

    if (NDOT == 1) {
        double *const out[1] = {a.out0};
        double v[1] = {acc[0]};
        grid_reduce<1>(v, a.partials, a.ticket, out);
    } else if (NDOT == 2) {
        double *const out[2] = {a.out0, a.out1};
        double v[2] = {acc[0], acc[NDOT > 1 ? 1 : 0]};
        grid_reduce<2>(v, a.partials, a.ticket, out);
    }
}

// Small-shape launches name the shared-memory carve-out they want (the driver's default sizes it for
// the largest residency the register count allows, which leaves the L1 too small for the gathers):
// kSmallCtasPerSm CTAs of two stages each, rounded up by the driver to the next configuration.
constexpr int kSmallCtasPerSm = SIGB_SMALL_TILE_CTAS;

template <int MODE, int NDOT, bool HALO, class CFG>
int launch_csr_cfg(const CsrKernelArgs &a, cudaStream_t st)
{
    int grid = 0;
    const size_t smem = 2 * (size_t)CFG::kStageBytes;
    constexpr bool small = CFG::kNnz != kTileNnz;
    if (small) {
        static thread_local int cached = 0;   // per host thread = per device
        if (cached == 0) {
            auto kernel = csr_tma_kernel<MODE, NDOT, HALO, CFG>;
            const int ctas = env_int("SIGB_SMALL_TILE_CTAS", kSmallCtasPerSm);
            const int carve_kb = env_int("SIGB_SMALL_TILE_CARVEOUT_KB", (int)((ctas * (smem + 1024) + 1023) / 1024));
            SIGB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SIGB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           std::min(100, (carve_kb * 100 + 227) / 228)));
            int per_sm = 0;
            SIGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
            per_sm = std::max(1, std::min(per_sm, ctas));
            cached = std::min(per_sm * ctx().num_sms, (kMaxGrid / ctx().num_sms) * ctx().num_sms);
        }
        grid = cached;
    } else {
        SIGB_CHECK((occupancy_grid<csr_tma_kernel<MODE, NDOT, HALO, CFG>>(smem, &grid)));
    }
    // one CTA per tile is enough; the communication CTAs of a row-sharded operator come on top --
    // unless every CTA pushes (HaloSync::push_all): then the whole grid is named as pushers
    const bool push_all = HALO && a.sync.win != nullptr && a.sync.push_all != 0;
    const int comm = (HALO && a.sync.win != nullptr && !push_all) ? a.sync.push_ctas : 0;
    if (a.ntiles + comm < grid) grid = a.ntiles + comm;
    if (grid < comm + 1) grid = comm + 1;
    if (push_all) {
        CsrKernelArgs b = a;
        b.sync.push_ctas = grid;
        csr_tma_kernel<MODE, NDOT, HALO, CFG><<<grid, kThreads, smem, st>>>(b);
    } else {
        csr_tma_kernel<MODE, NDOT, HALO, CFG><<<grid, kThreads, smem, st>>>(a);
    }
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

template <int MODE, int NDOT, bool HALO>
int launch_csr_t(const CsrKernelArgs &a, cudaStream_t st, int32_t tile_nnz)
{
    if (tile_nnz == kTileNnzSmall) return launch_csr_cfg<MODE, NDOT, HALO, TileCfgSmall>(a, st);
    return launch_csr_cfg<MODE, NDOT, HALO, TileCfgLarge>(a, st);
}

template <int MODE, int NDOT, int W>
int launch_ell_t(const EllKernelArgs &a, cudaStream_t st)
{
    int grid = 0;
    SIGB_CHECK((occupancy_grid<ell_kernel<MODE, NDOT, W>>(0, &grid)));
    const int need = ((a.n_pad >> 1) + kThreads - 1) / kThreads;
    if (need < grid) grid = need;
    if (grid < 1) grid = 1;
    ell_kernel<MODE, NDOT, W><<<grid, kThreads, 0, st>>>(a);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

template <int MODE, int NDOT>
int launch_ell_w(const EllKernelArgs &a, cudaStream_t st)
{
    switch (a.max_d) {
    case 1: return launch_ell_t<MODE, NDOT, 1>(a, st);
    case 2: return launch_ell_t<MODE, NDOT, 2>(a, st);
    case 3: return launch_ell_t<MODE, NDOT, 3>(a, st);
    case 4: return launch_ell_t<MODE, NDOT, 4>(a, st);
    case 5: return launch_ell_t<MODE, NDOT, 5>(a, st);
    case 6: return launch_ell_t<MODE, NDOT, 6>(a, st);
    case 7: return launch_ell_t<MODE, NDOT, 7>(a, st);
    case 8: return launch_ell_t<MODE, NDOT, 8>(a, st);
    default: return launch_ell_t<MODE, NDOT, 0>(a, st);
    }
}

}  // namespace

// Which tile shape a pattern gets.  Long rows mean many gathers per row, and a row of a banded /
// stencil operator is short; the measured cross-over is not sharp (Poisson, 5 per row: large shape
// 0.97 of HBM; Erdos-Renyi, 22 per row: gather-bound, small shape), so the rule is the mean row
// length.  SIGB_TILE_CLASS = 0 / 1 forces the large / small shape (A/B runs).
TileShape tile_shape_for(int64_t nnz, int64_t nrows)
{
    static const int forced = env_int("SIGB_TILE_CLASS", -1);
    static const int min_mean = env_int("SIGB_SMALL_TILE_MIN_ROW", 12);
    bool small = nrows > 0 && nnz >= (int64_t)min_mean * nrows;
    if (forced == 0) small = false;
    if (forced == 1) small = true;
    return tile_shape_of(small ? kTileNnzSmall : kTileNnz);
}

// Greedy row tiling: consecutive rows while the tile holds <= cap entries
// and <= rows rows; a longer row gets a tile of its own.  ptr is monotone,
// so the end of a tile is found by bisection inside the row window (the
// row-by-row walk cost 5 ms per 4 M rows on the host, more than the device side
// of a matrix copy).
int build_tiles_host(const int32_t *ptr1, int32_t nrows, std::vector<TileDesc> &tiles, TileShape *shape_out)
{
    const TileShape shape = tile_shape_for(nrows > 0 ? (int64_t)ptr1[nrows] - 1 : 0, nrows);
    if (shape_out) *shape_out = shape;
    tiles.clear();
    tiles.reserve((size_t)nrows / 256 + 16);
    int32_t s = 0;
    while (s < nrows) {
        const int64_t limit = (int64_t)ptr1[s] + shape.cap;          // ptr1[e] <= limit keeps the tile within the cap
        const int32_t hi = (int32_t)std::min<int64_t>((int64_t)s + shape.rows, nrows);
        // largest e in [s + 1, hi] with ptr1[e] <= limit; s + 1 when even the first row is too long
        const int32_t *first = ptr1 + s + 1, *last = ptr1 + hi + 1;
        int32_t e = (int32_t)(std::upper_bound(first, last, limit,
                                               [](int64_t lim, int32_t p) { return lim < (int64_t)p; }) - ptr1) - 1;
        if (e < s + 1) e = s + 1;
        TileDesc d;
        d.rs = s;
        d.re = e;
        d.ks = ptr1[s] - 1;
        d.ke = ptr1[e] - 1;
        tiles.push_back(d);
        s = e;
    }
    return SIGB_OK;
}

// Balanced tiling for operators whose tiles are shared round-robin by a KNOWN number of CTAs (`groups`:
// the compute CTAs of the persistent CG kernel on a row-sharded operator): exactly m * groups tiles of
// nearly equal entry counts, so that every CTA streams the same number of tiles -- with the greedy
// tiling a shard of 5130 tiles on 440 CTAs gives some of them 12 tiles and the others 11, and the
// whole grid waits at the barrier behind the phase for the twelfth (8 % of the SpMV phase).  Tile
// boundaries sit at the rows where the running entry count passes j * nnz / T.  Falls back to the
// greedy tiling when the caps cannot be met that way (long or empty rows, tiny matrices), and for
// the small tile shape (gather-bound operators run kernel per phase, on a grid of their own).  Tiling
// never changes a result: rows are summed by one thread each, in stored order, whatever the tile.
int build_tiles_balanced(const int32_t *ptr1, int32_t nrows, int groups, std::vector<TileDesc> &tiles,
                         TileShape *shape_out)
{
    const int64_t nnz = nrows > 0 ? (int64_t)ptr1[nrows] - 1 : 0;
    const TileShape shape = tile_shape_for(nnz, nrows);
    if (groups <= 0 || nrows <= 0 || nnz <= 0 || shape.nnz != kTileNnz) return build_tiles_host(ptr1, nrows, tiles, shape_out);
    if (shape_out) *shape_out = shape;
    int64_t maxrow = 0;
    for (int32_t i = 0; i < nrows; i++) maxrow = std::max<int64_t>(maxrow, ptr1[i + 1] - ptr1[i]);
    const int64_t m0 = std::max<int64_t>(1, (nnz + (int64_t)kTileCap * groups - 1) / ((int64_t)kTileCap * groups));
    for (int64_t m = m0; m <= m0 + 3; m++) {
        const int64_t T = m * groups;
        if (T > nrows) break;
        if ((nnz + T - 1) / T + maxrow > kTileCap) continue;
        tiles.clear();
        tiles.reserve((size_t)T);
        bool ok = true;
        int32_t s = 0;
        for (int64_t j = 1; j <= T && ok; j++) {
            int32_t e;
            if (j == T) {
                e = nrows;
            } else {
                // first row whose preceding entries reach j * nnz / T (as sigb_partition_rows does)
                const int64_t target = (nnz * j) / T + 1;     // compare with 1-based ptr
                e = (int32_t)(std::lower_bound(ptr1 + s, ptr1 + nrows, target,
                                               [](int32_t p_, int64_t t_) { return (int64_t)p_ < t_; }) - ptr1);
            }
            TileDesc d;
            d.rs = s;
            d.re = e;
            d.ks = ptr1[s] - 1;
            d.ke = ptr1[e] - 1;
            if (e <= s || e - s > kTileRows || d.ke - d.ks > kTileCap) ok = false;
            tiles.push_back(d);
            s = e;
        }
        if (ok && s == nrows) return SIGB_OK;
    }
    return build_tiles_host(ptr1, nrows, tiles, shape_out);
}

// Argument block of one SpMV over all tiles of A (which = 0), or over its
// interior / boundary sub-tables (which = 1 / 2, NCCL transport).
#ifdef SIGB_PHASE_TIMERS
// diagnostic build: one buffer per process for the per-tile cycle counters (spmv_device.cuh)
static unsigned long long *tile_dbg_buffer()
{
    static thread_local unsigned long long *buf = nullptr;
    if (!buf) {
        if (cudaMalloc((void **)&buf, sizeof(unsigned long long) * 6) != cudaSuccess) return nullptr;
        cudaMemset(buf, 0, sizeof(unsigned long long) * 6);
    }
    return buf;
}
#endif

static void fill_args(const CsrView &A, const double *val, const double *x, double *y, const DotSpec &dot,
                      int which, int ticket, CsrKernelArgs &a)
{
#ifdef SIGB_PHASE_TIMERS
    a.tile_dbg = tile_dbg_buffer();
#endif
    a.ptr = A.ptr;
    a.node = A.node;
    a.val = val;
    a.tiles = A.tiles;
    a.ntiles = A.ntiles;
    if (which == 1) { a.tiles = A.tiles_interior; a.ntiles = A.n_interior; }
    if (which == 2) { a.tiles = A.tiles_boundary; a.ntiles = A.n_boundary; }
    if (which == 3) { a.tiles = A.tiles_nonempty; a.ntiles = A.n_nonempty; }
    a.x1 = x - 1;
    a.y = y;
    a.u = dot.u;
    a.out0 = dot.out[0];
    a.out1 = dot.out[1];
    a.partials = ctx().partials + (size_t)ticket * kMaxGrid * kMaxDots;
    a.ticket = ctx().tickets + ticket;
    a.skip_flag = dot.skip_flag;
    a.scale = dot.row_scale;
    a.add0 = dot.addend[0];
    a.add1 = dot.addend[1];
    a.nloc = dot.nloc;
    a.h1 = dot.halo ? dot.halo - (dot.nloc + 1) : nullptr;
    if (dot.sync) a.sync = *dot.sync;
    if (dot.red) a.red = *dot.red;
    a.first_halo_tile = (which == 0 && dot.sync) ? A.n_interior : 0;
    a.fault = ctx().fault_dev;
}

int fill_csr_args(const CsrView &V, const double *val, const double *x, double *y, const DotSpec &dot,
                  CsrKernelArgs *out)
{
    *out = CsrKernelArgs();
    fill_args(V, val, x, y, dot, 0, 0, *out);
    return SIGB_OK;
}

int launch_csr_spmv(const CsrView &A, const double *val, const double *x, double *y,
                    SpmvMode mode, const DotSpec &dot, int which, cudaStream_t stream,
                    int ticket)
{
    const bool halo = dot.halo != nullptr || dot.sync != nullptr;
    // Accumulating forms without fused dots only have to visit tiles that hold
    // entries: csc_matvec_add (MODE_ACC_INIT) never touches a y(i) without
    // contributions, and csr_matvec_add's y(i) = y(i) + 0.0 is the identity unless
    // y(i) is -0.0, which the caller rules out with y_no_negative_zero.
    if (which == 0 && !halo && dot.ndot == 0 && A.tiles_nonempty != nullptr &&
        (mode == MODE_ACC_INIT || (mode == MODE_ADD_AFTER && dot.y_no_negative_zero)))
        which = 3;
    CsrKernelArgs a = CsrKernelArgs();
    fill_args(A, val, x, y, dot, which, ticket, a);
    cudaStream_t st = stream ? stream : ctx().stream;
    if (a.ntiles == 0 && dot.ndot == 0) return SIGB_OK;

#define SIGB_DISPATCH(M)                                                            \
    switch (dot.ndot) {                                                             \
    case 0: return halo ? launch_csr_t<M, 0, true>(a, st, A.tile_nnz) : launch_csr_t<M, 0, false>(a, st, A.tile_nnz);  \
    case 1: return halo ? launch_csr_t<M, 1, true>(a, st, A.tile_nnz) : launch_csr_t<M, 1, false>(a, st, A.tile_nnz);  \
    default: return halo ? launch_csr_t<M, 2, true>(a, st, A.tile_nnz) : launch_csr_t<M, 2, false>(a, st, A.tile_nnz); \
    }
    switch (mode) {
    case MODE_SET: SIGB_DISPATCH(MODE_SET)
    case MODE_ADD_AFTER: SIGB_DISPATCH(MODE_ADD_AFTER)
    default: SIGB_DISPATCH(MODE_ACC_INIT)
    }
#undef SIGB_DISPATCH
}

int launch_ell_spmv(int32_t n, int32_t n_pad, int32_t max_d, const int32_t *node_sm,
                    const double *val_sm, const double *x, double *y, SpmvMode mode,
                    const DotSpec &dot)
{
    EllKernelArgs a;
    a.node = node_sm;
    a.val = val_sm;
    a.n = n;
    a.n_pad = n_pad;
    a.max_d = max_d;
    a.x1 = x - 1;
    a.y = y;
    a.u = dot.u;
    a.out0 = dot.out[0];
    a.out1 = dot.out[1];
    a.partials = ctx().partials;
    a.ticket = ctx().tickets;
    a.skip_flag = dot.skip_flag;
    a.scale = dot.row_scale;
    cudaStream_t st = ctx().stream;
    const bool set = (mode == MODE_SET);
    switch (dot.ndot) {
    case 0: return set ? launch_ell_w<MODE_SET, 0>(a, st) : launch_ell_w<MODE_ADD_AFTER, 0>(a, st);
    case 1: return set ? launch_ell_w<MODE_SET, 1>(a, st) : launch_ell_w<MODE_ADD_AFTER, 1>(a, st);
    default: return set ? launch_ell_w<MODE_SET, 2>(a, st) : launch_ell_w<MODE_ADD_AFTER, 2>(a, st);
    }
}

}  // namespace sigb

extern "C" {

// Diagnostic: SM cycles of thread 0 of every CTA of the streaming CSR kernel, summed over
// the grid and over the launches since the last call (then reset): out[0] waiting for the
// staged tile, [1] products (gathers), [2] row sums, [3] whole pass, [4] staged tiles,
// [5] CTA passes.  *supported = 0 (and zeros) unless built with -DSIGB_PHASE_TIMERS.
int sigb_debug_spmv_tile_cycles(unsigned long long *out, int *supported)
{
    using namespace sigb;
    SIGB_REQUIRE(out && supported, SIGB_ERR_ARG, "sigb_debug_spmv_tile_cycles: bad argument");
    for (int k = 0; k < 6; k++) out[k] = 0ull;
    *supported = 0;
#ifdef SIGB_PHASE_TIMERS
    unsigned long long *buf = tile_dbg_buffer();
    SIGB_REQUIRE(buf, SIGB_ERR_CUDA, "sigb_debug_spmv_tile_cycles: no buffer");
    SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
    SIGB_CUDA(cudaMemcpy(out, buf, sizeof(unsigned long long) * 6, cudaMemcpyDeviceToHost));
    SIGB_CUDA(cudaMemset(buf, 0, sizeof(unsigned long long) * 6));
    *supported = 1;
#endif
    return SIGB_OK;
}

}  // extern "C"
