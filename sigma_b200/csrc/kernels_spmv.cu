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
// CSR kernel ("stream" layout).  Rows are grouped on the host into tiles of at
// most kTileNnz stored entries.  A CTA walks tiles round-robin; per tile
//   phase 1: the tile's slice of node/val is read with 128-bit streaming loads
//            (perfectly coalesced, independent of row lengths), each entry is
//            multiplied with its gathered x and the ROUNDED product is parked
//            in shared memory;
//   phase 2: one thread per row adds that row's products in STORED order.
// Because the product is rounded before it is added and the adds run in stored
// order, every y(i) is bit-identical to the reference's serial loop
// (z = z + val(k) * x(node(k)), no FMA) -- not merely within 1e-12.
// A row longer than a tile is reduced by the whole CTA with a fixed tree
// (the only case where the sum order differs from the reference).
//
// HBM traffic per SpMV = 12 B/entry + 4 B/row (ptr) + 8 B/row (x, compulsory)
// + 8 B/row (y) -- the algorithmic minimum; x re-reads
// are served by L1/L2.
#include "device_utils.cuh"

namespace sigb {

namespace {

struct CsrKernelArgs {
    const int32_t *ptr;       // 1-based, nrows + 1
    const int32_t *node;      // 1-based
    const double *val;
    const int32_t *tile_row;  // ntiles + 1
    const int32_t *tile_list; // optional indirection (subset launch) or null
    int32_t ntiles;           // tiles this launch covers
    const double *x1;         // x - 1 : indexable by 1-based column id
    double *y;
    const double *u;          // dot vector
    double *out0, *out1;
    double *partials;
    unsigned *ticket;
    const int *skip_flag;
    const double *scale;      // optional per-row scaling of the result
};

template <int MODE, int NDOT>
__global__ void __launch_bounds__(kThreads)
csr_stream_kernel(const CsrKernelArgs a)
{
    extern __shared__ __align__(16) double prod[];  // kTileNnz rounded products
    if (a.skip_flag != nullptr && *a.skip_flag != 0) return;

    const int tid = threadIdx.x;
    double acc[NDOT > 0 ? NDOT : 1];
#pragma unroll
    for (int d = 0; d < (NDOT > 0 ? NDOT : 1); d++) acc[d] = 0.0;

    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        const int tile = a.tile_list ? a.tile_list[t] : t;
        const int rs = a.tile_row[tile], re = a.tile_row[tile + 1];
        const int ks = a.ptr[rs] - 1, ke = a.ptr[re] - 1;  // 0-based entry range
        const int len = ke - ks;

        if (len <= kTileCap) {
            // ---- phase 1: stream the tile, stage rounded products ----------
            const int ka = ks & ~3;  // 16-byte aligned start of the node slice
#pragma unroll
            for (int it = 0; it < kTileNnz / (4 * kThreads); it++) {
                const int j = ka + 4 * (tid + it * kThreads);
                if (j < ke) {
                    const int4 c = ld_stream_i4(a.node + j);
                    const double2 v01 = ld_stream_d2(a.val + j);
                    const double2 v23 = ld_stream_d2(a.val + j + 2);
                    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
                    if (j >= ks) p0 = mul(v01.x, __ldg(a.x1 + c.x));
                    if (j + 1 >= ks && j + 1 < ke) p1 = mul(v01.y, __ldg(a.x1 + c.y));
                    if (j + 2 >= ks && j + 2 < ke) p2 = mul(v23.x, __ldg(a.x1 + c.z));
                    if (j + 3 >= ks && j + 3 < ke) p3 = mul(v23.y, __ldg(a.x1 + c.w));
                    double2 *dst = reinterpret_cast<double2 *>(prod + (j - ka));
                    dst[0] = make_double2(p0, p1);
                    dst[1] = make_double2(p2, p3);
                }
            }
            __syncthreads();
            // ---- phase 2: per-row sums in stored order ----------------------
            for (int r = rs + tid; r < re; r += kThreads) {
                const int b = a.ptr[r] - 1 - ka, e = a.ptr[r + 1] - 1 - ka;
                double z = (MODE == MODE_ACC_INIT) ? a.y[r] : 0.0;
                for (int k = b; k < e; k++) z = add(z, prod[k]);
                if (MODE == MODE_ADD_AFTER) z = add(a.y[r], z);
                if (MODE == MODE_SET && a.scale) z = mul(a.scale[r], z);
                a.y[r] = z;
                if (NDOT >= 1) acc[0] = add(acc[0], mul(a.u[r], z));
                if (NDOT >= 2) acc[NDOT > 1 ? 1 : 0] = add(acc[NDOT > 1 ? 1 : 0], mul(z, z));
            }
            __syncthreads();
        } else {
            // ---- one long row: CTA-wide fixed-tree reduction -----------------
            __shared__ double smr[1][kThreads / 32];
            double s[1] = {0.0};
            for (int k = ks + tid; k < ke; k += kThreads)
                s[0] = add(s[0], mul(a.val[k], __ldg(a.x1 + a.node[k])));
            block_tree<1>(s, smr);
            if (tid == 0) {
                double z = s[0];
                if (MODE == MODE_ACC_INIT || MODE == MODE_ADD_AFTER) z = add(a.y[rs], z);
                if (MODE == MODE_SET && a.scale) z = mul(a.scale[rs], z);
                a.y[rs] = z;
                if (NDOT >= 1) acc[0] = add(acc[0], mul(a.u[rs], z));
                if (NDOT >= 2) acc[NDOT > 1 ? 1 : 0] = add(acc[NDOT > 1 ? 1 : 0], mul(z, z));
            }
            __syncthreads();
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
#pragma unroll
            for (int k = 0; k < W; k++) {
                z0 = add(z0, mul(v[k].x, __ldg(a.x1 + c[k].x)));
                z1 = add(z1, mul(v[k].y, __ldg(a.x1 + c[k].y)));
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

template <typename K>
int occupancy_grid(K kernel, size_t smem, int *grid_out)
{
    static int cached = 0;  // one per template instantiation
    if (cached == 0) {
        int per_sm = 0;
        SIGB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SIGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
        if (per_sm < 1) per_sm = 1;
        int g = per_sm * ctx().num_sms;
        if (g > kMaxGrid) g = (kMaxGrid / ctx().num_sms) * ctx().num_sms;
        cached = g;
    }
    *grid_out = cached;
    return SIGB_OK;
}

template <int MODE, int NDOT>
int launch_csr_t(const CsrKernelArgs &a, cudaStream_t st)
{
    const size_t smem = (size_t)kTileNnz * sizeof(double);
    int grid = 0;
    SIGB_CHECK(occupancy_grid(csr_stream_kernel<MODE, NDOT>, smem, &grid));
    if (a.ntiles < grid) grid = a.ntiles;
    if (grid < 1) grid = 1;
    csr_stream_kernel<MODE, NDOT><<<grid, kThreads, smem, st>>>(a);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

template <int MODE, int NDOT, int W>
int launch_ell_t(const EllKernelArgs &a, cudaStream_t st)
{
    int grid = 0;
    SIGB_CHECK(occupancy_grid(ell_kernel<MODE, NDOT, W>, 0, &grid));
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

// Greedy row tiling: consecutive rows while the tile holds <= kTileCap entries
// and <= kTileRowsMax rows; a longer row gets a tile of its own.
int build_tiles_host(const int32_t *ptr1, int32_t nrows, std::vector<int32_t> &tile_row)
{
    tile_row.clear();
    tile_row.push_back(0);
    int32_t s = 0;
    while (s < nrows) {
        int32_t e = s + 1;
        const int64_t base = ptr1[s];
        while (e < nrows && (int64_t)ptr1[e + 1] - base <= kTileCap && e - s < kTileRowsMax) e++;
        tile_row.push_back(e);
        s = e;
    }
    return SIGB_OK;
}

int launch_csr_spmv(const CsrView &A, const double *val, const double *x, double *y,
                    SpmvMode mode, const DotSpec &dot, int which, cudaStream_t stream,
                    int ticket)
{
    CsrKernelArgs a;
    a.ptr = A.ptr;
    a.node = A.node;
    a.val = val;
    a.tile_row = A.tile_row;
    a.tile_list = nullptr;
    a.ntiles = A.ntiles;
    if (which == 1) { a.tile_list = A.tiles_interior; a.ntiles = A.n_interior; }
    if (which == 2) { a.tile_list = A.tiles_boundary; a.ntiles = A.n_boundary; }
    a.x1 = x - 1;
    a.y = y;
    a.u = dot.u;
    a.out0 = dot.out[0];
    a.out1 = dot.out[1];
    a.partials = ctx().partials + (size_t)ticket * kMaxGrid * kMaxDots;
    a.ticket = ctx().tickets + ticket;
    a.skip_flag = dot.skip_flag;
    a.scale = dot.row_scale;
    cudaStream_t st = stream ? stream : ctx().stream;
    if (a.ntiles == 0 && dot.ndot == 0) return SIGB_OK;

#define SIGB_DISPATCH(M)                                                   \
    switch (dot.ndot) {                                                    \
    case 0: return launch_csr_t<M, 0>(a, st);                              \
    case 1: return launch_csr_t<M, 1>(a, st);                              \
    default: return launch_csr_t<M, 2>(a, st);                             \
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
