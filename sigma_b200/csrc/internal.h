// internal.h -- private declarations shared by the sigma_b200 CUDA sources.
// Nothing here is part of the C-ABI (include/sigma_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/sigma_b200.h"

namespace sigb {

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define SIGB_CUDA(call)                                                       \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess)                                               \
            return ::sigb::cuda_fail(e__, #call, __FILE__, __LINE__);         \
    } while (0)

#define SIGB_CHECK(call)                                                      \
    do {                                                                      \
        int s__ = (call);                                                     \
        if (s__ != SIGB_OK) return s__;                                       \
    } while (0)

#define SIGB_REQUIRE(cond, code, ...)                                         \
    do {                                                                      \
        if (!(cond)) {                                                        \
            ::sigb::set_error(__VA_ARGS__);                                   \
            return (code);                                                    \
        }                                                                     \
    } while (0)

// ---------------------------------------------------------------------------
// runtime context: one per host THREAD (thread_local).  One process per GPU uses the main thread's;
// the single-process multi-GPU mode (mgpu.cu) gives every worker thread its own, bound to its GPU
// ---------------------------------------------------------------------------
struct Ctx {
    bool inited = false;
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;    // stream every launch goes to
    cudaStream_t aux_stream = nullptr;  // halo packing / exchange overlap
    int64_t launches = 0;
    // scratch for grid reductions
    double *partials = nullptr;       // kMaxGrid * kMaxDots doubles
    unsigned *tickets = nullptr;      // kNumTickets counters, zero between kernels
    void *pinned = nullptr;           // small pinned host staging area
    size_t pinned_bytes = 0;
    // grow-only device staging of the host-pointer matvec entry points (x | y)
    double *stage[2] = {nullptr, nullptr};
    size_t stage_len[2] = {0, 0};
    // device-side wait timeouts (device_utils.cuh spin_wait): mapped pinned host memory
    struct FaultBlock *fault = nullptr;       // host address
    struct FaultBlock *fault_dev = nullptr;   // the same block as the device sees it
};
Ctx &ctx();
int require_init();
// SIGB_OK, or SIGB_ERR_COMM once a device-side wait has timed out (sticky: the row-sharded
// state of this process can no longer be trusted).  Called wherever the host synchronises.
int check_fault(const char *where);

constexpr int kThreads = 256;          // every kernel in this library uses 256-thread CTAs
constexpr int kMaxGrid = 148 * 16;     // upper bound on persistent grid sizes
constexpr int kMaxDots = 4;
constexpr int kNumTickets = 8;

inline void count_launch(int n = 1) { ctx().launches += n; }

// Integer environment switch (opt-in paths and experiment knobs); callers keep the result in a
// function-local static, so each switch is read once per process.
inline int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

// ---------------------------------------------------------------------------
// device-side sparse structures
// ---------------------------------------------------------------------------

// How the row sum z meets y; reproduces the reference's accumulation order.
enum SpmvMode {
    MODE_SET = 0,        // y(i) = z                 (matvec: y was zero-filled)
    MODE_ADD_AFTER = 1,  // y(i) = y(i) + z, z from 0 (csr_matvec_add order)
    MODE_ACC_INIT = 2    // z starts from y(i)       (csc_matvec_add order)
};

// One row tile of the streaming CSR kernel: rows [rs, re) and their stored
// entries [ks, ke), all 0-based.  Tiles are built once on the host so the
// kernel finds everything it needs about a tile in ONE 16-byte load instead of
// chasing tile_row -> ptr -> data.
struct TileDesc {
    int32_t rs, re, ks, ke;
};

// CSR arrays exactly as the reference stores them (1-based) plus the row
// tiling used by the streaming kernel.
struct CsrView {
    int32_t nrows = 0, ncols = 0;
    int64_t nnz = 0;
    int32_t *ptr = nullptr;       // nrows + 1 (+ pad), 1-based
    int32_t *node = nullptr;      // nnz (+ pad), 1-based column ids, stored order
    TileDesc *tiles = nullptr;    // ntiles descriptors
    int32_t ntiles = 0;
    // optional tile subsets (row-sharded operators): interior tiles touch no
    // halo column, boundary tiles do.
    TileDesc *tiles_interior = nullptr, *tiles_boundary = nullptr;
    int32_t n_interior = 0, n_boundary = 0;
    // tiles that hold at least one stored entry (a view behind `tiles`): all an
    // accumulating matvec has to visit when rows without entries leave y as it is
    TileDesc *tiles_nonempty = nullptr;
    int32_t n_nonempty = 0;
    // entries per tile the table was built for (0 = kTileNnz): selects the kernel's tile shape
    int32_t tile_nnz = 0;
};

#ifndef SIGB_TILE_NNZ
#define SIGB_TILE_NNZ 2048
#endif
#ifndef SIGB_TILE_ROWS
#define SIGB_TILE_ROWS 512
#endif
constexpr int kTileNnz = SIGB_TILE_NNZ;   // entries staged per tile
// A tile's entry range is widened down to a 4-entry boundary so the slices
// can be moved with aligned 16-byte transfers; capping tiles at kTileNnz - 3
// entries keeps the widened range within kTileNnz.
constexpr int kTileCap = kTileNnz - 3;
constexpr int kTileRows = SIGB_TILE_ROWS;   // rows per tile (their ptr slice is staged too)
// Second tile shape, for operators whose SpMV is bound by uncoalesced x gathers (long rows with scattered
// columns): small stages leave most of the SM's unified shared-memory / L1 array to the L1, and the
// gather rate follows the L1 share (spmv_device.cuh, TileCfg).
#ifndef SIGB_TILE_NNZ_SMALL
#define SIGB_TILE_NNZ_SMALL 1024
#endif
constexpr int kTileNnzSmall = SIGB_TILE_NNZ_SMALL;
// caps of a row tiling: at most `cap` stored entries and `rows` rows per tile
struct TileShape {
    int32_t nnz;    // staged entries per tile the kernel is compiled for (kTileNnz or kTileNnzSmall)
    int32_t cap;    // nnz - 3 (the entry range is widened down to a 4-entry boundary)
    int32_t rows;
};
inline TileShape tile_shape_of(int32_t tile_nnz)
{
    const int32_t tn = tile_nnz > 0 ? tile_nnz : kTileNnz;
    return TileShape{tn, tn - 3, tn == kTileNnz ? kTileRows : (tn / 4 < 128 ? 128 : tn / 4)};
}
// the shape for a pattern with nnz entries in nrows rows (kernels_spmv.cu; SIGB_TILE_CLASS overrides)
TileShape tile_shape_for(int64_t nnz, int64_t nrows);
constexpr int kPad = 8;          // slack entries behind ptr / node / val arrays

enum GraphKind { G_CSR = 0, G_CSC = 1, G_ELL = 2 };

struct DistInfo;    // comm.cu
struct MgpuMatrix;  // mgpu.cu

// An operator expression over other operators (operators.cu): operator_sum,
// operator_product, operator_adjoint (src/linear_operator/linear_operator_
// {sums,products,adjoints}.f90) or the block composite sparse_matrix
// (src/matrix/sparse_matrix_composites.f90:41-49).
enum OpKind { OP_SUM = 1, OP_PRODUCT = 2, OP_ADJOINT = 3, OP_COMPOSITE = 4 };
struct OpInfo {
    int kind = 0;
    std::vector<sigb_matrix_s *> kids;       // summands / products / op / sub_mats(it, jt) row-major
    int32_t num_row_mats = 0, num_col_mats = 0;
    std::vector<int32_t> row_ptr, col_ptr;   // composite block offsets, 1-based like the reference
    int64_t temp_vec_size = 0;               // operator_product%temp_vec_size
    double *z1 = nullptr, *z2 = nullptr;     // operator_product%z1, z2 (device)
};

}  // namespace sigb

struct sigb_graph_s {
    int kind = 0;
    int refcount = 1;
    int32_t n = 0, m = 0;       // as in the reference graph (n lines of m possible ids)
    int64_t ne = 0;
    int32_t max_d = 0;
    // compressed sparse: arrays as stored, and the stable transpose
    sigb::CsrView stored;
    sigb::CsrView transposed;
    bool has_transposed = false;
    int32_t *perm_t = nullptr;  // transposed position -> stored entry index (0-based)
    // ellpack, slot-major on the device: node_sm[k * n_pad + i]
    int32_t *ell_node = nullptr;
    int32_t *ell_degrees = nullptr;  // degrees(n), as stored
    int32_t n_pad = 0;
    std::vector<int32_t> host_ptr_t;  // transposed ptr kept on the host for tiling
};

struct sigb_matrix_s {
    int refcount = 1;           // linear_operator%reference_count (linear_operator_interface.f90:285-302)
    sigb_graph_t g = nullptr;   // null for operator expressions
    double *val = nullptr;      // cs: ne (+pad) in stored order; ell: slot-major [max_d][n_pad]
    double *val_t = nullptr;    // values in transposed order (cs: perm gather; ell: from slots)
    bool val_t_valid = false;
    int32_t nrow = 0, ncol = 0;
    sigb::DistInfo *dist = nullptr;  // non-null for row-sharded operators
    sigb::MgpuMatrix *mg = nullptr;  // non-null for the single-process multi-GPU operator (one row block per GPU)
    sigb::OpInfo *op = nullptr;      // non-null for operator expressions (no graph, no values of their own)
};

namespace sigb {

// ---------------------------------------------------------------------------
// kernels_spmv.cu
// ---------------------------------------------------------------------------
struct DotSpec {
    int ndot = 0;                 // 0, 1 (u . y) or 2 (u . y and y . y)
    const double *u = nullptr;    // vector dotted with the result rows
    double *out[2] = {nullptr, nullptr};  // device scalars receiving the sums
    const int *skip_flag = nullptr;       // if non-null and *skip_flag != 0 the kernel is a no-op
    const double *row_scale = nullptr;    // y(i) = row_scale(i) * z (fused jacobi_solve), MODE_SET only
    const double *addend[2] = {nullptr, nullptr};  // partial sums of an earlier launch, added to the totals
    const double *halo = nullptr;         // row-sharded: values of columns nloc+1.. (halo landing buffer)
    int32_t nloc = 0;                     // row-sharded: number of owned columns
    const struct HaloSync *sync = nullptr;  // peer-memory transport: flags to wait on / acknowledge
    // y += A x only: the caller guarantees no y(i) is -0.0 (true of every y an
    // expression has already written), so rows without entries -- y(i) + 0.0 in
    // csr_matvec_add -- may be left untouched and tiles without entries skipped
    bool y_no_negative_zero = false;
    // row-sharded operators on the peer-memory transport: complete the cross-GPU part of the dot
    // products inside this kernel (its last CTA) instead of a separate all-reduce launch
    const struct RedFuse *red = nullptr;
};

// Peer-memory halo exchange, fused into the SpMV kernel (comm.cu builds it).
// Producer side: the first push_ctas CTAs of the grid (communication CTAs, no
// tiles of their own) store the owned entries other ranks need straight into
// those ranks' landing buffers and publish a sequence number.  Consumer side: a
// CTA waits for the peers' sequence numbers when it reaches its first boundary
// tile; the last CTA acknowledges consumption.
constexpr int kMaxRanks = 8;
struct HaloWin;  // device-resident, IPC-shared (device_utils.cuh)
struct RedWin;   // all-reduce inbox, IPC-shared (device_utils.cuh)
// All-reduce endpoints handed to a kernel that finishes its own reduction across the GPUs
// (device_utils.cuh grid_reduce): nranks <= 1 means "local sums only".
struct RedFuse {
    RedWin *win = nullptr;                // this rank's inbox
    RedWin *peer[kMaxRanks] = {};         // every rank's inbox, peer-mapped (peer[me] == win)
    int me = 0, nranks = 1;
    struct FaultBlock *fault = nullptr;   // wait timeouts (device_utils.cuh)
};
struct HaloSync {
    HaloWin *win = nullptr;               // this rank's window
    HaloWin *peer[kMaxRanks] = {};        // the peers' windows (peer-mapped)
    uint32_t src_mask = 0;                // ranks we receive halo entries from
    uint32_t dst_mask = 0;                // ranks we send entries to
    int me = 0;
    const double *halo_base = nullptr;    // two landing buffers, halo_stride apart
    int64_t halo_stride = 0;
    // push plan
    const int32_t *send_rows = nullptr;   // 1-based owned rows, grouped by destination
    int32_t total_send = 0;
    int32_t push_ctas = 0;                // communication CTAs (fixed when the operator is created)
    // Operators without locality (a random graph: nearly every tile reads halo columns, nearly all of x is
    // sent): there is no interior work to hide the transfer behind, so in a stand-alone launch EVERY CTA
    // first pushes its share of the send list and then takes tiles (the launcher sets push_ctas = grid).
    // The persistent CG kernel keeps dedicated communication CTAs and clears this.
    int32_t push_all = 0;
    int32_t send_off[kMaxRanks + 1] = {};
    double *dst[kMaxRanks] = {};          // peer landing buffer 0, offset to our slice
    int64_t dst_stride[kMaxRanks] = {};
};

// which: 0 = all tiles, 1 = interior subset, 2 = boundary subset, 3 = tiles with entries
int launch_csr_spmv(const CsrView &A, const double *val, const double *x,
                    double *y, SpmvMode mode, const DotSpec &dot, int which = 0,
                    cudaStream_t stream = nullptr, int ticket = 0);
int launch_ell_spmv(int32_t n, int32_t n_pad, int32_t max_d,
                    const int32_t *node_sm, const double *val_sm,
                    const double *x, double *y, SpmvMode mode,
                    const DotSpec &dot);
int build_tiles_host(const int32_t *ptr1, int32_t nrows, std::vector<TileDesc> &tiles, TileShape *shape_out = nullptr);
// exactly m * groups tiles of nearly equal entry counts when the caps allow it, else the greedy tiling
int build_tiles_balanced(const int32_t *ptr1, int32_t nrows, int groups, std::vector<TileDesc> &tiles,
                         TileShape *shape_out = nullptr);
// compute CTAs of the persistent CG kernel on this device (cg_persistent.cu), before communication CTAs are taken off
int persistent_grid_ctas();
// tiles_device.cu: the same tiling built on the device from a device-resident ptr (no read-back);
// also returns the extreme line lengths
int build_tiles_device(const int32_t *ptr1_dev, int32_t nrows, int64_t nnz, CsrView &v, int32_t *max_d, int32_t *min_d);

// ---------------------------------------------------------------------------
// transpose.cu
// ---------------------------------------------------------------------------
// Stable transpose of `ne` entries living in `nlines` lines (cs: ptr-delimited,
// ell: fixed width) into a CSR over `ntargets` rows.  Output arrays are
// allocated here.  perm[pos] = source entry index (0-based, ascending within a
// row => reference accumulation order).
int device_transpose_cs(const int32_t *ptr1, const int32_t *node1, int32_t nlines,
                        int32_t ntargets, int64_t ne, int32_t **ptr_t,
                        int32_t **node_t, int32_t **perm);
int device_transpose_ell(const int32_t *node_sm, int32_t n, int32_t n_pad,
                         int32_t max_d, int32_t ntargets, int32_t **ptr_t,
                         int32_t **node_t, int32_t **perm);
int gather_values(const double *val, const int32_t *perm, int64_t ne, double *val_t);
int ell_relayout_node(const int32_t *node_cm_dev, int32_t n, int32_t n_pad,
                      int32_t max_d, int32_t *node_sm);
int ell_relayout_val(const double *val_cm_dev, int32_t n, int32_t n_pad,
                     int32_t max_d, double *val_sm);
// transposed ell values: perm indexes the (line-major) stored layout
int gather_values_ell(const double *val_sm, const int32_t *perm, int64_t ne,
                      int32_t n_pad, int32_t max_d, double *val_t);
// exclusive scan of n int32 counts into n + 1 one-based offsets (ptr1[0] = 1)
int scan_to_ptr1(const int32_t *cnt, int64_t n, int32_t *ptr1);
int fill_i32(int32_t *p, int64_t n, int32_t v);
int fill_f64(double *p, int64_t n, double v);

// ---------------------------------------------------------------------------
// short-lived device scratch (api.cu): stream-ordered cudaMallocAsync / cudaFreeAsync on the
// library's stream from the device's default pool, kept cached (no release threshold), so a
// copy / transpose / assembly call does not pay a device-wide synchronisation per temporary.
// ---------------------------------------------------------------------------
cudaError_t tmp_alloc_bytes(void **p, size_t bytes);
cudaError_t tmp_free(void *p);
template <typename T>
inline cudaError_t tmp_alloc(T **p, size_t count)
{
    return tmp_alloc_bytes((void **)p, sizeof(T) * (count > 0 ? count : 1));
}

// ---------------------------------------------------------------------------
// matrix-level dispatch (api.cu)
// ---------------------------------------------------------------------------
int ensure_transposed(sigb_matrix_t A);
// upload the tile table of a CSR view (api.cu)
int upload_tiles(CsrView &v, const std::vector<TileDesc> &tiles);
// y = op(A) x with device vectors; the single entry every solver goes through
int matvec_dev(sigb_matrix_t A, int trans, const double *x, double *y,
               SpmvMode mode_csr_like, bool add, const DotSpec &dot);

// ---------------------------------------------------------------------------
// mgpu.cu: single-process multi-GPU mode (one worker thread per GPU behind the C-ABI)
// ---------------------------------------------------------------------------
bool mgpu_active();
void mgpu_matrix_free(sigb_matrix_t A);
int mgpu_set_values(sigb_matrix_t A, const double *val, int64_t count);
int mgpu_matvec(sigb_matrix_t A, int trans, const double *x, double *y, bool add);
int64_t mgpu_nnz(sigb_matrix_t A);
int64_t mgpu_launch_count();

// ---------------------------------------------------------------------------
// operators.cu: operator expressions
// ---------------------------------------------------------------------------
int op_matvec(sigb_matrix_t A, int trans, const double *x, double *y, bool add_to_y, const DotSpec &dot);
int op_jacobi_setup(sigb_matrix_t A, double *idiag);
void op_destroy(sigb_matrix_t A);

}  // namespace sigb
