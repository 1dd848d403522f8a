// spmv_device.cuh -- device side of the streaming CSR SpMV: argument block,
// TMA / mbarrier primitives and the per-CTA tile pass, shared by the stand-alone
// kernel (kernels_spmv.cu) and the persistent CG kernel (cg_persistent.cu).
#pragma once

#include <algorithm>

#include "device_utils.cuh"

namespace sigb {

struct CsrKernelArgs {
    const int32_t *ptr;       // 1-based, nrows + 1 (+ pad)
    const int32_t *node;      // 1-based (+ pad)
    const double *val;        // (+ pad)
    const TileDesc *tiles;
    int32_t ntiles;
    const double *x1;         // x - 1 : indexable by 1-based column id
    double *y;
    const double *u;          // dot vector
    double *out0, *out1;
    double *partials;
    unsigned *ticket;
    const int *skip_flag;
    const double *scale;      // optional per-row scaling of the result
    const double *h1;         // halo - (nloc + 1): indexable by column ids > nloc
    int32_t nloc;             // owned columns (HALO kernels)
    const double *add0, *add1;  // optional addends folded into the dot totals
    HaloSync sync;            // peer-memory transport (sync.win == nullptr: none)
    int32_t first_halo_tile;  // tiles from this index on read halo columns
    RedFuse red;              // finish the dots across the GPUs in this kernel (nranks <= 1: no)
    FaultBlock *fault;        // wait timeouts (device_utils.cuh)
#ifdef SIGB_PHASE_TIMERS
    // diagnostic build: SM cycles of thread 0 of every CTA, summed over the grid:
    // [0] waiting for the staged tile, [1] gathers issued + row sums of the previous tile,
    // [2] products (this is where the gathers are waited for), [3] the whole pass,
    // [4] staged tiles processed, [5] CTA passes
    unsigned long long *tile_dbg;
#endif
};

#ifdef SIGB_PHASE_TIMERS
#define SIGB_TCLK(var) const long long var = clock64()
#define SIGB_TACC(acc, t1, t0) acc += (unsigned long long)((t1) - (t0))
#else
#define SIGB_TCLK(var)
#define SIGB_TACC(acc, t1, t0)
#endif

__device__ __forceinline__ int4 load_desc(const TileDesc *t)
{
    return __ldg(reinterpret_cast<const int4 *>(t));
}

// ---- mbarrier / bulk-copy (TMA) primitives ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// The matrix arrays are read exactly once per SpMV: evict-first keeps them from
// displacing the vectors, which are re-read by the following kernels and fit
// in the 126 MB L2 once the operator is sharded.
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// shared-memory stage: val | node | ptr slices of one tile.  The kernels are compiled for two tile
// shapes (TileCfg<TN>, TN = staged entries per tile; a quarter as many rows, at least 128):
//   TN = 2048 (kTileNnz) -- short rows with locality (stencils, FEM): 4 CTAs x 52 KB per SM;
//   TN = kTileNnzSmall   -- long rows with scattered columns (random graphs): the SM's unified 256 KB
//     array is shared by shared memory and L1, and the RATE of uncoalesced gathers follows the L1 share
//     (pure-gather probe, profiles/r2_gather_probe.jsonl: 270 G gathers/s with <= 100 KB carved out for
//     shared memory, 247 at 132 KB, 197 at 164 KB, 101 at 228 KB -- whatever the number of warps), so a
//     gather-bound operator wants small stages, not deep ones.
// A table built for the smaller shape is valid for the larger kernel (the persistent CG kernel keeps
// the large one); the reverse is not.
template <int TN>
struct TileCfg {
    static constexpr int kNnz = TN;
    static constexpr int kCap = TN - 3;
    static constexpr int kRows = (TN / 4 < 128) ? 128 : TN / 4;
    static constexpr int kStageVal = TN * 8;
    static constexpr int kStageNode = TN * 4;
    static constexpr int kStagePtr = (kRows + 8) * 4;
    static constexpr int kStageBytes = (kStageVal + kStageNode + kStagePtr + 127) & ~127;
    static constexpr int kRowSlots = (kRows + kThreads - 1) / kThreads;   // rows of a tile a thread may own
    static constexpr int kEntrySlots = TN / kThreads;                     // entries of a tile a thread gathers
    static constexpr bool kBatchRowSums = TN != kTileNnz;                 // ordered_sum: products fetched eight at a time
    static_assert(TN % kThreads == 0 && TN >= kThreads, "tile entries: a multiple of the CTA size");
};
using TileCfgLarge = TileCfg<kTileNnz>;
using TileCfgSmall = TileCfg<kTileNnzSmall>;
static_assert(TileCfgLarge::kRows == kTileRows, "kTileRows is the row cap of the large shape");
constexpr int kStageBytes = TileCfgLarge::kStageBytes;   // callers that size the large shape's stages

template <class CFG>
__device__ __forceinline__ bool tile_staged(const int4 &d)
{
    return (d.w - d.z) <= CFG::kCap;  // else: one long row, streamed directly
}

// the vectors of one SpMV pass (the persistent kernel runs the same matrix over different ones)
struct SpmvVecs {
    const double *x1;   // x - 1
    double *y;
    const double *u;    // dot operand (NDOT >= 1)
};

// how the row sum z meets y(r) (SpmvMode), then the optional row scaling; returns what was stored
template <int MODE>
__device__ __forceinline__ double store_row(const CsrKernelArgs &a, const SpmvVecs &v, int r, double z)
{
    if (MODE == MODE_ADD_AFTER) z = add(v.y[r], z);
    if (MODE == MODE_SET && a.scale) z = mul(a.scale[r], z);
    v.y[r] = z;
    return z;
}
template <int NDOT>
__device__ __forceinline__ void dot_row(double ur, double z, double *acc)
{
    if (NDOT >= 1) acc[0] = add(acc[0], mul(ur, z));
    if (NDOT >= 2) acc[NDOT > 1 ? 1 : 0] = add(acc[NDOT > 1 ? 1 : 0], mul(z, z));
}

// XNC: x (and the dot operand u) may be read through the non-coherent read-only
// path.  True for stand-alone launches, where the vectors are constant for the
// kernel's lifetime; false inside the persistent CG kernel, which rewrites them
// between grid barriers and must use coherent loads.
template <bool XNC>
__device__ __forceinline__ double ld_x(const double *p)
{
    return XNC ? __ldg(p) : *p;
}

// one row longer than a tile: CTA-wide fixed-tree reduction, direct loads (the only case where
// the order of the additions differs from the reference's)
template <int MODE, int NDOT, bool HALO, bool XNC>
__device__ __forceinline__ void long_row(const CsrKernelArgs &a, const SpmvVecs &v, const int4 &d, const double *h1,
                                         double *acc)
{
    __shared__ double smr[1][kThreads / 32];
    double s[1] = {0.0};
    for (int k = d.z + threadIdx.x; k < d.w; k += kThreads) {
        const int c = a.node[k];
        const double xv = (HALO && c > a.nloc) ? __ldcg(h1 + c) : ld_x<XNC>(v.x1 + c);
        s[0] = add(s[0], mul(a.val[k], xv));
    }
    block_tree<1>(s, smr);
    if (threadIdx.x == 0) {
        double z = s[0];
        if (MODE == MODE_ACC_INIT) z = add(v.y[d.x], z);
        z = store_row<MODE>(a, v, d.x, z);
        dot_row<NDOT>(NDOT >= 1 ? ld_x<XNC>(v.u + d.x) : 0.0, z, acc);
    }
    __syncthreads();
}

template <int NDOT>
__device__ __forceinline__ void finish_dots(const CsrKernelArgs &a, double *acc)
{
    if (NDOT == 1) {
        double *const out[1] = {a.out0};
        const double *const addend[1] = {a.add0};
        double v[1] = {acc[0]};
        grid_reduce<1>(v, a.partials, a.ticket, out, addend, &a.red);
    } else if (NDOT == 2) {
        double *const out[2] = {a.out0, a.out1};
        const double *const addend[2] = {a.add0, a.add1};
        double v[2] = {acc[0], acc[NDOT > 1 ? 1 : 0]};
        grid_reduce<2>(v, a.partials, a.ticket, out, addend, &a.red);
    }
}

// z + p(b) + p(b+1) + ... + p(e-1), strictly left to right (the reference's z = z + val(k) * x(node(k))
// with the rounded products already formed).  The products are fetched from shared memory eight at a
// time BEFORE the dependent chain of additions starts: a plain loop exposes the shared-memory latency
// on every step (load, add, load, add ...), which made the row sums of matrices with ~20 entries per
// row the longest part of a tile (ncu on the Erdos-Renyi operator: warps mostly stalled at the CTA
// barrier behind the few threads that own rows, profiles/r2_ncu_er_2m_large_shape.txt).
// BATCH: the small tile shape (long rows); the large one keeps the plain loop -- with ~5 entries per row
// the batches never fill (A/B on the Poisson matrix: 225.3 us plain, 226.8 us batched,
// profiles/r2_visit_f_1gpu_summary.txt).
#ifndef SIGB_UR_STAGES
#define SIGB_UR_STAGES 3
#endif
template <bool BATCH>
__device__ __forceinline__ double ordered_sum(const double *sval, int b, int e, double z)
{
    if (!BATCH) {
        for (int k = b; k < e; k++) z = add(z, sval[k]);
        return z;
    }
    int k = b;
    for (; k + 8 <= e; k += 8) {
        double p[8];
#pragma unroll
        for (int j = 0; j < 8; j++) p[j] = sval[k + j];
#pragma unroll
        for (int j = 0; j < 8; j++) z = add(z, p[j]);
    }
    if (k < e) {
        double p[7];
#pragma unroll
        for (int j = 0; j < 7; j++) p[j] = (k + j < e) ? sval[k + j] : 0.0;
#pragma unroll
        for (int j = 0; j < 7; j++)
            if (k + j < e) z = add(z, p[j]);
    }
    return z;
}

// Per-CTA state of the double-buffered TMA pipeline, carried across calls.
struct TilePipe {
    unsigned sidx = 0;     // staged tiles consumed so far: stage = sidx & 1, mbarrier parity = (sidx >> 1) & 1
    bool primed = false;   // the first tile of the next pass has already been issued
};

// Row-sharded operators on the peer-memory transport: the first sync.push_ctas CTAs of the grid are
// COMMUNICATION CTAs.  They take no tiles: they store the owned entries other ranks need straight
// into those ranks' landing buffers (NVLink stores), fence at system scope and publish the sequence
// number of this SpMV -- off the critical path of the CTAs that stream the matrix.  (Round 1 let the
// first compute CTAs push before their own tiles, which made them the last to reach the end of the
// pass: 35.1 us against 30.3 us for the others, profiles/r2_visit_b_2gpu_summary.txt.)
__device__ __forceinline__ bool is_comm_cta(const CsrKernelArgs &a)
{
    return a.sync.win != nullptr && !a.sync.push_all && (int)blockIdx.x < a.sync.push_ctas;
}
// position of this CTA among the compute CTAs, and how many there are
__device__ __forceinline__ int compute_rank(const CsrKernelArgs &a)
{
    return (a.sync.win != nullptr && !a.sync.push_all) ? (int)blockIdx.x - a.sync.push_ctas : (int)blockIdx.x;
}
__device__ __forceinline__ int compute_ctas(const CsrKernelArgs &a)
{
    return (a.sync.win != nullptr && !a.sync.push_all) ? (int)gridDim.x - a.sync.push_ctas : (int)gridDim.x;
}

template <bool XNC>
__device__ __forceinline__ void halo_push(const CsrKernelArgs &a, const double *x1, unsigned long long hseq)
{
    const int tid = threadIdx.x;
    const int buf = (int)(hseq & 1);
    // landing buffer `buf` was last filled for SpMV hseq-2: wait until every consumer has
    // acknowledged reading it
    if (tid < kMaxRanks && hseq > 2 && (a.sync.dst_mask & (1u << tid))) {
        const unsigned long long *ack = &a.sync.win->ack[tid];
        spin_wait([&] { return ld_acquire_sys(ack) >= hseq - 2; }, a.fault, FAULT_HALO_ACK);
    }
    __syncthreads();
    for (int k = (int)blockIdx.x * kThreads + tid; k < a.sync.total_send; k += a.sync.push_ctas * kThreads) {
        int q = 0;
        while (k >= a.sync.send_off[q + 1]) q++;
        a.sync.dst[q][buf * a.sync.dst_stride[q] + (k - a.sync.send_off[q])] = ld_x<XNC>(x1 + a.sync.send_rows[k]);
    }
    __syncthreads();   // the CTA's stores happen-before thread 0's fence (cumulative)
    if (tid == 0) {
        __threadfence_system();
        const unsigned t0 = atomicAdd(&a.sync.win->push_ticket, 1u);
        if (t0 == (unsigned)a.sync.push_ctas - 1) {
            a.sync.win->push_ticket = 0u;
            __threadfence_system();
            for (int q = 0; q < kMaxRanks; q++)
                if (a.sync.dst_mask & (1u << q))
                    *reinterpret_cast<volatile unsigned long long *>(&a.sync.peer[q]->hflag[buf][a.sync.me]) = hseq;
        }
    }
}

// One SpMV pass of this CTA over its tiles (round-robin over the compute CTAs), including -- for
// row-sharded operators on the peer-memory transport -- the halo push by the communication CTAs and
// the lazy wait before the first boundary tile.  hseq is the sequence number of this SpMV (HALO
// only).  Dot partials accumulate in acc.
//
// Two-pass form, software-pipelined across tiles.  Per tile the TMA engine stages the val / node /
// ptr slices in shared memory (double-buffered, evict-first in L2);
//   gather  : every entry's x is gathered (kEntrySlots independent requests per thread);
//   product : the ROUNDED product replaces the value in shared memory;
//   row sums: one thread per row adds that row's products in STORED order
// and the gathers of tile t are issued BEFORE the row sums of tile t-1, so that their latency (one
// gather in five of the 5-point matrix is the first touch of an x line and comes from HBM, not from
// L1/L2) overlaps the row sums, the y stores and the barrier of the previous tile instead of stalling
// the product pass (round-2 diagnostic build of the unpipelined form: 46 % of a CTA's pass sat in
// the product pass, 2.6 % waiting for TMA):
//   iteration t:  wait stage(t) | gather x for t, load u for the rows of t-1 | row sums of t-1
//                 | barrier (stage(t-1) free) | TMA for t+1 into stage(t-1)
//                 | products of t into stage(t) | dot contributions of t-1 | barrier
// The fused dot is kept out of the row sums: u for the rows of tile t-1 is loaded next to the gathers of
// tile t and meets the finished row values only after the product pass of tile t, i.e. behind a point
// where every load of the iteration has landed
// -- inside the row sums it made them wait for the gathers in flight (288 us instead of 225 us).
// Same rounded products, added in the same order: bit-identical to the serial reference loop.
// Measured on the 4096^2 Poisson matrix: 226.6 -> 213.8 us (0.97 of the measured copy bandwidth),
// with the dot 229.3 -> 225.4 us (profiles/r2_visit_c_1gpu_summary.txt, r2_visit_d_1gpu_summary.txt).
template <int MODE, int NDOT, bool HALO, bool XNC, class CFG = TileCfgLarge>
__device__ __forceinline__ void spmv_phase(const CsrKernelArgs &a, const SpmvVecs &v, unsigned char *smem,
                                           uint64_t *mbar, TilePipe &pipe, double *acc, unsigned long long hseq,
                                           bool prime_next)
{
    constexpr int kStageBytes = CFG::kStageBytes, kStageVal = CFG::kStageVal, kStageNode = CFG::kStageNode;
    constexpr int kRowSlots = CFG::kRowSlots, kEntrySlots = CFG::kEntrySlots;
    const int tid = threadIdx.x;
    if (HALO && is_comm_cta(a)) {
        halo_push<XNC>(a, v.x1, hseq);
        return;
    }
    // no interior work to overlap with: every CTA pushes its share first (push_ctas == gridDim.x here)
    if (HALO && a.sync.win != nullptr && a.sync.push_all) halo_push<XNC>(a, v.x1, hseq);
    const int crank = HALO ? compute_rank(a) : (int)blockIdx.x;
    const int cstride = HALO ? compute_ctas(a) : (int)gridDim.x;
    const double *h1 = a.h1;
    bool halo_ready = !(HALO && a.sync.win != nullptr);

    // thread 0 is the producer: it programs the TMA engine for one tile
    const uint64_t stream_policy = policy_evict_first();
    auto issue = [&](int stage, const int4 &d) {
        unsigned char *base = smem + stage * kStageBytes;
        const int ka = d.z & ~3, ra = d.x & ~3;
        const uint32_t cnt = (uint32_t)((d.w - ka + 3) & ~3);
        const uint32_t rcnt = (uint32_t)((d.y + 1 - ra + 3) & ~3);
        mbar_expect_tx(&mbar[stage], cnt * 12u + rcnt * 4u);
        bulk_g2s(base, a.val + ka, cnt * 8u, &mbar[stage], stream_policy);
        bulk_g2s(base + kStageVal, a.node + ka, cnt * 4u, &mbar[stage], stream_policy);
        bulk_g2s(base + kStageVal + kStageNode, a.ptr + ra, rcnt * 4u, &mbar[stage], stream_policy);
    };

#ifdef SIGB_PHASE_TIMERS
    unsigned long long c_wait = 0, c_gather = 0, c_prod = 0, c_tiles = 0;
#endif
    SIGB_TCLK(tk_begin);
    int t = crank;
    int4 d_cur = make_int4(0, 0, 0, 0), d_next = make_int4(0, 0, 0, 0);
    if (t < a.ntiles) d_cur = load_desc(a.tiles + t);
    if (t + cstride < a.ntiles) d_next = load_desc(a.tiles + t + cstride);
    unsigned sidx = pipe.sidx;  // staged tiles consumed so far (identical in all threads)
    if (!pipe.primed && tid == 0 && t < a.ntiles && tile_staged<CFG>(d_cur)) issue((int)(sidx & 1u), d_cur);
    pipe.primed = false;

    // the tile whose products are parked in shared memory and whose rows have not been summed yet
    bool pending = false;
    int4 d_pend = make_int4(0, 0, 0, 0);
    int stage_pend = 0;
    // the tile whose rows have been summed and whose dot contributions have not been added yet
    bool dot_pending = false;
    int4 d_dot = make_int4(0, 0, 0, 0);
    double ur_dot[kRowSlots], z_dot[kRowSlots];
#pragma unroll
    for (int i = 0; i < kRowSlots; i++) { ur_dot[i] = 0.0; z_dot[i] = 0.0; }
#if SIGB_UR_STAGES == 3
    double ur_pend[kRowSlots];
#pragma unroll
    for (int i = 0; i < kRowSlots; i++) ur_pend[i] = 0.0;
#endif

    auto flush_dot = [&]() {
        if (NDOT >= 1 && dot_pending) {
#pragma unroll
            for (int i = 0; i < kRowSlots; i++)
                if (d_dot.x + tid + i * kThreads < d_dot.y) dot_row<NDOT>(ur_dot[i], z_dot[i], acc);
        }
        dot_pending = false;
    };
    auto load_u_pending = [&]() {
#if SIGB_UR_STAGES == 3
        if (NDOT >= 1) {
#pragma unroll
            for (int i = 0; i < kRowSlots; i++) ur_dot[i] = ur_pend[i];
        }
        return;
#endif
        if (NDOT >= 1) {
#pragma unroll
            for (int i = 0; i < kRowSlots; i++) {
                const int r = d_pend.x + tid + i * kThreads;
                ur_dot[i] = (r < d_pend.y) ? ld_x<XNC>(v.u + r) : 0.0;
            }
        }
    };
    // rows of the pending tile, summed in stored order
    auto row_sums = [&]() {
        unsigned char *base = smem + stage_pend * kStageBytes;
        const double *sval = reinterpret_cast<const double *>(base);
        const int32_t *sptr = reinterpret_cast<const int32_t *>(base + kStageVal + kStageNode);
        const int rs = d_pend.x, re = d_pend.y, ka = d_pend.z & ~3, ra = d_pend.x & ~3;
#pragma unroll
        for (int i = 0; i < kRowSlots; i++) {
            const int r = rs + tid + i * kThreads;
            if (r < re) {
                const int b = sptr[r - ra] - 1 - ka, e = sptr[r + 1 - ra] - 1 - ka;
                const double z = ordered_sum<CFG::kBatchRowSums>(sval, b, e, (MODE == MODE_ACC_INIT) ? v.y[r] : 0.0);
                z_dot[i] = store_row<MODE>(a, v, r, z);
            }
        }
        d_dot = d_pend;
        dot_pending = NDOT >= 1;
        pending = false;
        fence_proxy_async();           // the stage is overwritten by the async proxy next
    };

    for (; t < a.ntiles; t += cstride) {
        const int tn = t + cstride, tnn = tn + cstride;
        const bool have_next = tn < a.ntiles;
        int4 d_next2 = make_int4(0, 0, 0, 0);
        if (tnn < a.ntiles) d_next2 = load_desc(a.tiles + tnn);  // two tiles ahead, off the critical path
        const bool staged = tile_staged<CFG>(d_cur);

        // tiles [first_halo_tile, ntiles) read halo columns: wait for the peers' pushes of this
        // SpMV when the first one comes up (CTA-uniform; at the END of a CTA's pass, the tiles
        // being ordered interior first)
        if (HALO && !halo_ready && t >= a.first_halo_tile) {
            if (tid < kMaxRanks && (a.sync.src_mask & (1u << tid))) {
                const unsigned long long *flag = &a.sync.win->hflag[hseq & 1][tid];
                spin_wait([&] { return ld_acquire_sys(flag) >= hseq; }, a.fault, FAULT_HALO_FLAG);
            }
            __syncthreads();
            h1 = a.sync.halo_base + (hseq & 1) * a.sync.halo_stride - (a.nloc + 1);
            halo_ready = true;
        }

        if (staged) {
            const int stage = (int)(sidx & 1u);
            unsigned char *base = smem + stage * kStageBytes;
            double *sval = reinterpret_cast<double *>(base);
            const int32_t *snode = reinterpret_cast<const int32_t *>(base + kStageVal);
            const int ks = d_cur.z, ke = d_cur.w, ka = ks & ~3;
            const int cnt = ke - ka;
            SIGB_TCLK(tk0);
            mbar_wait(&mbar[stage], (sidx >> 1) & 1u);
            SIGB_TCLK(tk1);
            // ---- gathers of this tile (entries before ks belong to the previous tile and hold
            // valid columns, so every gather is in range)
            int c[kEntrySlots];
            double xv[kEntrySlots];
#pragma unroll
            for (int i = 0; i < kEntrySlots; i++) {
                const int k = tid + i * kThreads;
                c[i] = (k < cnt) ? snode[k] : 1;
            }
            // Only boundary tiles can hold halo columns, and whether a tile is one is uniform
            // over the CTA: interior tiles take the plain gather; in boundary tiles the address
            // is selected (no divergent branch between the gathers, which would serialise them)
            // and everything is read at L2 -- halo entries are written by peers, so L1 must not
            // serve them.
            if (HALO && t >= a.first_halo_tile) {
#pragma unroll
                for (int i = 0; i < kEntrySlots; i++) {
                    const double *src = (c[i] > a.nloc) ? h1 + c[i] : v.x1 + c[i];
                    xv[i] = __ldcg(src);
                }
            } else {
#pragma unroll
                for (int i = 0; i < kEntrySlots; i++) xv[i] = ld_x<XNC>(v.x1 + c[i]);
            }
            // u for the rows of the PENDING tile, next to the gathers: consumed after the product pass
            if (pending) load_u_pending();
#if SIGB_UR_STAGES == 3
            double ur_cur[kRowSlots];
#pragma unroll
            for (int i = 0; i < kRowSlots; i++) {
                const int r = d_cur.x + tid + i * kThreads;
                ur_cur[i] = (NDOT >= 1 && r < d_cur.y) ? ld_x<XNC>(v.u + r) : 0.0;
            }
#endif
            // ---- row sums of the previous tile while the gathers are in flight
            if (pending) row_sums();
            __syncthreads();                               // the other stage has been consumed by everybody
            if (tid == 0 && have_next && tile_staged<CFG>(d_next)) issue((int)((sidx + 1u) & 1u), d_next);
            SIGB_TCLK(tk2);
            // ---- products, in place
#pragma unroll
            for (int i = 0; i < kEntrySlots; i++) {
                const int k = tid + i * kThreads;
                if (k < cnt) sval[k] = mul(sval[k], xv[i]);
            }
            flush_dot();                                   // every load of this iteration has landed
            __syncthreads();
            SIGB_TCLK(tk3);
            SIGB_TACC(c_wait, tk1, tk0);
            SIGB_TACC(c_gather, tk2, tk1);
            SIGB_TACC(c_prod, tk3, tk2);
#ifdef SIGB_PHASE_TIMERS
            c_tiles++;
#endif
            pending = true;
            d_pend = d_cur;
            stage_pend = stage;
#if SIGB_UR_STAGES == 3
#pragma unroll
            for (int i = 0; i < kRowSlots; i++) ur_pend[i] = ur_cur[i];
#endif
            sidx++;
        } else {
            if (pending) {
                load_u_pending();
                row_sums();
                __syncthreads();
            }
            flush_dot();
            if (tid == 0 && have_next && tile_staged<CFG>(d_next)) issue((int)(sidx & 1u), d_next);   // both stages are free
            long_row<MODE, NDOT, HALO, XNC>(a, v, d_cur, h1, acc);
        }
        d_cur = d_next;
        d_next = d_next2;
    }
    if (pending) {
        load_u_pending();
        row_sums();
        __syncthreads();
    }
    flush_dot();
#ifdef SIGB_PHASE_TIMERS
    if (tid == 0 && a.tile_dbg != nullptr) {
        atomicAdd(a.tile_dbg + 0, c_wait);
        atomicAdd(a.tile_dbg + 1, c_gather);
        atomicAdd(a.tile_dbg + 2, c_prod);
        atomicAdd(a.tile_dbg + 3, (unsigned long long)(clock64() - tk_begin));
        atomicAdd(a.tile_dbg + 4, c_tiles);
        atomicAdd(a.tile_dbg + 5, 1ull);
    }
#endif
    pipe.sidx = sidx;
    // persistent callers: the matrix does not change between SpMVs, so the first
    // tile of the NEXT pass can already be in flight while other phases run
    if (prime_next) {
        if (crank < a.ntiles) {
            const int4 d0 = load_desc(a.tiles + crank);
            if (tid == 0 && tile_staged<CFG>(d0)) issue((int)(sidx & 1u), d0);
        }
        pipe.primed = true;
    }
}

// The tile a persistent caller primed for a pass that will not run must land before the CTA exits.
template <class CFG = TileCfgLarge>
__device__ __forceinline__ void drain_primed(const CsrKernelArgs &a, uint64_t *mbar, const TilePipe &pipe, bool halo)
{
    if (!pipe.primed) return;
    if (halo && is_comm_cta(a)) return;
    const int crank = halo ? compute_rank(a) : (int)blockIdx.x;
    if (crank < a.ntiles) {
        const int4 d0 = load_desc(a.tiles + crank);
        if (tile_staged<CFG>(d0)) mbar_wait(&mbar[pipe.sidx & 1u], (pipe.sidx >> 1) & 1u);
    }
}

}  // namespace sigb
