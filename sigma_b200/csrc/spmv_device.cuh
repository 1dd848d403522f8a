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
    RedFuse red;              // EXPERIMENTAL: finish the dots across the GPUs in this kernel (nranks <= 1: no)
#ifdef SIGB_PHASE_TIMERS
    // diagnostic build: SM cycles of thread 0 of every CTA, summed over the grid:
    // [0] waiting for the staged tile, [1] products (gathers), [2] row sums, [3] the whole pass,
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

// shared-memory stage: val | node | ptr slices of one tile
constexpr int kStageVal = kTileNnz * 8;
constexpr int kStageNode = kTileNnz * 4;
constexpr int kStagePtr = (kTileRows + 8) * 4;
constexpr int kStageBytes = (kStageVal + kStageNode + kStagePtr + 127) & ~127;

__device__ __forceinline__ bool tile_staged(const int4 &d)
{
    return (d.w - d.z) <= kTileCap;  // else: one long row, streamed directly
}

template <int MODE, int NDOT>
__device__ __forceinline__ void emit_row(const CsrKernelArgs &a, int r, double z, double ur, double *acc)
{
    if (MODE == MODE_ADD_AFTER) z = add(a.y[r], z);
    if (MODE == MODE_SET && a.scale) z = mul(a.scale[r], z);
    a.y[r] = z;
    if (NDOT >= 1) acc[0] = add(acc[0], mul(ur, z));
    if (NDOT >= 2) acc[NDOT > 1 ? 1 : 0] = add(acc[NDOT > 1 ? 1 : 0], mul(z, z));
}

// ---- EXPERIMENTAL fence-free halo (LL, opt-in through SIGB_HALO_LL=1, not yet run on a GPU) ------
// The landing buffers hold one 16-byte record per halo entry -- two words of 32 payload bits + the
// 32-bit sequence number of the SpMV, like the all-reduce inbox -- instead of bare doubles behind a
// per-source flag.  A word is delivered as a unit, so a record is either old or complete: the
// producer needs no system-scope fence and publishes nothing (that fence sits on the pushing CTAs'
// critical path today and makes them the last to reach the barrier), and a consumer simply polls
// the record it is about to gather.  Buffer reuse is still guarded by the acknowledgements.
__device__ __forceinline__ void halo_ll_store(RedEntry *e, unsigned seq, double v)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    st_word(&e->lo, (unsigned)bits, seq);
    st_word(&e->hi, (unsigned)(bits >> 32), seq);
}
__device__ __forceinline__ double halo_ll_load(const RedEntry *e, unsigned seq)
{
    uint2 lo, hi;
    unsigned spins = 0;
    do { lo = ld_word(&e->lo); } while (lo.y != seq && ++spins < kSpinLimit);
    do { hi = ld_word(&e->hi); } while (hi.y != seq && ++spins < kSpinLimit);
    return __longlong_as_double((long long)(((unsigned long long)hi.x << 32) | lo.x));
}

// one row longer than a tile: CTA-wide fixed-tree reduction, direct loads
// XNC: x (and the dot operand u) may be read through the non-coherent read-only
// path.  True for stand-alone launches, where the vectors are constant for the
// kernel's lifetime; false inside the persistent CG kernel, which rewrites them
// between grid barriers and must use coherent loads.
template <bool XNC>
__device__ __forceinline__ double ld_x(const double *p)
{
    return XNC ? __ldg(p) : *p;
}

template <int MODE, int NDOT, bool HALO, bool XNC, bool LL = false>
__device__ __forceinline__ void long_row(const CsrKernelArgs &a, const int4 &d, const double *h1, double *acc,
                                         const RedEntry *hll = nullptr, unsigned seq = 0)
{
    __shared__ double smr[1][kThreads / 32];
    double s[1] = {0.0};
    for (int k = d.z + threadIdx.x; k < d.w; k += kThreads) {
        const int c = a.node[k];
        double xv;
        if (LL && HALO && c > a.nloc) xv = halo_ll_load(hll + c, seq);
        else xv = (HALO && c > a.nloc) ? __ldcg(h1 + c) : ld_x<XNC>(a.x1 + c);
        s[0] = add(s[0], mul(a.val[k], xv));
    }
    block_tree<1>(s, smr);
    if (threadIdx.x == 0) {
        double z = s[0];
        if (MODE == MODE_ACC_INIT) z = add(a.y[d.x], z);
        emit_row<MODE, NDOT>(a, d.x, z, NDOT >= 1 ? ld_x<XNC>(a.u + d.x) : 0.0, acc);
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

// Per-CTA state of the double-buffered TMA pipeline, carried across calls.
struct TilePipe {
    unsigned sidx = 0;     // staged tiles consumed so far: stage = sidx & 1, mbarrier parity = (sidx >> 1) & 1
    bool primed = false;   // the first tile of the next pass has already been issued
};

// One SpMV pass of this CTA over its tiles (round-robin), including -- for
// row-sharded operators on the peer-memory transport -- the halo push in the
// prologue and the lazy wait before the first boundary tile.  hseq is the
// sequence number of this SpMV (HALO only).  Dot partials accumulate in acc.
//
// RD ("row direct", EXPERIMENTAL, opt-in through SIGB_SPMV_ROWDIRECT, not yet run on a GPU):
// skip the pass that parks the rounded products in shared memory; the thread that owns a row
// forms each product itself, in stored order, from the staged val / node slices.  Same
// rounded products added in the same order, so the result is bit-identical.  Why: with ~5
// entries per row the two-pass form costs ~28 bytes of shared-memory traffic per entry plus a
// gather whose 32 lanes touch 6-7 sectors, and L1TEX (which serves both) is the most utilised
// unit of this kernel (67-69 %, profiles/r1_ncu_csr_tma.txt); row-direct needs 12 bytes per
// entry, one barrier less per tile, and on banded matrices the j-th entries of consecutive
// rows are consecutive columns, so the gathers coalesce like the ELLPACK kernel's.  It loses
// when rows are long or ragged (a warp runs as long as its longest row), hence a per-matrix
// choice on the host.
template <int MODE, int NDOT, bool HALO, bool XNC, bool RD = false, bool LL = false>
__device__ __forceinline__ void spmv_phase(const CsrKernelArgs &a, unsigned char *smem, uint64_t *mbar,
                                           TilePipe &pipe, double *acc, unsigned long long hseq,
                                           bool prime_next)
{
    // ---- peer-memory halo exchange, producer side ------------------------
    // tiles [first_halo_tile, ntiles) read halo columns.  This SpMV has
    // sequence number hseq; halo_seq is only advanced by the last CTA to
    // finish, so every CTA reads the same value here.
    const int tid = threadIdx.x;
    const double *h1 = a.h1;
    const RedEntry *hll = nullptr;    // LL: records of the landing buffer, indexable by column ids > nloc
    bool halo_ready = !(HALO && a.sync.win != nullptr);
    bool push_pending = false;
    if (HALO && a.sync.win != nullptr) {
        const int push_rank = (int)blockIdx.x - a.sync.push_first;   // position among the pushing CTAs
        if (push_rank >= 0 && push_rank < a.sync.push_ctas) {
            const int buf = (int)(hseq & 1);
            // landing buffer `buf` was last filled for SpMV hseq-2: wait until
            // every consumer has acknowledged reading it
            if (tid == 0 && hseq > 2)
                for (int q = 0; q < kMaxRanks; q++)
                    if (a.sync.dst_mask & (1u << q)) {
                        unsigned spins = 0;
                        while (ld_acquire_sys(&a.sync.win->ack[q]) < hseq - 2 && ++spins < kSpinLimit) {}
                    }
            __syncthreads();
            for (int k = push_rank * kThreads + tid; k < a.sync.total_send; k += a.sync.push_ctas * kThreads) {
                int q = 0;
                while (k >= a.sync.send_off[q + 1]) q++;
                if (LL)
                    halo_ll_store(reinterpret_cast<RedEntry *>(a.sync.dst[q]) + buf * a.sync.dst_stride[q] +
                                      (k - a.sync.send_off[q]),
                                  (unsigned)hseq, ld_x<XNC>(a.x1 + a.sync.send_rows[k]));
                else
                    a.sync.dst[q][buf * a.sync.dst_stride[q] + (k - a.sync.send_off[q])] =
                        ld_x<XNC>(a.x1 + a.sync.send_rows[k]);
            }
            push_pending = !LL;    // published after the first tile, see below (LL: nothing to publish)
        }
    }
    // The stores above need a system-scope fence before the sequence number may
    // be published.  Issued right away that fence costs the pushing CTAs a few
    // microseconds of NVLink round trip before they touch their first tile (and
    // the whole grid waits for them at the end); issued after the first tile the
    // stores have long landed and the fence is cheap.  Consumers only look at
    // the flags when they reach their boundary tiles, at the END of their pass.
    auto publish_push = [&]() {
        __syncthreads();   // the CTA's stores happen-before thread 0's fence (cumulative)
        if (tid == 0) {
            const int buf = (int)(hseq & 1);
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
        push_pending = false;
    };

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
    unsigned long long c_wait = 0, c_prod = 0, c_rows = 0, c_tiles = 0;
#endif
    SIGB_TCLK(tk_begin);
    int t = blockIdx.x;
    int4 d_cur = make_int4(0, 0, 0, 0), d_next = make_int4(0, 0, 0, 0);
    if (t < a.ntiles) d_cur = load_desc(a.tiles + t);
    if (t + (int)gridDim.x < a.ntiles) d_next = load_desc(a.tiles + t + gridDim.x);
    unsigned sidx = pipe.sidx;  // staged tiles consumed so far (identical in all threads)
    if (!pipe.primed && tid == 0 && t < a.ntiles && tile_staged(d_cur)) issue((int)(sidx & 1u), d_cur);
    pipe.primed = false;

#ifdef SIGB_SPMV_PIPE
    // ---- software-pipelined two-pass form (build variant _pipe) -------------------------------
    // The x gathers of tile t are issued BEFORE the row sums of tile t-1 are formed, so that the
    // latency of the gathers (one in five is the first touch of an x line and comes from HBM, not
    // from L1/L2) overlaps the row sums, the y stores and the barrier of the previous tile instead
    // of stalling the product pass; the gathered values wait in registers.  Same rounded products,
    // added in the same stored order: results are bit-identical to the unpipelined form.
    //   iteration t:  wait stage(t) | gather x for t | row sums of t-1 | barrier (stage(t-1) free)
    //                 | TMA for t+1 into stage(t-1) | products of t into stage(t) | barrier
    if (!RD) {
        bool pending = false;              // a tile whose products are parked and whose rows are not summed yet
        int4 d_pend = make_int4(0, 0, 0, 0);
        int stage_pend = 0;
        auto row_sums = [&](const int4 &d, int stage) {
            unsigned char *base = smem + stage * kStageBytes;
            const double *sval = reinterpret_cast<const double *>(base);
            const int32_t *sptr = reinterpret_cast<const int32_t *>(base + kStageVal + kStageNode);
            const int rs = d.x, re = d.y, ka = d.z & ~3, ra = d.x & ~3;
            double ur[kTileRows / kThreads];
#pragma unroll
            for (int i = 0; i < kTileRows / kThreads; i++) {
                const int r = rs + tid + i * kThreads;
                ur[i] = (NDOT >= 1 && r < re) ? ld_x<XNC>(a.u + r) : 0.0;
            }
#pragma unroll
            for (int i = 0; i < kTileRows / kThreads; i++) {
                const int r = rs + tid + i * kThreads;
                if (r < re) {
                    const int b = sptr[r - ra] - 1 - ka, e = sptr[r + 1 - ra] - 1 - ka;
                    double z = (MODE == MODE_ACC_INIT) ? a.y[r] : 0.0;
                    for (int k = b; k < e; k++) z = add(z, sval[k]);
                    emit_row<MODE, NDOT>(a, r, z, ur[i], acc);
                }
            }
            fence_proxy_async();           // the stage is overwritten by the async proxy next
        };
        for (; t < a.ntiles; t += gridDim.x) {
            const int tn = t + gridDim.x, tnn = tn + gridDim.x;
            const bool have_next = tn < a.ntiles;
            int4 d_next2 = make_int4(0, 0, 0, 0);
            if (tnn < a.ntiles) d_next2 = load_desc(a.tiles + tnn);
            const bool staged = tile_staged(d_cur);
            if (HALO && !halo_ready && t >= a.first_halo_tile) {   // CTA-uniform, as in the unpipelined form
                if (LL) {
                    hll = reinterpret_cast<const RedEntry *>(a.sync.halo_base) + (hseq & 1) * a.sync.halo_stride -
                          (a.nloc + 1);
                } else {
                    if (push_pending) publish_push();
                    if (tid == 0) {
                        for (int q = 0; q < kMaxRanks; q++)
                            if (a.sync.src_mask & (1u << q)) {
                                unsigned spins = 0;
                                while (ld_acquire_sys(&a.sync.win->hflag[hseq & 1][q]) < hseq && ++spins < kSpinLimit) {}
                            }
                    }
                    __syncthreads();
                    h1 = a.sync.halo_base + (hseq & 1) * a.sync.halo_stride - (a.nloc + 1);
                }
                halo_ready = true;
            }
            if (staged) {
                const int stage = (int)(sidx & 1u);
                unsigned char *base = smem + stage * kStageBytes;
                double *sval = reinterpret_cast<double *>(base);
                const int32_t *snode = reinterpret_cast<const int32_t *>(base + kStageVal);
                const int ks = d_cur.z, ke = d_cur.w, ka = ks & ~3;
                const int cnt = ke - ka;
                mbar_wait(&mbar[stage], (sidx >> 1) & 1u);
                int c[kTileNnz / kThreads];
                double xv[kTileNnz / kThreads];
#pragma unroll
                for (int i = 0; i < kTileNnz / kThreads; i++) {
                    const int k = tid + i * kThreads;
                    c[i] = (k < cnt) ? snode[k] : 1;
                }
                if (LL && HALO && t >= a.first_halo_tile) {
#pragma unroll
                    for (int i = 0; i < kTileNnz / kThreads; i++) {
                        const int k = tid + i * kThreads;
                        xv[i] = (c[i] > a.nloc && k >= ks - ka && k < cnt) ? halo_ll_load(hll + c[i], (unsigned)hseq)
                                                                           : __ldcg(a.x1 + min(c[i], a.nloc));
                    }
                } else if (HALO && t >= a.first_halo_tile) {
#pragma unroll
                    for (int i = 0; i < kTileNnz / kThreads; i++) {
                        const double *src = (c[i] > a.nloc) ? h1 + c[i] : a.x1 + c[i];
                        xv[i] = __ldcg(src);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < kTileNnz / kThreads; i++) xv[i] = ld_x<XNC>(a.x1 + c[i]);
                }
                if (pending) row_sums(d_pend, stage_pend);     // ... while the gathers are in flight
                __syncthreads();                               // the other stage has been consumed by everybody
                if (tid == 0 && have_next && tile_staged(d_next)) issue((int)((sidx + 1u) & 1u), d_next);
#pragma unroll
                for (int i = 0; i < kTileNnz / kThreads; i++) {
                    const int k = tid + i * kThreads;
                    if (k < cnt) sval[k] = mul(sval[k], xv[i]);
                }
                __syncthreads();
                pending = true;
                d_pend = d_cur;
                stage_pend = stage;
                sidx++;
            } else {
                if (pending) {
                    row_sums(d_pend, stage_pend);
                    __syncthreads();
                    pending = false;
                }
                if (tid == 0 && have_next && tile_staged(d_next)) issue((int)(sidx & 1u), d_next);   // both stages are free
                long_row<MODE, NDOT, HALO, XNC, LL>(a, d_cur, h1, acc, hll, (unsigned)hseq);
            }
            d_cur = d_next;
            d_next = d_next2;
            if (HALO && push_pending) publish_push();
        }
        if (pending) {
            row_sums(d_pend, stage_pend);
            __syncthreads();
        }
    }
#endif
    for (; t < a.ntiles; t += gridDim.x) {
        const int tn = t + gridDim.x, tnn = tn + gridDim.x;
        const bool have_next = tn < a.ntiles;
        int4 d_next2 = make_int4(0, 0, 0, 0);
        if (tnn < a.ntiles) d_next2 = load_desc(a.tiles + tnn);  // two tiles ahead, off the critical path
        const bool staged = tile_staged(d_cur);
        const unsigned sidx_after = sidx + (staged ? 1u : 0u);
        if (tid == 0 && have_next && tile_staged(d_next)) issue((int)(sidx_after & 1u), d_next);

        if (HALO && !halo_ready && t >= a.first_halo_tile) {   // CTA-uniform
            // never wait on peers while our own push is unpublished (two ranks
            // whose pushing CTAs start on a boundary tile would wait forever)
            if (LL) {
                // nothing to wait for here: every gathered record is polled where it is read
                hll = reinterpret_cast<const RedEntry *>(a.sync.halo_base) + (hseq & 1) * a.sync.halo_stride -
                      (a.nloc + 1);
            } else {
                if (push_pending) publish_push();
                if (tid == 0) {
                    for (int q = 0; q < kMaxRanks; q++)
                        if (a.sync.src_mask & (1u << q)) {
                            unsigned spins = 0;
                            while (ld_acquire_sys(&a.sync.win->hflag[hseq & 1][q]) < hseq && ++spins < kSpinLimit) {}
                        }
                }
                __syncthreads();
                h1 = a.sync.halo_base + (hseq & 1) * a.sync.halo_stride - (a.nloc + 1);
            }
            halo_ready = true;
        }

        if (staged) {
            const int stage = (int)(sidx & 1u);
            unsigned char *base = smem + stage * kStageBytes;
            double *sval = reinterpret_cast<double *>(base);
            const int32_t *snode = reinterpret_cast<const int32_t *>(base + kStageVal);
            const int32_t *sptr = reinterpret_cast<const int32_t *>(base + kStageVal + kStageNode);
            const int rs = d_cur.x, re = d_cur.y, ks = d_cur.z, ke = d_cur.w;
            const int ka = ks & ~3, ra = rs & ~3;
            SIGB_TCLK(tk0);
            mbar_wait(&mbar[stage], (sidx >> 1) & 1u);
            SIGB_TCLK(tk1);
            if (RD) {
                // ---- row direct: one thread per row, products formed inline -----
                SIGB_TCLK(tk2);
                double ur[kTileRows / kThreads];
#pragma unroll
                for (int i = 0; i < kTileRows / kThreads; i++) {
                    const int r = rs + tid + i * kThreads;
                    ur[i] = (NDOT >= 1 && r < re) ? ld_x<XNC>(a.u + r) : 0.0;
                }
                const bool boundary = HALO && t >= a.first_halo_tile;   // CTA-uniform
#pragma unroll
                for (int i = 0; i < kTileRows / kThreads; i++) {
                    const int r = rs + tid + i * kThreads;
                    if (r < re) {
                        const int b = sptr[r - ra] - 1 - ka, e = sptr[r + 1 - ra] - 1 - ka;
                        double z = (MODE == MODE_ACC_INIT) ? a.y[r] : 0.0;
                        for (int k = b; k < e; k += 4) {
                            // Four entries per trip.  Indices past the row's end are clamped to its
                            // last entry (a valid read) and their products are not added, so there is
                            // no branch inside the trip: the four gathers are issued back to back
                            // (with per-entry branches the compiler serialised them behind the adds).
                            int c[4];
                            double v[4], xv[4], pr[4];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const int kk = min(k + j, e - 1);
                                c[j] = snode[kk];
                                v[j] = sval[kk];
                            }
                            if (boundary && LL) {
#pragma unroll
                                for (int j = 0; j < 4; j++)
                                    xv[j] = (c[j] > a.nloc) ? halo_ll_load(hll + c[j], (unsigned)hseq)
                                                            : __ldcg(a.x1 + c[j]);
                            } else if (boundary) {
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    const double *src = (c[j] > a.nloc) ? h1 + c[j] : a.x1 + c[j];
                                    xv[j] = __ldcg(src);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; j++) xv[j] = ld_x<XNC>(a.x1 + c[j]);
                            }
#pragma unroll
                            for (int j = 0; j < 4; j++) pr[j] = mul(v[j], xv[j]);
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const double zn = add(z, pr[j]);
                                z = (k + j < e) ? zn : z;
                            }
                        }
                        emit_row<MODE, NDOT>(a, r, z, ur[i], acc);
                    }
                }
                fence_proxy_async();
                __syncthreads();
                SIGB_TCLK(tk3);
                SIGB_TACC(c_wait, tk1, tk0);
                SIGB_TACC(c_prod, tk2, tk1);
                SIGB_TACC(c_rows, tk3, tk2);
#ifdef SIGB_PHASE_TIMERS
                c_tiles++;
#endif
            } else {
            // ---- phase 1: products, in place --------------------------------
            // (entries before ks belong to the previous tile and hold valid
            // columns, so every gather below is in range)
            const int cnt = ke - ka;
            int c[kTileNnz / kThreads];
            double v[kTileNnz / kThreads];
#pragma unroll
            for (int i = 0; i < kTileNnz / kThreads; i++) {
                const int k = tid + i * kThreads;
                c[i] = (k < cnt) ? snode[k] : 1;
                v[i] = (k < cnt) ? sval[k] : 0.0;
            }
            double xv[kTileNnz / kThreads];
            // Only boundary tiles can hold halo columns, and whether a tile is one
            // is uniform over the CTA: interior tiles take the plain gather; in
            // boundary tiles the address is selected (no divergent branch between
            // the gathers, which would serialise them) and everything is read at
            // L2 -- halo entries are written by peers, so L1 must not serve them.
            if (LL && HALO && t >= a.first_halo_tile) {
                // (entries of the previous tile that share this tile's first 16-byte group are
                //  multiplied too and never used: they must not be waited for)
#pragma unroll
                for (int i = 0; i < kTileNnz / kThreads; i++) {
                    const int k = tid + i * kThreads;
                    xv[i] = (c[i] > a.nloc && k >= ks - ka && k < cnt) ? halo_ll_load(hll + c[i], (unsigned)hseq)
                                                                       : __ldcg(a.x1 + min(c[i], a.nloc));
                }
            } else if (HALO && t >= a.first_halo_tile) {
#pragma unroll
                for (int i = 0; i < kTileNnz / kThreads; i++) {
                    const double *src = (c[i] > a.nloc) ? h1 + c[i] : a.x1 + c[i];
                    xv[i] = __ldcg(src);
                }
            } else {
#pragma unroll
                for (int i = 0; i < kTileNnz / kThreads; i++) xv[i] = ld_x<XNC>(a.x1 + c[i]);
            }
#pragma unroll
            for (int i = 0; i < kTileNnz / kThreads; i++) {
                const int k = tid + i * kThreads;
                if (k < cnt) sval[k] = mul(v[i], xv[i]);
            }
            __syncthreads();
            SIGB_TCLK(tk2);
            // ---- phase 2: per-row sums in stored order ----------------------
            // operands of the fused dot go through the same read-only path as
            // the gathers, so rows with a (near-)diagonal entry hit L1
            double ur[kTileRows / kThreads];
#pragma unroll
            for (int i = 0; i < kTileRows / kThreads; i++) {
                const int r = rs + tid + i * kThreads;
                ur[i] = (NDOT >= 1 && r < re) ? ld_x<XNC>(a.u + r) : 0.0;
            }
#pragma unroll
            for (int i = 0; i < kTileRows / kThreads; i++) {
                const int r = rs + tid + i * kThreads;
                if (r < re) {
                    const int b = sptr[r - ra] - 1 - ka, e = sptr[r + 1 - ra] - 1 - ka;
                    double z = (MODE == MODE_ACC_INIT) ? a.y[r] : 0.0;
                    for (int k = b; k < e; k++) z = add(z, sval[k]);
                    emit_row<MODE, NDOT>(a, r, z, ur[i], acc);
                }
            }
            // the stage is overwritten by the async proxy next: order our
            // generic-proxy accesses before it
            fence_proxy_async();
            __syncthreads();
            SIGB_TCLK(tk3);
            SIGB_TACC(c_wait, tk1, tk0);
            SIGB_TACC(c_prod, tk2, tk1);
            SIGB_TACC(c_rows, tk3, tk2);
#ifdef SIGB_PHASE_TIMERS
            c_tiles++;
#endif
            }   // !RD
        } else {
            long_row<MODE, NDOT, HALO, XNC, LL>(a, d_cur, h1, acc, hll, (unsigned)hseq);
        }
        sidx = sidx_after;
        d_cur = d_next;
        d_next = d_next2;
        if (HALO && push_pending) publish_push();
    }
    if (HALO && push_pending) publish_push();   // a CTA without tiles
#ifdef SIGB_PHASE_TIMERS
    if (tid == 0 && a.tile_dbg != nullptr) {
        atomicAdd(a.tile_dbg + 0, c_wait);
        atomicAdd(a.tile_dbg + 1, c_prod);
        atomicAdd(a.tile_dbg + 2, c_rows);
        atomicAdd(a.tile_dbg + 3, (unsigned long long)(clock64() - tk_begin));
        atomicAdd(a.tile_dbg + 4, c_tiles);
        atomicAdd(a.tile_dbg + 5, 1ull);
    }
#endif
    pipe.sidx = sidx;
    // persistent callers: the matrix does not change between SpMVs, so the first
    // tile of the NEXT pass can already be in flight while other phases run
    if (prime_next) {
        if (blockIdx.x < (unsigned)a.ntiles) {
            const int4 d0 = load_desc(a.tiles + blockIdx.x);
            if (tid == 0 && tile_staged(d0)) issue((int)(sidx & 1u), d0);
        }
        pipe.primed = true;
    }
}

}  // namespace sigb
