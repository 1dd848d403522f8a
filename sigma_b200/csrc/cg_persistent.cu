// cg_persistent.cu -- the whole CG / Jacobi-PCG loop of
//   cg_solve      src/solver/cg_solvers.f90:133-146
//   cg_solve_pc   src/solver/cg_solvers.f90:174-190  (pc = jacobi_solver)
// as ONE persistent cooperative kernel.
//
// Why: on a sharded operator an iteration is ~50 us of memory traffic per GPU;
// three kernel boundaries plus two dependent all-reduce launches cost as much
// again.  Here the CTAs stay resident (one cooperative wave, 4 per SM); the
// phases of an iteration are separated by grid barriers, the dot products are
// completed inside the barrier (every CTA adds the CTA partials in a fixed
// order; across GPUs CTA 0 stores the local sum into the peers' inboxes with
// fence-free 8-byte payload+flag words over NVLink and every CTA polls its own
// inbox), the halo push / wait / acknowledge run inside the SpMV phase, and
// the first matrix tile of the next SpMV is already in flight (TMA) while the
// vector phases run.  Recurrence, statement order, rounding (no FMA) and the
// stopping rule are those of the reference; all ranks compute bit-identical
// scalars and therefore leave the loop at the same iteration.
//
//   iteration:  A  q = A p, partial p.q            (spmv_phase, tiles round-robin)
//               -- barrier + all-reduce --         alpha = res2 / (p.q)
//               B  r -= alpha q ; [z = idiag r] ; partial r.r | r.z
//               -- barrier + all-reduce --         beta = dpr / res2 ; stop test
//               C  x += alpha p ; p = (r | z) + beta p     (p is read once for both)
//               -- barrier --                      (p complete before the next gathers)
// Per iteration this moves 12 nnz + 84 n bytes (the reference's statement order
// costs 92 n: it reads p for the x update and again for the p update).
#include <math.h>
#include <stdlib.h>

#include "krylov.cuh"
#include "solvers.h"
#include "spmv_device.cuh"

namespace sigb {

// Resident CTAs per SM the persistent kernels are compiled for.  4 is what shared memory allows and
// caps them at 64 registers, which they exceed (ptxas spills 36-124 bytes); the experiment build
// `make VARIANT=_pb3 DEFS=-DSIGB_PERSIST_MINBLOCKS=3` trades a quarter of the CTAs for 80 registers.
#ifndef SIGB_PERSIST_MINBLOCKS
#define SIGB_PERSIST_MINBLOCKS 4
#endif

constexpr int kPhaseSlots = 8;   // phases timed by SIGB_PHASE_TIMERS builds (see PhaseClock)
constexpr int kPhaseCtas = 3;    // first, middle, last CTA

struct CgPersistArgs {
    CsrKernelArgs A;            // matrix, tile table, x1 = p - 1, y = q, u = p, halo sync
    double *x, *p, *q, *r, *z;
    const double *idiag;        // PC only
    int64_t n;
    KState *st;
    unsigned long long *bar;    // grid barrier counter, zeroed before the launch
    double *partials;           // 2 x gridDim CTA partial sums
    long long max_iters;        // iterations this launch may run before handing back to the host
    RedWin *red;                // all-reduce inbox (nranks > 1)
    RedWin *peer_red[kMaxRanks];
    int me, nranks;
    unsigned long long *phase_dbg;  // SIGB_PHASE_TIMERS builds: kPhaseCtas x kPhaseSlots cycle counters
};

namespace {

struct Sync {
    unsigned long long *bar;
    unsigned long long epoch;   // barriers passed (uniform across the grid)
    int *abort_flag;
};

// Diagnostic build only (make VARIANT=_timers DEFS=-DSIGB_PHASE_TIMERS, loaded with
// SIGB_LIB_VARIANT=_timers): where an iteration of the persistent kernel spends its time.
// Three CTAs (first, middle, last) accumulate SM cycles per phase in registers and add them
// to phase_dbg on exit; the product build compiles all of it away.
//   0 spmv   1 barrier of reduction 1   2 cross-GPU part of reduction 1   3 phase B
//   4 barrier of reduction 2   5 cross-GPU part of reduction 2   6 phase C   7 closing barrier
struct PhaseClock {
#ifdef SIGB_PHASE_TIMERS
    long long last = 0;
    unsigned long long acc[kPhaseSlots] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool on = false;
    __device__ __forceinline__ void start(bool enable) { on = enable; if (on) last = clock64(); }
    __device__ __forceinline__ void stamp(int k)
    {
        if (on) { const long long t = clock64(); acc[k] += (unsigned long long)(t - last); last = t; }
    }
    __device__ __forceinline__ void flush(unsigned long long *dst, int which, long long iters)
    {
        if (on && dst) {
            for (int k = 0; k < kPhaseSlots; k++) atomicAdd(dst + which * (kPhaseSlots + 1) + k, acc[k]);
            atomicAdd(dst + which * (kPhaseSlots + 1) + kPhaseSlots, (unsigned long long)iters);
        }
    }
#else
    __device__ __forceinline__ void start(bool) {}
    __device__ __forceinline__ void stamp(int) {}
    __device__ __forceinline__ void flush(unsigned long long *, int, long long) {}
#endif
};

__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Grid-wide barrier with release/acquire semantics (the scheme cooperative
// groups uses: bar.sync ; thread 0: fence, arrive, wait, fence ; bar.sync), with
// arrival and release on different words: CTAs arrive with an atomic on bar[0];
// the last one to arrive publishes the epoch in bar[1], which is what everybody
// else polls -- the pollers do not compete with the arrival atomics.
// The trailing fence also drops stale L1 lines, so ordinary loads issued after
// the barrier observe what other SMs wrote before it.
__device__ __forceinline__ void grid_barrier(Sync &s)
{
    __syncthreads();
    s.epoch++;
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long arrived = atomicAdd(s.bar, 1ull) + 1;
        if (arrived == s.epoch * gridDim.x) {
            st_release_gpu(s.bar + 1, s.epoch);
        } else {
            unsigned spins = 0;
            while (ld_acquire_gpu(s.bar + 1) < s.epoch && ++spins < kSpinLimit) {}
        }
        __threadfence();
    }
    __syncthreads();
}

// Sum of one value per CTA over the grid (fixed order, identical in every
// CTA), then over the ranks (rank order).  Contains one grid barrier.
__device__ __forceinline__ double grid_allreduce(double v, const CgPersistArgs &a, Sync &s, int &pbuf,
                                                 unsigned long long &red_seq, double (*sm)[kThreads / 32],
                                                 double *s_bcast, PhaseClock &clk, int slot0)
{
    double acc[1] = {v};
    block_tree<1>(acc, sm);
    double *part = a.partials + (size_t)pbuf * gridDim.x;
    if (threadIdx.x == 0) part[blockIdx.x] = acc[0];
    grid_barrier(s);
    clk.stamp(slot0);
    double t[1] = {0.0};
    for (unsigned j = threadIdx.x; j < gridDim.x; j += kThreads) t[0] = add(t[0], __ldcg(part + j));
    block_tree<1>(t, sm);
    pbuf ^= 1;
    if (a.nranks > 1) {
        red_seq++;
        const int slot = (int)(red_seq & (kRedSlots - 1));
        const unsigned flag = (unsigned)red_seq;
        // published by the LAST CTA: CTA 0 also pushes halo entries and would be
        // the latest to get here
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x < 32) {
            const double local = __shfl_sync(0xffffffffu, t[0], 0);
            if ((int)threadIdx.x < a.nranks) {
                RedEntry *e = a.peer_red[threadIdx.x]->red[slot][a.me];
                const unsigned long long bits = (unsigned long long)__double_as_longlong(local);
                st_word(&e[0].lo, (unsigned)bits, flag);
                st_word(&e[0].hi, (unsigned)(bits >> 32), flag);
            }
        }
        if (threadIdx.x == 0) {
            double g = 0.0;
            for (int q = 0; q < a.nranks; q++) {
                const RedEntry *e = &a.red->red[slot][q][0];
                uint2 lo, hi;
                unsigned spins = 0;
                do { lo = ld_word(&e->lo); } while (lo.y != flag && ++spins < kSpinLimit);
                do { hi = ld_word(&e->hi); } while (hi.y != flag && ++spins < kSpinLimit);
                g = add(g, __longlong_as_double((long long)(((unsigned long long)hi.x << 32) | lo.x)));
            }
            *s_bcast = g;
        }
    } else {
        if (threadIdx.x == 0) *s_bcast = t[0];
    }
    __syncthreads();
    const double out = *s_bcast;
    __syncthreads();
    clk.stamp(slot0 + 1);
    return out;
}

template <bool HALO, bool PC, bool RD, bool LL = false>
__global__ void __launch_bounds__(kThreads, SIGB_PERSIST_MINBLOCKS)
cg_persistent_kernel(const CgPersistArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ double sm_red[1][kThreads / 32];
    __shared__ double s_bcast;

    KState *st = a.st;
    if (st->done[0]) return;   // loop test failed before the first pass (uniform)

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    double rr = st->rr[0];
    const double tol = st->tol;
    const long long cap = st->cap, it0 = st->iters;
    long long it = 0;
    Sync s{a.bar, 0ull, nullptr};
    int pbuf = 0;
    unsigned long long red_seq = a.nranks > 1 ? a.red->red_seq : 0ull;
    unsigned long long hseq = 0;
    if (HALO && a.A.sync.win != nullptr)
        hseq = *reinterpret_cast<volatile unsigned long long *>(&a.A.sync.win->halo_seq);
    TilePipe pipe;
    bool stop = false, capped = false;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    const double *rz = PC ? a.z : a.r;
    PhaseClock clk;
    const int clk_which = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x / 2 ? 1 : (blockIdx.x == gridDim.x - 1 ? 2 : -1));
    clk.start(tid == 0 && clk_which >= 0);

    for (;;) {
        // ---- A: q = A p, p.q ------------------------------------------------
        double acc[1] = {0.0};
        hseq++;
        spmv_phase<MODE_SET, 1, HALO, false, RD, LL>(a.A, smem, mbar, pipe, acc, hseq, true);
        clk.stamp(0);
        const double pq = grid_allreduce(acc[0], a, s, pbuf, red_seq, sm_red, &s_bcast, clk, 1);
        if (HALO && a.A.sync.win != nullptr && blockIdx.x == gridDim.x - 1 && tid < kMaxRanks &&
            (a.A.sync.src_mask & (1u << tid))) {
            // every CTA is past the barrier, i.e. has consumed this landing buffer
            // (its loads have completed); one lane per source rank acknowledges
            *reinterpret_cast<volatile unsigned long long *>(&a.A.sync.peer[tid]->ack[a.A.sync.me]) = hseq;
        }
        const double alpha = rr / pq;                                   // cg_solvers.f90:136

        // ---- B: r [, z], dpr -----------------------------------------------------
        // (x = x + alpha p is carried out in phase C, where p is read anyway:
        //  same arithmetic, one pass over p less)
        double dsum = 0.0;
        for (int64_t base = blockIdx.x * (int64_t)kThreads + tid; base < a.n; base += stride * 4) {
            double ri[4], qi[4], di[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t i = base + u * stride;
                if (i < a.n) {
                    ri[u] = a.r[i]; qi[u] = a.q[i];
                    if (PC) di[u] = a.idiag[i];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t i = base + u * stride;
                if (i < a.n) {
                    const double rn = sub(ri[u], mul(alpha, qi[u]));   // :138
                    a.r[i] = rn;
                    double zn = rn;
                    if (PC) { zn = mul(di[u], rn); a.z[i] = zn; }      // :181 (jacobi_solve)
                    dsum = add(dsum, mul(rn, zn));                     // :140 / :183
                }
            }
        }
        clk.stamp(3);
        const double dpr = grid_allreduce(dsum, a, s, pbuf, red_seq, sm_red, &s_bcast, clk, 4);
        const double beta = dpr / rr;                                   // :141
        it++;
        stop = !(sqrt(dpr) > tol);                                      // :133
        if (!stop && cap >= 0 && it0 + it >= cap) { stop = true; capped = true; }
        const bool pause = it >= a.max_iters;

        // ---- C: x = x + alpha p ; p = r + beta p ---------------------------------
        for (int64_t base = blockIdx.x * (int64_t)kThreads + tid; base < a.n; base += stride * 3) {
            double v0[3], p0[3], x0[3];
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int64_t i = base + u * stride;
                if (i < a.n) { v0[u] = rz[i]; p0[u] = a.p[i]; x0[u] = a.x[i]; }
            }
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int64_t i = base + u * stride;
                if (i < a.n) {
                    a.x[i] = add(x0[u], mul(alpha, p0[u]));            // :137
                    a.p[i] = add(v0[u], mul(beta, p0[u]));             // :142
                }
            }
        }
        rr = dpr;                                                       // :143
        clk.stamp(6);
        grid_barrier(s);
        clk.stamp(7);
        if (stop || pause) break;
    }
    clk.flush(a.phase_dbg, clk_which, it);

    // the tile primed for a pass that will not run must land before we exit
    if (pipe.primed && blockIdx.x < (unsigned)a.A.ntiles) {
        const int4 d0 = load_desc(a.A.tiles + blockIdx.x);
        if (tile_staged(d0)) mbar_wait(&mbar[pipe.sidx & 1u], (pipe.sidx >> 1) & 1u);
    }
    if (blockIdx.x == 0 && tid == 0) {
        st->iters = it0 + it;
        st->rr[0] = rr;
        st->rr[1] = rr;
        st->final_res2 = rr;
        st->done[0] = stop ? 1 : 0;
        st->done[1] = stop ? 1 : 0;
        if (capped) st->capped = 1;
        if (a.nranks > 1) a.red->red_seq = red_seq;
        if (HALO && a.A.sync.win != nullptr)
            *reinterpret_cast<volatile unsigned long long *>(&a.A.sync.win->halo_seq) = hseq;
    }
}

// ---------------------------------------------------------------------------
// EXPERIMENTAL, opt-in (SIGB_CG_SINGLE_REDUCE=1; not the default path, not yet
// run on a GPU): the Chronopoulos-Gear arrangement of the same CG iteration
// with ONE reduction per iteration instead of two.
//
//   given r, w = A r, gamma = r.r, delta = r.w      (one all-reduce of 2 values)
//   beta = gamma / gamma_old ; alpha = gamma / (delta - beta gamma / alpha_old)
//   p = r + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s
//
// In exact arithmetic p, x, r are those of cg_solve (s = A p by linearity);
// the stopping quantity is the same r.r tested at the same iteration.  It is
// NOT the reference's statement order, so results agree with the reference to
// rounding only (CPU emulation on the 2-D Poisson problems: identical
// iteration counts, solutions within 1e-14 relative) -- which is why it stays
// opt-in.  What it buys on a sharded operator: two grid barriers and one
// cross-GPU all-reduce per iteration instead of three and two, for 8 n more
// bytes of vector traffic (72 n instead of 64 n).
// Work vectors: p, s (the q slot), r, w (the z slot); unpreconditioned only.
// ---------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void grid_allreduce_n(double (&v)[NV], const CgPersistArgs &a, Sync &s, int &pbuf,
                                                 unsigned long long &red_seq, double (*sm)[kThreads / 32],
                                                 double *s_bcast, PhaseClock &clk, int slot0)
{
    static_assert(NV <= kRedVals, "the all-reduce inbox holds kRedVals values per slot");
    block_tree<NV>(v, sm);
    double *part = a.partials + (size_t)pbuf * NV * gridDim.x;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < NV; d++) part[(size_t)d * gridDim.x + blockIdx.x] = v[d];
    }
    grid_barrier(s);
    clk.stamp(slot0);
    double t[NV];
#pragma unroll
    for (int d = 0; d < NV; d++) t[d] = 0.0;
    for (unsigned j = threadIdx.x; j < gridDim.x; j += kThreads) {
#pragma unroll
        for (int d = 0; d < NV; d++) t[d] = add(t[d], __ldcg(part + (size_t)d * gridDim.x + j));
    }
    block_tree<NV>(t, sm);
    pbuf ^= 1;
    if (a.nranks > 1) {
        red_seq++;
        const int slot = (int)(red_seq & (kRedSlots - 1));
        const unsigned flag = (unsigned)red_seq;
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x < 32) {
#pragma unroll
            for (int d = 0; d < NV; d++) {
                const double local = __shfl_sync(0xffffffffu, t[d], 0);
                if ((int)threadIdx.x < a.nranks) {
                    RedEntry *e = a.peer_red[threadIdx.x]->red[slot][a.me];
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(local);
                    st_word(&e[d].lo, (unsigned)bits, flag);
                    st_word(&e[d].hi, (unsigned)(bits >> 32), flag);
                }
            }
        }
        if (threadIdx.x == 0) {
#pragma unroll
            for (int d = 0; d < NV; d++) {
                double g = 0.0;
                for (int q = 0; q < a.nranks; q++) {
                    const RedEntry *e = &a.red->red[slot][q][d];
                    uint2 lo, hi;
                    unsigned spins = 0;
                    do { lo = ld_word(&e->lo); } while (lo.y != flag && ++spins < kSpinLimit);
                    do { hi = ld_word(&e->hi); } while (hi.y != flag && ++spins < kSpinLimit);
                    g = add(g, __longlong_as_double((long long)(((unsigned long long)hi.x << 32) | lo.x)));
                }
                s_bcast[d] = g;
            }
        }
    } else if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < NV; d++) s_bcast[d] = t[d];
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < NV; d++) v[d] = s_bcast[d];
    __syncthreads();
    clk.stamp(slot0 + 1);
}

template <bool HALO>
__global__ void __launch_bounds__(kThreads, SIGB_PERSIST_MINBLOCKS)
cg_single_reduce_kernel(const CgPersistArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ double sm_red[2][kThreads / 32];
    __shared__ double s_bcast[2];

    KState *st = a.st;
    if (st->done[0]) return;

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // a.x = x, a.p = p, a.q = s, a.r = r, a.z = w ; a.A: x1 = r - 1, y = w, u = r
    double gamma = st->rr[0];
    const double tol = st->tol;
    const long long cap = st->cap, it0 = st->iters;
    long long it = 0;
    bool have = st->itc[0] != 0;        // resumed launch: w and delta of the last pass are still valid
    double delta = st->st, alpha_old = st->alpha[0], gamma_old = st->rho[0];
    bool first_of_solve = (it0 == 0 && !have);   // gamma comes from the host-side initial r.r
    double gpart = 0.0;                 // this thread's share of r.r, accumulated where r is updated
    Sync s{a.bar, 0ull, nullptr};
    int pbuf = 0;
    unsigned long long red_seq = a.nranks > 1 ? a.red->red_seq : 0ull;
    unsigned long long hseq = 0;
    if (HALO && a.A.sync.win != nullptr)
        hseq = *reinterpret_cast<volatile unsigned long long *>(&a.A.sync.win->halo_seq);
    TilePipe pipe;
    bool stop = false, capped = false;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    PhaseClock clk;   // slots used here: 0 spmv, 1 barrier, 2 cross-GPU part, 6 vector phase, 7 closing barrier
    const int clk_which = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x / 2 ? 1 : (blockIdx.x == gridDim.x - 1 ? 2 : -1));
    clk.start(tid == 0 && clk_which >= 0);

    for (;;) {
        if (!have) {
            // ---- A: w = A r, partial r.w ; all-reduce (r.r, r.w) ---------------
            double acc[1] = {0.0};
            hseq++;
            spmv_phase<MODE_SET, 1, HALO, false>(a.A, smem, mbar, pipe, acc, hseq, true);
            clk.stamp(0);
            double v[2] = {gpart, acc[0]};
            grid_allreduce_n<2>(v, a, s, pbuf, red_seq, sm_red, s_bcast, clk, 1);
            if (HALO && a.A.sync.win != nullptr && blockIdx.x == gridDim.x - 1 && tid < kMaxRanks &&
                (a.A.sync.src_mask & (1u << tid)))
                *reinterpret_cast<volatile unsigned long long *>(&a.A.sync.peer[tid]->ack[a.A.sync.me]) = hseq;
            delta = v[1];
            if (!first_of_solve) gamma = v[0];
        }
        have = false;
        first_of_solve = false;
        stop = !(sqrt(gamma) > tol);                                    // cg_solvers.f90:133
        if (!stop && cap >= 0 && it0 + it >= cap) { stop = true; capped = true; }
        if (stop || it >= a.max_iters) break;

        double alpha, beta;
        if (it0 + it == 0) {
            beta = 0.0;
            alpha = gamma / delta;
        } else {
            beta = gamma / gamma_old;
            alpha = gamma / sub(delta, mul(beta, gamma) / alpha_old);
        }

        // ---- V: p = r + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; partial r.r
        gpart = 0.0;
        for (int64_t base = blockIdx.x * (int64_t)kThreads + tid; base < a.n; base += stride * 2) {
            double ri[2], wi[2], pi[2], si[2], xi[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int64_t i = base + u * stride;
                if (i < a.n) { ri[u] = a.r[i]; wi[u] = a.z[i]; pi[u] = a.p[i]; si[u] = a.q[i]; xi[u] = a.x[i]; }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int64_t i = base + u * stride;
                if (i < a.n) {
                    const double pn = add(ri[u], mul(beta, pi[u]));
                    const double sn = add(wi[u], mul(beta, si[u]));
                    const double rn = sub(ri[u], mul(alpha, sn));
                    a.p[i] = pn;
                    a.q[i] = sn;
                    a.x[i] = add(xi[u], mul(alpha, pn));
                    a.r[i] = rn;
                    gpart = add(gpart, mul(rn, rn));
                }
            }
        }
        gamma_old = gamma;
        alpha_old = alpha;
        it++;
        clk.stamp(6);
        grid_barrier(s);   // r is complete before the next pass gathers (and pushes) it
        clk.stamp(7);
    }
    clk.flush(a.phase_dbg, clk_which, it);

    if (pipe.primed && blockIdx.x < (unsigned)a.A.ntiles) {
        const int4 d0 = load_desc(a.A.tiles + blockIdx.x);
        if (tile_staged(d0)) mbar_wait(&mbar[pipe.sidx & 1u], (pipe.sidx >> 1) & 1u);
    }
    if (blockIdx.x == 0 && tid == 0) {
        st->iters = it0 + it;
        st->rr[0] = gamma;
        st->rr[1] = gamma;
        st->final_res2 = gamma;
        st->done[0] = stop ? 1 : 0;
        st->done[1] = stop ? 1 : 0;
        if (capped) st->capped = 1;
        st->st = delta;                 // state of a paused solve: the pass that was reduced but not applied
        st->alpha[0] = alpha_old;
        st->rho[0] = gamma_old;
        st->itc[0] = stop ? 0 : 1;
        if (a.nranks > 1) a.red->red_seq = red_seq;
        if (HALO && a.A.sync.win != nullptr)
            *reinterpret_cast<volatile unsigned long long *>(&a.A.sync.win->halo_seq) = hseq;
    }
}

// Experiment knob (SIGB_CG_PERSIST_CTAS_PER_SM = 1..4, default: all that fit): fewer resident
// CTAs make the grid barriers and the partial-sum passes cheaper and the SpMV phase slower.
static int cap_persistent_grid(int grid)
{
    static const int per_sm = env_int("SIGB_CG_PERSIST_CTAS_PER_SM", 0);
    if (per_sm > 0 && per_sm * ctx().num_sms < grid) grid = per_sm * ctx().num_sms;
    return grid;
}

template <bool HALO>
int launch_single_reduce(const CgPersistArgs &a, cudaStream_t st)
{
    const size_t smem = 2 * (size_t)kStageBytes;
    int grid = 0;
    SIGB_CHECK((occupancy_grid<cg_single_reduce_kernel<HALO>>(smem, &grid)));
    grid = cap_persistent_grid(grid);
    CgPersistArgs b = a;
    if (HALO && b.A.sync.win != nullptr) {
        int pc = (b.A.sync.total_send + 2 * kThreads - 1) / (2 * kThreads);
        b.A.sync.push_ctas = b.A.sync.total_send > 0 ? std::max(1, std::min(pc, grid)) : 0;
        b.A.sync.push_first = halo_push_first(grid, b.A.sync.push_ctas);
    }
    void *params[] = {(void *)&b};
    SIGB_CUDA(cudaLaunchCooperativeKernel((const void *)cg_single_reduce_kernel<HALO>, dim3(grid), dim3(kThreads),
                                          params, smem, st));
    count_launch();
    return SIGB_OK;
}

template <bool HALO, bool PC, bool RD, bool LL = false>
int launch_persistent(const CgPersistArgs &a, cudaStream_t st)
{
    const size_t smem = 2 * (size_t)kStageBytes;
    int grid = 0;
    SIGB_CHECK((occupancy_grid<cg_persistent_kernel<HALO, PC, RD, LL>>(smem, &grid)));
    grid = cap_persistent_grid(grid);
    CgPersistArgs b = a;
    if (HALO && b.A.sync.win != nullptr) {
        int pc = (b.A.sync.total_send + 2 * kThreads - 1) / (2 * kThreads);
        b.A.sync.push_ctas = b.A.sync.total_send > 0 ? std::max(1, std::min(pc, grid)) : 0;
        b.A.sync.push_first = halo_push_first(grid, b.A.sync.push_ctas);
    }
    void *params[] = {(void *)&b};
    SIGB_CUDA(cudaLaunchCooperativeKernel((const void *)cg_persistent_kernel<HALO, PC, RD, LL>, dim3(grid), dim3(kThreads),
                                          params, smem, st));
    count_launch();
    return SIGB_OK;
}

}  // namespace

// SIGB_PHASE_TIMERS builds: one device buffer per process, (kPhaseSlots + 1) counters for each
// of the kPhaseCtas observed CTAs (the last one counts iterations); null in the product build.
static unsigned long long *phase_dbg_buffer()
{
#ifdef SIGB_PHASE_TIMERS
    static unsigned long long *buf = nullptr;
    if (!buf) {
        if (cudaMalloc((void **)&buf, sizeof(unsigned long long) * kPhaseCtas * (kPhaseSlots + 1)) != cudaSuccess)
            return nullptr;
        cudaMemset(buf, 0, sizeof(unsigned long long) * kPhaseCtas * (kPhaseSlots + 1));
    }
    return buf;
#else
    return nullptr;
#endif
}

// Fill the SpMV argument block the way launch_csr_spmv does (kernels_spmv.cu).
int fill_csr_args(const CsrView &V, const double *val, const double *x, double *y, const DotSpec &dot,
                  CsrKernelArgs *out);

int cg_persistent_run(sigb_solver_t s, const CsrView &V, const double *val, const DotSpec &halo, double *x,
                      double *p, double *q, double *r, double *z, const double *idiag, int64_t n, int grid_hint,
                      const PersistComm &pcomm, long long max_iters)
{
    (void)grid_hint;
    CgPersistArgs a;
    DotSpec d = halo;
    d.ndot = 1;
    d.u = p;
    SIGB_CHECK(fill_csr_args(V, val, p, q, d, &a.A));
    a.x = x; a.p = p; a.q = q; a.r = r; a.z = z;
    a.idiag = idiag;
    a.n = n;
    a.st = s->state;
    a.bar = s->bar;
    a.partials = s->pers_partials;
    a.max_iters = max_iters;
    a.red = (RedWin *)pcomm.red;
    for (int k = 0; k < kMaxRanks; k++) a.peer_red[k] = (RedWin *)pcomm.peer_red[k];
    a.me = pcomm.me;
    a.nranks = pcomm.nranks;
    a.phase_dbg = phase_dbg_buffer();
    cudaStream_t st = ctx().stream;
    SIGB_CUDA(cudaMemsetAsync(s->bar, 0, 2 * sizeof(unsigned long long), st));
    const bool halo_on = halo.sync != nullptr;
    if (halo_on && halo.halo_ll) {   // EXPERIMENTAL fence-free halo (SIGB_HALO_LL), see spmv_device.cuh
        if (spmv_rowdirect(V))
            return idiag ? launch_persistent<true, true, true, true>(a, st) : launch_persistent<true, false, true, true>(a, st);
        return idiag ? launch_persistent<true, true, false, true>(a, st) : launch_persistent<true, false, false, true>(a, st);
    }
    if (spmv_rowdirect(V)) {   // EXPERIMENTAL (SIGB_SPMV_ROWDIRECT), see spmv_device.cuh
        if (halo_on) return idiag ? launch_persistent<true, true, true>(a, st) : launch_persistent<true, false, true>(a, st);
        return idiag ? launch_persistent<false, true, true>(a, st) : launch_persistent<false, false, true>(a, st);
    }
    if (halo_on) return idiag ? launch_persistent<true, true, false>(a, st) : launch_persistent<true, false, false>(a, st);
    return idiag ? launch_persistent<false, true, false>(a, st) : launch_persistent<false, false, false>(a, st);
}

// EXPERIMENTAL single-reduction arrangement (see cg_single_reduce_kernel): x, p, r as in
// cg_persistent_run, s_vec = the q slot, w = the z slot.  The kernel leaves the loop with
// the last reduced pass stored in the KState, so a paused launch resumes without redoing it.
int cg_single_reduce_run(sigb_solver_t s, const CsrView &V, const double *val, const DotSpec &halo, double *x,
                         double *p, double *s_vec, double *r, double *w, int64_t n, const PersistComm &pcomm,
                         long long max_iters)
{
    CgPersistArgs a;
    DotSpec d = halo;
    d.ndot = 1;
    d.u = r;
    SIGB_CHECK(fill_csr_args(V, val, r, w, d, &a.A));
    a.x = x; a.p = p; a.q = s_vec; a.r = r; a.z = w;
    a.idiag = nullptr;
    a.n = n;
    a.st = s->state;
    a.bar = s->bar;
    a.partials = s->pers_partials;
    a.max_iters = max_iters;
    a.red = (RedWin *)pcomm.red;
    for (int k = 0; k < kMaxRanks; k++) a.peer_red[k] = (RedWin *)pcomm.peer_red[k];
    a.me = pcomm.me;
    a.nranks = pcomm.nranks;
    a.phase_dbg = phase_dbg_buffer();
    cudaStream_t st = ctx().stream;
    SIGB_CUDA(cudaMemsetAsync(s->bar, 0, 2 * sizeof(unsigned long long), st));
    return halo.sync != nullptr ? launch_single_reduce<true>(a, st) : launch_single_reduce<false>(a, st);
}

}  // namespace sigb

extern "C" {

// Diagnostic: cycles per phase of the persistent CG kernel accumulated since the last call
// (then reset), for the first / middle / last CTA: out[cta * 9 + k], k = 0..7 the phases
// listed at PhaseClock, k = 8 the iterations counted.  *supported = 0 (and zeros) unless the
// library was built with -DSIGB_PHASE_TIMERS.
int sigb_debug_cg_phase_cycles(unsigned long long *out, int *supported)
{
    using namespace sigb;
    SIGB_REQUIRE(out && supported, SIGB_ERR_ARG, "sigb_debug_cg_phase_cycles: bad argument");
    for (int k = 0; k < kPhaseCtas * (kPhaseSlots + 1); k++) out[k] = 0ull;
    *supported = 0;
#ifdef SIGB_PHASE_TIMERS
    unsigned long long *buf = phase_dbg_buffer();
    SIGB_REQUIRE(buf, SIGB_ERR_CUDA, "sigb_debug_cg_phase_cycles: no buffer");
    SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
    SIGB_CUDA(cudaMemcpy(out, buf, sizeof(unsigned long long) * kPhaseCtas * (kPhaseSlots + 1), cudaMemcpyDeviceToHost));
    SIGB_CUDA(cudaMemset(buf, 0, sizeof(unsigned long long) * kPhaseCtas * (kPhaseSlots + 1)));
    *supported = 1;
#endif
    return SIGB_OK;
}

}  // extern "C"

