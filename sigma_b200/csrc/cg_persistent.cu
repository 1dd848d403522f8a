// cg_persistent.cu -- the whole CG / Jacobi-PCG solve of
//   cg_solve      src/solver/cg_solvers.f90:116-150
//   cg_solve_pc   src/solver/cg_solvers.f90:155-194  (pc = jacobi_solver)
// as ONE persistent cooperative kernel, initial residual included.
//
// Why: on a sharded operator an iteration is ~50 us of memory traffic per GPU;
// three kernel boundaries plus two dependent all-reduce launches cost as much
// again, and the five launches that set a solve up cost several iterations.
// Here the CTAs stay resident (one cooperative wave, 3 per SM); the phases of an
// iteration are separated by grid barriers, the dot products are completed
// inside the barrier (every CTA adds the CTA partials in a fixed order; across
// GPUs the last CTA stores the local sum into the peers' inboxes with fence-free
// 8-byte payload+flag words over NVLink and every CTA polls its own inbox), the
// halo push / wait / acknowledge run inside the SpMV phase (communication CTAs,
// spmv_device.cuh), and the first matrix tile of the next SpMV is already in
// flight (TMA) while the vector phases run.  Recurrence, statement order,
// rounding (no FMA) and the stopping rule are those of the reference; all ranks
// compute bit-identical scalars and therefore leave the loop at the same
// iteration.
//
//   start:      0  q = A x                          (spmv_phase; x is the initial guess)
//               -- barrier --
//                  r = b - q ; [z = idiag r] ; p = r|z ; partial r.r | r.z      :128-131 / :167-172
//               -- barrier + all-reduce --          res2 ; first loop test      :133
//   iteration:  A  q = A p, partial p.q             (spmv_phase, tiles round-robin)
//               -- barrier + all-reduce --          alpha = res2 / (p.q)
//               B  r -= alpha q ; [z = idiag r] ; partial r.r | r.z
//               -- barrier + all-reduce --          beta = dpr / res2 ; stop test
//               C  x += alpha p ; p = (r | z) + beta p     (p is read once for both)
//               -- barrier --                       (p complete before the next gathers)
// Per iteration this moves 12 nnz + 84 n bytes (the reference's statement order
// costs 92 n: it reads p for the x update and again for the p update).
//
// 3 CTAs per SM: 4 is what shared memory allows, but that caps the kernel at 64 registers, which it
// exceeds (ptxas spilled 36-124 bytes per thread, most of it in the row-sharded instantiations);
// with 80 registers nothing spills and the grid barriers have a quarter fewer participants.
// Round-2 A/B, 2.1 M rows per GPU: 55.6 -> 50.5 us per iteration on one GPU, 67.1 -> 57.1 us on two
// (profiles/r2_visit_a_1gpu_summary.txt, r2_visit_b_2gpu_summary.txt).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "krylov.cuh"
#include "solvers.h"
#include "spmv_device.cuh"

namespace sigb {

#ifndef SIGB_PERSIST_MINBLOCKS
#define SIGB_PERSIST_MINBLOCKS 3
#endif

constexpr int kPhaseSlots = 8;   // phases timed by SIGB_PHASE_TIMERS builds (see PhaseClock)
constexpr int kPhaseCtas = 3;    // first compute CTA, middle, last CTA

struct CgPersistArgs {
    CsrKernelArgs A;            // matrix, tile table, halo sync, fault block (vectors are passed per pass)
    double *x, *p, *q, *r, *z;
    const double *b;            // right-hand side (initial residual)
    const double *idiag;        // PC only
    int64_t n;
    KState *st;
    unsigned long long *bar;    // grid barrier counter, zeroed before the launch
    double *partials;           // 2 x gridDim CTA partial sums
    long long max_iters;        // iterations this launch may run before handing back to the host
    RedWin *red;                // all-reduce inbox (nranks > 1)
    RedWin *peer_red[kMaxRanks];
    int me, nranks;
    unsigned long long *phase_dbg;  // SIGB_PHASE_TIMERS builds: kPhaseCtas x (kPhaseSlots + 1) cycle counters
};

namespace {

struct Sync {
    unsigned long long *bar;
    unsigned long long epoch;   // barriers passed (uniform across the grid)
    FaultBlock *fault;
    int *s_fault;               // shared: a wait of this CTA was abandoned
};

// Diagnostic build only (make VARIANT=_timers DEFS=-DSIGB_PHASE_TIMERS, loaded with
// SIGB_LIB_VARIANT=_timers): where an iteration of the persistent kernel spends its time.
// Three CTAs accumulate SM cycles per phase in registers and add them to phase_dbg on exit;
// the product build compiles all of it away.
//   0 spmv   1 barrier of reduction 1   2 sum + cross-GPU part of reduction 1   3 phase B
//   4 barrier of reduction 2   5 sum + cross-GPU part of reduction 2   6 phase C   7 closing barrier
struct PhaseClock {
#ifdef SIGB_PHASE_TIMERS
    long long last = 0;
    unsigned long long acc[kPhaseSlots] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool on = false;
    __device__ __forceinline__ void start(bool enable) { on = enable; if (on) last = clock64(); }
    __device__ __forceinline__ void restart() { if (on) last = clock64(); }
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
    __device__ __forceinline__ void restart() {}
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
            const unsigned long long want = s.epoch;
            const unsigned long long *flag = s.bar + 1;
            if (!spin_wait([&] { return ld_acquire_gpu(flag) >= want; }, s.fault, FAULT_GRID_BARRIER)) *s.s_fault = 1;
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
        // published by the LAST CTA (the first ones are the communication CTAs of the SpMV phase)
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x < 32) {
            const double local = __shfl_sync(0xffffffffu, t[0], 0);
            if ((int)threadIdx.x < a.nranks) red_entry_store(&a.peer_red[threadIdx.x]->red[slot][a.me][0], local, flag);
        }
        if (threadIdx.x < 32) {
            const double g = warp_rank_sum(a.red, slot, 0, a.nranks, flag, s.fault);
            if (threadIdx.x == 0) *s_bcast = g;
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

// The SpMV pass as a function of its own (not inlined into the loop below): its ~60 live registers
// (gathered values, tile descriptors, the batch of the ordered row sum) then do not compete with the
// unrolled vector phases for one allocation -- inlined, ptxas spilled 60-170 bytes per thread at the
// kernel's 80 registers.  The kernel arguments are __grid_constant__, so the callee reads them in
// place (no local copy of the argument block).
#ifndef SIGB_SPMV_NOINLINE
#define SIGB_SPMV_NOINLINE 0
#endif
#if SIGB_SPMV_NOINLINE
#define SIGB_PASS_ATTR __noinline__
#else
#define SIGB_PASS_ATTR __forceinline__
#endif
template <bool HALO>
__device__ SIGB_PASS_ATTR void spmv_pass(const CgPersistArgs &a, const double *xin, unsigned char *smem, uint64_t *mbar,
                                       TilePipe &pipe, double *acc, unsigned long long hseq)
{
    const SpmvVecs v{xin - 1, a.q, xin};
    spmv_phase<MODE_SET, 1, HALO, false>(a.A, v, smem, mbar, pipe, acc, hseq, true);
}

template <bool HALO, bool PC>
__global__ void __launch_bounds__(kThreads, SIGB_PERSIST_MINBLOCKS)
cg_persistent_kernel(const __grid_constant__ CgPersistArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ double sm_red[1][kThreads / 32];
    __shared__ double s_bcast;
    __shared__ int s_fault;

    KState *st = a.st;
    const bool resumed = st->started != 0;
    if (resumed && st->done[0]) return;   // nothing left to do (uniform)

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
        s_fault = 0;
    }
    __syncthreads();

    double rr = resumed ? st->rr[0] : 0.0;
    const double tol = st->tol;
    const long long cap = st->cap, it0 = resumed ? st->iters : 0;
    long long it = 0;
    Sync s{a.bar, 0ull, a.A.fault, &s_fault};
    int pbuf = 0;
    unsigned long long red_seq = a.nranks > 1 ? a.red->red_seq : 0ull;
    unsigned long long hseq = 0;
    const bool halo_on = HALO && a.A.sync.win != nullptr;
    if (halo_on) hseq = *reinterpret_cast<volatile unsigned long long *>(&a.A.sync.win->halo_seq);
    TilePipe pipe;
    bool stop = false, capped = false;
    bool first = !resumed;      // the pass that forms the initial residual
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    const double *rz = PC ? a.z : a.r;
    PhaseClock clk;
    const unsigned first_compute = halo_on ? (unsigned)a.A.sync.push_ctas : 0u;
    const int clk_which = blockIdx.x == first_compute ? 0 : (blockIdx.x == gridDim.x / 2 ? 1 : (blockIdx.x == gridDim.x - 1 ? 2 : -1));
    clk.start(tid == 0 && clk_which >= 0);

    for (;;) {
        // ---- A: q = A p, p.q   (first pass: q = A x, the dot is not used) ----------------
        double acc[1] = {0.0};
        hseq++;
        spmv_pass<HALO>(a, first ? a.x : a.p, smem, mbar, pipe, acc, hseq);
        if (first) {
            grid_barrier(s);   // q is written by the rows' owners and read element-wise below
            if (halo_on && blockIdx.x == gridDim.x - 1 && tid < kMaxRanks && (a.A.sync.src_mask & (1u << tid)))
                *reinterpret_cast<volatile unsigned long long *>(&a.A.sync.peer[tid]->ack[a.A.sync.me]) = hseq;
            // r = b - q ; [z = idiag r] ; p = r | z ; res2 = r.r | r.z      cg_solvers.f90:129-131 / :169-172
            double dsum = 0.0;
            for (int64_t i = blockIdx.x * (int64_t)kThreads + tid; i < a.n; i += stride) {
                const double ri = sub(a.b[i], a.q[i]);
                a.r[i] = ri;
                double zi = ri;
                if (PC) { zi = mul(a.idiag[i], ri); a.z[i] = zi; }
                a.p[i] = zi;
                dsum = add(dsum, mul(ri, zi));
            }
            rr = grid_allreduce(dsum, a, s, pbuf, red_seq, sm_red, &s_bcast, clk, 1);   // (p is complete behind its barrier)
            stop = !(sqrt(rr) > tol);                                   // :133 before the first pass
            if (!stop && cap == 0) { stop = true; capped = true; }
            if (s_fault) stop = true;
            first = false;
            clk.restart();
            if (stop) break;
            continue;
        }
        clk.stamp(0);
        const double pq = grid_allreduce(acc[0], a, s, pbuf, red_seq, sm_red, &s_bcast, clk, 1);
        if (halo_on && blockIdx.x == gridDim.x - 1 && tid < kMaxRanks && (a.A.sync.src_mask & (1u << tid))) {
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
            for (int u = 0; u < 4; u++) {                               // (every slot assigned: the arrays stay in registers)
                const int64_t i = base + u * stride;
                const bool in = i < a.n;
                ri[u] = in ? a.r[i] : 0.0;
                qi[u] = in ? a.q[i] : 0.0;
                di[u] = (PC && in) ? a.idiag[i] : 0.0;
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
                const bool in = i < a.n;
                v0[u] = in ? rz[i] : 0.0;
                p0[u] = in ? a.p[i] : 0.0;
                x0[u] = in ? a.x[i] : 0.0;
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
        if (s_fault) stop = true;
        if (stop || pause) break;
    }
    clk.flush(a.phase_dbg, clk_which, it);

    drain_primed(a.A, mbar, pipe, HALO);
    if (blockIdx.x == 0 && tid == 0) {
        st->iters = it0 + it;
        st->rr[0] = rr;
        st->rr[1] = rr;
        st->final_res2 = rr;
        st->done[0] = stop ? 1 : 0;
        st->done[1] = stop ? 1 : 0;
        if (capped) st->capped = 1;
        st->started = 1;
        if (a.nranks > 1) a.red->red_seq = red_seq;
        if (halo_on) *reinterpret_cast<volatile unsigned long long *>(&a.A.sync.win->halo_seq) = hseq;
    }
}

template <bool HALO, bool PC>
int launch_persistent(const CgPersistArgs &a, cudaStream_t st)
{
    const size_t smem = 2 * (size_t)kStageBytes;
    int grid = 0;
    SIGB_CHECK((occupancy_grid<cg_persistent_kernel<HALO, PC>>(smem, &grid)));
    void *params[] = {(void *)&a};
    static const bool verbose = env_int("SIGB_VERBOSE", 0) == 1;
    if (verbose) fprintf(stderr, "sigma_b200: cg_persistent_kernel<%d,%d> grid %d x %d threads, %zu B dynamic shared memory\n",
                         (int)HALO, (int)PC, grid, kThreads, smem);
    SIGB_CUDA(cudaLaunchCooperativeKernel((const void *)cg_persistent_kernel<HALO, PC>, dim3(grid), dim3(kThreads),
                                          params, smem, st));
    count_launch();
    return SIGB_OK;
}

}  // namespace

int persistent_grid_ctas() { return SIGB_PERSIST_MINBLOCKS * ctx().num_sms; }

// SIGB_PHASE_TIMERS builds: one device buffer per process, (kPhaseSlots + 1) counters for each
// of the kPhaseCtas observed CTAs (the last one counts iterations); null in the product build.
static unsigned long long *phase_dbg_buffer()
{
#ifdef SIGB_PHASE_TIMERS
    static thread_local unsigned long long *buf = nullptr;
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

// One launch of the persistent kernel: a fresh solve (KState::started == 0: the kernel forms the
// initial residual from x and b itself) or the continuation of a paused one.
int cg_persistent_run(sigb_solver_t s, const CsrView &V, const double *val, const DotSpec &halo, double *x,
                      const double *b, double *p, double *q, double *r, double *z, const double *idiag, int64_t n,
                      const PersistComm &pcomm, long long max_iters)
{
    CgPersistArgs a;
    DotSpec d = halo;
    d.ndot = 1;
    d.u = p;
    SIGB_CHECK(fill_csr_args(V, val, p, q, d, &a.A));
    a.A.sync.push_all = 0;     // dedicated communication CTAs inside the persistent kernel
    a.x = x; a.p = p; a.q = q; a.r = r; a.z = z;
    a.b = b;
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
    if (halo.sync != nullptr) return idiag ? launch_persistent<true, true>(a, st) : launch_persistent<true, false>(a, st);
    return idiag ? launch_persistent<false, true>(a, st) : launch_persistent<false, false>(a, st);
}

}  // namespace sigb

extern "C" {

// Diagnostic: cycles per phase of the persistent CG kernel accumulated since the last call
// (then reset), for its first compute / middle / last CTA: out[cta * 9 + k], k = 0..7 the phases
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
