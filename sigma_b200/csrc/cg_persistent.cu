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

#include "krylov.cuh"
#include "solvers.h"
#include "spmv_device.cuh"

namespace sigb {

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
};

namespace {

struct Sync {
    unsigned long long *bar;
    unsigned long long epoch;   // barriers passed (uniform across the grid)
    int *abort_flag;
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
                                                 double *s_bcast)
{
    double acc[1] = {v};
    block_tree<1>(acc, sm);
    double *part = a.partials + (size_t)pbuf * gridDim.x;
    if (threadIdx.x == 0) part[blockIdx.x] = acc[0];
    grid_barrier(s);
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
    return out;
}

template <bool HALO, bool PC>
__global__ void __launch_bounds__(kThreads, 4)
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

    for (;;) {
        // ---- A: q = A p, p.q ------------------------------------------------
        double acc[1] = {0.0};
        hseq++;
        spmv_phase<MODE_SET, 1, HALO, false>(a.A, smem, mbar, pipe, acc, hseq, true);
        const double pq = grid_allreduce(acc[0], a, s, pbuf, red_seq, sm_red, &s_bcast);
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
        const double dpr = grid_allreduce(dsum, a, s, pbuf, red_seq, sm_red, &s_bcast);
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
        grid_barrier(s);
        if (stop || pause) break;
    }

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

template <bool HALO, bool PC>
int launch_persistent(const CgPersistArgs &a, cudaStream_t st)
{
    const size_t smem = 2 * (size_t)kStageBytes;
    int grid = 0;
    SIGB_CHECK((occupancy_grid<cg_persistent_kernel<HALO, PC>>(smem, &grid)));
    CgPersistArgs b = a;
    if (HALO && b.A.sync.win != nullptr) {
        int pc = (b.A.sync.total_send + 2 * kThreads - 1) / (2 * kThreads);
        b.A.sync.push_ctas = b.A.sync.total_send > 0 ? std::max(1, std::min(pc, grid)) : 0;
    }
    void *params[] = {(void *)&b};
    SIGB_CUDA(cudaLaunchCooperativeKernel((const void *)cg_persistent_kernel<HALO, PC>, dim3(grid), dim3(kThreads),
                                          params, smem, st));
    count_launch();
    return SIGB_OK;
}

}  // namespace

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
    cudaStream_t st = ctx().stream;
    SIGB_CUDA(cudaMemsetAsync(s->bar, 0, 2 * sizeof(unsigned long long), st));
    const bool halo_on = halo.sync != nullptr;
    if (halo_on) return idiag ? launch_persistent<true, true>(a, st) : launch_persistent<true, false>(a, st);
    return idiag ? launch_persistent<false, true>(a, st) : launch_persistent<false, false>(a, st);
}

}  // namespace sigb
