// ldu.cu -- ILDU(0) preconditioner on the device.
//
// Replaces the bodies of
//   sparse_ldu_setup                          src/solver/ldu_solvers.f90:95-130
//   incomplete_ldu_sparsity_pattern           :396-441   (host index work, ldu_host.cpp)
//   sparse_static_pattern_ldu_factorization   :275-387
//   ldu_solve, lower_/upper_triangular_solve  :160-176, :208-263
//
// A ~= (I + L) D (I + U) with L, U csr matrices on the strict triangles of A's
// pattern.  Both the elimination and the triangular solves are sequential in
// the reference; their only dependencies are "row i needs rows k < i that are
// lower neighbours of i" (resp. upper neighbours for the backward solve), so
// rows are grouped into levels on the host once per pattern and each level is
// one launch in which every row runs the reference's own arithmetic, entry by
// entry in stored order, with rounded products (no FMA).  Factors, diagonal
// and every solve are therefore bit-identical to the serial loops.
//
// The sweeps are latency-bound by construction: a level of the 2-D five-point stencil in natural
// ordering is one anti-diagonal of the grid (2N - 1 levels of at most N rows).  One launch per
// level (round 1) costs ~4 us per level -- 16.8 ms per application at 1024^2, slower than one host
// core.  Deep schedules therefore run as CHUNKED SWEEPS (below): one launch per sweep, every
// thread walks a contiguous chunk of rows in order and waits only for the entries it reads.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "device_utils.cuh"
#include "dist.h"
#include "ldu_sweep.h"
#include "solvers.h"
#include "spmv_device.cuh"

namespace sigb {

// device side of a SweepPlan (ldu_sweep.h)
struct SweepDev {
    bool on = false;
    int32_t n = 0, backward = 0, R = 0, sigma = 0, C = 0, trips = 0, W = 0, S_max = 0, w16_max = 0;
    int32_t nstage = 0, stage_bytes = 0, threads = 0, has_far = 0;
    int64_t total = 0, total_s = 0;
    SweepTrip *trip = nullptr;
    int32_t *valmap = nullptr;              // entry of the factor behind every slot (-1: padding)
    unsigned char *slab = nullptr;          // what the trips read, trip by trip (sweep_static_kernel)
    double *xs = nullptr;                   // the solution in trip order
    void release()
    {
        cudaFree(trip); cudaFree(valmap); cudaFree(slab); cudaFree(xs);
        *this = SweepDev();
    }
};

struct LduInfo {
    int32_t n = 0;
    int64_t nL = 0, nU = 0, ne = 0;
    // patterns (1-based, device) and the combined value array [ Lval | Uval | D ]
    int32_t *Lptr = nullptr, *Lnode = nullptr, *Uptr = nullptr, *Unode = nullptr;
    double *fac = nullptr;
    int64_t *dest = nullptr;          // source entry -> offset in fac
    int32_t *frows = nullptr, *brows = nullptr;
    std::vector<int32_t> flev, blev;  // level pointers (host): launches are driven from here
    sigb_matrix_t rows = nullptr;     // csc / ellpack sources: their row form (device copy), else null
    // chunked sweeps: solution entries in flight as payload+flag words, the sequence number of the
    // last sweep, and the chunking (0 rows per chunk = one launch per level instead)
    RedEntry *xs = nullptr;
    unsigned sf_seq = 0;
    int32_t chunk_rows = 0, nchunks = 0;
    // statically scheduled sweeps (both or neither)
    SweepDev fsw, bsw;
    double *Lval() const { return fac; }
    double *Uval() const { return fac + nL; }
    double *D() const { return fac + nL + nU; }
};

namespace {

inline int grid_for(int64_t n)
{
    int64_t g = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)ctx().num_sms * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// "Copy A into L, D, U" (:307-324)
__global__ void __launch_bounds__(kThreads)
ldu_scatter_kernel(const double *__restrict__ aval, const int64_t *__restrict__ dest, int64_t ne,
                   double *__restrict__ fac)
{
    for (int64_t e = blockIdx.x * (int64_t)kThreads + threadIdx.x; e < ne; e += (int64_t)gridDim.x * kThreads)
        fac[dest[e]] = aval[e];
}

// M%get_value(i, j) on a csr pattern: last hit wins, 0 when absent (cs_matrices.f90:709-724)
__device__ __forceinline__ double get_value(const int32_t *ptr1, const int32_t *node1, const double *val, int32_t i,
                                            int32_t j)
{
    double z = 0.0;
    for (int32_t k = ptr1[i - 1] - 1; k < ptr1[i] - 1; k++)
        if (node1[k] == j) z = val[k];
    return z;
}

// One row of the elimination (:331-381).  Reads rows k < i of U and D(k) for the lower neighbours k of i,
// finished by an earlier launch.
__device__ __forceinline__ void factor_row(int32_t i, const int32_t *__restrict__ Lptr, const int32_t *__restrict__ Lnode,
                                           double *Lval, const int32_t *__restrict__ Uptr,
                                           const int32_t *__restrict__ Unode, double *Uval, double *D)
{
    const int32_t lb = Lptr[i - 1] - 1, dl = Lptr[i] - 1 - lb;
    const int32_t ub = Uptr[i - 1] - 1, du = Uptr[i] - 1 - ub;
    for (int32_t ind1 = 0; ind1 < dl; ind1++) {
        const int32_t k = Lnode[lb + ind1];
        double Lik = Lval[lb + ind1];                                  // L%get_value(i, k)   :342
        const double Uki = get_value(Uptr, Unode, Uval, k, i);     // :343
        const double Dk = D[k - 1];
        Lik = Lik / Dk;                                                // :345-346
        Lval[lb + ind1] = Lik;
        const double LikDk = mul(Lik, Dk);
        for (int32_t ind2 = 0; ind2 < dl; ind2++) {                    // :350-358
            const int32_t j = Lnode[lb + ind2];
            if (j > k) {
                const double Ukj = get_value(Uptr, Unode, Uval, k, j);
                Lval[lb + ind2] = add(Lval[lb + ind2], -mul(LikDk, Ukj));
            }
        }
        D[i - 1] = sub(D[i - 1], mul(LikDk, Uki));                     // :361
        for (int32_t ind2 = 0; ind2 < du; ind2++) {                    // :364-368
            const int32_t j = Unode[ub + ind2];
            const double Ukj = get_value(Uptr, Unode, Uval, k, j);
            Uval[ub + ind2] = add(Uval[ub + ind2], -mul(LikDk, Ukj));
        }
    }
    const double Di = D[i - 1];
    for (int32_t ind2 = 0; ind2 < du; ind2++) Uval[ub + ind2] = Uval[ub + ind2] / Di;   // :373-377
}

// rows of one level of the elimination, one thread per row
__global__ void __launch_bounds__(kThreads)
ldu_factor_level_kernel(const int32_t *__restrict__ rows, int32_t count, const int32_t *__restrict__ Lptr,
                        const int32_t *__restrict__ Lnode, double *Lval, const int32_t *__restrict__ Uptr,
                        const int32_t *__restrict__ Unode, double *Uval, double *D)
{
    for (int32_t t = blockIdx.x * kThreads + threadIdx.x; t < count; t += gridDim.x * kThreads)
        factor_row(rows[t], Lptr, Lnode, Lval, Uptr, Unode, Uval, D);
}

// rows of one level of lower_/upper_triangular_solve (:226-235, :254-263):
// z = x(i) ; z = z - M%val(k) * x(node(k)) in stored order ; x(i) = z
__global__ void __launch_bounds__(kThreads)
tri_level_kernel(const int32_t *__restrict__ rows, int32_t count, const int32_t *__restrict__ ptr1,
                 const int32_t *__restrict__ node1, const double *__restrict__ val, double *x, const int *skip)
{
    if (skip != nullptr && *skip != 0) return;
    for (int32_t t = blockIdx.x * kThreads + threadIdx.x; t < count; t += gridDim.x * kThreads) {
        const int32_t i = rows[t];
        double z = x[i - 1];
        for (int32_t k = ptr1[i - 1] - 1; k < ptr1[i] - 1; k++) z = sub(z, mul(val[k], x[node1[k] - 1]));
        x[i - 1] = z;
    }
}

// x = b, then the rows of forward level 0 need nothing else; x = x / D between the sweeps
__global__ void __launch_bounds__(kThreads)
copy_kernel(const double *__restrict__ b, double *__restrict__ x, int64_t n, const int *skip)
{
    if (skip != nullptr && *skip != 0) return;
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        x[i] = b[i];
}
__global__ void __launch_bounds__(kThreads)
divide_kernel(double *__restrict__ x, const double *__restrict__ D, int64_t n, const int *skip)
{
    if (skip != nullptr && *skip != 0) return;
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        x[i] = x[i] / D[i];
}

// ---------------------------------------------------------------------------
// Chunked sweeps: (I + L) x = b and (I + U) x = x / D in ONE launch each.
//
// Thread t owns the contiguous rows [t B, (t + 1) B) and solves them one after the other, in
// order (the backward sweep mirrors this from the last row down).  A finished x(i) is published as
// two 8-byte words, each 32 payload bits + the 32-bit sequence number of this sweep (the words of
// the all-reduce, device_utils.cuh): a word is delivered as a unit, so a reader needs no fence and
// no second round trip.  A row waits for exactly the entries it reads, in stored order; the entry
// it has just produced itself comes from a register.
// B = the bandwidth of the factor (max |i - j| over its entries): row i of thread t then depends on
// rows of thread t - 1 at the same position or earlier, so the threads run as a systolic wavefront
// one step apart and the sweep takes ~(B + n / B) dependent steps of one L2 round trip each instead
// of one launch (or one grid barrier) per level -- for the N x N five-point stencil B = N: 2 N steps.
// Lanes never spin on their own: a warp runs one loop in which every unfinished lane polls once
// and advances as far as it can; a dependency on a lower lane resolves on a later trip.  Progress:
// all threads are resident (the grid is sized for that) and every row depends on lower rows only,
// i.e. on its own thread's past or on lower threads.  Every row still does the reference's
// arithmetic in stored order with rounded products, so the solves stay bit-identical to the serial
// loops (ldu_solvers.f90:226-235, :254-263).  Waits are bounded through spin_check.
// ---------------------------------------------------------------------------
// payload+flag words as in device_utils.cuh, but at GPU scope: these sweeps never leave the device
__device__ __forceinline__ void st_word_gpu(unsigned int *p, unsigned int payload, unsigned int flag)
{
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(flag) : "memory");
}
__device__ __forceinline__ uint2 ld_word_gpu(const unsigned int *p)
{
    uint2 r;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ bool ll_try(const RedEntry *e, unsigned seq, double *out)
{
    const uint2 lo = ld_word_gpu(&e->lo), hi = ld_word_gpu(&e->hi);
    if (lo.y != seq || hi.y != seq) return false;
    *out = __longlong_as_double((long long)(((unsigned long long)hi.x << 32) | lo.x));
    return true;
}
__device__ __forceinline__ void ll_publish(RedEntry *e, unsigned seq, double v)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    st_word_gpu(&e->lo, (unsigned)bits, seq);
    st_word_gpu(&e->hi, (unsigned)(bits >> 32), seq);
}

constexpr int kSweepThreads = 32;    // one warp per CTA: the wavefront spreads over as many SMs as possible

// BACKWARD = false: (I + M) x = src, rows ascending.  BACKWARD = true: (I + M) x = src / D, rows
// descending (the x = x / D statement of ldu_solve :169 folded into the start of the sweep: same
// division, same operands).
template <bool BACKWARD>
__global__ void __launch_bounds__(kSweepThreads)
tri_chunked_kernel(int32_t n, int32_t chunk_rows, int32_t nchunks, const int32_t *__restrict__ ptr1,
                   const int32_t *__restrict__ node1, const double *__restrict__ val, const double *src,
                   const double *__restrict__ D, double *x, RedEntry *xs, unsigned seq, FaultBlock *fault,
                   const int *skip)
{
    if (skip != nullptr && *skip != 0) return;
    const int lane = threadIdx.x;
    const int32_t t = blockIdx.x * kSweepThreads + lane;      // chunk of this thread; the warp owns chunks [t - lane, t - lane + 32)
    // rows of this chunk, 1-based, in sweep order: i, i + step, ..., last
    int32_t i = 0, last = 0;
    const int32_t step = BACKWARD ? -1 : 1;
    bool done = true;
    if (t < nchunks) {
        const int64_t lo = (int64_t)t * chunk_rows, hi = min((int64_t)n, lo + chunk_rows);   // [lo, hi) counted from the sweep's start
        if (!BACKWARD) { i = (int32_t)lo + 1; last = (int32_t)hi; }
        else { i = n - (int32_t)lo; last = n - (int32_t)hi + 1; }
        done = false;
    }
    // chunk -> lane of this warp that owns row j (or a value outside 0..31)
    auto owner_lane = [&](int32_t j) {
        const unsigned pos = (unsigned)(BACKWARD ? n - j : j - 1);          // position from the sweep's start
        return (int)(pos / (unsigned)chunk_rows) - (int)(t - lane);
    };
    // Every lane walks its own chunk: the 32 lanes of a warp touch 32 different lines of ptr / node / val /
    // src per trip, one lane or another misses L1 on every trip, and the loads of a row are a dependent
    // chain (ptr -> node -> x) -- without help a trip costs several DRAM latencies (measured: ~5600 cycles).
    // The streams are sequential per lane, so the lines a lane will need ~32 rows from now are prefetched.
    const int64_t ne_total = (int64_t)ptr1[n] - 1;
    auto prefetch = [](const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); };
    int32_t k = 0, e = 0, prev_row = 0;
    double z = 0.0, prev_z = 0.0;
    bool have_row = false;
    unsigned trips = 0;
    unsigned long long t0 = 0ull;
#ifdef SIGB_SWEEP_STATS
    unsigned n_own = 0, n_shfl = 0, n_try = 0, n_fail = 0, n_rows = 0;
    const long long c_begin = clock64();
#define SIGB_STAT(x) x
#else
#define SIGB_STAT(x)
#endif
    for (;;) {
        if (!done && !have_row) {
            k = ptr1[i - 1] - 1;
            e = ptr1[i] - 1;
            z = BACKWARD ? src[i - 1] / D[i - 1] : src[i - 1];
            have_row = true;
            const int32_t ahead = min(max(i - 1 + 32 * step, 0), n - 1);           // a row ~32 steps ahead
            const int64_t kahead = min(max((int64_t)k + 96 * step, (int64_t)0), max(ne_total - 1, (int64_t)0));
            prefetch(ptr1 + ahead);
            prefetch(src + ahead);
            if (BACKWARD) prefetch(D + ahead);
            prefetch(node1 + kahead);
            prefetch(val + kahead);
        }
        // Two convergent rounds: the next entry of every lane is looked for, in this order, in the
        // lane's own last result (a register), in the last result of the lane of this warp that owns
        // it (a shuffle: in a banded matrix the neighbouring chunk finished exactly that row one trip
        // ago -- no round trip through L2), and in the published words.  z = z - M%val(k) * x(node(k)),
        // entries strictly in stored order.
#pragma unroll
        for (int round = 0; round < 2; round++) {
            const bool want = !done && k < e;
            const int32_t j = want ? node1[k] : 0;
            const int ol = want ? owner_lane(j) : lane;
            const int src_lane = (ol >= 0 && ol < 32) ? ol : lane;
            const int32_t their_row = __shfl_sync(0xffffffffu, prev_row, src_lane);
            const double their_z = __shfl_sync(0xffffffffu, prev_z, src_lane);
            if (want) {
                double xj = 0.0;
                bool got = false;
                if (j == prev_row) { xj = prev_z; got = true; SIGB_STAT(n_own++); }
                else if (their_row == j) { xj = their_z; got = true; SIGB_STAT(n_shfl++); }
                else { got = ll_try(xs + (j - 1), seq, &xj); SIGB_STAT(n_try++); SIGB_STAT(if (!got) n_fail++); }
                if (got) {
                    z = sub(z, mul(val[k], xj));
                    k++;
                }
            }
        }
        if (!done) {
            while (k < e) {                                   // further entries: own register or published words
                const int32_t j = node1[k];
                double xj;
                if (j == prev_row) xj = prev_z;
                else if (!ll_try(xs + (j - 1), seq, &xj)) break;
                z = sub(z, mul(val[k], xj));
                k++;
            }
            if (k == e) {
                x[i - 1] = z;
                ll_publish(xs + (i - 1), seq, z);
                prev_row = i;
                prev_z = z;
                have_row = false;
                SIGB_STAT(n_rows++);
                if (i == last) done = true;
                else i += step;
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
        if ((++trips & 0xfffu) == 0u && !spin_check(fault, &t0, FAULT_LDU_SWEEP)) break;
    }
#ifdef SIGB_SWEEP_STATS
    if ((blockIdx.x == 0 || blockIdx.x == gridDim.x / 2 || blockIdx.x == gridDim.x - 1) && (lane == 0 || lane == 31))
        printf("sweep%s cta %d lane %d: trips %u rows %u own %u shfl %u try %u fail %u cycles %lld (%.0f per trip)\n",
               BACKWARD ? "B" : "F", blockIdx.x, lane, trips, n_rows, n_own, n_shfl, n_try, n_fail,
               (long long)(clock64() - c_begin), (double)(clock64() - c_begin) / (trips + 1));
#endif
}


// ---------------------------------------------------------------------------
// Statically scheduled sweeps (ldu_sweep.h has the schedule): (I + M) x = rhs as ONE CTA.
//
// Thread u of trip t owns chunk vlo(t) + u and computes its row at position t - sigma * chunk: the
// reference's z = x(i); z = z - M%val(k) * x(node(k)) in stored order, rounded products (bit-identical
// to the serial loops, ldu_solvers.f90:226-235, :254-263).  x(node(k)) comes from the shared-memory ring
// (the last W positions of every chunk; slot computed on the host) or, further back, from the
// trip-ordered solution in global memory.  Everything a trip reads -- its right-hand sides, values,
// ring slots and row lengths -- is ONE contiguous slab [rhs | val | src | cnt] in global memory, staged
// by one bulk copy (TMA) nstage - 1 trips ahead; one barrier per trip; no polling anywhere: the host has
// shown that every value a trip reads was produced by an earlier trip.
//
// The sweep is bound by INSTRUCTION ISSUE, not by memory (diagnostic build _sweepstats, visit r2t: with a
// generic row loop ~140 instructions per row, 1 140 cycles per trip in the rows, 8 cycles at the barrier,
// ~120 waiting for the stage), so rows of at most two entries without far reads -- every stencil-like
// factor -- take a straight-line body, and the producer issues one copy per trip instead of four.
// The right-hand side and the solution live in trip order (coalesced in the sweep); two tiled transposes
// convert from and to the natural order.  (Measured and dropped: bulk stores through shared memory instead of
// plain stores in front of the barrier, L2 prefetch of the slabs 8 / 24 trips ahead -- no effect or slower,
// profiles/r2_visit_s_1gpu_summary.txt, r2_visit_pf_1gpu_summary.txt.)
// ---------------------------------------------------------------------------
struct SweepArgs {
    int32_t n, backward, R, sigma, C, trips, W, S_max, w16_max, nstage, stage_bytes;
    const SweepTrip *trip;
    unsigned char *slab;     // per trip [rhs 8 w16 | val 8 S w16 | src 4 S w16 | cnt w16] at byte 9 off + 12 soff
    double *xs;              // the solution in trip order
};

__device__ __forceinline__ long long slab_offset(const SweepTrip &T) { return 9ll * T.off + 12ll * T.soff; }

// The slots of a trip, from the device-resident pattern of the factor (one CTA per trip; the statements of
// build_sweep_plan's slot loop, ldu_host.cpp -- which tests/test_ldu_sweep_plan.py replays on the CPU):
// row length and ring slot / far index of every entry into the slab, the entry's position in the factor
// into valmap.  Once per pattern.
__global__ void __launch_bounds__(kThreads)
sweep_slots_kernel(const SweepArgs a, const int32_t *__restrict__ ptr1, const int32_t *__restrict__ node1,
                   int32_t *__restrict__ valmap)
{
    const int t = blockIdx.x;
    const SweepTrip T = a.trip[t];
    unsigned char *base = a.slab + slab_offset(T);
    int32_t *d_src = reinterpret_cast<int32_t *>(base + 8ll * T.w16 * (1 + T.S));
    unsigned char *d_cnt = base + 8ll * T.w16 * (1 + T.S) + 4ll * T.S * T.w16;
    for (int u = threadIdx.x; u < T.w16; u += kThreads) {
        const long long v = T.vlo + u, p = (long long)t - (long long)a.sigma * v, q = v * a.R + p;
        int c = kSweepNoRow;
        if (u < T.w && q < a.n) {
            const int i = a.backward ? (int)(a.n - q) : (int)(q + 1);          // 1-based row at sweep position q
            const int kb = ptr1[i - 1] - 1;
            c = ptr1[i] - 1 - kb;
            for (int s = 0; s < c; s++) {
                const int j = node1[kb + s];
                const long long q2 = a.backward ? (long long)a.n - j : (long long)j - 1;
                const long long v2 = q2 / a.R, p2 = q2 % a.R, t2 = p2 + (long long)a.sigma * v2;
                const long long d = t - t2;                                    // >= 1: finished in an earlier trip
                int src;
                if (d < a.W) {
                    src = (int)((p2 & (a.W - 1)) * a.C + v2);
                } else {
                    const SweepTrip T2 = a.trip[t2];
                    src = -(int)(1 + T2.off + (v2 - T2.vlo));
                }
                d_src[s * T.w16 + u] = src;
                valmap[T.soff + (long long)s * T.w16 + u] = kb + s;
            }
        }
        for (int s = (c == kSweepNoRow ? 0 : c); s < T.S; s++) {               // padding: ring slot 0, no entry
            d_src[s * T.w16 + u] = 0;
            valmap[T.soff + (long long)s * T.w16 + u] = -1;
        }
        d_cnt[u] = (unsigned char)c;
    }
}

// the factor's values into the slabs (every factorisation); one CTA per trip
__global__ void __launch_bounds__(kThreads)
sweep_pack_kernel(const SweepTrip *__restrict__ trip, const int32_t *__restrict__ valmap, const double *__restrict__ fac,
                  unsigned char *__restrict__ slab)
{
    const SweepTrip T = trip[blockIdx.x];
    double *val = reinterpret_cast<double *>(slab + slab_offset(T) + 8ll * T.w16);
    const int nslots = T.S * T.w16;
    for (int k = threadIdx.x; k < nslots; k += kThreads) {
        const int32_t m = valmap[T.soff + k];
        val[k] = m >= 0 ? fac[m] : 0.0;
    }
}

// natural order <-> trip order.  Tile of 32 trips x 32 chunks through shared memory: for a fixed chunk
// consecutive trips are consecutive rows (coalesced in the natural order), for a fixed trip consecutive
// chunks are consecutive trip-ordered entries.  IN: natural -> the rhs part of the slabs (BACKWARD folds
// x = x / D, ldu_solve :169, in); OUT: trip-ordered solution -> natural.
template <bool BACKWARD, bool OUT>
__global__ void __launch_bounds__(256)
sweep_transpose_kernel(const SweepArgs a, const double *__restrict__ src, const double *__restrict__ D,
                       double *__restrict__ dst, const int *skip)
{
    if (skip != nullptr && *skip != 0) return;
    __shared__ double tile[32][33];
    const int t0 = blockIdx.x * 32, v0 = blockIdx.y * 32;
    // position p = t - sigma * v over the tile: nothing to do when no (t, v) has 0 <= p < R
    const long long pmin = (long long)t0 - (long long)a.sigma * (v0 + 31), pmax = (long long)t0 + 31 - (long long)a.sigma * v0;
    if (pmax < 0 || pmin >= a.R) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    auto row_of = [&](int t, int v, long long *i0) {
        if (t >= a.trips || v >= a.C) return false;
        const long long p = (long long)t - (long long)a.sigma * v;
        if (p < 0 || p >= a.R) return false;
        const long long q = (long long)v * a.R + p;
        if (q >= a.n) return false;
        *i0 = BACKWARD ? (long long)a.n - 1 - q : q;
        return true;
    };
    if (!OUT) {
        for (int j = ty; j < 32; j += 8) {
            long long i0;
            if (row_of(t0 + tx, v0 + j, &i0)) tile[j][tx] = BACKWARD ? src[i0] / D[i0] : src[i0];
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            long long i0;
            const int t = t0 + j, v = v0 + tx;
            if (row_of(t, v, &i0)) {
                const SweepTrip T = a.trip[t];
                reinterpret_cast<double *>(a.slab + slab_offset(T))[v - T.vlo] = tile[tx][j];
            }
        }
    } else {
        for (int j = ty; j < 32; j += 8) {
            long long i0;
            const int t = t0 + j, v = v0 + tx;
            if (row_of(t, v, &i0)) tile[tx][j] = src[a.trip[t].off + (v - a.trip[t].vlo)];
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            long long i0;
            if (row_of(t0 + tx, v0 + j, &i0)) dst[i0] = tile[j][tx];
        }
    }
}

// SHORT: every row has at most two entries and none is read from global memory (straight-line body);
// otherwise the general loop.
template <bool SHORT>
__global__ void __launch_bounds__(1024, 1)
sweep_static_kernel(const __grid_constant__ SweepArgs a, const int *skip)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar[4];
    if (skip != nullptr && *skip != 0) return;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    double *ring = reinterpret_cast<double *>(smem + (size_t)a.nstage * a.stage_bytes);
    if (tid == 0) {
        for (int k = 0; k < a.nstage; k++) mbar_init(&mbar[k], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const uint64_t policy = policy_evict_first();
    // Warp 0 is the PRODUCER: its lane 0 programs one bulk copy per trip (descriptor read one trip earlier)
    // and takes no rows.  With the producer's ~100 instructions in a warp that also owned rows, that warp
    // was the last at every barrier: 800 of 1 585 cycles per trip (diagnostic build, visit r2u).
    const bool producer = tid < 32;
    const int ctid = tid - 32, ncompute = nthreads - 32;
    int t_issue = 0, stage_issue = 0;
    SweepTrip T_issue = a.trip[0];
    auto issue = [&]() {
        const uint32_t bytes = (uint32_t)T_issue.w16 * (9u + 12u * (uint32_t)T_issue.S);
        mbar_expect_tx(&mbar[stage_issue], bytes);
        if (bytes) bulk_g2s(smem + (size_t)stage_issue * a.stage_bytes, a.slab + slab_offset(T_issue), bytes, &mbar[stage_issue], policy);
        t_issue++;
        stage_issue = stage_issue + 1 == a.nstage ? 0 : stage_issue + 1;
        if (t_issue < a.trips) T_issue = a.trip[t_issue];
    };
    if (tid == 0)
        while (t_issue < a.nstage - 1 && t_issue < a.trips) issue();
    const int wmask = a.W - 1;
    // compute threads: (vlo, w, w16, S) and the offset in xs of the trip computed next, read one trip earlier
    int4 d = __ldg(reinterpret_cast<const int4 *>(a.trip));
    long long off = __ldg(&a.trip[0].off);
    int stage = 0;
    uint32_t parity = 0;
#ifdef SIGB_SWEEP_STATS
    long long c_issue = 0, c_wait = 0, c_rows = 0, c_bar = 0;
    const long long c_begin = clock64();
#define SWEEP_CLK(var) const long long var = clock64()
#define SWEEP_ACC(acc, t1, t0) acc += (t1) - (t0)
#else
#define SWEEP_CLK(var)
#define SWEEP_ACC(acc, t1, t0)
#endif
    for (int t = 0; t < a.trips; t++) {
        SWEEP_CLK(k0);
        int4 d_next = d;
        long long off_next = off;
        if (producer) {
            // the stage trip t - 1 used was released by the barrier that ended it
            if (tid == 0 && t_issue < a.trips) issue();
        } else if (t + 1 < a.trips) {
            d_next = __ldg(reinterpret_cast<const int4 *>(a.trip + t + 1));
            off_next = __ldg(&a.trip[t + 1].off);
        }
        SWEEP_CLK(k1);
        if (!producer) mbar_wait(&mbar[stage], parity);
        SWEEP_CLK(k2);
        const int w16 = d.z, S = d.w;
        const unsigned char *base = smem + (size_t)stage * a.stage_bytes;
        const double *s_rhs = reinterpret_cast<const double *>(base);
        const double *s_val = s_rhs + w16;
        const int32_t *s_src = reinterpret_cast<const int32_t *>(s_val + (size_t)S * w16);
        const unsigned char *s_cnt = reinterpret_cast<const unsigned char *>(s_src + (size_t)S * w16);
        for (int u = producer ? d.y : ctid; u < d.y; u += ncompute) {
            const int c = s_cnt[u];
            double z = s_rhs[u];
            if (SHORT) {
                // padded slots carry ring index 0 and are not applied (z - 0 * x would turn -0 into +0)
                const int i0 = S > 0 ? s_src[u] : 0, i1 = S > 1 ? s_src[w16 + u] : 0;
                const double a0 = S > 0 ? s_val[u] : 0.0, a1 = S > 1 ? s_val[w16 + u] : 0.0;
                const double x0 = ring[i0], x1 = ring[i1];
                if (c == kSweepNoRow) continue;
                if (c >= 1) z = sub(z, mul(a0, x0));
                if (c >= 2) z = sub(z, mul(a1, x1));
            } else {
                if (c == kSweepNoRow) continue;
                for (int s = 0; s < c; s++) {
                    const int src = s_src[s * w16 + u];
                    const double xj = src >= 0 ? ring[src] : __ldcg(a.xs + (-(long long)src - 1));
                    z = sub(z, mul(s_val[s * w16 + u], xj));
                }
            }
            const int v = d.x + u, p = t - a.sigma * v;
            a.xs[off + u] = z;
            ring[(p & wmask) * a.C + v] = z;
        }
        SWEEP_CLK(k3);
        // (the stage is only READ through the generic proxy, so the barrier alone orders it before the
        //  bulk copy that refills it; the barrier also makes this trip's x visible to the CTA)
        __syncthreads();
        SWEEP_CLK(k4);
        d = d_next;
        off = off_next;
        if (++stage == a.nstage) { stage = 0; parity ^= 1u; }
        SWEEP_ACC(c_issue, k1, k0);
        SWEEP_ACC(c_wait, k2, k1);
        SWEEP_ACC(c_rows, k3, k2);
        SWEEP_ACC(c_bar, k4, k3);
    }
#ifdef SIGB_SWEEP_STATS
    if (tid == 0 || tid == 32 || tid == nthreads - 1)
        printf("sweep %s thread %d of %d: trips %d, cycles per trip: issue+descriptors %.0f, wait for the stage %.0f, rows %.0f, "
               "barrier %.0f, all %.0f\n", a.backward ? "B" : "F", tid, nthreads, a.trips, (double)c_issue / a.trips,
               (double)c_wait / a.trips, (double)c_rows / a.trips, (double)c_bar / a.trips,
               (double)(clock64() - c_begin) / a.trips);
#endif
}

static SweepArgs sweep_args(const SweepDev &W)
{
    SweepArgs a;
    a.n = W.n; a.backward = W.backward; a.R = W.R; a.sigma = W.sigma; a.C = W.C; a.trips = W.trips; a.W = W.W;
    a.S_max = W.S_max; a.w16_max = W.w16_max; a.nstage = W.nstage; a.stage_bytes = W.stage_bytes;
    a.trip = W.trip; a.slab = W.slab; a.xs = W.xs;
    return a;
}

// one statically scheduled sweep: src (natural order; BACKWARD: / D) -> trip order, the sweep, -> x
template <bool BACKWARD>
int launch_sweep_static(const SweepDev &W, const double *src, const double *D, double *x, const int *skip)
{
    cudaStream_t st = ctx().stream;
    const SweepArgs a = sweep_args(W);
    const dim3 tg((unsigned)((W.trips + 31) / 32), (unsigned)((W.C + 31) / 32));
    sweep_transpose_kernel<BACKWARD, false><<<tg, 256, 0, st>>>(a, src, D, nullptr, skip);
    const size_t smem = (size_t)W.nstage * W.stage_bytes + (size_t)W.W * W.C * 8;
    static thread_local bool attr_set = false;   // per host thread = per device
    if (!attr_set) {
        SIGB_CUDA(cudaFuncSetAttribute(sweep_static_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        SIGB_CUDA(cudaFuncSetAttribute(sweep_static_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        attr_set = true;
    }
    // one producer warp + compute threads, each of which takes ceil(chunks / compute threads) rows per trip
    static const int max_threads = env_int("SIGB_LDU_SWEEP_THREADS", 512);
    const int threads = 32 + std::max(32, std::min(std::min(W.threads, 992), max_threads & ~31));
    if (W.S_max <= 2 && !W.has_far) sweep_static_kernel<true><<<1, threads, smem, st>>>(a, skip);
    else sweep_static_kernel<false><<<1, threads, smem, st>>>(a, skip);
    sweep_transpose_kernel<BACKWARD, true><<<tg, 256, 0, st>>>(a, W.xs, nullptr, x, skip);
    count_launch(3);
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

// plan -> device: the trip table is uploaded, the slots are computed here from the device-resident pattern
static int upload_sweep(const SweepPlan &P, const int32_t *ptr1_dev, const int32_t *node1_dev, SweepDev &W)
{
    W = SweepDev();
    if (!P.eligible) return SIGB_OK;
    W.n = P.n; W.backward = P.backward; W.R = P.R; W.sigma = P.sigma; W.C = P.C; W.trips = P.trips; W.W = P.W;
    W.S_max = P.S_max; W.w16_max = P.w16_max; W.nstage = P.nstage; W.stage_bytes = P.stage_bytes; W.threads = P.threads;
    W.total = P.total; W.total_s = P.total_s; W.has_far = P.has_far;
    cudaStream_t st = ctx().stream;
    const size_t ts = (size_t)std::max<int64_t>(P.total_s, 1), tt = (size_t)std::max<int64_t>(P.total, 1);
    SIGB_CUDA(cudaMalloc((void **)&W.trip, sizeof(SweepTrip) * P.trip.size()));
    SIGB_CUDA(cudaMalloc((void **)&W.valmap, sizeof(int32_t) * ts));
    SIGB_CUDA(cudaMalloc((void **)&W.slab, 9 * tt + 12 * ts));
    SIGB_CUDA(cudaMalloc((void **)&W.xs, sizeof(double) * tt));
    SIGB_CUDA(cudaMemcpyAsync(W.trip, P.trip.data(), sizeof(SweepTrip) * P.trip.size(), cudaMemcpyHostToDevice, st));
    // the padding of a trip's right-hand side / solution is staged and never used: defined values all the same
    SIGB_CUDA(cudaMemsetAsync(W.slab, 0, 9 * tt + 12 * ts, st));
    SIGB_CUDA(cudaMemsetAsync(W.xs, 0, sizeof(double) * tt, st));
    if (P.trips > 0) {
        sweep_slots_kernel<<<P.trips, kThreads, 0, st>>>(sweep_args(W), ptr1_dev, node1_dev, W.valmap);
        count_launch();
        SIGB_CUDA(cudaGetLastError());
    }
    SIGB_CUDA(cudaStreamSynchronize(st));    // the plan's trip table goes out of scope in the caller
    W.on = true;
    return SIGB_OK;
}

static int pack_sweep(const SweepDev &W, const double *fac_part)
{
    if (!W.on || W.total_s == 0 || W.trips == 0) return SIGB_OK;
    sweep_pack_kernel<<<W.trips, kThreads, 0, ctx().stream>>>(W.trip, W.valmap, fac_part, W.slab);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

// scratch of the chunked sweeps; hands out the next sequence number (0 = never written)
int sweep_prepare(LduInfo *F, unsigned *seq)
{
    cudaStream_t st = ctx().stream;
    const size_t n = (size_t)(F->n > 0 ? F->n : 1);
    if (!F->xs) {
        SIGB_CUDA(cudaMalloc((void **)&F->xs, sizeof(RedEntry) * n));
        F->sf_seq = 0;
    }
    if (F->sf_seq == 0 || F->sf_seq == 0xffffffffu) {   // first use, or the counter is about to wrap
        SIGB_CUDA(cudaMemsetAsync(F->xs, 0, sizeof(RedEntry) * n, st));
        F->sf_seq = 0;
    }
    *seq = ++F->sf_seq;
    return SIGB_OK;
}

template <bool BACKWARD>
int launch_tri_chunked(LduInfo *F, const int32_t *ptr1, const int32_t *node1, const double *val, const double *src,
                       double *x, const int *skip)
{
    unsigned seq = 0;
    SIGB_CHECK(sweep_prepare(F, &seq));
    const int grid = (F->nchunks + kSweepThreads - 1) / kSweepThreads;
    tri_chunked_kernel<BACKWARD><<<grid, kSweepThreads, 0, ctx().stream>>>(F->n, F->chunk_rows, F->nchunks, ptr1, node1,
                                                                           val, src, F->D(), x, F->xs, seq,
                                                                           ctx().fault_dev, skip);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

// Chunking of the sweeps from the patterns of L and U (host index work, once per pattern): chunk
// length = bandwidth of the factors, at least n / kMaxChunks.  Chunked sweeps are used when the level
// schedule is deep (one launch per level costs ~4 us per level) and the chunking leaves at least a
// warp of chunks; otherwise -- few, wide levels, e.g. a random graph -- the level launches stay.
constexpr int kMaxChunks = 4096;   // threads of one sweep; all must be resident (32 per CTA -> 128 CTAs)
void choose_chunking(LduInfo *F, const std::vector<int32_t> &Lptr, const std::vector<int32_t> &Lnode,
                     const std::vector<int32_t> &Uptr, const std::vector<int32_t> &Unode)
{
    const int32_t n = F->n;
    int64_t bw = 1;
    for (int32_t i = 1; i <= n; i++) {
        for (int32_t k = Lptr[(size_t)i - 1] - 1; k < Lptr[(size_t)i] - 1; k++) bw = std::max<int64_t>(bw, i - Lnode[(size_t)k]);
        for (int32_t k = Uptr[(size_t)i - 1] - 1; k < Uptr[(size_t)i] - 1; k++) bw = std::max<int64_t>(bw, Unode[(size_t)k] - i);
    }
    int64_t rows = std::max<int64_t>(bw, ((int64_t)n + kMaxChunks - 1) / kMaxChunks);
    const int64_t chunks = n > 0 ? ((int64_t)n + rows - 1) / rows : 0;
    const int64_t levels = std::max(F->flev.size(), F->blev.size());
    if (levels > 64 && chunks >= 32) {
        F->chunk_rows = (int32_t)rows;
        F->nchunks = (int32_t)chunks;
    } else {
        F->chunk_rows = 0;
        F->nchunks = 0;
    }
}

void free_ldu(LduInfo *F)
{
    if (!F) return;
    F->fsw.release();
    F->bsw.release();
    cudaFree(F->xs);
    cudaFree(F->Lptr); cudaFree(F->Lnode); cudaFree(F->Uptr); cudaFree(F->Unode);
    cudaFree(F->fac); cudaFree(F->dest); cudaFree(F->frows); cudaFree(F->brows);
    if (F->rows) sigb_matrix_destroy(F->rows);
    delete F;
}

template <typename T>
int upload(T **dst, const T *src, size_t count)
{
    SIGB_CUDA(cudaMalloc((void **)dst, sizeof(T) * (count > 0 ? count : 1)));
    if (count > 0)
        SIGB_CUDA(cudaMemcpyAsync(*dst, src, sizeof(T) * count, cudaMemcpyHostToDevice, ctx().stream));
    return SIGB_OK;
}

}  // namespace

void ldu_destroy_dev(sigb_solver_t s)
{
    free_ldu(s->ldu);
    s->ldu = nullptr;
}

// pc%setup(A): the pattern and the schedules once (solver%initialized, :113-121),
// the numeric factorisation every time (:124-126)
int ldu_setup_dev(sigb_solver_t s, sigb_matrix_t A)
{
    SIGB_REQUIRE(!A->op && !A->dist, SIGB_ERR_UNSUPPORTED,
                 "ldu setup needs a stored sparse matrix (sparse_ldu_setup selects on sparse_matrix_interface, "
                 "ldu_solvers.f90:111-112)");
    cudaStream_t st = ctx().stream;
    const int32_t n = A->nrow;
    LduInfo *F = s->ldu;
    if (F && (F->n != n || F->ne != A->g->ne)) {   // another operator: start over
        free_ldu(F);
        F = s->ldu = nullptr;
    }
    // the rows of A in iteration order: a csr_matrix as stored, otherwise its csr copy
    sigb_matrix_t R = A;
    if (A->g->kind != G_CSR) {
        if (F && F->rows) { sigb_matrix_destroy(F->rows); F->rows = nullptr; }
        sigb_matrix_t rows = nullptr;
        SIGB_CHECK(sigb_matrix_copy(A, SIGB_FMT_CSR, 0, &rows));
        R = rows;
    }
    int rc = SIGB_OK;
    if (!F) {
        F = new LduInfo();
        F->n = n;
        F->ne = R->g->ne;
        const int64_t ne = F->ne;
        std::vector<int32_t> ptr((size_t)n + 1), node((size_t)ne);
        std::vector<int32_t> Lptr((size_t)n + 1), Uptr((size_t)n + 1), Lnode((size_t)ne), Unode((size_t)ne);
        std::vector<int64_t> dest((size_t)ne);
        std::vector<int32_t> frows((size_t)n), brows((size_t)n), flev((size_t)n + 1), blev((size_t)n + 1);
        int32_t nf = 0, nb = 0;
        cudaError_t e = cudaStreamSynchronize(st);
        if (e == cudaSuccess)
            e = cudaMemcpy(ptr.data(), R->g->stored.ptr, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && ne > 0)
            e = cudaMemcpy(node.data(), R->g->stored.node, sizeof(int32_t) * (size_t)ne, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "ldu setup read-back", __FILE__, __LINE__);
        if (rc == SIGB_OK)
            rc = sigb_ldu_symbolic(n, ptr.data(), node.data(), Lptr.data(), Lnode.data(), Uptr.data(), Unode.data(),
                                   dest.data(), frows.data(), flev.data(), &nf, brows.data(), blev.data(), &nb);
        if (rc == SIGB_OK) {
            F->nL = (int64_t)Lptr[(size_t)n] - 1;
            F->nU = (int64_t)Uptr[(size_t)n] - 1;
            F->flev.assign(flev.begin(), flev.begin() + nf + 1);
            F->blev.assign(blev.begin(), blev.begin() + nb + 1);
            choose_chunking(F, Lptr, Lnode, Uptr, Unode);
            // statically scheduled sweeps where the pattern has a wavefront (both sweeps or neither)
            rc = upload(&F->Lptr, Lptr.data(), (size_t)n + 1);
            if (rc == SIGB_OK) rc = upload(&F->Uptr, Uptr.data(), (size_t)n + 1);
            if (rc == SIGB_OK) rc = upload(&F->Lnode, Lnode.data(), (size_t)F->nL);
            if (rc == SIGB_OK) rc = upload(&F->Unode, Unode.data(), (size_t)F->nU);
            if (rc == SIGB_OK) rc = upload(&F->dest, dest.data(), (size_t)ne);
            if (rc == SIGB_OK) rc = upload(&F->frows, frows.data(), (size_t)n);
            if (rc == SIGB_OK) rc = upload(&F->brows, brows.data(), (size_t)n);
            // statically scheduled sweeps where the pattern has a wavefront (both sweeps or neither): schedules
            // on the host, slots on the device from the patterns just uploaded
            if (rc == SIGB_OK && env_int("SIGB_LDU_STATIC", 1) != 0) {
                SweepPlan Pf, Pb;
                build_sweep_plan(n, Lptr.data(), Lnode.data(), 0, (int64_t)nf, Pf, false);
                if (Pf.eligible) build_sweep_plan(n, Uptr.data(), Unode.data(), 1, (int64_t)nb, Pb, false);
                if (Pf.eligible && Pb.eligible) {
                    rc = upload_sweep(Pf, F->Lptr, F->Lnode, F->fsw);
                    if (rc == SIGB_OK) rc = upload_sweep(Pb, F->Uptr, F->Unode, F->bsw);
                }
            }
            if (rc == SIGB_OK) {
                cudaError_t e2 = cudaMalloc((void **)&F->fac, sizeof(double) * (size_t)(F->nL + F->nU + n + 1));
                if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);   // the host vectors go out of scope below
                if (e2 != cudaSuccess) rc = cuda_fail(e2, "ldu setup", __FILE__, __LINE__);
            }
        }
        if (rc != SIGB_OK) {
            free_ldu(F);
            if (R != A) sigb_matrix_destroy(R);
            return rc;
        }
        s->ldu = F;
    }
    if (R != A) F->rows = R;

    // L%zero(), U%zero(), D = 0, then copy A in (:302-324)
    SIGB_CUDA(cudaMemsetAsync(F->fac, 0, sizeof(double) * (size_t)(F->nL + F->nU + n + 1), st));
    if (F->ne > 0) {
        ldu_scatter_kernel<<<grid_for(F->ne), kThreads, 0, st>>>(R->val, F->dest, F->ne, F->fac);
        count_launch();
    }
    // the elimination, level by level.  (Run on the forward sweep's static schedule as ONE CTA -- the same
    // dependencies -- it was no faster: 29.8 vs 28.4 ms at 1024^2, 113 vs 62 ms at 2048^2, a row of the
    // elimination being ~10x the instructions of a row of the sweep on one SM; visit r2y.)
    const int nlev = (int)F->flev.size() - 1;
    for (int l = 0; l < nlev; l++) {
        const int32_t b = F->flev[(size_t)l], cnt = F->flev[(size_t)l + 1] - b;
        ldu_factor_level_kernel<<<grid_for(cnt), kThreads, 0, st>>>(F->frows + b, cnt, F->Lptr, F->Lnode, F->Lval(),
                                                                   F->Uptr, F->Unode, F->Uval(), F->D());
        count_launch();
    }
    SIGB_CUDA(cudaGetLastError());
    // the factors in trip order for the statically scheduled sweeps
    SIGB_CHECK(pack_sweep(F->fsw, F->Lval()));
    SIGB_CHECK(pack_sweep(F->bsw, F->Uval()));
    return SIGB_OK;
}

// call pc%solve(A, x, b): x = b ; (I + L) x = x ; x = x / D ; (I + U) x = x   (:167-171)
int ldu_apply_dev(sigb_solver_t s, double *x, const double *b, const int *skip_flag)
{
    LduInfo *F = s->ldu;
    SIGB_REQUIRE(F, SIGB_ERR_STATE, "ldu solve: pc%%setup(A) has not been called");
    cudaStream_t st = ctx().stream;
    const int32_t n = F->n;
    if (F->fsw.on && F->bsw.on && n > 0) {
        // wavefront pattern: each sweep is one CTA on a schedule fixed at setup; x = b and x = x / D are
        // folded into the transposes that bring the right-hand sides into trip order
        SIGB_CHECK(launch_sweep_static<false>(F->fsw, b, nullptr, x, skip_flag));
        SIGB_CHECK(launch_sweep_static<true>(F->bsw, x, F->D(), x, skip_flag));
        return SIGB_OK;
    }
    if (F->nchunks > 0 && n > 0) {
        // deep schedule: two launches; x = b and x = x / D are folded into the sweeps' first reads
        SIGB_CHECK(launch_tri_chunked<false>(F, F->Lptr, F->Lnode, F->Lval(), b, x, skip_flag));
        SIGB_CHECK(launch_tri_chunked<true>(F, F->Uptr, F->Unode, F->Uval(), x, x, skip_flag));
        return SIGB_OK;
    }
    if (x != b) {
        copy_kernel<<<grid_for(n), kThreads, 0, st>>>(b, x, n, skip_flag);
        count_launch();
    }
    // level 0 of either sweep has no neighbours to subtract: nothing to do for it
    for (size_t l = 1; l + 1 < F->flev.size(); l++) {
        const int32_t b0 = F->flev[l], cnt = F->flev[l + 1] - b0;
        tri_level_kernel<<<grid_for(cnt), kThreads, 0, st>>>(F->frows + b0, cnt, F->Lptr, F->Lnode, F->Lval(), x,
                                                            skip_flag);
        count_launch();
    }
    divide_kernel<<<grid_for(n), kThreads, 0, st>>>(x, F->D(), n, skip_flag);
    count_launch();
    for (size_t l = 1; l + 1 < F->blev.size(); l++) {
        const int32_t b0 = F->blev[l], cnt = F->blev[l + 1] - b0;
        tri_level_kernel<<<grid_for(cnt), kThreads, 0, st>>>(F->brows + b0, cnt, F->Uptr, F->Unode, F->Uval(), x,
                                                            skip_flag);
        count_launch();
    }
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

int ldu_sizes(sigb_solver_t s, int32_t *n, int64_t *nL, int64_t *nU, int32_t *nflev, int32_t *nblev)
{
    LduInfo *F = s->ldu;
    SIGB_REQUIRE(F, SIGB_ERR_STATE, "ldu: pc%%setup(A) has not been called");
    if (n) *n = F->n;
    if (nL) *nL = F->nL;
    if (nU) *nU = F->nU;
    if (nflev) *nflev = (int32_t)F->flev.size() - 1;
    if (nblev) *nblev = (int32_t)F->blev.size() - 1;
    return SIGB_OK;
}

int ldu_read(sigb_solver_t s, int32_t *Lptr, int32_t *Lnode, double *Lval, int32_t *Uptr, int32_t *Unode,
             double *Uval, double *D)
{
    LduInfo *F = s->ldu;
    SIGB_REQUIRE(F, SIGB_ERR_STATE, "ldu: pc%%setup(A) has not been called");
    SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
    const size_t n = (size_t)F->n, nL = (size_t)F->nL, nU = (size_t)F->nU;
    if (Lptr) SIGB_CUDA(cudaMemcpy(Lptr, F->Lptr, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost));
    if (Uptr) SIGB_CUDA(cudaMemcpy(Uptr, F->Uptr, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost));
    if (Lnode && nL) SIGB_CUDA(cudaMemcpy(Lnode, F->Lnode, sizeof(int32_t) * nL, cudaMemcpyDeviceToHost));
    if (Unode && nU) SIGB_CUDA(cudaMemcpy(Unode, F->Unode, sizeof(int32_t) * nU, cudaMemcpyDeviceToHost));
    if (Lval && nL) SIGB_CUDA(cudaMemcpy(Lval, F->Lval(), sizeof(double) * nL, cudaMemcpyDeviceToHost));
    if (Uval && nU) SIGB_CUDA(cudaMemcpy(Uval, F->Uval(), sizeof(double) * nU, cudaMemcpyDeviceToHost));
    if (D && n) SIGB_CUDA(cudaMemcpy(D, F->D(), sizeof(double) * n, cudaMemcpyDeviceToHost));
    return SIGB_OK;
}

}  // namespace sigb
