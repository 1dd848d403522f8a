// device_utils.cuh -- device helpers shared by the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "internal.h"

namespace sigb {

// The reference is compiled without FMA contraction (default gfortran on
// x86-64): a*b is rounded, then added.  Every floating-point statement of the
// hot path is written with these two so nvcc can never contract them.
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

// Streaming (read-once) 128-bit loads: keep the matrix arrays out of L1 so the
// gathered x lines stay resident.
__device__ __forceinline__ int4 ld_stream_i4(const int32_t *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ int2 ld_stream_i2(const int32_t *p)
{
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld_stream_d2(const double *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                 : "=d"(r.x), "=d"(r.y)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream_d(const double *p)
{
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

// ---------------------------------------------------------------------------
// System-scope flag accesses for the peer-memory (NVLink) transport.  Data is
// written with plain stores, then published with a release store of a
// sequence number; consumers spin on an acquire load.  Peer-written data is
// read with ld.global.cg (L2 is the coherence point; L1 may be stale).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One fp64 value in flight, "LL" style: each 8-byte word carries 32 payload
// bits and a 32-bit sequence flag.  An aligned 8-byte store is delivered as a
// unit, so a word is either old or complete: no fence between payload and
// flag, one NVLink store latency per all-reduce.
struct RedEntry {
    unsigned int lo, flag_lo, hi, flag_hi;
};
constexpr int kRedSlots = 4;
constexpr int kRedVals = 3;
struct RedWin {
    RedEntry red[kRedSlots][kMaxRanks][kRedVals];  // [slot][source rank][value], written by the peers
    unsigned long long red_seq;                    // local: reductions completed
};

__device__ __forceinline__ void st_word(unsigned int *p, unsigned int payload, unsigned int flag)
{
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(flag) : "memory");
}
__device__ __forceinline__ uint2 ld_word(const unsigned int *p)
{
    uint2 r;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}

// ---------------------------------------------------------------------------
// Bounded waits.  Every device-side wait (grid barrier, all-reduce inbox, halo
// flags and acknowledgements, sync-free sweeps) goes through spin_wait: it polls
// flat out, and every 4096 polls looks at the clock and at the process-wide fault
// word.  A wait that outlives FaultBlock::limit_ns (default 30 s, environment
// SIGB_WAIT_TIMEOUT_MS) records its code in the fault word and gives up; once the
// word is set every later wait gives up after its first 4096 polls, the persistent
// kernels leave their loops, and the host turns the word into SIGB_ERR_COMM at
// its next synchronisation (api.cu check_fault) -- a late or dead peer produces an
// error, never a silently wrong result and never a hung GPU.  The block lives in
// mapped pinned host memory: the host reads it without a copy, the device only
// touches it on the slow path.
// ---------------------------------------------------------------------------
struct FaultBlock {
    unsigned int code;               // 0 = healthy; else the first FaultCode that timed out (sticky)
    unsigned int pad_;
    unsigned long long limit_ns;     // how long a single wait may last
};
enum FaultCode : unsigned {
    FAULT_GRID_BARRIER = 1, FAULT_ALLREDUCE = 2, FAULT_HALO_ACK = 3, FAULT_HALO_FLAG = 4, FAULT_LDU_SWEEP = 5
};

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Slow path of spin_wait, out of line so that the polling loops stay small.
// Returns false when the wait must be abandoned.
static __device__ __noinline__ bool spin_check(FaultBlock *fb, unsigned long long *t0, unsigned code)
{
    if (fb == nullptr) return true;
    if (*reinterpret_cast<volatile unsigned int *>(&fb->code) != 0u) return false;
    const unsigned long long now = global_timer_ns();
    if (*t0 == 0ull) { *t0 = now; return true; }
    if (now - *t0 > *reinterpret_cast<volatile unsigned long long *>(&fb->limit_ns)) {
        atomicCAS(&fb->code, 0u, code);
        __threadfence_system();
        return false;
    }
    return true;
}

// Polls ready() until it holds; false = gave up (fault recorded).
template <class Ready>
__device__ __forceinline__ bool spin_wait(Ready ready, FaultBlock *fb, unsigned code)
{
    unsigned spins = 0;
    unsigned long long t0 = 0ull;
    while (!ready()) {
        if ((++spins & 0xfffu) == 0u && !spin_check(fb, &t0, code)) return false;
    }
    return true;
}

// One fp64 through an inbox entry: two payload+flag words (see RedEntry).
__device__ __forceinline__ void red_entry_store(RedEntry *e, double v, unsigned int flag)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    st_word(&e->lo, (unsigned int)bits, flag);
    st_word(&e->hi, (unsigned int)(bits >> 32), flag);
}
__device__ __forceinline__ double red_entry_wait(const RedEntry *e, unsigned int flag, FaultBlock *fb)
{
    uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
    spin_wait([&] { lo = ld_word(&e->lo); return lo.y == flag; }, fb, FAULT_ALLREDUCE);
    spin_wait([&] { hi = ld_word(&e->hi); return hi.y == flag; }, fb, FAULT_ALLREDUCE);
    return __longlong_as_double((long long)(((unsigned long long)hi.x << 32) | lo.x));
}

// Warp-collective: lane q < nranks fetches rank q's contribution to value d of this slot (all ranks
// polled at once: one L2 round trip whatever the number of GPUs, where a single thread walking the
// ranks paid two dependent round trips per rank), then every lane adds the contributions in RANK ORDER,
// so all ranks -- and all lanes -- compute bit-identical totals.
__device__ __forceinline__ double warp_rank_sum(const RedWin *win, int slot, int d, int nranks, unsigned int flag,
                                                FaultBlock *fb)
{
    const int lane = threadIdx.x & 31;
    double mine = 0.0;
    if (lane < nranks) mine = red_entry_wait(&win->red[slot][lane][d], flag, fb);
    double g = 0.0;
    for (int q = 0; q < nranks; q++) g = add(g, __shfl_sync(0xffffffffu, mine, q));
    return g;
}

// Window of a row-sharded operator: everything peers write into this rank.
struct HaloWin {
    unsigned long long hflag[2][kMaxRanks];  // [buffer][source]: sequence number of the halo it holds
    unsigned long long ack[kMaxRanks];       // [consumer]: last sequence that consumer finished reading
    unsigned long long halo_seq;             // local: SpMVs completed (written by the last CTA)
    unsigned int push_ticket, done_ticket;   // local: last-CTA detection
    unsigned long long pad_[4];
    // followed by the two landing buffers
};

// ---------------------------------------------------------------------------
// Deterministic grid reduction.
//
// Each thread brings ND partial sums.  They are combined with a fixed shuffle
// tree inside the warp, a fixed order across the CTA's warps, and the CTA
// partial is parked in `partials`.  The CTA that takes the last ticket then
// adds the CTA partials in a fixed order and writes the totals.  For a given
// (grid, n) the summation tree is therefore identical from run to run.
// ---------------------------------------------------------------------------
template <int ND>
__device__ __forceinline__ void warp_tree(double (&v)[ND])
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int d = 0; d < ND; d++)
            v[d] = add(v[d], __shfl_down_sync(0xffffffffu, v[d], off));
    }
}

template <int ND>
__device__ __forceinline__ void block_tree(double (&v)[ND], double (*sm)[kThreads / 32])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_tree<ND>(v);
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < ND; d++) sm[d][warp] = v[d];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int d = 0; d < ND; d++) v[d] = (lane < kThreads / 32) ? sm[d][lane] : 0.0;
#pragma unroll
        for (int off = (kThreads / 64); off > 0; off >>= 1) {
#pragma unroll
            for (int d = 0; d < ND; d++)
                v[d] = add(v[d], __shfl_down_sync(0xffffffffu, v[d], off));
        }
    }
    __syncthreads();
}

// out[d] receive the grid totals (device pointers), plus *addend[d] when that
// pointer is non-null (a partial sum left by an earlier launch).  All threads
// of all CTAs must call this.  `ticket` must be zero on entry and is zero
// again on exit.
template <int ND>
__device__ __forceinline__ void grid_reduce(double (&acc)[ND], double *partials,
                                            unsigned *ticket, double *const (&out)[ND]);
template <int ND>
__device__ __forceinline__ void grid_reduce(double (&acc)[ND], double *partials,
                                            unsigned *ticket, double *const (&out)[ND],
                                            const double *const (&addend)[ND], const RedFuse *red = nullptr)
{
    __shared__ double sm[ND][kThreads / 32];
    __shared__ bool is_last;
    block_tree<ND>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < ND; d++) partials[(size_t)blockIdx.x * ND + d] = acc[d];
        __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double v[ND];
#pragma unroll
        for (int d = 0; d < ND; d++) v[d] = 0.0;
        for (unsigned j = threadIdx.x; j < gridDim.x; j += kThreads) {
#pragma unroll
            for (int d = 0; d < ND; d++)
                v[d] = add(v[d], __ldcg(&partials[(size_t)j * ND + d]));
        }
        block_tree<ND>(v, sm);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int d = 0; d < ND; d++) {
                v[d] = addend[d] ? add(__ldcg(addend[d]), v[d]) : v[d];
                *out[d] = v[d];
            }
            *ticket = 0u;
        }
        // Row-sharded operators on the peer-memory transport: the cross-GPU part, by warp 0 of this last CTA --
        // what red_kernel (comm.cu) does in a launch of its own: lane q stores this rank's sums
        // into rank q's inbox as payload+flag words, lane d adds the nranks contributions to
        // value d in rank order (bit-identical totals on every rank) and overwrites *out[d].
        if (red != nullptr && red->nranks > 1 && threadIdx.x < 32) {
            const int lane = threadIdx.x;
            const unsigned long long seq = red->win->red_seq + 1;
            const int slot = (int)(seq & (kRedSlots - 1));
            const unsigned int flag = (unsigned int)seq;
#pragma unroll
            for (int d = 0; d < ND; d++) {
                const double local = __shfl_sync(0xffffffffu, v[d], 0);
                if (lane < red->nranks) red_entry_store(&red->peer[lane]->red[slot][red->me][d], local, flag);
            }
            __syncwarp();
#pragma unroll
            for (int d = 0; d < ND; d++) {
                const double g = warp_rank_sum(red->win, slot, d, red->nranks, flag, red->fault);
                if (lane == 0) *out[d] = g;
            }
            __syncwarp();
            if (lane == 0) red->win->red_seq = seq;
        }
    }
}

template <int ND>
__device__ __forceinline__ void grid_reduce(double (&acc)[ND], double *partials,
                                            unsigned *ticket, double *const (&out)[ND])
{
    const double *addend[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) addend[d] = nullptr;
    const double *const(&ref)[ND] = addend;
    grid_reduce<ND>(acc, partials, ticket, out, ref);
}

template <int ND>
__device__ __forceinline__ void grid_reduce(double (&acc)[ND], double *partials,
                                            unsigned *ticket, double *const (&out)[ND], const RedFuse *red)
{
    const double *addend[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) addend[d] = nullptr;
    const double *const(&ref)[ND] = addend;
    grid_reduce<ND>(acc, partials, ticket, out, ref, red);
}

// Persistent grid = resident CTAs per SM x SMs, queried once per kernel.  The
// kernel is a template VALUE parameter: kernels sharing a signature must not
// share the cache (their register counts, hence residency, differ).
template <auto Kernel>
int occupancy_grid(size_t smem, int *grid_out)
{
    static thread_local int cached = 0;   // per host thread = per device (cudaFuncSetAttribute is per device)
    if (cached == 0) {
        int per_sm = 0;
        SIGB_CUDA(cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SIGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, Kernel, kThreads, smem));
        if (per_sm < 1) per_sm = 1;
        int g = per_sm * ctx().num_sms;
        if (g > kMaxGrid) g = (kMaxGrid / ctx().num_sms) * ctx().num_sms;
        cached = g;
    }
    *grid_out = cached;
    return SIGB_OK;
}

}  // namespace sigb
