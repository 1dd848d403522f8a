// krylov.cuh -- device-resident Krylov state and the fused vector kernels.
//
// The reference solvers are sequences of whole-array statements, each a
// separate pass over memory (cg_solvers.f90:134-143 = 1 SpMV + 2 dots + 3
// updates + 1 zero fill).  Here each iteration is 3 (CG) or 5 (BiCGSTAB)
// kernels; every scalar of the recurrence (alpha, beta, rho, omega, res2)
// stays on the device, every thread derives it from the reduced dot products,
// and the loop test `do while (dsqrt(res2) > tolerance)` is evaluated on the
// device after EVERY iteration and latched in `done[]`, so the solve stops at
// exactly the iteration the reference rule dictates even though the host only
// looks at the state once per batch of launches.
//
// Element-wise arithmetic follows the reference expression trees with
// explicit round-to-nearest mul/add (no FMA contraction); only the dot
// products differ from the serial reference, by summation order.
#pragma once

#include "device_utils.cuh"

namespace sigb {

struct KState {
    double pq;         // CG: p.q          BiCGSTAB: r0.v
    double rr[2];      // stopping quantity res2, parity-indexed (CG-PC: r.z)
    double rho[2];     // BiCGSTAB: r0.r, parity-indexed (rho / rho_old)
    double st, tt;     // BiCGSTAB: s.t, t.t
    double alpha[2];   // BiCGSTAB: alpha of the iteration with that parity
    double omega[2];
    double tol;
    double final_res2;
    long long iters;   // iterations performed by the current solve
    long long cap;     // safety cap (< 0: none) -- not in the reference
    long long itc[2];  // BiCGSTAB: parity-indexed copy of iters (race-free reads)
    int done[2];       // latch of the loop test, parity-indexed
    int capped;
    int started;       // persistent CG kernel: the initial residual has been formed (a later launch resumes)
    double lz[8];      // Lanczos scalars: [0] alpha, [1] beta, [2] c / norm2
};

// ---------------------------------------------------------------------------
// generic fused element-wise kernel: Op supplies
//   static constexpr int ND, NIN       dot products produced / values read per element
//   __device__ bool begin()            scalar prologue; false => whole grid exits
//   __device__ void load(i, in)        read element i's inputs (no stores)
//   __device__ void compute(i, in, acc) arithmetic + stores + dot contributions
//   __device__ double *out(d)          where dot d goes
// Loads of kUnroll elements are issued back to back before any arithmetic or
// store, so each thread keeps kUnroll * NIN independent 8-byte requests in
// flight (every request is a fully coalesced 256-byte warp access).
// ---------------------------------------------------------------------------
constexpr int kUnroll = 4;

template <class Op>
__global__ void __launch_bounds__(kThreads)
ew_kernel(Op op, int64_t n, double *partials, unsigned *ticket)
{
    if (!op.begin()) return;
    constexpr int ND = Op::ND;
    constexpr int NIN = Op::NIN > 0 ? Op::NIN : 1;
    double acc[ND > 0 ? ND : 1];
#pragma unroll
    for (int d = 0; d < (ND > 0 ? ND : 1); d++) acc[d] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t base = blockIdx.x * (int64_t)kThreads + threadIdx.x; base < n;
         base += stride * kUnroll) {
        double in[kUnroll][NIN];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const int64_t i = base + u * stride;
            if (i < n) op.load(i, in[u]);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const int64_t i = base + u * stride;
            if (i < n) op.compute(i, in[u], acc);
        }
    }
    if constexpr (ND == 1) {
        double *const out[1] = {op.out(0)};
        double v[1] = {acc[0]};
        grid_reduce<1>(v, partials, ticket, out);
    } else if constexpr (ND == 2) {
        double *const out[2] = {op.out(0), op.out(1)};
        double v[2] = {acc[0], acc[1]};
        grid_reduce<2>(v, partials, ticket, out);
    }
}

inline int ew_grid(int64_t n)
{
    int64_t g = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)ctx().num_sms * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

template <class Op>
int launch_ew(const Op &op, int64_t n, cudaStream_t st = nullptr)
{
    // one persistent wave: resident CTAs x SMs (register use differs per Op)
    int grid = 0;
    SIGB_CHECK((occupancy_grid<ew_kernel<Op>>(0, &grid)));
    const int need = ew_grid(n);
    if (need < grid) grid = need;
    ew_kernel<Op><<<grid, kThreads, 0, st ? st : ctx().stream>>>(op, n, ctx().partials, ctx().tickets);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

// Row-sharded operators, peer-memory transport: the same kernel with the Op's dot products all-reduced
// across the GPUs by its last CTA (device_utils.cuh grid_reduce) instead of a separate launch.
// A twin rather than a parameter of ew_kernel, so that the product kernels keep their code.
template <class Op>
__global__ void __launch_bounds__(kThreads)
ew_fused_kernel(Op op, int64_t n, double *partials, unsigned *ticket, const __grid_constant__ RedFuse red)
{
    if (!op.begin()) return;
    constexpr int ND = Op::ND;
    static_assert(ND == 1 || ND == 2, "fused all-reduce needs an Op that produces dot products");
    constexpr int NIN = Op::NIN > 0 ? Op::NIN : 1;
    double acc[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) acc[d] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t base = blockIdx.x * (int64_t)kThreads + threadIdx.x; base < n;
         base += stride * kUnroll) {
        double in[kUnroll][NIN];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const int64_t i = base + u * stride;
            if (i < n) op.load(i, in[u]);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const int64_t i = base + u * stride;
            if (i < n) op.compute(i, in[u], acc);
        }
    }
    if constexpr (ND == 1) {
        double *const out[1] = {op.out(0)};
        grid_reduce<1>(acc, partials, ticket, out, &red);
    } else {
        double *const out[2] = {op.out(0), op.out(1)};
        grid_reduce<2>(acc, partials, ticket, out, &red);
    }
}

template <class Op>
int launch_ew_fused(const Op &op, int64_t n, const RedFuse &red)
{
    int grid = 0;
    SIGB_CHECK((occupancy_grid<ew_fused_kernel<Op>>(0, &grid)));
    const int need = ew_grid(n);
    if (need < grid) grid = need;
    ew_fused_kernel<Op><<<grid, kThreads, 0, ctx().stream>>>(op, n, ctx().partials, ctx().tickets, red);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

__device__ __forceinline__ bool first_thread() { return blockIdx.x == 0 && threadIdx.x == 0; }

}  // namespace sigb
