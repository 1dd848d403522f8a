// solvers.cu -- device-resident CG, BiCGSTAB, Jacobi and Lanczos.
//
// Replaces the bodies of
//   cg_solve / cg_solve_pc               src/solver/cg_solvers.f90:116-194
//   bicgstab_solve / bicgstab_solve_pc   src/solver/bicgstab_solvers.f90:124-237
//   jacobi_setup / jacobi_solve          src/solver/jacobi_solvers.f90:37-81
//   lanczos / eigensolve                 src/eigensolver.f90:27-90,160-184
// keeping their stopping rules, work-vector sets and iteration counters.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "krylov.cuh"
#include "dist.h"
#include "solvers.h"

namespace sigb {

// ===========================================================================
// CG kernels
// ===========================================================================

// r = b - q ; p = r (or: z = idiag*r ; p = z) ; res2 = r.r (or r.z)
// cg_solvers.f90:129-131 / :169-172
struct CgInitOp {
    static constexpr int ND = 1;
    const double *__restrict__ b, *__restrict__ q, *__restrict__ idiag;
    double *__restrict__ r, *__restrict__ p, *__restrict__ z;
    KState *st;
    static constexpr int NIN = 3;
    __device__ bool begin() { return true; }
    __device__ void load(int64_t i, double *in)
    {
        in[0] = b[i];
        in[1] = q[i];
        if (idiag) in[2] = idiag[i];
    }
    __device__ void compute(int64_t i, const double *in, double *acc)
    {
        const double ri = sub(in[0], in[1]);
        r[i] = ri;
        double zi = ri;
        if (idiag) { zi = mul(in[2], ri); z[i] = zi; }
        p[i] = zi;
        acc[0] = add(acc[0], mul(ri, zi));
    }
    __device__ double *out(int) { return &st->rr[0]; }
};

// latch of the first loop test (cg_solvers.f90:133 before the first pass)
__global__ void latch0_kernel(KState *st)
{
    st->iters = 0;
    st->capped = 0;
    const bool go = sqrt(st->rr[0]) > st->tol;
    int done = go ? 0 : 1;
    if (go && st->cap == 0) { done = 1; st->capped = 1; }
    st->done[0] = done;
    st->done[1] = done;
    st->final_res2 = st->rr[0];
}

// r = r - alpha q ; [z = idiag r] ; dpr = r.r (r.z)       cg_solvers.f90:136,138-140 / :178-183
// (x = x + alpha p, :137, is carried out by CgDirectionOp, which reads p anyway:
//  identical arithmetic, one pass over p less -- 84 n instead of 92 n bytes of
//  vector traffic per iteration)
struct CgUpdateOp {
    static constexpr int ND = 1;
    const double *__restrict__ q, *__restrict__ idiag;
    double *__restrict__ r, *__restrict__ z;
    KState *st;
    int par;
    double alpha;
    __device__ bool begin()
    {
        if (st->done[par]) return false;
        alpha = st->rr[par] / st->pq;   // alpha = res2 / dpr
        return true;
    }
    static constexpr int NIN = 3;
    __device__ void load(int64_t i, double *in)
    {
        in[0] = r[i];
        in[1] = q[i];
        if (idiag) in[2] = idiag[i];
    }
    __device__ void compute(int64_t i, const double *in, double *acc)
    {
        const double ri = sub(in[0], mul(alpha, in[1]));
        r[i] = ri;
        double zi = ri;
        if (idiag) { zi = mul(in[2], ri); z[i] = zi; }
        acc[0] = add(acc[0], mul(ri, zi));
    }
    __device__ double *out(int) { return &st->rr[par ^ 1]; }
};

// x = x + alpha p (:137) ; beta = dpr / res2 ; p = r + beta p (p = z + beta p) ;
// res2 = dpr ; iterations += 1 ; evaluate the loop test for the next pass.
// cg_solvers.f90:141-145 / :184-189
struct CgDirectionOp {
    static constexpr int ND = 0;
    const double *__restrict__ rz;  // r, or z when preconditioned
    double *__restrict__ p, *__restrict__ x;
    KState *st;
    int par;
    double beta, alpha;
    __device__ bool begin()
    {
        if (st->done[par]) {
            if (first_thread()) st->done[par ^ 1] = 1;
            return false;
        }
        const double dpr = st->rr[par ^ 1];
        beta = dpr / st->rr[par];
        alpha = st->rr[par] / st->pq;   // the alpha of this iteration, recomputed from the same operands
        if (first_thread()) {
            const long long it = st->iters + 1;
            st->iters = it;
            st->final_res2 = dpr;
            int done = (sqrt(dpr) > st->tol) ? 0 : 1;
            if (!done && st->cap >= 0 && it >= st->cap) { done = 1; st->capped = 1; }
            st->done[par ^ 1] = done;
        }
        return true;
    }
    static constexpr int NIN = 3;
    __device__ void load(int64_t i, double *in) { in[0] = rz[i]; in[1] = p[i]; in[2] = x[i]; }
    __device__ void compute(int64_t i, const double *in, double *)
    {
        x[i] = add(in[2], mul(alpha, in[1]));
        p[i] = add(in[0], mul(beta, in[1]));
    }
    __device__ double *out(int) { return nullptr; }
};

// ---- CG with a preconditioner that is applied by its own kernels (ldu) -----
// p = z ; res2 = r.z   (cg_solvers.f90:171-172)
struct PcInitOp {
    static constexpr int ND = 1;
    static constexpr int NIN = 2;
    const double *__restrict__ r, *__restrict__ z;
    double *__restrict__ p;
    KState *st;
    __device__ bool begin() { return true; }
    __device__ void load(int64_t i, double *in) { in[0] = r[i]; in[1] = z[i]; }
    __device__ void compute(int64_t i, const double *in, double *acc)
    {
        p[i] = in[1];
        acc[0] = add(acc[0], mul(in[0], in[1]));
    }
    __device__ double *out(int) { return &st->rr[0]; }
};

// dpr = r.z   (cg_solvers.f90:183), skipped past the stopping latch
struct PcDotOp {
    static constexpr int ND = 1;
    static constexpr int NIN = 2;
    const double *__restrict__ r, *__restrict__ z;
    KState *st;
    int par;
    __device__ bool begin() { return st->done[par] == 0; }
    __device__ void load(int64_t i, double *in) { in[0] = r[i]; in[1] = z[i]; }
    __device__ void compute(int64_t, const double *in, double *acc) { acc[0] = add(acc[0], mul(in[0], in[1])); }
    __device__ double *out(int) { return &st->rr[par ^ 1]; }
};

// ===========================================================================
// BiCGSTAB kernels
// ===========================================================================

// r0 = b - q (or idiag*(b - q)) ; r = r0 ; v = 0 ; p = 0 ; res2 = r.r ; rho = r0.r
// bicgstab_solvers.f90:141-152 / :200-212
struct BicgInitOp {
    static constexpr int ND = 2;
    const double *__restrict__ b, *__restrict__ q, *__restrict__ idiag;
    double *__restrict__ r, *__restrict__ r0, *__restrict__ v, *__restrict__ p, *__restrict__ z;
    KState *st;
    static constexpr int NIN = 3;
    __device__ bool begin() { return true; }
    __device__ void load(int64_t i, double *in)
    {
        in[0] = b[i];
        in[1] = q[i];
        if (idiag) in[2] = idiag[i];
    }
    __device__ void compute(int64_t i, const double *in, double *acc)
    {
        double ri = sub(in[0], in[1]);
        if (idiag) { z[i] = ri; ri = mul(in[2], ri); }
        r0[i] = ri;
        r[i] = ri;
        v[i] = 0.0;
        p[i] = 0.0;
        acc[0] = add(acc[0], mul(ri, ri));
        acc[1] = add(acc[1], mul(ri, ri));
    }
    __device__ double *out(int d) { return d == 0 ? &st->rr[0] : &st->rho[0]; }
};

__global__ void bicg_latch0_kernel(KState *st)
{
    st->iters = 0;
    st->capped = 0;
    // rho_old = alpha = omega = 1 (bicgstab_solvers.f90:144-147) live in the
    // "previous iteration" slots
    st->rho[1] = 1.0;
    st->alpha[1] = 1.0;
    st->omega[1] = 1.0;
    const bool go = sqrt(st->rr[0]) > st->tol;
    int done = go ? 0 : 1;
    if (go && st->cap == 0) { done = 1; st->capped = 1; }
    // done[1] plays "previous iteration not finished" for the first direction
    // update; done[0] is written by that update
    st->done[1] = done;
    st->done[0] = done;
    st->itc[0] = 0;
    st->itc[1] = 0;
    st->final_res2 = st->rr[0];
}

// Closes iteration `par^1`... see BicgDirectionOp::begin: evaluates the loop
// test on rr[cur], then beta = rho/rho_old*alpha/omega ;
// p = r + beta*(p - omega*v)   (bicgstab_solvers.f90:154-157)
struct BicgDirectionOp {
    static constexpr int ND = 0;
    const double *__restrict__ r, *__restrict__ v;
    double *__restrict__ p;
    KState *st;
    int cur;      // parity of the iteration this p belongs to
    int initial;  // 1: first direction of a solve (no iteration to close)
    double beta, omega_old;
    __device__ bool begin()
    {
        const int prev = cur ^ 1;
        if (initial) {
            if (st->done[cur]) return false;
        } else {
            if (st->done[prev]) {
                if (first_thread()) st->done[cur] = 1;
                return false;
            }
            const double res2 = st->rr[cur];
            int done = (sqrt(res2) > st->tol) ? 0 : 1;
            int capped = 0;
            // the counter is parity-indexed so that no thread reads a word
            // another thread of this grid writes
            const long long it = st->itc[prev] + 1;
            if (!done && st->cap >= 0 && it >= st->cap) { done = 1; capped = 1; }
            if (first_thread()) {
                st->itc[cur] = it;
                st->iters = it;
                st->final_res2 = res2;
                if (capped) st->capped = 1;
                st->done[cur] = done;
            }
            if (done) return false;
        }
        omega_old = st->omega[prev];
        beta = st->rho[cur] / st->rho[prev] * st->alpha[prev] / omega_old;
        return true;
    }
    static constexpr int NIN = 3;
    __device__ void load(int64_t i, double *in) { in[0] = r[i]; in[1] = p[i]; in[2] = v[i]; }
    __device__ void compute(int64_t i, const double *in, double *)
    {
        p[i] = add(in[0], mul(beta, sub(in[1], mul(omega_old, in[2]))));
    }
    __device__ double *out(int) { return nullptr; }
};

// alpha = rho / (r0.v) ; s = r - alpha v     (bicgstab_solvers.f90:160-161)
struct BicgSOp {
    static constexpr int ND = 0;
    const double *__restrict__ r, *__restrict__ v;
    double *__restrict__ s;
    KState *st;
    int par;
    double alpha;
    __device__ bool begin()
    {
        if (st->done[par]) return false;
        alpha = st->rho[par] / st->pq;
        if (first_thread()) st->alpha[par] = alpha;
        return true;
    }
    static constexpr int NIN = 2;
    __device__ void load(int64_t i, double *in) { in[0] = r[i]; in[1] = v[i]; }
    __device__ void compute(int64_t i, const double *in, double *) { s[i] = sub(in[0], mul(alpha, in[1])); }
    __device__ double *out(int) { return nullptr; }
};

// omega = (s.t)/(t.t) [isnan -> 0] ; x = x + alpha p + omega s ; r = s - omega t ;
// res2 = r.r ; (next) rho = r0.r         (bicgstab_solvers.f90:164-169,155)
struct BicgUpdateOp {
    static constexpr int ND = 2;
    const double *__restrict__ p, *__restrict__ s, *__restrict__ t, *__restrict__ r0;
    double *__restrict__ x, *__restrict__ r;
    KState *st;
    int par;
    int nan_guard;  // unpreconditioned solver only (bicgstab_solvers.f90:165)
    double alpha, omega;
    __device__ bool begin()
    {
        if (st->done[par]) return false;
        alpha = st->alpha[par];
        omega = st->st / st->tt;
        if (nan_guard && isnan(omega)) omega = 0.0;
        if (first_thread()) st->omega[par] = omega;
        return true;
    }
    static constexpr int NIN = 5;
    __device__ void load(int64_t i, double *in)
    {
        in[0] = s[i];
        in[1] = x[i];
        in[2] = p[i];
        in[3] = t[i];
        in[4] = r0[i];
    }
    __device__ void compute(int64_t i, const double *in, double *acc)
    {
        const double si = in[0];
        x[i] = add(add(in[1], mul(alpha, in[2])), mul(omega, si));
        const double ri = sub(si, mul(omega, in[3]));
        r[i] = ri;
        acc[0] = add(acc[0], mul(ri, ri));
        acc[1] = add(acc[1], mul(in[4], ri));
    }
    __device__ double *out(int d) { return d == 0 ? &st->rr[par ^ 1] : &st->rho[par ^ 1]; }
};

// ===========================================================================
// Jacobi
// ===========================================================================

// x = idiag * b   (jacobi_solvers.f90:77)
struct JacobiApplyOp {
    static constexpr int ND = 0;
    const double *__restrict__ idiag, *__restrict__ b;
    double *__restrict__ x;
    static constexpr int NIN = 2;
    __device__ bool begin() { return true; }
    __device__ void load(int64_t i, double *in) { in[0] = idiag[i]; in[1] = b[i]; }
    __device__ void compute(int64_t i, const double *in, double *) { x[i] = mul(in[0], in[1]); }
    __device__ double *out(int) { return nullptr; }
};

// idiag(i) = 1 / A%get_value(i, i): scan line i of the stored arrays for id i,
// last hit wins, 0 when absent (csr_matrix_get_value cs_matrices.f90:709-724,
// csc_matrix_get_value :729-744).
__global__ void __launch_bounds__(kThreads)
jacobi_setup_cs_kernel(const int32_t *__restrict__ ptr1, const int32_t *__restrict__ node1,
                       const double *__restrict__ val, int32_t n, double *__restrict__ idiag)
{
    for (int32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        double z = 0.0;
        for (int32_t k = ptr1[i] - 1; k < ptr1[i + 1] - 1; k++)
            if (node1[k] == i + 1) z = val[k];
        idiag[i] = 1.0 / z;
    }
}

// ellpack_matrix_get_value ellpack_matrices.f90:220-237: first degrees(i) slots
__global__ void __launch_bounds__(kThreads)
jacobi_setup_ell_kernel(const int32_t *__restrict__ node_sm, const double *__restrict__ val_sm,
                        const int32_t *__restrict__ degrees, int32_t n, int32_t n_pad,
                        double *__restrict__ idiag)
{
    for (int32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        double z = 0.0;
        const int32_t d = degrees[i];
        for (int32_t k = 0; k < d; k++)
            if (node_sm[(size_t)k * n_pad + i] == i + 1) z = val_sm[(size_t)k * n_pad + i];
        idiag[i] = 1.0 / z;
    }
}

// ===========================================================================
// Lanczos kernels
// ===========================================================================

// dot(a, b) -> *out
struct DotOp {
    static constexpr int ND = 1;
    const double *__restrict__ a, *__restrict__ b;
    double *o;
    static constexpr int NIN = 2;
    __device__ bool begin() { return true; }
    __device__ void load(int64_t i, double *in) { in[0] = a[i]; in[1] = b[i]; }
    __device__ void compute(int64_t, const double *in, double *acc) { acc[0] = add(acc[0], mul(in[0], in[1])); }
    __device__ double *out(int) { return o; }
};

// dst = src / sqrt(*norm2)  (eigensolver.f90:52 ; :59,:79 with beta = sqrt(w.w))
// optionally records alpha / beta into T(:, col)  (:60-62, :80-82)
struct ScaleOp {
    static constexpr int ND = 0;
    const double *src;  // may alias dst
    double *dst;
    const double *norm2;
    const double *alpha;  // may be null
    double *Tcol;         // T(1:3, col) or null
    double *beta_out;     // where beta is kept for the next step, or null
    double d;
    __device__ bool begin()
    {
        d = sqrt(*norm2);
        if (first_thread()) {
            if (Tcol) { Tcol[1] = *alpha; Tcol[2] = d; Tcol[0] = d; }
            if (beta_out) *beta_out = d;
        }
        return true;
    }
    static constexpr int NIN = 1;
    __device__ void load(int64_t i, double *in) { in[0] = src[i]; }
    __device__ void compute(int64_t i, const double *in, double *) { dst[i] = in[0] / d; }
    __device__ double *out(int) { return nullptr; }
};

// w = w - alpha*qi [- beta*qim1] ; then dot(w, nxt or w) -> *o
// eigensolver.f90:57-58 / :70 (+ the first dot of the sweep :75 or of :78)
struct LanczosRecurOp {
    static constexpr int ND = 1;
    double *__restrict__ w;
    const double *__restrict__ qi, *__restrict__ qim1;  // qim1 may be null
    const double *__restrict__ nxt;                     // null => w.w
    const double *alpha_p, *beta_p;
    double *o;
    double alpha, beta;
    __device__ bool begin()
    {
        alpha = *alpha_p;
        beta = qim1 ? *beta_p : 0.0;
        return true;
    }
    static constexpr int NIN = 4;
    __device__ void load(int64_t i, double *in)
    {
        in[0] = w[i];
        in[1] = qi[i];
        if (qim1) in[2] = qim1[i];
        if (nxt) in[3] = nxt[i];
    }
    __device__ void compute(int64_t i, const double *in, double *acc)
    {
        double wi = sub(in[0], mul(alpha, in[1]));
        if (qim1) wi = sub(wi, mul(beta, in[2]));
        w[i] = wi;
        acc[0] = add(acc[0], mul(nxt ? in[3] : wi, wi));
    }
    __device__ double *out(int) { return o; }
};

// w = w - c*qk ; then dot(nxt or w, w) -> *o     (eigensolver.f90:75)
struct LanczosOrthoOp {
    static constexpr int ND = 1;
    double *__restrict__ w;
    const double *__restrict__ qk, *__restrict__ nxt;
    const double *c_p;
    double *o;
    double c;
    __device__ bool begin() { c = *c_p; return true; }
    static constexpr int NIN = 3;
    __device__ void load(int64_t i, double *in)
    {
        in[0] = w[i];
        in[1] = qk[i];
        if (nxt) in[2] = nxt[i];
    }
    __device__ void compute(int64_t i, const double *in, double *acc)
    {
        const double wi = sub(in[0], mul(c, in[1]));
        w[i] = wi;
        acc[0] = add(acc[0], mul(nxt ? in[2] : wi, wi));
    }
    __device__ double *out(int) { return o; }
};

// generalized Lanczos (eigensolver.f90:130-131 / :146-147):
// v = w - beta*zprev ; dot(v, q) -> *o
struct GenLanczosVOp {
    static constexpr int ND = 1;
    const double *__restrict__ w, *__restrict__ zprev, *__restrict__ q;
    double *__restrict__ v;
    const double *beta_p;
    double *o;
    double beta;
    __device__ bool begin() { beta = *beta_p; return true; }
    static constexpr int NIN = 3;
    __device__ void load(int64_t i, double *in) { in[0] = w[i]; in[1] = zprev[i]; in[2] = q[i]; }
    __device__ void compute(int64_t i, const double *in, double *acc)
    {
        const double vi = sub(in[0], mul(beta, in[1]));
        v[i] = vi;
        acc[0] = add(acc[0], mul(vi, in[2]));
    }
    __device__ double *out(int) { return o; }
};

// v = v - alpha*z   (eigensolver.f90:132)
struct GenLanczosAxpyOp {
    static constexpr int ND = 0;
    double *__restrict__ v;
    const double *__restrict__ z;
    const double *alpha_p;
    double alpha;
    __device__ bool begin() { alpha = *alpha_p; return true; }
    static constexpr int NIN = 2;
    __device__ void load(int64_t i, double *in) { in[0] = v[i]; in[1] = z[i]; }
    __device__ void compute(int64_t i, const double *in, double *) { v[i] = sub(in[0], mul(alpha, in[1])); }
    __device__ double *out(int) { return nullptr; }
};

// beta = sqrt(w.v) ; q_next = w / beta ; z_next = v / beta ; T(:, i)   (eigensolver.f90:136-142)
struct GenLanczosScaleOp {
    static constexpr int ND = 0;
    const double *__restrict__ w, *__restrict__ v;
    double *__restrict__ qn, *__restrict__ zn;
    const double *norm2, *alpha;
    double *Tcol, *beta_out;
    double d;
    __device__ bool begin()
    {
        d = sqrt(*norm2);
        if (first_thread()) {
            Tcol[1] = *alpha; Tcol[2] = d; Tcol[0] = d;
            *beta_out = d;
        }
        return true;
    }
    static constexpr int NIN = 2;
    __device__ void load(int64_t i, double *in) { in[0] = w[i]; in[1] = v[i]; }
    __device__ void compute(int64_t i, const double *in, double *) { qn[i] = in[0] / d; zn[i] = in[1] / d; }
    __device__ double *out(int) { return nullptr; }
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// random_number(Q(:,1)) ; Q(:,1) = 2*Q(:,1) - 1   (eigensolver.f90:50-51);
// counter-based so the vector does not depend on the launch shape or sharding
struct RandomOp {
    static constexpr int ND = 0;
    double *__restrict__ q;
    uint64_t seed;
    int64_t offset;  // global index of element 0
    static constexpr int NIN = 0;
    __device__ bool begin() { return true; }
    __device__ void load(int64_t, double *) {}
    __device__ void compute(int64_t i, const double *, double *)
    {
        const uint64_t h = splitmix64(seed ^ splitmix64((uint64_t)(i + offset)));
        const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
        q[i] = sub(mul(2.0, u), 1.0);
    }
    __device__ double *out(int) { return nullptr; }
};

// V2(l, j) = sum_k V(l, k) * Qm(k, j)   (V = matmul(V, Q), eigensolver.f90:176)
__global__ void __launch_bounds__(kThreads)
ritz_kernel(const double *__restrict__ V, const double *__restrict__ Qm, int64_t nr, int32_t n,
            double *__restrict__ V2)
{
    extern __shared__ double qs[];  // Qm, n*n
    for (int k = threadIdx.x; k < n * n; k += kThreads) qs[k] = Qm[k];
    __syncthreads();
    for (int64_t l = blockIdx.x * (int64_t)kThreads + threadIdx.x; l < nr;
         l += (int64_t)gridDim.x * kThreads) {
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int k = 0; k < n; k++) s = add(s, mul(V[(size_t)k * nr + l], qs[(size_t)j * n + k]));
            V2[(size_t)j * nr + l] = s;
        }
    }
}

// V(:, i) = V(1, i)/|V(1, i)| * V(:, i)   (eigensolver.f90:178-180)
__global__ void __launch_bounds__(kThreads)
sign_kernel(double *__restrict__ V, int64_t nr, int32_t n, const double *__restrict__ first_row)
{
    for (int j = 0; j < n; j++) {
        const double f = first_row[j];
        const double sg = f / fabs(f);
        for (int64_t l = blockIdx.x * (int64_t)kThreads + threadIdx.x; l < nr;
             l += (int64_t)gridDim.x * kThreads)
            V[(size_t)j * nr + l] = mul(sg, V[(size_t)j * nr + l]);
    }
}

__global__ void grab_first_row_kernel(const double *V, int64_t nr, int32_t n, double *first_row)
{
    for (int j = threadIdx.x; j < n; j += blockDim.x) first_row[j] = V[(size_t)j * nr];
}

// ---------------------------------------------------------------------------
// Strict-order dot products (sigb_solver_set_strict_order): a parity aid, not a fast path.
// Everything in the solvers except the dot products is element-wise and reproduces the reference
// statement for statement; the dot products differ from the serial loops only by the ORDER of the
// additions, which on an ill-conditioned nonsymmetric operator is enough to move the BiCGSTAB
// iteration count by several per cent.  With this switch every dot product the
// recurrence uses is recomputed after its producing kernel as sum = sum + a(i) * b(i), i = 1..n, by
// ONE thread (rounded product, then rounded add -- `sum(a * b)` of the reference, no FMA), so that the
// whole solve -- iterates, scalars, stopping iteration -- must equal the serial restatement BIT FOR
// BIT.  One GPU, kernel-per-phase path; used by the parity tests on parity-sized instances.
// ---------------------------------------------------------------------------
__global__ void seq_dot_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t n, double *out,
                               const int *skip)
{
    if (skip != nullptr && *skip != 0) return;
    double s = 0.0;
    int64_t i = 0;
    for (; i + 8 <= n; i += 8) {
        double pa[8], pb[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { pa[j] = a[i + j]; pb[j] = b[i + j]; }
#pragma unroll
        for (int j = 0; j < 8; j++) s = add(s, mul(pa[j], pb[j]));
    }
    for (; i < n; i++) s = add(s, mul(a[i], b[i]));
    *out = s;
}

static int strict_dot(sigb_solver_t s, const double *a, const double *b, double *out, const int *skip)
{
    if (!s->strict_order) return SIGB_OK;
    seq_dot_kernel<<<1, 1, 0, ctx().stream>>>(a, b, s->nn, out, skip);
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

// ===========================================================================
// host drivers
// ===========================================================================

static int sync_state(sigb_solver_t s)
{
    SIGB_CUDA(cudaMemcpyAsync(s->state_host, s->state, sizeof(KState), cudaMemcpyDeviceToHost,
                              ctx().stream));
    SIGB_CUDA(cudaStreamSynchronize(ctx().stream));
    return check_fault("solve");   // a device-side wait that timed out invalidates what was computed
}

static int push_state(sigb_solver_t s)
{
    memset(s->state_host, 0, sizeof(KState));
    s->state_host->tol = s->tol;
    s->state_host->cap = s->cap;
    SIGB_CUDA(cudaMemcpyAsync(s->state, s->state_host, sizeof(KState), cudaMemcpyHostToDevice,
                              ctx().stream));
    return SIGB_OK;
}

int jacobi_setup_dev(sigb_solver_t s, sigb_matrix_t A)
{
    if (A->op) return op_jacobi_setup(A, s->work);
    sigb_graph_t g = A->g;
    const int32_t n = A->nrow;
    cudaStream_t st = ctx().stream;
    const int grid = ew_grid(n);
    if (g->kind == G_ELL) {
        jacobi_setup_ell_kernel<<<grid, kThreads, 0, st>>>(g->ell_node, A->val, g->ell_degrees, n,
                                                          g->n_pad, s->work);
    } else {
        jacobi_setup_cs_kernel<<<grid, kThreads, 0, st>>>(g->stored.ptr, g->stored.node, A->val, n,
                                                         s->work);
    }
    count_launch();
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

int jacobi_apply_dev(sigb_solver_t s, double *x, const double *b)
{
    JacobiApplyOp op{s->work, b, x};
    return launch_ew(op, s->nn);
}

static int finish_solve(sigb_solver_t s, bool state_is_current = false)
{
    if (!state_is_current) SIGB_CHECK(sync_state(s));
    s->iterations += s->state_host->iters;
    s->res2 = s->state_host->final_res2;
    s->capped = s->state_host->capped;
    return SIGB_OK;
}

// The persistent kernel removes kernel boundaries and runs the all-reduces
// in-kernel, which wins when an iteration is short (sharded operators); its
// vector phases run at 4 CTAs/SM with a 64-register cap and stream a little
// slower than the dedicated kernels, which wins when an iteration is long.
// Measured cross-over on B200 (profiles/r1_size_sweep_persistent.jsonl): ~3 M
// rows per GPU.  SIGB_CG_PERSISTENT=1 / 0 forces one or the other.
static bool persistent_enabled(int64_t n_local, int nranks, int solver_choice)
{
    if (solver_choice >= 0) return solver_choice != 0;   // sigb_solver_set_persistent
    static int v = -2;
    if (v == -2) {
        const char *e = getenv("SIGB_CG_PERSISTENT");
        v = e ? (atoi(e) != 0 ? 1 : 0) : -1;
    }
    if (v >= 0) return v != 0;
    // cross-over on B200 with the 3-CTA kernel: ~4.2 M rows on one GPU (107.9 vs 108.8 us per iteration,
    // profiles/r2_visit_d_1gpu_summary.txt); on a sharded operator the persistent form also saves the
    // launches between the phases, so 4 GPUs x 4.2 M rows take it too
    return n_local <= (nranks > 1 ? 4500000 : 4000000);
}

// L2 persistence for the solver's work vectors (north_star: "x-vector reuse
// staged through shared memory and L2 persistence").  On a sharded operator the
// vectors of a rank (p, q, r, z: 4 x 17 MB at 2.1 M rows) fit in the 126 MB L2
// while the matrix streams through it with an evict-first policy; marking the
// work buffer persisting lets the vector phases and the SpMV gathers hit L2.
// Only applied when the whole buffer fits the device's persisting carve-out.
struct L2Window {
    bool on = false;
    cudaStream_t stream = nullptr;
    int begin(void *base, size_t bytes)
    {
        static thread_local int enabled = -1;
        static thread_local size_t max_persist = 0, max_window = 0;
        if (enabled < 0) {
            // measured on B200 (gpurun visit r1s, 2.1 M and 4.2 M rows): the carve-out
            // starves the matrix stream and the iteration gets 15-50 % SLOWER
            // (54 -> 63 us, 123 -> 188 us); the evict-first policy on the TMA
            // matrix loads is what keeps the vectors resident.  Opt-in only.
            const char *e = getenv("SIGB_L2_PERSIST");
            enabled = (e && atoi(e) == 1) ? 1 : 0;
            cudaDeviceProp prop;
            SIGB_CUDA(cudaGetDeviceProperties(&prop, ctx().device));
            max_persist = (size_t)prop.persistingL2CacheMaxSize;
            max_window = (size_t)prop.accessPolicyMaxWindowSize;
            if (enabled && max_persist > 0) SIGB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, max_persist));
        }
        if (!enabled || max_persist == 0 || bytes > max_persist || bytes > max_window) return SIGB_OK;
        stream = ctx().stream;
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = base;
        attr.accessPolicyWindow.num_bytes = bytes;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        SIGB_CUDA(cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        on = true;
        return SIGB_OK;
    }
    void end()
    {
        if (!on) return;
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        on = false;
    }
};

// iterations launched between two looks at the device state
static int batch_size(int64_t n)
{
    if (n < (1 << 16)) return 64;
    if (n < (1 << 20)) return 32;
    return 16;
}

static int cg_solve_body(sigb_solver_t s, sigb_matrix_t A, double *x, const double *b, sigb_solver_t pc);

int cg_solve_dev(sigb_solver_t s, sigb_matrix_t A, double *x, const double *b, sigb_solver_t pc)
{
    L2Window win;
    SIGB_CHECK(win.begin(s->work, sizeof(double) * (size_t)s->nvec * s->nwork));
    const int rc = cg_solve_body(s, A, x, b, pc);
    win.end();
    return rc;
}

// cg_solve_pc (cg_solvers.f90:155-194) with a preconditioner applied by its own kernels
// between the residual update and the direction update: call pc%solve(A, z, r) (:170, :181).
// Same device-resident control as the fused path: every kernel of an iteration launched
// past the stopping latch is a no-op.
static int cg_solve_ldu_pc(sigb_solver_t s, sigb_matrix_t A, double *x, const double *b, sigb_solver_t pc)
{
    const int64_t n = s->nn, nv = s->nvec;
    double *p = s->work, *q = p + nv, *r = q + nv, *z = r + nv;
    KState *st = s->state;
    SIGB_CHECK(push_state(s));
    DotSpec none;
    SIGB_CHECK(solver_matvec(A, x, q, none, false));            // z = x ; q = A z        :167-168
    CgInitOp init{b, q, nullptr, r, p, z, st};                  // r = b - q              :169
    SIGB_CHECK(launch_ew(init, n));
    SIGB_CHECK(ldu_apply_dev(pc, z, r, nullptr));               // call pc%solve(A, z, r) :170
    PcInitOp pinit{r, z, p, st};                                // p = z ; res2 = r.z     :171-172
    SIGB_CHECK(launch_ew(pinit, n));
    latch0_kernel<<<1, 1, 0, ctx().stream>>>(st);
    count_launch();
    const int nb = batch_size(n);
    int par = 0;
    for (;;) {
        for (int it = 0; it < nb; it++) {
            DotSpec d;
            d.ndot = 1;
            d.u = p;
            d.out[0] = &st->pq;
            d.skip_flag = &st->done[par];
            SIGB_CHECK(solver_matvec(A, p, q, d, false));               // q = A p ; dpr = p.q   :175-176
            CgUpdateOp up{q, nullptr, r, z, st, par, 0.0};              // r = r - alpha q       :179
            SIGB_CHECK(launch_ew(up, n));                               // (its r.r lands in rr[par^1] and is
            SIGB_CHECK(ldu_apply_dev(pc, z, r, &st->done[par]));        //  replaced by r.z below)   :181
            PcDotOp dz{r, z, st, par};                                  // dpr = r.z             :183
            SIGB_CHECK(launch_ew(dz, n));
            CgDirectionOp dir{z, p, x, st, par, 0.0, 0.0};              // x, beta, p = z + beta p  :178,184-189
            SIGB_CHECK(launch_ew(dir, n));
            par ^= 1;
        }
        SIGB_CHECK(sync_state(s));
        if (s->state_host->done[par]) break;
    }
    return finish_solve(s);
}

static int cg_solve_body(sigb_solver_t s, sigb_matrix_t A, double *x, const double *b, sigb_solver_t pc)
{
    if (pc && pc->kind == S_LDU) return cg_solve_ldu_pc(s, A, x, b, pc);
    const int64_t n = s->nn, nv = s->nvec;
    double *p = s->work, *q = p + nv, *r = q + nv, *z = r + nv;
    const double *idiag = pc ? pc->work : nullptr;
    KState *st = s->state;
    SIGB_CHECK(push_state(s));

    // ---- the whole solve as one persistent cooperative kernel ---------------
    // (CSR-shaped operators; sigb_solver_set_persistent / SIGB_CG_PERSISTENT=0 select the
    // kernel-per-phase path below, which is also what ELLPACK, operator expressions and the
    // NCCL transport use).  The kernel forms the initial residual itself: one launch per solve.
    {
        const CsrView *V = nullptr;
        const double *val = nullptr;
        PersistComm pcomm;
        DotSpec halo;
        bool eligible = true;
        SIGB_CHECK(dist_persist_info(A, &pcomm, &halo, &eligible));
        eligible = eligible && persistent_enabled(n, pcomm.nranks, s->persistent) && !s->strict_order;
        if (eligible && !A->op) {   // operator expressions run kernel-per-phase
            sigb_graph_t g = A->g;
            if (g->kind == G_CSR) {
                V = &g->stored;
                val = A->val;
            } else if (g->kind == G_CSC) {
                SIGB_CHECK(ensure_transposed(A));
                V = &g->transposed;
                val = A->val_t;
            }
        }
        // gather-bound operators (small tile shape): kernel per phase, where the SpMV gets the L1 share
        // it needs; the persistent kernel is compiled for the large stages
        if (V != nullptr && V->tile_nnz == kTileNnzSmall && s->persistent < 0) V = nullptr;
        if (V != nullptr) {
            if (!s->bar) {
                SIGB_CUDA(cudaMalloc((void **)&s->bar, 2 * sizeof(unsigned long long)));
                SIGB_CUDA(cudaMalloc((void **)&s->pers_partials, sizeof(double) * 4 * kMaxGrid));
            }
            for (;;) {
                SIGB_CHECK(cg_persistent_run(s, *V, val, halo, x, b, p, q, r, z, idiag, n, pcomm, 4096));
                SIGB_CHECK(sync_state(s));
                if (s->state_host->done[0]) break;
            }
            return finish_solve(s, /*state_is_current=*/true);
        }
    }

    DotSpec none;
    // q = A x ; r = b - q ; [z = M r] ; p = r|z ; res2 = r.r | r.z
    SIGB_CHECK(solver_matvec(A, x, q, none, /*x_has_halo=*/false));
    CgInitOp init{b, q, idiag, r, p, z, st};
    SIGB_CHECK(launch_ew(init, n));
    SIGB_CHECK(dist_allreduce(A, &st->rr[0], 1));
    SIGB_CHECK(strict_dot(s, r, idiag ? z : r, &st->rr[0], nullptr));
    latch0_kernel<<<1, 1, 0, ctx().stream>>>(st);
    count_launch();

    const int nb = batch_size(n);
    int par = 0;
    RedFuse rf;
    const bool fused = dist_red_fuse(A, &rf);   // peer-memory transport: the two all-reduces inside their producers
    for (;;) {
        for (int it = 0; it < nb; it++) {
            DotSpec d;
            d.ndot = 1;
            d.u = p;
            d.out[0] = &st->pq;
            d.skip_flag = &st->done[par];
            if (fused) d.red = &rf;
            SIGB_CHECK(solver_matvec(A, p, q, d, /*x_has_halo=*/true));   // q = A p ; dpr = p.q
            if (!fused) SIGB_CHECK(dist_allreduce(A, &st->pq, 1, &st->done[par]));
            SIGB_CHECK(strict_dot(s, p, q, &st->pq, &st->done[par]));
            CgUpdateOp up{q, idiag, r, z, st, par, 0.0};
            if (fused) SIGB_CHECK(launch_ew_fused(up, n, rf));
            else SIGB_CHECK(launch_ew(up, n));
            if (!fused) SIGB_CHECK(dist_allreduce(A, &st->rr[par ^ 1], 1, &st->done[par]));
            SIGB_CHECK(strict_dot(s, r, idiag ? z : r, &st->rr[par ^ 1], &st->done[par]));
            CgDirectionOp dir{idiag ? z : r, p, x, st, par, 0.0, 0.0};
            SIGB_CHECK(launch_ew(dir, n));
            par ^= 1;
        }
        SIGB_CHECK(sync_state(s));
        if (s->state_host->done[par]) break;
    }
    return finish_solve(s);
}

// ---------------------------------------------------------------------------
// bicgstab_solve_pc
// (bicgstab_solvers.f90:182-237) with pc = ldu().  The preconditioner is applied by its own
// kernels (ldu.cu), so the three products of an iteration are plain SpMVs into z followed by
// call pc%solve(A, ., z), and the dot products that the Jacobi form fuses into the SpMV run
// as separate passes.  Same device-resident control: everything launched past the stopping
// latch is a no-op.
// ---------------------------------------------------------------------------
// z = b - q   (:200)
struct BicgResidualOp {
    static constexpr int ND = 0;
    static constexpr int NIN = 2;
    const double *__restrict__ b, *__restrict__ q;
    double *__restrict__ z;
    __device__ bool begin() { return true; }
    __device__ void load(int64_t i, double *in) { in[0] = b[i]; in[1] = q[i]; }
    __device__ void compute(int64_t i, const double *in, double *) { z[i] = sub(in[0], in[1]); }
    __device__ double *out(int) { return nullptr; }
};

// r = r0 ; v = 0 ; p = 0 ; res2 = r.r ; rho = r0.r   (:202-212), r0 = pc%solve(A, ., z) done
struct BicgInitR0Op {
    static constexpr int ND = 2;
    static constexpr int NIN = 1;
    const double *__restrict__ r0;
    double *__restrict__ r, *__restrict__ v, *__restrict__ p;
    KState *st;
    __device__ bool begin() { return true; }
    __device__ void load(int64_t i, double *in) { in[0] = r0[i]; }
    __device__ void compute(int64_t i, const double *in, double *acc)
    {
        r[i] = in[0];
        v[i] = 0.0;
        p[i] = 0.0;
        acc[0] = add(acc[0], mul(in[0], in[0]));
        acc[1] = add(acc[1], mul(in[0], in[0]));
    }
    __device__ double *out(int d) { return d == 0 ? &st->rr[0] : &st->rho[0]; }
};

// r0.v -> pq   (:221), skipped past the stopping latch
struct BicgDotOp {
    static constexpr int ND = 1;
    static constexpr int NIN = 2;
    const double *__restrict__ a, *__restrict__ b;
    KState *st;
    int par;
    __device__ bool begin() { return st->done[par] == 0; }
    __device__ void load(int64_t i, double *in) { in[0] = a[i]; in[1] = b[i]; }
    __device__ void compute(int64_t, const double *in, double *acc) { acc[0] = add(acc[0], mul(in[0], in[1])); }
    __device__ double *out(int) { return &st->pq; }
};

// s.t -> st, t.t -> tt   (:225)
struct BicgDot2Op {
    static constexpr int ND = 2;
    static constexpr int NIN = 2;
    const double *__restrict__ s, *__restrict__ t;
    KState *st;
    int par;
    __device__ bool begin() { return st->done[par] == 0; }
    __device__ void load(int64_t i, double *in) { in[0] = s[i]; in[1] = t[i]; }
    __device__ void compute(int64_t, const double *in, double *acc)
    {
        acc[0] = add(acc[0], mul(in[0], in[1]));
        acc[1] = add(acc[1], mul(in[1], in[1]));
    }
    __device__ double *out(int d) { return d == 0 ? &st->st : &st->tt; }
};

static int bicgstab_solve_ldu_pc(sigb_solver_t s, sigb_matrix_t A, double *x, const double *b, sigb_solver_t pc)
{
    const int64_t n = s->nn, nv = s->nvec;
    double *p = s->work, *q = p + nv, *r = q + nv, *r0 = r + nv, *v = r0 + nv, *sv = v + nv,
           *t = sv + nv, *z = t + nv;
    KState *st = s->state;
    SIGB_CHECK(push_state(s));
    DotSpec none;
    SIGB_CHECK(solver_matvec(A, x, q, none, false));            // call A%matvec(x, q)      :199
    BicgResidualOp res{b, q, z};                                // z = b - q                :200
    SIGB_CHECK(launch_ew(res, n));
    SIGB_CHECK(ldu_apply_dev(pc, r0, z, nullptr));              // call pc%solve(A, r0, z)  :201
    BicgInitR0Op init{r0, r, v, p, st};                         // :202-212
    SIGB_CHECK(launch_ew(init, n));
    bicg_latch0_kernel<<<1, 1, 0, ctx().stream>>>(st);
    count_launch();
    {
        BicgDirectionOp dir{r, v, p, st, 0, 1, 0.0, 0.0};
        SIGB_CHECK(launch_ew(dir, n));
    }
    const int nb = batch_size(n);
    int par = 0;
    for (;;) {
        for (int it = 0; it < nb; it++) {
            DotSpec latched;
            latched.skip_flag = &st->done[par];
            SIGB_CHECK(solver_matvec(A, p, z, latched, false));             // call A%matvec(p, z)     :218
            SIGB_CHECK(ldu_apply_dev(pc, v, z, &st->done[par]));            // call pc%solve(A, v, z)  :219
            BicgDotOp d1{r0, v, st, par};                                   // r0.v                    :221
            SIGB_CHECK(launch_ew(d1, n));
            BicgSOp sop{r, v, sv, st, par, 0.0};                            // alpha ; s = r - alpha v :221-222
            SIGB_CHECK(launch_ew(sop, n));
            SIGB_CHECK(solver_matvec(A, sv, z, latched, false));            // call A%matvec(s, z)     :223
            SIGB_CHECK(ldu_apply_dev(pc, t, z, &st->done[par]));            // call pc%solve(A, t, z)  :224
            BicgDot2Op d2{sv, t, st, par};                                  // s.t, t.t                :225
            SIGB_CHECK(launch_ew(d2, n));
            BicgUpdateOp up{p, sv, t, r0, x, r, st, par, 0, 0.0, 0.0};      // omega, x, r, res2, rho  :225-230,215
            SIGB_CHECK(launch_ew(up, n));
            BicgDirectionOp dir{r, v, p, st, par ^ 1, 0, 0.0, 0.0};         // loop test, beta, p      :214-217
            SIGB_CHECK(launch_ew(dir, n));
            par ^= 1;
        }
        SIGB_CHECK(sync_state(s));
        if (s->state_host->done[par]) break;
    }
    return finish_solve(s);
}

int bicgstab_solve_dev(sigb_solver_t s, sigb_matrix_t A, double *x, const double *b,
                       sigb_solver_t pc)
{
    if (pc && pc->kind == S_LDU) return bicgstab_solve_ldu_pc(s, A, x, b, pc);
    const int64_t n = s->nn, nv = s->nvec;
    double *p = s->work, *q = p + nv, *r = q + nv, *r0 = r + nv, *v = r0 + nv, *sv = v + nv,
           *t = sv + nv, *z = t + nv;
    const double *idiag = pc ? pc->work : nullptr;
    KState *st = s->state;
    SIGB_CHECK(push_state(s));

    DotSpec none;
    SIGB_CHECK(solver_matvec(A, x, q, none, false));
    BicgInitOp init{b, q, idiag, r, r0, v, p, z, st};
    SIGB_CHECK(launch_ew(init, n));
    SIGB_CHECK(dist_allreduce(A, &st->rr[0], 3));  // rr[0], rr[1], rho[0] are contiguous
    SIGB_CHECK(strict_dot(s, r, r, &st->rr[0], nullptr));
    SIGB_CHECK(strict_dot(s, r0, r, &st->rho[0], nullptr));
    bicg_latch0_kernel<<<1, 1, 0, ctx().stream>>>(st);
    count_launch();
    {
        BicgDirectionOp dir{r, v, p, st, 0, 1, 0.0, 0.0};
        SIGB_CHECK(launch_ew(dir, n));
    }

    const int nb = batch_size(n);
    int par = 0;
    RedFuse rf;
    const bool fused = dist_red_fuse(A, &rf);   // peer-memory transport: the three all-reduces inside their producers
    for (;;) {
        for (int it = 0; it < nb; it++) {
            DotSpec d1;                      // v = [M] A p ; r0.v
            d1.ndot = 1;
            d1.u = r0;
            d1.out[0] = &st->pq;
            d1.skip_flag = &st->done[par];
            d1.row_scale = idiag;
            if (fused) d1.red = &rf;
            SIGB_CHECK(solver_matvec(A, p, v, d1, true));
            if (!fused) SIGB_CHECK(dist_allreduce(A, &st->pq, 1, &st->done[par]));
            SIGB_CHECK(strict_dot(s, r0, v, &st->pq, &st->done[par]));
            BicgSOp sop{r, v, sv, st, par, 0.0};
            SIGB_CHECK(launch_ew(sop, n));
            DotSpec d2;                      // t = [M] A s ; s.t, t.t
            d2.ndot = 2;
            d2.u = sv;
            d2.out[0] = &st->st;
            d2.out[1] = &st->tt;
            d2.skip_flag = &st->done[par];
            d2.row_scale = idiag;
            if (fused) d2.red = &rf;
            SIGB_CHECK(solver_matvec(A, sv, t, d2, true));
            if (!fused) SIGB_CHECK(dist_allreduce(A, &st->st, 2, &st->done[par]));
            SIGB_CHECK(strict_dot(s, sv, t, &st->st, &st->done[par]));
            SIGB_CHECK(strict_dot(s, t, t, &st->tt, &st->done[par]));
            BicgUpdateOp up{p, sv, t, r0, x, r, st, par, pc ? 0 : 1, 0.0, 0.0};
            if (fused) SIGB_CHECK(launch_ew_fused(up, n, rf));
            else SIGB_CHECK(launch_ew(up, n));
            if (!fused) SIGB_CHECK(dist_allreduce2(A, &st->rr[par ^ 1], &st->rho[par ^ 1], &st->done[par]));
            SIGB_CHECK(strict_dot(s, r, r, &st->rr[par ^ 1], &st->done[par]));
            SIGB_CHECK(strict_dot(s, r0, r, &st->rho[par ^ 1], &st->done[par]));
            BicgDirectionOp dir{r, v, p, st, par ^ 1, 0, 0.0, 0.0};
            SIGB_CHECK(launch_ew(dir, n));
            par ^= 1;
        }
        SIGB_CHECK(sync_state(s));
        if (s->state_host->done[par]) break;
    }
    return finish_solve(s);
}

// n-step Lanczos on device arrays: Q is nr x n column-major (ld = nr), T is 3 x n.
int lanczos_dev(sigb_matrix_t A, int32_t n, const double *q1, uint64_t seed, int64_t row_offset,
                double *T, double *Q, double *w, KState *st)
{
    const int64_t nr = A->nrow;
    cudaStream_t stream = ctx().stream;
    SIGB_CUDA(cudaMemsetAsync(T, 0, sizeof(double) * 3 * (size_t)n, stream));
    SIGB_CUDA(cudaMemsetAsync(Q, 0, sizeof(double) * (size_t)nr * n, stream));
    SIGB_CUDA(cudaMemsetAsync(w, 0, sizeof(double) * (size_t)nr, stream));
    auto col = [&](int c) { return Q + (size_t)(c - 1) * nr; };  // 1-based column

    if (q1) {
        SIGB_CUDA(cudaMemcpyAsync(col(1), q1, sizeof(double) * (size_t)nr, cudaMemcpyDeviceToDevice, stream));
    } else {
        RandomOp rnd{col(1), seed, row_offset};
        SIGB_CHECK(launch_ew(rnd, nr));
    }
    // Q(:,1) = Q(:,1) / dsqrt(sum(Q(:,1)*Q(:,1)))          :52
    {
        DotOp d{col(1), col(1), &st->lz[2]};
        SIGB_CHECK(launch_ew(d, nr));
        SIGB_CHECK(dist_allreduce(A, &st->lz[2], 1));
        ScaleOp sc{col(1), col(1), &st->lz[2], nullptr, nullptr, nullptr, 0.0};
        SIGB_CHECK(launch_ew(sc, nr));
    }
    if (n == 1) {
        DotSpec d;
        d.ndot = 1; d.u = col(1); d.out[0] = &T[1];
        SIGB_CHECK(solver_matvec(A, col(1), w, d, false));
        SIGB_CHECK(dist_allreduce(A, &T[1], 1));
        return SIGB_OK;
    }
    for (int i = 1; i <= n - 1; i++) {
        // w = A q_i ; alpha = q_i . w                         :55-56 / :68-69
        DotSpec d;
        d.ndot = 1; d.u = col(i); d.out[0] = &st->lz[0];
        SIGB_CHECK(solver_matvec(A, col(i), w, d, false));
        SIGB_CHECK(dist_allreduce(A, &st->lz[0], 1));
        // w = w - alpha q_i - beta q_{i-1}, fused with the next dot  :57 / :70
        const int nsweep = i - 2;  // re-orthogonalise against q_1 .. q_{i-2}   :74
        int slot = 3;
        {
            LanczosRecurOp op{w, col(i), i >= 2 ? col(i - 1) : nullptr, nsweep >= 1 ? col(1) : nullptr,
                              &st->lz[0], &st->lz[1], nsweep >= 1 ? &st->lz[slot] : &st->lz[2], 0.0, 0.0};
            SIGB_CHECK(launch_ew(op, nr));
            SIGB_CHECK(dist_allreduce(A, nsweep >= 1 ? &st->lz[slot] : &st->lz[2], 1));
        }
        for (int k = 1; k <= nsweep; k++) {
            const bool last = (k == nsweep);
            const int nslot = (slot == 3) ? 4 : 3;
            LanczosOrthoOp op{w, col(k), last ? nullptr : col(k + 1), &st->lz[slot],
                              last ? &st->lz[2] : &st->lz[nslot], 0.0};
            SIGB_CHECK(launch_ew(op, nr));
            SIGB_CHECK(dist_allreduce(A, last ? &st->lz[2] : &st->lz[nslot], 1));
            slot = nslot;
        }
        // beta = sqrt(w.w) ; q_{i+1} = w / beta ; T(:, i)     :58-62 / :78-82
        ScaleOp sc{w, col(i + 1), &st->lz[2], &st->lz[0], T + 3 * (size_t)(i - 1), &st->lz[1], 0.0};
        SIGB_CHECK(launch_ew(sc, nr));
    }
    // T(2, n) = q_n . (A q_n)                                 :87-88
    DotSpec d;
    d.ndot = 1; d.u = col(n); d.out[0] = &T[3 * (size_t)(n - 1) + 1];
    SIGB_CHECK(solver_matvec(A, col(n), w, d, false));
    SIGB_CHECK(dist_allreduce(A, &T[3 * (size_t)(n - 1) + 1], 1));
    return SIGB_OK;
}

// n-step generalized Lanczos for A x = lambda B x (eigensolver.f90:95-155) on
// device arrays.  `call B%solve(w, v)` (:134) runs the attached solver -- the
// device CG / PCG -- with w (= A q_i) as the initial guess, exactly like the
// reference's facade (linear_operator_interface.f90:213-233).  Only z_{i-1},
// z_i, z_{i+1} are live, so three rotating buffers replace the reference's
// z(:, 0:n).  zbuf: 3*nr doubles, w and v: nr each.
int generalized_lanczos_dev(sigb_matrix_t A, sigb_matrix_t B, sigb_solver_t bs, sigb_solver_t bpc, int32_t n,
                            const double *q1, uint64_t seed, int64_t row_offset, double *T, double *Q, double *w,
                            double *v, double *zbuf, KState *st)
{
    const int64_t nr = A->nrow;
    cudaStream_t stream = ctx().stream;
    SIGB_CUDA(cudaMemsetAsync(T, 0, sizeof(double) * 3 * (size_t)n, stream));
    SIGB_CUDA(cudaMemsetAsync(Q, 0, sizeof(double) * (size_t)nr * n, stream));
    SIGB_CUDA(cudaMemsetAsync(zbuf, 0, sizeof(double) * 3 * (size_t)nr, stream));
    SIGB_CUDA(cudaMemsetAsync(st, 0, sizeof(KState), stream));     // alpha = beta = 0 (:125-126)
    auto col = [&](int c) { return Q + (size_t)(c - 1) * nr; };    // 1-based column
    auto zc = [&](int c) { return zbuf + (size_t)(((c % 3) + 3) % 3) * nr; };   // z(:, c), c = 0..n

    if (q1) {
        SIGB_CUDA(cudaMemcpyAsync(col(1), q1, sizeof(double) * (size_t)nr, cudaMemcpyDeviceToDevice, stream));
    } else {
        RandomOp rnd{col(1), seed, row_offset};
        SIGB_CHECK(launch_ew(rnd, nr));
    }
    {   // w = B q1 ; q1 = q1 / sqrt(w.q1) ; z(:,1) = B q1           :121-123
        DotSpec d;
        d.ndot = 1; d.u = col(1); d.out[0] = &st->lz[2];
        SIGB_CHECK(solver_matvec(B, col(1), w, d, false));
        SIGB_CHECK(dist_allreduce(B, &st->lz[2], 1));
        ScaleOp sc{col(1), col(1), &st->lz[2], nullptr, nullptr, nullptr, 0.0};
        SIGB_CHECK(launch_ew(sc, nr));
        DotSpec none;
        SIGB_CHECK(solver_matvec(B, col(1), zc(1), none, false));
    }
    DotSpec none;
    for (int i = 1; i <= n - 1; i++) {
        SIGB_CHECK(solver_matvec(A, col(i), w, none, false));                       // :129
        GenLanczosVOp vop{w, zc(i - 1), col(i), v, &st->lz[1], &st->lz[0], 0.0};      // :130-131
        SIGB_CHECK(launch_ew(vop, nr));
        SIGB_CHECK(dist_allreduce(A, &st->lz[0], 1));
        GenLanczosAxpyOp ax{v, zc(i), &st->lz[0], 0.0};                               // :132
        SIGB_CHECK(launch_ew(ax, nr));
        SIGB_CHECK(sigb_solver_solve_dev(bs, B, w, v, bpc));                          // :134
        DotOp dop{w, v, &st->lz[2]};                                                  // :136
        SIGB_CHECK(launch_ew(dop, nr));
        SIGB_CHECK(dist_allreduce(A, &st->lz[2], 1));
        GenLanczosScaleOp sc{w, v, col(i + 1), zc(i + 1), &st->lz[2], &st->lz[0], T + 3 * (size_t)(i - 1),
                             &st->lz[1], 0.0};                                         // :137-142
        SIGB_CHECK(launch_ew(sc, nr));
    }
    // v = A q_n - beta z(:,n) ; T(2,n) = q_n . v                                    :145-147
    SIGB_CHECK(solver_matvec(A, col(n), w, none, false));
    GenLanczosVOp vop{w, zc(n), col(n), v, &st->lz[1], &T[3 * (size_t)(n - 1) + 1], 0.0};
    SIGB_CHECK(launch_ew(vop, nr));
    SIGB_CHECK(dist_allreduce(A, &T[3 * (size_t)(n - 1) + 1], 1));
    return SIGB_OK;
}

// Symmetric tridiagonal eigen-solve standing in for LAPACK dstev('V')
// (eigensolver.f90:174; LAPACK is not vendored by the reference).  Implicit QL
// with Wilkinson shifts; eigenvalues ascending, eigenvectors in the columns of
// Z (column-major n x n).  Runs on the host: n is the number of Lanczos steps.
int tridiag_eig_host(int n, double *d, double *e, double *Z)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) Z[(size_t)j * n + i] = (i == j) ? 1.0 : 0.0;
    if (n == 1) return 0;
    e[n - 1] = 0.0;
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < n; l++) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; m++) {
                const double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= eps * dd) break;
            }
            if (m != l) {
                if (iter++ == 60) return 1;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + copysign(r, g));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; i--) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    r = hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    for (int k = 0; k < n; k++) {
                        double *zi1 = Z + (size_t)(i + 1) * n + k, *zi = Z + (size_t)i * n + k;
                        f = *zi1;
                        *zi1 = s * (*zi) + c * f;
                        *zi = c * (*zi) - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    for (int i = 0; i < n - 1; i++) {
        int k = i;
        double p = d[i];
        for (int j = i + 1; j < n; j++)
            if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i];
            d[i] = p;
            for (int j = 0; j < n; j++) {
                const double tmp = Z[(size_t)i * n + j];
                Z[(size_t)i * n + j] = Z[(size_t)k * n + j];
                Z[(size_t)k * n + j] = tmp;
            }
        }
    }
    return 0;
}

size_t kstate_bytes() { return sizeof(KState); }

// A: the operator the vectors belong to -- on a row-sharded operator the first row of V lives on the rank
// that owns global row 1; the other ranks contribute zeros and the n entries are summed across the ranks
// (exact: x + 0), so that every rank scales by the same signs.
int ritz_vectors_dev(sigb_matrix_t A, double *V, double *V2, const double *Qm_dev, int64_t nr, int32_t n,
                     double *first_row_dev)
{
    cudaStream_t st = ctx().stream;
    const size_t smem = sizeof(double) * (size_t)n * n;
    SIGB_REQUIRE(smem <= 200 * 1024, SIGB_ERR_UNSUPPORTED,
                 "eigensolve: %d Lanczos steps exceed the on-chip Ritz-matrix budget", n);
    SIGB_CUDA(cudaFuncSetAttribute(ritz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ritz_kernel<<<ew_grid(nr), kThreads, smem, st>>>(V, Qm_dev, nr, n, V2);
    SIGB_CUDA(cudaMemcpyAsync(V, V2, sizeof(double) * (size_t)nr * n, cudaMemcpyDeviceToDevice, st));
    count_launch(1);
    if (first_row_dev) {   // sign normalisation of eigensolve (:178-180); generalized_eigensolve has none
        const bool sharded = A != nullptr && A->dist != nullptr;
        if (sharded && (dist_row_offset(A) != 0 || nr == 0)) {
            SIGB_CUDA(cudaMemsetAsync(first_row_dev, 0, sizeof(double) * (size_t)n, st));
        } else {
            grab_first_row_kernel<<<1, 128, 0, st>>>(V, nr, n, first_row_dev);
            count_launch();
        }
        if (sharded)
            for (int32_t c = 0; c < n; c += 3) SIGB_CHECK(dist_allreduce(A, first_row_dev + c, std::min(3, n - c), nullptr));
        sign_kernel<<<ew_grid(nr), kThreads, 0, st>>>(V, nr, n, first_row_dev);
        count_launch(1);
    }
    SIGB_CUDA(cudaGetLastError());
    return SIGB_OK;
}

}  // namespace sigb
