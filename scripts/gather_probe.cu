// gather_probe.cu -- what the memory system gives a kernel that does NOTHING but the x gathers of
// an irregular SpMV: sum_k x[node[k]] with uniformly random node, 8-byte elements.  The ceiling the
// CSR kernel's gather pass is measured against on the Erdos-Renyi operators.
//
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -o scripts/_build/gather_probe scripts/gather_probe.cu
//   scripts/_build/gather_probe            # JSON lines
//
// Variants: load flavour (plain / read-only / L2-only), gathers in flight per thread (U), resident
// CTAs per SM, and HALF = every load instruction carries 16 active lanes instead of 32 (does the
// replay of one divergent instruction cost more than two half-populated ones?).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("cuda error %s at line %d\n", cudaGetErrorString(e_), __LINE__); std::exit(1); } } while (0)

template <int FLAVOUR>
__device__ __forceinline__ double ld(const double *p)
{
    if (FLAVOUR == 0) return *p;
    if (FLAVOUR == 1) return __ldg(p);
    if (FLAVOUR == 2) return __ldcg(p);
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

template <int FLAVOUR, int U, bool HALF, int MINB>
__global__ void __launch_bounds__(256, MINB) gather_kernel(const int32_t *__restrict__ node, const double *x, long nnz, double *out)
{
    double s = 0.0;
    const long stride = (long)gridDim.x * 256 * U;
    const bool lo = (threadIdx.x & 16) == 0;
    for (long base = (long)blockIdx.x * 256 * U; base < nnz; base += stride) {
        int c[U];
        double v[U];
#pragma unroll
        for (int i = 0; i < U; i++) {
            const long k = base + i * 256 + threadIdx.x;
            c[i] = (k < nnz) ? __ldcs(node + k) : 0;
        }
        if (!HALF) {
#pragma unroll
            for (int i = 0; i < U; i++) v[i] = ld<FLAVOUR>(x + c[i]);
        } else {
#pragma unroll
            for (int i = 0; i < U; i++) {
                v[i] = 0.0;
                if (lo) v[i] = ld<FLAVOUR>(x + c[i]);
                if (!lo) v[i] = ld<FLAVOUR>(x + c[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < U; i++) s += v[i];
    }
    // keep the sum alive
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s == 1.2345e-300) out[0] = s;
}

template <int FLAVOUR, int U, bool HALF, int MINB = 4>
static void run(const char *name, const int32_t *node, const double *x, long nnz, long n, int ctas_per_sm, double *out,
                int smem_kb = -1)
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    // smem_kb < 0: residency is capped with dynamic shared memory, 220 KB / ctas_per_sm each (which also
    // shrinks the L1: the unified array is 256 KB).  smem_kb >= 0: that much per CTA, residency set by the
    // grid alone (one wave of sms * ctas_per_sm CTAs) -- the sweep that separates L1 size from warp count.
    const int smem = smem_kb >= 0 ? smem_kb * 1024 : (ctas_per_sm >= 8 ? 0 : (220 * 1024) / ctas_per_sm);
    CK(cudaFuncSetAttribute(gather_kernel<FLAVOUR, U, HALF, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; w++) gather_kernel<FLAVOUR, U, HALF, MINB><<<grid, 256, smem>>>(node, x, nnz, out);
    CK(cudaDeviceSynchronize());
    const int reps = 5;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) gather_kernel<FLAVOUR, U, HALF, MINB><<<grid, 256, smem>>>(node, x, nnz, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1000.0 / reps;
    std::printf("{\"probe\": \"gather\", \"load\": \"%s\", \"in_flight_per_thread\": %d, \"half_populated\": %s, \"ctas_per_sm\": %d, \"smem_kb_per_cta\": %.1f, "
                "\"x_mb\": %.0f, \"gathers\": %ld, \"us\": %.1f, \"g_gathers_per_s\": %.1f, \"cycles_per_gather_per_sm_at_1965mhz\": %.2f}\n",
                name, U, HALF ? "true" : "false", ctas_per_sm, smem / 1024.0, n * 8.0 / 1e6, nnz, us, nnz / us / 1e3,
                us * 1965.0 * sms / nnz);
    std::fflush(stdout);
}

int main(int argc, char **argv)
{
    const bool sweep = argc > 1 && std::string(argv[1]) == "sweep";   // L1-size sweep instead of the flavour table
    const long nnz = 1L << 26;  // 67 M gathers
    const long sizes[2] = {2000000L, 20000000L};
    double *out;
    CK(cudaMalloc(&out, 8));
    for (int si = 0; si < 2; si++) {
        const long n = sizes[si];
        std::vector<int32_t> h(nnz);
        uint64_t st = 0x9E3779B97F4A7C15ull + si;
        for (long k = 0; k < nnz; k++) {
            st ^= st << 13; st ^= st >> 7; st ^= st << 17;
            h[k] = (int32_t)(st % (uint64_t)n);
        }
        int32_t *node;
        double *x;
        CK(cudaMalloc(&node, nnz * 4));
        CK(cudaMalloc(&x, n * 8));
        CK(cudaMemcpy(node, h.data(), nnz * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(x, 0, n * 8));
        if (sweep) {
            const int kb8[] = {0, 4, 8, 12, 16, 20, 24, 27};
            for (int kb : kb8) run<1, 8, false, 8>("ldg.nc", node, x, nnz, n, 8, out, kb);
            const int kb4[] = {0, 8, 16, 24, 32, 40, 48, 54};
            for (int kb : kb4) run<1, 8, false, 8>("ldg.nc", node, x, nnz, n, 4, out, kb);
            const int kb6[] = {0, 8, 16, 21, 26, 32, 36};
            for (int kb : kb6) run<1, 8, false, 8>("ldg.nc", node, x, nnz, n, 6, out, kb);
            for (int kb : kb8) run<2, 8, false, 8>("ldcg", node, x, nnz, n, 8, out, kb);
            CK(cudaFree(node));
            CK(cudaFree(x));
            continue;
        }
        run<0, 8, false>("plain", node, x, nnz, n, 4, out);
        run<1, 8, false>("ldg.nc", node, x, nnz, n, 4, out);
        run<2, 8, false>("ldcg", node, x, nnz, n, 4, out);
        run<3, 8, false>("nc.no_allocate", node, x, nnz, n, 4, out);
        run<1, 4, false, 8>("ldg.nc", node, x, nnz, n, 8, out);
        run<1, 8, false, 8>("ldg.nc", node, x, nnz, n, 8, out);
        run<1, 16, false, 8>("ldg.nc", node, x, nnz, n, 8, out);
        run<1, 16, false>("ldg.nc", node, x, nnz, n, 4, out);
        run<1, 8, false>("ldg.nc", node, x, nnz, n, 2, out);
        run<1, 8, false>("ldg.nc", node, x, nnz, n, 1, out);
        run<1, 8, true>("ldg.nc", node, x, nnz, n, 4, out);
        run<2, 8, true, 8>("ldcg", node, x, nnz, n, 8, out);
        run<2, 16, false, 4>("ldcg", node, x, nnz, n, 4, out);
        CK(cudaFree(node));
        CK(cudaFree(x));
    }
    return 0;
}
