// fb_peak.cu — measures the two issue-rate ceilings this path is bound by, on the device it runs on:
// dense FP32 FMA throughput and SFU (MUFU.RSQ) throughput.  bench.py uses them as roofline
// denominators because MEASURED_PEAKS.json only carries HBM and bf16-tensor peaks and this path
// touches neither ("Not HBM, not tensor cores").
#include <cuda_runtime.h>

#include "../../include/fuzzyblue.h"
#include "fb_kernels.h"

namespace fb {

template <int ITERS>
__global__ void __launch_bounds__(256) k_peak_fma(float* out, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;   // never true; keeps the chains alive
}

__device__ __forceinline__ float rsq(float x) {
    float y;
    asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int ITERS>
__global__ void __launch_bounds__(256) k_peak_sfu(float* out, float seed) {
    float x0 = seed + threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) {
        x0 = rsq(x0); x1 = rsq(x1); x2 = rsq(x2); x3 = rsq(x3);
        x4 = rsq(x4); x5 = rsq(x5); x6 = rsq(x6); x7 = rsq(x7);
    }
    float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;
}

cudaError_t measure_peaks(int sm_count, double* fma_tflops, double* sfu_gops) {
    constexpr int ITERS = 8192;
    float* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 64);
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int blocks = sm_count * 8, threads = 256;
    double best_fma = 0, best_sfu = 0;
    for (int rep = 0; rep < 6; ++rep) {
        float ms = 0;
        cudaEventRecord(t0);
        k_peak_fma<ITERS><<<blocks, threads>>>(d, 1.0000001f, 1e-9f);
        cudaEventRecord(t1);
        cudaEventSynchronize(t1);
        cudaEventElapsedTime(&ms, t0, t1);
        double fl = (double)blocks * threads * ITERS * 8 * 2;
        if (rep && ms > 0 && fl / (ms * 1e-3) / 1e12 > best_fma) best_fma = fl / (ms * 1e-3) / 1e12;
        cudaEventRecord(t0);
        k_peak_sfu<ITERS><<<blocks, threads>>>(d, 1.5f);
        cudaEventRecord(t1);
        cudaEventSynchronize(t1);
        cudaEventElapsedTime(&ms, t0, t1);
        double ops = (double)blocks * threads * ITERS * 8;
        if (rep && ms > 0 && ops / (ms * 1e-3) / 1e9 > best_sfu) best_sfu = ops / (ms * 1e-3) / 1e9;
    }
    e = cudaDeviceSynchronize();
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(d);
    if (fma_tflops) *fma_tflops = best_fma;
    if (sfu_gops) *sfu_gops = best_sfu;
    return e;
}

}  // namespace fb
