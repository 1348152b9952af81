// Microbenchmark: issue cost of the packed fp32 instructions of sm_100 (FFMA2 / FMUL2 / FADD2, PTX fma.rn.f32x2 ...)
// against their scalar forms, alone and interleaved with integer work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float mul1(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) { unsigned r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// MODE 0: 8 scalar FFMA chains; 1: 8 FFMA2 chains; 2: 8 FFMA2 + 8 IADD; 3: 8 FFMA + 8 IADD; 4: 8 FMUL2; 5: 8 FADD2; 6: 8 FMUL;
// 7: 4 FFMA2 + 8 FFMA; 8: 16 FFMA + 8 IADD (the scalar form of mode 2)
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    float f[8], g[8]; u64 p[8]; unsigned x[8];
    for (int j = 0; j < 8; ++j) { f[j] = seed + j + threadIdx.x; g[j] = f[j] * 0.25f; p[j] = ((u64)__float_as_uint(f[j]) << 32) | __float_as_uint(f[j] * 0.5f); x[j] = threadIdx.x * 7 + j; }
    const float a = 0.999f, b = 0.001f;
    const u64 a2 = ((u64)__float_as_uint(a) << 32) | __float_as_uint(a), b2 = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0 || MODE == 3 || MODE == 7 || MODE == 8) f[j] = fma1(f[j], a, b);
            if (MODE == 8) g[j] = fma1(g[j], a, b);
            if (MODE == 1 || MODE == 2) p[j] = fma2(p[j], a2, b2);
            if (MODE == 7 && j < 4) p[j] = fma2(p[j], a2, b2);
            if (MODE == 2 || MODE == 3 || MODE == 8) x[j] = lop(x[j], x[(j + 1) & 7]);
            if (MODE == 4) p[j] = mul2(p[j], a2);
            if (MODE == 5) p[j] = add2(p[j], b2);
            if (MODE == 6) f[j] = mul1(f[j], a);
        }
    }
    float acc = 0; for (int j = 0; j < 8; ++j) acc += f[j] + g[j] + __uint_as_float((unsigned)p[j]) + __uint_as_float((unsigned)(p[j] >> 32)) + (float)x[j];
    if (acc == 12345.f) out[0] = acc;
}
template <int MODE> void run(const char* name, float* d, int per_iter) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 200000;
    k<MODE><<<148 * 4, 256>>>(d, 1000, 1.f);
    cudaEventRecord(a); k<MODE><<<148 * 4, 256>>>(d, iters, 1.f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // per SM sub-partition: 4 blocks * 8 warps / 4 SMSPs = 8 warps, each `per_iter` instructions per iteration
    const double cyc = ms * 1e-3 * 1.965e9, inst = 8.0 * per_iter * iters;
    printf("%-34s %.3f ms  %.3f cycles per warp-instruction per SMSP (at 1965 MHz)\n", name, ms, cyc / inst);
}
int main() {
    float* d; cudaMalloc(&d, 64);
    run<0>("8 FFMA", d, 8); run<1>("8 FFMA2", d, 8); run<2>("8 FFMA2 + 8 IADD", d, 16); run<3>("8 FFMA + 8 IADD", d, 16);
    run<8>("16 FFMA + 8 IADD", d, 24);
    run<4>("8 FMUL2", d, 8); run<5>("8 FADD2", d, 8); run<6>("8 FMUL", d, 8); run<7>("8 FFMA + 4 FFMA2", d, 12);
    return 0;
}
