// Microbenchmark: shared-memory wavefronts of LDS.32/64/128 under broadcast patterns (run under ncu).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int WIDTH, int PATTERN>
__global__ void k(float* out, int iters) {
    __shared__ __align__(16) float s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i;
    __syncthreads();
    int lane = threadIdx.x & 31;
    // PATTERN 0: all lanes same address; 1: 8 distinct 16B knots within one 128B row (lane%8 .. pseudo-random);
    // 2: each lane distinct (lane*WIDTH words); 3: 3 distinct knots per quarter-warp (typical)
    int idx;
    if (PATTERN == 0) idx = 0;
    else if (PATTERN == 1) idx = ((lane * 5) & 7) * 4;
    else if (PATTERN == 2) idx = lane * WIDTH;
    else idx = ((lane >> 2) % 3) * 4;
    float acc = 0;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(s) + idx * 4;
    for (int i = 0; i < iters; ++i) {
        uint32_t a = base + ((i & 7) << 9);
        if (WIDTH == 4) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); acc += v.x + v.y + v.z + v.w; }
        if (WIDTH == 2) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); acc += v.x + v.y; }
        if (WIDTH == 1) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); acc += v; }
    }
    if (acc == 12345.f) out[0] = acc;
}
template <int W, int P> void run(const char* name, float* d) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<W, P><<<148 * 4, 256>>>(d, 1000);
    cudaEventRecord(a); k<W, P><<<148 * 4, 256>>>(d, 100000); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // per SM: 4 blocks * 8 warps * 100000 loads
    double cyc_per_warp_load = ms * 1e-3 * 1.965e9 / (4.0 * 8 * 100000);
    printf("%-28s %.3f ms  ~%.2f cycles per warp-load per SM\n", name, ms, cyc_per_warp_load);
}
// The density kernel's access pattern: lane = phi sample, entry k(lane) = floor(c + a cos(phi)) -- neighbouring lanes mostly
// share an entry, a warp spans 2-3 entries.  LAYOUT 0: entries of 48 bytes read by three LDS.128 (array of structures);
// 1: the same words as three 128-byte planes of eight 16-byte chunks (structure of arrays); 2: entries of 24 bytes read by
// three LDS.64; 3: three 64-byte planes of eight 8-byte chunks.
template <int LAYOUT>
__global__ void kd(float* out, int iters, float amp) {
    __shared__ __align__(128) float s[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int k = (int)floorf(3.5f + amp * cosf(6.2831853f * (lane + 0.5f) / 32.f));
    float acc = 0;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(s);
    for (int i = 0; i < iters; ++i) {
        const uint32_t row = base + ((i & 7) * (LAYOUT < 2 ? 384 : 192));
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            if (LAYOUT == 0) { float4 v; uint32_t a = row + k * 48 + p * 16; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); acc += v.x + v.y + v.z + v.w; }
            if (LAYOUT == 1) { float4 v; uint32_t a = row + p * 128 + k * 16; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); acc += v.x + v.y + v.z + v.w; }
            if (LAYOUT == 2) { float2 v; uint32_t a = row + k * 24 + p * 8; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); acc += v.x + v.y; }
            if (LAYOUT == 3) { float2 v; uint32_t a = row + p * 64 + k * 8; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); acc += v.x + v.y; }
        }
    }
    if (acc == 12345.f) out[0] = acc;
}
template <int LAYOUT> void rund(const char* name, float* d, float amp) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kd<LAYOUT><<<148 * 4, 256>>>(d, 1000, amp);
    cudaEventRecord(a); kd<LAYOUT><<<148 * 4, 256>>>(d, 40000, amp); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-44s amp %.1f  %.3f ms  ~%.2f cycles per warp-load per SM\n", name, amp, ms, ms * 1e-3 * 1.965e9 / (4.0 * 8 * 40000 * 3));
}
int main() {
    { float* d; cudaMalloc(&d, 64);
      for (float amp : {0.2f, 0.8f, 1.4f, 3.4f}) {
          rund<0>("density pattern, 48 B entries, LDS.128", d, amp); rund<1>("density pattern, 128 B planes, LDS.128", d, amp);
          rund<2>("density pattern, 24 B entries, LDS.64", d, amp); rund<3>("density pattern, 64 B planes, LDS.64", d, amp);
      } }
    float* d; cudaMalloc(&d, 64);
    run<4, 0>("LDS.128 uniform", d); run<4, 1>("LDS.128 8 knots in a row", d); run<4, 3>("LDS.128 3 knots", d); run<4, 2>("LDS.128 distinct", d);
    run<2, 0>("LDS.64 uniform", d); run<2, 1>("LDS.64 8 knots in a row", d); run<2, 2>("LDS.64 distinct", d);
    run<1, 0>("LDS.32 uniform", d); run<1, 1>("LDS.32 8 knots in a row", d); run<1, 2>("LDS.32 distinct", d);
    return 0;
}
