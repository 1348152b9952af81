// Microbenchmark: which lane -> address patterns let the shared-memory pipe of sm_100 serve a 64- / 128-bit warp load in
// fewer wavefronts.  k(lane) comes from a table, entry k sits at k * STRIDE bytes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__constant__ int KTAB[32];
template <int WIDTH, int STRIDE>
__global__ void k(float* out, int iters) {
    __shared__ __align__(128) float s[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(s) + KTAB[lane] * STRIDE;
    float acc = 0;
    for (int i = 0; i < iters; ++i) {
        const uint32_t a = base + ((i & 7) * 3072);
        if (WIDTH == 4) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); acc += v.x + v.y + v.z + v.w; }
        if (WIDTH == 2) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); acc += v.x + v.y; }
        if (WIDTH == 1) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); acc += v; }
    }
    if (acc == 12345.f) out[0] = acc;
}
template <int W, int STRIDE> double run(float* d) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<W, STRIDE><<<148 * 4, 256>>>(d, 1000);
    cudaEventRecord(a); k<W, STRIDE><<<148 * 4, 256>>>(d, 50000); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms * 1e-3 * 1.965e9 / (4.0 * 8 * 50000);
}
int main() {
    float* d; cudaMalloc(&d, 64);
    struct Pat { const char* name; int k[32]; };
    Pat pats[40]; int np = 0;
    auto add = [&](const char* n, auto f) { pats[np].name = n; for (int l = 0; l < 32; ++l) pats[np].k[l] = f(l); ++np; };
    add("all lanes one entry", [](int l) { return 0; });
    add("pairs share, 16 entries (l>>1)", [](int l) { return l >> 1; });
    add("pairs straddle ((l+1)>>1)", [](int l) { return (l + 1) >> 1; });
    add("quads share, 8 entries (l>>2)", [](int l) { return l >> 2; });
    add("quads straddle, pairs share ((l+2)>>2)", [](int l) { return (l + 2) >> 2; });
    add("octets share, 4 entries (l>>3)", [](int l) { return l >> 3; });
    add("halves share, 2 entries (l>>4)", [](int l) { return l >> 4; });
    add("alternating 2 entries (l&1)", [](int l) { return l & 1; });
    add("pairs alternate 2 entries ((l>>1)&1)", [](int l) { return (l >> 1) & 1; });
    add("3 entries, even boundaries 6,10,22,26", [](int l) { return (l < 6 || l >= 26) ? 2 : (l < 10 || l >= 22) ? 1 : 0; });
    add("3 entries, odd boundaries 5,11,21,27", [](int l) { return (l < 5 || l >= 27) ? 2 : (l < 11 || l >= 21) ? 1 : 0; });
    add("3 entries, one odd boundary (5), rest even", [](int l) { return (l < 5 || l >= 26) ? 2 : (l < 10 || l >= 22) ? 1 : 0; });
    add("2 entries, boundary at 16", [](int l) { return l < 16 ? 0 : 1; });
    add("2 entries, boundary at 15", [](int l) { return l < 15 ? 0 : 1; });
    add("2 entries, boundary at 7", [](int l) { return l < 7 ? 0 : 1; });
    add("2 entries, boundary at 8", [](int l) { return l < 8 ? 0 : 1; });
    add("pairs (0,1) x8 then (2,3) x8", [](int l) { return (l & 1) + 2 * (l >> 4); });
    add("pairs (0,1),(2,3) alternating", [](int l) { return (l & 1) + 2 * ((l >> 1) & 1); });
    add("first half l&1, second half one entry", [](int l) { return l < 16 ? (l & 1) : 0; });
    add("pairs (i, i+8): 16 entries, all pairs differ", [](int l) { return (l >> 1) + 8 * (l & 1); });
    add("pairs (A,B) except one pair (A,A)", [](int l) { return l == 5 ? 0 : (l & 1); });
    add("pairs (A,B) except one pair (B,A)", [](int l) { return (l >> 1) == 2 ? 1 - (l & 1) : (l & 1); });
    add("even lanes entry 0, odd lanes l>>3", [](int l) { return (l & 1) ? 1 + (l >> 3) : 0; });
    add("even lanes l>>3, odd lanes 4+(l>>3)", [](int l) { return (l & 1) ? 4 + (l >> 3) : (l >> 3); });
    add("lanes l and l+16 share, 16 entries (l&15)", [](int l) { return l & 15; });
    add("l&3 (4 entries, period 4)", [](int l) { return l & 3; });
    add("l&7 (8 entries, period 8)", [](int l) { return l & 7; });
    add("2 entries, boundary at 1", [](int l) { return l < 1 ? 0 : 1; });
    add("2 entries, boundary at 2", [](int l) { return l < 2 ? 0 : 1; });
    add("2 entries, lanes 3 and 4 only differ", [](int l) { return (l == 3 || l == 4) ? 1 : 0; });
    add("2 entries, lanes 4 and 5 only differ", [](int l) { return (l == 4 || l == 5) ? 1 : 0; });
    printf("%-46s %10s %10s %10s %10s %10s\n", "cycles per warp-load per SM", "128b/48B", "128b/16B", "64b/24B", "64b/8B", "32b/4B");
    for (int i = 0; i < np; ++i) {
        cudaMemcpyToSymbol(KTAB, pats[i].k, sizeof(int) * 32);
        printf("%-46s %10.2f %10.2f %10.2f %10.2f %10.2f\n", pats[i].name, run<4, 48>(d), run<4, 16>(d), run<2, 24>(d), run<2, 8>(d), run<1, 4>(d));
    }
    return 0;
}
