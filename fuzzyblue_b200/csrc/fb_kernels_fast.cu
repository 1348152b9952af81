// fb_kernels_fast.cu — FB_KERNELS_FAST: the product kernels for sm_100a.
//
// Same shaders, same quadrature nodes, same image formats as fb_kernels_ref.cu; what changes is
// *where* each quantity is computed.  Rule of the house (DESIGN.md "Numerics"):
//   * every ill-conditioned scalar (ray lengths, discriminants, r*r - bottom*bottom, texel -> (r,mu,mu_s,nu))
//     is evaluated in fb::xf exactly as the GLSL writes it — those cancel catastrophically and a
//     different rounding moves a texel by per cents;
//   * anything that is invariant across a block (per-r, per-(r,theta), per-(r,mu) quantities, whole
//     interpolated table rows) is computed once, bit-identically, and shared through shared memory;
//   * only convex interpolation, phase functions and the accumulation of positive terms run in plain
//     fp32 with FMA contraction and a different summation order (relative effect ~1e-7).
#include "fb_kernels.h"

#include <type_traits>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include "fb_shader_math.cuh"

namespace fb {
namespace fast {

typedef xf F;
typedef V3<xf> V;

static inline Tex2 tex2(const float4* p, int w, int h) { Tex2 t; t.p = p; t.w = w; t.h = h; return t; }
static inline Tex3 tex3(const uint2* p, int w, int h, int d) { Tex3 t; t.p = p; t.w = w; t.h = h; t.d = d; return t; }
static inline Tex2 texT(const LaunchCtx& c) { return tex2(c.img.transmittance, c.P.transmittance_mu_size, c.P.transmittance_r_size); }
static inline Tex3 texS(const LaunchCtx& c, const uint2* p) {
    return tex3(p, c.P.scattering_nu_size * c.P.scattering_mu_s_size, c.P.scattering_mu_size, c.P.scattering_r_size);
}

// ---------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + 1-D TMA bulk copy (global -> shared), sm_90+/sm_100a
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// scattering_density.comp — the dominant kernel (~90 % of the reference's work).
//
// For a texel (r, mu, mu_s, nu) the shader sums, over 16 theta x 32 phi directions w_i,
//     L(r, cos theta, mu_s, nu1 = w_s . w_i) * Phase(nu2 = w . w_i) * dw
// and inside that loop re-derives, per sample, quantities that do not depend on the sample:
//   * the 4-D look-up coordinates u_r(r), u_mu(r, theta), u_mu_s(mu_s)  — only u_nu(nu1) varies;
//   * the phase/extinction weight — a function of (r, mu, theta, phi) only, shared by the whole
//     (mu_s, nu) plane of the table;
//   * the ground term's transmittance, distance and normal — functions of (r, theta[, phi]).
// Layout here:
//   k_density_prep   once per order, per r: for every (theta_l, mu_s, nu-slice k) the trilinear tap the shader
//                    would take (exact, as written) -> a [r][tile][l][mu_s][k] table in global scratch (2 MiB at
//                    default dims; x2 at order 2 where Rayleigh and Mie are looked up separately), plus the
//                    per-(r, theta) ground constants.
//   k_density_main   one CTA per (r, mu, mu_s-tile): the CTA's 64 / 128 KiB slice of that table arrives by ONE
//                    TMA bulk copy (cp.async.bulk + mbarrier) while the CTA computes texel geometry and the 512
//                    phase weights.  Then one WARP per texel, lane = phi index (32 lanes = the 32 phi samples),
//                    unrolled loop over the 16 theta: per sample 2 LDS.128 from one 128-byte table row
//                    (one wavefront: every lane reads the same (l, mu_s) row), a nu-lerp and 3 FMAs against the
//                    lane's register-resident weights; the phi-sum is a 5-step __shfl_xor tree.
// ---------------------------------------------------------------------------------------------
constexpr int DL = 16;   // theta samples, scattering_density.comp:34
// cos / sin of theta_l = (l + 0.5) * pi / 16 as literals: inside the unrolled theta loop they become FFMA immediates
// (only used where a 1-ulp difference from the host cosf() table is harmless; the exact paths take fb::Trig).
__device__ constexpr float CT16[DL] = {0.99518472f, 0.956940353f, 0.881921291f, 0.773010433f, 0.634393275f, 0.471396744f, 0.290284663f, 0.0980171412f, -0.0980171412f, -0.290284663f, -0.471396744f, -0.634393275f, -0.773010433f, -0.881921291f, -0.956940353f, -0.99518472f};
__device__ constexpr float ST16[DL] = {0.0980171412f, 0.290284663f, 0.471396744f, 0.634393275f, 0.773010433f, 0.881921291f, 0.956940353f, 0.99518472f, 0.99518472f, 0.956940353f, 0.881921291f, 0.773010433f, 0.634393275f, 0.471396744f, 0.290284663f, 0.0980171412f};

// Texels per CTA are a compile-time constant so that every table offset inside the unrolled theta loop is an
// immediate: 256 texels at order >= 3 (one look-up table), 128 at order 2 (Rayleigh + Mie tables): 96 KiB either way.
//
// Table entries are (intercept, slope) pairs per channel, so the nu interpolation is one FFMA per channel and
// a sample reads exactly the 6 (12 at order 2) words it needs: the shared-memory data pipe is the binding unit of this
// kernel and an LDS costs one wavefront per 4 bytes per lane (tools/lds_bench.cu; the cheaper path the pipe has for
// quads of lanes reading x,x,y,y or x,y,x,y -- tools/lds_pair_bench.cu -- is out of reach of a lane = phi staircase).
// The red and green channels travel as the two halves of one 64-bit operand (packed fp32 pairs, FFMA2: see P2 below):
//   order >= 3 entry, 24 B:  [i.r i.g | s.r s.g | i.b s.b]                                   3 x LDS.64
//   order 2    entry, 48 B:  [Ri.r Ri.g Rs.r Rs.g | Mi.r Mi.g Ms.r Ms.g | Ri.b Rs.b Mi.b Ms.b]  3 x LDS.128
// (i = intercept, s = slope of the nu segment; R = Rayleigh, M = Mie table)
#ifndef FB_PAIR_O2
#define FB_PAIR_O2 1   // round 1 measured the order-2 pair body slower; with the packed (red, green) channels it wins: 936 -> 904 us (64-bit / 128-bit loads only, see PAIRED)
#endif
#ifndef FB_PAIR_O3
#define FB_PAIR_O3 1
#endif
template <bool ORDER2> struct DensityCfg {
    static constexpr bool PAIRED = ORDER2 ? FB_PAIR_O2 != 0 : FB_PAIR_O3 != 0;   // mirror-sample body for coplanar texels
    static constexpr int ENT = ORDER2 ? 12 : 6;       // floats per table entry
    static constexpr int T = ORDER2 ? 128 : 256;      // texels per CTA = ms_tile * nu
    static constexpr int NWARPS = 8;
};
struct DensityDims {
    int nu, ms_tile, tiles;             // ms_tile = T / nu mu_s columns per CTA (any nu: 3, 6, 12 ... need not be a power of two)
    uint32_t ms_mul;                    // ceil(2^20 / ms_tile): t / ms_tile == (t * ms_mul) >> 20 for every texel index t < 256
};
// what the restructured density kernels cover: a CTA needs at least one whole mu_s column (nu <= 128 texels at order 2) and
// the eight ground rows must fit beside the table slice.  Anything else runs the one-thread-per-texel transcription, ~30x
// slower -- reported through fb_params_slow_stages / fb_pending_slow_stages, never silently.
bool density_is_fast(const FbParams& P) {
    const int nu = P.scattering_nu_size;
    return nu >= 2 && nu <= 128 && P.irradiance_mu_s_size <= 512;
}
static inline bool density_supported(const FbParams& P) { return density_is_fast(P); }
static inline DensityDims density_dims(const FbParams& P, int T) {
    DensityDims d;
    d.nu = P.scattering_nu_size;
    d.ms_tile = T / d.nu;
    d.ms_mul = d.ms_tile > 0 ? (uint32_t)(((1u << 20) + d.ms_tile - 1) / d.ms_tile) : 0u;   // exact while t * ms_tile < 2^20
    d.tiles = (P.scattering_mu_s_size + d.ms_tile - 1) / d.ms_tile;
    return d;
}
// scratch: table [R][tiles][DL][T][ENT] floats (sized for the order-2 layout, the larger one) + ground-hit flags
// [R][DL] + ground rows [R][DL/2][nE][3] float2: row 0 of delta_irradiance as (value, delta) pairs, multiplied
// by the ground factor of (r, theta_l) for the eight downward theta rows (the others cannot reach the ground)
static inline size_t density_tab_floats(const FbParams& P) {
    DensityDims d = density_dims(P, DensityCfg<true>::T);
    return (size_t)P.scattering_r_size * d.tiles * DL * DensityCfg<true>::T * DensityCfg<true>::ENT;
}
static inline size_t density_grow_floats(const FbParams& P) { return (size_t)P.scattering_r_size * (DL / 2) * P.irradiance_mu_s_size * 6; }

size_t scratch_bytes(const FbParams& P) {
    if (!density_supported(P)) return 0;
    return (density_tab_floats(P) + density_grow_floats(P) + (size_t)P.scattering_r_size * DL) * sizeof(float);
}

template <bool ORDER2>
__global__ void __launch_bounds__(256) k_density_prep(const __grid_constant__ FbParams P, const __grid_constant__ Trig tg, Tex2 T,
                                                      Tex3 A0, Tex3 A1, DensityDims dd, float* __restrict__ tab,
                                                      float* __restrict__ hit, const float4* __restrict__ dE_row0,
                                                      float2* __restrict__ grow, int r0, int sw32) {
    constexpr int ENT = DensityCfg<ORDER2>::ENT, TT = DensityCfg<ORDER2>::T;
    const int l = blockIdx.y, z = r0 + blockIdx.z;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int W = P.scattering_nu_size * P.scattering_mu_s_size;
    const bool active = e < W;
    const int k = e % dd.nu, ms = active ? e / dd.nu : 0;
    A<F> a(P);
    F r, mu_unused, mu_s, nu_unused;
    bool hits_unused;
    // r and mu_s exactly as the consuming texels derive them (scattering.h:116-137)
    a.TexelToRMuMuSNu((unsigned)ms, 0u, (unsigned)z, r, mu_unused, mu_s, nu_unused, hits_unused);
    const F ct = F(tg.ct16[l]), st = F(tg.st16[l]);
    const bool hits = a.RayIntersectsGround(r, ct);                                   // scattering_density.comp:45-46
    if (blockIdx.x == 0) {   // per-(r, theta) ground terms, scattering_density.comp:50-60, :81-87 (the whole block: r is uniform)
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (hits) {
            F dg = a.DistanceToBottom(r, ct);
            V tgr = a.Transmittance(T, r, ct, dg, true);
            V G = tgr * V(P.ground_albedo) * (F(1.f) / F(FB_PI_F));
            gx = G.x.v; gy = G.y.v; gz = G.z.v;
        }
        if (threadIdx.x == 0) hit[(size_t)z * DL + l] = hits ? 1.f : 0.f;   // the main kernel's ground-row mask
        if (l >= DL / 2) {   // GetIrradiance(bottom, .) only ever reads row 0 of delta_irradiance (irradiance.h:20-30)
            const int nE = P.irradiance_mu_s_size;
            float2* o = grow + ((size_t)z * (DL / 2) + (l - DL / 2)) * nE * 3;
            for (int i = threadIdx.x; i < nE; i += blockDim.x) {
                const float4 v0 = __ldg(dE_row0 + i), v1 = __ldg(dE_row0 + min(i + 1, nE - 1));
                // (intercept, slope) of the segment [i, i + 1): value(te) = intercept + te * slope, no fractional part needed
                const float dx_ = gx * (v1.x - v0.x), dy_ = gy * (v1.y - v0.y), dz_ = gz * (v1.z - v0.z), fi = (float)i;
                o[i * 3] = make_float2(fmaf(-fi, dx_, gx * v0.x), fmaf(-fi, dy_, gy * v0.y));   // (red, green) intercepts
                o[i * 3 + 1] = make_float2(dx_, dy_);                                            // (red, green) slopes
                o[i * 3 + 2] = make_float2(fmaf(-fi, dz_, gz * v0.z), dz_);                      // blue (intercept, slope)
            }
        }
    }
    if (!active) return;
    F uvwz[4];
    a.ScatteringUvwz(r, ct, mu_s, F(0.f), hits, uvwz);                                // scattering.h:144-145
    const F nn = F((float)dd.nu);
    const F ux0 = (F((float)k) + uvwz[1]) / nn;                                       // scattering.h:149, tex_x = k
    const F ux1 = (F((float)k) + F(1.f) + uvwz[1]) / nn;                              // scattering.h:151, tex_x + 1
    const bool last = k + 1 >= dd.nu;                                                 // no knot beyond: delta 0
    const int tile = ms / dd.ms_tile, msl = ms % dd.ms_tile;
    float* o = tab + ((((size_t)z * dd.tiles + tile) * DL + l) * TT + (size_t)msl * dd.nu + k) * ENT;
    const V4<F> s0 = sample<F>(A0, ux0, uvwz[2], uvwz[3]);
    const V4<F> s1 = last ? s0 : sample<F>(A0, ux1, uvwz[2], uvwz[3]);
    // Entries are (intercept, slope) of the segment [k, k + 1) of the nu axis: value(tcx) = intercept + tcx * slope with
    // tcx = u_nu * (nu - 1) the un-floored coordinate, so the consumer needs floor(tcx) for the address only.
    const float fk = (float)k;
    auto seg = [fk](F v0, F v1) { const float d = v1.v - v0.v; return make_float2(fmaf(-fk, d, v0.v), d); };
    if (ORDER2) {
        const V4<F> m0 = sample<F>(A1, ux0, uvwz[2], uvwz[3]);
        const V4<F> m1 = last ? m0 : sample<F>(A1, ux1, uvwz[2], uvwz[3]);
        const float2 rx = seg(s0.x, s1.x), ry = seg(s0.y, s1.y), rz = seg(s0.z, s1.z);
        const float2 mx = seg(m0.x, m1.x), my = seg(m0.y, m1.y), mz = seg(m0.z, m1.z);
        // wide tables (nu > 16): the words of each group of four at position j ^ s, s = (k >> 3) & 3 (tab12<true> reads them so)
        const int sx = sw32 ? (k >> 3) & 3 : 0;
        // words: Rayleigh (red, green) intercepts, slopes | Mie (red, green) intercepts, slopes | blue: Rayleigh, Mie pairs --
        // the red and green channels of a sample are evaluated by one packed instruction (FFMA2, see P2 below)
        const float w12[12] = {rx.x, ry.x, rx.y, ry.y, mx.x, my.x, mx.y, my.y, rz.x, rz.y, mz.x, mz.y};
#pragma unroll
        for (int j = 0; j < 12; ++j) o[(j & ~3) | ((j & 3) ^ sx)] = w12[j];
    } else {
        float2* o2 = reinterpret_cast<float2*>(o);
        // wide tables (nu > 16): entries 16..31, 48..63, ... as (slope, intercept) -- the bank swizzle tab3<true> reads
        const bool swp = sw32 && (k & 16);
        auto put = [swp](float2 v) { return swp ? make_float2(v.y, v.x) : v; };
        const float2 sr = seg(s0.x, s1.x), sg = seg(s0.y, s1.y);
        o2[0] = put(make_float2(sr.x, sg.x));              // (red, green) intercepts
        o2[1] = put(make_float2(sr.y, sg.y));              // (red, green) slopes
        o2[2] = put(seg(s0.z, s1.z));                      // blue (intercept, slope)
    }
}

template <int OFF> __device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(addr), "n"(OFF));
    return v;
}
// Packed fp32 pairs (sm_100: FFMA2, FMUL2, FADD2).  Each half is the scalar IEEE operation, so results keep their bits;
// one warp-instruction produces the red and the green channel (tools/ffma2_bench.cu: an FFMA2 holds the FMA pipe for two
// cycles like two FFMA but takes ONE issue slot).  add2 / mul2 are only ever applied to results of fused multiply-adds
// or feed one as the addend, where ptxas has nothing to contract (it does contract mul.rn.f32x2 + add.rn.f32x2).
struct P2 { unsigned long long v; };
__device__ __forceinline__ P2 pk(float lo, float hi) { P2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(P2 p, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p.v)); }
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) { P2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ P2 mul2(P2 a, P2 b) { P2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2 add2(P2 a, P2 b) { P2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2 sub2(P2 a, P2 b) { P2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2 bc(float f) { return pk(f, f); }
template <int OFF> __device__ __forceinline__ P2 lds64p(uint32_t addr) {
    P2 v;
    asm volatile("ld.shared.b64 %0, [%1+%2];" : "=l"(v.v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ void lds128p(uint32_t addr, P2& lo, P2& hi) {
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2+%3];" : "=l"(lo.v), "=l"(hi.v) : "r"(addr), "n"(OFF));
}
template <int OFF> __device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}
// One (intercept, slope) entry of the order >= 3 table: three 64-bit loads -- or, for tables of more than 16 nu knots
// (SW32), six 32-bit loads from a bank-swizzled entry.  With 64-bit accesses a wavefront serves 16 lanes over 16 bank
// pairs, and entries k and k + 16 of a row (24-byte stride) share a pair: at nu = 32 the phi samples of a warp reach
// across more than 16 segments and 23 % of the wavefronts of the high-resolution launch were replays
// (l1tex__data_bank_conflicts_pipe_lsu_mem_shared 1.2e9 of 5.3e9; 1 % at nu = 8).  k_density_prep therefore stores the
// pairs of entries 16..31 (48..63, ...) as (slope, intercept): read as 32-bit words, entry k + 16 lands on the odd bank
// of the pair entry k uses, and any set of entries 0..31 is conflict-free.  `sw` = 4 for a swapped entry, else 0.
template <bool SW32, int OFF>
__device__ __forceinline__ void tab3(uint32_t addr, uint32_t sw, P2& irg, P2& srg, float2& cb) {   // (r, g) intercepts, slopes; blue pair
    if (!SW32) {
        irg = lds64p<OFF>(addr); srg = lds64p<OFF + 8>(addr); cb = lds64<OFF + 16>(addr);
    } else {
        const uint32_t aI = addr + sw, aS = aI ^ 4u;        // addr is 8-byte aligned
        irg = pk(lds32<OFF>(aI), lds32<OFF>(aS)); srg = pk(lds32<OFF + 8>(aI), lds32<OFF + 8>(aS));
        cb.x = lds32<OFF + 16>(aI); cb.y = lds32<OFF + 16>(aS);
    }
}
__device__ __forceinline__ uint32_t tab_swap(float tm) { return (__float_as_uint(tm) >> 2) & 4u; }   // bit 4 of floor(tcx)
// The order-2 entry (twelve words, 48-byte stride): three 128-bit loads -- or, SW32, twelve 32-bit loads from an entry
// whose words are permuted inside each aligned group of four by XOR with s = (k >> 3) & 3 (k_density_prep writes them
// so): entries k, k + 8, k + 16, k + 24 of a row start on the same bank, and the permutation spreads the same logical
// word of the four over the four banks of its group, so any set of entries 0..31 is conflict-free.
template <bool SW32, int OFF>
__device__ __forceinline__ void tab12(uint32_t addr, float tm, P2& Ri, P2& Rs, P2& Mi, P2& Ms, float4& tb) {
    if (!SW32) {
        lds128p<OFF>(addr, Ri, Rs); lds128p<OFF + 16>(addr, Mi, Ms); tb = lds128<OFF + 32>(addr);
    } else {
        const uint32_t a0 = addr + ((__float_as_uint(tm) >> 1) & 12u);      // addr is 16-byte aligned: word j of a group sits at (4 j) ^ (4 s)
        const uint32_t a1 = a0 ^ 4u, a2 = a0 ^ 8u, a3 = a0 ^ 12u;
        Ri = pk(lds32<OFF>(a0), lds32<OFF>(a1)); Rs = pk(lds32<OFF>(a2), lds32<OFF>(a3));
        Mi = pk(lds32<OFF + 16>(a0), lds32<OFF + 16>(a1)); Ms = pk(lds32<OFF + 16>(a2), lds32<OFF + 16>(a3));
        tb.x = lds32<OFF + 32>(a0); tb.y = lds32<OFF + 32>(a1); tb.z = lds32<OFF + 32>(a2); tb.w = lds32<OFF + 32>(a3);
    }
}

__device__ __forceinline__ float rsqrt_fast(float x) {   // one MUFU.RSQ; callers guarantee a normal, positive x
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One sample of the integrand for the paired body's slow branch (a table knot between two mirror samples): the look-up
// of FB_DENSITY_STEP as a function.  TOFF / GOFF are the immediate offsets of the theta row and of the ground rows.
template <bool ORDER2, bool SW32, int TOFF, int GOFF>
__device__ __forceinline__ void density_tap(uint32_t addr, uint32_t sw, float tmf, float f, float nu1, float kR, float kMR, float g2p1, float m2g, bool gnd,
                                            uint32_t ea, float te, float& Lr, float& Lg, float& Lb) {
    if (ORDER2) {
        P2 Ri, Rs, Mi, Ms; float4 tb;
        tab12<SW32, TOFF>(addr, tmf, Ri, Rs, Mi, Ms, tb);
        const float pr = fmaf(nu1 * kR, nu1, kR);
        const float rs = rsqrt_fast(fmaf(m2g, nu1, g2p1));
        const float pm = pr * kMR * (rs * rs * rs);
        upk(fma2(fma2(bc(f), Ms, Mi), bc(pm), mul2(fma2(bc(f), Rs, Ri), bc(pr))), Lr, Lg);
        Lb = fmaf(fmaf(f, tb.w, tb.z), pm, fmaf(f, tb.y, tb.x) * pr);
    } else {
        P2 irg, srg; float2 cb;
        tab3<SW32, TOFF>(addr, sw, irg, srg, cb);
        upk(fma2(bc(f), srg, irg), Lr, Lg); Lb = fmaf(f, cb.y, cb.x);
    }
    if (gnd) {
        const P2 ei = lds64p<GOFF>(ea), es = lds64p<GOFF + 8>(ea);
        const float2 eb = lds64<GOFF + 16>(ea);
        float gr, gg;
        upk(fma2(bc(te), es, ei), gr, gg);
        Lr += gr; Lg += gg; Lb += fmaf(te, eb.y, eb.x);
    }
}

// Ground normals of the downward theta rows, (n_x, n_z) * scale per (r, l): CTA-uniform values the ground term needs
// per sample.  As a kernel parameter they live in the constant bank and reach the FFMAs through uniform registers
// (ULDC) instead of costing shared-memory wavefronts — the unit that binds this kernel.  One launch covers at most
// GN_MAXR altitude levels.  They feed a look-up coordinate only, so the host evaluates them in double precision.
constexpr int GN_MAXR = 64;
struct GroundNormals { float2 n[GN_MAXR][DL / 2]; };


template <bool ORDER2, bool SW32>
__global__ void __launch_bounds__(256, 2)
k_density_main(const __grid_constant__ FbParams P, const __grid_constant__ Trig tg, DensityDims dd, const float* __restrict__ tabG,
               const float* __restrict__ hitG, const float2* __restrict__ growG, uint2* __restrict__ out, int r0,
               uint32_t magic_tab, uint32_t magic_row, const __grid_constant__ GroundNormals GN
               ) {
    typedef DensityCfg<ORDER2> C;
    constexpr int ENT = C::ENT, TT = C::T, NWARPS = C::NWARPS;
    // order 2 with bank-swizzled tables (32-bit loads) keeps the general body: there the pair body measured slower
    // (high-resolution dims, 16 levels: 31.5 -> 32.7 ms)
    constexpr bool PAIRED = C::PAIRED && !(ORDER2 && SW32);
    constexpr int ENT_B = ENT * 4;                       // bytes per table entry
    constexpr int L_STRIDE = TT * ENT_B;                 // bytes per theta row block
    constexpr int TAB_B = DL * L_STRIDE;                 // 96 KiB
    // shared memory map (bytes): [table TAB_B][geo TT*16][2 mbarriers, per-warp leader counts and follower masks: 128]
    //                            [Wt DL*32*16 during the prologue, then the 8 ground rows (DL/2)*nE*24]
    constexpr int GEO_OFF = TAB_B, BAR_OFF = GEO_OFF + TT * 16, GR_OFF = BAR_OFF + 128;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* geoS = reinterpret_cast<float4*>(smem_raw + GEO_OFF);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + BAR_OFF);
    float4* WtS = reinterpret_cast<float4*>(smem_raw + GR_OFF);
    int* cntS = reinterpret_cast<int*>(smem_raw + BAR_OFF + 16);           // per warp: general | paired << 16 leaders
    uint32_t* folS = reinterpret_cast<uint32_t*>(smem_raw + BAR_OFF + 64);  // per warp: follower mask of its 32 texels
    const int nE = P.irradiance_mu_s_size;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, y = blockIdx.y, z = r0 + blockIdx.z;
    const int W = P.scattering_nu_size * P.scattering_mu_s_size;

    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, TAB_B);
        tma_bulk_g2s(smem_raw, tabG + ((size_t)z * dd.tiles + tile) * DL * TT * ENT, TAB_B, bar);
    }

    A<F> a(P);
    // ---- texel geometry, exact (scattering_density.comp:28-32) ------------------------------------------
    // One evaluation per thread: thread t < TT takes texel t of the tile; r and mu (functions of y, z only) come out of
    // the same call, the remaining threads (and padding texels of a partial tile) evaluate texel 0 of the row for them.
    F r, mu;
    {
        const int t = tid < TT ? tid : 0;
        const int nui = (int)(((uint32_t)t * dd.ms_mul) >> 20), msl = t - nui * dd.ms_tile;
        const int ms = tile * dd.ms_tile + msl;
        const bool valid = tid < TT && nui < dd.nu && ms < P.scattering_mu_s_size;   // nui >= nu: padding of a tile with T % nu != 0
        F mu_s, nu;
        bool hu;
        a.TexelToRMuMuSNu(valid ? (unsigned)(nui * P.scattering_mu_s_size + ms) : 0u, (unsigned)y, (unsigned)z, r, mu, mu_s, nu, hu);
        if (tid < TT) {
            float4 g = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
            if (valid) {
                F ox = f_sqrt(F(1.f) - mu * mu);
                F sx = ox == F(0.f) ? F(0.f) : (nu - mu * mu_s) / ox;
                F sy = f_sqrt(f_max(F(1.f) - sx * sx - mu_s * mu_s, F(0.f)));
                g = make_float4(sx.v, sy.v, mu_s.v, __int_as_float(msl * dd.nu * ENT_B));   // .w: byte offset of the texel's rows
            }
            geoS[tid] = g;
        }
    }
    // ---- the 512 phase weights of this (r, mu) row, once per CTA (scattering_density.comp:93-103) -------
    // Phase functions are smooth and well-conditioned: contracted fp32 with x^-1.5 = rsqrt(x)^3, the form the order-2
    // samples use for nu1 (util.h:26-34).  The density profiles stay exact.
    {
        const float ox = f_sqrt(F(1.f) - mu * mu).v, muf = mu.v;
        const F ray_rho = A<F>::ProfileDensity(P.rayleigh_density, r - a.bottom());
        const F mie_rho = A<F>::ProfileDensity(P.mie_density, r - a.bottom());
        const float dd_ = (F(FB_PI_F) / F(16.f) * (F(FB_PI_F) / F(16.f))).v;        // dtheta * dphi
        const float g = P.mie_phase_function_g;
        const float kRw = 3.f / (16.f * FB_PI_F), kMw = 3.f / (8.f * FB_PI_F) * (1.f - g * g) / (2.f + g * g);
        const float g2p1w = 1.f + g * g, m2gw = -2.f * g;
        const float rr_ = P.rayleigh_scattering[0] * ray_rho.v, rg_ = P.rayleigh_scattering[1] * ray_rho.v,
                    rb_ = P.rayleigh_scattering[2] * ray_rho.v;
        const float mr_ = P.mie_scattering[0] * mie_rho.v, mg_ = P.mie_scattering[1] * mie_rho.v, mb_ = P.mie_scattering[2] * mie_rho.v;
        for (int e = tid; e < DL * 32; e += NWARPS * 32) {
            const int l = e >> 5, m = e & 31;
            const float st = tg.st16[l], ct = tg.ct16[l];
            const float nu2 = fmaf(ox, tg.cp32[m] * st, muf * ct);
            const float dw = dd_ * st;
            const float pr = fmaf(nu2 * kRw, nu2, kRw) * dw;
            const float rs = rsqrt_fast(fmaf(m2gw, nu2, g2p1w));
            const float pm = fmaf(nu2 * kMw, nu2, kMw) * (rs * rs * rs) * dw;
            WtS[e] = make_float4(fmaf(rr_, pr, mr_ * pm), fmaf(rg_, pr, mg_ * pm), fmaf(rb_, pr, mb_ * pm), 0.f);
        }
    }
    uint32_t gmask = 0;                                                        // theta rows that reach the ground (CTA-uniform)
    for (int l = DL / 2; l < DL; ++l)
        if (__ldg(hitG + (size_t)z * DL + l) != 0.f) gmask |= 1u << l;
    __syncthreads();

    // ---- per-lane constants: the weights of phi sample `lane` for the 16 theta rows (registers) ------------
    P2 Wrg[DL];                                                                // (red, green) as one packed operand
    float Wb[DL];
#pragma unroll
    for (int l = 0; l < DL; ++l) {
        const float4 w = WtS[l * 32 + lane];
        Wrg[l] = pk(w.x, w.y); Wb[l] = w.z;
    }
    // ---- work list ------------------------------------------------------------------------------------------
    // Every nu knot outside [mu mu_s - s, mu mu_s + s] is clamped onto the same bound (scattering.h:133-136), so a
    // texel whose (sx, sy, mu_s) equal those of the previous nu slice has bit-identical inputs: it is a FOLLOWER and
    // receives its leader's result (5.5 % of the texels at default dims).  A leader whose sun direction lies in the
    // (zenith, view) plane -- sy ~ 0: every clamped texel, 25 % of the table -- takes the paired body below.
    const float hn = 0.5f * (float)(dd.nu - 1);
    static_assert(TT <= NWARPS * 32 && TT * ENT_B <= (1 << 14), "one thread classifies one texel; packed record fields");
    int nGen = 0, nPair = 0;
    {
        bool isGen = false, isPair = false, follower = false;
        const int t = tid;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < TT) {
            g = geoS[t];
            if (__float_as_int(g.w) >= 0) {
                if (t >= dd.ms_tile) {
                    const float4* pg = geoS + (t - dd.ms_tile);
                    follower = __float_as_int(pg->x) == __float_as_int(g.x) && __float_as_int(pg->y) == __float_as_int(g.y) &&
                               __float_as_int(pg->z) == __float_as_int(g.z);
                }
                if (follower) {}
                else if (PAIRED && g.y * hn <= 0.00390625f) isPair = true;
                else isGen = true;
            }
        }
        // ordered compaction (ballot + per-warp counts): the records are sorted by texel, so the half-warp a paired texel
        // lands in -- and with it the last-ulp rounding of its mirror trig values -- is the same on every run
        const uint32_t bg = __ballot_sync(0xffffffffu, isGen), bp = __ballot_sync(0xffffffffu, isPair);
        const uint32_t bf = __ballot_sync(0xffffffffu, follower);
        if (lane == 0) { cntS[warp] = __popc(bg) | (__popc(bp) << 16); folS[warp] = bf; }
        __syncthreads();                      // counts and follower masks visible; nobody reads the un-compacted records any more
        int before = 0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) {
            const int c = cntS[w];
            if (w < warp) before += c;
            nGen += c & 0xffff; nPair += c >> 16;
        }
        if (isGen || isPair) {
            int nf = 0;                       // followers: the run of flagged texels in the next nu slices
            for (int u = t + dd.ms_tile; u < TT && ((folS[u >> 5] >> (u & 31)) & 1u); u += dd.ms_tile) ++nf;
            // record .w: byte offset of the texel's table rows | texel << 14 | followers << 22
            g.w = __int_as_float(__float_as_int(g.w) | (t << 14) | (nf << 22));
            const uint32_t lt = (1u << lane) - 1u;
            geoS[isGen ? (before & 0xffff) + __popc(bg & lt) : TT - 1 - ((before >> 16) + __popc(bp & lt))] = g;
        }
    }
    __syncthreads();                          // every warp holds its weights: their region now receives the ground rows
    const uint32_t grow_b = (uint32_t)nE * 24u;                                // bytes per ground row
    if (gmask && tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic reads above, async-proxy writes below
        mbar_expect_tx(bar + 1, (DL / 2) * grow_b);
        tma_bulk_g2s(smem_raw + GR_OFF, growG + (size_t)z * (DL / 2) * nE * 3, (DL / 2) * grow_b, bar + 1);
    }
    const float cp = tg.cp32[lane], sp = tg.sp32[lane];
    const float MAGIC = 8388608.f;   // 2^23: x + MAGIC rounded down leaves floor(x) in the low mantissa bits
    // GetIrradiance at r = bottom lands on row 0: u*N - 0.5 = (mu_s*0.5 + 0.5)*(N - 1)
    const float e_c = 0.5f * (float)(nE - 1);
    float kR = 0.f, kMR = 0.f, g2p1 = 0.f, m2g = 0.f;
    if (ORDER2) {
        const float g = P.mie_phase_function_g;
        kR = 3.f / (16.f * FB_PI_F);
        const float kM = 3.f / (8.f * FB_PI_F) * (1.f - g * g) / (2.f + g * g);
        kMR = kM / kR;
        g2p1 = 1.f + g * g;
        m2g = -2.f * g;
    }
    // bits(x + MAGIC) * S == 0x4B000000 * S + S * floor(x)  (mod 2^32): the constant is folded into the base address.
    // It arrives as a kernel argument on purpose: as a literal ptxas splits it off again and pays an IADD3 per load.
    uint32_t sbase = smem_u32(smem_raw);
    asm volatile("" : "+r"(sbase));           // one register for every region; offsets below are immediates
    const uint32_t tab_base = sbase - magic_tab;
    const uint32_t erow_t = sbase - magic_row;
    const int out_row = (z * P.scattering_mu_size + y) * W;

    mbar_wait(bar, 0);
    if (gmask) mbar_wait(bar + 1, 0);

    // ---- one warp per texel, lane = phi sample -------------------------------------------------------------
    // The theta rows that reach the ground are a suffix l >= L0 of the downward rows (a steeper ray hits if a shallower
    // one does).  The two patterns an Earth-like shell produces are compiled with L0 fixed: the 16 unrolled steps then
    // form one basic block and the loads of later steps are scheduled across the ground terms of earlier ones.  Any
    // other pattern takes the L0 = -1 body, which tests the mask per step.
    // The texel and its followers (the next nu slices): every lane of the group holds the sums after the butterfly, so
    // lane `sub` of `group` writes follower sub, sub + group, ... -- one predicated store instead of a serial loop.
    auto store = [&](int rec, float ar, float ag, float ab, int sub, int group) {
        const uint2 v = pack_half4(ar, ag, ab, 0.f);
        const int t = (rec >> 14) & 0xff;
        const int tn = (int)(((uint32_t)t * dd.ms_mul) >> 20);
        uint2* o = out + (out_row + tn * P.scattering_mu_s_size + tile * dd.ms_tile + (t - tn * dd.ms_tile));
        for (int k = sub, nf = rec >> 22; k <= nf; k += group) o[(size_t)k * P.scattering_mu_s_size] = v;
    };
    auto texels = [&](auto L0c) __attribute__((always_inline)) {
    constexpr int L0 = decltype(L0c)::value;
    // the 512 samples of one texel: lane = phi sample, 16 unrolled theta rows; per-lane partial sums come back in (ar, ag, ab)
    auto texel_sums = [&](const float4 geo, float& ar, float& ag, float& ab) __attribute__((always_inline)) {
        // w_s . w_i = sin(theta_l) * q + mu_s * cos(theta_l) with q = sx cos(phi) + sy sin(phi) per (texel, lane).
        // |w_s . w_i| <= sqrt(q^2 + mu_s^2) for every theta; pull (q, mu_s) inside the unit disc by a hair so that
        // neither the nu look-up below nor the ground look-up can step outside its table row — no per-sample clamps.
        float q = fmaf(geo.x, cp, geo.y * sp), mus = geo.z;
        {
            const float n2 = fmaf(q, q, mus * mus);
            const float sc = n2 > 0.999998f ? 0.999999f * rsqrt_fast(n2) : 1.f;
            q *= sc; mus *= sc;
        }
        if (!ORDER2) { q *= hn; mus *= hn; }                                      // tcx = hn * nu1 + hn in two FFMAs
        const uint32_t row_t = tab_base + ((uint32_t)__float_as_int(geo.w) & 0x3fffu);
        P2 arg = pk(0.f, 0.f);
        ab = 0.f;
#define FB_DENSITY_STEP(l)                                                                                              \
        {                                                                                                               \
            float nu1, tcx;                                                       /* scattering.h:146, in [0, nu-1) */  \
            if (ORDER2) { nu1 = fmaf(mus, CT16[l], q * ST16[l]); tcx = fmaf(nu1, hn, hn); }                             \
            else { nu1 = 0.f; tcx = fmaf(mus, CT16[l], fmaf(q, ST16[l], hn)); }                                         \
            const float tm = __fadd_rd(tcx, MAGIC);                               /* floor(tcx) in the low mantissa */  \
            const float f = tcx;                                                  /* entries are (intercept, slope) */  \
            const uint32_t addr = row_t + __float_as_uint(tm) * (uint32_t)ENT_B;                                        \
            P2 Lrg; float Lb;                                                     /* (red, green) packed, blue */       \
            if (ORDER2) {                                                                                               \
                P2 Ri, Rs, Mi, Ms; float4 tb;                                                                           \
                tab12<SW32, (l) * L_STRIDE>(addr, tm, Ri, Rs, Mi, Ms, tb);                                              \
                const float pr = fmaf(nu1 * kR, nu1, kR);                         /* util.h:26-29 */                    \
                const float rs = rsqrt_fast(fmaf(m2g, nu1, g2p1));                                                      \
                const float pm = pr * kMR * (rs * rs * rs);                       /* util.h:31-34: x^-1.5 = rsqrt^3 */  \
                const P2 f2 = bc(f);                                                                                    \
                Lrg = fma2(fma2(f2, Ms, Mi), bc(pm), mul2(fma2(f2, Rs, Ri), bc(pr)));   /* scattering.h:172-173 */      \
                Lb = fmaf(fmaf(f, tb.w, tb.z), pm, fmaf(f, tb.y, tb.x) * pr);                                           \
            } else {                                                                                                    \
                P2 irg, srg; float2 cb;                                                                                 \
                tab3<SW32, (l) * L_STRIDE>(addr, tab_swap(tm), irg, srg, cb);                                           \
                Lrg = fma2(bc(f), srg, irg); Lb = fmaf(f, cb.y, cb.x);                                                  \
            }                                                                                                           \
            if (L0 >= 0 ? (l) >= L0 : ((l) >= DL / 2 && (gmask & (1u << (l))) != 0)) {  /* CTA-uniform */              \
                const float2 N = GN.n[blockIdx.z][(l) - DL / 2];                  /* ground normal, uniform regs */     \
                /* (dot(ground_normal, omega_s) + 1) * e_c, irradiance.h:20-30 at r = bottom; N is pre-scaled */        \
                const float te = fmaf(q, N.x, fmaf(mus, N.y, e_c));                                                     \
                const float em = __fadd_rd(te, MAGIC);                                                                  \
                const float fe = te;                                                                                    \
                const uint32_t ea = erow_t + (uint32_t)((l) - DL / 2) * grow_b + __float_as_uint(em) * 24u;             \
                const P2 ei = lds64p<GR_OFF>(ea), es = lds64p<GR_OFF + 8>(ea);                                          \
                const float2 eb = lds64<GR_OFF + 16>(ea);                                                               \
                Lrg = add2(Lrg, fma2(bc(fe), es, ei));                            /* rows carry the ground factor */    \
                Lb += fmaf(fe, eb.y, eb.x);                                                                             \
            }                                                                                                           \
            arg = fma2(Lrg, Wrg[l], arg); ab = fmaf(Lb, Wb[l], ab);                                                     \
        }
        FB_DENSITY_STEP(0) FB_DENSITY_STEP(1) FB_DENSITY_STEP(2) FB_DENSITY_STEP(3)
        FB_DENSITY_STEP(4) FB_DENSITY_STEP(5) FB_DENSITY_STEP(6) FB_DENSITY_STEP(7)
        FB_DENSITY_STEP(8) FB_DENSITY_STEP(9) FB_DENSITY_STEP(10) FB_DENSITY_STEP(11)
        FB_DENSITY_STEP(12) FB_DENSITY_STEP(13) FB_DENSITY_STEP(14) FB_DENSITY_STEP(15)
#undef FB_DENSITY_STEP
        upk(arg, ar, ag);
    };
    // Two texels per iteration: their partial sums meet in ONE butterfly -- the first exchange hands texel 0 to lanes 0-15
    // and texel 1 to lanes 16-31 (3 shuffles for both texels instead of 6), the remaining four steps run on half-warps.
    // Each texel's sum is associated exactly as by a full-warp butterfly of its own (lane l + lane l ^ 16 first).
    for (int i = warp; i < nGen; i += 2 * NWARPS) {
        const bool two = i + NWARPS < nGen;                                       // warp-uniform
        const float4 geo0 = geoS[i], geo1 = geoS[two ? i + NWARPS : i];
        float a0r, a0g, a0b, a1r = 0.f, a1g = 0.f, a1b = 0.f;
        texel_sums(geo0, a0r, a0g, a0b);
        if (two) texel_sums(geo1, a1r, a1g, a1b);
        const bool up = lane >= 16;
        float ar = up ? a1r : a0r, ag = up ? a1g : a0g, ab = up ? a1b : a0b;      // what this half-warp keeps
        ar += __shfl_xor_sync(0xffffffffu, up ? a0r : a1r, 16);
        ag += __shfl_xor_sync(0xffffffffu, up ? a0g : a1g, 16);
        ab += __shfl_xor_sync(0xffffffffu, up ? a0b : a1b, 16);
#pragma unroll
        for (int s = 8; s > 0; s >>= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, s);
            ag += __shfl_xor_sync(0xffffffffu, ag, s);
            ab += __shfl_xor_sync(0xffffffffu, ab, s);
        }
        if (!up || two) store(__float_as_int(up ? geo1.w : geo0.w), ar, ag, ab, lane & 15, 16);
    }
    // ---- paired body: two texels per warp (lanes 0-15 / 16-31), a lane takes the mirror samples phi_m and 2 pi - phi_m
    // (same cos phi, hence the same phase weight; sin phi flips).  With sy ~ 0 both fall into the same table segment
    // [k, k + 1) except when a knot lies between them, and on a segment L(a) + L(b) = 2 (intercept + slope * mean):
    // one table entry serves two samples.  The warp checks `same segment` per step and otherwise gives each sample its
    // own entry, so the result is the general body's up to summation order.
#ifndef FB_PAIR_UNROLL
#define FB_PAIR_UNROLL 1
#endif
    constexpr int PU = FB_PAIR_UNROLL;
    if (PAIRED)
#pragma unroll PU
    for (int i = warp; 2 * i < nPair; i += NWARPS) {
        const int idx = 2 * i + (lane >> 4);
        const bool valid = idx < nPair;                                          // odd count: the last half-warp idles
        const float4 geo = geoS[TT - 1 - (valid ? idx : idx - 1)];
        const float qm0 = geo.x * cp, qd0 = geo.y * sp;
        float mus = geo.z;
        const float qx = fmaxf(fabsf(qm0 + qd0), fabsf(qm0 - qd0));
        const float n2 = fmaf(qx, qx, mus * mus);
        float sc = n2 > 0.999998f ? 0.999999f * rsqrt_fast(n2) : 1.f;            // one pull for both samples of the pair
        if (!ORDER2) sc *= hn;
        const float qm = qm0 * sc, qd = qd0 * sc, qdn = -qd;
        mus *= sc;
        const uint32_t row_t = tab_base + ((uint32_t)__float_as_int(geo.w) & 0x3fffu);
        P2 arg = pk(0.f, 0.f);
        float ab = 0.f;
        uint32_t straddle = 0;                                                   // bit l: table knot, bit 16 + l: ground-row knot
#define FB_PAIR_STEP(l)                                                                                                 \
        {                                                                                                               \
            const bool gnd = L0 >= 0 ? (l) >= L0 : ((l) >= DL / 2 && (gmask & (1u << (l))) != 0);   /* CTA-uniform */   \
            float nu1a = 0.f, nu1b = 0.f, fa, fb, fm = 0.f;                                                             \
            if (ORDER2) {                                                                                               \
                const float nm = fmaf(mus, CT16[l], qm * ST16[l]);                                                      \
                nu1a = fmaf(qd, ST16[l], nm); nu1b = fmaf(qdn, ST16[l], nm);                                            \
                fa = fmaf(nu1a, hn, hn); fb = fmaf(nu1b, hn, hn);                                                       \
            } else {                                                                                                    \
                fm = fmaf(mus, CT16[l], fmaf(qm, ST16[l], hn));                                                         \
                fa = fmaf(qd, ST16[l], fm); fb = fmaf(qdn, ST16[l], fm);                                                \
            }                                                                                                           \
            const float ta = __fadd_rd(fa, MAGIC), tb = __fadd_rd(fb, MAGIC);                                           \
            if (ta != tb) straddle |= 1u << (l);                                  /* a nu knot between the two */       \
            const uint32_t addr = row_t + __float_as_uint(ta) * (uint32_t)ENT_B;                                        \
            P2 Lrg; float Lb;                                                                                           \
            if (ORDER2) {                                                                                               \
                P2 Ri, Rs, Mi, Ms; float4 tb;                                                                           \
                tab12<SW32, (l) * L_STRIDE>(addr, ta, Ri, Rs, Mi, Ms, tb);                                              \
                const float pra = fmaf(nu1a * kR, nu1a, kR), prb = fmaf(nu1b * kR, nu1b, kR);                           \
                const float rsa = rsqrt_fast(fmaf(m2g, nu1a, g2p1)), rsb = rsqrt_fast(fmaf(m2g, nu1b, g2p1));           \
                const float pma = pra * kMR * (rsa * rsa * rsa), pmb = prb * kMR * (rsb * rsb * rsb);                   \
                const P2 fa2 = bc(fa), fb2 = bc(fb);                                                                    \
                Lrg = add2(fma2(fma2(fa2, Ms, Mi), bc(pma), mul2(fma2(fa2, Rs, Ri), bc(pra))),                          \
                           fma2(fma2(fb2, Ms, Mi), bc(pmb), mul2(fma2(fb2, Rs, Ri), bc(prb))));                         \
                Lb = fmaf(fmaf(fa, tb.w, tb.z), pma, fmaf(fa, tb.y, tb.x) * pra) +                                      \
                     fmaf(fmaf(fb, tb.w, tb.z), pmb, fmaf(fb, tb.y, tb.x) * prb);                                       \
            } else {                          /* half sums: the texel's total is doubled after the loop */              \
                P2 irg, srg; float2 cb;                                                                                 \
                tab3<SW32, (l) * L_STRIDE>(addr, tab_swap(ta), irg, srg, cb);                                           \
                Lrg = fma2(bc(fm), srg, irg); Lb = fmaf(fm, cb.y, cb.x);                                                \
            }                                                                                                           \
            if (gnd) {                                                                                                  \
                const float2 N = GN.n[blockIdx.z][(l) >= DL / 2 ? (l) - DL / 2 : 0];                                    \
                const float tem = fmaf(qm, N.x, fmaf(mus, N.y, e_c));                                                   \
                const float ema = __fadd_rd(fmaf(qd, N.x, tem), MAGIC), emb = __fadd_rd(fmaf(qdn, N.x, tem), MAGIC);    \
                if (ema != emb) straddle |= 0x10000u << (l);                      /* an irradiance knot between them */ \
                const uint32_t ea = erow_t + (uint32_t)((l) - DL / 2) * grow_b + __float_as_uint(ema) * 24u;            \
                const P2 ei = lds64p<GR_OFF>(ea), es = lds64p<GR_OFF + 8>(ea);                                          \
                const float2 eb = lds64<GR_OFF + 16>(ea);                                                               \
                if (ORDER2) {                                                                                           \
                    Lrg = fma2(bc(2.f), fma2(bc(tem), es, ei), Lrg);                                                    \
                    Lb = fmaf(2.f, fmaf(tem, eb.y, eb.x), Lb);                                                          \
                } else {                                                                                                \
                    Lrg = add2(Lrg, fma2(bc(tem), es, ei)); Lb += fmaf(tem, eb.y, eb.x);                                \
                }                                                                                                       \
            }                                                                                                           \
            arg = fma2(Lrg, Wrg[l], arg); ab = fmaf(Lb, Wb[l], ab);                                                     \
        }
        FB_PAIR_STEP(0) FB_PAIR_STEP(1) FB_PAIR_STEP(2) FB_PAIR_STEP(3)
        FB_PAIR_STEP(4) FB_PAIR_STEP(5) FB_PAIR_STEP(6) FB_PAIR_STEP(7)
        FB_PAIR_STEP(8) FB_PAIR_STEP(9) FB_PAIR_STEP(10) FB_PAIR_STEP(11)
        FB_PAIR_STEP(12) FB_PAIR_STEP(13) FB_PAIR_STEP(14) FB_PAIR_STEP(15)
#undef FB_PAIR_STEP
        // The rare sample b that sits beyond a knot was evaluated on a's segment: add (its own segment - a's segment) at
        // its coordinate.  A compact loop over the affected theta rows (run-time l: the weights come out of the register
        // file through a select chain); per-lane predicates only, so a texel's result does not depend on its partner.
        float ar, ag;
        upk(arg, ar, ag);
        {
            const uint32_t any = __reduce_or_sync(0xffffffffu, straddle);
            uint32_t steps = (any | (any >> 16)) & 0xffffu;
            while (steps) {
                const int l = __ffs((int)steps) - 1;
                steps &= steps - 1;
                P2 wrg = Wrg[0];
                float wr, wg, wb = Wb[0];
#pragma unroll
                for (int j = 1; j < DL; ++j)
                    if (l == j) { wrg = Wrg[j]; wb = Wb[j]; }
                upk(wrg, wr, wg);
                const float ct = CT16[l], st = ST16[l];
                const float hs = ORDER2 ? 1.f : 0.5f;
                float dr = 0.f, dg = 0.f, db = 0.f;
                if (straddle & (1u << l)) {
                    float nu1b = 0.f, fa, fb;
                    if (ORDER2) {
                        const float nm = fmaf(mus, ct, qm * st);
                        nu1b = fmaf(qdn, st, nm);
                        fa = fmaf(fmaf(qd, st, nm), hn, hn); fb = fmaf(nu1b, hn, hn);
                    } else {
                        const float fm = fmaf(mus, ct, fmaf(qm, st, hn));
                        fa = fmaf(qd, st, fm); fb = fmaf(qdn, st, fm);
                    }
                    const uint32_t lrow = row_t + (uint32_t)(l * L_STRIDE);
                    const float tma = __fadd_rd(fa, MAGIC), tmb = __fadd_rd(fb, MAGIC);
                    const uint32_t addra = lrow + __float_as_uint(tma) * (uint32_t)ENT_B;
                    const uint32_t addrb = lrow + __float_as_uint(tmb) * (uint32_t)ENT_B;
                    float r1, g1, b1, r2, g2, b2;
                    density_tap<ORDER2, SW32, 0, GR_OFF>(addrb, tab_swap(tmb), tmb, fb, nu1b, kR, kMR, g2p1, m2g, false, 0u, 0.f, r1, g1, b1);
                    density_tap<ORDER2, SW32, 0, GR_OFF>(addra, tab_swap(tma), tma, fb, nu1b, kR, kMR, g2p1, m2g, false, 0u, 0.f, r2, g2, b2);
                    dr = hs * (r1 - r2); dg = hs * (g1 - g2); db = hs * (b1 - b2);
                }
                if (straddle & (0x10000u << l)) {                                  // only ever set for l >= DL / 2
                    const float2 N = GN.n[blockIdx.z][(l - DL / 2) & (DL / 2 - 1)];
                    const float tem = fmaf(qm, N.x, fmaf(mus, N.y, e_c));
                    const float teb = fmaf(qdn, N.x, tem);
                    const uint32_t grow_l = erow_t + (uint32_t)(l - DL / 2) * grow_b;
                    const uint32_t ea = grow_l + __float_as_uint(__fadd_rd(fmaf(qd, N.x, tem), MAGIC)) * 24u;
                    const uint32_t eb_ = grow_l + __float_as_uint(__fadd_rd(teb, MAGIC)) * 24u;
                    // ground-row entries: (red, green) intercepts, (red, green) slopes, blue (intercept, slope)
                    const float2 ei = lds64<GR_OFF>(ea), es = lds64<GR_OFF + 8>(ea), eb = lds64<GR_OFF + 16>(ea);
                    const float2 fi = lds64<GR_OFF>(eb_), fs = lds64<GR_OFF + 8>(eb_), fbb = lds64<GR_OFF + 16>(eb_);
                    dr += hs * (fmaf(teb, fs.x, fi.x) - fmaf(teb, es.x, ei.x));
                    dg += hs * (fmaf(teb, fs.y, fi.y) - fmaf(teb, es.y, ei.y));
                    db += hs * (fmaf(teb, fbb.y, fbb.x) - fmaf(teb, eb.y, eb.x));
                }
                ar = fmaf(dr, wr, ar); ag = fmaf(dg, wg, ag); ab = fmaf(db, wb, ab);
            }
        }
        if (!ORDER2) { ar += ar; ag += ag; ab += ab; }
#pragma unroll
        for (int s = 8; s > 0; s >>= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, s);
            ag += __shfl_xor_sync(0xffffffffu, ag, s);
            ab += __shfl_xor_sync(0xffffffffu, ab, s);
        }
        if (valid) store(__float_as_int(geo.w), ar, ag, ab, lane & 15, 16);
    }
    };
    if (gmask == 0xFF00u) texels(std::integral_constant<int, 8>());
    else if (gmask == 0xFE00u) texels(std::integral_constant<int, 9>());
    else texels(std::integral_constant<int, -1>());
}

template <bool ORDER2> static size_t density_smem(const FbParams& P) {
    typedef DensityCfg<ORDER2> C;
    const size_t wt = (size_t)DL * 32 * 16, gr = (size_t)(DL / 2) * P.irradiance_mu_s_size * 24;
    return (size_t)DL * C::T * C::ENT * 4 + C::T * 16 + 128 + (wt > gr ? wt : gr);
}

template <bool ORDER2, bool SW32>
static cudaError_t density_launch(const LaunchCtx& c, Tex3 A0, Tex3 A1, int r0, int r1, cudaEvent_t after_prep) {
    typedef DensityCfg<ORDER2> C;
    const FbParams& P = c.P;
    const DensityDims d = density_dims(P, C::T);
    const size_t smem = density_smem<ORDER2>(P);
    float* tab = c.img.scratch;
    float2* grow = reinterpret_cast<float2*>(tab + density_tab_floats(P));
    float* hit = tab + density_tab_floats(P) + density_grow_floats(P);
    const int W = P.scattering_nu_size * P.scattering_mu_s_size;
    dim3 gp((W + 255) / 256, DL, r1 - r0);
    k_density_prep<ORDER2><<<gp, 256, 0, c.stream>>>(P, c.trig, texT(c), A0, A1, d, tab, hit, c.img.delta_irradiance, grow, r0, SW32 ? 1 : 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // from here on nothing reads delta_irradiance: indirect_irradiance may run concurrently with the main kernel
    if (after_prep && (e = cudaEventRecord(after_prep, c.stream)) != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_density_main<ORDER2, SW32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // ground normals (scattering_density.comp:81-87) per (r, downward theta row); the look-up coordinate the kernel forms
    // is te = (q n_x + mu_s n_z + 1) e_c with e_c = (nE - 1) / 2, and at order >= 3 its q, mu_s carry hn = (nu - 1) / 2
    const double bot = P.bottom_radius, top = P.top_radius, Hh = std::sqrt(top * top - bot * bot);
    const double e_c = 0.5 * (P.irradiance_mu_s_size - 1), ksc = ORDER2 ? e_c : e_c / (0.5 * (d.nu - 1));
    for (int zc = r0; zc < r1; zc += GN_MAXR) {
        const int nz_ = std::min(GN_MAXR, r1 - zc);
        GroundNormals gn;
        std::memset(&gn, 0, sizeof gn);
        for (int i = 0; i < nz_; ++i) {
            const double rho = P.scattering_r_size > 1 ? Hh * (zc + i) / (P.scattering_r_size - 1) : 0.0;   // scattering.h:64-67
            const double r = std::sqrt(rho * rho + bot * bot);
            for (int l = DL / 2; l < DL; ++l) {
                const double ct = c.trig.ct16[l], st = c.trig.st16[l];
                const double dg = std::max(-r * ct - std::sqrt(std::max(r * r * (ct * ct - 1.0) + bot * bot, 0.0)), 0.0);
                const double vx = st * dg, vz = r + ct * dg, len = std::sqrt(vx * vx + vz * vz);
                gn.n[i][l - DL / 2] = make_float2((float)(vx / len * ksc), (float)(vz / len * ksc));
            }
        }
        dim3 gm(d.tiles, P.scattering_mu_size, nz_);
        k_density_main<ORDER2, SW32><<<gm, C::NWARPS * 32, smem, c.stream>>>(P, c.trig, d, tab, hit, grow, c.img.scattering_density, zc,
                                                                      0x4B000000u * (uint32_t)(C::ENT * 4), 0x4B000000u * 24u, gn
                                                                      );
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t scattering_density(const LaunchCtx& c, int order, int r0, int r1, cudaEvent_t after_prep) {
    if (!density_supported(c.P) || !c.img.scratch) {
        cudaError_t e = ref::scattering_density(c, order, r0, r1);   // reads delta_irradiance throughout: no early event
        if (e == cudaSuccess && after_prep) e = cudaEventRecord(after_prep, c.stream);
        return e;
    }
    // more than 16 nu knots: bank-swizzled table entries read with 32-bit loads (tab3 / tab12)
    if (order == 2 && c.P.scattering_nu_size > 16)
        return density_launch<true, true>(c, texS(c, c.img.delta_rayleigh), texS(c, c.img.delta_mie), r0, r1, after_prep);
    if (order == 2) return density_launch<true, false>(c, texS(c, c.img.delta_rayleigh), texS(c, c.img.delta_mie), r0, r1, after_prep);
    if (c.P.scattering_nu_size > 16)
        return density_launch<false, true>(c, texS(c, c.img.delta_multiple_scattering), texS(c, c.img.delta_multiple_scattering), r0, r1, after_prep);
    return density_launch<false, false>(c, texS(c, c.img.delta_multiple_scattering), texS(c, c.img.delta_multiple_scattering), r0, r1, after_prep);
}

// ---------------------------------------------------------------------------------------------
// transmittance.comp — one WARP per texel: the 501 trapezoid nodes are strided over the lanes (the
// reference walks them serially in one thread: 16 384 threads x 1 503 dependent steps, latency-bound on 148
// SMs), the square root is shared by the three density profiles, the three partial sums meet in a
// __shfl_xor tree.  Per-node arithmetic is the shader's, in xf.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_transmittance(const __grid_constant__ FbParams P, float4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int texel = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int W = P.transmittance_mu_size;
    if (texel >= W * P.transmittance_r_size) return;                                   // warp-uniform
    const int x = texel % W, y = texel / W;
    A<F> a(P);
    F r, mu;
    a.RMuFromUnitRanges(F((float)x) / F((float)(W - 1)), F((float)y) / F((float)(P.transmittance_r_size - 1)), r, mu);
    const int N = 500;                                                                 // transmittance.comp:14
    const F dx = a.DistanceToTop(r, mu) / F((float)N);
    const F c2 = F(2.f) * r * mu, rr = r * r;
    F sR = F(0.f), sM = F(0.f), sA = F(0.f);
    for (int i = lane; i <= N; i += 32) {
        const F d_i = F((float)i) * dx;
        const F r_i = f_sqrt(d_i * d_i + c2 * d_i + rr);
        const F h = r_i - a.bottom();
        const F w_i = (i == 0 || i == N) ? F(0.5f) : F(1.f);
        sR += A<F>::ProfileDensity(P.rayleigh_density, h) * w_i * dx;
        sM += A<F>::ProfileDensity(P.mie_density, h) * w_i * dx;
        sA += A<F>::ProfileDensity(P.absorption_density, h) * w_i * dx;
    }
    float fR = sR.v, fM = sM.v, fA = sA.v;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        fR += __shfl_xor_sync(0xffffffffu, fR, s);
        fM += __shfl_xor_sync(0xffffffffu, fM, s);
        fA += __shfl_xor_sync(0xffffffffu, fA, s);
    }
    if (lane == 0) {
        V tau = V(P.rayleigh_scattering) * F(fR) + V(P.mie_extinction) * F(fM) + V(P.absorption_extinction) * F(fA);
        out[texel] = make_float4(expf(-tau.x.v), expf(-tau.y.v), expf(-tau.z.v), 1.f);
    }
}

cudaError_t transmittance(const LaunchCtx& c) {
    const int texels = c.P.transmittance_mu_size * c.P.transmittance_r_size;
    k_transmittance<<<(texels + 7) / 8, 256, 0, c.stream>>>(c.P, c.img.transmittance);
    return cudaGetLastError();
}

cudaError_t direct_irradiance(const LaunchCtx& c) { return ref::direct_irradiance(c); }   // 1 024 threads, one tap each

// ---------------------------------------------------------------------------------------------
// indirect_irradiance.comp — one 4-warp CTA per texel (1 024 texels x 1 024 directions: one thread per texel
// would leave 140 SMs idle).  A thread takes one phi sample of 8 theta rows; each sample is the shader's 4-D
// look-up, as written, in xf; the hemisphere sum is a __shfl_xor tree + a 4-way shared-memory add.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_indirect_irradiance(const __grid_constant__ FbParams P, const __grid_constant__ Trig tg,
                                                             Tex3 dR, Tex3 dM, Tex3 dMS, int order, float4* __restrict__ dE,
                                                             float4* __restrict__ E, int texel0) {
    // one CTA (4 warps) per texel: thread t takes phi sample (t & 63) of the theta rows j = (t >> 6), +2, +4, ...
    __shared__ float part[4][3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int texel = texel0 + blockIdx.x;
    const int W = P.irradiance_mu_s_size;
    const int x = texel % W, y = texel / W;
    A<F> a(P);
    F r, mu_s;
    a.RMuSFromIrradianceUnit(F((float)x) / F((float)(W - 1)), F((float)y) / F((float)(P.irradiance_r_size - 1)), r, mu_s);
    const F dphi = F(FB_PI_F) / F(32.f), dtheta = F(FB_PI_F) / F(32.f);
    const V omega_s(f_sqrt(F(1.f) - mu_s * mu_s), F(0.f), mu_s);
    const int i = tid & 63;
    const F cp = F(tg.cp64[i]), sp = F(tg.sp64[i]);
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int j = tid >> 6; j < 16; j += 2) {
        const F ct = F(tg.ct32[j]), st = F(tg.st32[j]);
        const F dw = dtheta * dphi * st;
        const V w(cp * st, sp * st, ct);
        const F nu = dot(w, omega_s);
        const V t = a.ScatteringOrder(dR, dM, dMS, r, w.z, mu_s, nu, false, order) * w.z * dw;
        ax += t.x.v; ay += t.y.v; az += t.z.v;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, s);
        ay += __shfl_xor_sync(0xffffffffu, ay, s);
        az += __shfl_xor_sync(0xffffffffu, az, s);
    }
    if (lane == 0) { part[warp][0] = ax; part[warp][1] = ay; part[warp][2] = az; }
    __syncthreads();
    if (tid == 0) {
        ax = (part[0][0] + part[1][0]) + (part[2][0] + part[3][0]);
        ay = (part[0][1] + part[1][1]) + (part[2][1] + part[3][1]);
        az = (part[0][2] + part[1][2]) + (part[2][2] + part[3][2]);
        dE[texel] = make_float4(ax, ay, az, 0.f);                                      // indirect_irradiance.comp:72
        const float4 e = E[texel];                                                     // :73
        E[texel] = make_float4(__fadd_rn(ax, e.x), __fadd_rn(ay, e.y), __fadd_rn(az, e.z), __fadd_rn(0.f, e.w));
    }
}

cudaError_t indirect_irradiance(const LaunchCtx& c, int order, int row0, int row1) {   // rows [row0, row1) of the irradiance table
    const int texels = c.P.irradiance_mu_s_size * (row1 - row0);
    k_indirect_irradiance<<<texels, 128, 0, c.stream>>>(c.P, c.trig, texS(c, c.img.delta_rayleigh), texS(c, c.img.delta_mie),
                                                        texS(c, c.img.delta_multiple_scattering), order,
                                                        c.img.delta_irradiance, c.img.irradiance, row0 * c.P.irradiance_mu_s_size);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// single_scattering.comp / multiple_scattering.comp — one CTA per (r, mu) row of the table, one thread per
// (nu, mu_s) texel of the row.  The 51 trapezoid nodes of a ray depend on (r, mu) only, so everything the
// shader recomputes per texel from (r, mu, d_i) — the node radius r_d, both transmittance look-ups of
// GetTransmittance(r, mu, d_i), the density profiles, the horizon terms of GetTransmittanceToSun, the
// (mu, r) interpolation cell of the 4-D look-up — is evaluated ONCE per node by thread i, exactly as written,
// and shared through shared memory.  What stays per texel and node is the sun-angle part.
// ---------------------------------------------------------------------------------------------
constexpr int NS = 51;   // SAMPLE_COUNT + 1, single_scattering.comp:42 / multiple_scattering.comp:21

__device__ __forceinline__ float rcp_fast(float x) {    // one MUFU.RCP; callers guarantee a normal x
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_fast(float x) {   // MUFU-based, ~1 ulp; used only where the result is well-conditioned
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct SingleNode {      // per trapezoid node, block-uniform; four 16-byte words, each one broadcast LDS.128
    float4 g0;           // d, 1/r_d, r_d, r_d^2
    float4 g1;           // d_min, 1/(d_max - d_min), cos_h + e0 (smoothstep origin), 1/(e1 - e0)
    float4 t;            // GetTransmittance(r, mu, d_i).rgb, fy (row fraction of r_d in the transmittance table)
    float rho_r, rho_m;  // density * trapezoid weight
    int row0, row1;      // texel offsets of the two transmittance rows bracketing r_d
};

// The per-node quantities (ray geometry, both transmittance taps of GetTransmittance(r, mu, d_i), density profiles,
// horizon terms, the table rows bracketing r_d) are exact (xf, as written).  The per-texel-per-node remainder — the
// sun-angle cosine at the node, its u coordinate in the transmittance table, the bilinear blend, the smoothstep —
// is a smooth, positive, well-conditioned chain and runs in contracted fp32 with MUFU reciprocal / square root.
template <bool STAGED>   // STAGED: row pairs pre-blended in shared memory (pays when a CTA covers at least one table row of texels)
__global__ void __launch_bounds__(256) k_single_scattering(const __grid_constant__ FbParams P, Tex2 T, uint2* __restrict__ dR,
                                                           uint2* __restrict__ dM, uint2* __restrict__ S, int r0, int CH) {
    __shared__ SingleNode nodes[NS];
    __shared__ float s_dx;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* slab = reinterpret_cast<float4*>(smem_raw);   // [CH][transmittance_mu_size]: per node, the row pair pre-blended
    const int W = P.scattering_nu_size * P.scattering_mu_s_size;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = r0 + blockIdx.z;
    A<F> a(P);
    F r, mu, mu_s, nu;
    bool hits;
    a.TexelToRMuMuSNu((unsigned)min(x, W - 1), (unsigned)y, (unsigned)z, r, mu, mu_s, nu, hits);
    const F H = f_sqrt(a.top() * a.top() - a.bottom() * a.bottom());
    // two threads per node (i and 64 + i): the transmittance look-up of GetTransmittance(r, mu, d_i) and the r_d-only
    // terms are independent halves of the CTA's serial prologue
    const bool split = blockDim.x >= 128;
    for (int j = threadIdx.x; j < (split ? 128 : NS); j += blockDim.x) {
        const int i = split ? (j & 63) : j;
        if (i >= NS) continue;
        const bool partA = !split || j < 64, partB = !split || j >= 64;
        const F dx = a.DistanceToNearest(r, mu, hits) / F(50.f);
        const F d = F((float)i) * dx;
        if (partA) {
            const V tr = a.Transmittance(T, r, mu, d, hits);                            // :19-21
            nodes[i].t.x = tr.x.v; nodes[i].t.y = tr.y.v; nodes[i].t.z = tr.z.v;
            if (i == 0) s_dx = dx.v;
        }
        if (partB) {
            const F r_d = a.ClampRadius(f_sqrt(d * d + F(2.f) * r * mu * d + r * r));  // single_scattering.comp:16
            nodes[i].g0 = make_float4(d.v, (F(1.f) / r_d).v, r_d.v, (r_d * r_d).v);
            // GetTransmittanceToSun(r_d, .) and GetTransmittanceTextureUvFromRMu(r_d, .): the r_d-only parts
            const F rho = A<F>::SafeSqrt(r_d * r_d - a.bottom() * a.bottom());          // transmittance.h:14
            const F d_min = a.top() - r_d, d_max = rho + H;
            const F v = A<F>::CoordFromUnit(rho / H, P.transmittance_r_size);           // transmittance.h:21-23
            int y0, y1;
            F fy;
            tex_axis(v, P.transmittance_r_size, y0, y1, fy);
            nodes[i].row0 = y0 * P.transmittance_mu_size; nodes[i].row1 = y1 * P.transmittance_mu_size;
            const F sin_h = a.bottom() / r_d;                                           // transmittance.h:67-73
            const F cos_h = -f_sqrt(f_max(F(1.f) - sin_h * sin_h, F(0.f)));
            const F al = F(P.sun_angular_radius);
            const F e0 = -sin_h * al, e1 = sin_h * al;
            nodes[i].g1 = make_float4(d_min.v, (F(1.f) / (d_max - d_min)).v, (cos_h + e0).v, (F(1.f) / (e1 - e0)).v);
            nodes[i].t.w = fy.v;
            const F w = (i == 0 || i == NS - 1) ? F(0.5f) : F(1.f);                     // a power of two: folding it is exact
            nodes[i].rho_r = (A<F>::ProfileDensity(P.rayleigh_density, r_d - a.bottom()) * w).v;   // :24-27
            nodes[i].rho_m = (A<F>::ProfileDensity(P.mie_density, r_d - a.bottom()) * w).v;
        }
    }
    const bool active = x < W;
    const float r_mu_s = (r * mu_s).v, nuf = nu.v, tt = P.top_radius * P.top_radius;
    const int TW = P.transmittance_mu_size;
    const float un = (float)(TW - 1);                                                   // u*N - 0.5 == x_mu * (N - 1)
    const float umax = __int_as_float(__float_as_int(un) - 1);
    float rsr = 0.f, rsg = 0.f, rsb = 0.f, msr = 0.f, msg = 0.f, msb = 0.f;
    // GetTransmittanceToSun(r_d, mu_s_d) reads the two table rows bracketing r_d -- a per-NODE pair -- at a column that
    // depends on the texel: the 32 lanes of a warp (32 mu_s values) spread over the whole row, 19 cache lines per
    // load.  Per chunk of CH nodes the CTA therefore blends each node's row pair once (coalesced), folds in
    // GetTransmittance(r, mu, d_i), and the per-sample look-up becomes two shared-memory taps.
    if (!STAGED) {                                // direct look-ups: 4 global taps per sample
        __syncthreads();
        if (active) {
#pragma unroll 3
            for (int i = 0; i < NS; ++i) {
                const float4 g0 = nodes[i].g0, g1 = nodes[i].g1, nt = nodes[i].t;
                const float mu_s_d = fminf(fmaxf(fmaf(g0.x, nuf, r_mu_s) * g0.y, -1.f), 1.f);   // :17
                const float disc = fmaf(g0.w, fmaf(mu_s_d, mu_s_d, -1.f), tt);                  // params.h:105-110
                const float dtop = fmaxf(fmaf(-g0.z, mu_s_d, sqrt_fast(fmaxf(disc, 0.f))), 0.f);
                const float tu = fminf(fmaxf((dtop - g1.x) * g1.y * un, 0.f), umax);            // transmittance.h:20-22
                const float tm = __fadd_rd(tu, 8388608.f);
                const int j = __float_as_int(tm) - 0x4B000000;
                const float fx = tu - (tm - 8388608.f);
                const float4 a00 = __ldg(T.p + nodes[i].row0 + j), a10 = __ldg(T.p + nodes[i].row0 + j + 1);
                const float4 a01 = __ldg(T.p + nodes[i].row1 + j), a11 = __ldg(T.p + nodes[i].row1 + j + 1);
                float sm = fminf(fmaxf((mu_s_d - g1.z) * g1.w, 0.f), 1.f);                      // smoothstep, transmittance.h:71-73
                sm = sm * sm * fmaf(-2.f, sm, 3.f);
                const float b0r = fmaf(fx, a10.x - a00.x, a00.x), b0g = fmaf(fx, a10.y - a00.y, a00.y), b0b = fmaf(fx, a10.z - a00.z, a00.z);
                const float b1r = fmaf(fx, a11.x - a01.x, a01.x), b1g = fmaf(fx, a11.y - a01.y, a01.y), b1b = fmaf(fx, a11.z - a01.z, a01.z);
                const float tr_ = fmaf(nt.w, b1r - b0r, b0r) * sm * nt.x, tg_ = fmaf(nt.w, b1g - b0g, b0g) * sm * nt.y,
                            tb_ = fmaf(nt.w, b1b - b0b, b0b) * sm * nt.z;
                const float rr = nodes[i].rho_r, rm = nodes[i].rho_m;
                rsr = fmaf(tr_, rr, rsr); rsg = fmaf(tg_, rr, rsg); rsb = fmaf(tb_, rr, rsb);
                msr = fmaf(tr_, rm, msr); msg = fmaf(tg_, rm, msg); msb = fmaf(tb_, rm, msb);
            }
        }
    } else
    for (int c0 = 0; c0 < NS; c0 += CH) {
        const int cn = min(CH, NS - c0);
        __syncthreads();                          // nodes ready (first pass) / previous chunk consumed
        for (int e0 = 0; e0 < cn; e0 += 3) {
            for (int j = threadIdx.x; j < TW; j += blockDim.x) {
                float4 ra[3], rb[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) {     // 6 row loads in flight per thread
                    const SingleNode& n = nodes[c0 + min(e0 + u, cn - 1)];
                    ra[u] = __ldg(T.p + n.row0 + j); rb[u] = __ldg(T.p + n.row1 + j);
                }
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    if (e0 + u < cn) {
                        const float4 nt = nodes[c0 + e0 + u].t;
                        slab[(e0 + u) * TW + j] = make_float4(fmaf(nt.w, rb[u].x - ra[u].x, ra[u].x) * nt.x, fmaf(nt.w, rb[u].y - ra[u].y, ra[u].y) * nt.y,
                                                              fmaf(nt.w, rb[u].z - ra[u].z, ra[u].z) * nt.z, 0.f);
                    }
                }
            }
        }
        __syncthreads();
        if (active) {
#pragma unroll 3
            for (int e = 0; e < cn; ++e) {
                const int i = c0 + e;
                const float4 g0 = nodes[i].g0, g1 = nodes[i].g1;
                const float mu_s_d = fminf(fmaxf(fmaf(g0.x, nuf, r_mu_s) * g0.y, -1.f), 1.f);   // :17
                // DistanceToTopAtmosphereBoundary(r_d, mu_s_d), params.h:105-110
                const float disc = fmaf(g0.w, fmaf(mu_s_d, mu_s_d, -1.f), tt);
                const float dtop = fmaxf(fmaf(-g0.z, mu_s_d, sqrt_fast(fmaxf(disc, 0.f))), 0.f);
                const float tu = fminf(fmaxf((dtop - g1.x) * g1.y * un, 0.f), umax);            // transmittance.h:20-22
                const float tm = __fadd_rd(tu, 8388608.f);
                const int j = __float_as_int(tm) - 0x4B000000;
                const float fx = tu - (tm - 8388608.f);
                // (the 32 columns of a warp are irregularly spaced: ~2.2 wavefronts per ideal one; padding does not help)
                const float4 p0 = slab[e * TW + j], p1 = slab[e * TW + j + 1];
                float sm = fminf(fmaxf((mu_s_d - g1.z) * g1.w, 0.f), 1.f);                      // smoothstep, transmittance.h:71-73
                sm = sm * sm * fmaf(-2.f, sm, 3.f);
                const float tr_ = fmaf(fx, p1.x - p0.x, p0.x) * sm, tg_ = fmaf(fx, p1.y - p0.y, p0.y) * sm, tb_ = fmaf(fx, p1.z - p0.z, p0.z) * sm;
                const float rr = nodes[i].rho_r, rm = nodes[i].rho_m;
                rsr = fmaf(tr_, rr, rsr); rsg = fmaf(tg_, rr, rsg); rsb = fmaf(tb_, rr, rsb);
                msr = fmaf(tr_, rm, msr); msg = fmaf(tg_, rm, msg); msb = fmaf(tb_, rm, msb);
            }
        }
    }
    if (!active) return;
    const F dx = F(s_dx);
    const V ray = V(F(rsr), F(rsg), F(rsb)) * dx * V(P.solar_irradiance) * V(P.rayleigh_scattering);   // :62-64
    const V mie = V(F(msr), F(msg), F(msb)) * dx * V(P.solar_irradiance) * V(P.mie_scattering);
    const size_t o = ((size_t)z * P.scattering_mu_size + y) * W + x;
    dR[o] = pack_half4(ray.x.v, ray.y.v, ray.z.v, 0.f);
    dM[o] = pack_half4(mie.x.v, mie.y.v, mie.z.v, 0.f);
    S[o] = pack_half4(ray.x.v, ray.y.v, ray.z.v, mie.x.v);
}

cudaError_t single_scattering(const LaunchCtx& c, int r0, int r1) {
    const int W = c.P.scattering_nu_size * c.P.scattering_mu_s_size;
    const int nt = W >= 256 ? 256 : ((W + 31) / 32) * 32;
    dim3 g((W + nt - 1) / nt, c.P.scattering_mu_size, r1 - r0);
    const int TW = c.P.transmittance_mu_size;
    int CH = 3072 / TW;                           // nodes staged per pass: CH * TW * 16 B <= 48 KiB
    CH = CH < 1 ? 1 : (CH > 17 ? 17 : CH);
    const size_t smem = (size_t)CH * TW * sizeof(float4);
    // staging touches TW entries per node, the samples nt texels per node: stage when a CTA's texels cover the row
    if (TW > nt || smem > 160 * 1024) {
        k_single_scattering<false><<<g, nt, 0, c.stream>>>(c.P, texT(c), c.img.delta_rayleigh, c.img.delta_mie, c.img.scattering, r0, 0);
        return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(k_single_scattering<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_single_scattering<true><<<g, nt, smem, c.stream>>>(c.P, texT(c), c.img.delta_rayleigh, c.img.delta_mie, c.img.scattering, r0, CH);
    return cudaGetLastError();
}

struct MultiNode {       // per trapezoid node, block-uniform; 16-byte words so each is one (broadcast) LDS.128
    float4 w;            // bilinear weights of the (mu, r) cell of GetScattering(r_i, mu_i, ., ., hits): (y0z0, y1z0, y0z1, y1z1)
    uint4 off;           // texel offsets of those four rows in the density table
    float4 t;            // GetTransmittance(r, mu, d_i) * dx * trapezoid weight (rgb), node distance d_i
    float inv_r, pad0, pad1, pad2;   // 1 / r_i
};


#ifndef FB_MS_U
#define FB_MS_U 4        // nodes whose density rows are in flight together during staging
#endif
#ifndef FB_MS_MINB
#define FB_MS_MINB 4     // CTAs per SM of the 256-thread variant
#endif
#ifndef FB_MS_CU
#define FB_MS_CU 4       // unroll of the per-texel node loop
#endif
template <int TPT, int NTMAX>   // TPT texels per thread: the CTA covers the whole (nu, mu_s) row, W <= TPT * blockDim.x
__global__ void __launch_bounds__(NTMAX, NTMAX == 256 ? FB_MS_MINB : 1) k_multiple_scattering(const __grid_constant__ FbParams P, Tex2 T, const uint2* __restrict__ dens,
                                                              uint2* __restrict__ dMS, uint2* __restrict__ S, int r0, int CH) {
    constexpr int CU = FB_MS_CU;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MultiNode* nodes = reinterpret_cast<MultiNode*>(smem_raw);
    float4* slab = reinterpret_cast<float4*>(smem_raw + sizeof(MultiNode) * NS);
    const int NU = P.scattering_nu_size, MS = P.scattering_mu_s_size, W = NU * MS;
    const int y = blockIdx.y, z = r0 + blockIdx.z;
    A<F> a(P);
    // per-texel constants of the 4-D look-up: the nu slice pair (scattering.h:146-152) and the sun-angle terms
    float ln[TPT], rmus[TPT], nuf[TPT], ar[TPT], ag[TPT], ab[TPT];
    int kx0[TPT], kx1[TPT];
    bool two[TPT];       // this texel interpolates between two nu slices (scattering.h:150-154)
    F r, mu;
    bool hits;
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        const int x = min((int)(threadIdx.x + k * blockDim.x), W - 1);
        F mu_s, nu;
        a.TexelToRMuMuSNu((unsigned)x, (unsigned)y, (unsigned)z, r, mu, mu_s, nu, hits);
        const F tcx = (nu + F(1.f)) / F(2.f) * F((float)(NU - 1));
        const F txf = f_floor(tcx);
        ln[k] = (tcx - txf).v;
        int tx = min(max((int)txf.v, 0), NU - 1);
        // An un-clamped nu is a knot value, but (nu + 1) / 2 * (NU - 1) may round to k - 1 ulp: the as-written blend is then
        // v[k - 1] * 6e-8 + v[k] * (1 - 6e-8).  Such a texel is put onto its knot (a 1e-7-relative change of a convex
        // blend, the same class as FMA contraction), so that only clamped texels ever read two nu slices.
        if (ln[k] > 0.999999f) { tx = min(tx + 1, NU - 1); ln[k] = 0.f; }
        else if (ln[k] < 0.000001f) ln[k] = 0.f;
        kx0[k] = tx * MS; kx1[k] = min(tx + 1, NU - 1) * MS;
        rmus[k] = (r * mu_s).v; nuf[k] = nu.v;
        ar[k] = ag[k] = ab[k] = 0.f;
        // a texel whose nu was not clamped sits on a nu knot: lerp == 0 exactly and fma(0, v1 - v0, v0) == v0, so the
        // second slice need not be fetched (bit-identical; 65 % of the texels at default dims)
        two[k] = ln[k] != 0.f;
    }
    // Texels of one mu_s column whose nu knots were clamped onto the same bound (scattering.h:133-136) have bit-identical
    // inputs.  A run of equal nu that contains slice 0 is led by slice 0, any other run by its top slice: at default dims
    // the leaders of clamped runs then all sit in the first and the last warp, and the six warps in between hold only
    // un-clamped texels (one slice, warp-uniformly) and followers, which wait for their leader's result.
    int leader[TPT];
    {
        // (nu, r mu_s) of every texel; aliases the slab: consumed before the first staging pass.  mu_s is derived from the
        // texel's own x coordinate and may differ in its last bit between nu slices: both inputs must match bit for bit.
        int2* nuS = reinterpret_cast<int2*>(slab);
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
            const int x = threadIdx.x + k * blockDim.x;
            if (x < W) nuS[x] = make_int2(__float_as_int(nuf[k]), __float_as_int(rmus[k]));
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
            const int x = threadIdx.x + k * blockDim.x;
            leader[k] = x;
            if (x < W) {
                const int2 me = make_int2(__float_as_int(nuf[k]), __float_as_int(rmus[k]));
                int lo = x, hi = x;
                while (lo >= MS && nuS[lo - MS].x == me.x && nuS[lo - MS].y == me.y) lo -= MS;
                while (hi + MS < W && nuS[hi + MS].x == me.x && nuS[hi + MS].y == me.y) hi += MS;
                leader[k] = lo < MS ? lo : hi;
            }
        }
    }
    // node records: the transmittance half and the (mu, r)-cell half of a node are independent, so two threads share a
    // node (threads i and 64 + i when the CTA has them): the serial prologue of the 51 node threads is the CTA's
    // critical path
    const bool split = blockDim.x >= 128;
    for (int j = threadIdx.x; j < (split ? 128 : NS); j += blockDim.x) {
        const int i = split ? (j & 63) : j;
        if (i >= NS) continue;
        const bool partA = !split || j < 64, partB = !split || j >= 64;
        const F dx = a.DistanceToNearest(r, mu, hits) / F(50.f);                        // multiple_scattering.comp:23-26
        const F d = F((float)i) * dx;
        if (partA) {
            const V tr = a.Transmittance(T, r, mu, d, hits) * dx;                       // :45-48
            const F w = (i == 0 || i == NS - 1) ? F(0.5f) : F(1.f);
            nodes[i].t = make_float4((tr.x * w).v, (tr.y * w).v, (tr.z * w).v, d.v);
        }
        if (partB) {
            const F r_i = a.ClampRadius(f_sqrt(d * d + F(2.f) * r * mu * d + r * r));  // :35-37
            const F mu_i = A<F>::ClampCosine((r * mu + d) / r_i);
            F uvwz[4];
            a.ScatteringUvwz(r_i, mu_i, F(0.f), F(0.f), hits, uvwz);                    // u_mu, u_r of scattering.h:7-60
            int y0, y1, z0, z1;
            F fy, fz;
            tex_axis(uvwz[2], P.scattering_mu_size, y0, y1, fy);
            tex_axis(uvwz[3], P.scattering_r_size, z0, z1, fz);
            const float gy = 1.f - fy.v, gz = 1.f - fz.v;
            nodes[i].w = make_float4(gy * gz, fy.v * gz, gy * fz.v, fy.v * fz.v);
            nodes[i].off = make_uint4((unsigned)((z0 * P.scattering_mu_size + y0) * W), (unsigned)((z0 * P.scattering_mu_size + y1) * W),
                                      (unsigned)((z1 * P.scattering_mu_size + y0) * W), (unsigned)((z1 * P.scattering_mu_size + y1) * W));
            nodes[i].inv_r = (F(1.f) / r_i).v; nodes[i].pad0 = nodes[i].pad1 = nodes[i].pad2 = 0.f;
        }
    }
    // the mu_s mapping constants (scattering.h:48-56)
    const float bot = P.bottom_radius, top = P.top_radius;
    const float H2 = top * top - bot * bot, b2 = bot * bot, Hh = sqrtf(H2);
    const float dmin = top - bot, inv_span = 1.f / (Hh - dmin);
    const float Ac = -2.f * P.mu_s_min * bot / (Hh - dmin), minvA = -1.f / Ac;
    // u * MS - 0.5 with u = 0.5/MS + xx * (1 - 1/MS)  ==  xx * (MS - 1)
    const float msm1 = (float)(MS - 1);
    const float tmax = __int_as_float(__float_as_int(msm1) - 1);
    for (int c0 = 0; c0 < NS; c0 += CH) {
        const int cn = min(CH, NS - c0);
        __syncthreads();                          // nodes ready (first pass) / previous chunk consumed
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
            const int x = threadIdx.x + k * blockDim.x;
            if (x < W) {
                const uint32_t xi = (uint32_t)x;
                // stage: (mu, r)-bilinear of the density table at this x, per node.  The four row loads of FB_MS_U nodes
                // are issued before the first blend, so 4 * FB_MS_U loads are in flight per thread (one node at a time
                // left the loop waiting on L2 latency every iteration).
                for (int e0 = 0; e0 < cn; e0 += FB_MS_U) {
                    uint2 raw[FB_MS_U][4];
#pragma unroll
                    for (int u = 0; u < FB_MS_U; ++u) {
                        const uint4 o = nodes[c0 + min(e0 + u, cn - 1)].off;   // 32-bit texel indices: one IMAD.WIDE per address
                        raw[u][0] = __ldg(dens + (o.x + xi)); raw[u][1] = __ldg(dens + (o.y + xi));
                        raw[u][2] = __ldg(dens + (o.z + xi)); raw[u][3] = __ldg(dens + (o.w + xi));
                    }
#pragma unroll
                    for (int u = 0; u < FB_MS_U; ++u) {
                        if (e0 + u < cn) {
                            const float4 w = nodes[c0 + e0 + u].w;
                            const float4 a00 = unpack_half4(raw[u][0]), a10 = unpack_half4(raw[u][1]);
                            const float4 a01 = unpack_half4(raw[u][2]), a11 = unpack_half4(raw[u][3]);
                            float4 v;
                            v.x = fmaf(a11.x, w.w, fmaf(a01.x, w.z, fmaf(a10.x, w.y, a00.x * w.x)));
                            v.y = fmaf(a11.y, w.w, fmaf(a01.y, w.z, fmaf(a10.y, w.y, a00.y * w.x)));
                            v.z = fmaf(a11.z, w.w, fmaf(a01.z, w.z, fmaf(a10.z, w.y, a00.z * w.x)));
                            v.w = 0.f;
                            slab[(e0 + u) * W + x] = v;
                        }
                    }
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
            const int xk = threadIdx.x + k * blockDim.x;
            const bool act = xk < W && leader[k] == xk;
            const bool any_two = __any_sync(0xffffffffu, act && two[k]);   // warp-uniform: does any lane need slice 2
            if (act) {
#define FB_MS_SAMPLE(TWO)                                                                                                   \
                for (int e = 0; e < cn; ++e) {                                                                              \
                    const float4 nt = nodes[c0 + e].t;                                                                      \
                    const float inv_r = nodes[c0 + e].inv_r;                                                                \
                    const float mus_i = fminf(fmaxf(fmaf(nt.w, nuf[k], rmus[k]) * inv_r, -1.f), 1.f);   /* :38 */           \
                    /* DistanceToTopAtmosphereBoundary(bottom, mu_s_i), params.h:105-110: b^2 (mu^2 - 1) + top^2 > 0 */     \
                    const float dd = fmaf(-bot, mus_i, sqrt_fast(fmaf(b2, mus_i * mus_i, H2)));                             \
                    const float aa = (dd - dmin) * inv_span;                                                                \
                    const float xx = fmaxf(fmaf(aa, minvA, 1.f), 0.f) * rcp_fast(1.f + aa);   /* 1 + a is in [1, 2] */      \
                    const float t = fminf(fmaxf(xx * msm1, 0.f), tmax);                                                     \
                    const float tm = __fadd_rd(t, 8388608.f);                                                               \
                    const int j = __float_as_int(tm) - 0x4B000000;                                                          \
                    const float fx = t - (tm - 8388608.f);                                                                  \
                    const float4* s = slab + e * W + j;                                                                     \
                    const float4 p00 = s[kx0[k]], p01 = s[kx0[k] + 1];                                                      \
                    float vr = fmaf(fx, p01.x - p00.x, p00.x), vg = fmaf(fx, p01.y - p00.y, p00.y), vb = fmaf(fx, p01.z - p00.z, p00.z); \
                    if (TWO) {                                                                                              \
                        const float4 p10 = s[kx1[k]], p11 = s[kx1[k] + 1];                                                  \
                        const float v1r = fmaf(fx, p11.x - p10.x, p10.x), v1g = fmaf(fx, p11.y - p10.y, p10.y), v1b = fmaf(fx, p11.z - p10.z, p10.z); \
                        vr = fmaf(ln[k], v1r - vr, vr); vg = fmaf(ln[k], v1g - vg, vg); vb = fmaf(ln[k], v1b - vb, vb);     \
                    }                                                                                                       \
                    ar[k] = fmaf(vr, nt.x, ar[k]);                                                                          \
                    ag[k] = fmaf(vg, nt.y, ag[k]);                                                                          \
                    ab[k] = fmaf(vb, nt.z, ab[k]);                                                                          \
                }
                if (any_two) {
#pragma unroll CU
                    FB_MS_SAMPLE(two[k])
                } else {
#pragma unroll CU
                    FB_MS_SAMPLE(false)
                }
#undef FB_MS_SAMPLE
            }
        }
    }
    __syncthreads();                              // last chunk consumed: the slab now carries the leaders' results
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        const int x = threadIdx.x + k * blockDim.x;
        if (x < W && leader[k] == x) slab[x] = make_float4(ar[k], ag[k], ab[k], 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        const int x = threadIdx.x + k * blockDim.x;
        if (x >= W) continue;
        if (leader[k] != x) { const float4 v = slab[leader[k]]; ar[k] = v.x; ag[k] = v.y; ab[k] = v.z; }
        const size_t o = ((size_t)z * P.scattering_mu_size + y) * W + x;
        dMS[o] = pack_half4(ar[k], ag[k], ab[k], 0.f);                                  // multiple_scattering.comp:91
        const F pr = A<F>::RayleighPhase(F(nuf[k]));                                    // :92
        const float4 s = unpack_half4(S[o]);
        S[o] = pack_half4((F(ar[k]) / pr + F(s.x)).v, (F(ag[k]) / pr + F(s.y)).v, (F(ab[k]) / pr + F(s.z)).v, __fadd_rn(0.f, s.w));
    }
}

template <int TPT, int NTMAX>
static cudaError_t multiple_launch(const LaunchCtx& c, int nt, int CH, size_t smem, int r0, int r1) {
    cudaError_t e = cudaFuncSetAttribute(k_multiple_scattering<TPT, NTMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 g(1, c.P.scattering_mu_size, r1 - r0);
    k_multiple_scattering<TPT, NTMAX><<<g, nt, smem, c.stream>>>(c.P, texT(c), c.img.scattering_density, c.img.delta_multiple_scattering,
                                                          c.img.scattering, r0, CH);
    return cudaGetLastError();
}

// what the row-shared multiple-scattering kernel covers: a (nu, mu_s) row of at most 8192 texels whose staging slab fits
// in shared memory (W = 8192: 17 KiB of node records + 128 KiB of slab).  Wider rows run the transcription (reported
// through fb_params_slow_stages / fb_pending_slow_stages).
// nodes staged per pass.  Rows of up to 1024 texels: a slab of <= 48 KiB, so that four CTAs share an SM.  Wider rows
// run one 1024-thread CTA per SM anyway: their slab takes what the SM has (up to 3 nodes in <= 196 KiB), which cuts the
// stage / consume barrier pairs of a row from 51 to 17 (FUZZYBLUE_B200_MS_WIDE_CH overrides, for measurements).
static int multiple_chunk(int W) {
    int CH = 3072 / W;
    if (W > 1024) {
        static const int forced = [] { const char* e = std::getenv("FUZZYBLUE_B200_MS_WIDE_CH"); return e ? std::atoi(e) : 0; }();
        const int fit = (int)((196 * 1024) / ((size_t)W * sizeof(float4)));
        CH = std::max(CH, std::min(forced > 0 ? forced : 3, fit));
    }
    return CH < 1 ? 1 : (CH > 17 ? 17 : CH);
}
bool multiple_is_fast(const FbParams& P) {
    const int W = P.scattering_nu_size * P.scattering_mu_s_size;
    if (W <= 0 || W > 8192 || P.scattering_mu_s_size < 2) return false;   // W <= 0: a block fb_params_validate rejects
    return sizeof(MultiNode) * NS + (size_t)multiple_chunk(W) * W * sizeof(float4) <= 200 * 1024;
}

cudaError_t multiple_scattering(const LaunchCtx& c, int r0, int r1) {
    const int W = c.P.scattering_nu_size * c.P.scattering_mu_s_size;
    if (!multiple_is_fast(c.P)) return ref::multiple_scattering(c, r0, r1);
    // default dims: 256 threads, 1 texel each; larger rows: up to 1024 threads x {1, 2, 4, 8} texels
    const int nt = W <= 1024 ? ((W + 31) / 32) * 32 : 1024;
    const int tpt = (W + nt - 1) / nt;
    const int CH = multiple_chunk(W);             // nodes staged per pass
    const size_t smem = sizeof(MultiNode) * NS + (size_t)CH * W * sizeof(float4);
    if (nt <= 256) return multiple_launch<1, 256>(c, nt, CH, smem, r0, r1);
    if (tpt == 1) return multiple_launch<1, 1024>(c, nt, CH, smem, r0, r1);
    if (tpt == 2) return multiple_launch<2, 1024>(c, nt, CH, smem, r0, r1);
    if (tpt <= 4) return multiple_launch<4, 1024>(c, nt, CH, smem, r0, r1);
    return multiple_launch<8, 1024>(c, nt, CH, smem, r0, r1);
}

int launches_per_stage(const FbParams& P, int stage, int r_count) {
    if (stage != FB_STAGE_SCATTERING_DENSITY || !density_supported(P)) return 1;
    return 1 + (r_count + GN_MAXR - 1) / GN_MAXR;   // preparation + one main launch per GN_MAXR levels
}

}  // namespace fast
}  // namespace fb
