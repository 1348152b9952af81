// fb_kernels_ref.cu — FB_KERNELS_REFERENCE: one thread per texel, every shader statement kept,
// all arithmetic in fb::xf (un-fusable fp32).  This family exists as the on-device statement of
// "the reference's shaders in fp32": it is what the restructured kernels in fb_kernels_fast.cu are
// diffed against at full LUT sizes, where the CPU oracle takes minutes.
//
// Grid shapes differ from the reference's 8x8 / 4x4x4 workgroups on purpose: x (nu*mu_s) is the
// fastest-varying image axis, so a warp covers 32 consecutive x texels and the fp16x4 / fp32x4
// stores coalesce into full 256 B / 512 B segments.  Unlike the reference's truncating dispatch
// (src/precompute.rs:1740-1745) every texel is written even when a size is not a multiple of the
// workgroup.
#include "fb_kernels.h"
#include "fb_shader_math.cuh"

#include <math.h>

namespace fb {

void make_trig(Trig* t) {
    const float PI = FB_PI_F;
    {
        const float dtheta = PI / 16.f, dphi = PI / 16.f;   // scattering_density.comp:34-36
        for (int l = 0; l < 16; ++l) {
            float th = ((float)l + 0.5f) * dtheta;
            t->ct16[l] = cosf(th);
            t->st16[l] = sinf(th);
        }
        for (int m = 0; m < 32; ++m) {
            float ph = ((float)m + 0.5f) * dphi;
            t->cp32[m] = cosf(ph);
            t->sp32[m] = sinf(ph);
        }
    }
    {
        const float dtheta = PI / 32.f, dphi = PI / 32.f;   // indirect_irradiance.comp:20-22
        for (int j = 0; j < 16; ++j) {
            float th = ((float)j + 0.5f) * dtheta;
            t->ct32[j] = cosf(th);
            t->st32[j] = sinf(th);
        }
        for (int i = 0; i < 64; ++i) {
            float ph = ((float)i + 0.5f) * dphi;
            t->cp64[i] = cosf(ph);
            t->sp64[i] = sinf(ph);
        }
    }
}

namespace ref {

typedef xf F;
typedef V3<xf> V;

// ---- transmittance.comp ----------------------------------------------------------------------
__device__ F optical_length(const A<F>& a, const FbDensityProfile& prof, F r, F mu) {   // :8-32
    const int N = 500;
    F dx = a.DistanceToTop(r, mu) / F((float)N);
    F acc = F(0.f);
    for (int i = 0; i <= N; ++i) {
        F d_i = F((float)i) * dx;
        F r_i = f_sqrt(d_i * d_i + F(2.f) * r * mu * d_i + r * r);
        F y_i = A<F>::ProfileDensity(prof, r_i - a.bottom());
        F w_i = (i == 0 || i == N) ? F(0.5f) : F(1.f);
        acc += y_i * w_i * dx;
    }
    return acc;
}

__global__ void __launch_bounds__(64) k_transmittance(const __grid_constant__ FbParams P, float4* __restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.transmittance_mu_size || y >= P.transmittance_r_size) return;   // :72-74
    A<F> a(P);
    F r, mu;
    a.RMuFromUnitRanges(F((float)x) / F((float)(P.transmittance_mu_size - 1)),
                        F((float)y) / F((float)(P.transmittance_r_size - 1)), r, mu);
    V tau = V(P.rayleigh_scattering) * optical_length(a, P.rayleigh_density, r, mu) +          // :34-46
            V(P.mie_extinction) * optical_length(a, P.mie_density, r, mu) +
            V(P.absorption_extinction) * optical_length(a, P.absorption_density, r, mu);
    out[(size_t)y * P.transmittance_mu_size + x] =
        make_float4(expf(-tau.x.v), expf(-tau.y.v), expf(-tau.z.v), 1.f);
}

// ---- direct_irradiance.comp ------------------------------------------------------------------
__global__ void k_direct_irradiance(const __grid_constant__ FbParams P, Tex2 T, float4* __restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.irradiance_mu_s_size || y >= P.irradiance_r_size) return;
    A<F> a(P);
    F r, mu_s;
    a.RMuSFromIrradianceUnit(F((float)x) / F((float)(P.irradiance_mu_s_size - 1)),
                             F((float)y) / F((float)(P.irradiance_r_size - 1)), r, mu_s);
    F al = F(P.sun_angular_radius);                                                           // :17-22
    F avg = mu_s < -al ? F(0.f) : (mu_s > al ? mu_s : (mu_s + al) * (mu_s + al) / (F(4.f) * al));
    V e = V(P.solar_irradiance) * a.TransmittanceToTop(T, r, mu_s) * avg;
    out[(size_t)y * P.irradiance_mu_s_size + x] = make_float4(e.x.v, e.y.v, e.z.v, 0.f);
}

// ---- single_scattering.comp ------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_single_scattering(const __grid_constant__ FbParams P, Tex2 T, uint2* __restrict__ dR,
                                                           uint2* __restrict__ dM, uint2* __restrict__ S, int r0) {
    const int W = P.scattering_nu_size * P.scattering_mu_s_size;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = r0 + blockIdx.z;
    if (x >= W) return;
    A<F> a(P);
    F r, mu, mu_s, nu;
    bool hits;
    a.TexelToRMuMuSNu(x, y, z, r, mu, mu_s, nu, hits);
    const int N = 50;                                                                          // :42-65
    F dx = a.DistanceToNearest(r, mu, hits) / F((float)N);
    V rs(F(0.f)), ms(F(0.f));
    for (int i = 0; i <= N; ++i) {
        F d = F((float)i) * dx;
        F r_d = a.ClampRadius(f_sqrt(d * d + F(2.f) * r * mu * d + r * r));                    // :16-27
        F mu_s_d = A<F>::ClampCosine((r * mu_s + d * nu) / r_d);
        V t = a.Transmittance(T, r, mu, d, hits) * a.TransmittanceToSun(T, r_d, mu_s_d);
        V ri = t * A<F>::ProfileDensity(P.rayleigh_density, r_d - a.bottom());
        V mi = t * A<F>::ProfileDensity(P.mie_density, r_d - a.bottom());
        F w = (i == 0 || i == N) ? F(0.5f) : F(1.f);
        rs = rs + ri * w;
        ms = ms + mi * w;
    }
    V ray = rs * dx * V(P.solar_irradiance) * V(P.rayleigh_scattering);
    V mie = ms * dx * V(P.solar_irradiance) * V(P.mie_scattering);
    size_t o = ((size_t)z * P.scattering_mu_size + y) * W + x;
    dR[o] = pack_half4(ray.x.v, ray.y.v, ray.z.v, 0.f);                                        // :98-100
    dM[o] = pack_half4(mie.x.v, mie.y.v, mie.z.v, 0.f);
    S[o] = pack_half4(ray.x.v, ray.y.v, ray.z.v, mie.x.v);
}

// ---- scattering_density.comp -----------------------------------------------------------------
__global__ void __launch_bounds__(128) k_scattering_density(const __grid_constant__ FbParams P, const __grid_constant__ Trig tg,
                                                            Tex2 T, Tex3 dR, Tex3 dM, Tex3 dMS, Tex2 dE, int order,
                                                            uint2* __restrict__ out, int r0) {
    const int W = P.scattering_nu_size * P.scattering_mu_s_size;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = r0 + blockIdx.z;
    if (x >= W) return;
    A<F> a(P);
    F r, mu, mu_s, nu;
    bool hits_unused;
    a.TexelToRMuMuSNu(x, y, z, r, mu, mu_s, nu, hits_unused);
    V omega(f_sqrt(F(1.f) - mu * mu), F(0.f), mu);                                             // :28-32
    F sx = omega.x == F(0.f) ? F(0.f) : (nu - mu * mu_s) / omega.x;
    F sy = f_sqrt(f_max(F(1.f) - sx * sx - mu_s * mu_s, F(0.f)));
    V omega_s(sx, sy, mu_s);
    const F dphi = F(FB_PI_F) / F(16.f), dtheta = F(FB_PI_F) / F(16.f);
    F ray_rho = A<F>::ProfileDensity(P.rayleigh_density, r - a.bottom());                     // :94-97
    F mie_rho = A<F>::ProfileDensity(P.mie_density, r - a.bottom());
    V acc(F(0.f));
    for (int l = 0; l < 16; ++l) {
        F ct = F(tg.ct16[l]), st = F(tg.st16[l]);
        bool hits = a.RayIntersectsGround(r, ct);
        F dist_ground = F(0.f);
        V t_ground(F(0.f)), albedo(F(0.f));
        if (hits) {                                                                            // :53-60
            dist_ground = a.DistanceToBottom(r, ct);
            t_ground = a.Transmittance(T, r, ct, dist_ground, true);
            albedo = V(P.ground_albedo);
        }
        for (int m = 0; m < 32; ++m) {
            V wi(F(tg.cp32[m]) * st, F(tg.sp32[m]) * st, ct);
            F dw = dtheta * dphi * st;
            F nu1 = dot(omega_s, wi);
            V L = a.ScatteringOrder(dR, dM, dMS, r, wi.z, mu_s, nu1, hits, order - 1);
            V gn = V(F(0.f), F(0.f), r) + wi * dist_ground;                                    // :81-87
            gn = gn / f_sqrt(dot(gn, gn));
            V gE = a.Irradiance(dE, a.bottom(), dot(gn, omega_s));
            L = L + t_ground * albedo * (F(1.f) / F(FB_PI_F)) * gE;
            F nu2 = dot(omega, wi);
            acc = acc + L * (V(P.rayleigh_scattering) * ray_rho * A<F>::RayleighPhase(nu2) +
                             V(P.mie_scattering) * mie_rho * A<F>::MiePhase(F(P.mie_phase_function_g), nu2)) * dw;
        }
    }
    out[((size_t)z * P.scattering_mu_size + y) * W + x] = pack_half4(acc.x.v, acc.y.v, acc.z.v, 0.f);
}

// ---- indirect_irradiance.comp ----------------------------------------------------------------
__global__ void __launch_bounds__(32) k_indirect_irradiance(const __grid_constant__ FbParams P, const __grid_constant__ Trig tg,
                                                            Tex3 dR, Tex3 dM, Tex3 dMS, int order, float4* __restrict__ dE,
                                                            float4* __restrict__ E, int row0) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = row0 + blockIdx.y;
    if (x >= P.irradiance_mu_s_size || y >= P.irradiance_r_size) return;
    A<F> a(P);
    F r, mu_s;
    a.RMuSFromIrradianceUnit(F((float)x) / F((float)(P.irradiance_mu_s_size - 1)),
                             F((float)y) / F((float)(P.irradiance_r_size - 1)), r, mu_s);
    const F dphi = F(FB_PI_F) / F(32.f), dtheta = F(FB_PI_F) / F(32.f);
    V omega_s(f_sqrt(F(1.f) - mu_s * mu_s), F(0.f), mu_s);
    V acc(F(0.f));
    for (int j = 0; j < 16; ++j) {
        F ct = F(tg.ct32[j]), st = F(tg.st32[j]);
        for (int i = 0; i < 64; ++i) {
            V w(F(tg.cp64[i]) * st, F(tg.sp64[i]) * st, ct);
            F dw = dtheta * dphi * st;
            F nu = dot(w, omega_s);
            acc = acc + a.ScatteringOrder(dR, dM, dMS, r, w.z, mu_s, nu, false, order) * w.z * dw;
        }
    }
    size_t o = (size_t)y * P.irradiance_mu_s_size + x;
    dE[o] = make_float4(acc.x.v, acc.y.v, acc.z.v, 0.f);                                       // :72
    float4 e = E[o];                                                                           // :73
    E[o] = make_float4(__fadd_rn(acc.x.v, e.x), __fadd_rn(acc.y.v, e.y), __fadd_rn(acc.z.v, e.z), __fadd_rn(0.f, e.w));
}

// ---- multiple_scattering.comp ----------------------------------------------------------------
__global__ void __launch_bounds__(128) k_multiple_scattering(const __grid_constant__ FbParams P, Tex2 T, Tex3 dens,
                                                             uint2* __restrict__ dMS, uint2* __restrict__ S, int r0) {
    const int W = P.scattering_nu_size * P.scattering_mu_s_size;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = r0 + blockIdx.z;
    if (x >= W) return;
    A<F> a(P);
    F r, mu, mu_s, nu;
    bool hits;
    a.TexelToRMuMuSNu(x, y, z, r, mu, mu_s, nu, hits);
    const int N = 50;                                                                          // :21-53
    F dx = a.DistanceToNearest(r, mu, hits) / F((float)N);
    V acc(F(0.f));
    for (int i = 0; i <= N; ++i) {
        F d = F((float)i) * dx;
        F r_i = a.ClampRadius(f_sqrt(d * d + F(2.f) * r * mu * d + r * r));
        F mu_i = A<F>::ClampCosine((r * mu + d) / r_i);
        F mu_s_i = A<F>::ClampCosine((r * mu_s + d * nu) / r_i);
        V v = a.Scattering4(dens, r_i, mu_i, mu_s_i, nu, hits).rgb() * a.Transmittance(T, r, mu, d, hits) * dx;
        F w = (i == 0 || i == N) ? F(0.5f) : F(1.f);
        acc = acc + v * w;
    }
    size_t o = ((size_t)z * P.scattering_mu_size + y) * W + x;
    dMS[o] = pack_half4(acc.x.v, acc.y.v, acc.z.v, 0.f);                                       // :91
    F pr = A<F>::RayleighPhase(nu);                                                            // :92
    float4 s = unpack_half4(S[o]);
    S[o] = pack_half4(((acc.x / pr) + F(s.x)).v, ((acc.y / pr) + F(s.y)).v, ((acc.z / pr) + F(s.z)).v, __fadd_rn(0.f, s.w));
}

// ---- launchers ----------------------------------------------------------------------------------
static inline Tex2 tex2(const float4* p, int w, int h) { Tex2 t; t.p = p; t.w = w; t.h = h; return t; }
static inline Tex3 tex3(const uint2* p, int w, int h, int d) { Tex3 t; t.p = p; t.w = w; t.h = h; t.d = d; return t; }
static inline Tex2 texT(const LaunchCtx& c) { return tex2(c.img.transmittance, c.P.transmittance_mu_size, c.P.transmittance_r_size); }
static inline Tex3 texS(const LaunchCtx& c, const uint2* p) {
    return tex3(p, c.P.scattering_nu_size * c.P.scattering_mu_s_size, c.P.scattering_mu_size, c.P.scattering_r_size);
}
static inline dim3 grid3(const LaunchCtx& c, int block, int r0, int r1) {
    int W = c.P.scattering_nu_size * c.P.scattering_mu_s_size;
    return dim3((W + block - 1) / block, c.P.scattering_mu_size, r1 - r0);
}

cudaError_t transmittance(const LaunchCtx& c) {
    dim3 g((c.P.transmittance_mu_size + 63) / 64, c.P.transmittance_r_size);
    k_transmittance<<<g, 64, 0, c.stream>>>(c.P, c.img.transmittance);
    return cudaGetLastError();
}
cudaError_t direct_irradiance(const LaunchCtx& c) {
    dim3 g((c.P.irradiance_mu_s_size + 63) / 64, c.P.irradiance_r_size);
    k_direct_irradiance<<<g, 64, 0, c.stream>>>(c.P, texT(c), c.img.delta_irradiance);
    return cudaGetLastError();
}
cudaError_t single_scattering(const LaunchCtx& c, int r0, int r1) {
    k_single_scattering<<<grid3(c, 128, r0, r1), 128, 0, c.stream>>>(c.P, texT(c), c.img.delta_rayleigh, c.img.delta_mie,
                                                                      c.img.scattering, r0);
    return cudaGetLastError();
}
cudaError_t scattering_density(const LaunchCtx& c, int order, int r0, int r1) {
    k_scattering_density<<<grid3(c, 128, r0, r1), 128, 0, c.stream>>>(
        c.P, c.trig, texT(c), texS(c, c.img.delta_rayleigh), texS(c, c.img.delta_mie), texS(c, c.img.delta_multiple_scattering),
        tex2(c.img.delta_irradiance, c.P.irradiance_mu_s_size, c.P.irradiance_r_size), order, c.img.scattering_density, r0);
    return cudaGetLastError();
}
cudaError_t indirect_irradiance(const LaunchCtx& c, int order, int row0, int row1) {
    dim3 g((c.P.irradiance_mu_s_size + 31) / 32, row1 - row0);
    k_indirect_irradiance<<<g, 32, 0, c.stream>>>(c.P, c.trig, texS(c, c.img.delta_rayleigh), texS(c, c.img.delta_mie),
                                                  texS(c, c.img.delta_multiple_scattering), order, c.img.delta_irradiance,
                                                  c.img.irradiance, row0);
    return cudaGetLastError();
}
cudaError_t multiple_scattering(const LaunchCtx& c, int r0, int r1) {
    k_multiple_scattering<<<grid3(c, 128, r0, r1), 128, 0, c.stream>>>(c.P, texT(c), texS(c, c.img.scattering_density),
                                                                        c.img.delta_multiple_scattering, c.img.scattering, r0);
    return cudaGetLastError();
}

}  // namespace ref
}  // namespace fb
