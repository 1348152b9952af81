// fb_shader_math.cuh — device-side restatement of the reference's shared shader library
// (shaders/{util,params,transmittance,scattering,irradiance,render_sky,render_lighting}.h),
// templated on the scalar type F:
//
//   F = xf     every + - * / is one correctly rounded fp32 operation that ptxas can NOT contract
//              into an FMA (__fadd_rn/__fmul_rn/...).  This is "the GLSL as written, in fp32".
//              The reference's formulas cancel catastrophically on horizon-grazing rays
//              (r*r*(mu*mu-1)+bottom*bottom, r*r-bottom*bottom, ...: a different rounding of one
//              product moves a ray length by 1-2 km and a texel by several per cent), so every
//              geometry quantity that feeds a table is evaluated in xf.
//   F = float  ordinary fp32, contraction allowed — only for well-conditioned hot loops.
//
// Each function cites the GLSL it follows (file:line, relative to the reference checkout).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fuzzyblue.h"

namespace fb {

// ---------------------------------------------------------------------------------------------
// xf: fp32 with un-fusable arithmetic
// ---------------------------------------------------------------------------------------------
struct xf {
    float v;
    __device__ __forceinline__ xf() {}
    __device__ __forceinline__ constexpr xf(float x) : v(x) {}
};
#ifdef FB_XF_CONTRACT_EXPERIMENT   // measurement only: what exactness costs the sky evaluation
__device__ __forceinline__ xf operator+(xf a, xf b) { return xf(a.v + b.v); }
__device__ __forceinline__ xf operator-(xf a, xf b) { return xf(a.v - b.v); }
__device__ __forceinline__ xf operator*(xf a, xf b) { return xf(a.v * b.v); }
#else
__device__ __forceinline__ xf operator+(xf a, xf b) { return xf(__fadd_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator-(xf a, xf b) { return xf(__fsub_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator*(xf a, xf b) { return xf(__fmul_rn(a.v, b.v)); }
#endif
__device__ __forceinline__ xf operator/(xf a, xf b) { return xf(__fdiv_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator-(xf a) { return xf(-a.v); }
__device__ __forceinline__ xf& operator+=(xf& a, xf b) { a = a + b; return a; }
__device__ __forceinline__ bool operator<(xf a, xf b) { return a.v < b.v; }
__device__ __forceinline__ bool operator>(xf a, xf b) { return a.v > b.v; }
__device__ __forceinline__ bool operator<=(xf a, xf b) { return a.v <= b.v; }
__device__ __forceinline__ bool operator>=(xf a, xf b) { return a.v >= b.v; }
__device__ __forceinline__ bool operator==(xf a, xf b) { return a.v == b.v; }

__device__ __forceinline__ float raw(float a) { return a; }
__device__ __forceinline__ float raw(xf a) { return a.v; }

__device__ __forceinline__ float f_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ xf f_sqrt(xf a) { return xf(__fsqrt_rn(a.v)); }
__device__ __forceinline__ float f_floor(float a) { return floorf(a); }
__device__ __forceinline__ xf f_floor(xf a) { return xf(floorf(a.v)); }
__device__ __forceinline__ float f_exp(float a) { return expf(a); }
__device__ __forceinline__ xf f_exp(xf a) { return xf(expf(a.v)); }
__device__ __forceinline__ float f_sin(float a) { return sinf(a); }
__device__ __forceinline__ xf f_sin(xf a) { return xf(sinf(a.v)); }
__device__ __forceinline__ float f_cos(float a) { return cosf(a); }
__device__ __forceinline__ xf f_cos(xf a) { return xf(cosf(a.v)); }
__device__ __forceinline__ float f_pow15(float a) { return powf(a, 1.5f); }
__device__ __forceinline__ xf f_pow15(xf a) { return xf(powf(a.v, 1.5f)); }
__device__ __forceinline__ float f_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ xf f_min(xf a, xf b) { return xf(fminf(a.v, b.v)); }
__device__ __forceinline__ float f_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ xf f_max(xf a, xf b) { return xf(fmaxf(a.v, b.v)); }
__device__ __forceinline__ bool f_isinf(float a) { return isinf(a); }
__device__ __forceinline__ bool f_isinf(xf a) { return isinf(a.v); }
template <class F> __device__ __forceinline__ F f_clamp(F v, F lo, F hi) { return f_min(f_max(v, lo), hi); }
template <class F> __device__ __forceinline__ F f_smoothstep(F e0, F e1, F x) {
    F t = f_clamp<F>((x - e0) / (e1 - e0), F(0.f), F(1.f));
    return t * t * (F(3.f) - F(2.f) * t);
}

template <class F> struct V3 {
    F x, y, z;
    __device__ __forceinline__ V3() {}
    __device__ __forceinline__ V3(F a, F b, F c) : x(a), y(b), z(c) {}
    __device__ __forceinline__ explicit V3(F a) : x(a), y(a), z(a) {}
    __device__ __forceinline__ explicit V3(const float* p) : x(F(p[0])), y(F(p[1])), z(F(p[2])) {}
};
template <class F> __device__ __forceinline__ V3<F> operator+(V3<F> a, V3<F> b) { return V3<F>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class F> __device__ __forceinline__ V3<F> operator-(V3<F> a, V3<F> b) { return V3<F>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class F> __device__ __forceinline__ V3<F> operator*(V3<F> a, V3<F> b) { return V3<F>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <class F> __device__ __forceinline__ V3<F> operator/(V3<F> a, V3<F> b) { return V3<F>(a.x / b.x, a.y / b.y, a.z / b.z); }
template <class F> __device__ __forceinline__ V3<F> operator*(V3<F> a, F s) { return V3<F>(a.x * s, a.y * s, a.z * s); }
template <class F> __device__ __forceinline__ V3<F> operator/(V3<F> a, F s) { return V3<F>(a.x / s, a.y / s, a.z / s); }
template <class F> __device__ __forceinline__ F dot(V3<F> a, V3<F> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class F> __device__ __forceinline__ V3<F> min1(V3<F> a) {
    return V3<F>(f_min(a.x, F(1.f)), f_min(a.y, F(1.f)), f_min(a.z, F(1.f)));
}
template <class F> struct V4 {
    F x, y, z, w;
    __device__ __forceinline__ V3<F> rgb() const { return V3<F>(x, y, z); }
};
template <class F> __device__ __forceinline__ V4<F> operator+(V4<F> a, V4<F> b) { return V4<F>{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
template <class F> __device__ __forceinline__ V4<F> operator*(V4<F> a, F s) { return V4<F>{a.x * s, a.y * s, a.z * s, a.w * s}; }

// ---------------------------------------------------------------------------------------------
// Sampled images: linear device memory, VkSampler{LINEAR, CLAMP_TO_EDGE, normalised}
// (src/precompute.rs:85-98) restated with exact fp32 weights.  The hardware texture unit is
// deliberately not used: its 8-bit fixed-point weights are ~2e-3 texels coarse, which on the
// horizon rows of the transmittance table (neighbouring texels a factor 100 apart) is far
// outside the 1e-3 parity budget (SURVEY.md §7 hard part 2).
// ---------------------------------------------------------------------------------------------
struct Tex2 {   // RGBA32F
    const float4* p; int w, h;
    template <class F> __device__ __forceinline__ V4<F> texel(int x, int y) const {
        float4 t = __ldg(p + (size_t)y * w + x);
        return V4<F>{F(t.x), F(t.y), F(t.z), F(t.w)};
    }
};
struct Tex3 {   // RGBA16F
    const uint2* p; int w, h, d;
    template <class F> __device__ __forceinline__ V4<F> texel(int x, int y, int z) const {
        uint2 t = __ldg(p + ((size_t)z * h + y) * w + x);
        float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
        float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
        return V4<F>{F(a.x), F(a.y), F(b.x), F(b.y)};
    }
};
template <class F> __device__ __forceinline__ void tex_axis(F u, int n, int& i0, int& i1, F& f) {
    F t = u * F((float)n) - F(0.5f);
    F fl = f_floor(t);
    f = t - fl;
    float c = fminf(fmaxf(raw(fl), -1.f), (float)n);
    int i = (int)c;
    i0 = min(max(i, 0), n - 1);
    i1 = min(max(i + 1, 0), n - 1);
}
template <class F> __device__ __forceinline__ V4<F> sample(const Tex2& t, F u, F v) {
    int x0, x1, y0, y1; F fx, fy;
    tex_axis(u, t.w, x0, x1, fx); tex_axis(v, t.h, y0, y1, fy);
    F gx = F(1.f) - fx, gy = F(1.f) - fy;
    V4<F> a = t.texel<F>(x0, y0) * gx + t.texel<F>(x1, y0) * fx;
    V4<F> b = t.texel<F>(x0, y1) * gx + t.texel<F>(x1, y1) * fx;
    return a * gy + b * fy;
}
template <class F> __device__ __forceinline__ V4<F> sample(const Tex3& t, F u, F v, F s) {
    int x0, x1, y0, y1, z0, z1; F fx, fy, fz;
    tex_axis(u, t.w, x0, x1, fx); tex_axis(v, t.h, y0, y1, fy); tex_axis(s, t.d, z0, z1, fz);
    F gx = F(1.f) - fx, gy = F(1.f) - fy, gz = F(1.f) - fz;
    V4<F> a = t.texel<F>(x0, y0, z0) * gx + t.texel<F>(x1, y0, z0) * fx;
    V4<F> b = t.texel<F>(x0, y1, z0) * gx + t.texel<F>(x1, y1, z0) * fx;
    V4<F> c = t.texel<F>(x0, y0, z1) * gx + t.texel<F>(x1, y0, z1) * fx;
    V4<F> e = t.texel<F>(x0, y1, z1) * gx + t.texel<F>(x1, y1, z1) * fx;
    V4<F> ab = a * gy + b * fy;
    V4<F> ce = c * gy + e * fy;
    return ab * gz + ce * fz;
}

// image stores in the image-view format (src/precompute.rs:1170,1191,1215)
__device__ __forceinline__ uint2 pack_half4(float a, float b, float c, float d) {
    __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&lo);
    r.y = *reinterpret_cast<uint32_t*>(&hi);
    return r;
}
__device__ __forceinline__ float4 unpack_half4(uint2 t) {
    float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

#define FB_PI_F 3.14159265358979323846f   /* util.h:4 */

// ---------------------------------------------------------------------------------------------
// The shader library proper.  A<F> wraps the uniform block (params.h:26-87).
// ---------------------------------------------------------------------------------------------
template <class F> struct A {
    const FbParams& P;
    __device__ __forceinline__ explicit A(const FbParams& p) : P(p) {}
    __device__ __forceinline__ F bottom() const { return F(P.bottom_radius); }
    __device__ __forceinline__ F top() const { return F(P.top_radius); }

    // util.h
    __device__ __forceinline__ static F ClampCosine(F mu) { return f_clamp<F>(mu, F(-1.f), F(1.f)); }   // :6-8
    __device__ __forceinline__ static F ClampDistance(F d) { return f_max(d, F(0.f)); }                 // :10-12
    __device__ __forceinline__ static F SafeSqrt(F a) { return f_sqrt(f_max(a, F(0.f))); }              // :14-16
    __device__ __forceinline__ static F CoordFromUnit(F x, int n) {                                    // :18-20
        return F(0.5f) / F((float)n) + x * (F(1.f) - F(1.f) / F((float)n));
    }
    __device__ __forceinline__ static F UnitFromCoord(F u, int n) {                                    // :22-24
        return (u - F(0.5f) / F((float)n)) / (F(1.f) - F(1.f) / F((float)n));
    }
    __device__ __forceinline__ static F RayleighPhase(F nu) {                                          // :26-29
        F k = F(3.f) / (F(16.f) * F(FB_PI_F));
        return k * (F(1.f) + nu * nu);
    }
    __device__ __forceinline__ static F MiePhase(F g, F nu) {                                          // :31-34
        F k = F(3.f) / (F(8.f) * F(FB_PI_F)) * (F(1.f) - g * g) / (F(2.f) + g * g);
        return k * (F(1.f) + nu * nu) / f_pow15(F(1.f) + g * g - F(2.f) * g * nu);
    }
    __device__ __forceinline__ static F FragCoordFromTexel(unsigned x, unsigned n) {                   // :36-38
        return F((float)n) * CoordFromUnit(F((float)x) / F((float)(n - 1)), (int)n);
    }

    // params.h:89-133
    __device__ __forceinline__ static F LayerDensity(const FbDensityProfileLayer& l, F h) {            // :89-93
        F d = F(l.exp_term) * f_exp(F(l.exp_scale) * h) + F(l.linear_term) * h + F(l.constant_term);
        return f_clamp<F>(d, F(0.f), F(1.f));
    }
    __device__ __forceinline__ static F ProfileDensity(const FbDensityProfile& p, F h) {               // :95-99
        return h < F(p.layers[0].width) ? LayerDensity(p.layers[0], h) : LayerDensity(p.layers[1], h);
    }
    __device__ __forceinline__ F ClampRadius(F r) const { return f_clamp<F>(r, bottom(), top()); }      // :101-103
    __device__ __forceinline__ F DistanceToTop(F r, F mu) const {                                      // :105-110
        F disc = r * r * (mu * mu - F(1.f)) + top() * top();
        return ClampDistance(-r * mu + SafeSqrt(disc));
    }
    __device__ __forceinline__ F DistanceToBottom(F r, F mu) const {                                   // :112-117
        F disc = r * r * (mu * mu - F(1.f)) + bottom() * bottom();
        return ClampDistance(-r * mu - SafeSqrt(disc));
    }
    __device__ __forceinline__ bool RayIntersectsGround(F r, F mu) const {                             // :119-124
        return mu < F(0.f) && r * r * (mu * mu - F(1.f)) + bottom() * bottom() >= F(0.f);
    }
    __device__ __forceinline__ F DistanceToNearest(F r, F mu, bool hits) const {                       // :126-133
        return hits ? DistanceToBottom(r, mu) : DistanceToTop(r, mu);
    }

    // transmittance.h
    __device__ __forceinline__ void TransmittanceUv(F r, F mu, F& u, F& v) const {                     // :7-24
        F H = f_sqrt(top() * top() - bottom() * bottom());
        F rho = SafeSqrt(r * r - bottom() * bottom());
        F d = DistanceToTop(r, mu);
        F d_min = top() - r;
        F d_max = rho + H;
        F x_mu = (d - d_min) / (d_max - d_min);
        F x_r = rho / H;
        u = CoordFromUnit(x_mu, P.transmittance_mu_size);
        v = CoordFromUnit(x_r, P.transmittance_r_size);
    }
    __device__ __forceinline__ V3<F> TransmittanceToTop(const Tex2& T, F r, F mu) const {              // :26-33
        F u, v;
        TransmittanceUv(r, mu, u, v);
        return sample<F>(T, u, v).rgb();
    }
    __device__ __forceinline__ V3<F> Transmittance(const Tex2& T, F r, F mu, F d, bool hits) const {   // :35-61
        F r_d = ClampRadius(f_sqrt(d * d + F(2.f) * r * mu * d + r * r));
        F mu_d = ClampCosine((r * mu + d) / r_d);
        V3<F> q = hits ? TransmittanceToTop(T, r_d, -mu_d) / TransmittanceToTop(T, r, -mu)
                       : TransmittanceToTop(T, r, mu) / TransmittanceToTop(T, r_d, mu_d);
        return min1(q);
    }
    __device__ __forceinline__ V3<F> TransmittanceToSun(const Tex2& T, F r, F mu_s) const {            // :63-74
        F sin_h = bottom() / r;
        F cos_h = -f_sqrt(f_max(F(1.f) - sin_h * sin_h, F(0.f)));
        F a = F(P.sun_angular_radius);
        return TransmittanceToTop(T, r, mu_s) * f_smoothstep<F>(-sin_h * a, sin_h * a, mu_s - cos_h);
    }

    // transmittance.comp:48-64
    __device__ __forceinline__ void RMuFromUnitRanges(F x_mu, F x_r, F& r, F& mu) const {
        F H = f_sqrt(top() * top() - bottom() * bottom());
        F rho = H * x_r;
        r = f_sqrt(rho * rho + bottom() * bottom());
        F d_min = top() - r;
        F d_max = rho + H;
        F d = d_min + x_mu * (d_max - d_min);
        mu = d == F(0.f) ? F(1.f) : (H * H - rho * rho - d * d) / (F(2.f) * r * d);
        mu = ClampCosine(mu);
    }

    // irradiance.h
    __device__ __forceinline__ void RMuSFromIrradianceUnit(F x_mu_s, F x_r, F& r, F& mu_s) const {     // :7-18
        r = bottom() + x_r * (top() - bottom());
        mu_s = ClampCosine(F(2.f) * x_mu_s - F(1.f));
    }
    __device__ __forceinline__ V3<F> Irradiance(const Tex2& E, F r, F mu_s) const {                    // :20-38
        F x_r = (r - bottom()) / (top() - bottom());
        F x_mu_s = mu_s * F(0.5f) + F(0.5f);
        return sample<F>(E, CoordFromUnit(x_mu_s, P.irradiance_mu_s_size), CoordFromUnit(x_r, P.irradiance_r_size)).rgb();
    }

    // scattering.h
    __device__ __forceinline__ void ScatteringUvwz(F r, F mu, F mu_s, F nu, bool hits, F uvwz[4]) const {   // :7-60
        F H = f_sqrt(top() * top() - bottom() * bottom());
        F rho = SafeSqrt(r * r - bottom() * bottom());
        F u_r = CoordFromUnit(rho / H, P.scattering_r_size);
        F r_mu = r * mu;
        F disc = r_mu * r_mu - r * r + bottom() * bottom();
        F u_mu;
        if (hits) {
            F d = -r_mu - SafeSqrt(disc);
            F d_min = r - bottom();
            F d_max = rho;
            u_mu = F(0.5f) - F(0.5f) * CoordFromUnit(d_max == d_min ? F(0.f) : (d - d_min) / (d_max - d_min),
                                                     P.scattering_mu_size / 2);
        } else {
            F d = -r_mu + SafeSqrt(disc + H * H);
            F d_min = top() - r;
            F d_max = rho + H;
            u_mu = F(0.5f) + F(0.5f) * CoordFromUnit((d - d_min) / (d_max - d_min), P.scattering_mu_size / 2);
        }
        F d = DistanceToTop(bottom(), mu_s);
        F d_min = top() - bottom();
        F d_max = H;
        F a = (d - d_min) / (d_max - d_min);
        F Ac = F(-2.f) * F(P.mu_s_min) * bottom() / (d_max - d_min);
        F u_mu_s = CoordFromUnit(f_max(F(1.f) - a / Ac, F(0.f)) / (F(1.f) + a), P.scattering_mu_s_size);
        F u_nu = (nu + F(1.f)) / F(2.f);
        uvwz[0] = u_nu; uvwz[1] = u_mu_s; uvwz[2] = u_mu; uvwz[3] = u_r;
    }
    __device__ __forceinline__ void RMuMuSNuFromUvwz(const F uvwz[4], F& r, F& mu, F& mu_s, F& nu, bool& hits) const {   // :62-114
        F H = f_sqrt(top() * top() - bottom() * bottom());
        F rho = H * UnitFromCoord(uvwz[3], P.scattering_r_size);
        r = f_sqrt(rho * rho + bottom() * bottom());
        if (uvwz[2] < F(0.5f)) {
            F d_min = r - bottom();
            F d_max = rho;
            F d = d_min + (d_max - d_min) * UnitFromCoord(F(1.f) - F(2.f) * uvwz[2], P.scattering_mu_size / 2);
            mu = d == F(0.f) ? F(-1.f) : ClampCosine(-(rho * rho + d * d) / (F(2.f) * r * d));
            hits = true;
        } else {
            F d_min = top() - r;
            F d_max = rho + H;
            F d = d_min + (d_max - d_min) * UnitFromCoord(F(2.f) * uvwz[2] - F(1.f), P.scattering_mu_size / 2);
            mu = d == F(0.f) ? F(1.f) : ClampCosine((H * H - rho * rho - d * d) / (F(2.f) * r * d));
            hits = false;
        }
        F x_mu_s = UnitFromCoord(uvwz[1], P.scattering_mu_s_size);
        F d_min = top() - bottom();
        F d_max = H;
        F Ac = F(-2.f) * F(P.mu_s_min) * bottom() / (d_max - d_min);
        F a = (Ac - x_mu_s * Ac) / (F(1.f) + x_mu_s * Ac);
        F d = d_min + f_min(a, Ac) * (d_max - d_min);
        mu_s = d == F(0.f) ? F(1.f) : ClampCosine((H * H - d * d) / (F(2.f) * bottom() * d));
        nu = ClampCosine(uvwz[0] * F(2.f) - F(1.f));
    }
    // GetScatteringFragCoord :181-190 + GetRMuMuSNuFromScatteringTextureFragCoord :116-137
    __device__ __forceinline__ void TexelToRMuMuSNu(unsigned x, unsigned y, unsigned z, F& r, F& mu, F& mu_s, F& nu,
                                                    bool& hits) const {
        F fx = FragCoordFromTexel(x, (unsigned)(P.scattering_nu_size * P.scattering_mu_s_size));
        F fy = FragCoordFromTexel(y, (unsigned)P.scattering_mu_size);
        F fz = FragCoordFromTexel(z, (unsigned)P.scattering_r_size);
        F ms = F((float)P.scattering_mu_s_size);
        F f_nu = f_floor(fx / ms);
        F f_mu_s = fx - ms * f_floor(fx / ms);   // GLSL mod()
        F uvwz[4] = {f_nu / F((float)(P.scattering_nu_size - 1)), f_mu_s / ms, fy / F((float)P.scattering_mu_size),
                     fz / F((float)P.scattering_r_size)};
        RMuMuSNuFromUvwz(uvwz, r, mu, mu_s, nu, hits);
        F s = f_sqrt((F(1.f) - mu * mu) * (F(1.f) - mu_s * mu_s));
        nu = f_clamp<F>(nu, mu * mu_s - s, mu * mu_s + s);
    }
    __device__ __forceinline__ V4<F> Scattering4(const Tex3& S, F r, F mu, F mu_s, F nu, bool hits) const {   // :139-155
        F uvwz[4];
        ScatteringUvwz(r, mu, mu_s, nu, hits, uvwz);
        F tcx = uvwz[0] * F((float)(P.scattering_nu_size - 1));
        F tx = f_floor(tcx);
        F l = tcx - tx;
        F nn = F((float)P.scattering_nu_size);
        V4<F> a = sample<F>(S, (tx + uvwz[1]) / nn, uvwz[2], uvwz[3]);
        V4<F> b = sample<F>(S, (tx + F(1.f) + uvwz[1]) / nn, uvwz[2], uvwz[3]);
        return a * (F(1.f) - l) + b * l;
    }
    __device__ __forceinline__ V3<F> ScatteringOrder(const Tex3& dR, const Tex3& dM, const Tex3& dMS, F r, F mu, F mu_s,
                                                     F nu, bool hits, int order) const {               // :157-179
        if (order == 1) {
            V3<F> ray = Scattering4(dR, r, mu, mu_s, nu, hits).rgb();
            V3<F> mie = Scattering4(dM, r, mu, mu_s, nu, hits).rgb();
            return ray * RayleighPhase(nu) + mie * MiePhase(F(P.mie_phase_function_g), nu);
        }
        return Scattering4(dMS, r, mu, mu_s, nu, hits).rgb();
    }

    // render_sky.h
    __device__ __forceinline__ V3<F> ExtrapolatedSingleMie(V4<F> s) const {                            // :9-19
        if (s.x <= F(0.f)) return V3<F>(F(0.f));
        V3<F> bR(P.rayleigh_scattering), bM(P.mie_scattering);
        return s.rgb() * s.w / s.x * (bR.x / bM.x) * (bM / bR);
    }
    __device__ __forceinline__ V3<F> CombinedScattering(const Tex3& S, F r, F mu, F mu_s, F nu, bool hits,
                                                        V3<F>& single_mie) const {                     // :21-43
        V4<F> c = Scattering4(S, r, mu, mu_s, nu, hits);
        single_mie = ExtrapolatedSingleMie(c);
        return c.rgb();
    }
    __device__ __forceinline__ V3<F> SkyRadianceToPoint(const Tex2& T, const Tex3& S, V3<F> camera, V3<F> view,
                                                        V3<F> point, V3<F> sun, V3<F>& transmittance) const {   // :111-191
        F r = f_sqrt(dot(camera, camera));
        F rmu = dot(camera, view);
        F to_top = -rmu - f_sqrt(rmu * rmu - r * r + top() * top());   // NaN when the ray misses
        if (to_top > F(0.f)) {
            camera = camera + view * to_top;
            r = top();
            rmu = rmu + to_top;
        } else if (r > top()) {
            transmittance = V3<F>(F(1.f));
            return V3<F>(F(0.f));
        }
        F mu = rmu / r;
        F mu_s = dot(camera, sun) / r;
        F nu = dot(view, sun);
        V3<F> pc = point - camera;
        F d = f_sqrt(dot(pc, pc));
        bool hits = RayIntersectsGround(r, mu);
        transmittance = Transmittance(T, r, mu, d, hits);
        V3<F> single_mie;
        V3<F> scat = CombinedScattering(S, r, mu, mu_s, nu, hits, single_mie);
        if (!f_isinf(d)) {
            F r_p = ClampRadius(f_sqrt(d * d + F(2.f) * r * mu * d + r * r));
            F mu_p = (r * mu + d) / r_p;
            F mu_s_p = (r * mu_s + d * nu) / r_p;
            V3<F> single_mie_p;
            V3<F> scat_p = CombinedScattering(S, r_p, mu_p, mu_s_p, nu, hits, single_mie_p);
            scat = scat - transmittance * scat_p;
            single_mie = single_mie - transmittance * single_mie_p;
            single_mie = ExtrapolatedSingleMie(V4<F>{scat.x, scat.y, scat.z, single_mie.x});
            single_mie = single_mie * f_smoothstep<F>(F(0.f), F(0.01f), mu_s);
        }
        return scat * RayleighPhase(nu) + single_mie * MiePhase(F(P.mie_phase_function_g), nu);
    }
    __device__ __forceinline__ V3<F> SkyRadiance(const Tex2& T, const Tex3& S, V3<F> camera, V3<F> view, V3<F> sun,
                                                 V3<F>& transmittance) const {                         // :45-109
        F r = f_sqrt(dot(camera, camera));
        F rmu = dot(camera, view);
        F to_top = -rmu - f_sqrt(rmu * rmu - r * r + top() * top());
        if (to_top > F(0.f)) {
            camera = camera + view * to_top;
            r = top();
            rmu = rmu + to_top;
        } else if (r > top()) {
            transmittance = V3<F>(F(1.f));
            return V3<F>(F(0.f));
        }
        F mu = rmu / r;
        F mu_s = dot(camera, sun) / r;
        F nu = dot(view, sun);
        bool hits = RayIntersectsGround(r, mu);
        transmittance = hits ? V3<F>(F(0.f)) : TransmittanceToTop(T, r, mu);
        V3<F> single_mie;
        V3<F> scat = CombinedScattering(S, r, mu, mu_s, nu, hits, single_mie);
        return scat * RayleighPhase(nu) + single_mie * MiePhase(F(P.mie_phase_function_g), nu);
    }
    // render_lighting.h:10-28
    __device__ __forceinline__ V3<F> SunAndSkyIrradiance(const Tex2& T, const Tex2& E, V3<F> point, V3<F> normal, V3<F> sun,
                                                         V3<F>& sky) const {
        F r = f_sqrt(dot(point, point));
        F mu_s = dot(point, sun) / r;
        sky = Irradiance(E, r, mu_s) * (F(1.f) + dot(normal, point) / r) * F(0.5f);
        return V3<F>(P.solar_irradiance) * TransmittanceToSun(T, r, mu_s) * f_max(dot(normal, sun), F(0.f));
    }
};

}  // namespace fb
