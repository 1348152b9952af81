// fb_oracle.cpp — CPU restatement of fuzzyblue's atmosphere precompute + sky evaluation.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it is imported
// by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
// legs as the *checker* and the *CPU baseline*.  The shipped path (fuzzyblue_b200/) never
// links, loads or calls it.
//
// PARITY PIN: the reference has no golden vectors or known-answer tests of its own, and the crate cannot be built
// in this image (no cargo / shaderc / Vulkan ICD) -- but its SHADERS can be run: oracle/glsl_ref compiles
// /root/reference/shaders/*.h, *.comp and render_sky.frag as C++ (a GLSL-subset header + a purely syntactic
// translator) into oracle/_ref/libfb_glsl_ref.so, and mode 0 of this file reproduces the output of that library BIT FOR
// BIT for every stage, the whole per-order schedule and the sky evaluation (tests/test_reference_glsl.py; the same pin
// as committed golden vectors generated from the reference's shaders: tests/golden/reference_*.npz,
// tests/test_golden_cpu.py).  What both sides share by construction is what Vulkan leaves to the driver: binary32
// arithmetic with the C library's exp / pow / sin / cos, the LINEAR / CLAMP_TO_EDGE sampler with exact weights, and
// round-to-nearest-even binary16 image stores.  Independently of the reference it is pinned by closed-form
// known-answer tests for all six stages (tests/test_oracle_kat.py, tests/test_kat_3d.py) and cross-checked against a
// second scalar numpy restatement (oracle/numpy_check.py, tests/test_oracle_numpy.py).
//
// Every routine restates one GLSL function of /root/reference/shaders (cited per function
// as file:line) in a scalar type R chosen by the caller:
//   mode 0  R = float , tables quantised where the reference stores them
//                       (RGBA32F 2-D tables, RGBA16F 3-D tables: src/precompute.rs:1170,1191,1215)
//           each GLSL operation is one correctly rounded fp32 operation, no FMA contraction
//           (build with -ffp-contract=off) — "the shaders as written, in fp32".
//   mode 1  R = double, same quantisation points (fp64 arithmetic, fp16/fp32 storage)
//   mode 2  R = double, no quantisation anywhere ("ideal")
// The sampler restates VkSampler{LINEAR, CLAMP_TO_EDGE, normalised coords}
// (src/precompute.rs:85-98) with exact (non fixed-point) weights.
//
// Tables cross the C interface as double arrays [z][y][x][4] (x fastest: the read-back
// layout of examples/dump.rs:175-193); a value that was quantised is exactly representable.

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include <omp.h>

namespace {

// ---------------------------------------------------------------------------------------
// fp16 rounding (round-to-nearest-even, subnormals kept, overflow -> inf), from double so a
// double never rounds twice.
// ---------------------------------------------------------------------------------------
double round_to_half(double v) {
    if (!(v == v) || std::isinf(v) || v == 0.0) return v;
    double a = std::fabs(v);
    int e;
    std::frexp(a, &e);          // a = m * 2^e, m in [0.5,1)
    int exp2 = e - 1;           // a in [2^exp2, 2^(exp2+1))
    if (exp2 < -14) exp2 = -14; // subnormal range shares the spacing of the lowest binade
    double ulp = std::ldexp(1.0, exp2 - 10);
    double q = std::nearbyint(a / ulp) * ulp;   // default rounding mode: ties to even
    if (q >= 65520.0) q = std::numeric_limits<double>::infinity();
    return v < 0 ? -q : q;
}

// 320-byte uniform block, shaders/params.h:9-87 / src/precompute.rs:937-1033.
struct RawLayer { float width, exp_term, exp_scale, linear_term, constant_term, pad[3]; };
struct RawProfile { RawLayer layers[2]; };
struct RawParams {
    float solar_irradiance[3];   float sun_angular_radius;
    float rayleigh_scattering[3]; float bottom_radius;
    float mie_scattering[3];     float top_radius;
    float mie_extinction[3];     float mie_phase_function_g;
    float ground_albedo[3];      float mu_s_min;
    float absorption_extinction[3];
    int32_t transmittance_mu_size, transmittance_r_size;
    int32_t scattering_r_size, scattering_mu_size, scattering_mu_s_size, scattering_nu_size;
    int32_t irradiance_mu_s_size, irradiance_r_size;
    int32_t pad;
    RawProfile rayleigh_density, mie_density, absorption_density;
};
static_assert(sizeof(RawParams) == 320, "std140 block is 320 bytes");
static_assert(offsetof(RawParams, transmittance_mu_size) == 92, "sizes start at 92");
static_assert(offsetof(RawParams, rayleigh_density) == 128, "profiles start at 128");

template <class R> struct V3 {
    R x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(R a, R b, R c) : x(a), y(b), z(c) {}
    explicit V3(R a) : x(a), y(a), z(a) {}
    V3 operator+(const V3& o) const { return V3(x + o.x, y + o.y, z + o.z); }
    V3 operator-(const V3& o) const { return V3(x - o.x, y - o.y, z - o.z); }
    V3 operator*(const V3& o) const { return V3(x * o.x, y * o.y, z * o.z); }
    V3 operator/(const V3& o) const { return V3(x / o.x, y / o.y, z / o.z); }
    V3 operator*(R s) const { return V3(x * s, y * s, z * s); }
    V3 operator/(R s) const { return V3(x / s, y / s, z / s); }
    V3& operator+=(const V3& o) { x += o.x; y += o.y; z += o.z; return *this; }
};
template <class R> R dot(const V3<R>& a, const V3<R>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class R> V3<R> vmin1(const V3<R>& a) {
    return V3<R>(std::min(a.x, R(1)), std::min(a.y, R(1)), std::min(a.z, R(1)));
}

template <class R> struct V4 {
    R x, y, z, w;
    V4 operator+(const V4& o) const { return V4{x + o.x, y + o.y, z + o.z, w + o.w}; }
    V4 operator*(R s) const { return V4{x * s, y * s, z * s, w * s}; }
    V3<R> rgb() const { return V3<R>(x, y, z); }
};

template <class R> R clampv(R v, R lo, R hi) { return std::min(std::max(v, lo), hi); }
template <class R> R smoothstep(R e0, R e1, R x) {
    R t = clampv<R>((x - e0) / (e1 - e0), R(0), R(1));
    return t * t * (R(3) - R(2) * t);
}

// A table as the shaders see it: texels already hold the values the image format can hold.
template <class R> struct Table {
    int w = 0, h = 0, d = 1;
    std::vector<R> px;   // [z][y][x][4]
    void resize(int W, int H, int D) { w = W; h = H; d = D; px.assign(size_t(W) * H * D * 4, R(0)); }
    V4<R> texel(int x, int y, int z) const {
        const R* p = &px[((size_t(z) * h + y) * w + x) * 4];
        return V4<R>{p[0], p[1], p[2], p[3]};
    }
    R* at(int x, int y, int z) { return &px[((size_t(z) * h + y) * w + x) * 4]; }
    // VK_FILTER_LINEAR + CLAMP_TO_EDGE on normalised coordinates (src/precompute.rs:85-98):
    // unnormalised u*N - 0.5, floor / fract, both taps clamped into [0, N-1].
    static void axis(R u, int n, int& i0, int& i1, R& f) {
        R t = u * R(n) - R(0.5);
        R fl = std::floor(t);
        f = t - fl;
        // clamp in floating point first so wild coordinates cannot overflow the int cast
        R lo = std::min(std::max(fl, R(-1)), R(n));
        int i = int(lo);
        i0 = std::min(std::max(i, 0), n - 1);
        i1 = std::min(std::max(i + 1, 0), n - 1);
    }
    V4<R> sample2(R u, R v) const {
        int x0, x1, y0, y1; R fx, fy;
        axis(u, w, x0, x1, fx); axis(v, h, y0, y1, fy);
        V4<R> a = texel(x0, y0, 0) * (R(1) - fx) + texel(x1, y0, 0) * fx;
        V4<R> b = texel(x0, y1, 0) * (R(1) - fx) + texel(x1, y1, 0) * fx;
        return a * (R(1) - fy) + b * fy;
    }
    V4<R> sample3(R u, R v, R s) const {
        int x0, x1, y0, y1, z0, z1; R fx, fy, fz;
        axis(u, w, x0, x1, fx); axis(v, h, y0, y1, fy); axis(s, d, z0, z1, fz);
        V4<R> a = texel(x0, y0, z0) * (R(1) - fx) + texel(x1, y0, z0) * fx;
        V4<R> b = texel(x0, y1, z0) * (R(1) - fx) + texel(x1, y1, z0) * fx;
        V4<R> c = texel(x0, y0, z1) * (R(1) - fx) + texel(x1, y0, z1) * fx;
        V4<R> e = texel(x0, y1, z1) * (R(1) - fx) + texel(x1, y1, z1) * fx;
        V4<R> ab = a * (R(1) - fy) + b * fy;
        V4<R> ce = c * (R(1) - fy) + e * fy;
        return ab * (R(1) - fz) + ce * fz;
    }
};

template <class R> struct Layer { R width, exp_term, exp_scale, linear_term, constant_term; };
template <class R> struct Profile { Layer<R> layers[2]; };

template <class R> struct Oracle {
    // --- shaders/params.h:26-87 in R -----------------------------------------------------
    V3<R> solar_irradiance, rayleigh_scattering, mie_scattering, mie_extinction, ground_albedo,
        absorption_extinction;
    R sun_angular_radius, bottom_radius, top_radius, mie_phase_function_g, mu_s_min;
    int T_mu, T_r, S_r, S_mu, S_mu_s, S_nu, E_mu_s, E_r;
    Profile<R> rayleigh_density, mie_density, absorption_density;
    bool quantise;   // false only in mode 2
    R PI;

    static Profile<R> conv(const RawProfile& p) {
        Profile<R> o;
        for (int i = 0; i < 2; ++i) {
            const RawLayer& l = p.layers[i];
            o.layers[i] = Layer<R>{R(l.width), R(l.exp_term), R(l.exp_scale), R(l.linear_term),
                                   R(l.constant_term)};
        }
        return o;
    }
    Oracle(const RawParams& p, bool q) : quantise(q) {
        auto v3 = [](const float* f) { return V3<R>(R(f[0]), R(f[1]), R(f[2])); };
        solar_irradiance = v3(p.solar_irradiance);
        rayleigh_scattering = v3(p.rayleigh_scattering);
        mie_scattering = v3(p.mie_scattering);
        mie_extinction = v3(p.mie_extinction);
        ground_albedo = v3(p.ground_albedo);
        absorption_extinction = v3(p.absorption_extinction);
        sun_angular_radius = R(p.sun_angular_radius);
        bottom_radius = R(p.bottom_radius);
        top_radius = R(p.top_radius);
        mie_phase_function_g = R(p.mie_phase_function_g);
        mu_s_min = R(p.mu_s_min);
        T_mu = p.transmittance_mu_size; T_r = p.transmittance_r_size;
        S_r = p.scattering_r_size; S_mu = p.scattering_mu_size;
        S_mu_s = p.scattering_mu_s_size; S_nu = p.scattering_nu_size;
        E_mu_s = p.irradiance_mu_s_size; E_r = p.irradiance_r_size;
        rayleigh_density = conv(p.rayleigh_density);
        mie_density = conv(p.mie_density);
        absorption_density = conv(p.absorption_density);
        PI = R(3.14159265358979323846);   // util.h:4 (a float literal in GLSL)
    }

    // image-format rounding at the imageStore points
    R store32(R v) const { return quantise ? R(float(v)) : v; }          // RGBA32F tables
    R store16(R v) const { return quantise ? R(round_to_half(double(v))) : v; }  // RGBA16F tables

    // --- shaders/util.h ---------------------------------------------------------------------
    static R ClampCosine(R mu) { return clampv<R>(mu, R(-1), R(1)); }                // :6-8
    static R ClampDistance(R d) { return std::max(d, R(0)); }                        // :10-12
    static R SafeSqrt(R a) { return std::sqrt(std::max(a, R(0))); }                  // :14-16
    static R CoordFromUnit(R x, int n) {                                             // :18-20
        return R(0.5) / R(n) + x * (R(1) - R(1) / R(n));
    }
    static R UnitFromCoord(R u, int n) {                                             // :22-24
        return (u - R(0.5) / R(n)) / (R(1) - R(1) / R(n));
    }
    R RayleighPhase(R nu) const {                                                    // :26-29
        R k = R(3) / (R(16) * PI);
        return k * (R(1) + nu * nu);
    }
    R MiePhase(R g, R nu) const {                                                    // :31-34
        R k = R(3) / (R(8) * PI) * (R(1) - g * g) / (R(2) + g * g);
        return k * (R(1) + nu * nu) / std::pow(R(1) + g * g - R(2) * g * nu, R(1.5));
    }
    static R FragCoordFromTexel(unsigned x, unsigned n) {                            // :36-38
        return R(n) * CoordFromUnit(R(x) / R(n - 1), int(n));
    }

    // --- shaders/params.h:89-133 -------------------------------------------------------------
    static R LayerDensity(const Layer<R>& l, R h) {                                  // :89-93
        R d = l.exp_term * std::exp(l.exp_scale * h) + l.linear_term * h + l.constant_term;
        return clampv<R>(d, R(0), R(1));
    }
    static R ProfileDensity(const Profile<R>& p, R h) {                              // :95-99
        return h < p.layers[0].width ? LayerDensity(p.layers[0], h) : LayerDensity(p.layers[1], h);
    }
    R ClampRadius(R r) const { return clampv<R>(r, bottom_radius, top_radius); }     // :101-103
    R DistanceToTop(R r, R mu) const {                                               // :105-110
        R disc = r * r * (mu * mu - R(1)) + top_radius * top_radius;
        return ClampDistance(-r * mu + SafeSqrt(disc));
    }
    R DistanceToBottom(R r, R mu) const {                                            // :112-117
        R disc = r * r * (mu * mu - R(1)) + bottom_radius * bottom_radius;
        return ClampDistance(-r * mu - SafeSqrt(disc));
    }
    bool RayIntersectsGround(R r, R mu) const {                                      // :119-124
        return mu < R(0) && r * r * (mu * mu - R(1)) + bottom_radius * bottom_radius >= R(0);
    }
    R DistanceToNearest(R r, R mu, bool hits) const {                                // :126-133
        return hits ? DistanceToBottom(r, mu) : DistanceToTop(r, mu);
    }

    // --- shaders/transmittance.h -------------------------------------------------------------
    void TransmittanceUv(R r, R mu, R& u, R& v) const {                              // :7-24
        R H = std::sqrt(top_radius * top_radius - bottom_radius * bottom_radius);
        R rho = SafeSqrt(r * r - bottom_radius * bottom_radius);
        R d = DistanceToTop(r, mu);
        R d_min = top_radius - r;
        R d_max = rho + H;
        R x_mu = (d - d_min) / (d_max - d_min);
        R x_r = rho / H;
        u = CoordFromUnit(x_mu, T_mu);
        v = CoordFromUnit(x_r, T_r);
    }
    V3<R> TransmittanceToTop(const Table<R>& T, R r, R mu) const {                   // :26-33
        R u, v;
        TransmittanceUv(r, mu, u, v);
        return T.sample2(u, v).rgb();
    }
    V3<R> Transmittance(const Table<R>& T, R r, R mu, R d, bool hits) const {        // :35-61
        R r_d = ClampRadius(std::sqrt(d * d + R(2) * r * mu * d + r * r));
        R mu_d = ClampCosine((r * mu + d) / r_d);
        V3<R> q = hits ? TransmittanceToTop(T, r_d, -mu_d) / TransmittanceToTop(T, r, -mu)
                       : TransmittanceToTop(T, r, mu) / TransmittanceToTop(T, r_d, mu_d);
        return vmin1(q);
    }
    V3<R> TransmittanceToSun(const Table<R>& T, R r, R mu_s) const {                 // :63-74
        R sin_h = bottom_radius / r;
        R cos_h = -std::sqrt(std::max(R(1) - sin_h * sin_h, R(0)));
        return TransmittanceToTop(T, r, mu_s) *
               smoothstep<R>(-sin_h * sun_angular_radius, sin_h * sun_angular_radius, mu_s - cos_h);
    }

    // --- shaders/transmittance.comp ----------------------------------------------------------
    R OpticalLength(const Profile<R>& prof, R r, R mu) const {                       // :8-32
        const int N = 500;
        R dx = DistanceToTop(r, mu) / R(N);
        R acc = R(0);
        for (int i = 0; i <= N; ++i) {
            R d_i = R(i) * dx;
            R r_i = std::sqrt(d_i * d_i + R(2) * r * mu * d_i + r * r);
            R y_i = ProfileDensity(prof, r_i - bottom_radius);
            R w_i = (i == 0 || i == N) ? R(0.5) : R(1);
            acc += y_i * w_i * dx;
        }
        return acc;
    }
    V3<R> ComputeTransmittanceToTop(R r, R mu) const {                               // :34-46
        V3<R> tau = rayleigh_scattering * OpticalLength(rayleigh_density, r, mu) +
                    mie_extinction * OpticalLength(mie_density, r, mu) +
                    absorption_extinction * OpticalLength(absorption_density, r, mu);
        return V3<R>(std::exp(-tau.x), std::exp(-tau.y), std::exp(-tau.z));
    }
    void RMuFromUnitRanges(R x_mu, R x_r, R& r, R& mu) const {                       // :48-64
        R H = std::sqrt(top_radius * top_radius - bottom_radius * bottom_radius);
        R rho = H * x_r;
        r = std::sqrt(rho * rho + bottom_radius * bottom_radius);
        R d_min = top_radius - r;
        R d_max = rho + H;
        R d = d_min + x_mu * (d_max - d_min);
        mu = d == R(0) ? R(1) : (H * H - rho * rho - d * d) / (R(2) * r * d);
        mu = ClampCosine(mu);
    }
    void transmittance_pass(Table<R>& T) const {                                     // :71-80
        T.resize(T_mu, T_r, 1);
#pragma omp parallel for schedule(dynamic, 1)
        for (int y = 0; y < T_r; ++y)
            for (int x = 0; x < T_mu; ++x) {
                R r, mu;
                RMuFromUnitRanges(R(unsigned(x)) / R(T_mu - 1), R(unsigned(y)) / R(T_r - 1), r, mu);
                V3<R> t = ComputeTransmittanceToTop(r, mu);
                R* o = T.at(x, y, 0);
                o[0] = store32(t.x); o[1] = store32(t.y); o[2] = store32(t.z); o[3] = R(1);
            }
    }

    // --- shaders/irradiance.h ----------------------------------------------------------------
    void RMuSFromIrradianceUnit(R x_mu_s, R x_r, R& r, R& mu_s) const {              // :7-18
        r = bottom_radius + x_r * (top_radius - bottom_radius);
        mu_s = ClampCosine(R(2) * x_mu_s - R(1));
    }
    V3<R> Irradiance(const Table<R>& E, R r, R mu_s) const {                         // :20-38
        R x_r = (r - bottom_radius) / (top_radius - bottom_radius);
        R x_mu_s = mu_s * R(0.5) + R(0.5);
        return E.sample2(CoordFromUnit(x_mu_s, E_mu_s), CoordFromUnit(x_r, E_r)).rgb();
    }

    // --- shaders/direct_irradiance.comp ------------------------------------------------------
    void direct_irradiance_pass(const Table<R>& T, Table<R>& dE) const {             // :10-46
        dE.resize(E_mu_s, E_r, 1);
        for (int y = 0; y < E_r; ++y)
            for (int x = 0; x < E_mu_s; ++x) {
                R r, mu_s;
                RMuSFromIrradianceUnit(R(unsigned(x)) / R(E_mu_s - 1), R(unsigned(y)) / R(E_r - 1), r, mu_s);
                R a = sun_angular_radius;
                R avg = mu_s < -a ? R(0)
                                  : (mu_s > a ? mu_s : (mu_s + a) * (mu_s + a) / (R(4) * a));
                V3<R> e = solar_irradiance * TransmittanceToTop(T, r, mu_s) * avg;
                R* o = dE.at(x, y, 0);
                o[0] = store32(e.x); o[1] = store32(e.y); o[2] = store32(e.z); o[3] = R(0);
            }
    }

    // --- shaders/scattering.h ----------------------------------------------------------------
    void ScatteringUvwz(R r, R mu, R mu_s, R nu, bool hits, R uvwz[4]) const {       // :7-60
        R H = std::sqrt(top_radius * top_radius - bottom_radius * bottom_radius);
        R rho = SafeSqrt(r * r - bottom_radius * bottom_radius);
        R u_r = CoordFromUnit(rho / H, S_r);
        R r_mu = r * mu;
        R disc = r_mu * r_mu - r * r + bottom_radius * bottom_radius;
        R u_mu;
        if (hits) {
            R d = -r_mu - SafeSqrt(disc);
            R d_min = r - bottom_radius;
            R d_max = rho;
            u_mu = R(0.5) - R(0.5) * CoordFromUnit(d_max == d_min ? R(0) : (d - d_min) / (d_max - d_min),
                                                   S_mu / 2);
        } else {
            R d = -r_mu + SafeSqrt(disc + H * H);
            R d_min = top_radius - r;
            R d_max = rho + H;
            u_mu = R(0.5) + R(0.5) * CoordFromUnit((d - d_min) / (d_max - d_min), S_mu / 2);
        }
        R d = DistanceToTop(bottom_radius, mu_s);
        R d_min = top_radius - bottom_radius;
        R d_max = H;
        R a = (d - d_min) / (d_max - d_min);
        R A = R(-2) * mu_s_min * bottom_radius / (d_max - d_min);
        R u_mu_s = CoordFromUnit(std::max(R(1) - a / A, R(0)) / (R(1) + a), S_mu_s);
        R u_nu = (nu + R(1)) / R(2);
        uvwz[0] = u_nu; uvwz[1] = u_mu_s; uvwz[2] = u_mu; uvwz[3] = u_r;
    }
    void RMuMuSNuFromUvwz(const R uvwz[4], R& r, R& mu, R& mu_s, R& nu, bool& hits) const {  // :62-114
        R H = std::sqrt(top_radius * top_radius - bottom_radius * bottom_radius);
        R rho = H * UnitFromCoord(uvwz[3], S_r);
        r = std::sqrt(rho * rho + bottom_radius * bottom_radius);
        if (uvwz[2] < R(0.5)) {
            R d_min = r - bottom_radius;
            R d_max = rho;
            R d = d_min + (d_max - d_min) * UnitFromCoord(R(1) - R(2) * uvwz[2], S_mu / 2);
            mu = d == R(0) ? R(-1) : ClampCosine(-(rho * rho + d * d) / (R(2) * r * d));
            hits = true;
        } else {
            R d_min = top_radius - r;
            R d_max = rho + H;
            R d = d_min + (d_max - d_min) * UnitFromCoord(R(2) * uvwz[2] - R(1), S_mu / 2);
            mu = d == R(0) ? R(1) : ClampCosine((H * H - rho * rho - d * d) / (R(2) * r * d));
            hits = false;
        }
        R x_mu_s = UnitFromCoord(uvwz[1], S_mu_s);
        R d_min = top_radius - bottom_radius;
        R d_max = H;
        R A = R(-2) * mu_s_min * bottom_radius / (d_max - d_min);
        R a = (A - x_mu_s * A) / (R(1) + x_mu_s * A);
        R d = d_min + std::min(a, A) * (d_max - d_min);
        mu_s = d == R(0) ? R(1) : ClampCosine((H * H - d * d) / (R(2) * bottom_radius * d));
        nu = ClampCosine(uvwz[0] * R(2) - R(1));
    }
    // texel -> (r, mu, mu_s, nu): GetScatteringFragCoord :181-190 + ...FromScatteringTextureFragCoord :116-137
    void TexelToRMuMuSNu(unsigned x, unsigned y, unsigned z, R& r, R& mu, R& mu_s, R& nu, bool& hits) const {
        R fx = FragCoordFromTexel(x, unsigned(S_nu * S_mu_s));
        R fy = FragCoordFromTexel(y, unsigned(S_mu));
        R fz = FragCoordFromTexel(z, unsigned(S_r));
        R f_nu = std::floor(fx / R(S_mu_s));
        R f_mu_s = fx - R(S_mu_s) * std::floor(fx / R(S_mu_s));   // GLSL mod()
        R uvwz[4] = {f_nu / R(S_nu - 1), f_mu_s / R(S_mu_s), fy / R(S_mu), fz / R(S_r)};
        RMuMuSNuFromUvwz(uvwz, r, mu, mu_s, nu, hits);
        R s = std::sqrt((R(1) - mu * mu) * (R(1) - mu_s * mu_s));
        nu = clampv<R>(nu, mu * mu_s - s, mu * mu_s + s);
    }
    V4<R> Scattering4(const Table<R>& S, R r, R mu, R mu_s, R nu, bool hits) const {  // :139-155 (all 4 channels)
        R uvwz[4];
        ScatteringUvwz(r, mu, mu_s, nu, hits, uvwz);
        R tcx = uvwz[0] * R(S_nu - 1);
        R tx = std::floor(tcx);
        R l = tcx - tx;
        V4<R> a = S.sample3((tx + uvwz[1]) / R(S_nu), uvwz[2], uvwz[3]);
        V4<R> b = S.sample3((tx + R(1) + uvwz[1]) / R(S_nu), uvwz[2], uvwz[3]);
        return a * (R(1) - l) + b * l;
    }
    V3<R> ScatteringOrder(const Table<R>& dR, const Table<R>& dM, const Table<R>& dMS, R r, R mu, R mu_s,
                          R nu, bool hits, int order) const {                        // :157-179
        if (order == 1) {
            V3<R> ray = Scattering4(dR, r, mu, mu_s, nu, hits).rgb();
            V3<R> mie = Scattering4(dM, r, mu, mu_s, nu, hits).rgb();
            return ray * RayleighPhase(nu) + mie * MiePhase(mie_phase_function_g, nu);
        }
        return Scattering4(dMS, r, mu, mu_s, nu, hits).rgb();
    }

    // --- shaders/single_scattering.comp ------------------------------------------------------
    void SingleScatteringAt(const Table<R>& T, R r, R mu, R mu_s, R nu, bool hits, V3<R>& ray, V3<R>& mie) const {
        const int N = 50;                                                            // :30-65
        R dx = DistanceToNearest(r, mu, hits) / R(N);
        V3<R> rs, ms;
        for (int i = 0; i <= N; ++i) {
            R d = R(i) * dx;
            // integrand :10-28
            R r_d = ClampRadius(std::sqrt(d * d + R(2) * r * mu * d + r * r));
            R mu_s_d = ClampCosine((r * mu_s + d * nu) / r_d);
            V3<R> t = Transmittance(T, r, mu, d, hits) * TransmittanceToSun(T, r_d, mu_s_d);
            V3<R> ri = t * ProfileDensity(rayleigh_density, r_d - bottom_radius);
            V3<R> mi = t * ProfileDensity(mie_density, r_d - bottom_radius);
            R w = (i == 0 || i == N) ? R(0.5) : R(1);
            rs += ri * w;
            ms += mi * w;
        }
        ray = rs * dx * solar_irradiance * rayleigh_scattering;
        mie = ms * dx * solar_irradiance * mie_scattering;
    }
    // :89-101; idx == nullptr -> every texel, else the n linear texel indices (outputs are [n][4] rows)
    void single_scattering_texels(const Table<R>& T, const int64_t* idx, int64_t n, double* dR, double* dM, double* S) const {
        int W = S_nu * S_mu_s;
        int64_t total = idx ? n : int64_t(W) * S_mu * S_r;
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t k = 0; k < total; ++k) {
            int64_t lin = idx ? idx[k] : k;
            unsigned x = unsigned(lin % W), y = unsigned((lin / W) % S_mu), z = unsigned(lin / (int64_t(W) * S_mu));
            R r, mu, mu_s, nu; bool hits;
            TexelToRMuMuSNu(x, y, z, r, mu, mu_s, nu, hits);
            V3<R> ray, mie;
            SingleScatteringAt(T, r, mu, mu_s, nu, hits, ray, mie);
            double* a = dR + k * 4; double* b = dM + k * 4; double* c = S + k * 4;
            a[0] = double(store16(ray.x)); a[1] = double(store16(ray.y)); a[2] = double(store16(ray.z)); a[3] = 0.0;
            b[0] = double(store16(mie.x)); b[1] = double(store16(mie.y)); b[2] = double(store16(mie.z)); b[3] = 0.0;
            c[0] = a[0]; c[1] = a[1]; c[2] = a[2]; c[3] = b[0];
        }
    }

    // --- shaders/scattering_density.comp -----------------------------------------------------
    V3<R> ScatteringDensityAt(const Table<R>& T, const Table<R>& dR, const Table<R>& dM, const Table<R>& dMS,
                              const Table<R>& dE, R r, R mu, R mu_s, R nu, int order) const {   // :10-107
        V3<R> omega(std::sqrt(R(1) - mu * mu), R(0), mu);
        R sx = omega.x == R(0) ? R(0) : (nu - mu * mu_s) / omega.x;
        R sy = std::sqrt(std::max(R(1) - sx * sx - mu_s * mu_s, R(0)));
        V3<R> omega_s(sx, sy, mu_s);
        const int N = 16;
        const R dphi = PI / R(N), dtheta = PI / R(N);
        V3<R> acc;
        // r-only factors of :94-97 (GLSL recomputes them per sample; same value every time)
        R ray_rho = ProfileDensity(rayleigh_density, r - bottom_radius);
        R mie_rho = ProfileDensity(mie_density, r - bottom_radius);
        for (int l = 0; l < N; ++l) {
            R theta = (R(l) + R(0.5)) * dtheta;
            R ct = std::cos(theta), st = std::sin(theta);
            bool hits = RayIntersectsGround(r, ct);
            R dist_ground = R(0);
            V3<R> t_ground, albedo;
            if (hits) {
                dist_ground = DistanceToBottom(r, ct);
                t_ground = Transmittance(T, r, ct, dist_ground, true);
                albedo = ground_albedo;
            }
            for (int m = 0; m < 2 * N; ++m) {
                R phi = (R(m) + R(0.5)) * dphi;
                V3<R> wi(std::cos(phi) * st, std::sin(phi) * st, ct);
                R dw = dtheta * dphi * std::sin(theta);
                R nu1 = dot(omega_s, wi);
                V3<R> L = ScatteringOrder(dR, dM, dMS, r, wi.z, mu_s, nu1, hits, order - 1);
                V3<R> gn = V3<R>(R(0), R(0), r) + wi * dist_ground;
                gn = gn / std::sqrt(dot(gn, gn));                                    // normalize()
                V3<R> gE = Irradiance(dE, bottom_radius, dot(gn, omega_s));
                L += t_ground * albedo * (R(1) / PI) * gE;
                R nu2 = dot(omega, wi);
                acc += L * (rayleigh_scattering * ray_rho * RayleighPhase(nu2) +
                            mie_scattering * mie_rho * MiePhase(mie_phase_function_g, nu2)) * dw;
            }
        }
        return acc;
    }

    // --- shaders/indirect_irradiance.comp ----------------------------------------------------
    V3<R> IndirectIrradianceAt(const Table<R>& dR, const Table<R>& dM, const Table<R>& dMS, R r, R mu_s,
                               int order) const {                                    // :10-44
        const int N = 32;
        const R dphi = PI / R(N), dtheta = PI / R(N);
        V3<R> acc;
        V3<R> omega_s(std::sqrt(R(1) - mu_s * mu_s), R(0), mu_s);
        for (int j = 0; j < N / 2; ++j) {
            R theta = (R(j) + R(0.5)) * dtheta;
            for (int i = 0; i < 2 * N; ++i) {
                R phi = (R(i) + R(0.5)) * dphi;
                V3<R> w(std::cos(phi) * std::sin(theta), std::sin(phi) * std::sin(theta), std::cos(theta));
                R dw = dtheta * dphi * std::sin(theta);
                R nu = dot(w, omega_s);
                acc += ScatteringOrder(dR, dM, dMS, r, w.z, mu_s, nu, false, order) * w.z * dw;
            }
        }
        return acc;
    }

    // --- shaders/multiple_scattering.comp ----------------------------------------------------
    V3<R> MultipleScatteringAt(const Table<R>& T, const Table<R>& dens, R r, R mu, R mu_s, R nu, bool hits) const {
        const int N = 50;                                                            // :9-54
        R dx = DistanceToNearest(r, mu, hits) / R(N);
        V3<R> acc;
        for (int i = 0; i <= N; ++i) {
            R d = R(i) * dx;
            R r_i = ClampRadius(std::sqrt(d * d + R(2) * r * mu * d + r * r));
            R mu_i = ClampCosine((r * mu + d) / r_i);
            R mu_s_i = ClampCosine((r * mu_s + d * nu) / r_i);
            V3<R> v = Scattering4(dens, r_i, mu_i, mu_s_i, nu, hits).rgb() * Transmittance(T, r, mu, d, hits) * dx;
            R w = (i == 0 || i == N) ? R(0.5) : R(1);
            acc += v * w;
        }
        return acc;
    }

    // --- shaders/render_sky.h ----------------------------------------------------------------
    V3<R> ExtrapolatedSingleMie(const V4<R>& s) const {                              // :9-19
        if (s.x <= R(0)) return V3<R>();
        return s.rgb() * s.w / s.x * (rayleigh_scattering.x / mie_scattering.x) *
               (mie_scattering / rayleigh_scattering);
    }
    V3<R> CombinedScattering(const Table<R>& S, R r, R mu, R mu_s, R nu, bool hits, V3<R>& single_mie) const {
        V4<R> c = Scattering4(S, r, mu, mu_s, nu, hits);   // :21-43 (mix(a,b,t) == a(1-t)+bt)
        single_mie = ExtrapolatedSingleMie(c);
        return c.rgb();
    }
    V3<R> SkyRadianceToPoint(const Table<R>& T, const Table<R>& S, V3<R> camera, V3<R> view, V3<R> point,
                             V3<R> sun, V3<R>& transmittance) const {                // :111-191
        R r = std::sqrt(dot(camera, camera));
        R rmu = dot(camera, view);
        R to_top = -rmu - std::sqrt(rmu * rmu - r * r + top_radius * top_radius);   // NaN if no hit
        if (to_top > R(0)) {
            camera = camera + view * to_top;
            r = top_radius;
            rmu += to_top;
        } else if (r > top_radius) {
            transmittance = V3<R>(R(1));
            return V3<R>();
        }
        R mu = rmu / r;
        R mu_s = dot(camera, sun) / r;
        R nu = dot(view, sun);
        V3<R> pc = point - camera;
        R d = std::sqrt(dot(pc, pc));
        bool hits = RayIntersectsGround(r, mu);
        transmittance = Transmittance(T, r, mu, d, hits);
        V3<R> single_mie;
        V3<R> scat = CombinedScattering(S, r, mu, mu_s, nu, hits, single_mie);
        if (!std::isinf(d)) {
            R r_p = ClampRadius(std::sqrt(d * d + R(2) * r * mu * d + r * r));
            R mu_p = (r * mu + d) / r_p;
            R mu_s_p = (r * mu_s + d * nu) / r_p;
            V3<R> single_mie_p;
            V3<R> scat_p = CombinedScattering(S, r_p, mu_p, mu_s_p, nu, hits, single_mie_p);
            scat = scat - transmittance * scat_p;
            single_mie = single_mie - transmittance * single_mie_p;
            single_mie = ExtrapolatedSingleMie(V4<R>{scat.x, scat.y, scat.z, single_mie.x});
            single_mie = single_mie * smoothstep<R>(R(0), R(0.01), mu_s);
        }
        return scat * RayleighPhase(nu) + single_mie * MiePhase(mie_phase_function_g, nu);
    }
    // GetSkyRadiance, render_sky.h:45-109 (no-depth variant; library function)
    V3<R> SkyRadiance(const Table<R>& T, const Table<R>& S, V3<R> camera, V3<R> view, V3<R> sun,
                      V3<R>& transmittance) const {
        R r = std::sqrt(dot(camera, camera));
        R rmu = dot(camera, view);
        R to_top = -rmu - std::sqrt(rmu * rmu - r * r + top_radius * top_radius);
        if (to_top > R(0)) {
            camera = camera + view * to_top;
            r = top_radius;
            rmu += to_top;
        } else if (r > top_radius) {
            transmittance = V3<R>(R(1));
            return V3<R>();
        }
        R mu = rmu / r;
        R mu_s = dot(camera, sun) / r;
        R nu = dot(view, sun);
        bool hits = RayIntersectsGround(r, mu);
        transmittance = hits ? V3<R>() : TransmittanceToTop(T, r, mu);
        V3<R> single_mie;
        V3<R> scat = CombinedScattering(S, r, mu, mu_s, nu, hits, single_mie);
        return scat * RayleighPhase(nu) + single_mie * MiePhase(mie_phase_function_g, nu);
    }
    // GetSunAndSkyIrradiance, render_lighting.h:10-28
    V3<R> SunAndSkyIrradiance(const Table<R>& T, const Table<R>& E, V3<R> point, V3<R> normal, V3<R> sun,
                              V3<R>& sky) const {
        R r = std::sqrt(dot(point, point));
        R mu_s = dot(point, sun) / r;
        sky = Irradiance(E, r, mu_s) * (R(1) + dot(normal, point) / r) * R(0.5);
        return solar_irradiance * TransmittanceToSun(T, r, mu_s) * std::max(dot(normal, sun), R(0));
    }
};

// ------------------------------------------------------------------------------------------
// C interface plumbing: double arrays <-> Table<R>
// ------------------------------------------------------------------------------------------
template <class R> void load(Table<R>& t, const double* src, int w, int h, int d) {
    t.resize(w, h, d);
    if (src) for (size_t i = 0; i < t.px.size(); ++i) t.px[i] = R(src[i]);
}
template <class R> void save(const Table<R>& t, double* dst) {
    if (dst) for (size_t i = 0; i < t.px.size(); ++i) dst[i] = double(t.px[i]);
}

template <class R> struct Run {
    Oracle<R> o;
    int W;
    Run(const RawParams& p, bool q) : o(p, q), W(p.scattering_nu_size * p.scattering_mu_s_size) {}

    void transmittance(double* T) { Table<R> t; o.transmittance_pass(t); save(t, T); }
    void direct_irradiance(const double* T, double* dE) {
        Table<R> t, e; load(t, T, o.T_mu, o.T_r, 1); o.direct_irradiance_pass(t, e); save(e, dE);
    }
    void single_scattering(const double* T, const int64_t* idx, int64_t n, double* dR, double* dM, double* S) {
        Table<R> t; load(t, T, o.T_mu, o.T_r, 1);
        o.single_scattering_texels(t, idx, n, dR, dM, S);
    }
    // texel subset: idx == nullptr -> every texel (n ignored), out is [n][4] or the whole table
    void scattering_density(const double* T, const double* dR, const double* dM, const double* dMS,
                            const double* dE, int order, const int64_t* idx, int64_t n, double* out) {
        Table<R> t, a, b, c, e;
        load(t, T, o.T_mu, o.T_r, 1); load(a, dR, W, o.S_mu, o.S_r); load(b, dM, W, o.S_mu, o.S_r);
        load(c, dMS, W, o.S_mu, o.S_r); load(e, dE, o.E_mu_s, o.E_r, 1);
        int64_t total = idx ? n : int64_t(W) * o.S_mu * o.S_r;
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t k = 0; k < total; ++k) {
            int64_t lin = idx ? idx[k] : k;
            unsigned x = unsigned(lin % W), y = unsigned((lin / W) % o.S_mu), z = unsigned(lin / (int64_t(W) * o.S_mu));
            R r, mu, mu_s, nu; bool hits;
            o.TexelToRMuMuSNu(x, y, z, r, mu, mu_s, nu, hits);
            V3<R> v = o.ScatteringDensityAt(t, a, b, c, e, r, mu, mu_s, nu, order);
            double* p = out + k * 4;
            p[0] = double(o.store16(v.x)); p[1] = double(o.store16(v.y)); p[2] = double(o.store16(v.z)); p[3] = 0.0;
        }
    }
    // delta_irradiance is overwritten, irradiance accumulated (indirect_irradiance.comp:72-73)
    void indirect_irradiance(const double* dR, const double* dM, const double* dMS, int order, double* dE,
                             double* E) {
        Table<R> a, b, c;
        load(a, dR, W, o.S_mu, o.S_r); load(b, dM, W, o.S_mu, o.S_r); load(c, dMS, W, o.S_mu, o.S_r);
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
        for (int y = 0; y < o.E_r; ++y)
            for (int x = 0; x < o.E_mu_s; ++x) {
                R r, mu_s;
                o.RMuSFromIrradianceUnit(R(unsigned(x)) / R(o.E_mu_s - 1), R(unsigned(y)) / R(o.E_r - 1), r, mu_s);
                V3<R> v = o.IndirectIrradianceAt(a, b, c, r, mu_s, order);
                size_t k = (size_t(y) * o.E_mu_s + x) * 4;
                R res[3] = {v.x, v.y, v.z};
                for (int ch = 0; ch < 3; ++ch) {
                    dE[k + ch] = double(o.store32(res[ch]));
                    E[k + ch] = double(o.store32(res[ch] + R(E[k + ch])));
                }
                dE[k + 3] = 0.0;
                E[k + 3] = double(o.store32(R(0) + R(E[k + 3])));
            }
    }
    // delta_multiple_scattering written, scattering accumulated in its storage format
    // (multiple_scattering.comp:91-92)
    void multiple_scattering(const double* T, const double* dens, const int64_t* idx, int64_t n, double* dMS,
                             double* S) {
        Table<R> t, dn;
        load(t, T, o.T_mu, o.T_r, 1); load(dn, dens, W, o.S_mu, o.S_r);
        int64_t total = idx ? n : int64_t(W) * o.S_mu * o.S_r;
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t k = 0; k < total; ++k) {
            int64_t lin = idx ? idx[k] : k;
            unsigned x = unsigned(lin % W), y = unsigned((lin / W) % o.S_mu), z = unsigned(lin / (int64_t(W) * o.S_mu));
            R r, mu, mu_s, nu; bool hits;
            o.TexelToRMuMuSNu(x, y, z, r, mu, mu_s, nu, hits);
            V3<R> ms = o.MultipleScatteringAt(t, dn, r, mu, mu_s, nu, hits);
            R pr = o.RayleighPhase(nu);
            R res[3] = {ms.x, ms.y, ms.z};
            double* pd = dMS + k * 4;
            double* ps = S + k * 4;   // S is [n][4] when idx != nullptr (caller pre-fills with old values)
            for (int ch = 0; ch < 3; ++ch) {
                pd[ch] = double(o.store16(res[ch]));
                ps[ch] = double(o.store16(res[ch] / pr + R(ps[ch])));
            }
            pd[3] = 0.0;
            ps[3] = double(o.store16(R(0) + R(ps[3])));
        }
    }
    // render_sky.frag:24-35 with fullscreen.vert:5-8 (screen_coords = (pixel + 0.5) / size)
    void render(const double* T, const double* S, const float* draw, const float* depth, int w, int h,
                const int64_t* idx, int64_t n, double* color, double* transm) {
        Table<R> t, s; load(t, T, o.T_mu, o.T_r, 1); load(s, S, W, o.S_mu, o.S_r);
        R M[4][4];   // column-major: M[col][row]
        for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) M[c][r] = R(draw[c * 4 + r]);
        V3<R> cam{R(draw[16]), R(draw[17]), R(draw[18])};
        V3<R> sun{R(draw[20]), R(draw[21]), R(draw[22])};
        int64_t total = idx ? n : int64_t(w) * h;
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t k = 0; k < total; ++k) {
            int64_t lin = idx ? idx[k] : k;
            int px = int(lin % w), py = int(lin / w);
            R sx = (R(px) + R(0.5)) / R(w), sy = (R(py) + R(0.5)) / R(h);
            R nx = R(2) * sx - R(1), ny = R(2) * sy - R(1);
            auto xf = [&](R zc, R out[4]) {
                for (int r = 0; r < 4; ++r) out[r] = M[0][r] * nx + M[1][r] * ny + M[2][r] * zc + M[3][r] * R(1);
            };
            R v0[4], v1[4];
            xf(R(0), v0);
            V3<R> view(v0[0], v0[1], v0[2]);
            view = view / std::sqrt(dot(view, view));
            xf(R(depth[lin]), v1);
            V3<R> world = V3<R>(v1[0] / v1[3], v1[1] / v1[3], v1[2] / v1[3]) * R(1e-3);
            V3<R> tr;
            V3<R> c = o.SkyRadianceToPoint(t, s, cam, view, world, sun, tr);
            double* pc = color + k * 4; double* pt = transm + k * 4;
            pc[0] = double(c.x); pc[1] = double(c.y); pc[2] = double(c.z); pc[3] = 0.0;
            pt[0] = double(tr.x); pt[1] = double(tr.y); pt[2] = double(tr.z); pt[3] = 1.0;
        }
    }
    // library functions (render_sky.h:45-109, render_lighting.h:10-28), n independent queries
    void sky_radiance(const double* T, const double* S, const double* cam, const double* view, const double* sun,
                      int64_t n, double* radiance, double* transm) {
        Table<R> t, s; load(t, T, o.T_mu, o.T_r, 1); load(s, S, W, o.S_mu, o.S_r);
        for (int64_t k = 0; k < n; ++k) {
            auto v3 = [&](const double* p) { return V3<R>(R(float(p[k * 3])), R(float(p[k * 3 + 1])), R(float(p[k * 3 + 2]))); };
            V3<R> tr;
            V3<R> c = o.SkyRadiance(t, s, v3(cam), v3(view), v3(sun), tr);
            radiance[k * 3] = double(c.x); radiance[k * 3 + 1] = double(c.y); radiance[k * 3 + 2] = double(c.z);
            transm[k * 3] = double(tr.x); transm[k * 3 + 1] = double(tr.y); transm[k * 3 + 2] = double(tr.z);
        }
    }
    void sun_sky_irradiance(const double* T, const double* E, const double* point, const double* normal,
                            const double* sun, int64_t n, double* direct, double* sky) {
        Table<R> t, e; load(t, T, o.T_mu, o.T_r, 1); load(e, E, o.E_mu_s, o.E_r, 1);
        for (int64_t k = 0; k < n; ++k) {
            auto v3 = [&](const double* p) { return V3<R>(R(float(p[k * 3])), R(float(p[k * 3 + 1])), R(float(p[k * 3 + 2]))); };
            V3<R> sk;
            V3<R> c = o.SunAndSkyIrradiance(t, e, v3(point), v3(normal), v3(sun), sk);
            direct[k * 3] = double(c.x); direct[k * 3 + 1] = double(c.y); direct[k * 3 + 2] = double(c.z);
            sky[k * 3] = double(sk.x); sky[k * 3 + 1] = double(sk.y); sky[k * 3 + 2] = double(sk.z);
        }
    }
};

template <class F> int dispatch(const void* params, int mode, F&& f) {
    RawParams p;
    std::memcpy(&p, params, sizeof p);
    if (mode == 0) { Run<float> r(p, true); f(r); return 0; }
    if (mode == 1) { Run<double> r(p, true); f(r); return 0; }
    if (mode == 2) { Run<double> r(p, false); f(r); return 0; }
    return -1;
}

}  // namespace

extern "C" {

int fbo_params_size(void) { return int(sizeof(RawParams)); }
// worker threads of the texel loops (launchers such as torchrun export OMP_NUM_THREADS=1 to their children)
void fbo_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int fbo_max_threads(void) { return omp_get_max_threads(); }
double fbo_round_to_half(double v) { return round_to_half(v); }

int fbo_transmittance(const void* params, int mode, double* T) {
    return dispatch(params, mode, [&](auto& r) { r.transmittance(T); });
}
int fbo_direct_irradiance(const void* params, int mode, const double* T, double* dE) {
    return dispatch(params, mode, [&](auto& r) { r.direct_irradiance(T, dE); });
}
int fbo_single_scattering(const void* params, int mode, const double* T, const int64_t* idx, int64_t n, double* dR,
                          double* dM, double* S) {
    return dispatch(params, mode, [&](auto& r) { r.single_scattering(T, idx, n, dR, dM, S); });
}
int fbo_scattering_density(const void* params, int mode, int order, const double* T, const double* dR,
                           const double* dM, const double* dMS, const double* dE, const int64_t* idx, int64_t n,
                           double* out) {
    return dispatch(params, mode, [&](auto& r) { r.scattering_density(T, dR, dM, dMS, dE, order, idx, n, out); });
}
int fbo_indirect_irradiance(const void* params, int mode, int order, const double* dR, const double* dM,
                            const double* dMS, double* dE, double* E) {
    return dispatch(params, mode, [&](auto& r) { r.indirect_irradiance(dR, dM, dMS, order, dE, E); });
}
int fbo_multiple_scattering(const void* params, int mode, const double* T, const double* dens, const int64_t* idx,
                            int64_t n, double* dMS, double* S) {
    return dispatch(params, mode, [&](auto& r) { r.multiple_scattering(T, dens, idx, n, dMS, S); });
}
int fbo_render(const void* params, int mode, const double* T, const double* S, const float* draw92,
               const float* depth, int w, int h, const int64_t* idx, int64_t n, double* color, double* transm) {
    return dispatch(params, mode, [&](auto& r) { r.render(T, S, draw92, depth, w, h, idx, n, color, transm); });
}
int fbo_sky_radiance(const void* params, int mode, const double* T, const double* S, const double* cam,
                     const double* view, const double* sun, int64_t n, double* radiance, double* transm) {
    return dispatch(params, mode, [&](auto& r) { r.sky_radiance(T, S, cam, view, sun, n, radiance, transm); });
}
int fbo_sun_sky_irradiance(const void* params, int mode, const double* T, const double* E, const double* point,
                           const double* normal, const double* sun, int64_t n, double* direct, double* sky) {
    return dispatch(params, mode, [&](auto& r) { r.sun_sky_irradiance(T, E, point, normal, sun, n, direct, sky); });
}

// debugging aid for tests: texel -> (r, mu, mu_s, nu, hits) and the single-scattering integrand samples
int fbo_debug_single(const void* params, int mode, const double* T, int x, int y, int z, double* geom5,
                     double* per_sample /* [51][8]: d, r_d, mu_s_d, Tpath.rgb, Tsun.r, rho_ray */) {
    return dispatch(params, mode, [&](auto& run) {
        auto& o = run.o;
        using R = decltype(o.PI);
        Table<R> t; load(t, T, o.T_mu, o.T_r, 1);
        R r, mu, mu_s, nu; bool hits;
        o.TexelToRMuMuSNu(unsigned(x), unsigned(y), unsigned(z), r, mu, mu_s, nu, hits);
        geom5[0] = double(r); geom5[1] = double(mu); geom5[2] = double(mu_s); geom5[3] = double(nu); geom5[4] = hits;
        R dx = o.DistanceToNearest(r, mu, hits) / R(50);
        for (int i = 0; i <= 50; ++i) {
            R d = R(i) * dx;
            R r_d = o.ClampRadius(std::sqrt(d * d + R(2) * r * mu * d + r * r));
            R mu_s_d = o.ClampCosine((r * mu_s + d * nu) / r_d);
            auto tp = o.Transmittance(t, r, mu, d, hits);
            auto ts = o.TransmittanceToSun(t, r_d, mu_s_d);
            double* ps = per_sample + i * 8;
            ps[0] = double(d); ps[1] = double(r_d); ps[2] = double(mu_s_d); ps[3] = double(tp.x); ps[4] = double(tp.y);
            ps[5] = double(tp.z); ps[6] = double(ts.x); ps[7] = double(o.ProfileDensity(o.rayleigh_density, r_d - o.bottom_radius));
        }
    });
}

}  // extern "C"
