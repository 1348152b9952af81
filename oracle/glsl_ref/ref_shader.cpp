// ref_shader.cpp — one translation unit per shader of the reference (compile with -DSHADER_<NAME>): includes the
// reference's own (syntax-translated) source from the build-time scratch directory and runs its main() over a dispatch on the CPU.
// TEST INFRASTRUCTURE ONLY; see glsl_compat.h for what is the reference's and what is the "driver" defined here.
#include <cstdint>
#include <cstring>
#include <vector>

#include "glsl_compat.h"

namespace glsl {
namespace {   // internal linkage: every translation unit holds its own copy of the shared shader headers
#if defined(SHADER_TRANSMITTANCE)
#include "transmittance.comp"
#elif defined(SHADER_DIRECT_IRRADIANCE)
#include "direct_irradiance.comp"
#elif defined(SHADER_SINGLE_SCATTERING)
#include "single_scattering.comp"
#elif defined(SHADER_SCATTERING_DENSITY)
#include "scattering_density.comp"
#elif defined(SHADER_INDIRECT_IRRADIANCE)
#include "indirect_irradiance.comp"
#elif defined(SHADER_MULTIPLE_SCATTERING)
#include "multiple_scattering.comp"
#elif defined(SHADER_RENDER_SKY)
#include "render_sky.frag"
#else
#error "define one SHADER_<NAME>"
#endif

// The 320-byte std140 block (shaders/params.h:26-87 as laid out by ParamsRaw, src/precompute.rs:937-1033) -> the struct
static void load_params(const void* raw) {
    const float* f = static_cast<const float*>(raw);
    const int32_t* i = static_cast<const int32_t*>(raw);
    AtmosphereParameters& a = atmosphere;
    a.solar_irradiance = vec3(f[0], f[1], f[2]);      a.sun_angular_radius = f[3];
    a.rayleigh_scattering = vec3(f[4], f[5], f[6]);   a.bottom_radius = f[7];
    a.mie_scattering = vec3(f[8], f[9], f[10]);       a.top_radius = f[11];
    a.mie_extinction = vec3(f[12], f[13], f[14]);     a.mie_phase_function_g = f[15];
    a.ground_albedo = vec3(f[16], f[17], f[18]);      a.mu_s_min = f[19];
    a.absorption_extinction = vec3(f[20], f[21], f[22]);
    a.transmittance_texture_mu_size = i[23]; a.transmittance_texture_r_size = i[24];
    a.scattering_texture_r_size = i[25];     a.scattering_texture_mu_size = i[26];
    a.scattering_texture_mu_s_size = i[27];  a.scattering_texture_nu_size = i[28];
    a.irradiance_texture_mu_s_size = i[29];  a.irradiance_texture_r_size = i[30];
    DensityProfile* prof[3] = {&a.rayleigh_density, &a.mie_density, &a.absorption_density};
    for (int p = 0; p < 3; ++p)
        for (int l = 0; l < 2; ++l) {
            const float* q = f + 32 + (p * 2 + l) * 8;    // profiles at byte 128, 32 bytes per layer
            DensityProfileLayer& L = prof[p]->layers[l];
            L.width = q[0]; L.exp_term = q[1]; L.exp_scale = q[2]; L.linear_term = q[3]; L.constant_term = q[4];
        }
}
static Image img(const double* p, int w, int h, int d, bool half) {
    Image im; im.p = const_cast<double*>(p); im.w = w; im.h = h; im.d = d; im.half = half; return im;
}
static int SW() { return atmosphere.scattering_texture_nu_size * atmosphere.scattering_texture_mu_s_size; }
static int SH() { return atmosphere.scattering_texture_mu_size; }
static int SD() { return atmosphere.scattering_texture_r_size; }
static Image tex3(const double* p, bool half = true) { return img(p, SW(), SH(), SD(), half); }
static Image texT(const double* p) { return img(p, atmosphere.transmittance_texture_mu_size, atmosphere.transmittance_texture_r_size, 1, false); }
static Image texE(const double* p) { return img(p, atmosphere.irradiance_texture_mu_s_size, atmosphere.irradiance_texture_r_size, 1, false); }

// run main() over the 2-D extent (w, h)
static void dispatch2(int w, int h) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t k = 0; k < (int64_t)w * h; ++k) {
        gl_GlobalInvocationID = uvec3((uint)(k % w), (uint)(k / w), 0u);
        shader_main();
    }
}
// run main() over the whole 3-D table, or over the texels idx[0..n)
static void dispatch3(const int64_t* idx, int64_t n) {
    const int64_t W = SW(), H = SH(), total = W * H * SD();
    const int64_t count = idx ? n : total;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < count; ++k) {
        const int64_t t = idx ? idx[k] : k;
        gl_GlobalInvocationID = uvec3((uint)(t % W), (uint)((t / W) % H), (uint)(t / (W * H)));
        shader_main();
    }
}
// gather / scatter the rows idx[0..n) of a full table
static void gather(const std::vector<double>& full, const int64_t* idx, int64_t n, double* rows) {
    for (int64_t k = 0; k < n; ++k) std::memcpy(rows + 4 * k, full.data() + 4 * idx[k], 4 * sizeof(double));
}
}  // anonymous namespace
}  // namespace glsl

using namespace glsl;


#if defined(SHADER_TRANSMITTANCE)
extern "C" int fbr_transmittance(const void* params, double* T) {
    load_params(params);
    table = texT(T);
    dispatch2(table.w, table.h);
    return 0;
}
#elif defined(SHADER_DIRECT_IRRADIANCE)
extern "C" int fbr_direct_irradiance(const void* params, const double* T, double* dE) {
    load_params(params);
    transmittance_texture = texT(T);
    delta_irradiance = texE(dE);
    dispatch2(delta_irradiance.w, delta_irradiance.h);
    return 0;
}
#elif defined(SHADER_SINGLE_SCATTERING)
extern "C" int fbr_single_scattering(const void* params, const double* T, const int64_t* idx, int64_t n, double* dR, double* dM, double* S) {
    load_params(params);
    transmittance = texT(T);
    const size_t total = (size_t)SW() * SH() * SD() * 4;
    std::vector<double> fR, fM, fS;
    if (idx) { fR.assign(total, 0.0); fM.assign(total, 0.0); fS.assign(total, 0.0); }
    delta_rayleigh = tex3(idx ? fR.data() : dR); delta_mie = tex3(idx ? fM.data() : dM); scattering = tex3(idx ? fS.data() : S);
    dispatch3(idx, n);
    if (idx) { gather(fR, idx, n, dR); gather(fM, idx, n, dM); gather(fS, idx, n, S); }
    return 0;
}
#elif defined(SHADER_SCATTERING_DENSITY)
extern "C" int fbr_scattering_density(const void* params, int order, const double* T, const double* dR, const double* dM, const double* dMS,
                                      const double* dE, const int64_t* idx, int64_t n, double* out) {
    load_params(params);
    scattering_order = order;                                   // push constant, src/precompute.rs:1892-1898
    transmittance_texture = texT(T);
    single_rayleigh_scattering_texture = tex3(dR); single_mie_scattering_texture = tex3(dM);
    multiple_scattering_texture = tex3(dMS);
    irradiance_texture = texE(dE);                              // binding 4 = delta_irradiance, :1417-1429
    std::vector<double> full;
    if (idx) full.assign((size_t)SW() * SH() * SD() * 4, 0.0);
    scattering_density = tex3(idx ? full.data() : out);
    dispatch3(idx, n);
    if (idx) gather(full, idx, n, out);
    return 0;
}
#elif defined(SHADER_INDIRECT_IRRADIANCE)
extern "C" int fbr_indirect_irradiance(const void* params, int order, const double* dR, const double* dM, const double* dMS, double* dE, double* E) {
    load_params(params);
    scattering_order = order;                                   // push constant, :1941-1947
    single_rayleigh_scattering_texture = tex3(dR); single_mie_scattering_texture = tex3(dM);
    multiple_scattering_texture = tex3(dMS);
    delta_irradiance = texE(dE);
    irradiance = texE(E);
    dispatch2(irradiance.w, irradiance.h);
    return 0;
}
#elif defined(SHADER_MULTIPLE_SCATTERING)
extern "C" int fbr_multiple_scattering(const void* params, const double* T, const double* dens, const int64_t* idx, int64_t n, double* dMS, double* S) {
    load_params(params);
    transmittance_texture = texT(T);
    scattering_density_texture = tex3(dens);
    const size_t total = (size_t)SW() * SH() * SD() * 4;
    std::vector<double> fD, fS;
    if (idx) {
        fD.assign(total, 0.0); fS.assign(total, 0.0);
        for (int64_t k = 0; k < n; ++k) std::memcpy(fS.data() + 4 * idx[k], S + 4 * k, 4 * sizeof(double));   // S arrives as gathered rows
    }
    delta_multiple_scattering = tex3(idx ? fD.data() : dMS);
    scattering = tex3(idx ? fS.data() : S);
    dispatch3(idx, n);
    if (idx) { gather(fD, idx, n, dMS); gather(fS, idx, n, S); }
    return 0;
}
#elif defined(SHADER_RENDER_SKY)
extern "C" int fbr_render(const void* params, const double* T, const double* S, const float* draw, const float* depth, int w, int h,
                          const int64_t* idx, int64_t n, double* color, double* transm) {
    load_params(params);
    transmittance_texture = texT(T);
    scattering_texture = tex3(S);
    for (int c = 0; c < 4; ++c) inverse_viewproj.c[c] = vec4(draw[4 * c], draw[4 * c + 1], draw[4 * c + 2], draw[4 * c + 3]);   // DrawParamsRaw, render.rs:254-260
    camera_position = vec3(draw[16], draw[17], draw[18]);
    sun_direction = vec3(draw[20], draw[21], draw[22]);
    std::vector<double> d4((size_t)w * h * 4, 0.0);
    for (size_t k = 0; k < (size_t)w * h; ++k) d4[4 * k] = depth[k];
    depth_buffer = img(d4.data(), w, h, 1, false);
    const int64_t count = idx ? n : (int64_t)w * h;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < count; ++k) {
        const int64_t t = idx ? idx[k] : k;
        const int px = (int)(t % w), py = (int)(t / w);
        frag_pixel = ivec2(px, py);
        screen_coords = vec2(((float)px + 0.5f) / (float)w, ((float)py + 0.5f) / (float)h);     // fullscreen.vert:6 at the pixel centre
        shader_main();
        const float c[4] = {color_out.x, color_out.y, color_out.z, color_out.w}, tr[4] = {transmittance_out.x, transmittance_out.y, transmittance_out.z, transmittance_out.w};
        for (int j = 0; j < 4; ++j) { color[4 * k + j] = c[j]; transm[4 * k + j] = tr[j]; }
    }
    return 0;
}
#endif
