// fb_kernels.h — host-callable launchers shared by the kernel translation units and fb_api.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fuzzyblue.h"

namespace fb {

// Quadrature direction tables, computed once on the host in fp32 with the C library's
// cosf/sinf: theta_l = (l + 0.5) * pi/16, phi_m = (m + 0.5) * pi/16 for scattering_density.comp:34-66
// and theta_j = (j + 0.5) * pi/32, phi_i = (i + 0.5) * pi/32 for indirect_irradiance.comp:20-33.
struct Trig {
    float ct16[16], st16[16], cp32[32], sp32[32];
    float ct32[16], st32[16], cp64[64], sp64[64];
};
void make_trig(Trig* t);

// Device images of one in-flight precompute (linear layouts, see include/fuzzyblue.h).
struct Images {
    float4* transmittance;
    float4* irradiance;
    uint2* scattering;
    float4* delta_irradiance;
    uint2* delta_rayleigh;
    uint2* delta_mie;
    uint2* scattering_density;
    uint2* delta_multiple_scattering;
    // scratch of the FAST kernels (not images of the reference)
    float* scratch;
    size_t scratch_bytes;
};

struct LaunchCtx {
    FbParams P;
    Trig trig;
    Images img;
    int sm_count;
    cudaStream_t stream;
};

// One-thread-per-texel, contraction-free transcription of the six compute shaders.
namespace ref {
cudaError_t transmittance(const LaunchCtx& c);
cudaError_t direct_irradiance(const LaunchCtx& c);
cudaError_t single_scattering(const LaunchCtx& c, int r0, int r1);
cudaError_t scattering_density(const LaunchCtx& c, int order, int r0, int r1);
cudaError_t indirect_irradiance(const LaunchCtx& c, int order, int row0, int row1);   // rows [row0, row1) of the irradiance table
cudaError_t multiple_scattering(const LaunchCtx& c, int r0, int r1);
}  // namespace ref

// Restructured sm_100a kernels (the product path).
namespace fast {
size_t scratch_bytes(const FbParams& P);
cudaError_t transmittance(const LaunchCtx& c);
cudaError_t direct_irradiance(const LaunchCtx& c);
cudaError_t single_scattering(const LaunchCtx& c, int r0, int r1);
// `after_prep` (optional) is recorded on c.stream once the stage no longer reads delta_irradiance
cudaError_t scattering_density(const LaunchCtx& c, int order, int r0, int r1, cudaEvent_t after_prep = nullptr);
cudaError_t indirect_irradiance(const LaunchCtx& c, int order, int row0, int row1);
cudaError_t multiple_scattering(const LaunchCtx& c, int r0, int r1);
// false when a stage of these dims runs the one-thread-per-texel transcription instead (a ~30x cliff the API reports)
bool density_is_fast(const FbParams& P);
bool multiple_is_fast(const FbParams& P);
int launches_per_stage(const FbParams& P, int stage, int r_count);   // kernels one stage launches over r_count altitude levels
}  // namespace fast

// render_sky.frag and the library queries.
// `draws_host` holds `views` draw blocks in host memory; a sweep (views > 1) also needs `view_records_dev`, device
// scratch of views * render_view_record_bytes() that receives the draw blocks and their per-view constants.
size_t render_view_record_bytes();
// `expanded` (may be NULL): render_expanded_bytes() of device memory holding render_expand_scattering()'s output for this
// scattering table — the FAST path's private (value, delta-to-next-x) fp32 form of it.
size_t render_expanded_bytes(const FbParams& P);
cudaError_t render_expand_scattering(const FbParams& P, const float4* transmittance, const uint2* scattering, void* expanded,
                                     cudaStream_t s);
// `view_tables_dev` (may be NULL): views * render_view_table_bytes() of device scratch for the per-view sky tables of the
// FAST path (fb_render.cu: k_view_tables), rewritten by every call.
size_t render_view_table_bytes(const FbParams& P);
cudaError_t render_sky(const FbParams& P, const float4* transmittance, const uint2* scattering, const void* expanded,
                       const FbDrawParams* draws_host, void* view_records_dev, void* view_tables_dev, uint32_t views, const float* depth,
                       float4* color, float4* transm, float4* blend_fb, uint32_t w, uint32_t h, int kernels, cudaStream_t s);
cudaError_t sky_radiance(const FbParams& P, const float4* transmittance, const uint2* scattering, const float* cam,
                         const float* view, const float* sun, uint64_t n, float* radiance, float* transm, cudaStream_t s);
cudaError_t sun_sky_irradiance(const FbParams& P, const float4* transmittance, const float4* irradiance, const float* point,
                               const float* normal, const float* sun, uint64_t n, float* direct, float* sky, cudaStream_t s);

cudaError_t measure_peaks(int sm_count, double* fma_tflops, double* sfu_gops);

}  // namespace fb
