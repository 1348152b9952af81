// fb_kernels_fast.cu — FB_KERNELS_FAST (placeholder while the restructured kernels land: forwards
// to the reference family so the ABI is complete).
#include "fb_kernels.h"

namespace fb {
namespace fast {
size_t scratch_bytes(const FbParams&) { return 0; }
cudaError_t transmittance(const LaunchCtx& c) { return ref::transmittance(c); }
cudaError_t direct_irradiance(const LaunchCtx& c) { return ref::direct_irradiance(c); }
cudaError_t single_scattering(const LaunchCtx& c, int r0, int r1) { return ref::single_scattering(c, r0, r1); }
cudaError_t scattering_density(const LaunchCtx& c, int order, int r0, int r1) { return ref::scattering_density(c, order, r0, r1); }
cudaError_t indirect_irradiance(const LaunchCtx& c, int order) { return ref::indirect_irradiance(c, order); }
cudaError_t multiple_scattering(const LaunchCtx& c, int r0, int r1) { return ref::multiple_scattering(c, r0, r1); }
int launches_per_stage(int) { return 1; }
}  // namespace fast
}  // namespace fb
