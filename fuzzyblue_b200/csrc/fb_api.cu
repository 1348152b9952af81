// fb_api.cu — the extern "C" boundary declared in include/fuzzyblue.h: object lifetimes, the
// per-order command stream of Atmosphere::build (src/precompute.rs:1671-2048) as stream-ordered
// launches or a replayable CUDA graph, read-back, the renderer front-end and batches.
//
// There is no CPU fallback anywhere in this file: without a CUDA device fb_builder_create fails
// with FB_ERR_NO_DEVICE and nothing else can be constructed.
#include <cuda.h>           // driver-API TYPES only (virtual memory management); entry points come from cudaGetDriverEntryPoint
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/fuzzyblue.h"
#include "fb_internal.h"
#include "fb_kernels.h"

using namespace fb;

static_assert(sizeof(FbParams) == 320, "FbParams must mirror the 320-byte std140 block (precompute.rs:1675)");
static_assert(offsetof(FbParams, transmittance_mu_size) == 92, "sizes at 92");
static_assert(offsetof(FbParams, rayleigh_density) == 128, "profiles at 128");
static_assert(offsetof(FbParams, absorption_density) == 256, "absorption profile at 256");
static_assert(sizeof(FbDrawParams) == 92, "FbDrawParams must mirror the 92-byte push-constant block (render.rs:231)");
static_assert(offsetof(FbDrawParams, sun_direction) == 80, "sun_direction after the pad");

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

namespace fb {
int fail(int status, const std::string& msg) {
    g_last_error = msg;
    return status;
}
int cuda_fail(cudaError_t e, const char* what) {
    (void)cudaGetLastError();   // clear the sticky-less error so later calls are not poisoned
    int st = (e == cudaErrorMemoryAllocation) ? FB_ERR_OUT_OF_MEMORY : FB_ERR_CUDA;
    return fail(st, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
}  // namespace fb

const char* fb_status_string(int s) {
    switch (s) {
        case FB_OK: return "FB_OK";
        case FB_ERR_INVALID_ARGUMENT: return "FB_ERR_INVALID_ARGUMENT";
        case FB_ERR_CUDA: return "FB_ERR_CUDA";
        case FB_ERR_OUT_OF_MEMORY: return "FB_ERR_OUT_OF_MEMORY";
        case FB_ERR_NO_DEVICE: return "FB_ERR_NO_DEVICE";
        case FB_ERR_NOT_READY: return "FB_ERR_NOT_READY";
        default: return "FB_ERR_UNKNOWN";
    }
}
const char* fb_last_error(void) { return g_last_error.c_str(); }
const char* fb_version(void) { return "fuzzyblue_b200 0.1.0 (sm_100a)"; }

// ---------------------------------------------------------------------------------------------
// parameters
// ---------------------------------------------------------------------------------------------
int fb_params_default(FbParams* p) {   // src/precompute.rs:849-935
    if (!p) return fail(FB_ERR_INVALID_ARGUMENT, "fb_params_default: NULL");
    std::memset(p, 0, sizeof *p);
    const float solar[3] = {1.474f, 1.850f, 1.91198f};
    const float ray[3] = {0.005802f, 0.013558f, 0.033100f};
    const float ozone[3] = {6.5e-4f, 1.881e-3f, 8.5e-5f};
    for (int i = 0; i < 3; ++i) {
        p->solar_irradiance[i] = solar[i];
        p->rayleigh_scattering[i] = ray[i];
        p->mie_scattering[i] = 0.003996f;
        p->mie_extinction[i] = 0.004440f;
        p->ground_albedo[i] = 0.1f;
        p->absorption_extinction[i] = ozone[i];
    }
    p->sun_angular_radius = 0.004675f;
    p->bottom_radius = 6360.0f;
    p->top_radius = 6420.0f;
    p->mie_phase_function_g = 0.8f;
    p->mu_s_min = -0.207912f;
    p->transmittance_mu_size = 256;
    p->transmittance_r_size = 64;
    p->scattering_r_size = 32;
    p->scattering_mu_size = 128;
    p->scattering_mu_s_size = 32;
    p->scattering_nu_size = 8;
    p->irradiance_mu_s_size = 64;
    p->irradiance_r_size = 16;
    p->rayleigh_density.layers[1].exp_term = 1.0f;
    p->rayleigh_density.layers[1].exp_scale = -0.125f;
    p->mie_density.layers[1].exp_term = 1.0f;
    p->mie_density.layers[1].exp_scale = -0.833333f;
    p->absorption_density.layers[0].width = 25.0f;
    p->absorption_density.layers[0].linear_term = 0.066667f;
    p->absorption_density.layers[0].constant_term = -0.666667f;
    p->absorption_density.layers[1].linear_term = -0.066667f;
    p->absorption_density.layers[1].constant_term = 2.666667f;
    return FB_OK;
}
uint32_t fb_params_default_order(void) { return 4; }   // precompute.rs:856

int fb_params_validate(const FbParams* p) {
    if (!p) return fail(FB_ERR_INVALID_ARGUMENT, "params: NULL");
    const int32_t s[8] = {p->transmittance_mu_size, p->transmittance_r_size, p->scattering_r_size, p->scattering_mu_size,
                          p->scattering_mu_s_size, p->scattering_nu_size, p->irradiance_mu_s_size, p->irradiance_r_size};
    for (int i = 0; i < 8; ++i)
        if (s[i] < 2 || s[i] > 16384)
            return fail(FB_ERR_INVALID_ARGUMENT, "params: every LUT size must be in [2, 16384] (x/(size-1) mappings, "
                                                 "scattering.h:120-131)");
    if (p->scattering_mu_size % 2) return fail(FB_ERR_INVALID_ARGUMENT, "params: scattering_mu_size must be even (scattering.h:36)");
    if (p->scattering_mu_size > 65535 || p->scattering_r_size > 65535)
        return fail(FB_ERR_INVALID_ARGUMENT, "params: scattering mu/r size exceeds the launch grid");
    int64_t texels = (int64_t)p->scattering_nu_size * p->scattering_mu_s_size * p->scattering_mu_size * p->scattering_r_size;
    if (texels > ((int64_t)1 << 31)) return fail(FB_ERR_INVALID_ARGUMENT, "params: scattering table exceeds 2^31 texels");
    if (!(p->top_radius > p->bottom_radius) || !(p->bottom_radius > 0.f))
        return fail(FB_ERR_INVALID_ARGUMENT, "params: need 0 < bottom_radius < top_radius");
    // The restructured kernels address shared-memory tables from geometry without per-sample clamps, which is safe for
    // finite geometry only: a NaN / infinite scalar would become a wild address (a sticky device fault) where the
    // reference merely writes NaN texels.  Reject such blocks up front.
    const float* f = reinterpret_cast<const float*>(p);
    for (int i = 0; i < 23; ++i)                                        // the 92 bytes of floats ahead of the sizes
        if (!std::isfinite(f[i])) return fail(FB_ERR_INVALID_ARGUMENT, "params: non-finite value in the parameter block");
    const FbDensityProfile* prof[3] = {&p->rayleigh_density, &p->mie_density, &p->absorption_density};
    for (const FbDensityProfile* d : prof)
        for (const FbDensityProfileLayer& l : d->layers)
            if (!std::isfinite(l.width) || !std::isfinite(l.exp_term) || !std::isfinite(l.exp_scale) || !std::isfinite(l.linear_term) ||
                !std::isfinite(l.constant_term))
                return fail(FB_ERR_INVALID_ARGUMENT, "params: non-finite value in a density profile");
    if (!(p->mu_s_min >= -1.f && p->mu_s_min <= 0.f))
        return fail(FB_ERR_INVALID_ARGUMENT, "params: mu_s_min must lie in [-1, 0] (cosine of the largest sun zenith angle tabulated, scattering.h:50-56)");
    if (!(p->mie_phase_function_g > -1.f && p->mie_phase_function_g < 1.f))
        return fail(FB_ERR_INVALID_ARGUMENT, "params: mie_phase_function_g must lie in (-1, 1) (util.h:31-34)");
    return FB_OK;
}
int fb_params_transmittance_extent(const FbParams* p, FbExtent2D* o) {   // precompute.rs:772-777
    if (!p || !o) return fail(FB_ERR_INVALID_ARGUMENT, "extent: NULL");
    o->width = p->transmittance_mu_size; o->height = p->transmittance_r_size;
    return FB_OK;
}
int fb_params_irradiance_extent(const FbParams* p, FbExtent2D* o) {      // :779-784
    if (!p || !o) return fail(FB_ERR_INVALID_ARGUMENT, "extent: NULL");
    o->width = p->irradiance_mu_s_size; o->height = p->irradiance_r_size;
    return FB_OK;
}
int fb_params_scattering_extent(const FbParams* p, FbExtent3D* o) {      // :786-792
    if (!p || !o) return fail(FB_ERR_INVALID_ARGUMENT, "extent: NULL");
    o->width = p->scattering_nu_size * p->scattering_mu_s_size; o->height = p->scattering_mu_size; o->depth = p->scattering_r_size;
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------------------------
// Completion and BlockCache (fb_internal.h): stream-ordered reuse of released device blocks.
namespace fb {
void Completion::note(cudaStream_t s) {
    std::lock_guard<std::mutex> g(m);
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess) { (void)cudaGetLastError(); unknown = true; return; }
    if (cs != cudaStreamCaptureStatusNone) { unknown = true; return; }   // the caller records us into a graph of its own
    for (auto& kv : ev)
        if (kv.first == s) {
            if (cudaEventRecord(kv.second, s) != cudaSuccess) { (void)cudaGetLastError(); unknown = true; }
            return;
        }
    if (ev.size() >= 32) { unknown = true; return; }
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(e, s) != cudaSuccess) {
        (void)cudaGetLastError();
        if (e) cudaEventDestroy(e);
        unknown = true;
        return;
    }
    ev.emplace_back(s, e);
}
cudaError_t Completion::query() {
    std::lock_guard<std::mutex> g(m);
    if (unknown) return cudaDeviceSynchronize();      // untracked work: the only safe answer is to drain the device
    for (auto& kv : ev) {
        const cudaError_t e = cudaEventQuery(kv.second);
        if (e != cudaSuccess) return e;               // cudaErrorNotReady or a real error
    }
    return cudaSuccess;
}
cudaError_t Completion::wait() {
    std::lock_guard<std::mutex> g(m);
    if (unknown) {
        const cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) unknown = false;
        return e;
    }
    for (auto& kv : ev) {
        const cudaError_t e = cudaEventSynchronize(kv.second);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
cudaError_t Completion::stream_wait(cudaStream_t s) {
    std::unique_lock<std::mutex> g(m);
    if (unknown) { g.unlock(); return wait(); }
    for (auto& kv : ev) {
        if (kv.first == s) continue;                  // stream order already covers it
        const cudaError_t e = cudaStreamWaitEvent(s, kv.second, 0);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
Completion::~Completion() {
    for (auto& kv : ev) cudaEventDestroy(kv.second);
}

cudaError_t BlockCache::get(void** p, size_t n) {
    Block blk{nullptr, nullptr};
    {
        std::lock_guard<std::mutex> g(m);
        auto it = free_blocks.find(n);
        if (it != free_blocks.end()) {
            blk = it->second;
            free_blocks.erase(it);
            cached -= n;
        }
    }
    if (blk.p) {
        // the previous owner's kernels may still be in flight on its streams: wait for exactly that work
        if (blk.busy) {
            const cudaError_t e = blk.busy->wait();
            if (e != cudaSuccess) { cudaFree(blk.p); return e; }
        }
        *p = blk.p;
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, n);
    if (e == cudaErrorMemoryAllocation) {   // give the cached blocks back to the driver and retry once
        (void)cudaGetLastError();
        trim();
        e = cudaMalloc(p, n);
    }
    return e;
}
void BlockCache::put(void* p, size_t n, const std::shared_ptr<Completion>& busy) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> g(m);
        if (cached + n <= limit) {
            free_blocks.emplace(n, Block{p, busy});
            cached += n;
            return;
        }
    }
    cudaFree(p);                            // synchronises the device itself
}
void BlockCache::trim() {
    std::lock_guard<std::mutex> g(m);
    for (auto& kv : free_blocks) cudaFree(kv.second.p);
    free_blocks.clear();
    cached = 0;
}
BlockCache::~BlockCache() {
    int prev = -1;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) cudaSetDevice(device);
    trim();
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
}
}  // namespace fb

// ---------------------------------------------------------------------------------------------
// Exportable device memory: the kept block as a CUDA virtual-memory allocation whose POSIX
// file descriptor a Vulkan caller imports (VkImportMemoryFdInfoKHR, OPAQUE_FD) — the counterpart of
// Atmosphere::{transmittance,scattering,irradiance}() handing out vk::Image (precompute.rs:2075-2101).  The driver
// entry points are fetched at run time, so the library carries no link-time dependency on libcuda.
// ---------------------------------------------------------------------------------------------
struct Vmm {
    CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
    CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
    CUresult (*Release)(CUmemGenericAllocationHandle);
    CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
    CUresult (*AddressFree)(CUdeviceptr, size_t);
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
    CUresult (*Unmap)(CUdeviceptr, size_t);
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
    CUresult (*Export)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
    CUresult (*Import)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
    bool ok;
};
static const Vmm& vmm_api() {
    static Vmm v = [] {
        Vmm t;
        std::memset(&t, 0, sizeof t);
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        t.ok = get("cuMemGetAllocationGranularity", (void**)&t.GetGranularity) && get("cuMemCreate", (void**)&t.Create) &&
               get("cuMemRelease", (void**)&t.Release) && get("cuMemAddressReserve", (void**)&t.AddressReserve) &&
               get("cuMemAddressFree", (void**)&t.AddressFree) && get("cuMemMap", (void**)&t.Map) &&
               get("cuMemUnmap", (void**)&t.Unmap) && get("cuMemSetAccess", (void**)&t.SetAccess) &&
               get("cuMemExportToShareableHandle", (void**)&t.Export) && get("cuMemImportFromShareableHandle", (void**)&t.Import);
        (void)cudaGetLastError();
        return t;
    }();
    return v;
}
static CUmemAllocationProp vmm_prop(int device) {
    CUmemAllocationProp prop;
    std::memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}
// map `handle` (size bytes) read-write on `device`; 0 on failure
static CUdeviceptr vmm_map(const Vmm& v, int device, CUmemGenericAllocationHandle handle, size_t size) {
    CUdeviceptr ptr = 0;
    if (v.AddressReserve(&ptr, size, 0, 0, 0) != CUDA_SUCCESS) return 0;
    if (v.Map(ptr, size, 0, handle, 0) != CUDA_SUCCESS) { v.AddressFree(ptr, size); return 0; }
    CUmemAccessDesc acc;
    std::memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (v.SetAccess(ptr, size, &acc, 1) != CUDA_SUCCESS) { v.Unmap(ptr, size); v.AddressFree(ptr, size); return 0; }
    return ptr;
}
static bool vmm_alloc(int device, size_t bytes, void** out, unsigned long long* handle, size_t* alloc_bytes) {
    const Vmm& v = vmm_api();
    if (!v.ok) return false;
    (void)cudaFree(0);                               // make sure the primary context exists and is current
    const CUmemAllocationProp prop = vmm_prop(device);
    size_t gran = 0;
    if (v.GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || !gran) return false;
    const size_t size = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    if (v.Create(&h, size, &prop, 0) != CUDA_SUCCESS) return false;
    const CUdeviceptr ptr = vmm_map(v, device, h, size);
    if (!ptr) { v.Release(h); return false; }
    *out = reinterpret_cast<void*>(ptr);
    *handle = (unsigned long long)h;
    *alloc_bytes = size;
    return true;
}
static void vmm_free(void* p, unsigned long long handle, size_t alloc_bytes) {
    const Vmm& v = vmm_api();
    if (!v.ok || !p) return;
    (void)cudaDeviceSynchronize();                   // cudaFree semantics: nothing may still be using the block
    v.Unmap((CUdeviceptr)p, alloc_bytes);
    v.AddressFree((CUdeviceptr)p, alloc_bytes);
    v.Release((CUmemGenericAllocationHandle)handle);
}

static std::atomic<uint64_t> g_atmosphere_serial{1};
constexpr int RB_SLABS = 4;
static int rb_slabs() {      // FUZZYBLUE_B200_RB_SLABS=1..4 overrides the slab count (tuning experiments)
    static const int n = [] {
        const char* e = std::getenv("FUZZYBLUE_B200_RB_SLABS");
        const int v = e ? std::atoi(e) : RB_SLABS;
        return v < 1 ? 1 : (v > RB_SLABS ? RB_SLABS : v);
    }();
    return n;
}

struct FbRenderer {
    int device;
    int kernels;
    void* sweep_draws;           // device copy of a sweep's draw parameters + per-view constants
    uint32_t sweep_capacity;
    void* view_tables;           // per-view sky tables of the FAST path (fb_render.cu: k_view_tables), rewritten by every draw
    size_t view_tables_bytes;
    // the FAST path's expanded copy of one atmosphere's scattering table (fb_render.cu: Tex3X), rebuilt when a draw
    // names other table contents; draws on other streams wait for the rebuild through `expanded_ready`
    void* expanded;
    size_t expanded_bytes;
    uint64_t expanded_serial, expanded_version;
    cudaEvent_t expanded_ready;
    bool no_expand;              // FUZZYBLUE_B200_RENDER_FP16_TABLE set at creation: always tap the fp16 table (A/B tests)
    std::shared_ptr<Completion> draws;   // every draw issued through this renderer, per stream: what a table rebuild waits for
};
constexpr size_t EXPANDED_MAX_BYTES = (size_t)256 << 20;

static size_t bytes2d(int w, int h) { return (size_t)w * h * sizeof(float4); }
static size_t bytes3d(const FbParams& P) {
    return (size_t)P.scattering_nu_size * P.scattering_mu_s_size * P.scattering_mu_size * P.scattering_r_size * sizeof(uint2);
}
namespace fb {
size_t image_bytes(const FbParams& P, int image) {
    switch (image) {
        case FB_IMAGE_TRANSMITTANCE: return bytes2d(P.transmittance_mu_size, P.transmittance_r_size);
        case FB_IMAGE_IRRADIANCE:
        case FB_IMAGE_DELTA_IRRADIANCE: return bytes2d(P.irradiance_mu_s_size, P.irradiance_r_size);
        case FB_IMAGE_SCATTERING:
        case FB_IMAGE_DELTA_RAYLEIGH:
        case FB_IMAGE_DELTA_MIE:
        case FB_IMAGE_SCATTERING_DENSITY:
        case FB_IMAGE_DELTA_MULTIPLE_SCATTERING: return bytes3d(P);
        default: return 0;
    }
}
void* image_ptr(FbPending* p, int image) {
    switch (image) {
        case FB_IMAGE_TRANSMITTANCE: return p->img.transmittance;
        case FB_IMAGE_IRRADIANCE: return p->img.irradiance;
        case FB_IMAGE_SCATTERING: return p->img.scattering;
        case FB_IMAGE_DELTA_IRRADIANCE: return p->img.delta_irradiance;
        case FB_IMAGE_DELTA_RAYLEIGH: return p->img.delta_rayleigh;
        case FB_IMAGE_DELTA_MIE: return p->img.delta_mie;
        case FB_IMAGE_SCATTERING_DENSITY: return p->img.scattering_density;
        case FB_IMAGE_DELTA_MULTIPLE_SCATTERING: return p->img.delta_multiple_scattering;
        default: return nullptr;
    }
}
}  // namespace fb

int fb_builder_create(int device, FbBuilder** out) {
    if (!out) return fail(FB_ERR_INVALID_ARGUMENT, "fb_builder_create: NULL out");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        return fail(FB_ERR_NO_DEVICE, "fb_builder_create: no CUDA device is visible; this library has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(FB_ERR_INVALID_ARGUMENT, "fb_builder_create: device ordinal out of range");
    cudaDeviceProp prop;
    FB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(FB_ERR_NO_DEVICE, std::string("fb_builder_create: device '") + prop.name +
                                          "' is not an sm_100 part; the kernels are built for sm_100a only");
    FbBuilder* b = new (std::nothrow) FbBuilder();
    if (!b) return fail(FB_ERR_OUT_OF_MEMORY, "fb_builder_create: host allocation");
    b->device = device;
    b->sm_count = prop.multiProcessorCount;
    b->kernels = FB_KERNELS_FAST;
    b->exportable = 0;
    b->cache = std::make_shared<BlockCache>(device);
    make_trig(&b->trig);
    *out = b;
    return FB_OK;
}
void fb_builder_destroy(FbBuilder* b) { delete b; }
int fb_builder_set_kernels(FbBuilder* b, int k) {
    if (!b || (k != FB_KERNELS_FAST && k != FB_KERNELS_REFERENCE)) return fail(FB_ERR_INVALID_ARGUMENT, "fb_builder_set_kernels");
    b->kernels = k;
    return FB_OK;
}
int fb_builder_measure_peaks(FbBuilder* b, double* fp32_fma_tflops, double* sfu_gops) {
    if (!b) return fail(FB_ERR_INVALID_ARGUMENT, "fb_builder_measure_peaks: NULL");
    DeviceGuard g(b->device);
    cudaError_t e = measure_peaks(b->sm_count, fp32_fma_tflops, sfu_gops);
    return e == cudaSuccess ? FB_OK : cuda_fail(e, "measure_peaks");
}
int fb_builder_trim(FbBuilder* b) {
    if (!b) return fail(FB_ERR_INVALID_ARGUMENT, "fb_builder_trim: NULL");
    DeviceGuard g(b->device);
    b->cache->trim();
    return FB_OK;
}
int fb_builder_set_exportable(FbBuilder* b, int on) {
    if (!b) return fail(FB_ERR_INVALID_ARGUMENT, "fb_builder_set_exportable: NULL");
    if (on && !vmm_api().ok) return fail(FB_ERR_CUDA, "fb_builder_set_exportable: the driver lacks the virtual memory management API");
    b->exportable = on ? 1 : 0;
    return FB_OK;
}

int fb_atmosphere_export_fd(const FbAtmosphere* a, int* fd, FbExportLayout* layout) {
    if (!a || !fd) return fail(FB_ERR_INVALID_ARGUMENT, "fb_atmosphere_export_fd: NULL");
    *fd = -1;
    if (!a->vmm) return fail(FB_ERR_INVALID_ARGUMENT, "fb_atmosphere_export_fd: build the atmosphere after fb_builder_set_exportable(b, 1)");
    DeviceGuard g(a->device);
    int out = -1;
    const CUresult r = vmm_api().Export(&out, (CUmemGenericAllocationHandle)a->vmm_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) return fail(FB_ERR_CUDA, "cuMemExportToShareableHandle failed (" + std::to_string((int)r) + ")");
    *fd = out;
    if (layout) {
        layout->allocation_bytes = a->vmm_bytes;
        layout->scattering_offset = 0;
        layout->scattering_bytes = bytes3d(a->P);
        layout->transmittance_offset = a->off_transmittance;
        layout->transmittance_bytes = image_bytes(a->P, FB_IMAGE_TRANSMITTANCE);
        layout->irradiance_offset = a->off_irradiance;
        layout->irradiance_bytes = image_bytes(a->P, FB_IMAGE_IRRADIANCE);
    }
    return FB_OK;
}

int fb_external_memory_read_fd(int device, int fd, size_t allocation_bytes, size_t offset, void* host, size_t bytes) {
    if (fd < 0 || !host || offset + bytes > allocation_bytes) return fail(FB_ERR_INVALID_ARGUMENT, "fb_external_memory_read_fd");
    const Vmm& v = vmm_api();
    if (!v.ok) return fail(FB_ERR_CUDA, "the driver lacks the virtual memory management API");
    DeviceGuard g(device);
    if (!g.ok) return fail(FB_ERR_CUDA, "cudaSetDevice failed");
    (void)cudaFree(0);
    CUmemGenericAllocationHandle h;
    CUresult r = v.Import(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r != CUDA_SUCCESS) return fail(FB_ERR_CUDA, "cuMemImportFromShareableHandle failed (" + std::to_string((int)r) + ")");
    const CUdeviceptr ptr = vmm_map(v, device, h, allocation_bytes);
    if (!ptr) { v.Release(h); return fail(FB_ERR_CUDA, "mapping the imported allocation failed"); }
    const cudaError_t e = cudaMemcpy(host, reinterpret_cast<const char*>(ptr) + offset, bytes, cudaMemcpyDeviceToHost);
    v.Unmap(ptr, allocation_bytes);
    v.AddressFree(ptr, allocation_bytes);
    v.Release(h);
    return e == cudaSuccess ? FB_OK : cuda_fail(e, "cudaMemcpy from the imported allocation");
}

struct FbExternalSemaphore {
    int device;
    cudaExternalSemaphore_t sem;
    int timeline;
};
int fb_external_semaphore_import_fd(int device, int fd, int is_timeline, FbExternalSemaphore** out) {
    if (!out || fd < 0) return fail(FB_ERR_INVALID_ARGUMENT, "fb_external_semaphore_import_fd");
    *out = nullptr;
    DeviceGuard g(device);
    if (!g.ok) return fail(FB_ERR_CUDA, "cudaSetDevice failed");
    cudaExternalSemaphoreHandleDesc d;
    std::memset(&d, 0, sizeof d);
    d.type = is_timeline ? cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd : cudaExternalSemaphoreHandleTypeOpaqueFd;
    d.handle.fd = fd;                                // ownership of the descriptor passes to CUDA on success
    cudaExternalSemaphore_t sem;
    FB_CUDA(cudaImportExternalSemaphore(&sem, &d));
    FbExternalSemaphore* s = new (std::nothrow) FbExternalSemaphore();
    if (!s) { cudaDestroyExternalSemaphore(sem); return fail(FB_ERR_OUT_OF_MEMORY, "host allocation"); }
    s->device = device; s->sem = sem; s->timeline = is_timeline ? 1 : 0;
    *out = s;
    return FB_OK;
}
int fb_external_semaphore_signal(FbExternalSemaphore* s, uint64_t value, void* stream) {
    if (!s) return fail(FB_ERR_INVALID_ARGUMENT, "fb_external_semaphore_signal: NULL");
    DeviceGuard g(s->device);
    cudaExternalSemaphoreSignalParams p;
    std::memset(&p, 0, sizeof p);
    p.params.fence.value = s->timeline ? value : 0;
    FB_CUDA(cudaSignalExternalSemaphoresAsync(&s->sem, &p, 1, (cudaStream_t)stream));
    return FB_OK;
}
int fb_external_semaphore_wait(FbExternalSemaphore* s, uint64_t value, void* stream) {
    if (!s) return fail(FB_ERR_INVALID_ARGUMENT, "fb_external_semaphore_wait: NULL");
    DeviceGuard g(s->device);
    cudaExternalSemaphoreWaitParams p;
    std::memset(&p, 0, sizeof p);
    p.params.fence.value = s->timeline ? value : 0;
    FB_CUDA(cudaWaitExternalSemaphoresAsync(&s->sem, &p, 1, (cudaStream_t)stream));
    return FB_OK;
}
void fb_external_semaphore_destroy(FbExternalSemaphore* s) {
    if (!s) return;
    DeviceGuard g(s->device);
    cudaDestroyExternalSemaphore(s->sem);
    delete s;
}

int fb_builder_device(const FbBuilder* b) { return b ? b->device : -1; }
int fb_builder_sm_count(const FbBuilder* b) { return b ? b->sm_count : 0; }

static void free_pending_temps(FbPending* p) {
    // work enqueued on the internal streams always joins the caller's stream before a submission returns, but an
    // enqueue that failed half-way may have left some behind: make them part of what the block's next owner waits for
    if (p->side) p->done->note(p->side);
    for (cudaStream_t q : p->rb_stream) if (q) p->done->note(q);
    if (p->comm) p->done->note(p->comm);
    p->cache->put(p->temp_block, p->temp_bytes, p->done);
    p->temp_block = nullptr;
    p->img.delta_irradiance = nullptr;
    p->img.delta_rayleigh = p->img.delta_mie = p->img.scattering_density = p->img.delta_multiple_scattering = nullptr;
    p->img.scratch = nullptr;
    if (p->graph) { cudaGraphExecDestroy(p->graph); p->graph = nullptr; }
    if (p->side) { cudaStreamDestroy(p->side); p->side = nullptr; }
    if (p->ev_fork) { cudaEventDestroy(p->ev_fork); p->ev_fork = nullptr; }
    if (p->ev_join) { cudaEventDestroy(p->ev_join); p->ev_join = nullptr; }
    for (cudaStream_t& q : p->rb_stream) if (q) { cudaStreamDestroy(q); q = nullptr; }
    if (p->rb_ev) { cudaEventDestroy(p->rb_ev); p->rb_ev = nullptr; }
    if (p->comm) { cudaStreamDestroy(p->comm); p->comm = nullptr; }
    if (p->ev_comm) { cudaEventDestroy(p->ev_comm); p->ev_comm = nullptr; }
}

void fb_atmosphere_destroy(FbAtmosphere* a) {   // Drop, precompute.rs:1045-1073
    if (!a) return;
    DeviceGuard g(a->device);
    if (a->vmm) vmm_free(a->block, a->vmm_handle, a->vmm_bytes);
    else a->cache->put(a->block, a->block_bytes, a->done);   // draws / reads still in flight: the next owner waits for them
    delete a;
}

void fb_pending_destroy(FbPending* p) {         // Drop, precompute.rs:2122-2140
    if (!p) return;
    DeviceGuard g(p->device);
    free_pending_temps(p);
    if (p->inner) fb_atmosphere_destroy(p->inner);
    delete p;
}

int fb_atmosphere_allocate(FbBuilder* b, const FbParams* params, uint32_t order, FbPending** out) {
    if (!b || !params || !out) return fail(FB_ERR_INVALID_ARGUMENT, "fb_atmosphere_allocate: NULL argument");
    *out = nullptr;
    int st = fb_params_validate(params);
    if (st != FB_OK) return st;
    if (order < 1 || order > 64) return fail(FB_ERR_INVALID_ARGUMENT, "order must be in [1, 64]");
    DeviceGuard g(b->device);
    if (!g.ok) return fail(FB_ERR_CUDA, "cudaSetDevice failed");
    FbPending* p = new (std::nothrow) FbPending();
    FbAtmosphere* a = new (std::nothrow) FbAtmosphere();
    if (!p || !a) { delete p; delete a; return fail(FB_ERR_OUT_OF_MEMORY, "host allocation"); }
    std::memset(&p->img, 0, sizeof p->img);
    p->device = b->device; p->kernels = b->kernels; p->sm_count = b->sm_count; p->trig = b->trig;
    p->done = std::make_shared<Completion>();
    a->done = p->done;
    p->slow_stages = 0;
    p->comm = nullptr; p->ev_comm = nullptr;
    p->cache = b->cache;
    a->cache = b->cache;
    p->P = *params;
    p->order = order;
    p->inner = a;
    p->graph = nullptr;
    p->launches = 0;
    p->side = nullptr; p->ev_fork = nullptr; p->ev_join = nullptr;
    p->rb_T = p->rb_S = p->rb_E = nullptr;
    for (cudaStream_t& q : p->rb_stream) q = nullptr;
    p->rb_ev = nullptr;
    a->device = b->device;
    a->serial = g_atmosphere_serial.fetch_add(1);
    a->version = 0;
    a->kernels = b->kernels;
    a->P = *params;
    a->transmittance = nullptr; a->irradiance = nullptr; a->scattering = nullptr;
    a->block = nullptr; p->temp_block = nullptr;
    // Two device blocks instead of the reference's one VkDeviceMemory per image (precompute.rs:1167-1234, :589-635):
    // what the Atmosphere keeps, and what dies with the PendingAtmosphere.  Sub-allocations are 256-byte aligned.
    auto up = [](size_t n) { return (n + 255) & ~(size_t)255; };
    const size_t b2t = up(image_bytes(*params, FB_IMAGE_TRANSMITTANCE)), b2e = up(image_bytes(*params, FB_IMAGE_IRRADIANCE)),
                 b3 = up(bytes3d(*params));
    p->img.scratch_bytes = fast::scratch_bytes(*params);
    a->block_bytes = b3 + b2t + b2e;
    p->temp_bytes = 4 * b3 + b2e + up(p->img.scratch_bytes);
    a->vmm = false; a->vmm_handle = 0; a->vmm_bytes = 0;
    a->off_transmittance = b3; a->off_irradiance = b3 + b2t;
    cudaError_t e = cudaSuccess;
    if (b->exportable) {
        if (!vmm_alloc(b->device, a->block_bytes, &a->block, &a->vmm_handle, &a->vmm_bytes)) {
            a->block = nullptr;
            fb_pending_destroy(p);
            return fail(FB_ERR_CUDA, "exportable allocation failed (cuMemCreate with a POSIX file-descriptor handle type)");
        }
        a->vmm = true;
    } else {
        e = b->cache->get(&a->block, a->block_bytes);
    }
    if (e == cudaSuccess) e = b->cache->get(&p->temp_block, p->temp_bytes);
    if (e != cudaSuccess) {
        fb_pending_destroy(p);
        return cuda_fail(e, "cudaMalloc(images)");
    }
    char* kb = static_cast<char*>(a->block);
    a->scattering = reinterpret_cast<uint2*>(kb);
    a->transmittance = reinterpret_cast<float4*>(kb + b3);
    a->irradiance = reinterpret_cast<float4*>(kb + b3 + b2t);
    char* tb = static_cast<char*>(p->temp_block);
    p->img.delta_rayleigh = reinterpret_cast<uint2*>(tb);
    p->img.delta_mie = reinterpret_cast<uint2*>(tb + b3);
    p->img.scattering_density = reinterpret_cast<uint2*>(tb + 2 * b3);
    p->img.delta_multiple_scattering = reinterpret_cast<uint2*>(tb + 3 * b3);
    p->img.delta_irradiance = reinterpret_cast<float4*>(tb + 4 * b3);
    p->img.scratch = p->img.scratch_bytes ? reinterpret_cast<float*>(tb + 4 * b3 + b2e) : nullptr;
    p->img.transmittance = a->transmittance;
    p->img.irradiance = a->irradiance;
    p->img.scattering = a->scattering;
    *out = p;
    return FB_OK;
}

namespace fb {
LaunchCtx make_ctx(FbPending* p, cudaStream_t s) {
    LaunchCtx c;
    c.P = p->P;
    c.trig = p->trig;
    c.img = p->img;
    c.sm_count = p->sm_count;
    c.stream = s;
    return c;
}

int run_stage(FbPending* p, const LaunchCtx& c, int stage, uint32_t order, int r0, int r1, int* launches) {
    const bool fastk = p->kernels == FB_KERNELS_FAST;
    if (fastk) {   // a stage these dims push onto the transcription kernels: remember it (fb_pending_slow_stages)
        if (stage == FB_STAGE_SCATTERING_DENSITY && (!fast::density_is_fast(c.P) || !c.img.scratch)) p->slow_stages |= 1u << stage;
        if (stage == FB_STAGE_MULTIPLE_SCATTERING && !fast::multiple_is_fast(c.P)) p->slow_stages |= 1u << stage;
    }
    // a launch reports errors through cudaGetLastError(): make sure an error some earlier, unrelated runtime call in
    // this thread left behind is not attributed to this stage
    (void)cudaGetLastError();
    cudaError_t e = cudaSuccess;
    int n = 1;
    switch (stage) {
        case FB_STAGE_TRANSMITTANCE: e = fastk ? fast::transmittance(c) : ref::transmittance(c); break;
        case FB_STAGE_DIRECT_IRRADIANCE: e = fastk ? fast::direct_irradiance(c) : ref::direct_irradiance(c); break;
        case FB_STAGE_SINGLE_SCATTERING: e = fastk ? fast::single_scattering(c, r0, r1) : ref::single_scattering(c, r0, r1); break;
        case FB_STAGE_SCATTERING_DENSITY:
            if (order < 2) return fail(FB_ERR_INVALID_ARGUMENT, "scattering_density needs order >= 2 (scattering_density.comp:22)");
            e = fastk ? fast::scattering_density(c, (int)order, r0, r1) : ref::scattering_density(c, (int)order, r0, r1);
            break;
        case FB_STAGE_INDIRECT_IRRADIANCE:
            if (order < 1) return fail(FB_ERR_INVALID_ARGUMENT, "indirect_irradiance needs order >= 1 (indirect_irradiance.comp:18)");
            if (r1 <= r0 || r1 > c.P.irradiance_r_size) return fail(FB_ERR_INVALID_ARGUMENT, "indirect_irradiance: bad row range");
            e = fastk ? fast::indirect_irradiance(c, (int)order, r0, r1) : ref::indirect_irradiance(c, (int)order, r0, r1);
            break;
        case FB_STAGE_MULTIPLE_SCATTERING: e = fastk ? fast::multiple_scattering(c, r0, r1) : ref::multiple_scattering(c, r0, r1); break;
        case FB_STAGE_CLEAR_IRRADIANCE:
            e = cudaMemsetAsync(c.img.irradiance, 0, image_bytes(c.P, FB_IMAGE_IRRADIANCE), c.stream);
            break;
        default: return fail(FB_ERR_INVALID_ARGUMENT, "unknown stage");
    }
    if (fastk && stage <= FB_STAGE_MULTIPLE_SCATTERING) n = fast::launches_per_stage(c.P, stage, r1 - r0);
    if (e != cudaSuccess) return cuda_fail(e, (std::string("stage ") + std::to_string(stage) + " launch").c_str());
    if (launches) *launches += n;
    return FB_OK;
}
}  // namespace fb

// The recorded command stream, src/precompute.rs:1671-2048.  Stream order replaces the pipeline
// barriers; the parameter block travels as a kernel argument instead of vkCmdUpdateBuffer.
static int enqueue_all(FbPending* p, cudaStream_t s, int* launches) {
    if (p->inner) ++p->inner->version;
    LaunchCtx c = make_ctx(p, s);
    const int R = p->P.scattering_r_size, ER = p->P.irradiance_r_size;
    int st;
    const bool fastk = p->kernels == FB_KERNELS_FAST;
    const bool overlap = fastk && p->order >= 2;
    const bool rb = p->rb_T || p->rb_S || p->rb_E;
    if ((overlap || rb) && !p->side) {
        FB_CUDA(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
        FB_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
        FB_CUDA(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
    }
    if (rb && !p->rb_ev) {
        int least = 0, greatest = 0;
        FB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));      // numerically lower = more urgent
        for (int i = 0; i < RB_SLABS; ++i)
            FB_CUDA(cudaStreamCreateWithPriority(&p->rb_stream[i], cudaStreamNonBlocking, std::min(greatest + i, least)));
        FB_CUDA(cudaEventCreateWithFlags(&p->rb_ev, cudaEventDisableTiming));
    }
    bool side_used = false;
    // a finished table leaves for host memory on the side stream while the main stream carries on
    auto copy_out = [&](cudaStream_t after, void* host, const void* dev, size_t bytes) -> int {
        if (!host) return FB_OK;
        if (after != p->side) {
            FB_CUDA(cudaEventRecord(p->rb_ev, after));
            FB_CUDA(cudaStreamWaitEvent(p->side, p->rb_ev, 0));
        }
        FB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, p->side));
        side_used = true;
        return FB_OK;
    };
    const size_t bT = image_bytes(p->P, FB_IMAGE_TRANSMITTANCE), bE = image_bytes(p->P, FB_IMAGE_IRRADIANCE),
                 bS = image_bytes(p->P, FB_IMAGE_SCATTERING);
#define STAGE(stage, ord) if ((st = run_stage(p, c, stage, ord, 0, R, launches)) != FB_OK) return st
#define COPY_OUT(after, host, dev, bytes) if ((st = copy_out(after, host, dev, bytes)) != FB_OK) return st
    STAGE(FB_STAGE_TRANSMITTANCE, 0);          // :1726-1745
    COPY_OUT(s, p->rb_T, p->img.transmittance, bT);
    if (!overlap) {
        STAGE(FB_STAGE_DIRECT_IRRADIANCE, 0);      // :1760-1779  -> delta_irradiance
        STAGE(FB_STAGE_SINGLE_SCATTERING, 0);      // :1781-1800
        STAGE(FB_STAGE_CLEAR_IRRADIANCE, 0);       // :1802-1831  direct irradiance is not accumulated
    } else {
        // K2 and the clear touch the two irradiance images only, K3 reads the transmittance table only: the two small
        // launches run on the side stream next to K3 and join before the order loop
        FB_CUDA(cudaEventRecord(p->ev_fork, s));
        FB_CUDA(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
        LaunchCtx cs = c;
        cs.stream = p->side;
        if ((st = run_stage(p, cs, FB_STAGE_DIRECT_IRRADIANCE, 0, 0, R, launches)) != FB_OK) return st;
        if ((st = run_stage(p, cs, FB_STAGE_CLEAR_IRRADIANCE, 0, 0, R, launches)) != FB_OK) return st;
        FB_CUDA(cudaEventRecord(p->ev_join, p->side));
        STAGE(FB_STAGE_SINGLE_SCATTERING, 0);
        FB_CUDA(cudaStreamWaitEvent(s, p->ev_join, 0));
    }
    if (p->order < 2) {
        COPY_OUT(s, p->rb_S, p->img.scattering, bS);
        COPY_OUT(s, p->rb_E, p->img.irradiance, bE);
    }
    for (uint32_t order = 2; order <= p->order; ++order) {   // :1853
        const bool last = order == p->order;
        if (!overlap) {
            STAGE(FB_STAGE_SCATTERING_DENSITY, order);           // :1878-1904, push constant `order`
            if ((st = run_stage(p, c, FB_STAGE_INDIRECT_IRRADIANCE, order - 1, 0, ER, launches)) != FB_OK) return st;   // :1927-1953, push constant `order - 1`
            if (last) COPY_OUT(s, p->rb_E, p->img.irradiance, bE);
        } else {
            // K4 reads delta_irradiance (row 0) only in its preparation kernel, K5 overwrites that image and reads
            // nothing K4 writes: K5 runs on a side stream next to K4's main kernel and joins before K6 (which
            // overwrites the delta_multiple_scattering K5 reads).
            if (!fast::density_is_fast(c.P) || !c.img.scratch) p->slow_stages |= 1u << FB_STAGE_SCATTERING_DENSITY;
            cudaError_t e = fast::scattering_density(c, (int)order, 0, R, p->ev_fork);
            if (e != cudaSuccess) return cuda_fail(e, "scattering_density launch");
            if (launches) *launches += fast::launches_per_stage(c.P, FB_STAGE_SCATTERING_DENSITY, R);
            FB_CUDA(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
            side_used = true;
            LaunchCtx cs = c;
            cs.stream = p->side;
            if ((st = run_stage(p, cs, FB_STAGE_INDIRECT_IRRADIANCE, order - 1, 0, ER, launches)) != FB_OK) return st;
            FB_CUDA(cudaEventRecord(p->ev_join, p->side));
            FB_CUDA(cudaStreamWaitEvent(s, p->ev_join, 0));
            if (last) COPY_OUT(p->side, p->rb_E, p->img.irradiance, bE);   // after the join point: K6 does not wait for it
        }
        const int nsl = (last && p->rb_S && R % rb_slabs() == 0) ? rb_slabs() : 1;
        if (nsl == 1) {
            STAGE(FB_STAGE_MULTIPLE_SCATTERING, 0);              // :1979-1998
            if (last) COPY_OUT(s, p->rb_S, p->img.scattering, bS);
        } else {
            // K6 is r-local in what it writes (multiple_scattering.comp:77-92): the last pass runs as r-slabs on
            // streams of descending priority (the block scheduler drains them in that order, later slabs fill the
            // tail of earlier ones), each followed by the read-back of its slab of `scattering`.
            FB_CUDA(cudaEventRecord(p->rb_ev, s));
            const size_t slab_b = bS / nsl;
            for (int i = 0; i < nsl; ++i) {
                cudaStream_t q = p->rb_stream[i];
                FB_CUDA(cudaStreamWaitEvent(q, p->rb_ev, 0));
                LaunchCtx cq = c;
                cq.stream = q;
                if ((st = run_stage(p, cq, FB_STAGE_MULTIPLE_SCATTERING, 0, i * (R / nsl), (i + 1) * (R / nsl), launches)) != FB_OK)
                    return st;
            }
            // every slab's kernel is enqueued before the first copy: a copy into pageable memory blocks the calling thread
            for (int i = 0; i < nsl; ++i)
                FB_CUDA(cudaMemcpyAsync(static_cast<char*>(p->rb_S) + i * slab_b,
                                        reinterpret_cast<const char*>(p->img.scattering) + i * slab_b, slab_b,
                                        cudaMemcpyDeviceToHost, p->rb_stream[i]));
            for (int i = 0; i < nsl; ++i) {
                FB_CUDA(cudaEventRecord(p->ev_fork, p->rb_stream[i]));
                FB_CUDA(cudaStreamWaitEvent(s, p->ev_fork, 0));
            }
        }
    }
#undef STAGE
#undef COPY_OUT
    if (rb && side_used) {                     // the copies on the side stream are part of the command stream
        FB_CUDA(cudaEventRecord(p->ev_join, p->side));
        FB_CUDA(cudaStreamWaitEvent(s, p->ev_join, 0));
    }
    return FB_OK;
}

int fb_pending_set_readback(FbPending* p, void* host_transmittance, void* host_scattering, void* host_irradiance) {
    if (!p || !p->inner) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_set_readback: NULL / already taken");
    DeviceGuard g(p->device);
    p->rb_T = host_transmittance; p->rb_S = host_scattering; p->rb_E = host_irradiance;
    if (p->graph) { cudaGraphExecDestroy(p->graph); p->graph = nullptr; }   // re-recorded by the next resubmit
    return FB_OK;
}

int fb_atmosphere_build(FbBuilder* b, const FbParams* params, uint32_t order, void* stream, FbPending** out) {
    int st = fb_atmosphere_allocate(b, params, order, out);
    if (st != FB_OK) return st;
    DeviceGuard g(b->device);
    int launches = 0;
    st = enqueue_all(*out, (cudaStream_t)stream, &launches);
    (*out)->launches = launches;
    (*out)->done->note((cudaStream_t)stream);   // also after a failed enqueue: earlier stages are in flight on the blocks
    if (st != FB_OK) { fb_pending_destroy(*out); *out = nullptr; }
    return st;
}

int fb_pending_resubmit(FbPending* p, void* stream) {
    if (!p || !p->inner) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_resubmit: NULL / already taken");
    DeviceGuard g(p->device);
    if (!p->graph) {
        cudaStream_t cap;
        FB_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
        int launches = 0, st = FB_OK;
        if (e == cudaSuccess) st = enqueue_all(p, cap, &launches);
        cudaGraph_t graph = nullptr;
        cudaError_t e2 = (e == cudaSuccess) ? cudaStreamEndCapture(cap, &graph) : e;
        cudaStreamDestroy(cap);
        if (st != FB_OK) { if (graph) cudaGraphDestroy(graph); return st; }
        if (e2 != cudaSuccess) return cuda_fail(e2, "stream capture");
        e = cudaGraphInstantiate(&p->graph, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { p->graph = nullptr; return cuda_fail(e, "cudaGraphInstantiate"); }
        p->launches = launches;
    }
    ++p->inner->version;
    FB_CUDA(cudaGraphLaunch(p->graph, (cudaStream_t)stream));
    p->done->note((cudaStream_t)stream);
    return FB_OK;
}
int fb_pending_launch_count(const FbPending* p) { return p ? p->launches : 0; }

int fb_pending_run_stage(FbPending* p, int stage, uint32_t order, uint32_t r_begin, uint32_t r_end, void* stream) {
    if (!p) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_run_stage: NULL");
    // the slab counts altitude levels of the scattering table, or rows of the irradiance table for indirect_irradiance
    const uint32_t R = (uint32_t)(stage == FB_STAGE_INDIRECT_IRRADIANCE ? p->P.irradiance_r_size : p->P.scattering_r_size);
    if (r_end == 0) r_end = R;
    if (r_begin >= r_end || r_end > R) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_run_stage: bad r slab");
    DeviceGuard g(p->device);
    LaunchCtx c = make_ctx(p, (cudaStream_t)stream);
    int n = 0;
    if (p->inner) ++p->inner->version;
    const int st = run_stage(p, c, stage, order, (int)r_begin, (int)r_end, &n);
    p->launches += n;   // stage-driven pendings accumulate; build / resubmit overwrite with the per-submit count
    p->done->note((cudaStream_t)stream);
    return st;
}
uint32_t fb_pending_slow_stages(const FbPending* p) { return p ? p->slow_stages : 0; }
uint32_t fb_params_slow_stages(const FbParams* p) {
    if (!p || fb_params_validate(p) != FB_OK) return 0;     // nothing runs with such a block: fb_params_validate says why
    uint32_t m = 0;
    if (!fast::density_is_fast(*p)) m |= 1u << FB_STAGE_SCATTERING_DENSITY;
    if (!fast::multiple_is_fast(*p)) m |= 1u << FB_STAGE_MULTIPLE_SCATTERING;
    return m;
}

int fb_pending_image(FbPending* p, int image, void** dev_ptr, size_t* bytes) {
    if (!p || image < 0 || image >= FB_IMAGE_COUNT) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_image");
    if (dev_ptr) *dev_ptr = image_ptr(p, image);
    if (bytes) *bytes = image_bytes(p->P, image);
    return FB_OK;
}
int fb_pending_upload(FbPending* p, int image, const void* host, size_t bytes, void* stream) {
    if (!p || !host || image < 0 || image >= FB_IMAGE_COUNT) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_upload");
    if (bytes != image_bytes(p->P, image)) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_upload: size mismatch");
    DeviceGuard g(p->device);
    if (p->inner) ++p->inner->version;
    FB_CUDA(cudaMemcpyAsync(image_ptr(p, image), host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    p->done->note((cudaStream_t)stream);
    return FB_OK;
}
int fb_pending_download(FbPending* p, int image, void* host, size_t bytes, void* stream) {
    if (!p || !host || image < 0 || image >= FB_IMAGE_COUNT) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_download");
    if (bytes != image_bytes(p->P, image)) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_download: size mismatch");
    DeviceGuard g(p->device);
    FB_CUDA(cudaMemcpyAsync(host, image_ptr(p, image), bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    p->done->note((cudaStream_t)stream);
    return FB_OK;
}

int fb_pending_atmosphere(FbPending* p, const FbAtmosphere** out) {
    if (!p || !out || !p->inner) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_atmosphere");
    *out = p->inner;
    return FB_OK;
}
int fb_pending_wait(FbPending* p) {
    if (!p) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_wait: NULL");
    DeviceGuard g(p->device);
    const cudaError_t e = p->done->wait();
    return e == cudaSuccess ? FB_OK : cuda_fail(e, "waiting for the precompute");
}
int fb_pending_assert_ready(FbPending* p, int check, FbAtmosphere** out) {
    if (!p || !out || !p->inner) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_assert_ready");
    DeviceGuard g(p->device);
    if (check) {
        // Every submission left an event behind on its stream: a too-early call (undefined behaviour in the reference,
        // precompute.rs:2208-2211) is an error code here, found without synchronising anything.  `p` stays valid.
        const cudaError_t e = p->done->query();
        if (e == cudaErrorNotReady) {
            (void)cudaGetLastError();
            return fail(FB_ERR_NOT_READY, "fb_pending_assert_ready: the precompute has not finished on its stream(s)");
        }
        if (e != cudaSuccess) return cuda_fail(e, "querying the precompute's completion events");
    }
    // check = 0: the temporaries go back to the block cache carrying the completion events, so their next owner waits
    // for whatever is still in flight (stream-ordered reuse, fb_internal.h)
    *out = p->inner;
    p->inner = nullptr;
    fb_pending_destroy(p);
    return FB_OK;
}

int fb_atmosphere_transmittance(const FbAtmosphere* a, const void** ptr, FbExtent2D* e) {
    if (!a) return fail(FB_ERR_INVALID_ARGUMENT, "atmosphere: NULL");
    if (ptr) *ptr = a->transmittance;
    if (e) fb_params_transmittance_extent(&a->P, e);
    return FB_OK;
}
int fb_atmosphere_scattering(const FbAtmosphere* a, const void** ptr, FbExtent3D* e) {
    if (!a) return fail(FB_ERR_INVALID_ARGUMENT, "atmosphere: NULL");
    if (ptr) *ptr = a->scattering;
    if (e) fb_params_scattering_extent(&a->P, e);
    return FB_OK;
}
int fb_atmosphere_irradiance(const FbAtmosphere* a, const void** ptr, FbExtent2D* e) {
    if (!a) return fail(FB_ERR_INVALID_ARGUMENT, "atmosphere: NULL");
    if (ptr) *ptr = a->irradiance;
    if (e) fb_params_irradiance_extent(&a->P, e);
    return FB_OK;
}
int fb_atmosphere_params(const FbAtmosphere* a, FbParams* out) {
    if (!a || !out) return fail(FB_ERR_INVALID_ARGUMENT, "atmosphere: NULL");
    *out = a->P;
    return FB_OK;
}
static int read_back(const FbAtmosphere* a, const void* src, size_t want, void* host, size_t bytes, void* stream) {
    if (!a || !host) return fail(FB_ERR_INVALID_ARGUMENT, "read: NULL");
    if (bytes != want) return fail(FB_ERR_INVALID_ARGUMENT, "read: size mismatch");
    DeviceGuard g(a->device);
    FB_CUDA(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    a->done->note((cudaStream_t)stream);
    return FB_OK;
}
int fb_atmosphere_read_transmittance(const FbAtmosphere* a, void* host, size_t bytes, void* stream) {
    return a ? read_back(a, a->transmittance, image_bytes(a->P, FB_IMAGE_TRANSMITTANCE), host, bytes, stream)
             : fail(FB_ERR_INVALID_ARGUMENT, "atmosphere: NULL");
}
int fb_atmosphere_read_scattering(const FbAtmosphere* a, void* host, size_t bytes, void* stream) {
    return a ? read_back(a, a->scattering, image_bytes(a->P, FB_IMAGE_SCATTERING), host, bytes, stream)
             : fail(FB_ERR_INVALID_ARGUMENT, "atmosphere: NULL");
}
int fb_atmosphere_read_irradiance(const FbAtmosphere* a, void* host, size_t bytes, void* stream) {
    return a ? read_back(a, a->irradiance, image_bytes(a->P, FB_IMAGE_IRRADIANCE), host, bytes, stream)
             : fail(FB_ERR_INVALID_ARGUMENT, "atmosphere: NULL");
}

int fb_precompute_host(FbBuilder* b, const FbParams* p, uint32_t order, void* T, void* S, void* E) {
    if (!b) return fail(FB_ERR_INVALID_ARGUMENT, "fb_precompute_host: NULL builder");
    DeviceGuard g(b->device);
    cudaStream_t s;
    FB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    FbPending* pend = nullptr;
    int st = fb_atmosphere_allocate(b, p, order, &pend);
    if (st == FB_OK) {
        // the read-backs ride in the command stream, each table leaving as soon as its last writer has run
        pend->rb_T = T; pend->rb_S = S; pend->rb_E = E;
        int launches = 0;
        st = enqueue_all(pend, s, &launches);
        pend->done->note(s);
        cudaError_t e = cudaStreamSynchronize(s);
        if (st == FB_OK && e != cudaSuccess) st = cuda_fail(e, "cudaStreamSynchronize");
        fb_pending_destroy(pend);
    }
    cudaStreamDestroy(s);
    return st;
}

int fb_atmosphere_build_batch(FbBuilder* b, const FbParams* params, uint32_t n, uint32_t order, void* stream, FbPending** out) {
    if (!b || !params || !out) return fail(FB_ERR_INVALID_ARGUMENT, "fb_atmosphere_build_batch: NULL");
    DeviceGuard g(b->device);
    for (uint32_t i = 0; i < n; ++i) out[i] = nullptr;
    int st = FB_OK;
    for (uint32_t i = 0; i < n && st == FB_OK; ++i) st = fb_atmosphere_allocate(b, params + i, order, out + i);
    // independent atmospheres overlap on side streams that fork from and join into `stream`
    const uint32_t lanes = n < 8 ? n : 8;
    std::vector<cudaStream_t> side(lanes, nullptr);
    cudaEvent_t fork = nullptr;
    cudaError_t e = cudaSuccess;
    if (st == FB_OK && lanes) {
        e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(fork, (cudaStream_t)stream);
        for (uint32_t l = 0; l < lanes && e == cudaSuccess; ++l) {
            e = cudaStreamCreateWithFlags(&side[l], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(side[l], fork, 0);
        }
        for (uint32_t i = 0; i < n && e == cudaSuccess && st == FB_OK; ++i) {
            int launches = 0;
            st = enqueue_all(out[i], side[i % lanes], &launches);
            out[i]->launches = launches;
            out[i]->done->note(side[i % lanes]);   // also after a failed enqueue: its earlier stages are in flight
        }
        for (uint32_t l = 0; l < lanes && e == cudaSuccess; ++l) {
            cudaEvent_t join;
            e = cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(join, side[l]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent((cudaStream_t)stream, join, 0);
            cudaEventDestroy(join);
        }
    }
    for (uint32_t l = 0; l < lanes; ++l) if (side[l]) cudaStreamDestroy(side[l]);   // deferred until the work drains
    if (fork) cudaEventDestroy(fork);
    if (e != cudaSuccess && st == FB_OK) st = cuda_fail(e, "batch stream fork/join");
    if (st != FB_OK)
        for (uint32_t i = 0; i < n; ++i) { fb_pending_destroy(out[i]); out[i] = nullptr; }
    return st;
}

// ---------------------------------------------------------------------------------------------
// renderer
// ---------------------------------------------------------------------------------------------
int fb_renderer_create(FbBuilder* b, FbRenderer** out) {
    if (!b || !out) return fail(FB_ERR_INVALID_ARGUMENT, "fb_renderer_create: NULL");
    FbRenderer* r = new (std::nothrow) FbRenderer();
    if (!r) return fail(FB_ERR_OUT_OF_MEMORY, "host allocation");
    r->device = b->device;
    r->kernels = b->kernels;
    r->sweep_draws = nullptr;
    r->sweep_capacity = 0;
    r->view_tables = nullptr; r->view_tables_bytes = 0;
    r->expanded = nullptr; r->expanded_bytes = 0; r->expanded_serial = 0; r->expanded_version = 0; r->expanded_ready = nullptr;
    r->no_expand = std::getenv("FUZZYBLUE_B200_RENDER_FP16_TABLE") != nullptr;
    r->draws = std::make_shared<Completion>();
    *out = r;
    return FB_OK;
}
void fb_renderer_destroy(FbRenderer* r) {
    if (!r) return;
    DeviceGuard g(r->device);
    cudaFree(r->sweep_draws);
    cudaFree(r->view_tables);
    cudaFree(r->expanded);
    if (r->expanded_ready) cudaEventDestroy(r->expanded_ready);
    delete r;
}

static int draw_common(FbRenderer* r, const FbAtmosphere* a, const FbDrawParams* d, uint32_t views, const float* depth,
                       float* color, float* transm, float* blend, uint32_t w, uint32_t h, void* stream) {
    if (!r || !a || !d || !depth) return fail(FB_ERR_INVALID_ARGUMENT, "draw: NULL argument");
    if (r->device != a->device) return fail(FB_ERR_INVALID_ARGUMENT, "draw: renderer and atmosphere live on different devices");
    if (h > 65535 || views > 65535) return fail(FB_ERR_INVALID_ARGUMENT, "draw: height / views exceed the launch grid");
    DeviceGuard g(r->device);
    if (views > 1 && views > r->sweep_capacity) {
        cudaFree(r->sweep_draws);
        r->sweep_draws = nullptr;
        r->sweep_capacity = 0;
        FB_CUDA(cudaMalloc(&r->sweep_draws, (size_t)views * render_view_record_bytes()));
        r->sweep_capacity = views;
    }
    // FAST path: keep a (value, delta) fp32 expansion of this atmosphere's scattering table (bit-identical look-ups with
    // half the instructions per tap).  Tables another API can write behind our back (exportable blocks) and tables whose
    // expansion would not stay cache-resident anyway use the fp16 table directly.
    const void* expanded = nullptr;
    const size_t xb = render_expanded_bytes(a->P);
    const uint64_t texels = xb / 32;                        // (value, delta) float4 pairs; the trailing constants slot rounds away
    if (r->kernels != FB_KERNELS_REFERENCE && !r->no_expand && !a->vmm && xb <= EXPANDED_MAX_BYTES && texels < (1ull << 31) &&
        (uint64_t)w * h * views >= texels) {                // tiny draws: the expansion would cost more than it saves
        if (r->expanded_bytes < xb) {
            cudaFree(r->expanded);
            r->expanded = nullptr; r->expanded_bytes = 0; r->expanded_serial = 0;
            FB_CUDA(cudaMalloc(&r->expanded, xb));
            r->expanded_bytes = xb;
        }
        if (!r->expanded_ready) FB_CUDA(cudaEventCreateWithFlags(&r->expanded_ready, cudaEventDisableTiming));
        if (r->expanded_serial != a->serial || r->expanded_version != a->version) {
            // earlier draws (any stream) may still read the old contents: the rebuild waits, on the device, for the
            // completion event each of those streams carries -- no host or device-wide synchronisation inside a draw
            if (r->expanded_serial) FB_CUDA(r->draws->stream_wait((cudaStream_t)stream));
            cudaError_t ee = render_expand_scattering(a->P, a->transmittance, a->scattering, r->expanded, (cudaStream_t)stream);
            if (ee != cudaSuccess) return cuda_fail(ee, "render_expand_scattering launch");
            FB_CUDA(cudaEventRecord(r->expanded_ready, (cudaStream_t)stream));
            r->expanded_serial = a->serial; r->expanded_version = a->version;
        } else {
            FB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, r->expanded_ready, 0));
        }
        expanded = r->expanded;
    }
    // Per-view sky tables: scratch the draw's pre-pass rewrites.  A draw on another stream may still be reading the previous
    // contents (and a sweep's view records): wait for it on the device, as a table rebuild does.
    void* vtabs = nullptr;
    if (r->kernels != FB_KERNELS_REFERENCE) {
        const size_t need = (size_t)views * render_view_table_bytes(a->P);
        if (r->view_tables_bytes < need) {
            cudaFree(r->view_tables);                       // synchronises the device: no earlier draw still reads it
            r->view_tables = nullptr; r->view_tables_bytes = 0;
            FB_CUDA(cudaMalloc(&r->view_tables, need));
            r->view_tables_bytes = need;
        }
        FB_CUDA(r->draws->stream_wait((cudaStream_t)stream));
        vtabs = r->view_tables;
    }
    cudaError_t e = render_sky(a->P, a->transmittance, a->scattering, expanded, d, views > 1 ? r->sweep_draws : nullptr, vtabs, views, depth,
                               (float4*)color, (float4*)transm, (float4*)blend, w, h, r->kernels, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "render_sky launch");
    r->draws->note((cudaStream_t)stream);
    a->done->note((cudaStream_t)stream);   // the tables must outlive this draw: their block's next owner waits for it
    return FB_OK;
}

int fb_renderer_draw(FbRenderer* r, const FbAtmosphere* a, const FbDrawParams* d, const float* depth, float* color,
                     float* transm, uint32_t w, uint32_t h, void* stream) {
    return draw_common(r, a, d, 1, depth, color, transm, nullptr, w, h, stream);
}
int fb_renderer_draw_blend(FbRenderer* r, const FbAtmosphere* a, const FbDrawParams* d, const float* depth, float* fbuf,
                           uint32_t w, uint32_t h, void* stream) {
    if (!fbuf) return fail(FB_ERR_INVALID_ARGUMENT, "draw_blend: NULL framebuffer");
    return draw_common(r, a, d, 1, depth, nullptr, nullptr, fbuf, w, h, stream);
}
int fb_renderer_draw_sweep(FbRenderer* r, const FbAtmosphere* a, const FbDrawParams* d, uint32_t views, const float* depth,
                           float* color, float* transm, uint32_t w, uint32_t h, void* stream) {
    if (views == 0) return FB_OK;
    return draw_common(r, a, d, views, depth, color, transm, nullptr, w, h, stream);
}
int fb_renderer_draw_host(FbRenderer* r, const FbAtmosphere* a, const FbDrawParams* d, const float* depth_host, float* color_host,
                          float* transm_host, uint32_t w, uint32_t h) {
    if (!r || !a || !d || !depth_host) return fail(FB_ERR_INVALID_ARGUMENT, "draw_host: NULL argument");
    DeviceGuard g(r->device);
    const size_t npx = (size_t)w * h;
    float *depth = nullptr, *color = nullptr, *transm = nullptr;
    cudaStream_t s = nullptr;
    int st = FB_OK;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&depth, npx * sizeof(float));
    if (e == cudaSuccess && color_host) e = cudaMalloc((void**)&color, npx * sizeof(float4));
    if (e == cudaSuccess && transm_host) e = cudaMalloc((void**)&transm, npx * sizeof(float4));
    if (e == cudaSuccess) e = cudaMemcpyAsync(depth, depth_host, npx * sizeof(float), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) st = fb_renderer_draw(r, a, d, depth, color, transm, w, h, s);
    if (e == cudaSuccess && st == FB_OK && color_host) e = cudaMemcpyAsync(color_host, color, npx * sizeof(float4), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && st == FB_OK && transm_host) e = cudaMemcpyAsync(transm_host, transm, npx * sizeof(float4), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(depth); cudaFree(color); cudaFree(transm);
    if (s) cudaStreamDestroy(s);
    if (e != cudaSuccess) return cuda_fail(e, "draw_host");
    return st;
}

int fb_sky_radiance(const FbAtmosphere* a, const float* camera, const float* view_ray, const float* sun_direction, uint64_t n,
                    float* radiance_out, float* transmittance_out, void* stream) {
    if (!a || !camera || !view_ray || !sun_direction || !radiance_out || !transmittance_out)
        return fail(FB_ERR_INVALID_ARGUMENT, "fb_sky_radiance: NULL argument");
    DeviceGuard g(a->device);
    cudaError_t e = sky_radiance(a->P, a->transmittance, a->scattering, camera, view_ray, sun_direction, n, radiance_out,
                                 transmittance_out, (cudaStream_t)stream);
    a->done->note((cudaStream_t)stream);
    return e == cudaSuccess ? FB_OK : cuda_fail(e, "sky_radiance launch");
}
int fb_sun_and_sky_irradiance(const FbAtmosphere* a, const float* point, const float* normal, const float* sun_direction,
                              uint64_t n, float* sun_out, float* sky_out, void* stream) {
    if (!a || !point || !normal || !sun_direction || !sun_out || !sky_out)
        return fail(FB_ERR_INVALID_ARGUMENT, "fb_sun_and_sky_irradiance: NULL argument");
    DeviceGuard g(a->device);
    cudaError_t e = sun_sky_irradiance(a->P, a->transmittance, a->irradiance, point, normal, sun_direction, n, sun_out, sky_out,
                                       (cudaStream_t)stream);
    a->done->note((cudaStream_t)stream);
    return e == cudaSuccess ? FB_OK : cuda_fail(e, "sun_sky_irradiance launch");
}
