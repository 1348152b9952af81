/* fuzzyblue.h — C ABI of the B200-native atmosphere LUT precompute + sky evaluation.
 *
 * This is the drop-in boundary for the one hot path of Ralith/fuzzyblue: what a Rust shim
 * crate (rust/src/lib.rs), the C++ mirror (include/fuzzyblue.hpp) and the
 * Python ctypes mirror (fuzzyblue_b200/api.py) bind.  Plain pointers and sizes only; no torch,
 * no Vulkan types.  Every entry point names the reference interface it replaces
 * (paths relative to the reference checkout, file:line).
 *
 * Execution model: the reference only *records* into a caller-owned VkCommandBuffer
 * (src/precompute.rs:1077-1081, :2108-2110); here every `stream` argument is a caller-owned
 * cudaStream_t (passed as void*; NULL = the legacy default stream) and the call enqueues work on
 * it and returns without waiting.  Completion is the caller's business, exactly as
 * vkQueueSubmit + fence is in the reference (tests/smoke.rs:147-155).
 *
 * Errors: the reference panics through .unwrap(); this ABI returns an FbStatus and never
 * aborts.  fb_last_error() gives a thread-local message for the last non-OK status.
 *
 * LUT memory (read-back layout of examples/dump.rs:175-193, x fastest, tightly packed):
 *   transmittance  [T_r][T_mu][4]                 float32  (R32G32B32A32_SFLOAT, precompute.rs:1170)
 *   irradiance     [E_r][E_mu_s][4]               float32  (precompute.rs:1191)
 *   scattering     [S_r][S_mu][S_nu*S_mu_s][4]    float16  (R16G16B16A16_SFLOAT, precompute.rs:1215)
 */
#ifndef FUZZYBLUE_H_
#define FUZZYBLUE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FB_API __attribute__((visibility("default")))
#else
#define FB_API
#endif

typedef enum FbStatus {
    FB_OK = 0,
    FB_ERR_INVALID_ARGUMENT = 1, /* NULL handle, bad dims (nu_size < 2, odd mu_size, ...), bad slab */
    FB_ERR_CUDA = 2,             /* a CUDA runtime call failed; see fb_last_error() */
    FB_ERR_OUT_OF_MEMORY = 3,
    FB_ERR_NO_DEVICE = 4,        /* no CUDA device / not an sm_100 part: there is NO CPU fallback */
    FB_ERR_NOT_READY = 5         /* fb_pending_assert_ready() while the stream is still running */
} FbStatus;

/* ---- the atmosphere parameter block: shaders/params.h:9-87, host mirror ParamsRaw
 *      src/precompute.rs:937-1033.  Byte-identical std140 layout, 320 bytes. ------------------- */
typedef struct FbDensityProfileLayer { /* params.h:9-15; 32-byte stride (precompute.rs:1012-1021) */
    float width, exp_term, exp_scale, linear_term, constant_term;
    float _pad[3];
} FbDensityProfileLayer;

typedef struct FbDensityProfile { /* params.h:21-23 */
    FbDensityProfileLayer layers[2];
} FbDensityProfile;

typedef struct FbParams {
    float solar_irradiance[3];      /*   0 */
    float sun_angular_radius;       /*  12 */
    float rayleigh_scattering[3];   /*  16 */
    float bottom_radius;            /*  28  km */
    float mie_scattering[3];        /*  32 */
    float top_radius;               /*  44  km */
    float mie_extinction[3];        /*  48 */
    float mie_phase_function_g;     /*  60 */
    float ground_albedo[3];         /*  64 */
    float mu_s_min;                 /*  76 */
    float absorption_extinction[3]; /*  80  (host field is spelled `absorbtion_extinction`) */
    int32_t transmittance_mu_size;  /*  92 */
    int32_t transmittance_r_size;   /*  96 */
    int32_t scattering_r_size;      /* 100 */
    int32_t scattering_mu_size;     /* 104 */
    int32_t scattering_mu_s_size;   /* 108 */
    int32_t scattering_nu_size;     /* 112 */
    int32_t irradiance_mu_s_size;   /* 116 */
    int32_t irradiance_r_size;      /* 120 */
    int32_t _pad;                   /* 124 */
    FbDensityProfile rayleigh_density;   /* 128 */
    FbDensityProfile mie_density;        /* 192 */
    FbDensityProfile absorption_density; /* 256 */
} FbParams;                              /* 320 */

/* DrawParamsRaw, src/render.rs:254-271 == push constants of shaders/render_sky.frag:16-20. 92 bytes. */
typedef struct FbDrawParams {
    float inverse_viewproj[4][4]; /* [column][row]: (projection * view)^-1, world units metres */
    float camera_position[3];     /* km, planet frame */
    uint32_t _pad;
    float sun_direction[3];
} FbDrawParams;

typedef struct FbExtent2D { uint32_t width, height; } FbExtent2D;
typedef struct FbExtent3D { uint32_t width, height, depth; } FbExtent3D;

typedef struct FbBuilder FbBuilder;       /* Builder            src/precompute.rs:32-58  */
typedef struct FbPending FbPending;       /* PendingAtmosphere  src/precompute.rs:2111-2120 */
typedef struct FbAtmosphere FbAtmosphere; /* Atmosphere         src/precompute.rs:1036-1043 */
typedef struct FbRenderer FbRenderer;     /* Renderer           src/render.rs:13-19 */

/* Images an in-flight precompute owns (descriptor map src/precompute.rs:1259-1601). */
typedef enum FbImage {
    FB_IMAGE_TRANSMITTANCE = 0,             /* f32x4 2-D, kept   */
    FB_IMAGE_IRRADIANCE = 1,                /* f32x4 2-D, kept   */
    FB_IMAGE_SCATTERING = 2,                /* f16x4 3-D, kept   */
    FB_IMAGE_DELTA_IRRADIANCE = 3,          /* f32x4 2-D, temp   */
    FB_IMAGE_DELTA_RAYLEIGH = 4,            /* f16x4 3-D, temp   */
    FB_IMAGE_DELTA_MIE = 5,                 /* f16x4 3-D, temp   */
    FB_IMAGE_SCATTERING_DENSITY = 6,        /* f16x4 3-D, temp   */
    FB_IMAGE_DELTA_MULTIPLE_SCATTERING = 7, /* f16x4 3-D, temp   */
    FB_IMAGE_COUNT = 8
} FbImage;

/* The six compute shaders (shaders/<name>.comp) as individually launchable stages. */
typedef enum FbStage {
    FB_STAGE_TRANSMITTANCE = 0,       /* transmittance.comp:71-80 */
    FB_STAGE_DIRECT_IRRADIANCE = 1,   /* direct_irradiance.comp:36-46 */
    FB_STAGE_SINGLE_SCATTERING = 2,   /* single_scattering.comp:89-101 */
    FB_STAGE_SCATTERING_DENSITY = 3,  /* scattering_density.comp:144-155, push constant = order */
    FB_STAGE_INDIRECT_IRRADIANCE = 4, /* indirect_irradiance.comp:59-74, push constant = order */
    FB_STAGE_MULTIPLE_SCATTERING = 5, /* multiple_scattering.comp:81-93 */
    FB_STAGE_CLEAR_IRRADIANCE = 6     /* vkCmdClearColorImage, precompute.rs:1802-1831 */
} FbStage;

/* Kernel families.  Both follow the same shaders and produce the same tables within the parity
 * tolerance; FAST is the product, REFERENCE the one-thread-per-texel, contraction-free
 * transcription kept as the on-device cross-check. */
typedef enum FbKernels { FB_KERNELS_FAST = 0, FB_KERNELS_REFERENCE = 1 } FbKernels;

FB_API const char* fb_status_string(int status);
FB_API const char* fb_last_error(void);
FB_API const char* fb_version(void);

/* Parameters::default(), src/precompute.rs:849-935 (Earth).  `order` is not part of the uniform
 * block in the reference either (Parameters.order, :700); its default is 4. */
FB_API int fb_params_default(FbParams* out);
FB_API uint32_t fb_params_default_order(void);
/* Parameters::{transmittance,irradiance,scattering}_extent(), src/precompute.rs:771-793 */
FB_API int fb_params_transmittance_extent(const FbParams* p, FbExtent2D* out);
FB_API int fb_params_irradiance_extent(const FbParams* p, FbExtent2D* out);
FB_API int fb_params_scattering_extent(const FbParams* p, FbExtent3D* out);
/* FB_OK when the block is usable: every size >= 2, scattering_mu_size even, products fit int32, every float finite,
 * 0 < bottom_radius < top_radius, mu_s_min in [-1, 0], |g| < 1 (the reference checks nothing: a bad block there is
 * undefined behaviour in the shaders). */
FB_API int fb_params_validate(const FbParams* p);
/* Bit s (FbStage) is set when stage s of these dims is NOT covered by the restructured kernels and runs the
 * one-thread-per-texel transcription instead (~30x slower): scattering_density for scattering_nu_size > 128 or
 * irradiance_mu_s_size > 512, multiple_scattering for rows of more than 8192 texels.  0 for every config of
 * BASELINE.json, and 0 for a block fb_params_validate rejects (nothing runs with it).  fb_pending_slow_stages reports
 * what a pending actually ran. */
FB_API uint32_t fb_params_slow_stages(const FbParams* p);

/* Builder::new, src/precompute.rs:61-68: one-time, device-wide setup.  `device` is a CUDA ordinal. */
FB_API int fb_builder_create(int device, FbBuilder** out);
FB_API void fb_builder_destroy(FbBuilder* b);
FB_API int fb_builder_set_kernels(FbBuilder* b, int kernels /* FbKernels */);
/* Device memory released by finished precomputes is cached (up to 8 GiB) for the next build of the same dims;
 * fb_builder_trim returns the cached blocks to the driver.  Everything is released with the last object built from
 * the builder. */
FB_API int fb_builder_trim(FbBuilder* b);
FB_API int fb_builder_device(const FbBuilder* b);
FB_API int fb_builder_sm_count(const FbBuilder* b);
/* Measured issue-rate ceilings of this device (dense FP32 FMA TFLOP/s, SFU Gop/s): the roofline
 * denominators of this transcendental-heavy path (no reference analogue). Takes ~50 ms. */
FB_API int fb_builder_measure_peaks(FbBuilder* b, double* fp32_fma_tflops, double* sfu_gops);

/* Atmosphere::build, src/precompute.rs:1077-2073: allocate the 3 kept + 5 temporary images and
 * enqueue the whole command stream of :1671-2048 on `stream`. */
FB_API int fb_atmosphere_build(FbBuilder* b, const FbParams* p, uint32_t order, void* stream, FbPending** out);
/* Same allocation, nothing enqueued: the caller drives stages itself (tests, r-slab sharding). */
FB_API int fb_atmosphere_allocate(FbBuilder* b, const FbParams* p, uint32_t order, FbPending** out);
/* Re-enqueue the identical command stream on the images `p` already owns — the analogue of
 * re-submitting the pre-recorded command buffer, which is what benches/precompute.rs:138-148
 * times.  The stream is replayed from a CUDA graph instantiated on first use. */
FB_API int fb_pending_resubmit(FbPending* p, void* stream);
/* Make the read-back part of the recorded command stream (what examples/dump.rs:175-193 does with
 * vkCmdCopyImageToBuffer after the precompute): every later fb_pending_resubmit also copies the finished tables,
 * tightly packed, to these host buffers (pinned memory for an asynchronous copy; NULL = leave that table on the
 * device).  Each table leaves as soon as its last writer has run: transmittance after the first kernel, irradiance
 * after the last indirect_irradiance pass, and the last multiple_scattering pass runs in r-slabs whose slices of
 * `scattering` are copied while the remaining slabs compute.  The buffers must stay valid until the stream has
 * finished; results on the device are unchanged. */
FB_API int fb_pending_set_readback(FbPending* p, void* host_transmittance, void* host_scattering, void* host_irradiance);
/* Number of kernel launches / memset nodes of the last build / resubmit; for a pending driven stage by stage
 * (fb_atmosphere_allocate + fb_pending_run_stage) the running total since allocation. */
FB_API int fb_pending_launch_count(const FbPending* p);

/* One stage on the slab r in [r_begin, r_end) of the scattering r axis (3-D stages), on the ROWS [r_begin, r_end) of
 * the irradiance table (indirect_irradiance) or on the whole table (the other 2-D stages; slab ignored).  r_end = 0
 * means "to the end".  `order` is the push constant the reference passes for that stage (precompute.rs:1897, :1946);
 * ignored elsewhere. */
FB_API int fb_pending_run_stage(FbPending* p, int stage, uint32_t order, uint32_t r_begin, uint32_t r_end, void* stream);
/* Stages (bit s = FbStage s) this pending ran on the transcription kernels although the builder asked for FAST. */
FB_API uint32_t fb_pending_slow_stages(const FbPending* p);
/* Device pointer + byte size of one image (linear layout above), for collectives and interop. */
FB_API int fb_pending_image(FbPending* p, int image, void** dev_ptr, size_t* bytes);
/* Host <-> image copies (async on `stream`; `bytes` must equal the image size). */
FB_API int fb_pending_upload(FbPending* p, int image, const void* host, size_t bytes, void* stream);
FB_API int fb_pending_download(FbPending* p, int image, void* host, size_t bytes, void* stream);

/* PendingAtmosphere::atmosphere, :2203-2206 — borrowed, valid while `p` lives. */
FB_API int fb_pending_atmosphere(FbPending* p, const FbAtmosphere** out);
/* Blocks the calling thread until everything submitted through `p` (on any stream) has finished — the analogue of
 * waiting for the fence of the submission that carried the command buffer (tests/smoke.rs:147-155). */
FB_API int fb_pending_wait(FbPending* p);
/* PendingAtmosphere::assert_ready, :2208-2211 — consumes `p`, frees the five temporaries.  The caller asserts the
 * stream has finished.  check=1 verifies it with an event query (no synchronisation): FB_ERR_NOT_READY if work is
 * still in flight, and `p` stays valid.  check=0 trusts the caller as the reference does; the temporaries are then
 * recycled stream-ordered, so a too-early call cannot corrupt a later precompute. */
FB_API int fb_pending_assert_ready(FbPending* p, int check, FbAtmosphere** out);
/* Drop for PendingAtmosphere, :2122-2140 (also drops the atmosphere if it was never taken). */
FB_API void fb_pending_destroy(FbPending* p);

/* Atmosphere::{transmittance,scattering,irradiance}{,_extent}, :2075-2101.  Device pointers to the
 * linear tables described at the top of this header. */
FB_API int fb_atmosphere_transmittance(const FbAtmosphere* a, const void** dev_ptr, FbExtent2D* extent);
FB_API int fb_atmosphere_scattering(const FbAtmosphere* a, const void** dev_ptr, FbExtent3D* extent);
FB_API int fb_atmosphere_irradiance(const FbAtmosphere* a, const void** dev_ptr, FbExtent2D* extent);
FB_API int fb_atmosphere_params(const FbAtmosphere* a, FbParams* out);
/* examples/dump.rs:110-193: tightly packed read-back into host memory (async on `stream`). */
FB_API int fb_atmosphere_read_transmittance(const FbAtmosphere* a, void* host, size_t bytes, void* stream);
FB_API int fb_atmosphere_read_scattering(const FbAtmosphere* a, void* host, size_t bytes, void* stream);
FB_API int fb_atmosphere_read_irradiance(const FbAtmosphere* a, void* host, size_t bytes, void* stream);
/* Drop for Atmosphere, :1045-1073 */
FB_API void fb_atmosphere_destroy(FbAtmosphere* a);

/* ---- Vulkan interop (what keeps Atmosphere::{transmittance,scattering,irradiance}() -> vk::Image, :2075-2101, and the
 * caller-chosen final layout / stage / access of Parameters, :691-698, meaningful for a Vulkan caller) ---------------
 * After fb_builder_set_exportable(b, 1) the block an Atmosphere keeps (its three tables) is a CUDA virtual-memory
 * allocation with a POSIX-file-descriptor handle.  fb_atmosphere_export_fd returns a NEW descriptor per call (the caller
 * closes it, or hands it to vkAllocateMemory with VkImportMemoryFdInfoKHR { handleType = OPAQUE_FD }, which takes it
 * over) and where each table sits in the allocation; the Vulkan side binds a VkBuffer to the imported memory and
 * copies the tightly packed tables (layout at the top of this header) into its images with vkCmdCopyBufferToImage,
 * choosing layout, stage and access itself.  Exportable blocks bypass the builder's block cache. */
typedef struct FbExportLayout {
    size_t allocation_bytes;                     /* size to pass as VkMemoryAllocateInfo::allocationSize */
    size_t scattering_offset, scattering_bytes;  /* RGBA16F [r][mu][nu*mu_s] */
    size_t transmittance_offset, transmittance_bytes; /* RGBA32F [r][mu] */
    size_t irradiance_offset, irradiance_bytes;  /* RGBA32F [r][mu_s] */
} FbExportLayout;
FB_API int fb_builder_set_exportable(FbBuilder* b, int on);
FB_API int fb_atmosphere_export_fd(const FbAtmosphere* a, int* fd, FbExportLayout* layout);
/* What an importing process does, with CUDA as the importer: import `fd`, map it, copy `bytes` at `offset` to host
 * memory, unmap.  Does not close `fd`.  Used by the tests to prove the exported allocation carries the tables. */
FB_API int fb_external_memory_read_fd(int device, int fd, size_t allocation_bytes, size_t offset, void* host, size_t bytes);
/* The caller's VkSemaphore (exported with VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT; is_timeline = 1 for a
 * timeline semaphore) as a stream operation: signal it after the precompute / wait for it before a draw, in place of
 * the queue-family ownership transfer of PendingAtmosphere::acquire_ownership (:2147-2201).  On success the
 * descriptor belongs to CUDA.  NOT exercised in this image (no Vulkan loader or ICD): thin wrappers over
 * cudaImportExternalSemaphore / cudaSignalExternalSemaphoresAsync / cudaWaitExternalSemaphoresAsync. */
typedef struct FbExternalSemaphore FbExternalSemaphore;
FB_API int fb_external_semaphore_import_fd(int device, int fd, int is_timeline, FbExternalSemaphore** out);
FB_API int fb_external_semaphore_signal(FbExternalSemaphore* s, uint64_t value, void* stream);
FB_API int fb_external_semaphore_wait(FbExternalSemaphore* s, uint64_t value, void* stream);
FB_API void fb_external_semaphore_destroy(FbExternalSemaphore* s);

/* Host-buffer convenience = what a caller of the reference does end to end (build, submit, wait,
 * read back as examples/dump.rs does): params in host memory, three tables out to host memory.
 * Any output pointer may be NULL.  Synchronous. */
FB_API int fb_precompute_host(FbBuilder* b, const FbParams* p, uint32_t order, void* transmittance_f32,
                              void* scattering_f16, void* irradiance_f32);

/* Renderer::new, src/render.rs:34-40 (render pass / subpass / frame count have no CUDA meaning).
 * A renderer keeps device scratch of its own: the draw blocks of a sweep and, for draws of at least one pixel per
 * scattering texel, an fp32 (value, delta) expansion of the scattering table it last drew from (32 bytes per texel,
 * up to 256 MiB; bit-identical look-ups at half the instructions).  The expansion is re-derived when a draw names other
 * table contents (the rebuild waits on the device for earlier draws, no host synchronisation), so alternate between
 * atmospheres with one renderer each.
 * A draw also writes small per-view tables (the look-ups of sky pixels that depend on the camera alone, blended once
 * per view) into scratch the renderer owns; a draw therefore waits, on the device, for the renderer's earlier draws on
 * other streams: draws through ONE renderer execute in submission order.  Use one renderer per stream to overlap draws.
 * As with the reference's Renderer (descriptor sets per frame, render.rs:34-40), calls on one renderer are externally
 * synchronised by the caller. */
FB_API int fb_renderer_create(FbBuilder* b, FbRenderer** out);
FB_API void fb_renderer_destroy(FbRenderer* r);
/* Renderer::draw, src/render.rs:209-236 + shaders/render_sky.frag:24-35, one thread per pixel.
 *   depth        [h][w] float32 device pointer — the input attachment (render_sky.frag:22)
 *   color_out    [h][w][4] float32: (radiance.rgb, 0)        (location 0, index 0)
 *   transm_out   [h][w][4] float32: (transmittance.rgb, 1)   (location 0, index 1)
 * Either output may be NULL. */
FB_API int fb_renderer_draw(FbRenderer* r, const FbAtmosphere* a, const FbDrawParams* d, const float* depth,
                            float* color_out, float* transm_out, uint32_t width, uint32_t height, void* stream);
/* The fixed-function dual-source blend of src/render.rs:124-137 applied to a caller framebuffer:
 * dst.rgb = color.rgb + dst.rgb * transmittance.rgb, dst.a kept. */
FB_API int fb_renderer_draw_blend(FbRenderer* r, const FbAtmosphere* a, const FbDrawParams* d, const float* depth,
                                  float* framebuffer_rgba, uint32_t width, uint32_t height, void* stream);
/* A sweep of `views` draws over one depth buffer per view ([views][h][w]); outputs [views][h][w][4]. */
FB_API int fb_renderer_draw_sweep(FbRenderer* r, const FbAtmosphere* a, const FbDrawParams* d, uint32_t views,
                                  const float* depth, float* color_out, float* transm_out, uint32_t width,
                                  uint32_t height, void* stream);
/* Host-buffer variant (pageable or pinned host pointers; copies inside; synchronous). */
FB_API int fb_renderer_draw_host(FbRenderer* r, const FbAtmosphere* a, const FbDrawParams* d, const float* depth_host,
                                 float* color_host, float* transm_host, uint32_t width, uint32_t height);

/* Shader-library functions downstream engines include (no in-tree entry point in the reference):
 * GetSkyRadiance, shaders/render_sky.h:45-109, and GetSunAndSkyIrradiance,
 * shaders/render_lighting.h:10-28, evaluated for n independent queries.  All pointers are device
 * pointers to [n][3] float32 arrays. */
FB_API int fb_sky_radiance(const FbAtmosphere* a, const float* camera, const float* view_ray, const float* sun_direction,
                           uint64_t n, float* radiance_out, float* transmittance_out, void* stream);
FB_API int fb_sun_and_sky_irradiance(const FbAtmosphere* a, const float* point, const float* normal,
                                     const float* sun_direction, uint64_t n, float* sun_irradiance_out,
                                     float* sky_irradiance_out, void* stream);

/* Batch of independent atmospheres (BASELINE.json config 4): `n` parameter blocks, one precompute
 * each, round-robin over an internal pool of streams that all fork from / join into `stream`. */
FB_API int fb_atmosphere_build_batch(FbBuilder* b, const FbParams* params, uint32_t n, uint32_t order, void* stream,
                                     FbPending** out /* [n] */);

/* ---- One atmosphere built by several GPUs (BASELINE.json configs[2]): the scattering table slabbed along r ---------
 * The multi-device replacement for Atmosphere::build (src/precompute.rs:1077-1081).  Rank `rank` of `world` owns the
 * altitude levels [rank * R / world, (rank + 1) * R / world) of every 3-D image (R = scattering_r_size must divide by
 * `world`); a slab is one contiguous byte range of the linear layout.  Per scattering order the ranks exchange
 *   - an all-gather of scattering_density before multiple_scattering (its ray march crosses every r:
 *     multiple_scattering.comp:35-44), pipelined in sub-slabs behind the density kernels,
 *   - a one-slice halo of delta_multiple_scattering (delta_rayleigh / delta_mie once) with each neighbour: a density
 *     texel reads the previous order's table at its own r +- one slice (scattering_density.comp:72-75, scattering.h:17-22),
 *   - the irradiance rows, each computed by the rank whose slab brackets its altitude, as 1 KiB broadcasts.
 * The schedule is data: fb_sharded_plan writes the steps rank `rank` executes (no device needed); fb_pending_run_sharded
 * executes them with NCCL on `stream` + an internal communication stream.  Results are bit-identical to the
 * single-GPU build (tests/test_sharded_gpu.py). */
typedef enum FbShardOp {
    FB_SHARD_STAGE = 0,      /* fb_pending_run_stage(stage, order, begin, end)                                             */
    FB_SHARD_ALLGATHER = 1,  /* 3-D `image`: rows [begin, end) RELATIVE to each rank's slab, every rank to every rank     */
    FB_SHARD_HALO = 2,       /* 3-D `image`: my first slice -> rank - 1, my last slice -> rank + 1; theirs into my halo    */
    FB_SHARD_BCAST_ROWS = 3, /* 2-D `image`: rows [begin, end) from rank `root` to everyone                               */
    FB_SHARD_JOIN = 4        /* later stages wait for every exchange issued so far                                        */
} FbShardOp;
typedef struct FbShardStep {
    int32_t op;     /* FbShardOp */
    int32_t stage;  /* FbStage (FB_SHARD_STAGE) */
    int32_t image;  /* FbImage (exchanges) */
    uint32_t order; /* push constant of the stage */
    uint32_t begin, end;
    int32_t root;
    int32_t _pad;
} FbShardStep;
#define FB_SHARD_GATHER_RESULT 1u /* finish with an all-gather of `scattering`: every rank ends with the whole table */
#define FB_SHARD_NO_PIPELINE 2u   /* exchange whole slabs (one ncclAllGather) instead of sub-slabs behind the kernels */
#define FB_SHARD_PIPELINE_ALWAYS 4u /* sub-slab exchanges even for tables whose sub-slabs are below 16 MiB (tests) */
/* steps = NULL: only *count is written.  Transmittance and irradiance always end complete on every rank. */
FB_API int fb_sharded_plan(const FbParams* p, uint32_t order, int rank, int world, uint32_t flags, FbShardStep* steps,
                           uint32_t capacity, uint32_t* count);
/* `nccl_comm` is an ncclComm_t of `world` ranks in which this process is `rank` (NULL allowed when world == 1).
 * NCCL is loaded at run time (libnccl.so.2, or $FUZZYBLUE_B200_NCCL_LIB): no link-time dependency. */
FB_API int fb_atmosphere_build_sharded(FbBuilder* b, const FbParams* p, uint32_t order, void* nccl_comm, int rank, int world,
                                       uint32_t flags, void* stream, FbPending** out);
/* The same schedule again on the images `p` already owns (what a benchmark loop times). */
FB_API int fb_pending_run_sharded(FbPending* p, void* nccl_comm, int rank, int world, uint32_t flags, void* stream);
/* For callers without NCCL bindings of their own (ctypes, C): thin wrappers over ncclGetUniqueId / ncclCommInitRank /
 * ncclCommDestroy.  Rank 0 makes the id and ships its 128 bytes to the other ranks by any means. */
#define FB_NCCL_UNIQUE_ID_BYTES 128
FB_API int fb_nccl_version(int* version);
FB_API int fb_nccl_unique_id(void* id128);
FB_API int fb_nccl_comm_create(int device, int world, int rank, const void* id128, void** nccl_comm_out);
FB_API int fb_nccl_comm_destroy(void* nccl_comm);

#ifdef __cplusplus
}
#endif
#endif /* FUZZYBLUE_H_ */
