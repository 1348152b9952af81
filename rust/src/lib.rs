//! Drop-in shim: the six public names of Ralith/fuzzyblue (`src/lib.rs:8-12`) implemented over the C ABI of
//! `include/fuzzyblue.h`, with the reference's signatures (names, arity, argument order):
//!
//! | reference (ash 0.31)                                                                  | here                                   |
//! |---|---|
//! | `Builder::new(&Instance, Arc<Device>, PipelineCache, PhysicalDevice, u32, Option<u32>)` | same six arguments; the Vulkan handles are generic and ignored, the CUDA ordinal comes from `$FUZZYBLUE_B200_DEVICE` (default 0) or `Builder::with_cuda_device` |
//! | `Atmosphere::build(Arc<Builder>, vk::CommandBuffer, &Parameters)`                     | same; `vk::CommandBuffer` is [`vk::CommandBuffer`] = a CUDA stream handle (null = default stream) |
//! | `PendingAtmosphere::{acquire_ownership, atmosphere, assert_ready}`                     | same |
//! | `Atmosphere::{transmittance, scattering, irradiance}{,_extent}()`                      | device pointers to the linear tables instead of `vk::Image` (`*_view()` returns the same pointer) |
//! | `Renderer::new(&Builder, PipelineCache, RenderPass, subpass, frames)`                  | same five arguments |
//! | `Renderer::set_depth_buffer(&mut self, frame, &vk::DescriptorImageInfo)`               | same; [`vk::DescriptorImageInfo`] carries the device pointer of the `[h][w]` f32 depth buffer |
//! | `Renderer::draw(&self, cmd, &Atmosphere, frame, &DrawParameters)`                      | same four arguments; the two fragment outputs go to the targets set with `Renderer::set_targets` (the render pass attachments of the reference) |
//!
//! The four Vulkan-only fields of `Parameters` are kept (as plain integers) so struct-update call sites such as
//! `tests/smoke.rs:136-142` compile.
//!
//! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Rust toolchain).  `tests/test_rust_ffi.py` parses the `extern "C"` block
//! and the `#[repr(C)]` structs below and checks them against `include/fuzzyblue.h` (names, arity, argument types,
//! struct sizes); the ABI itself is exercised by the Python, C and C++ callers in the test-suite.
use std::os::raw::{c_char, c_int, c_void};
use std::sync::Arc;

/// Stand-ins for the `ash::vk` types the reference's signatures name.
pub mod vk {
    use std::os::raw::c_void;
    /// Where the reference records into a command buffer, work is enqueued on this CUDA stream (`cudaStream_t`).
    pub type CommandBuffer = *mut c_void;
    #[derive(Debug, Default, Copy, Clone)] pub struct PipelineCache;
    #[derive(Debug, Default, Copy, Clone)] pub struct PhysicalDevice;
    #[derive(Debug, Default, Copy, Clone)] pub struct RenderPass;
    /// `set_depth_buffer` takes the depth attachment: here the device pointer of a `[height][width]` f32 buffer.
    #[derive(Debug, Copy, Clone)] pub struct DescriptorImageInfo { pub depth: *const f32, pub width: u32, pub height: u32 }
    #[derive(Debug, Default, Copy, Clone, PartialEq, Eq)] pub struct Extent2D { pub width: u32, pub height: u32 }
    #[derive(Debug, Default, Copy, Clone, PartialEq, Eq)] pub struct Extent3D { pub width: u32, pub height: u32, pub depth: u32 }
}

#[allow(non_camel_case_types)]
pub mod ffi {
    use super::*;
    #[repr(C)] #[derive(Copy, Clone, Default)]
    pub struct FbDensityProfileLayer { pub width: f32, pub exp_term: f32, pub exp_scale: f32, pub linear_term: f32, pub constant_term: f32, pub _pad: [f32; 3] }
    #[repr(C)] #[derive(Copy, Clone, Default)]
    pub struct FbDensityProfile { pub layers: [FbDensityProfileLayer; 2] }
    /// Byte-identical to `ParamsRaw` (src/precompute.rs:937-1033): 320 bytes.
    #[repr(C)] #[derive(Copy, Clone, Default)]
    pub struct FbParams {
        pub solar_irradiance: [f32; 3], pub sun_angular_radius: f32,
        pub rayleigh_scattering: [f32; 3], pub bottom_radius: f32,
        pub mie_scattering: [f32; 3], pub top_radius: f32,
        pub mie_extinction: [f32; 3], pub mie_phase_function_g: f32,
        pub ground_albedo: [f32; 3], pub mu_s_min: f32,
        pub absorption_extinction: [f32; 3],
        pub transmittance_mu_size: i32, pub transmittance_r_size: i32,
        pub scattering_r_size: i32, pub scattering_mu_size: i32, pub scattering_mu_s_size: i32, pub scattering_nu_size: i32,
        pub irradiance_mu_s_size: i32, pub irradiance_r_size: i32, pub _pad: i32,
        pub rayleigh_density: FbDensityProfile, pub mie_density: FbDensityProfile, pub absorption_density: FbDensityProfile,
    }
    /// Byte-identical to `DrawParamsRaw` (src/render.rs:254-260): 92 bytes.
    #[repr(C)] #[derive(Copy, Clone)]
    pub struct FbDrawParams { pub inverse_viewproj: [[f32; 4]; 4], pub camera_position: [f32; 3], pub _pad: u32, pub sun_direction: [f32; 3] }
    #[repr(C)] #[derive(Copy, Clone, Default)] pub struct FbExtent2D { pub width: u32, pub height: u32 }
    #[repr(C)] #[derive(Copy, Clone, Default)] pub struct FbExtent3D { pub width: u32, pub height: u32, pub depth: u32 }
    #[repr(C)] #[derive(Copy, Clone, Default)]
    pub struct FbExportLayout { pub allocation_bytes: usize, pub scattering_offset: usize, pub scattering_bytes: usize,
                                pub transmittance_offset: usize, pub transmittance_bytes: usize,
                                pub irradiance_offset: usize, pub irradiance_bytes: usize }
    #[repr(C)] #[derive(Copy, Clone, Default)]
    pub struct FbShardStep { pub op: i32, pub stage: i32, pub image: i32, pub order: u32, pub begin: u32, pub end: u32, pub root: i32, pub _pad: i32 }
    pub enum FbBuilder {} pub enum FbPending {} pub enum FbAtmosphere {} pub enum FbRenderer {} pub enum FbExternalSemaphore {}
    // every FB_API function of include/fuzzyblue.h (generated by tools/gen_rust_ffi.py; checked by tests/test_rust_ffi.py)
    extern "C" {
        pub fn fb_status_string(status: c_int) -> *const c_char;
        pub fn fb_last_error() -> *const c_char;
        pub fn fb_version() -> *const c_char;
        pub fn fb_params_default(out: *mut FbParams) -> c_int;
        pub fn fb_params_default_order() -> u32;
        pub fn fb_params_transmittance_extent(p: *const FbParams, out: *mut FbExtent2D) -> c_int;
        pub fn fb_params_irradiance_extent(p: *const FbParams, out: *mut FbExtent2D) -> c_int;
        pub fn fb_params_scattering_extent(p: *const FbParams, out: *mut FbExtent3D) -> c_int;
        pub fn fb_params_validate(p: *const FbParams) -> c_int;
        pub fn fb_params_slow_stages(p: *const FbParams) -> u32;
        pub fn fb_builder_create(device: c_int, out: *mut *mut FbBuilder) -> c_int;
        pub fn fb_builder_destroy(b: *mut FbBuilder);
        pub fn fb_builder_set_kernels(b: *mut FbBuilder, kernels: c_int) -> c_int;
        pub fn fb_builder_trim(b: *mut FbBuilder) -> c_int;
        pub fn fb_builder_device(b: *const FbBuilder) -> c_int;
        pub fn fb_builder_sm_count(b: *const FbBuilder) -> c_int;
        pub fn fb_builder_measure_peaks(b: *mut FbBuilder, fp32_fma_tflops: *mut f64, sfu_gops: *mut f64) -> c_int;
        pub fn fb_atmosphere_build(b: *mut FbBuilder, p: *const FbParams, order: u32, stream: *mut c_void, out: *mut *mut FbPending) -> c_int;
        pub fn fb_atmosphere_allocate(b: *mut FbBuilder, p: *const FbParams, order: u32, out: *mut *mut FbPending) -> c_int;
        pub fn fb_pending_resubmit(p: *mut FbPending, stream: *mut c_void) -> c_int;
        pub fn fb_pending_set_readback(p: *mut FbPending, host_transmittance: *mut c_void, host_scattering: *mut c_void, host_irradiance: *mut c_void) -> c_int;
        pub fn fb_pending_launch_count(p: *const FbPending) -> c_int;
        pub fn fb_pending_run_stage(p: *mut FbPending, stage: c_int, order: u32, r_begin: u32, r_end: u32, stream: *mut c_void) -> c_int;
        pub fn fb_pending_slow_stages(p: *const FbPending) -> u32;
        pub fn fb_pending_image(p: *mut FbPending, image: c_int, dev_ptr: *mut *mut c_void, bytes: *mut usize) -> c_int;
        pub fn fb_pending_upload(p: *mut FbPending, image: c_int, host: *const c_void, bytes: usize, stream: *mut c_void) -> c_int;
        pub fn fb_pending_download(p: *mut FbPending, image: c_int, host: *mut c_void, bytes: usize, stream: *mut c_void) -> c_int;
        pub fn fb_pending_atmosphere(p: *mut FbPending, out: *mut *const FbAtmosphere) -> c_int;
        pub fn fb_pending_wait(p: *mut FbPending) -> c_int;
        pub fn fb_pending_assert_ready(p: *mut FbPending, check: c_int, out: *mut *mut FbAtmosphere) -> c_int;
        pub fn fb_pending_destroy(p: *mut FbPending);
        pub fn fb_atmosphere_transmittance(a: *const FbAtmosphere, dev_ptr: *mut *const c_void, extent: *mut FbExtent2D) -> c_int;
        pub fn fb_atmosphere_scattering(a: *const FbAtmosphere, dev_ptr: *mut *const c_void, extent: *mut FbExtent3D) -> c_int;
        pub fn fb_atmosphere_irradiance(a: *const FbAtmosphere, dev_ptr: *mut *const c_void, extent: *mut FbExtent2D) -> c_int;
        pub fn fb_atmosphere_params(a: *const FbAtmosphere, out: *mut FbParams) -> c_int;
        pub fn fb_atmosphere_read_transmittance(a: *const FbAtmosphere, host: *mut c_void, bytes: usize, stream: *mut c_void) -> c_int;
        pub fn fb_atmosphere_read_scattering(a: *const FbAtmosphere, host: *mut c_void, bytes: usize, stream: *mut c_void) -> c_int;
        pub fn fb_atmosphere_read_irradiance(a: *const FbAtmosphere, host: *mut c_void, bytes: usize, stream: *mut c_void) -> c_int;
        pub fn fb_atmosphere_destroy(a: *mut FbAtmosphere);
        pub fn fb_builder_set_exportable(b: *mut FbBuilder, on: c_int) -> c_int;
        pub fn fb_atmosphere_export_fd(a: *const FbAtmosphere, fd: *mut c_int, layout: *mut FbExportLayout) -> c_int;
        pub fn fb_external_memory_read_fd(device: c_int, fd: c_int, allocation_bytes: usize, offset: usize, host: *mut c_void, bytes: usize) -> c_int;
        pub fn fb_external_semaphore_import_fd(device: c_int, fd: c_int, is_timeline: c_int, out: *mut *mut FbExternalSemaphore) -> c_int;
        pub fn fb_external_semaphore_signal(s: *mut FbExternalSemaphore, value: u64, stream: *mut c_void) -> c_int;
        pub fn fb_external_semaphore_wait(s: *mut FbExternalSemaphore, value: u64, stream: *mut c_void) -> c_int;
        pub fn fb_external_semaphore_destroy(s: *mut FbExternalSemaphore);
        pub fn fb_precompute_host(b: *mut FbBuilder, p: *const FbParams, order: u32, transmittance_f32: *mut c_void, scattering_f16: *mut c_void, irradiance_f32: *mut c_void) -> c_int;
        pub fn fb_renderer_create(b: *mut FbBuilder, out: *mut *mut FbRenderer) -> c_int;
        pub fn fb_renderer_destroy(r: *mut FbRenderer);
        pub fn fb_renderer_draw(r: *mut FbRenderer, a: *const FbAtmosphere, d: *const FbDrawParams, depth: *const f32, color_out: *mut f32, transm_out: *mut f32, width: u32, height: u32, stream: *mut c_void) -> c_int;
        pub fn fb_renderer_draw_blend(r: *mut FbRenderer, a: *const FbAtmosphere, d: *const FbDrawParams, depth: *const f32, framebuffer_rgba: *mut f32, width: u32, height: u32, stream: *mut c_void) -> c_int;
        pub fn fb_renderer_draw_sweep(r: *mut FbRenderer, a: *const FbAtmosphere, d: *const FbDrawParams, views: u32, depth: *const f32, color_out: *mut f32, transm_out: *mut f32, width: u32, height: u32, stream: *mut c_void) -> c_int;
        pub fn fb_renderer_draw_host(r: *mut FbRenderer, a: *const FbAtmosphere, d: *const FbDrawParams, depth_host: *const f32, color_host: *mut f32, transm_host: *mut f32, width: u32, height: u32) -> c_int;
        pub fn fb_sky_radiance(a: *const FbAtmosphere, camera: *const f32, view_ray: *const f32, sun_direction: *const f32, n: u64, radiance_out: *mut f32, transmittance_out: *mut f32, stream: *mut c_void) -> c_int;
        pub fn fb_sun_and_sky_irradiance(a: *const FbAtmosphere, point: *const f32, normal: *const f32, sun_direction: *const f32, n: u64, sun_irradiance_out: *mut f32, sky_irradiance_out: *mut f32, stream: *mut c_void) -> c_int;
        pub fn fb_atmosphere_build_batch(b: *mut FbBuilder, params: *const FbParams, n: u32, order: u32, stream: *mut c_void, out: *mut *mut FbPending) -> c_int;
        pub fn fb_sharded_plan(p: *const FbParams, order: u32, rank: c_int, world: c_int, flags: u32, steps: *mut FbShardStep, capacity: u32, count: *mut u32) -> c_int;
        pub fn fb_atmosphere_build_sharded(b: *mut FbBuilder, p: *const FbParams, order: u32, nccl_comm: *mut c_void, rank: c_int, world: c_int, flags: u32, stream: *mut c_void, out: *mut *mut FbPending) -> c_int;
        pub fn fb_pending_run_sharded(p: *mut FbPending, nccl_comm: *mut c_void, rank: c_int, world: c_int, flags: u32, stream: *mut c_void) -> c_int;
        pub fn fb_nccl_version(version: *mut c_int) -> c_int;
        pub fn fb_nccl_unique_id(id128: *mut c_void) -> c_int;
        pub fn fb_nccl_comm_create(device: c_int, world: c_int, rank: c_int, id128: *const c_void, nccl_comm_out: *mut *mut c_void) -> c_int;
        pub fn fb_nccl_comm_destroy(nccl_comm: *mut c_void) -> c_int;
    }
}

fn check(status: c_int) {
    // the reference panics through .unwrap() on every Vulkan error (e.g. src/precompute.rs:83,98,529)
    if status != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::fb_last_error()) }.to_string_lossy().into_owned();
        panic!("fuzzyblue_b200: status {}: {}", status, msg);
    }
}

/// src/precompute.rs:660-666
#[derive(Debug, Copy, Clone, Default)]
pub struct DensityProfileLayer { pub width: f32, pub exp_term: f32, pub exp_scale: f32, pub linear_term: f32, pub constant_term: f32 }
/// src/precompute.rs:674-676
#[derive(Debug, Copy, Clone, Default)]
pub struct DensityProfile { pub layers: [DensityProfileLayer; 2] }

/// src/precompute.rs:690-769.  `usage`, `dst_stage_mask`, `dst_access_mask`, `layout` are accepted and ignored.
#[derive(Debug, Clone)]
pub struct Parameters {
    pub usage: u32, pub dst_stage_mask: u32, pub dst_access_mask: u32, pub layout: i32,
    pub order: u32,
    pub transmittance_mu_size: u32, pub transmittance_r_size: u32,
    pub scattering_r_size: u32, pub scattering_mu_size: u32, pub scattering_mu_s_size: u32, pub scattering_nu_size: u32,
    pub irradiance_mu_s_size: u32, pub irradiance_r_size: u32,
    pub solar_irradiance: [f32; 3], pub sun_angular_radius: f32, pub bottom_radius: f32, pub top_radius: f32,
    pub rayleigh_density: DensityProfile, pub rayleigh_scattering: [f32; 3],
    pub mie_density: DensityProfile, pub mie_scattering: [f32; 3], pub mie_extinction: [f32; 3], pub mie_phase_function_g: f32,
    pub absorbtion_density: DensityProfile, pub absorbtion_extinction: [f32; 3],
    pub ground_albedo: [f32; 3], pub mu_s_min: f32,
}

impl Default for Parameters {
    /// src/precompute.rs:849-935 (Earth)
    fn default() -> Self {
        let exp = |scale: f32| DensityProfile { layers: [DensityProfileLayer::default(),
            DensityProfileLayer { width: 0.0, exp_term: 1.0, exp_scale: scale, linear_term: 0.0, constant_term: 0.0 }] };
        Self {
            usage: 0, dst_stage_mask: 0x80, dst_access_mask: 0x20, layout: 5, order: 4,
            transmittance_mu_size: 256, transmittance_r_size: 64,
            scattering_r_size: 32, scattering_mu_size: 128, scattering_mu_s_size: 32, scattering_nu_size: 8,
            irradiance_mu_s_size: 64, irradiance_r_size: 16,
            solar_irradiance: [1.474, 1.850, 1.91198], sun_angular_radius: 0.004675, bottom_radius: 6360.0, top_radius: 6420.0,
            rayleigh_density: exp(-0.125), rayleigh_scattering: [0.005802, 0.013558, 0.033100],
            mie_density: exp(-0.833333), mie_scattering: [0.003996; 3], mie_extinction: [0.004440; 3], mie_phase_function_g: 0.8,
            absorbtion_density: DensityProfile { layers: [
                DensityProfileLayer { width: 25.0, exp_term: 0.0, exp_scale: 0.0, linear_term: 0.066667, constant_term: -0.666667 },
                DensityProfileLayer { width: 0.0, exp_term: 0.0, exp_scale: 0.0, linear_term: -0.066667, constant_term: 2.666667 }] },
            absorbtion_extinction: [6.5e-4, 1.881e-3, 8.5e-5], ground_albedo: [0.1; 3], mu_s_min: -0.207912,
        }
    }
}

impl Parameters {
    /// src/precompute.rs:771-793
    pub fn transmittance_extent(&self) -> vk::Extent2D { vk::Extent2D { width: self.transmittance_mu_size, height: self.transmittance_r_size } }
    pub fn irradiance_extent(&self) -> vk::Extent2D { vk::Extent2D { width: self.irradiance_mu_s_size, height: self.irradiance_r_size } }
    pub fn scattering_extent(&self) -> vk::Extent3D { vk::Extent3D { width: self.scattering_nu_size * self.scattering_mu_s_size, height: self.scattering_mu_size, depth: self.scattering_r_size } }
    /// ParamsRaw::new, src/precompute.rs:964-991
    fn raw(&self) -> ffi::FbParams {
        let prof = |p: &DensityProfile| { let mut o = ffi::FbDensityProfile::default(); for i in 0..2 { let l = &p.layers[i];
            o.layers[i] = ffi::FbDensityProfileLayer { width: l.width, exp_term: l.exp_term, exp_scale: l.exp_scale, linear_term: l.linear_term, constant_term: l.constant_term, _pad: [0.0; 3] }; } o };
        ffi::FbParams {
            solar_irradiance: self.solar_irradiance, sun_angular_radius: self.sun_angular_radius,
            rayleigh_scattering: self.rayleigh_scattering, bottom_radius: self.bottom_radius,
            mie_scattering: self.mie_scattering, top_radius: self.top_radius,
            mie_extinction: self.mie_extinction, mie_phase_function_g: self.mie_phase_function_g,
            ground_albedo: self.ground_albedo, mu_s_min: self.mu_s_min, absorption_extinction: self.absorbtion_extinction,
            transmittance_mu_size: self.transmittance_mu_size as i32, transmittance_r_size: self.transmittance_r_size as i32,
            scattering_r_size: self.scattering_r_size as i32, scattering_mu_size: self.scattering_mu_size as i32,
            scattering_mu_s_size: self.scattering_mu_s_size as i32, scattering_nu_size: self.scattering_nu_size as i32,
            irradiance_mu_s_size: self.irradiance_mu_s_size as i32, irradiance_r_size: self.irradiance_r_size as i32, _pad: 0,
            rayleigh_density: prof(&self.rayleigh_density), mie_density: prof(&self.mie_density), absorption_density: prof(&self.absorbtion_density),
        }
    }
}

/// `Builder::new` (src/precompute.rs:61-68).  The Vulkan instance / device / cache / physical device / queue families
/// have no CUDA meaning and are ignored; the CUDA device is `$FUZZYBLUE_B200_DEVICE` (default ordinal 0).
pub struct Builder { raw: *mut ffi::FbBuilder }
impl Builder {
    pub fn new<I, D>(_instance: &I, _device: Arc<D>, _cache: vk::PipelineCache, _physical: vk::PhysicalDevice,
                     _gfx_queue_family: u32, _compute_queue_family: Option<u32>) -> Self {
        let ordinal = std::env::var("FUZZYBLUE_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        Self::with_cuda_device(ordinal)
    }
    /// Not in the reference: pick the CUDA device explicitly.
    pub fn with_cuda_device(device: i32) -> Self { let mut raw = std::ptr::null_mut(); check(unsafe { ffi::fb_builder_create(device, &mut raw) }); Self { raw } }
}
impl Drop for Builder { fn drop(&mut self) { unsafe { ffi::fb_builder_destroy(self.raw) } } }

/// src/precompute.rs:1036-1043
pub struct Atmosphere { _builder: Arc<Builder>, raw: *const ffi::FbAtmosphere, owned: bool }
impl Atmosphere {
    /// `Atmosphere::build(builder, cmd, &params)` (src/precompute.rs:1077-1081): enqueue on the stream `cmd`, return at once.
    pub unsafe fn build(builder: Arc<Builder>, cmd: vk::CommandBuffer, params: &Parameters) -> PendingAtmosphere {
        let mut raw = std::ptr::null_mut();
        check(ffi::fb_atmosphere_build(builder.raw, &params.raw(), params.order, cmd, &mut raw));
        PendingAtmosphere { builder, raw }
    }
    /// Device pointers to the linear tables (layout: include/fuzzyblue.h) -- the `vk::Image` / `vk::ImageView` getters of :2075-2101.
    pub fn transmittance(&self) -> *const c_void { let mut p = std::ptr::null(); check(unsafe { ffi::fb_atmosphere_transmittance(self.raw, &mut p, std::ptr::null_mut()) }); p }
    pub fn transmittance_view(&self) -> *const c_void { self.transmittance() }
    pub fn transmittance_extent(&self) -> vk::Extent2D { let mut e = ffi::FbExtent2D::default(); check(unsafe { ffi::fb_atmosphere_transmittance(self.raw, std::ptr::null_mut(), &mut e) }); vk::Extent2D { width: e.width, height: e.height } }
    pub fn scattering(&self) -> *const c_void { let mut p = std::ptr::null(); check(unsafe { ffi::fb_atmosphere_scattering(self.raw, &mut p, std::ptr::null_mut()) }); p }
    pub fn scattering_view(&self) -> *const c_void { self.scattering() }
    pub fn scattering_extent(&self) -> vk::Extent3D { let mut e = ffi::FbExtent3D::default(); check(unsafe { ffi::fb_atmosphere_scattering(self.raw, std::ptr::null_mut(), &mut e) }); vk::Extent3D { width: e.width, height: e.height, depth: e.depth } }
    pub fn irradiance(&self) -> *const c_void { let mut p = std::ptr::null(); check(unsafe { ffi::fb_atmosphere_irradiance(self.raw, &mut p, std::ptr::null_mut()) }); p }
    pub fn irradiance_view(&self) -> *const c_void { self.irradiance() }
    pub fn irradiance_extent(&self) -> vk::Extent2D { let mut e = ffi::FbExtent2D::default(); check(unsafe { ffi::fb_atmosphere_irradiance(self.raw, std::ptr::null_mut(), &mut e) }); vk::Extent2D { width: e.width, height: e.height } }
}
impl Drop for Atmosphere { fn drop(&mut self) { if self.owned { unsafe { ffi::fb_atmosphere_destroy(self.raw as *mut _) } } } }

/// src/precompute.rs:2103-2120.  Must outlive the work enqueued by `Atmosphere::build`.
pub struct PendingAtmosphere { builder: Arc<Builder>, raw: *mut ffi::FbPending }
impl PendingAtmosphere {
    /// Queue-family ownership transfer (:2147-2201) has no CUDA analogue.
    pub unsafe fn acquire_ownership(&self, _cmd: vk::CommandBuffer, _compute_queue_family: u32, _gfx_queue_family: u32) {}
    pub unsafe fn atmosphere(&self) -> Atmosphere { let mut a = std::ptr::null(); check(ffi::fb_pending_atmosphere(self.raw, &mut a)); Atmosphere { _builder: self.builder.clone(), raw: a, owned: false } }
    /// Caller asserts the stream has completed (:2208-2211); as in the reference nothing is verified (check = 0): the
    /// temporaries are recycled stream-ordered, so even a too-early call cannot corrupt a later precompute.
    pub unsafe fn assert_ready(mut self) -> Atmosphere {
        let mut a = std::ptr::null_mut(); check(ffi::fb_pending_assert_ready(self.raw, 0, &mut a)); self.raw = std::ptr::null_mut();
        Atmosphere { _builder: self.builder.clone(), raw: a, owned: true }
    }
    /// Not in the reference: block until the precompute has finished (the fence wait of tests/smoke.rs:147-155).
    pub fn wait(&self) { check(unsafe { ffi::fb_pending_wait(self.raw) }) }
    /// Re-submit the recorded stream (what benches/precompute.rs:138-148 does with its command buffer).
    pub unsafe fn resubmit(&self, cmd: vk::CommandBuffer) { check(ffi::fb_pending_resubmit(self.raw, cmd)) }
    /// Later `resubmit`s also copy the finished tables to these pinned host buffers (null = keep on the device),
    /// overlapped with the last kernels of the command stream.
    pub unsafe fn set_readback(&self, transmittance: *mut c_void, scattering: *mut c_void, irradiance: *mut c_void) {
        check(ffi::fb_pending_set_readback(self.raw, transmittance, scattering, irradiance))
    }
}
impl Drop for PendingAtmosphere { fn drop(&mut self) { if !self.raw.is_null() { unsafe { ffi::fb_pending_destroy(self.raw) } } } }

/// src/render.rs:246-252
#[derive(Debug, Copy, Clone)]
pub struct DrawParameters { pub inverse_viewproj: [[f32; 4]; 4], pub camera_position: [f32; 3], pub sun_direction: [f32; 3] }

/// src/render.rs:13-19.  Pipeline cache / render pass / subpass have no CUDA meaning; `frames` sizes the per-frame state.
pub struct Renderer { raw: *mut ffi::FbRenderer, depth: Vec<vk::DescriptorImageInfo>, color: *mut f32, transmittance: *mut f32 }
impl Renderer {
    /// `Renderer::new(&builder, cache, render_pass, subpass, frames)`, src/render.rs:34-40
    pub fn new(builder: &Builder, _cache: vk::PipelineCache, _render_pass: vk::RenderPass, _subpass: u32, frames: u32) -> Self {
        let mut raw = std::ptr::null_mut(); check(unsafe { ffi::fb_renderer_create(builder.raw, &mut raw) });
        let none = vk::DescriptorImageInfo { depth: std::ptr::null(), width: 0, height: 0 };
        Self { raw, depth: vec![none; frames as usize], color: std::ptr::null_mut(), transmittance: std::ptr::null_mut() }
    }
    /// src/render.rs:194-207: the depth attachment of a frame.
    pub unsafe fn set_depth_buffer(&mut self, frame: u32, image: &vk::DescriptorImageInfo) { self.depth[frame as usize] = *image; }
    /// Not in the reference (there the two fragment outputs land in the render pass's colour attachment through the
    /// dual-source blend, src/render.rs:124-137): `[h][w][4]` f32 device buffers for colour and transmittance.
    pub unsafe fn set_targets(&mut self, color: *mut f32, transmittance: *mut f32) { self.color = color; self.transmittance = transmittance; }
    /// `Renderer::draw(&self, cmd, &atmosphere, frame, &params)`, src/render.rs:209-215
    pub fn draw(&self, cmd: vk::CommandBuffer, atmosphere: &Atmosphere, frame: u32, params: &DrawParameters) {
        let raw = ffi::FbDrawParams { inverse_viewproj: params.inverse_viewproj, camera_position: params.camera_position, _pad: 0, sun_direction: params.sun_direction };
        let d = self.depth[frame as usize];
        check(unsafe { ffi::fb_renderer_draw(self.raw, atmosphere.raw, &raw, d.depth, self.color, self.transmittance, d.width, d.height, cmd) });
    }
}
impl Drop for Renderer { fn drop(&mut self) { unsafe { ffi::fb_renderer_destroy(self.raw) } } }
