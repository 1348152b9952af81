// fuzzyblue.hpp — header-only C++ mirror of the reference crate's public API over the C ABI of fuzzyblue.h.
// Same names as /root/reference/src/lib.rs:8-12: Builder, Parameters, Atmosphere, PendingAtmosphere,
// DrawParameters, Renderer.  RAII where the reference has Drop; a cudaStream_t (void*) where it takes a
// vk::CommandBuffer; exceptions (fuzzyblue::Error) where it panics through .unwrap().
#ifndef FUZZYBLUE_HPP_
#define FUZZYBLUE_HPP_

#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "fuzzyblue.h"

namespace fuzzyblue {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& m) : std::runtime_error(std::string(fb_status_string(s)) + ": " + m), status(s) {}
};
inline void check(int status) {
    if (status != FB_OK) throw Error(status, fb_last_error());
}

using DensityProfileLayer = FbDensityProfileLayer;   // src/precompute.rs:660-666
using DensityProfile = FbDensityProfile;             // src/precompute.rs:674-676

// src/precompute.rs:690-769.  Default-constructed == Parameters::default() (:849-935).
struct Parameters {
    uint32_t usage = 0, dst_stage_mask = 0x80, dst_access_mask = 0x20;   // Vulkan-only fields: kept, ignored
    int32_t layout = 5;
    uint32_t order = fb_params_default_order();
    FbParams raw{};   // every physical field and LUT size, in the layout of shaders/params.h
    Parameters() { check(fb_params_default(&raw)); }
    std::pair<uint32_t, uint32_t> transmittance_extent() const { return {(uint32_t)raw.transmittance_mu_size, (uint32_t)raw.transmittance_r_size}; }
    std::pair<uint32_t, uint32_t> irradiance_extent() const { return {(uint32_t)raw.irradiance_mu_s_size, (uint32_t)raw.irradiance_r_size}; }
    std::array<uint32_t, 3> scattering_extent() const {
        return {(uint32_t)(raw.scattering_nu_size * raw.scattering_mu_s_size), (uint32_t)raw.scattering_mu_size, (uint32_t)raw.scattering_r_size};
    }
};

using DrawParameters = FbDrawParams;   // src/render.rs:246-252 (with the raw block's padding)

class Builder {   // src/precompute.rs:32-68
public:
    explicit Builder(int device = 0) { check(fb_builder_create(device, &h_)); }
    ~Builder() { fb_builder_destroy(h_); }
    Builder(const Builder&) = delete;
    Builder& operator=(const Builder&) = delete;
    void set_kernels(int kernels) { check(fb_builder_set_kernels(h_, kernels)); }
    void set_exportable(bool on) { check(fb_builder_set_exportable(h_, on ? 1 : 0)); }   // Vulkan interop, see fuzzyblue.h
    FbBuilder* handle() const { return h_; }
private:
    FbBuilder* h_ = nullptr;
};

class Atmosphere {   // src/precompute.rs:1036-1101
public:
    Atmosphere(std::shared_ptr<Builder> b, const FbAtmosphere* a, bool owned) : builder_(std::move(b)), a_(a), owned_(owned) {}
    Atmosphere(Atmosphere&& o) noexcept : builder_(std::move(o.builder_)), a_(o.a_), owned_(o.owned_) { o.a_ = nullptr; o.owned_ = false; }
    Atmosphere(const Atmosphere&) = delete;
    ~Atmosphere() { if (owned_ && a_) fb_atmosphere_destroy(const_cast<FbAtmosphere*>(a_)); }
    const void* transmittance() const { const void* p; check(fb_atmosphere_transmittance(a_, &p, nullptr)); return p; }
    FbExtent2D transmittance_extent() const { FbExtent2D e; check(fb_atmosphere_transmittance(a_, nullptr, &e)); return e; }
    const void* scattering() const { const void* p; check(fb_atmosphere_scattering(a_, &p, nullptr)); return p; }
    FbExtent3D scattering_extent() const { FbExtent3D e; check(fb_atmosphere_scattering(a_, nullptr, &e)); return e; }
    const void* irradiance() const { const void* p; check(fb_atmosphere_irradiance(a_, &p, nullptr)); return p; }
    FbExtent2D irradiance_extent() const { FbExtent2D e; check(fb_atmosphere_irradiance(a_, nullptr, &e)); return e; }
    // examples/dump.rs:110-193
    std::vector<float> read_transmittance(void* stream = nullptr) const {
        FbExtent2D e = transmittance_extent();
        std::vector<float> v((size_t)e.width * e.height * 4);
        check(fb_atmosphere_read_transmittance(a_, v.data(), v.size() * 4, stream));
        return v;
    }
    std::vector<uint16_t> read_scattering(void* stream = nullptr) const {   // IEEE half bits
        FbExtent3D e = scattering_extent();
        std::vector<uint16_t> v((size_t)e.width * e.height * e.depth * 4);
        check(fb_atmosphere_read_scattering(a_, v.data(), v.size() * 2, stream));
        return v;
    }
    std::vector<float> read_irradiance(void* stream = nullptr) const {
        FbExtent2D e = irradiance_extent();
        std::vector<float> v((size_t)e.width * e.height * 4);
        check(fb_atmosphere_read_irradiance(a_, v.data(), v.size() * 4, stream));
        return v;
    }
    // Vulkan interop: a descriptor of the block holding the three tables (the caller owns it) and where they sit
    int export_fd(FbExportLayout* layout = nullptr) const { int fd = -1; check(fb_atmosphere_export_fd(a_, &fd, layout)); return fd; }
    const FbAtmosphere* handle() const { return a_; }
private:
    std::shared_ptr<Builder> builder_;
    const FbAtmosphere* a_;
    bool owned_;
};

class PendingAtmosphere {   // src/precompute.rs:2103-2211
public:
    PendingAtmosphere(std::shared_ptr<Builder> b, FbPending* p) : builder_(std::move(b)), p_(p) {}
    PendingAtmosphere(PendingAtmosphere&& o) noexcept : builder_(std::move(o.builder_)), p_(o.p_) { o.p_ = nullptr; }
    PendingAtmosphere(const PendingAtmosphere&) = delete;
    ~PendingAtmosphere() { if (p_) fb_pending_destroy(p_); }
    void acquire_ownership(void* /*stream*/, uint32_t /*compute_queue_family*/, uint32_t /*gfx_queue_family*/) const {}   // :2147-2201: no CUDA analogue
    Atmosphere atmosphere() const { const FbAtmosphere* a; check(fb_pending_atmosphere(p_, &a)); return Atmosphere(builder_, a, false); }
    Atmosphere assert_ready(bool check_stream = true) && {
        FbAtmosphere* a;
        check(fb_pending_assert_ready(p_, check_stream ? 1 : 0, &a));
        p_ = nullptr;
        return Atmosphere(builder_, a, true);
    }
    void resubmit(void* stream) { check(fb_pending_resubmit(p_, stream)); }   // benches/precompute.rs:138-148
    // later resubmits also copy the finished tables to these (pinned) host buffers, overlapped with the last kernels
    void set_readback(void* transmittance, void* scattering, void* irradiance) {
        check(fb_pending_set_readback(p_, transmittance, scattering, irradiance));
    }
    FbPending* handle() const { return p_; }
private:
    std::shared_ptr<Builder> builder_;
    FbPending* p_;
};

// Atmosphere::build(builder, cmd, &params), src/precompute.rs:1077-1081
inline PendingAtmosphere build(std::shared_ptr<Builder> builder, void* stream, const Parameters& params) {
    FbPending* p;
    check(fb_atmosphere_build(builder->handle(), &params.raw, params.order, stream, &p));
    return PendingAtmosphere(std::move(builder), p);
}

class Renderer {   // src/render.rs:13-236
public:
    explicit Renderer(const Builder& b, uint32_t frames = 1) : depth_(frames, nullptr) { check(fb_renderer_create(b.handle(), &r_)); }
    ~Renderer() { fb_renderer_destroy(r_); }
    Renderer(const Renderer&) = delete;
    void set_depth_buffer(uint32_t frame, const float* depth_dev) { depth_.at(frame) = depth_dev; }   // :194-207
    void draw(void* stream, const Atmosphere& a, uint32_t frame, const DrawParameters& d, float* color, float* transmittance,
              uint32_t width, uint32_t height) const {                                                 // :209-236
        check(fb_renderer_draw(r_, a.handle(), &d, depth_.at(frame), color, transmittance, width, height, stream));
    }
private:
    FbRenderer* r_ = nullptr;
    std::vector<const float*> depth_;
};

}  // namespace fuzzyblue
#endif
