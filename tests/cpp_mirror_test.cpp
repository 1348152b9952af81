// C++ mirror smoke test: mirrors /root/reference/tests/smoke.rs (reduced dims, build, wait, sanity) through
// include/fuzzyblue.hpp.  Exit 0 = pass, 77 = no GPU (expected on the CPU box: the library has no fallback).
#include <cmath>
#include <cstdio>
#include <cuda_runtime_api.h>
#include "../include/fuzzyblue.hpp"

int main() {
    using namespace fuzzyblue;
    static_assert(sizeof(FbParams) == 320 && sizeof(FbDrawParams) == 92, "ABI block sizes");
    Parameters params;   // Parameters::default()
    if (params.raw.bottom_radius != 6360.0f || params.order != 4 || params.scattering_extent()[0] != 256) return 2;
    std::shared_ptr<Builder> builder;
    try {
        builder = std::make_shared<Builder>(0);
    } catch (const Error& e) {
        std::printf("no device: %s\n", e.what());
        return e.status == FB_ERR_NO_DEVICE ? 77 : 3;
    }
    // "Simplified for speed", tests/smoke.rs:135-142
    params.raw.scattering_r_size = 8; params.raw.scattering_mu_size = 32; params.raw.scattering_mu_s_size = 8; params.raw.scattering_nu_size = 2;
    cudaStream_t stream;
    cudaStreamCreate(&stream);
    PendingAtmosphere pending = build(builder, stream, params);
    cudaStreamSynchronize(stream);
    // the read-back recorded into the command stream (examples/dump.rs:175-193 records its copies the same way)
    const size_t nT = (size_t)params.raw.transmittance_mu_size * params.raw.transmittance_r_size * 4;
    const size_t nS = (size_t)params.raw.scattering_nu_size * params.raw.scattering_mu_s_size * params.raw.scattering_mu_size *
                      params.raw.scattering_r_size * 4;
    float* hT = nullptr;
    unsigned short* hS = nullptr;
    if (cudaMallocHost((void**)&hT, nT * sizeof(float)) != cudaSuccess || cudaMallocHost((void**)&hS, nS * 2) != cudaSuccess) return 7;
    for (size_t i = 0; i < nS; ++i) hS[i] = 0xffffu;
    pending.set_readback(hT, hS, nullptr);
    pending.resubmit(stream);
    cudaStreamSynchronize(stream);
    Atmosphere atm = std::move(pending).assert_ready();
    auto T = atm.read_transmittance(stream);
    auto E = atm.read_irradiance(stream);
    cudaStreamSynchronize(stream);
    // KAT (iii): r = top, mu = 1 -> T = 1; alpha = 1
    size_t last_row = (size_t)(atm.transmittance_extent().height - 1) * atm.transmittance_extent().width * 4;
    if (T[last_row] != 1.0f || T[last_row + 3] != 1.0f) return 4;
    for (size_t i = 0; i < nT; ++i) if (hT[i] != T[i]) return 8;
    for (size_t i = 3; i < nS; i += 4) if (hS[i] == 0xffffu) return 9;     // every texel of every r-slab arrived
    cudaFreeHost(hT); cudaFreeHost(hS);
    double sum = 0;
    for (float v : E) { if (!std::isfinite(v)) return 5; sum += v; }
    if (!(sum > 0)) return 6;
    std::printf("cpp mirror ok: irradiance sum %.6f\n", sum);
    return 0;
}
