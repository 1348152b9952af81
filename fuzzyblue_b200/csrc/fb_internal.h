// fb_internal.h — object layouts and helpers shared by the translation units behind include/fuzzyblue.h
// (fb_api.cu: lifetimes, the recorded command stream, the renderer front-end; fb_sharded.cu: the r-slab sharded build).
#pragma once

#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/fuzzyblue.h"
#include "fb_kernels.h"

namespace fb {

int fail(int status, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define FB_CUDA(call)                                          \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return ::fb::cuda_fail(e__, #call); \
    } while (0)

// "Everything submitted so far that touches this object": one event per stream work was enqueued on, re-recorded after
// every submission.  It is what fb_pending_assert_ready(check = 1) queries (FB_ERR_NOT_READY instead of a device-wide
// synchronisation) and what makes the block cache stream-ordered: a released block carries the Completion of the work
// that may still be writing it, and the next owner waits for exactly that work — not for the whole device.
struct Completion {
    std::mutex m;
    std::vector<std::pair<cudaStream_t, cudaEvent_t>> ev;
    bool unknown = false;   // some work cannot be tracked (the caller is capturing the stream, a dead stream handle, > 32 streams)
    void note(cudaStream_t s);
    cudaError_t query();    // cudaSuccess, cudaErrorNotReady, or an error
    cudaError_t wait();     // blocks the calling thread until that work has finished
    // make `s` wait (on the device) for that work; falls back to a host wait when it cannot be tracked
    cudaError_t stream_wait(cudaStream_t s);
    ~Completion();
};

// Device blocks released by a finished precompute are kept (up to `limit` bytes) and handed to the next one of the
// same dims: cudaMalloc / cudaFree cost ~0.2 ms each and cudaFree synchronises the whole device, which would serialise
// a batch of independent atmospheres.  Shared by the builder and everything built from it, so an Atmosphere may
// outlive its Builder (the reference holds an Arc<Builder> for the same reason, precompute.rs:1037).
// Reuse is ordered: put() takes the Completion of the work that used the block, get() waits for it before handing the
// block out, so kernels still in flight on the previous owner's streams can never write into the next owner's tables.
struct BlockCache {
    struct Block {
        void* p;
        std::shared_ptr<Completion> busy;
    };
    int device;
    size_t limit, cached;
    std::mutex m;
    std::multimap<size_t, Block> free_blocks;
    explicit BlockCache(int dev) : device(dev), limit((size_t)8 << 30), cached(0) {}
    cudaError_t get(void** p, size_t n);
    void put(void* p, size_t n, const std::shared_ptr<Completion>& busy);
    void trim();
    ~BlockCache();
};

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(false) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

size_t image_bytes(const FbParams& P, int image);

}  // namespace fb

struct FbBuilder {
    int device;
    int sm_count;
    int kernels;
    int exportable;          // kept blocks come from the virtual-memory API as fd-exportable allocations
    fb::Trig trig;
    std::shared_ptr<fb::BlockCache> cache;
};

struct FbAtmosphere {
    int device;
    int kernels;
    std::shared_ptr<fb::BlockCache> cache;
    std::shared_ptr<fb::Completion> done;   // shared with the PendingAtmosphere that builds it; draws and reads add to it
    void* block;             // one device block: [scattering | transmittance | irradiance]
    size_t block_bytes;
    bool vmm;                // block is an exportable virtual-memory allocation (not from the cache)
    unsigned long long vmm_handle;
    size_t vmm_bytes;        // allocation size (block_bytes rounded up to the granularity)
    size_t off_transmittance, off_irradiance;
    FbParams P;
    float4* transmittance;
    float4* irradiance;
    uint2* scattering;
    // identity of the table CONTENTS for a renderer's derived copy: `serial` is unique per atmosphere, `version` counts
    // the submissions that (re)write the tables through the owning PendingAtmosphere
    uint64_t serial, version;
};

struct FbPending {
    // what the stages need of the Builder, copied at allocation: a pending may outlive the builder it came from
    int device, kernels, sm_count;
    fb::Trig trig;
    std::shared_ptr<fb::BlockCache> cache;
    std::shared_ptr<fb::Completion> done;
    void* temp_block;        // one device block for the five temporaries and the kernel scratch
    size_t temp_bytes;
    FbParams P;
    uint32_t order;
    fb::Images img;
    FbAtmosphere* inner;     // Option<Atmosphere>, precompute.rs:2114
    cudaGraphExec_t graph;   // pre-recorded command stream, instantiated lazily
    int launches;
    uint32_t slow_stages;    // bit s: stage s ran the one-thread-per-texel transcription although FAST kernels were asked for
    cudaStream_t side;       // indirect_irradiance overlaps the density main kernel here (FAST family)
    cudaEvent_t ev_fork, ev_join;
    // fb_pending_set_readback: host destinations recorded into the command stream.  The last multiple-scattering
    // pass runs as RB_SLABS r-slabs on streams of descending priority, each followed by the copy of its slab of
    // `scattering`, so the 8 MiB read-back hides behind the remaining slabs' kernels.
    void *rb_T, *rb_S, *rb_E;
    cudaStream_t rb_stream[4];
    cudaEvent_t rb_ev;
    // fb_pending_run_sharded: communication stream + ordering events of the r-slab exchanges
    cudaStream_t comm;
    cudaEvent_t ev_comm;
};

namespace fb {
LaunchCtx make_ctx(FbPending* p, cudaStream_t s);
// one stage on [r0, r1) (3-D stages: altitude levels; indirect irradiance: rows of the irradiance table)
int run_stage(FbPending* p, const LaunchCtx& c, int stage, uint32_t order, int r0, int r1, int* launches);
void* image_ptr(FbPending* p, int image);
}  // namespace fb
