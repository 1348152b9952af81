// fb_sharded.cu — ONE atmosphere built by `world` ranks, one GPU each, the scattering table slabbed along its altitude
// (r) axis (BASELINE.json configs[2]).  The replacement for Atmosphere::build
// (src/precompute.rs:1077-1081) when the five 3-D images (2 GiB each at the high-resolution dims) are to be produced by
// several devices.
//
// Which stage needs which exchange follows from what each shader reads:
//   transmittance, direct_irradiance     tiny 2-D tables: computed replicated on every rank
//   single_scattering at r               r-local (reads only the transmittance table)
//   scattering_density at r              reads the previous order's 3-D table at the SAME r: the texel -> r -> u_r round
//                                        trip (scattering.h:62-67 then :17-22) lands within 1.2e-2 texels of the texel
//                                        centre, so the trilinear tap touches slice z and ONE neighbour
//                                          -> a one-slice HALO from each neighbouring rank, not an all-gather;
//                                        plus row 0 of delta_irradiance (GetIrradiance at r = bottom, irradiance.h:32-38)
//   multiple_scattering at r             marches along the ray through ALL r of scattering_density
//                                        (multiple_scattering.comp:35-44)  -> the one mandatory ALL-GATHER per order
//   indirect_irradiance row j            r_j is linear in j (irradiance.h:15-16); it reads the previous order's table at
//                                        the two slices bracketing r_j  -> each row is computed by the rank whose slab
//                                        (+ halo) holds them, and the finished rows (64 texels each) are broadcast
//   scattering += ...                    r-local
//
// The schedule is DATA: fb_sharded_plan() emits it as a list of steps (pure host logic, no device needed), and two
// executors run the same list — fb_pending_run_sharded() below with NCCL over NVLink / NVSwitch, and the CPU test's gloo
// executor over the oracle (tests/test_sharded_cpu.py), which proves every cross-slab dependency is covered by a step.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2", which finds the copy a host process such as PyTorch has already
// loaded): the library has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>   // types and enums only; every entry point goes through the table below

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fb_internal.h"

using namespace fb;

// ---------------------------------------------------------------------------------------------
// NCCL entry points, resolved lazily
// ---------------------------------------------------------------------------------------------
namespace {
struct Nccl {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommCount)(const ncclComm_t, int*);
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int*);
    bool ok;
    std::string why;
};
const Nccl& nccl_api() {
    static Nccl n = [] {
        Nccl t;
        std::memset(static_cast<void*>(&t), 0, offsetof(Nccl, ok));
        t.ok = false;
        void* h = nullptr;
        const char* env = std::getenv("FUZZYBLUE_B200_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { t.why = std::string("libnccl.so.2 could not be loaded (") + (dlerror() ? dlerror() : "?") + "); set FUZZYBLUE_B200_NCCL_LIB"; return t; }
        bool all = true;
        auto get = [&](const char* name, void** fn) {
            *fn = dlsym(h, name);
            if (!*fn) { all = false; t.why = std::string("libnccl lacks ") + name; }
        };
        get("ncclGetUniqueId", (void**)&t.GetUniqueId); get("ncclCommInitRank", (void**)&t.CommInitRank);
        get("ncclCommDestroy", (void**)&t.CommDestroy); get("ncclCommCount", (void**)&t.CommCount);
        get("ncclCommUserRank", (void**)&t.CommUserRank); get("ncclAllGather", (void**)&t.AllGather);
        get("ncclBroadcast", (void**)&t.Broadcast); get("ncclSend", (void**)&t.Send); get("ncclRecv", (void**)&t.Recv);
        get("ncclGroupStart", (void**)&t.GroupStart); get("ncclGroupEnd", (void**)&t.GroupEnd);
        get("ncclGetErrorString", (void**)&t.GetErrorString); get("ncclGetVersion", (void**)&t.GetVersion);
        t.ok = all;
        return t;
    }();
    return n;
}
int nccl_fail(ncclResult_t r, const char* what) {
    const Nccl& n = nccl_api();
    (void)cudaGetLastError();
    return fail(FB_ERR_CUDA, std::string(what) + ": NCCL " + (n.ok ? n.GetErrorString(r) : "unavailable"));
}
#define FB_NCCL(call)                                             \
    do {                                                          \
        ncclResult_t r__ = (call);                                \
        if (r__ != ncclSuccess) return nccl_fail(r__, #call);     \
    } while (0)
}  // namespace

int fb_nccl_version(int* version) {
    const Nccl& n = nccl_api();
    if (!n.ok) return fail(FB_ERR_CUDA, n.why);
    int v = 0;
    FB_NCCL(n.GetVersion(&v));
    if (version) *version = v;
    return FB_OK;
}
int fb_nccl_unique_id(void* id128) {
    if (!id128) return fail(FB_ERR_INVALID_ARGUMENT, "fb_nccl_unique_id: NULL");
    const Nccl& n = nccl_api();
    if (!n.ok) return fail(FB_ERR_CUDA, n.why);
    static_assert(sizeof(ncclUniqueId) == FB_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    FB_NCCL(n.GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return FB_OK;
}
int fb_nccl_comm_create(int device, int world, int rank, const void* id128, void** comm_out) {
    if (!id128 || !comm_out || world < 1 || rank < 0 || rank >= world) return fail(FB_ERR_INVALID_ARGUMENT, "fb_nccl_comm_create");
    *comm_out = nullptr;
    const Nccl& n = nccl_api();
    if (!n.ok) return fail(FB_ERR_CUDA, n.why);
    DeviceGuard g(device);
    if (!g.ok) return fail(FB_ERR_CUDA, "cudaSetDevice failed");
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    ncclComm_t c = nullptr;
    FB_NCCL(n.CommInitRank(&c, world, id, rank));
    *comm_out = c;
    return FB_OK;
}
int fb_nccl_comm_destroy(void* comm) {
    if (!comm) return FB_OK;
    const Nccl& n = nccl_api();
    if (!n.ok) return fail(FB_ERR_CUDA, n.why);
    FB_NCCL(n.CommDestroy((ncclComm_t)comm));
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------
// the plan
// ---------------------------------------------------------------------------------------------
namespace {
// rows [e0, e1) of the irradiance table whose two bracketing scattering slices lie in [a - 1, b] (slab + one-slice halo):
// row j sits at r_j = bottom + j / (E_r - 1) (top - bottom) (irradiance.h:15-16), i.e. at the slice coordinate
// x_j = rho_j / H (R - 1) of the scattering table (scattering.h:17-22 with Vulkan's u N - 0.5).  Rank q owns the rows with
// a_q - 0.5 <= x_j < b_q - 0.5: the kernel's fp32 coordinate differs from x_j by at most ~1e-2, so its floor() stays within
// [a - 1, b - 1] and the tap within [a - 1, b].
void irradiance_rows_of(const FbParams& P, int rank, int world, int* e0, int* e1) {
    const int R = P.scattering_r_size, ER = P.irradiance_r_size, n = R / world;
    const double bot = P.bottom_radius, top = P.top_radius, H = std::sqrt(top * top - bot * bot);
    const double lo = rank == 0 ? -1e300 : rank * n - 0.5, hi = rank == world - 1 ? 1e300 : (rank + 1) * n - 0.5;
    int b = ER, e = ER;
    for (int j = ER - 1; j >= 0; --j) {
        const double r = bot + (double)j / (ER - 1) * (top - bot);
        const double x = std::sqrt(std::max(r * r - bot * bot, 0.0)) / H * (R - 1);
        if (x >= hi) e = j;          // rows are monotone in x: everything from j on belongs to later ranks
        if (x >= lo) b = j;
    }
    *e0 = b; *e1 = e;
}
void push(std::vector<FbShardStep>& v, int op, int stage, int image, uint32_t order, uint32_t begin, uint32_t end, int root = 0) {
    FbShardStep s;
    s.op = op; s.stage = stage; s.image = image; s.order = order; s.begin = begin; s.end = end; s.root = root; s._pad = 0;
    v.push_back(s);
}
int make_plan(const FbParams& P, uint32_t order, int rank, int world, uint32_t flags, std::vector<FbShardStep>& v) {
    const int R = P.scattering_r_size;
    if (world < 1 || rank < 0 || rank >= world) return fail(FB_ERR_INVALID_ARGUMENT, "sharded: bad rank / world");
    if (R % world) return fail(FB_ERR_INVALID_ARGUMENT, "sharded: scattering_r_size must divide by the number of ranks (32 / 8 and 128 / 8 do)");
    const int n = R / world, a = rank * n, b = a + n;
    const bool multi = world > 1;
    // the two big producers run in sub-slabs whose exchange overlaps the next sub-slab's kernels; splitting only pays
    // when a sub-slab is a real transfer (>= 16 MiB), small tables go in one piece
    const size_t slice = image_bytes(P, FB_IMAGE_SCATTERING) / R;
    int chunks = 1;
    static const int max_chunks = [] { const char* e = std::getenv("FUZZYBLUE_B200_SHARD_CHUNKS"); const int v = e ? std::atoi(e) : 0; return v > 0 ? v : 4; }();
    if (multi && !(flags & FB_SHARD_NO_PIPELINE))
        for (int c = std::min(max_chunks, n); c >= 1; --c)
            if (n % c == 0 && ((flags & FB_SHARD_PIPELINE_ALWAYS) || (size_t)(n / c) * slice >= ((size_t)16 << 20))) { chunks = c; break; }
    const int m = n / chunks;
    push(v, FB_SHARD_STAGE, FB_STAGE_TRANSMITTANCE, 0, 0, 0, 0);
    push(v, FB_SHARD_STAGE, FB_STAGE_DIRECT_IRRADIANCE, 0, 0, 0, 0);
    push(v, FB_SHARD_STAGE, FB_STAGE_SINGLE_SCATTERING, 0, 0, a, b);
    push(v, FB_SHARD_STAGE, FB_STAGE_CLEAR_IRRADIANCE, 0, 0, 0, 0);
    if (order >= 2 && multi) {       // order-2 density and the order-1 irradiance rows read the single-scattering tables
        push(v, FB_SHARD_HALO, 0, FB_IMAGE_DELTA_RAYLEIGH, 0, 0, 0);
        push(v, FB_SHARD_HALO, 0, FB_IMAGE_DELTA_MIE, 0, 0, 0);
        push(v, FB_SHARD_JOIN, 0, 0, 0, 0, 0);
    }
    int e0 = 0, e1 = P.irradiance_r_size;
    if (multi) irradiance_rows_of(P, rank, world, &e0, &e1);
    bool result_gathered = false;
    for (uint32_t o = 2; o <= order; ++o) {
        for (int c = 0; c < chunks; ++c) {
            push(v, FB_SHARD_STAGE, FB_STAGE_SCATTERING_DENSITY, 0, o, a + c * m, a + (c + 1) * m);
            if (multi) push(v, FB_SHARD_ALLGATHER, 0, FB_IMAGE_SCATTERING_DENSITY, 0, c * m, (c + 1) * m);
        }
        if (e1 > e0) push(v, FB_SHARD_STAGE, FB_STAGE_INDIRECT_IRRADIANCE, 0, o - 1, e0, e1);
        // The next density pass looks delta_irradiance up at r = bottom (the ground term, irradiance.h:32-38): row 0, and --
        // where (0.5 / E_r) * E_r - 0.5 does not round to exactly 0 in fp32 -- row 1 with a weight of ~1e-8.  Both rows
        // travel from their owners so that the sharded tables stay bit-identical to the single-GPU ones.
        if (multi && o < order)
            for (int j = 0; j < std::min(2, (int)P.irradiance_r_size); ++j)
                for (int q = 0; q < world; ++q) {
                    int q0, q1;
                    irradiance_rows_of(P, q, world, &q0, &q1);
                    if (j >= q0 && j < q1) push(v, FB_SHARD_BCAST_ROWS, 0, FB_IMAGE_DELTA_IRRADIANCE, 0, j, j + 1, q);
                }
        if (multi) push(v, FB_SHARD_JOIN, 0, 0, 0, 0, 0);
        if (multi && o == order && (flags & FB_SHARD_GATHER_RESULT) && chunks > 1) {
            // the last pass finishes `scattering`: its sub-slabs leave for the other ranks behind the next sub-slab's kernel
            for (int c = 0; c < chunks; ++c) {
                push(v, FB_SHARD_STAGE, FB_STAGE_MULTIPLE_SCATTERING, 0, 0, a + c * m, a + (c + 1) * m);
                push(v, FB_SHARD_ALLGATHER, 0, FB_IMAGE_SCATTERING, 0, c * m, (c + 1) * m);
            }
            result_gathered = true;
        } else {
            push(v, FB_SHARD_STAGE, FB_STAGE_MULTIPLE_SCATTERING, 0, 0, a, b);
        }
        if (multi && o < order) {    // next order: density halo + irradiance rows read delta_multiple_scattering
            push(v, FB_SHARD_HALO, 0, FB_IMAGE_DELTA_MULTIPLE_SCATTERING, 0, 0, 0);
            push(v, FB_SHARD_JOIN, 0, 0, 0, 0, 0);
        }
    }
    if (multi) {
        if (order >= 2)              // every rank ends with the whole irradiance table
            for (int q = 0; q < world; ++q) {
                int q0, q1;
                irradiance_rows_of(P, q, world, &q0, &q1);
                if (q1 > q0) push(v, FB_SHARD_BCAST_ROWS, 0, FB_IMAGE_IRRADIANCE, 0, q0, q1, q);
            }
        if ((flags & FB_SHARD_GATHER_RESULT) && !result_gathered) push(v, FB_SHARD_ALLGATHER, 0, FB_IMAGE_SCATTERING, 0, 0, n);
        push(v, FB_SHARD_JOIN, 0, 0, 0, 0, 0);
    }
    return FB_OK;
}
}  // namespace

int fb_sharded_plan(const FbParams* p, uint32_t order, int rank, int world, uint32_t flags, FbShardStep* steps, uint32_t capacity,
                    uint32_t* count) {
    if (!p || !count) return fail(FB_ERR_INVALID_ARGUMENT, "fb_sharded_plan: NULL");
    int st = fb_params_validate(p);
    if (st != FB_OK) return st;
    if (order < 1 || order > 64) return fail(FB_ERR_INVALID_ARGUMENT, "order must be in [1, 64]");
    std::vector<FbShardStep> v;
    if ((st = make_plan(*p, order, rank, world, flags, v)) != FB_OK) return st;
    *count = (uint32_t)v.size();
    if (steps) {
        if (capacity < v.size()) return fail(FB_ERR_INVALID_ARGUMENT, "fb_sharded_plan: capacity too small (call with steps = NULL for the count)");
        std::memcpy(steps, v.data(), v.size() * sizeof(FbShardStep));
    }
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------
// the NCCL executor
// ---------------------------------------------------------------------------------------------
int fb_pending_run_sharded(FbPending* p, void* nccl_comm, int rank, int world, uint32_t flags, void* stream) {
    if (!p || !p->inner) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_run_sharded: NULL / already taken");
    std::vector<FbShardStep> plan;
    int st = make_plan(p->P, p->order, rank, world, flags, plan);
    if (st != FB_OK) return st;
    const Nccl& N = nccl_api();
    ncclComm_t comm = (ncclComm_t)nccl_comm;
    if (world > 1) {
        if (!comm) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_run_sharded: world > 1 needs an NCCL communicator");
        if (!N.ok) return fail(FB_ERR_CUDA, N.why);
        int cw = 0, cr = -1;
        FB_NCCL(N.CommCount(comm, &cw));
        FB_NCCL(N.CommUserRank(comm, &cr));
        if (cw != world || cr != rank) return fail(FB_ERR_INVALID_ARGUMENT, "fb_pending_run_sharded: rank / world do not match the communicator");
    }
    DeviceGuard g(p->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (world > 1 && !p->comm) {
        FB_CUDA(cudaStreamCreateWithFlags(&p->comm, cudaStreamNonBlocking));
        FB_CUDA(cudaEventCreateWithFlags(&p->ev_comm, cudaEventDisableTiming));
    }
    ++p->inner->version;
    LaunchCtx c = make_ctx(p, s);
    const int R = p->P.scattering_r_size, n = R / world, a = rank * n, b = a + n;
    int launches = 0;
    auto base = [&](int image) { return static_cast<char*>(image_ptr(p, image)); };
    // an exchange starts when everything enqueued so far on the compute stream has finished (its producer and every
    // earlier reader of the regions it overwrites)
    auto comm_after_compute = [&]() -> int {
        FB_CUDA(cudaEventRecord(p->ev_comm, s));
        FB_CUDA(cudaStreamWaitEvent(p->comm, p->ev_comm, 0));
        return FB_OK;
    };
    for (const FbShardStep& k : plan) {
        switch (k.op) {
            case FB_SHARD_STAGE:
                if ((st = run_stage(p, c, k.stage, k.order, (int)k.begin, (int)k.end, &launches)) != FB_OK) return st;
                break;
            case FB_SHARD_ALLGATHER: {   // rows [begin, end) of every rank's slab, in place
                if ((st = comm_after_compute()) != FB_OK) return st;
                const size_t slice = image_bytes(p->P, k.image) / R;
                char* img = base(k.image);
                if (k.begin == 0 && (int)k.end == n) {
                    FB_NCCL(N.AllGather(img + (size_t)a * slice, img, (size_t)n * slice, ncclInt8, comm, p->comm));
                } else {                 // a sub-slab is strided across ranks: one in-place broadcast per owner, fused in a group
                    FB_NCCL(N.GroupStart());
                    for (int q = 0; q < world; ++q) {
                        char* part = img + ((size_t)q * n + k.begin) * slice;
                        const ncclResult_t r = N.Broadcast(part, part, (size_t)(k.end - k.begin) * slice, ncclInt8, q, comm, p->comm);
                        if (r != ncclSuccess) { N.GroupEnd(); return nccl_fail(r, "ncclBroadcast (sub-slab)"); }
                    }
                    FB_NCCL(N.GroupEnd());
                }
                break;
            }
            case FB_SHARD_HALO: {        // first / last slice of the slab to the neighbours, their edge slices into the halo
                if ((st = comm_after_compute()) != FB_OK) return st;
                const size_t slice = image_bytes(p->P, k.image) / R;
                char* img = base(k.image);
                FB_NCCL(N.GroupStart());
                ncclResult_t r = ncclSuccess;
                if (rank > 0) {
                    r = N.Send(img + (size_t)a * slice, slice, ncclInt8, rank - 1, comm, p->comm);
                    if (r == ncclSuccess) r = N.Recv(img + (size_t)(a - 1) * slice, slice, ncclInt8, rank - 1, comm, p->comm);
                }
                if (r == ncclSuccess && rank < world - 1) {
                    r = N.Send(img + (size_t)(b - 1) * slice, slice, ncclInt8, rank + 1, comm, p->comm);
                    if (r == ncclSuccess) r = N.Recv(img + (size_t)b * slice, slice, ncclInt8, rank + 1, comm, p->comm);
                }
                if (r != ncclSuccess) { N.GroupEnd(); return nccl_fail(r, "ncclSend / ncclRecv (halo)"); }
                FB_NCCL(N.GroupEnd());
                break;
            }
            case FB_SHARD_BCAST_ROWS: {  // rows [begin, end) of a 2-D image from their owner
                if ((st = comm_after_compute()) != FB_OK) return st;
                const size_t row = (size_t)p->P.irradiance_mu_s_size * sizeof(float4);
                char* part = base(k.image) + (size_t)k.begin * row;
                FB_NCCL(N.Broadcast(part, part, (size_t)(k.end - k.begin) * row, ncclInt8, k.root, comm, p->comm));
                break;
            }
            case FB_SHARD_JOIN:          // later stages on the compute stream wait for every exchange issued so far
                FB_CUDA(cudaEventRecord(p->ev_comm, p->comm));
                FB_CUDA(cudaStreamWaitEvent(s, p->ev_comm, 0));
                break;
            default: return fail(FB_ERR_INVALID_ARGUMENT, "sharded: unknown plan step");
        }
    }
    p->launches = launches;
    p->done->note(s);
    if (p->comm) p->done->note(p->comm);
    return FB_OK;
}

int fb_atmosphere_build_sharded(FbBuilder* b, const FbParams* params, uint32_t order, void* nccl_comm, int rank, int world,
                                uint32_t flags, void* stream, FbPending** out) {
    int st = fb_atmosphere_allocate(b, params, order, out);
    if (st != FB_OK) return st;
    st = fb_pending_run_sharded(*out, nccl_comm, rank, world, flags, stream);
    if (st != FB_OK) {
        (*out)->done->note((cudaStream_t)stream);   // stages enqueued before the failure are in flight on the blocks
        fb_pending_destroy(*out);
        *out = nullptr;
    }
    return st;
}
