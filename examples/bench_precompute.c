/* bench_precompute.c — the counterpart of the reference's benches/precompute.rs in plain C over the C ABI.
 *
 * benches/precompute.rs builds one atmosphere at reduced dims (scattering 8x32x8x2, :126-133), then times
 * `queue_submit(pre-recorded command buffer) + device_wait_idle` in a bencher loop (:138-148).  Here: one
 * fb_atmosphere_build, then fb_pending_resubmit + cudaStreamSynchronize per iteration.  Pass `default` as the first
 * argument for Parameters::default() dims (BASELINE.json configs[1]).
 *
 *   gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/bench_precompute.c -o bench_precompute \
 *       -Lfuzzyblue_b200/csrc -lfuzzyblue_b200 -Wl,-rpath,$PWD/fuzzyblue_b200/csrc -L/usr/local/cuda/lib64 -lcudart
 */
#define _POSIX_C_SOURCE 199309L
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include "fuzzyblue.h"

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int st_ = (call);                                                                   \
        if (st_ != FB_OK) {                                                                 \
            fprintf(stderr, "%s: %s: %s\n", #call, fb_status_string(st_), fb_last_error()); \
            return st_ == FB_ERR_NO_DEVICE ? 77 : 1;                                        \
        }                                                                                   \
    } while (0)

static double now_ns(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e9 + ts.tv_nsec;
}

int main(int argc, char** argv) {
    FbParams params;
    CHECK(fb_params_default(&params));
    if (!(argc > 1 && strcmp(argv[1], "default") == 0)) { /* "Simplified for speed", benches/precompute.rs:126 */
        params.scattering_r_size = 8;
        params.scattering_mu_size = 32;
        params.scattering_mu_s_size = 8;
        params.scattering_nu_size = 2;
    }
    FbBuilder* builder;
    CHECK(fb_builder_create(0, &builder));
    cudaStream_t stream;
    if (cudaStreamCreate(&stream) != cudaSuccess) return 1;
    FbPending* pending;
    CHECK(fb_atmosphere_build(builder, &params, fb_params_default_order(), stream, &pending));
    cudaStreamSynchronize(stream);
    const int iters = 200;
    for (int i = 0; i < 10; ++i) CHECK(fb_pending_resubmit(pending, stream));
    cudaStreamSynchronize(stream);
    double t0 = now_ns();
    for (int i = 0; i < iters; ++i) {
        CHECK(fb_pending_resubmit(pending, stream)); /* queue_submit */
        cudaStreamSynchronize(stream);               /* device_wait_idle */
    }
    double ns = (now_ns() - t0) / iters;
    printf("test precompute ... bench: %12.0f ns/iter (%d launches per submit)\n", ns, fb_pending_launch_count(pending));
    FbAtmosphere* atmosphere;
    CHECK(fb_pending_assert_ready(pending, 1, &atmosphere));
    fb_atmosphere_destroy(atmosphere);
    fb_builder_destroy(builder);
    cudaStreamDestroy(stream);
    return 0;
}
