// ref_common.cpp — per-invocation state shared by the shader translation units (see ref_shader.cpp).
#include <omp.h>

#include "glsl_compat.h"

namespace glsl {
thread_local uvec3 gl_GlobalInvocationID;
thread_local ivec2 frag_pixel;
}  // namespace glsl

extern "C" void fbr_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
extern "C" int fbr_max_threads(void) { return omp_get_max_threads(); }
extern "C" const char* fbr_source(void) { return "/root/reference/shaders (GLSL compiled as C++ through oracle/glsl_ref)"; }
