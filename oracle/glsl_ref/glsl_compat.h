// glsl_compat.h — the small GLSL 450 subset the reference's shaders use, as C++, so that the reference's OWN shader
// sources (/root/reference/shaders/*.h, *.comp, render_sky.frag) can be compiled by g++ and executed on the CPU.
//
// TEST INFRASTRUCTURE ONLY (part of oracle/): it pins the CPU oracle against the reference itself.  Nothing here or in
// the generated files is shipped, linked or imported by fuzzyblue_b200.
//
// What comes from the reference: every line of shader logic (included, at build time, from where it lies under
// /root/reference; oracle/glsl_ref/translate.py only rewrites GLSL syntax that is not C++: `out` parameters, swizzles,
// interface blocks, float literals).  What is defined HERE, because in Vulkan it belongs to the driver and not to the
// shader source:
//   * scalar arithmetic: IEEE binary32, one rounding per operation (build with -ffp-contract=off, no fast-math);
//     exp / pow / sin / cos / sqrt = the C library's float functions;
//   * built-ins: clamp = min(max(x, lo), hi); mix(x, y, a) = x (1 - a) + y a; smoothstep, mod, dot (left to right),
//     length = sqrt(dot), normalize(v) = v / length(v) — the formulas of the GLSL specification;
//   * texture(): VkSampler{LINEAR, CLAMP_TO_EDGE, normalised coordinates} (src/precompute.rs:85-98): unnormalised
//     coordinate u N - 0.5, floor / fract, both taps clamped to the edge, separable blend a (1 - f) + b f per axis in fp32;
//   * imageStore(): the image VIEW formats the host creates (src/precompute.rs:1170, :1191, :1215): RGBA32F for the two
//     2-D tables, RGBA16F (round to nearest even) for the 3-D tables; imageLoad() returns what was stored.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl {

typedef unsigned int uint;

// ---------------------------------------------------------------------------------------------
// vectors
// ---------------------------------------------------------------------------------------------
struct vec2 {
    union { struct { float x, y; }; struct { float r, g; }; };
    vec2() : x(0), y(0) {}
    explicit vec2(float a) : x(a), y(a) {}
    vec2(float a, float b) : x(a), y(b) {}
    vec2 xy() const { return *this; }
};
struct vec4;
struct vec3 {
    union { struct { float x, y, z; }; struct { float r, g, b; }; };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(const vec4& v);
    vec3 xyz() const { return *this; }
    vec3 rgb() const { return *this; }
    vec2 xy() const { return vec2(x, y); }
};
struct vec4 {
    union { struct { float x, y, z, w; }; struct { float r, g, b, a; }; };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a_, float b_, float c_, float d_) : x(a_), y(b_), z(c_), w(d_) {}
    vec4(const vec3& v, float d_) : x(v.x), y(v.y), z(v.z), w(d_) {}
    vec4(const vec2& v, float c_, float d_) : x(v.x), y(v.y), z(c_), w(d_) {}
    vec3 xyz() const { return vec3(x, y, z); }
    vec3 rgb() const { return vec3(x, y, z); }
    vec2 xy() const { return vec2(x, y); }
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

struct uvec2 { uint x, y; uvec2() : x(0), y(0) {} uvec2(uint a, uint b) : x(a), y(b) {} };
struct uvec3 {
    uint x, y, z;
    uvec3() : x(0), y(0), z(0) {}
    uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
    uvec2 xy() const { return uvec2(x, y); }
};
struct ivec2 { int x, y; ivec2() : x(0), y(0) {} ivec2(int a, int b) : x(a), y(b) {}
               explicit ivec2(const uvec2& v) : x((int)v.x), y((int)v.y) {} explicit ivec2(const uvec3& v) : x((int)v.x), y((int)v.y) {} };
struct ivec3 { int x, y, z; ivec3() : x(0), y(0), z(0) {} ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
               explicit ivec3(const uvec3& v) : x((int)v.x), y((int)v.y), z((int)v.z) {} };
struct bvec2 { bool x, y; };
struct bvec3 { bool x, y, z; };

#define GLSL_VEC_OPS(V, ...)                                                                                       \
    inline V operator+(const V& a, const V& b) { V o; __VA_ARGS__(o, a, b, +) return o; }                         \
    inline V operator-(const V& a, const V& b) { V o; __VA_ARGS__(o, a, b, -) return o; }                         \
    inline V operator*(const V& a, const V& b) { V o; __VA_ARGS__(o, a, b, *) return o; }                         \
    inline V operator/(const V& a, const V& b) { V o; __VA_ARGS__(o, a, b, /) return o; }                         \
    inline V operator+(const V& a, float s) { return a + V(s); }                                                   \
    inline V operator-(const V& a, float s) { return a - V(s); }                                                   \
    inline V operator*(const V& a, float s) { return a * V(s); }                                                   \
    inline V operator/(const V& a, float s) { return a / V(s); }                                                   \
    inline V operator+(float s, const V& a) { return V(s) + a; }                                                   \
    inline V operator-(float s, const V& a) { return V(s) - a; }                                                   \
    inline V operator*(float s, const V& a) { return V(s) * a; }                                                   \
    inline V operator/(float s, const V& a) { return V(s) / a; }                                                   \
    inline V operator-(const V& a) { return neg(a); }                                                              \
    inline V& operator+=(V& a, const V& b) { a = a + b; return a; }                                                \
    inline V& operator-=(V& a, const V& b) { a = a - b; return a; }                                                \
    inline V& operator*=(V& a, const V& b) { a = a * b; return a; }                                                \
    inline V& operator*=(V& a, float s) { a = a * s; return a; }                                                   \
    inline V& operator/=(V& a, float s) { a = a / s; return a; }
#define GLSL_E2(o, a, b, op) o.x = a.x op b.x; o.y = a.y op b.y;
#define GLSL_E3(o, a, b, op) o.x = a.x op b.x; o.y = a.y op b.y; o.z = a.z op b.z;
#define GLSL_E4(o, a, b, op) o.x = a.x op b.x; o.y = a.y op b.y; o.z = a.z op b.z; o.w = a.w op b.w;
inline vec2 neg(const vec2& a) { return vec2(-a.x, -a.y); }
inline vec3 neg(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 neg(const vec4& a) { return vec4(-a.x, -a.y, -a.z, -a.w); }
GLSL_VEC_OPS(vec2, GLSL_E2)
GLSL_VEC_OPS(vec3, GLSL_E3)
GLSL_VEC_OPS(vec4, GLSL_E4)

struct mat4 {                 // column-major: c[i] is column i
    vec4 c[4];
};
inline vec4 operator*(const mat4& m, const vec4& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z + m.c[3] * v.w; }

// ---------------------------------------------------------------------------------------------
// built-in functions (GLSL specification, chapter 8)
// ---------------------------------------------------------------------------------------------
inline float sqrt(float x) { return ::sqrtf(x); }
inline float exp(float x) { return ::expf(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float floor(float x) { return ::floorf(x); }
inline float abs(float x) { return ::fabsf(x); }
inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float x, float y, float a) { return x * (1.f - a) + y * a; }
inline float mod(float x, float y) { return x - y * floor(x / y); }
inline bool isinf(float x) { return std::isinf(x); }
inline float smoothstep(float e0, float e1, float x) {
    float t = clamp((x - e0) / (e1 - e0), 0.f, 1.f);
    return t * t * (3.f - 2.f * t);
}
inline vec3 exp(const vec3& v) { return vec3(exp(v.x), exp(v.y), exp(v.z)); }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline vec4 mix(const vec4& x, const vec4& y, float a) { return x * (1.f - a) + y * a; }
inline vec3 mix(const vec3& x, const vec3& y, float a) { return x * (1.f - a) + y * a; }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(const vec3& v) { return sqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v / length(v); }
inline bvec2 greaterThanEqual(const uvec2& a, const uvec2& b) { return bvec2{a.x >= b.x, a.y >= b.y}; }
inline bvec3 greaterThanEqual(const uvec3& a, const uvec3& b) { return bvec3{a.x >= b.x, a.y >= b.y, a.z >= b.z}; }
inline bool any(const bvec2& b) { return b.x || b.y; }
inline bool any(const bvec3& b) { return b.x || b.y || b.z; }

// ---------------------------------------------------------------------------------------------
// images and samplers over host arrays of doubles (every stored value is exactly a float or a half)
// ---------------------------------------------------------------------------------------------
inline float round_to_half(float v) {      // IEEE binary16 round-to-nearest-even of a binary32, returned as binary32
    if (std::isnan(v)) return v;
    const float av = ::fabsf(v);
    if (av >= 65520.f) return std::copysign(INFINITY, v);
    if (av < 6.103515625e-05f) {           // subnormal halves: multiples of 2^-24
        const float q = ::nearbyintf(av * 16777216.f);
        return std::copysign(q / 16777216.f, v);
    }
    int e;
    const float m = ::frexpf(av, &e);      // av = m 2^e, m in [0.5, 1): 11 significant bits = m 2^11 rounded
    const float q = ::nearbyintf(m * 2048.f);
    return std::copysign(::ldexpf(q, e - 11), v);
}

struct Image {                             // storage: [z][y][x][4] doubles
    double* p = nullptr;
    int w = 0, h = 0, d = 1;
    bool half = false;                     // RGBA16F view: imageStore rounds to binary16
    double* at(int x, int y, int z) const { return p + (((size_t)z * h + y) * w + x) * 4; }
    vec4 texel(int x, int y, int z) const { const double* t = at(x, y, z); return vec4((float)t[0], (float)t[1], (float)t[2], (float)t[3]); }
};
typedef Image image2D;
typedef Image image3D;
typedef Image sampler2D;
typedef Image sampler3D;
typedef Image subpassInput;

inline void imageStore(const Image& im, const ivec2& c, const vec4& v) {
    double* t = im.at(c.x, c.y, 0);
    const float f[4] = {v.x, v.y, v.z, v.w};
    for (int i = 0; i < 4; ++i) t[i] = im.half ? (double)round_to_half(f[i]) : (double)f[i];
}
inline void imageStore(const Image& im, const ivec3& c, const vec4& v) {
    double* t = im.at(c.x, c.y, c.z);
    const float f[4] = {v.x, v.y, v.z, v.w};
    for (int i = 0; i < 4; ++i) t[i] = im.half ? (double)round_to_half(f[i]) : (double)f[i];
}
inline vec4 imageLoad(const Image& im, const ivec2& c) { return im.texel(c.x, c.y, 0); }
inline vec4 imageLoad(const Image& im, const ivec3& c) { return im.texel(c.x, c.y, c.z); }

inline void tex_axis(float u, int n, int& i0, int& i1, float& f) {
    const float t = u * (float)n - 0.5f;
    const float fl = floor(t);
    f = t - fl;
    const float lo = min(max(fl, -1.f), (float)n);     // clamp before the integer cast
    const int i = (int)lo;
    i0 = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
    i1 = i + 1 < 0 ? 0 : (i + 1 > n - 1 ? n - 1 : i + 1);
}
inline vec4 texture(const Image& s, const vec2& uv) {
    int x0, x1, y0, y1; float fx, fy;
    tex_axis(uv.x, s.w, x0, x1, fx); tex_axis(uv.y, s.h, y0, y1, fy);
    const vec4 a = s.texel(x0, y0, 0) * (1.f - fx) + s.texel(x1, y0, 0) * fx;
    const vec4 b = s.texel(x0, y1, 0) * (1.f - fx) + s.texel(x1, y1, 0) * fx;
    return a * (1.f - fy) + b * fy;
}
inline vec4 texture(const Image& s, const vec3& uvw) {
    int x0, x1, y0, y1, z0, z1; float fx, fy, fz;
    tex_axis(uvw.x, s.w, x0, x1, fx); tex_axis(uvw.y, s.h, y0, y1, fy); tex_axis(uvw.z, s.d, z0, z1, fz);
    const vec4 a = s.texel(x0, y0, z0) * (1.f - fx) + s.texel(x1, y0, z0) * fx;
    const vec4 b = s.texel(x0, y1, z0) * (1.f - fx) + s.texel(x1, y1, z0) * fx;
    const vec4 c = s.texel(x0, y0, z1) * (1.f - fx) + s.texel(x1, y0, z1) * fx;
    const vec4 e = s.texel(x0, y1, z1) * (1.f - fx) + s.texel(x1, y1, z1) * fx;
    const vec4 ab = a * (1.f - fy) + b * fy;
    const vec4 ce = c * (1.f - fy) + e * fy;
    return ab * (1.f - fz) + ce * fz;
}

// the invocation a shader's main() runs for (gl_GlobalInvocationID, or the pixel of the fragment shader's attachment)
extern thread_local uvec3 gl_GlobalInvocationID;
extern thread_local ivec2 frag_pixel;
inline vec4 subpassLoad(const Image& im) { return im.texel(frag_pixel.x, frag_pixel.y, 0); }

}  // namespace glsl
