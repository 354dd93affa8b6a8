// TEST INFRASTRUCTURE -- what a Vulkan device supplies to the reference's shaders when oracle/glsl2cpp.py has turned
// their text into C++: GLSL's vector types and built-ins (glm, the reference's own vendored copy), storage images,
// samplers, gl_GlobalInvocationID.  Included by oracle/refshader_*.cpp only; never part of the product.
//
// Two things here are NOT the reference's code and cannot be: (1) the fixed-function sampler -- LINEAR / REPEAT,
// UNORM8 -> float -- which is the oracle's own (mto_sample3d / mto_sample2d, exported by libmeteoros_oracle.so), so a
// difference between this build and the oracle isolates the shader arithmetic; (2) the precision of GLSL built-ins,
// which Vulkan leaves to the driver.  By default the built-ins are glm's (an independent implementation of the GLSL
// spec).  -DMTREF_CANONICAL_BUILTINS renames the handful where glm's formula differs from the canonical semantics
// DESIGN.md section 2 fixes (the GLSL spec's own formulas) to the definitions below; with it the oracle must match
// this build bit for bit, without it to rounding.
#pragma once
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>

#include <cmath>
#include <cstdint>
#include <cstring>

using namespace glm;

extern "C" {
void mto_sample3d(const uint8_t* vol, int W, int H, int D, float s, float t, float r, float out[4]);
void mto_sample2d(const uint8_t* img, int W, int H, float s, float t, float out[4]);
void mto_sample2d_f32_border(const float* img, int W, int H, float s, float t, float out[4]);
void mto_sample2d_unorm8_border(const uint8_t* img, int W, int H, float s, float t, float out[4]);
void mto_load_unorm8(const uint8_t* img, int W, int H, int x, int y, float out[4]);
uint8_t mto_to_unorm8(float v);
}

// Storage image.  rgba32f: `texels` [y][x][4] floats, loads and stores hit the same memory.  rgba8: loads come from
// `ldr_in`, stores go to `ldr_out` (UNORM8 conversion both ways) -- the TXAA fragment shader reads its neighbourhood
// from the image it is writing (a data race on a GPU); canonical: it sees the image as it was before the pass.
struct image2D { float* texels; int w, h; const uint8_t* ldr_in; uint8_t* ldr_out; };
struct sampler3D { const uint8_t* texels; int w, h, d; };  // RGBA8_UNORM, LINEAR, REPEAT (Texture3D.cpp:92-134)
enum SamplerKind {
    UNORM8_REPEAT,   // curl / weather: RGBA8_UNORM, LINEAR, REPEAT (Image.cpp:305-347)
    F32_BORDER,      // god-ray mask: float image, LINEAR, CLAMP_TO_BORDER opaque black (Texture2D.cpp:75)
    F32_TEXEL,       // tone-map input: sampled at texel centres -> the texel itself
    UNORM8_BORDER,   // TXAA history: RGBA8 image, LINEAR, CLAMP_TO_BORDER opaque black
};
struct sampler2D { const void* texels; int w, h; SamplerKind kind; };

inline ivec2 imageSize(const image2D& im) { return ivec2(im.w, im.h); }
inline void imageStore(image2D& im, ivec2 p, vec4 v)
{
    if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return;  // out-of-bounds stores are discarded
    const size_t k = ((size_t)p.y * im.w + p.x) * 4;
    if (im.ldr_out)
        for (int c = 0; c < 4; ++c) im.ldr_out[k + c] = mto_to_unorm8(v[c]);
    if (!im.texels) return;   // rgba8 image: optional unquantised copy of what was stored, for the tests
    float* t = im.texels + k;
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
}
inline vec4 imageLoad(const image2D& im, ivec2 p)
{
    if (im.ldr_in) {
        float o[4];
        mto_load_unorm8(im.ldr_in, im.w, im.h, p.x, p.y, o);   // outside the image: 0
        return vec4(o[0], o[1], o[2], o[3]);
    }
    if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return vec4(0.0f);
    const float* t = im.texels + ((size_t)p.y * im.w + p.x) * 4;
    return vec4(t[0], t[1], t[2], t[3]);
}
inline vec4 texture(const sampler3D& s, vec3 p)
{
    float o[4];
    mto_sample3d(s.texels, s.w, s.h, s.d, p.x, p.y, p.z, o);
    return vec4(o[0], o[1], o[2], o[3]);
}
inline vec4 texture(const sampler2D& s, vec2 p)
{
    float o[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
    switch (s.kind) {
    case UNORM8_REPEAT: mto_sample2d((const uint8_t*)s.texels, s.w, s.h, p.x, p.y, o); break;
    case F32_BORDER: mto_sample2d_f32_border((const float*)s.texels, s.w, s.h, p.x, p.y, o); break;
    case UNORM8_BORDER: mto_sample2d_unorm8_border((const uint8_t*)s.texels, s.w, s.h, p.x, p.y, o); break;
    case F32_TEXEL: {
        int x = (int)std::floor(p.x * (float)s.w), y = (int)std::floor(p.y * (float)s.h);
        x = x < 0 ? 0 : (x >= s.w ? s.w - 1 : x); y = y < 0 ? 0 : (y >= s.h ? s.h - 1 : y);
        memcpy(o, (const float*)s.texels + ((size_t)y * s.w + x) * 4, 16);
        break;
    }
    }
    return vec4(o[0], o[1], o[2], o[3]);
}

// GLSL's implicit int -> float conversions, which C++ templates do not perform
inline vec2 operator/(vec2 a, ivec2 b) { return a / vec2(b); }
inline vec2 operator*(vec2 a, ivec2 b) { return a * vec2(b); }
inline float mod(int a, int b) { return glm::mod((float)a, (float)b); }
inline float mod(float a, int b) { return glm::mod(a, (float)b); }

static thread_local uvec3 gl_GlobalInvocationID;

// Fragment stage: the full-screen triangle of postProcess_GenericVertShader.vert:15-16 interpolates
// in_uv = ((x + 0.5) / W, (y + 0.5) / H) at the centre of pixel (x, y).
#define MTREF_FOR_EACH_FRAGMENT(W, H, body)                                              \
    _Pragma("omp parallel for schedule(static)") for (int fy = 0; fy < (H); ++fy)      \
        for (int fx = 0; fx < (W); ++fx) {                                              \
            in_uv = vec2(((float)fx + 0.5f) / (float)(W), ((float)fy + 0.5f) / (float)(H)); \
            body;                                                                       \
        }

#ifdef MTREF_CANONICAL_BUILTINS
// GLSL 4.50 spec section 8.3: mix(x, y, a) = x * (1 - a) + y * a   (glm: x + a * (y - x))
inline float canon_mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 canon_mix(vec3 x, vec3 y, float a) { return x * (1.0f - a) + y * a; }
inline vec4 canon_mix(vec4 x, vec4 y, float a) { return x * (1.0f - a) + y * a; }
// round(): "the fraction 0.5 rounds in a direction chosen by the implementation" -> half to even (glm: away from zero)
inline float canon_round(float x) { return rintf(x); }
inline vec2 canon_round(vec2 v) { return vec2(rintf(v.x), rintf(v.y)); }
// dot(): "x[0]*y[0] + x[1]*y[1] + ..." summed left to right (glm sums a vec4 pairwise: (x + y) + (z + w))
inline float canon_dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float canon_dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float canon_dot(vec4 a, vec4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
// mat4 * vec4: the four products summed left to right, the order a GPU's multiply-add chain gives them
// (glm: (m0*x + m1*y) + (m2*z + m3*w)).  A non-template overload wins over glm's template.
inline vec4 operator*(const mat4& m, const vec4& v) { return ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3] * v.w; }
#define mix canon_mix
#define round canon_round
#define dot canon_dot
#endif

#ifdef MTREF_FRAGMENT_STAGE
// The post shaders compute pixelPos = ivec2(round(dim * in_uv)): dim * in_uv = x + 0.5 is an exact tie, which round()
// may resolve either way (GLSL 4.50 8.3) -- on a device the write lands on pixel x or x + 1 at the driver's whim.
// Canonical resolution = the evident intent, the fragment's own pixel.
inline float frag_round(float x) { return std::floor(x); }
inline vec2 frag_round(vec2 v) { return vec2(std::floor(v.x), std::floor(v.y)); }
#undef round
#define round frag_round
#endif
