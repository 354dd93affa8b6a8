// post_core.cuh -- per-pixel bodies of the Reprojection compute pass (reprojection.comp:193-244), the god-ray
// post pass (postProcess_GodRays.frag:66-150) and the tone-map post pass (postProcess_ToneMap.frag:68-84).
#pragma once

#include "mt_params.h"

// ---- Reprojection ------------------------------------------------------------------------------------------------
// Returns the ten clamped tap positions (linear index y*W + x) of pixel (x, y); the caller gathers and averages.
MT_DEVICE void reproject_taps(const ReprojParams& P, const RayBasis& B, int x, int y, int taps[10])
{
    const float fw = (float)P.W, fh = (float)P.H;
    float u = (float)x / fw;
    float v = (float)y / fh;  // no y flip here (reprojection.comp:200-201)
    // getJitterOffset of THIS shader: index >= 4 re-reads haltonSeq1/2 (reprojection.comp:83-88)
    int hj = (P.tm.frameCountMod16 >> 1) & 3;
    float jx = P.tm.halton[hj] / fw;
    float jy = P.tm.halton[4 + hj] / fh;
    f3 eye = mk3(-P.cam.eye[0], -P.cam.eye[1], -P.cam.eye[2]);
    f3 dir = cast_ray_dir(P.cam, B, eye, u, v, jx, jy);
    f3 ec = mk3(eye.x, -MT_EARTH_RADIUS, eye.z);
    ShellHit hit = ray_shell(eye, dir, ec, MT_R_INNER);
    const float* m = P.camOld.view;
    f3 p = hit.point;
    f3 q = mk3(((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12] * 1.0f,
               ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13] * 1.0f,
               ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14] * 1.0f);
    q = norm3(q);
    q = q / (-q.z);
    float old_u = (q.x / P.cam.tanFovBy2[0]) * 0.5f + 0.5f;
    float old_v = (q.y / P.cam.tanFovBy2[1]) * 0.5f + 0.5f;
    float mx = old_u - u, my = old_v - v;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const float f = (float)i / 10.0f;
        float ix = rintf((old_u - mx * f) * fw);
        float iy = rintf((old_v - my * f) * fh);
        int cx = min(max(mt_f2i(ix), 0), P.W - 1);
        int cy = min(max(mt_f2i(iy), 0), P.H - 1);
        taps[i] = cy * P.W + cx;
    }
}

// ---- God rays ----------------------------------------------------------------------------------------------------
struct GodRayFrame {  // per-frame values of postProcess_GodRays.frag:74-91
    float blend;      // dot(normalize(sun - eye), camForward); < 0 => the pass writes nothing
    float sunx, suny; // clamped screen-space sun position
};

MT_DEVICE GodRayFrame godray_frame(const CamU& cam)
{
    GodRayFrame g;
    f3 toSun = norm3(mk3(0.0f, 1.0f, 0.0f) - mk3(cam.eye[0], cam.eye[1], cam.eye[2]));
    f3 fwd = neg(norm3(mk3(cam.view[2], cam.view[6], cam.view[10])));
    g.blend = dot3(toSun, fwd);
    // (proj * view) * vec4(sun, 1): the matrix product first, then the vector (GLSL is left-associative)
    float ndc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float pv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
            pv[c] = ((cam.proj[0 * 4 + r] * cam.view[c * 4 + 0] + cam.proj[1 * 4 + r] * cam.view[c * 4 + 1]) +
                     cam.proj[2 * 4 + r] * cam.view[c * 4 + 2]) + cam.proj[3 * 4 + r] * cam.view[c * 4 + 3];
        ndc[r] = ((pv[0] * 0.0f + pv[1] * 1.0f) + pv[2] * 0.0f) + pv[3] * 1.0f;
    }
    g.sunx = sat1((ndc[0] + 1.0f) / 2.0f);
    g.suny = sat1((ndc[1] + 1.0f) / 2.0f);
    return g;
}

// dot(texel, 1/bitEnc): postProcess_GodRays.frag:36-43.
MT_DEVICE float mask_texel_decode(F4 t)
{
    return ((t.x * (1.0f / 1.0f) + t.y * (1.0f / 255.0f)) + t.z * (1.0f / 65025.0f)) + t.w * (1.0f / 16581375.0f);
}
// CLAMP_TO_BORDER, VK_BORDER_COLOR_INT_OPAQUE_BLACK = (0,0,0,1) (Texture2D.cpp:75, Image.cpp:322): decodes to 1/16581375.
#define MT_MASK_BORDER_DECODED (((0.0f * (1.0f / 1.0f) + 0.0f * (1.0f / 255.0f)) + 0.0f * (1.0f / 65025.0f)) + 1.0f * (1.0f / 16581375.0f))

// extract32fFromRGBA8f (postProcess_GodRays.frag:39-43).  The shader filters the four ENCODED channels bilinearly and
// then takes the dot product with 1/bitEnc; both steps are linear, so the kernel decodes each texel once
// (mask_decode_kernel, 16 B read -> 4 B written per pixel) and filters the decoded scalar: 16 instead of 64 bytes
// and 4 instead of 16 multiply-adds per tap, 100 taps per pixel.  The two orders agree to rounding (~1e-7 relative on
// a term that is itself <= 2.5 % of the pixel); no decision depends on it.  `dec` is the (W+2) x (H+2) decoded image
// whose one-texel ring holds the border value, so the taps need no bounds tests.
MT_DEVICE float mask_decode(const float* dec, int W, int H, float s, float t)
{
    float u = s * (float)W - 0.5f, v = t * (float)H - 0.5f;
    int x0 = mt_floor2i(u), y0 = mt_floor2i(v);
    float ax = u - (float)x0, ay = v - (float)y0;
    // uv stays inside [0,1] (the march runs from the pixel towards the clamped sun position), so x0 in [-1, W-1]
    x0 = min(max(x0, -1), W - 1);
    y0 = min(max(y0, -1), H - 1);
    const float* p = dec + (size_t)(y0 + 1) * (size_t)(W + 2) + (size_t)(x0 + 1);
    float a = MT_LDG(p), b = MT_LDG(p + 1), c = MT_LDG(p + (W + 2)), d = MT_LDG(p + (W + 3));
    float w00 = (1.0f - ax) * (1.0f - ay), w01 = ax * (1.0f - ay), w10 = (1.0f - ax) * ay, w11 = ax * ay;
    return fmaf(w11, d, fmaf(w10, c, fmaf(w01, b, w00 * a)));
}

// The radial accumulation of one fragment; returns the colour to ADD to the HDR pixel (already * blend).
MT_DEVICE F4 godray_pixel(const GodRayParams& P, const GodRayFrame& G, int x, int y)
{
    float u = ((float)x + 0.5f) / (float)P.W;
    float v = ((float)y + 0.5f) / (float)P.H;
    const float du = ((u - G.sunx) / 100.0f) * 1.0f;
    const float dv = ((v - G.suny) / 100.0f) * 1.0f;
    float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f;
    for (int i = 0; i < 100; ++i) {
        float a = mask_decode(P.decoded, P.W, P.H, u, v);
        acc0 += (P.lightColor[0] * a) * (1.0f * 0.001f);
        acc1 += (P.lightColor[1] * a) * (1.0f * 0.001f);
        acc2 += (P.lightColor[2] * a) * (1.0f * 0.001f);
        u -= du;
        v -= dv;
    }
    F4 o;
    o.x = (acc0 * 1.0f) * G.blend;
    o.y = (acc1 * 1.0f) * G.blend;
    o.z = (acc2 * 1.0f) * G.blend;
    o.w = 1.0f * G.blend;
    return o;
}

// ---- Tone map ----------------------------------------------------------------------------------------------------
MT_DEVICE float uncharted2(float x)
{
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F;
}
MT_DEVICE unsigned wang_hash(unsigned u, unsigned v, unsigned s)
{
    unsigned seed = (u * 1664525u + v) + s;
    seed = (seed ^ 61u) ^ (seed >> 16u);
    seed *= 9u;
    seed = seed ^ (seed >> 4u);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15u);
    return seed;
}
MT_DEVICE unsigned unorm8(float v)
{
#if defined(MT_HOSTSIM)
    if (v != v) return 0u;
    return (unsigned)rintf(sat1(v) * 255.0f);
#else
    return __float2uint_rn(__saturatef(v) * 255.0f);  // NaN -> 0 in both steps
#endif
}
MT_DEVICE unsigned tonemap_pixel(const ToneMapParams& P, F4 in, int x, int y)
{
    const float whitemap = 1.0f / uncharted2(100.0f);
    const float invGamma = 1.0f / 2.2f;
    float noise = ((float)wang_hash((unsigned)x, (unsigned)y, P.seed) * (1.0f / 4294967296.0f)) * 0.01f;
    float r = MT_POWF(uncharted2(2.5f * in.x) * whitemap, invGamma) + noise;
    float g = MT_POWF(uncharted2(2.5f * in.y) * whitemap, invGamma) + noise;
    float b = MT_POWF(uncharted2(2.5f * in.z) * whitemap, invGamma) + noise;
    return unorm8(r) | (unorm8(g) << 8) | (unorm8(b) << 16) | (255u << 24);
}
