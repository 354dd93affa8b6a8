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

MT_DEVICE F4 mask_fetch(const F4* mask, int W, int H, int x, int y)
{
    F4 t;
    if (x < 0 || y < 0 || x >= W || y >= H) {  // CLAMP_TO_BORDER, opaque black (Texture2D.cpp:75, Image.cpp:322)
        t.x = t.y = t.z = 0.0f; t.w = 1.0f;
        return t;
    }
#if defined(MT_HOSTSIM)
    return mask[(size_t)y * W + x];
#else
    float4 v = __ldg(reinterpret_cast<const float4*>(mask) + ((size_t)y * W + x));
    t.x = v.x; t.y = v.y; t.z = v.z; t.w = v.w;
    return t;
#endif
}

// extract32fFromRGBA8f (postProcess_GodRays.frag:39-43): bilinear fetch of the ENCODED channels, then decode.
MT_DEVICE float mask_decode(const F4* mask, int W, int H, float s, float t)
{
    float u = s * (float)W - 0.5f, v = t * (float)H - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float ax = u - fu, ay = v - fv;
    int x0 = mt_f2i(fu), y0 = mt_f2i(fv);
    float w00 = (1.0f - ax) * (1.0f - ay), w01 = ax * (1.0f - ay), w10 = (1.0f - ax) * ay, w11 = ax * ay;
    F4 a = mask_fetch(mask, W, H, x0, y0), b = mask_fetch(mask, W, H, x0 + 1, y0);
    F4 c = mask_fetch(mask, W, H, x0, y0 + 1), d = mask_fetch(mask, W, H, x0 + 1, y0 + 1);
    float r0 = fmaf(w11, d.x, fmaf(w10, c.x, fmaf(w01, b.x, w00 * a.x)));
    float r1 = fmaf(w11, d.y, fmaf(w10, c.y, fmaf(w01, b.y, w00 * a.y)));
    float r2 = fmaf(w11, d.z, fmaf(w10, c.z, fmaf(w01, b.z, w00 * a.z)));
    float r3 = fmaf(w11, d.w, fmaf(w10, c.w, fmaf(w01, b.w, w00 * a.w)));
    return ((r0 * (1.0f / 1.0f) + r1 * (1.0f / 255.0f)) + r2 * (1.0f / 65025.0f)) + r3 * (1.0f / 16581375.0f);
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
        float a = mask_decode(P.mask, P.W, P.H, u, v);
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
