// post_core.cuh -- per-pixel bodies of the Reprojection compute pass (reprojection.comp:193-244), the god-ray
// post pass (postProcess_GodRays.frag:66-150) and the tone-map post pass (postProcess_ToneMap.frag:68-84).
#pragma once

#include "mt_params.h"

// ---- Reprojection ------------------------------------------------------------------------------------------------
// ReprojFrame (mt_params.h): the per-frame values of reprojection.comp:203-211, evaluated by the host per dispatch.
MT_HD ReprojFrame reproject_frame(const ReprojParams& P)
{
    ReprojFrame F;
    F.basis = ray_basis(P.cam);
    F.eye = mk3(-P.cam.eye[0], -P.cam.eye[1], -P.cam.eye[2]);
    F.ec = mk3(F.eye.x, -MT_EARTH_RADIUS, F.eye.z);
    F.o = (F.eye - F.ec) / MT_R_INNER;
    F.C = dot3(F.o, F.o) - 1.0f;
    // getJitterOffset of THIS shader: index >= 4 re-reads haltonSeq1/2 (reprojection.comp:83-88)
    const int hj = (P.tm.frameCountMod16 >> 1) & 3;
    F.jx = P.tm.halton[hj] / (float)P.W;
    F.jy = P.tm.halton[4 + hj] / (float)P.H;
    F.uMax = ((float)P.W - 1.0f) / (float)P.W;
    F.vMax = ((float)P.H - 1.0f) / (float)P.H;
    return F;
}

// Returns the ten clamped tap positions (linear index y*W + x, plus MT_TAP_BIAS) of the pixel whose uv is (u, v) = (x/W, y/H).
// Device fast path (MT_REPROJ_FAST): a tap lies between old_uv and uv, so when old_uv is inside [0, (dim-1)/dim] no tap can
// round outside the image and the four integer clamps per tap are dead; ivec2(round(c)) then comes from adding 2^23 (the
// integer lands in the low mantissa bits, round-half-even like F2I.RN) instead of two conversions on the quarter-rate pipe:
// 12 instead of 16 instructions per tap, the same indices.  The sums are scalar adds: a packed add fed by the packed
// multiply would be contracted into one FFMA2 (mt_math.cuh) and round c * dim + 2^23 once instead of twice.  Every index carries
// the bias the fast path leaves in it (0x4B000000, the bits of 2^23); the kernel takes it out of the image's base address.
#ifndef MT_REPROJ_FAST
#define MT_REPROJ_FAST 1
#endif
#if defined(MT_HOSTSIM) || !MT_REPROJ_FAST
#define MT_TAP_BIAS 0
#else
#define MT_TAP_BIAS 0x4B000000
#endif
// old_uv of a pixel: castRay, the inner-shell intersection, the hit point through the previous frame's view matrix and back to
// screen space (reprojection.comp:203-232; postProcess_TXAA.frag does the same with its own uv / jitter).  Seven IEEE divisions
// and three square roots per pixel, each with its range test, branch to a slow path and reconvergence pair.  NICE = the same
// operations through the fast paths only (mt_math.cuh: div_nice / sqrt_nice / nice_rcp -- correctly rounded whenever no operand
// or intermediate leaves the normal range), for frames whose constants the host has checked (mt_post_nice_ok, post_passes.cu:
// orthonormal bases, both eyes inside the inner shell with 100 m to spare, tan(fov/2) in [1e-3, 1e3]): then |p - eye| >= 0.7,
// 2A ~ 2, the discriminant is 0 or >= 1e-18, |q| >= 100 m; the one divisor no frame constant bounds, -q.z, is tested per pixel.
template <bool NICE>
MT_DEVICE void reproject_old_uv(const CamU& cam, const float* m, const RayBasis& basis, f3 eye, f3 ec, f3 o, float C, float u, float v,
                                float jx, float jy, float& old_u, float& old_v)
{
    f3 dir;
    {   // castRay (mt_math.cuh, cast_ray_dir) with the normalisation on the fast path
        const float nx = (u * 2.0f - 1.0f) + jx, ny = (v * 2.0f - 1.0f) + jy;
        const f3 pe = (((eye + basis.look) + basis.right * (nx * cam.tanFovBy2[0])) + basis.up * (ny * cam.tanFovBy2[1])) - eye;
        dir = NICE ? norm3_nice(pe) : norm3(pe);
    }
    // raySphereIntersection (reprojection.comp:147-191) with the pixel-independent terms hoisted; only .point is used
    f3 p = mk3(0.0f, 0.0f, 0.0f);
    {
        const float A = dot3(dir, dir);
        const float B = 2.0f * dot3(dir, o);
        const float disc = B * B - (4.0f * A) * C;
        if (!(disc < 0.0f)) {
            const float sq = NICE ? sqrt_nice(disc) : sqrtf(disc);
            float t = NICE ? div_nice(-B - sq, 2.0f * A) : (-B - sq) / (2.0f * A);
            if (t < 0.0f) t = NICE ? div_nice(-B + sq, 2.0f * A) : (-B + sq) / (2.0f * A);
            if (t >= 0.0f) p = ((o + dir * t) * MT_R_INNER) + ec;
        }
    }
    f3 q = mk3(((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12] * 1.0f,
               ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13] * 1.0f,
               ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14] * 1.0f);
    q = NICE ? norm3_nice(q) : norm3(q);
    if (NICE && fabsf(q.z) >= 1e-30f) {  // (else: zero, denormal or NaN divisor -- the IEEE sequence below)
        const float nz = -q.z, rz = nice_rcp(nz);
        old_u = div_nice(div_nice_r(q.x, nz, rz), cam.tanFovBy2[0]) * 0.5f + 0.5f;
        old_v = div_nice(div_nice_r(q.y, nz, rz), cam.tanFovBy2[1]) * 0.5f + 0.5f;
        return;
    }
    q = q / (-q.z);
    old_u = (q.x / cam.tanFovBy2[0]) * 0.5f + 0.5f;
    old_v = (q.y / cam.tanFovBy2[1]) * 0.5f + 0.5f;
}

template <bool NICE = false>
MT_DEVICE void reproject_taps(const ReprojParams& P, const ReprojFrame& F, float u, float v, int taps[10])
{
    const float fw = (float)P.W, fh = (float)P.H;  // no y flip here (reprojection.comp:200-201)
    float old_u, old_v;
    reproject_old_uv<NICE>(P.cam, P.camOld.view, F.basis, F.eye, F.ec, F.o, F.C, u, v, F.jx, F.jy, old_u, old_v);
    const P2 mv = pk2(old_u - u, old_v - v), dim = pk2(fw, fh);
#if !defined(MT_HOSTSIM) && MT_REPROJ_FAST
    // (dim - 1) / dim rounded: old * dim <= dim - 1 + 5e-4 and the taps' own roundings move them by < 1e-3 of a pixel,
    // so every tap rounds into [0, dim - 1]; NaN fails the comparisons
    if (old_u >= 0.0f && old_u <= F.uMax && old_v >= 0.0f && old_v <= F.vMax) {
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const float f = (float)i / 10.0f;
            const P2 b = mul2(mv, bc2(f));
            const P2 c = mul2(pk2(old_u - lo2(b), old_v - hi2(b)), dim);
            const unsigned tx = __float_as_uint(lo2(c) + 8388608.0f), ty = __float_as_uint(hi2(c) + 8388608.0f);
            taps[i] = (int)((ty & 0x007fffffu) * (unsigned)P.W + tx);  // cy * W + cx + MT_TAP_BIAS
        }
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const float f = (float)i / 10.0f;
        // (old_uv - motion * (i/10)) * dim; the subtraction is scalar (a mul2 feeding a sub2 would be contracted)
        const P2 b = mul2(mv, bc2(f));
        const P2 c = mul2(pk2(old_u - lo2(b), old_v - hi2(b)), dim);
        const int cx = min(max(mt_round2i(lo2(c)), 0), P.W - 1);
        const int cy = min(max(mt_round2i(hi2(c)), 0), P.H - 1);
        taps[i] = cy * P.W + cx + MT_TAP_BIAS;
    }
}

// ---- God rays ----------------------------------------------------------------------------------------------------
MT_HD GodRayFrame godray_frame(const CamU& cam)
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
// (mask_decode_kernel) and filters the decoded scalar: 100 taps per pixel.  The god-ray term is radiance only -- no decision
// depends on it -- and at most 2.5 % of the pixel, so it is held to "equal to rounding" (tests: <= 2e-5 of the term), not to
// bit-exactness; what is kept exact is the tap POSITION sequence in uv (uv -= delta, 100 roundings), whose drift would
// otherwise add up along the march.  The texel coordinate of a tap, uv * dim - 0.5, is formed in one rounding on the device
// (MT_GODRAY_FMA_POS: within half an ulp, 3e-5 texel at 1080p, of the shader's two-rounding form; the filter is continuous).
// `dec` is the (W+2) x (H+2) decoded image whose one-texel ring holds the border value, so the taps need no bounds tests;
// each element is the PAIR (d(x, y), d(x+1, y)), so a tap is two 8-byte loads.  Per tap (device): the floor of both
// coordinates comes from ONE packed add in round-down mode against 1.5 * 2^23 (the integer lands in the low mantissa bits,
// the float floor is the same value minus the constant) -- no F2I / I2F on the quarter-rate pipe; the filter is
// a + ay (c - a) on both columns at once, then in x: ~19 instructions per tap instead of 31 (profiles/r2_passes_1080p.md).
#ifndef MT_GODRAY_WIDE
#define MT_GODRAY_WIDE 0
#endif
#ifndef MT_GODRAY_FMA_POS
#define MT_GODRAY_FMA_POS 1  /* 1080p: 166.4 us with 0, 154.1 us with 1 (profiles/r2c_ab.md) */
#endif
#ifndef MT_GODRAY_SCALAR
#define MT_GODRAY_SCALAR 0
#endif
#if MT_GODRAY_SCALAR && !defined(MT_HOSTSIM)
// A/B: the same tap with scalar instructions only (is the FMA pipe, which executes an fp32x2 instruction in two passes, the limit?)
template <int K>
MT_DEVICE float mask_decode(const GodRayParams& P, P2 st)
{
    const float mx = lo2(st) * (float)P.W, my = hi2(st) * (float)P.H;
    const float ux = mx - 0.5f, uy = my - 0.5f;
    float tx, ty;
    asm("add.rm.f32 %0, %1, %2;" : "=f"(tx) : "f"(ux), "f"(MT_FLOOR_MAGIC));
    asm("add.rm.f32 %0, %1, %2;" : "=f"(ty) : "f"(uy), "f"(MT_FLOOR_MAGIC));
    const float ax = ux - (tx - MT_FLOOR_MAGIC), ay = uy - (ty - MT_FLOOR_MAGIC);
    const int i0 = (__float_as_int(ty) - MT_FLOOR_MAGIC_BITS) * (P.W + 2) + __float_as_int(tx);
    const float2 top = MT_LDG(P.tapRow0 + i0), bot = MT_LDG(P.tapRow1 + i0);
    const float l = fmaf(ay, bot.x - top.x, top.x), r = fmaf(ay, bot.y - top.y, top.y);
    return fmaf(ax, r - l, l);
}
#else
template <int K>
MT_DEVICE float mask_decode(const GodRayParams& P, P2 st)
{
    const float2* dec = P.decoded;
    const int W = P.W, H = P.H;
#if MT_GODRAY_FMA_POS && !defined(MT_HOSTSIM)
    // u * dim - 0.5 in ONE rounding (FFMA2) instead of two (FMUL2 + two scalar adds): the position moves by at most half an ulp of
    // u * dim (3e-5 texel at 1080p) against the shader's two-rounding form, which the bilinear filter -- continuous across cell
    // boundaries -- turns into < 4e-5 of a tap on a full-contrast edge; the pass is radiance only and held to 2e-5 of its term.
    const P2 um = fma2(st, pk2((float)W, (float)H), bc2(-0.5f));
    const float ux = lo2(um), uy = hi2(um);
#else
    const P2 m = mul2(st, pk2((float)W, (float)H));
    const float ux = lo2(m) - 0.5f, uy = hi2(m) - 0.5f;  // scalar: mul2 -> sub2 would be contracted (mt_math.cuh)
#endif
    // No clamp: the march runs from the pixel centre towards the sun position, which main() clamps to [0,1]
    // (postProcess_GodRays.frag:90), so u*W - 0.5 lies in [-0.5, W - 0.5] and floor() in [-1, W-1] -- the ring.  The 100
    // roundings of `uv -= delta` move u*W by < 0.05 texel even at W = 7680, against a margin of 0.5.
    const int pitch = P.pitch;
#if defined(MT_HOSTSIM)
    const float fx = floorf(ux), fy = floorf(uy);
    const int x0 = (int)fx, y0 = (int)fy;
    const P2 fl = pk2(fx, fy);
    const float2* texel00 = dec + pitch + 1;                      // image texel (0, 0) inside the ring
    const int i0 = y0 * pitch + x0;
#else
    P2 t;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pk2(ux, uy)), "l"(bc2(MT_FLOOR_MAGIC)));
    const P2 fl = sub2(t, bc2(MT_FLOOR_MAGIC));                   // exact
    if (K > 0) {
        // Power-of-two pitch 2^K (K >= 10): the bits of t.y are 0x4B400000 + floor(y), and 0x4B400000 << K vanishes mod 2^32, so
        // (t.y << K) + t.x = (floor(y) << K) + floor(x) + 0x4B400000 in ONE LEA; times eight (mod 2^32) that is the texel's byte
        // offset plus the constant 0x5A000000, which the host has subtracted from the base.  The second row is an immediate
        // offset of the same address: one address computation on the ALU pipe (LEA + carry) for both loads, where the generic
        // path spends IADD3 + IMAD + 2 x IMAD.WIDE -- the latter three on the FMA-heavy pipe, the kernel's busiest (70 %,
        // profiles/r2_passes_1080p.md) because the packed fp32x2 instructions live there as well.
        unsigned tx, ty;
        asm("mov.b64 {%0, %1}, %2;" : "=r"(tx), "=r"(ty) : "l"(t));
        const unsigned i0 = (ty << K) + tx;
#if MT_GODRAY_WIDE   // A/B: one IMAD.WIDE (FMA-heavy pipe) instead of LEA + carry (ALU pipe)
        const char* p = P.tapBaseWide + (size_t)i0 * 8u;
#else
        const char* p = P.tapBase + (i0 << 3);
#endif
        const float2 top = MT_LDG(reinterpret_cast<const float2*>(p)), bot = MT_LDG(reinterpret_cast<const float2*>(p + ((size_t)8 << K)));
        const P2 a1 = sub2(pk2(ux, uy), fl);                      // (ax, ay)
        const P2 t2 = pk2(top.x, top.y), b2 = pk2(bot.x, bot.y);
        const P2 lr = fma2(bc2(hi2(a1)), sub2(b2, t2), t2);
        return fmaf(lo2(a1), hi2(lr) - lo2(lr), lo2(lr));
    }
    const int xb = (int)(unsigned)t, y0 = (int)(unsigned)(t >> 32) - MT_FLOOR_MAGIC_BITS;
    // the host passes the two row bases (image texel (0, 0) and (0, 1) inside the ring, the x bias of the magic constant
    // folded in): two opaque loop-invariant pointers, one IMAD.WIDE per load instead of 64-bit pointer arithmetic
    (void)dec;
    const float2* texel00 = P.tapRow0;
    const int i0 = y0 * pitch + xb;                               // < 2^31: xb <= 0x4B400000 + W, y0 * pitch <= 33e6 at 8K
#endif
    const P2 a1 = sub2(pk2(ux, uy), fl);                          // (ax, ay)
#if defined(MT_HOSTSIM)
    const float2* row1 = texel00 + pitch;
#else
    const float2* row1 = P.tapRow1;
#endif
    const float2 top = MT_LDG(texel00 + i0), bot = MT_LDG(row1 + i0);
    const P2 t2 = pk2(top.x, top.y), b2 = pk2(bot.x, bot.y);      // the pairs as loaded: (x0, x0+1) of each row
    const P2 lr = fma2(bc2(hi2(a1)), sub2(b2, t2), t2);           // both columns filtered in y: a + ay (c - a)
    return fmaf(lo2(a1), hi2(lr) - lo2(lr), lo2(lr));             // then in x
}
#endif

// The radial accumulation of one fragment; returns the colour to ADD to the HDR pixel (already * blend).
template <int K>
MT_DEVICE F4 godray_pixel_uv(const GodRayParams& P, const GodRayFrame& G, float u, float v);
template <int K = 0>
MT_DEVICE F4 godray_pixel(const GodRayParams& P, const GodRayFrame& G, int x, int y)
{
    return godray_pixel_uv<K>(P, G, ((float)x + 0.5f) / (float)P.W, ((float)y + 0.5f) / (float)P.H);
}
// (u, v) = the fragment's uv, ((x + .5) / W, (y + .5) / H): the kernel reads it from the context's uv table
template <int K>
MT_DEVICE F4 godray_pixel_uv(const GodRayParams& P, const GodRayFrame& G, float u, float v)
{
    const float du = ((u - G.sunx) / 100.0f) * 1.0f;
    const float dv = ((v - G.suny) / 100.0f) * 1.0f;
    P2 uv = pk2(u, v);
    const P2 duv = pk2(du, dv);
    // sum_i (lightColor * a_i) * 0.001 is accumulated as lightColor * 0.001 * sum_i a_i: the same real number, one add
    // per tap instead of seven instructions; the reordering moves the result by ~1e-7 of a term that is at most 2.5 %
    // of the pixel (radiance only, no decision depends on it).
    float sum = 0.0f;
#ifndef MT_GODRAY_UNROLL
#define MT_GODRAY_UNROLL 4
#endif
#define MT_GR_PRAGMA_(x) _Pragma(#x)
#define MT_GR_UNROLL_(n) MT_GR_PRAGMA_(unroll n)
    MT_GR_UNROLL_(MT_GODRAY_UNROLL)
    for (int i = 0; i < 100; ++i) {
        sum += mask_decode<K>(P, uv);
        uv = sub2(uv, duv);
    }
    F4 o;
    o.x = ((P.lightColor[0] * sum) * (1.0f * 0.001f)) * G.blend;
    o.y = ((P.lightColor[1] * sum) * (1.0f * 0.001f)) * G.blend;
    o.z = ((P.lightColor[2] * sum) * (1.0f * 0.001f)) * G.blend;
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

// ---- TXAA (postProcess_TXAA.frag:171-270) ---------------------------------------------------------------
MT_HD TxaaFrame txaa_frame(const TxaaParams& P)
{
    TxaaFrame F;
    F.basis = ray_basis(P.cam);
    F.eye = mk3(-P.cam.eye[0], -P.cam.eye[1], -P.cam.eye[2]);
    F.ec = mk3(F.eye.x, -MT_EARTH_RADIUS, F.eye.z);
    F.o = (F.eye - F.ec) / MT_R_INNER;
    F.C = dot3(F.o, F.o) - 1.0f;
    const int hj = P.tm.frameCountMod16 >> 1;
    const int hx = hj < 4 ? hj : hj + 4;
    F.jx = P.tm.halton[hx] / (float)P.W;
    F.jy = P.tm.halton[hx + 4] / (float)P.H;
    return F;
}

struct C4 {
    float c[4];
};
// MT_TXAA_FAST: bytes become floats without the quarter-rate conversion pipe -- __byte_perm builds 0x4B0000bb = 2^23 + b,
// subtracting 2^23 is exact -- and interior pixels skip the 13 bounds tests.  Same values, same order of operations:
// bit-identical by construction (and by test).  1080p: 70.1 -> 63.9 us (profiles/r1_ab.md).
#ifndef MT_TXAA_FAST
#define MT_TXAA_FAST 1
#endif
MT_DEVICE C4 ldr_unpack(uint32_t t)
{
    C4 r;
#if MT_TXAA_FAST && !defined(MT_HOSTSIM)
    r.c[0] = (__uint_as_float(__byte_perm(t, 0x4B000000u, 0x7650)) - 8388608.0f) * (1.0f / 255.0f);
    r.c[1] = (__uint_as_float(__byte_perm(t, 0x4B000000u, 0x7651)) - 8388608.0f) * (1.0f / 255.0f);
    r.c[2] = (__uint_as_float(__byte_perm(t, 0x4B000000u, 0x7652)) - 8388608.0f) * (1.0f / 255.0f);
    r.c[3] = (__uint_as_float(__byte_perm(t, 0x4B000000u, 0x7653)) - 8388608.0f) * (1.0f / 255.0f);
    return r;
#endif
    r.c[0] = (float)(t & 0xffu) * (1.0f / 255.0f);
    r.c[1] = (float)((t >> 8) & 0xffu) * (1.0f / 255.0f);
    r.c[2] = (float)((t >> 16) & 0xffu) * (1.0f / 255.0f);
    r.c[3] = (float)(t >> 24) * (1.0f / 255.0f);
    return r;
}
MT_DEVICE C4 ldr_load(const uint32_t* img, int W, int H, int x, int y)  // imageLoad: zero outside the image
{
    if (x < 0 || y < 0 || x >= W || y >= H) {
        C4 z;
        z.c[0] = z.c[1] = z.c[2] = z.c[3] = 0.0f;
        return z;
    }
    return ldr_unpack(MT_LDG(img + ((unsigned)y * (unsigned)W + (unsigned)x)));
}
MT_DEVICE C4 ldr_border_texel(const uint32_t* img, int W, int H, int x, int y)  // sampler: CLAMP_TO_BORDER, (0,0,0,1)
{
    if (x < 0 || y < 0 || x >= W || y >= H) {
        C4 z;
        z.c[0] = z.c[1] = z.c[2] = 0.0f; z.c[3] = 1.0f;
        return z;
    }
    return ldr_unpack(MT_LDG(img + ((unsigned)y * (unsigned)W + (unsigned)x)));
}

// One fragment of the TXAA pass; returns the packed RGBA8 result.
template <bool NICE = false>
MT_DEVICE uint32_t txaa_pixel(const TxaaParams& P, const TxaaFrame& F, int x, int y, float u, float v)
{
    // (u, v) = in_uv = ((x + .5) / W, (y + .5) / H), from the context's uv table
    float old_u, old_v;
    reproject_old_uv<NICE>(P.cam, P.camOld.view, F.basis, F.eye, F.ec, F.o, F.C, u, v, F.jx, F.jy, old_u, old_v);

    // 3x3 neighbourhood of the tone-mapped frame (order: tl tc tr ml mc mr bl bc br)
    C4 n[9];
    if (MT_TXAA_FAST && x >= 1 && y >= 1 && x < P.W - 1 && y < P.H - 1) {  // interior: no bounds tests
        const uint32_t* row = P.cur + ((unsigned)(y - 1) * (unsigned)P.W + (unsigned)(x - 1));
#pragma unroll
        for (int k = 0; k < 9; ++k) n[k] = ldr_unpack(MT_LDG(row + (unsigned)(k / 3) * (unsigned)P.W + (unsigned)(k % 3)));
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) n[k] = ldr_load(P.cur, P.W, P.H, x + (k % 3) - 1, y + (k / 3) - 1);
    }
    float cmin[4], cmax[4], cavg[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float mn = n[8].c[c], mx = n[8].c[c];
#pragma unroll
        for (int j = 7; j >= 0; --j) { mn = fminf(n[j].c[c], mn); mx = fmaxf(n[j].c[c], mx); }
        float sum = n[0].c[c];
#pragma unroll
        for (int j = 1; j < 9; ++j) sum += n[j].c[c];
        const float avg = MT_DIV_CONST(sum, 9.0f);   // exact 3-instruction division (tests/test_exact_tricks.py)
        const float mn5 = fminf(n[1].c[c], fminf(n[3].c[c], fminf(n[4].c[c], fminf(n[5].c[c], n[7].c[c]))));
        const float mx5 = fmaxf(n[1].c[c], fmaxf(n[3].c[c], fmaxf(n[4].c[c], fmaxf(n[5].c[c], n[7].c[c]))));
        const float avg5 = MT_DIV_CONST(((((n[1].c[c] + n[3].c[c]) + n[4].c[c]) + n[5].c[c]) + n[7].c[c]), 5.0f);
        cmin[c] = 0.5f * (mn + mn5);
        cmax[c] = 0.5f * (mx + mx5);
        cavg[c] = 0.5f * (avg + avg5);
    }
    // texture(prevFrameImage, old_uv): bilinear, border (0,0,0,1)
    float prevc[4];
    {
        const float tu = old_u * (float)P.W - 0.5f, tv = old_v * (float)P.H - 0.5f;
        const float fu = floorf(tu), fv = floorf(tv);
        const float ax = tu - fu, ay = tv - fv;
        const int x0 = mt_f2i(fu), y0 = mt_f2i(fv);
        C4 a, b, c2, d;
        if (MT_TXAA_FAST && x0 >= 0 && y0 >= 0 && x0 < P.W - 1 && y0 < P.H - 1) {  // all four texels inside the image
            const uint32_t* q00 = P.prev + ((unsigned)y0 * (unsigned)P.W + (unsigned)x0);
            a = ldr_unpack(MT_LDG(q00)); b = ldr_unpack(MT_LDG(q00 + 1));
            c2 = ldr_unpack(MT_LDG(q00 + P.W)); d = ldr_unpack(MT_LDG(q00 + P.W + 1));
        } else {
            a = ldr_border_texel(P.prev, P.W, P.H, x0, y0); b = ldr_border_texel(P.prev, P.W, P.H, x0 + 1, y0);
            c2 = ldr_border_texel(P.prev, P.W, P.H, x0, y0 + 1); d = ldr_border_texel(P.prev, P.W, P.H, x0 + 1, y0 + 1);
        }
        const float w00 = (1.0f - ax) * (1.0f - ay), w01 = ax * (1.0f - ay), w10 = (1.0f - ax) * ay, w11 = ax * ay;
#pragma unroll
        for (int c = 0; c < 4; ++c) prevc[c] = fmaf(w11, d.c[c], fmaf(w10, c2.c[c], fmaf(w01, b.c[c], w00 * a.c[c])));
    }
    // clip_aabb towards the neighbourhood box (:150-169)
    const float pw = clamp1(cavg[3], cmin[3], cmax[3]);
    float pclip[4], vclip[4], aunit[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        pclip[c] = 0.5f * (cmax[c] + cmin[c]);
        const float e = 0.5f * (cmax[c] - cmin[c]) + 0.0000000001f;
        vclip[c] = prevc[c] - pclip[c];
        aunit[c] = fabsf(div_nice(vclip[c], e));  // e >= 1e-10, |vclip| <= 1
    }
    const float ma = fmaxf(aunit[0], fmaxf(aunit[1], aunit[2]));
    pclip[3] = pw;
    vclip[3] = prevc[3] - pw;
    if (ma > 1.0f) {
#pragma unroll
        for (int c = 0; c < 4; ++c) prevc[c] = pclip[c] + div_nice(vclip[c], ma);  // ma > 1
    }
    const float* curr = n[4].c;
    const float lum0 = (curr[0] * 0.2125f + curr[1] * 0.7154f) + curr[2] * 0.0721f;
    const float lum1 = (prevc[0] * 0.2125f + prevc[1] * 0.7154f) + prevc[2] * 0.0721f;
    const float diff = div_nice(fabsf(lum0 - lum1), fmaxf(lum0, fmaxf(lum1, 0.2f)));
    const float wgt = 1.0f - diff;
    const float kfb = mix1(0.0f, 0.5f, wgt * wgt);
    uint32_t o = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) o |= unorm8(mix1(prevc[c], curr[c], kfb)) << (8 * c);
    return o;
}
