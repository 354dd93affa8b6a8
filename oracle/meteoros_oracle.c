/*
 * meteoros_oracle.c -- CPU restatement of the Meteoros cloud hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA kernels in meteoros_b200/csrc.  It is imported by tests/,
 * by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, and by nothing else:
 * the product library (libmeteoros_b200.so) never links, loads or calls it.
 *
 * PINNING: the reference ships no golden vectors or tests for this path and its
 * Vulkan application cannot run in the build container or on the GPU box (no loader/ICD, no glslang, no
 * SPIR-V).  The oracle follows the GLSL text op for op, and is pinned to that text directly: where
 * /root/reference exists, `make -C oracle refshaders` compiles the five shaders from their own source
 * (glsl2cpp.py, glm) and tests/test_reference_shaders.py requires this file to reproduce every byte they
 * store; the .npz fixtures under tests/golden carry reference-shader output to machines without the reference.  Unpinned:
 * what a Vulkan driver would add (built-in precision, FMA contraction, sampler fixed point) -- DESIGN.md 2.
 *
 * What it restates (paths relative to /root/reference/src/CloudScapes/shaders):
 *   cloudRayMarch.comp:106-132,153-273,295-306,331-388,401-467,489-563,565-688,690-779,824-825  -> mto_cloud
 *   reprojection.comp:72-244                                                                     -> mto_reproject
 *   postProcess_GodRays.frag:36-150 (+ postProcess_GenericVertShader.vert:15-16)                 -> mto_godrays
 *   postProcess_ToneMap.frag:32-84                                                               -> mto_tonemap
 * Grid shapes / pass order: Renderer.cpp:683-716, 122-192.  Sampler state: Texture3D.cpp:92-134,
 * Image.cpp:305-347 (LINEAR, REPEAT), Texture2D.cpp:75 (mask sampler CLAMP_TO_BORDER, opaque black).
 *
 * Canonical arithmetic (DESIGN.md "Canonical semantics"): everything is IEEE binary32, evaluated exactly
 * in the order written here, compiled with -ffp-contract=off -fno-fast-math; the only fused operations are
 * the explicit fmaf() calls of the texture filter.  GLSL built-ins are pinned as:
 *   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z           length(v) = sqrt(dot(v,v))
 *   normalize(v)  = v * (1.0f / sqrt(dot(v,v)))             mix(x,y,a) = x*(1-a) + y*a
 *   clamp(x,l,h)  = min(max(x,l),h)                         fract(x)  = x - floor(x)
 *   round(x)      = round-half-to-even                      int(x)/uint(x) = truncate, saturating, NaN -> 0
 *   mat4 * vec4   = ((m0*x + m1*y) + m2*z) + m3*w           a local never written reads 0
 *   texture()     = Vulkan linear filter, fp32 weights, see tex3d_linear / tex2d_linear below
 *   exp/pow/acos/cos = libm single precision (these never feed a discrete decision)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/meteoros_b200.h"
#include "meteoros_oracle.h"

/* Host threads of the OpenMP loops below.  torchrun exports OMP_NUM_THREADS=1 to every rank; the timing legs of bench.py
 * set the count explicitly and report what the runtime really uses (one libgomp per process: the setting also governs
 * oracle/_ref's loops). */
int mto_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}
int mto_num_threads(void)
{
#ifdef _OPENMP
    int n = 1;
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
    return n;
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------ */
/* small vector helpers, fp32, fixed evaluation order                                                */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;

static inline v3 V3(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 add3(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 div3(v3 a, v3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 scale3(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 divs3(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline v3 neg3(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float length3(v3 a) { return sqrtf(dot3(a, a)); }
static inline v3 normalize3(v3 a) { float r = 1.0f / sqrtf(dot3(a, a)); return scale3(a, r); }
static inline v3 cross3(v3 a, v3 b)
{
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
static inline float fractf(float x) { return x - floorf(x); }
static inline float smoothstepf(float e0, float e1, float x)
{
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
/* GLSL int(float): truncate; out-of-range saturates, NaN -> 0 (what the GPU's F2I does). */
static inline int32_t f2i(float x)
{
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}
static inline uint32_t f2u(float x)
{
    if (x != x || x <= 0.0f) return 0u;
    if (x >= 4294967296.0f) return UINT32_MAX;
    return (uint32_t)x;
}
/* pow as GLSL defines it for the domain the shaders use: x < 0 -> NaN, x == 0 -> 0 (y > 0). */
static inline float glsl_pow(float x, float y) { return powf(x, y); }

/* column-major mat4 (glm): m[c*4 + r] */
static inline v4 mat4_mul_v4(const float* m, v4 v)
{
    v4 r;
    r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
    r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
    r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
    r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
    return r;
}
static inline void mat4_mul_mat4(const float* a, const float* b, float* out)
{
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r)
            out[c * 4 + r] = ((a[0 * 4 + r] * b[c * 4 + 0] + a[1 * 4 + r] * b[c * 4 + 1]) + a[2 * 4 + r] * b[c * 4 + 2]) +
                             a[3 * 4 + r] * b[c * 4 + 3];
}

/* ------------------------------------------------------------------------------------------------ */
/* texture sampling: Vulkan linear filter on RGBA8_UNORM, REPEAT                                      */
/*   u = s*size - 0.5 ; i0 = floor(u) ; a = u - i0 ; i1 = i0 + 1 ; both wrapped modulo size           */
/*   weights w[k][j][i] = (wx_i * wy_j) * wz_k ;  value_c = (sum_{k,j,i} w * float(texel_c)) * (1/255) */
/*   the sum runs k-major, i fastest, first term a plain product, the rest fmaf                       */
/* ------------------------------------------------------------------------------------------------ */
static inline int wrapi(int i, int n)
{
    int m = i % n;
    return m < 0 ? m + n : m;
}

static v4 tex3d_linear(const uint8_t* vol, int W, int H, int D, float s, float t, float r)
{
    float u = s * (float)W - 0.5f, v = t * (float)H - 0.5f, w = r * (float)D - 0.5f;
    float fu = floorf(u), fv = floorf(v), fw = floorf(w);
    float ax = u - fu, ay = v - fv, az = w - fw;
    int x0 = wrapi(f2i(fu), W), y0 = wrapi(f2i(fv), H), z0 = wrapi(f2i(fw), D);
    int x1 = x0 + 1 == W ? 0 : x0 + 1, y1 = y0 + 1 == H ? 0 : y0 + 1, z1 = z0 + 1 == D ? 0 : z0 + 1;
    float wx[2] = { 1.0f - ax, ax }, wy[2] = { 1.0f - ay, ay }, wz[2] = { 1.0f - az, az };
    int xs[2] = { x0, x1 }, ys[2] = { y0, y1 }, zs[2] = { z0, z1 };
    float acc[4] = { 0, 0, 0, 0 };
    int first = 1;
    for (int k = 0; k < 2; ++k)
        for (int j = 0; j < 2; ++j)
            for (int i = 0; i < 2; ++i) {
                float wgt = (wx[i] * wy[j]) * wz[k];
                const uint8_t* p = vol + 4 * ((size_t)(zs[k] * H + ys[j]) * (size_t)W + (size_t)xs[i]);
                for (int c = 0; c < 4; ++c) {
                    if (first) acc[c] = wgt * (float)p[c];
                    else acc[c] = fmaf(wgt, (float)p[c], acc[c]);
                }
                first = 0;
            }
    const float inv255 = 1.0f / 255.0f;
    v4 o = { acc[0] * inv255, acc[1] * inv255, acc[2] * inv255, acc[3] * inv255 };
    return o;
}

static v4 tex2d_linear(const uint8_t* img, int W, int H, float s, float t)
{
    float u = s * (float)W - 0.5f, v = t * (float)H - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float ax = u - fu, ay = v - fv;
    int x0 = wrapi(f2i(fu), W), y0 = wrapi(f2i(fv), H);
    int x1 = x0 + 1 == W ? 0 : x0 + 1, y1 = y0 + 1 == H ? 0 : y0 + 1;
    float wx[2] = { 1.0f - ax, ax }, wy[2] = { 1.0f - ay, ay };
    int xs[2] = { x0, x1 }, ys[2] = { y0, y1 };
    float acc[4] = { 0, 0, 0, 0 };
    int first = 1;
    for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 2; ++i) {
            float wgt = wx[i] * wy[j];
            const uint8_t* p = img + 4 * ((size_t)ys[j] * (size_t)W + (size_t)xs[i]);
            for (int c = 0; c < 4; ++c) {
                if (first) acc[c] = wgt * (float)p[c];
                else acc[c] = fmaf(wgt, (float)p[c], acc[c]);
            }
            first = 0;
        }
    const float inv255 = 1.0f / 255.0f;
    v4 o = { acc[0] * inv255, acc[1] * inv255, acc[2] * inv255, acc[3] * inv255 };
    return o;
}

void mto_sample3d(const uint8_t* vol, int W, int H, int D, float s, float t, float r, float out[4])
{
    v4 o = tex3d_linear(vol, W, H, D, s, t, r);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
}
void mto_sample2d(const uint8_t* img, int W, int H, float s, float t, float out[4])
{
    v4 o = tex2d_linear(img, W, H, s, t);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
}

/* ------------------------------------------------------------------------------------------------ */
/* shared by CLOUD and REPROJ                                                                        */
/* ------------------------------------------------------------------------------------------------ */
#define EARTH_RADIUS 6371000.0f
#define ATMOSPHERE_RADIUS_INNER (EARTH_RADIUS + 7500.0f)   /* exact in fp32 */
#define ATMOSPHERE_RADIUS_OUTER (EARTH_RADIUS + 20000.0f)  /* exact in fp32 */
#define ATMOSPHERE_THICKNESS (ATMOSPHERE_RADIUS_OUTER - ATMOSPHERE_RADIUS_INNER)

typedef struct { v3 origin, direction; } Ray;
typedef struct { v3 normal, point; int valid; float t; } Intersection;

static inline float halton_at(const MtTimeUBO* tm, int seq, int i)
{
    const float* s = seq == 0 ? tm->haltonSeq1 : seq == 1 ? tm->haltonSeq2 : seq == 2 ? tm->haltonSeq3 : tm->haltonSeq4;
    return s[i];
}

/* cloudRayMarch.comp:114-132 (variant 0) / reprojection.comp:72-90 (variant 1: index>=4 reuses seq1/seq2). */
static inline void jitter_offset(const MtTimeUBO* tm, int index, float dimx, float dimy, int reproj_variant, float* jx, float* jy)
{
    index = index / 2;
    float x, y;
    if (index < 4) {
        x = halton_at(tm, 0, index);
        y = halton_at(tm, 1, index);
    } else {
        index -= 4;
        x = halton_at(tm, reproj_variant ? 0 : 2, index);
        y = halton_at(tm, reproj_variant ? 1 : 3, index);
    }
    *jx = x / dimx;
    *jy = y / dimy;
}

/* cloudRayMarch.comp:194-226, reprojection.comp:111-144 */
static Ray cast_ray(const MtCameraUBO* cam, const MtTimeUBO* tm, float sx, float sy, v3 eye, int pixelID, int W, int H, int reproj_variant)
{
    const float* view = cam->view; /* view[c][r] = view[c*4 + r] */
    v3 camRight = normalize3(V3(view[0 * 4 + 0], view[1 * 4 + 0], view[2 * 4 + 0]));
    v3 camUp = normalize3(V3(view[0 * 4 + 1], view[1 * 4 + 1], view[2 * 4 + 1]));
    v3 camLook = neg3(normalize3(V3(view[0 * 4 + 2], view[1 * 4 + 2], view[2 * 4 + 2])));

    float ndcx = sx * 2.0f - 1.0f;
    float ndcy = sy * 2.0f - 1.0f;
    float jx, jy;
    jitter_offset(tm, pixelID, (float)W, (float)H, reproj_variant, &jx, &jy);
    ndcx += jx;
    ndcy += jy;

    v3 cam_x = scale3(camRight, ndcx * cam->tanFovBy2[0]);
    v3 cam_y = scale3(camUp, ndcy * cam->tanFovBy2[1]);
    v3 ref = add3(eye, camLook);
    v3 p = add3(add3(ref, cam_x), cam_y);

    Ray r;
    r.origin = eye;
    r.direction = normalize3(sub3(p, eye));
    return r;
}

/* cloudRayMarch.comp:229-273 incl. the in-place rO overwrite that makes isect.t = |p_world - rO_normalised|. */
static Intersection ray_sphere(v3 rO, v3 rD, v3 c, float radius)
{
    Intersection is;
    is.valid = 0;
    is.point = V3(0, 0, 0);
    is.normal = V3(0, 1, 0);
    is.t = 0.0f; /* uninitialised in GLSL; canonical 0 */

    rO = sub3(rO, c);
    rO = divs3(rO, radius);

    float A = dot3(rD, rD);
    float B = 2.0f * dot3(rD, rO);
    float C = dot3(rO, rO) - 1.0f;
    float disc = B * B - (4.0f * A) * C;
    if (disc < 0.0f) return is;

    float sq = sqrtf(disc);
    float t = (-B - sq) / (2.0f * A);
    if (t < 0.0f) t = (-B + sq) / (2.0f * A);
    if (t >= 0.0f) {
        v3 p = add3(rO, scale3(rD, t));
        is.valid = 1;
        is.normal = normalize3(p);
        p = scale3(p, radius);
        p = add3(p, c);
        is.point = p;
        is.t = length3(sub3(p, rO));
    }
    return is;
}

/* ------------------------------------------------------------------------------------------------ */
/* CLOUD                                                                                             */
/* ------------------------------------------------------------------------------------------------ */
static inline float remapf(float v, float omin, float omax, float nmin, float nmax)
{
    return nmin + (((v - omin) / (omax - omin)) * (nmax - nmin));
}
static inline float remap_clamped(float v, float omin, float omax, float nmin, float nmax)
{
    return clampf(remapf(v, omin, omax, nmin, nmax), nmin, nmax);
}
static inline float remap_clamped_ba(float v, float omin, float omax, float nmin, float nmax)
{
    v = clampf(v, omin, omax);
    return clampf(remapf(v, omin, omax, nmin, nmax), nmin, nmax);
}

/* cloudRayMarch.comp:106-112 */
static v4 encode_float_rgba(float v)
{
    v4 e = { 1.0f * v, 255.0f * v, 65025.0f * v, 16581375.0f * v };
    e.x = fractf(e.x); e.y = fractf(e.y); e.z = fractf(e.z); e.w = fractf(e.w);
    /* enc -= enc.yzww * vec4(1/255, 1/255, 1/255, 0) */
    const float k = 1.0f / 255.0f;
    v4 o = { e.x - e.y * k, e.y - e.z * k, e.z - e.w * k, e.w - e.w * 0.0f };
    return o;
}

/* cloudRayMarch.comp:295-306 */
static float henyey_greenstein(float cosa, float g)
{
    float num = 1.0f - g * g;
    float den = glsl_pow((1.0f + g * g) - (2.0f * g) * cosa, 1.5f);
    return (num / den) * 0.07957747154594767f;
}
static float hg_modified(float cosa, float g, float silver_intensity, float silver_spread)
{
    return fmaxf(henyey_greenstein(cosa, g), silver_intensity * henyey_greenstein(cosa, 0.99f - silver_spread));
}

/* cloudRayMarch.comp:331-388 (only the live "THIRD INSTANCE" code) */
static float get_light_energy(float h, float dl, float ds_loded, float phase, float cosa, float brightness)
{
    float primary = expf(-dl);
    float secondary = expf(-dl);
    float att = fmaxf(remapf(cosa, 0.7f, 1.0f, secondary, secondary * 0.25f), primary);
    float depth = 0.05f + glsl_pow(ds_loded, clampf(remapf(h * 0.125f, 0.3f, 0.85f, 0.5f, 2.0f), 0.5f, 2.0f));
    float vertical = glsl_pow(clampf(remapf(h * 1.5f, 0.07f, 0.34f, 0.1f, 1.0f), 0.1f, 1.0f), 0.8f);
    float in_scatter = depth * vertical;
    return (((att * primary) * in_scatter) * phase) * brightness;
}

/* cloudRayMarch.comp:401-467, with BACKGROUND_SKY_SUN_LOCATION supplied through MtTuning. */
static v3 atmosphere_color(v3 dir, v3 sunDir, float sunIntensity, v3 skySunLoc)
{
    const float PI_F = 3.14159265f;
    const float E_F = 2.718281828459f;
    const v3 MIE_CONST = { 1.839991851443397f, 2.779802391966052f, 4.079047954386109f };
    const v3 RAYLEIGH_TOTAL = { 5.804542996261093E-6f, 1.3562911419845635E-5f, 3.0265902468824876E-5f };

    sunDir = normalize3(sunDir);
    /* calcSunIntensity(), :421-425 */
    float zenithAngleCos = clampf(normalize3(skySunLoc).y, -1.0f, 1.0f);
    float sunI = 1000.0f * fmaxf(0.0f, 1.0f - glsl_pow(E_F, -((1.6110731557f - acosf(zenithAngleCos)) / 1.5f)));
    float sunE = sunIntensity * sunI;
    /* calcSkyBetaR(), :406-411 */
    float sunFade = 1.0f - clampf(1.0f - expf(skySunLoc.y / 450000.0f), 0.0f, 1.0f);
    v3 BetaR = scale3(RAYLEIGH_TOTAL, (2.0f - 1.0f) + sunFade);
    /* calcSkyBetaV(), :413-419 */
    float c = (0.2f * 10.0f) * 10E-18f;
    v3 BetaM = scale3(scale3(MIE_CONST, 0.434f * c), 0.005f);

    float zenith = acosf(fmaxf(0.0f, dir.y));
    float inverse = 1.0f / (cosf(zenith) + 0.15f * glsl_pow(93.885f - ((zenith * 180.0f) / PI_F), -1.253f));
    float sR = 8.4E3f * inverse;
    float sM = 1.25E3f * inverse;

    v3 ex = add3(scale3(neg3(BetaR), sR), scale3(BetaM, sM));
    v3 fex = V3(expf(ex.x), expf(ex.y), expf(ex.z));

    float cosTheta = dot3(sunDir, dir);
    float rc = cosTheta * 0.5f + 0.5f;
    float rPhase = 0.05968310365946075f * (1.0f + rc * rc);
    v3 betaRTheta = scale3(BetaR, rPhase);
    float mPhase = henyey_greenstein(cosTheta, 0.8f);
    v3 betaMTheta = scale3(BetaM, mPhase);

    float yDot = 1.0f - sunDir.y;
    yDot *= ((yDot * yDot) * yDot) * yDot;
    v3 betas = div3(add3(betaRTheta, betaMTheta), add3(BetaR, BetaM));
    v3 one_m_fex = V3(1.0f - fex.x, 1.0f - fex.y, 1.0f - fex.z);
    v3 a = mul3(scale3(betas, sunE), one_m_fex);
    v3 Lin = V3(glsl_pow(a.x, 1.5f), glsl_pow(a.y, 1.5f), glsl_pow(a.z, 1.5f));
    v3 b = mul3(scale3(betas, sunE), fex);
    v3 bp = V3(glsl_pow(b.x, 0.5f), glsl_pow(b.y, 0.5f), glsl_pow(b.z, 0.5f));
    float ym = clampf(yDot, 0.0f, 1.0f);
    Lin = mul3(Lin, V3(mixf(1.0f, bp.x, ym), mixf(1.0f, bp.y, ym), mixf(1.0f, bp.z, ym)));

    v3 L0 = scale3(fex, 0.1f);
    const float SUN_ANGULAR_COS = 0.999956676946448443553574619906976478926848692873900859324f;
    float sunDisk = smoothstepf(SUN_ANGULAR_COS, SUN_ANGULAR_COS + 0.00002f, cosTheta);
    L0 = add3(L0, scale3(scale3(fex, sunE * 15000.0f), sunDisk));

    v3 color = add3(scale3(add3(Lin, L0), 0.04f), V3(0.0f, 0.0003f, 0.00075f));
    return color;
}

typedef struct {
    const MtCameraUBO* cam;
    const MtTimeUBO* tm;
    const MtTuning* tun;
    const MtoTextures* tex;
    int W, H;
} CloudEnv;

/* getDensityHeightGradientForPoint, cloudRayMarch.comp:475-487 (reached only through the weather path below) */
static float density_height_gradient(float relativeHeight, float cloudType)
{
    relativeHeight = clampf(relativeHeight, 0.0f, 1.0f);
    /* `cumulus` (:479) is computed by the shader but never used */
    float stratocumulus = fmaxf(0.0f, remapf(relativeHeight, 0.0f, 0.25f, 0.0f, 1.0f) * remapf(relativeHeight, 0.3f, 0.65f, 1.0f, 0.0f));
    float stratus = fmaxf(0.0f, remapf(relativeHeight, 0.0f, 0.1f, 0.0f, 1.0f) * remapf(relativeHeight, 0.2f, 0.3f, 1.0f, 0.0f));
    float a = mixf(stratus, stratocumulus, clampf(cloudType * 2.0f, 0.0f, 1.0f));
    float b = mixf(stratocumulus, stratus, clampf((cloudType - 0.5f) * 2.0f, 0.0f, 1.0f));
    return mixf(a, b, cloudType);
}

/* cloudRayMarch.comp:499-540.  With tun->use_weather the commented block :515-525 is live:
 * weather sampled at unskewedSamplePoint.xz (* weather_scale), base cloud scaled by the height gradient of the
 * weather's cloud type, coverage taken from the weather's red channel. */
static float sample_low_frequency(const CloudEnv* e, v3 p, v3 unskewed, float relativeHeight)
{
    v4 n = tex3d_linear(e->tex->low, e->tex->low_w, e->tex->low_h, e->tex->low_d, p.x, p.y, p.z);
    float fbm = (n.y * 0.625f + n.z * 0.25f) + n.w * 0.125f;
    fbm = clampf(fbm, 0.0f, 1.0f);
    float baseCloud = remap_clamped(n.x, fbm - 0.9f, 1.0f, 0.0f, 1.0f);
    float cov = e->tun->coverage;
    if (e->tun->use_weather && e->tex->weather) {
        float ws = e->tun->weather_scale;
        v4 wd = tex2d_linear(e->tex->weather, e->tex->weather_w, e->tex->weather_h, unskewed.x * ws, unskewed.z * ws);
        float grad = density_height_gradient(relativeHeight, wd.y);
        baseCloud *= grad * 0.5f;
        cov = wd.x;
    }
    float b = remap_clamped_ba(baseCloud, cov, 1.0f, 0.0f, 1.0f);
    b *= cov;
    return b;
}

/* cloudRayMarch.comp:542-563 */
static float erode_high_frequency(const CloudEnv* e, float baseCloud, v3 p, float h)
{
    v4 curl = tex2d_linear(e->tex->curl, e->tex->curl_w, e->tex->curl_h, p.x, p.y);
    p.x += (curl.x * (1.0f - h)) * 0.5f;
    p.y += (curl.y * (1.0f - h)) * 0.5f;
    v4 hf = tex3d_linear(e->tex->high, e->tex->high_w, e->tex->high_h, e->tex->high_d, p.x, p.y, p.z);
    float fbm = (hf.x * 0.625f + hf.y * 0.25f) + hf.z * 0.125f;
    float mod = clampf(mixf(fbm, 1.0f - fbm, clampf(h * 2.0f, 0.0f, 1.0f)), 0.0f, 1.0f);
    return remapf(baseCloud, mod * 0.005f, 1.0f, 0.0f, 1.0f);
}

/* cloudRayMarch.comp:565-688.  Colour is grey: one scalar. */
static float ray_march(const CloudEnv* e, Ray ray, v3 earthCenter, v3 startPos, float start_t, float end_t, int pixelID,
                       float* accumDensity, MtCounters* cnt, MtRayDebug* dbg)
{
    const MtTuning* tun = e->tun;
    float _dot = dot3(ray.direction, V3(0.0f, 1.0f, 0.0f));
    const float jitterfactor = 1.180f;
    const float baseDensityFactor = tun->base_density_factor;
    const float maxSteps = floorf(mixf(35.0f, 60.0f, 1.0f - _dot));
    const float atmosphereThickness = end_t - start_t;
    const float stepSize = atmosphereThickness / maxSteps;
    float transmittance = 1.0f;
    float returnColor = 0.0f;

    const v3 sunLoc = V3(tun->sun_location[0], tun->sun_location[1], tun->sun_location[2]);
    const v3 lightDir = normalize3(sub3(sunLoc, ray.origin));
    const float cos_angle = dot3(normalize3(ray.direction), lightDir);
    const float HG_light = hg_modified(cos_angle, 0.6f, 0.7f, 0.1f);

    v3 maxComp;
    float a0 = fabsf(lightDir.x), a1 = fabsf(lightDir.y), a2 = fabsf(lightDir.z);
    if (a0 > a1 && a0 > a2) maxComp = V3(a0, 0.0f, 0.0f);
    else if (a1 > a0 && a1 > a2) maxComp = V3(0.0f, a1, 0.0f);
    else maxComp = V3(0.0f, 0.0f, a2);
    v3 zC = cross3(lightDir, maxComp);
    v3 xC = cross3(zC, lightDir);
    /* mat3(xC, lightDir, zC) * v = xC*v.x + lightDir*v.y + zC*v.z */
    static const float K[6][3] = { { 0.1f, 0.25f, -0.15f }, { 0.2f, 0.5f, 0.2f },  { -0.2f, 0.1f, -0.1f },
                                   { -0.05f, 0.75f, 0.05f }, { -0.1f, 1.0f, 0.0f }, { 0.0f, 3.0f, 0.0f } };
    v3 kernel[6];
    for (int i = 0; i < 6; ++i)
        kernel[i] = add3(add3(scale3(xC, K[i][0]), scale3(lightDir, K[i][1])), scale3(zC, K[i][2]));

    if (dbg) { dbg->step_size = stepSize; dbg->steps = 0; dbg->jitter_hash = 2166136261u; }
    const v3 wind = V3(tun->wind_direction[0], tun->wind_direction[1], tun->wind_direction[2]);
    const float lengthToInner = length3(sub3(startPos, ray.origin));

    int iters = 0; /* maxSteps <= 60, so at most 61 iterations; the cap only guards degenerate shells and is never reached
                    * (MT_MAX_MARCH_ITERS: the same cap in every CUDA path) */
    for (float t = start_t; t < end_t && iters < 64; t += stepSize, ++iters) {
        /* int(mod(float(pixelID + int(t)), 16.0)) */
        float fi = (float)(pixelID + f2i(t));
        int _index = f2i(fi - 16.0f * floorf(fi / 16.0f));
        float jx, jy;
        jitter_offset(e->tm, _index, 75.0f, 75.0f, 0, &jx, &jy);
        v3 jdir = add3(ray.direction, V3(jx, (jx + jy) * jitterfactor, jy));
        v3 pos = add3(ray.origin, scale3(jdir, t));
        /* getRelativePositionInAtmosphere, :188-191, then /8 */
        v3 samplePoint = divs3(sub3(pos, V3(earthCenter.x, ATMOSPHERE_RADIUS_INNER - EARTH_RADIUS, earthCenter.z)), ATMOSPHERE_THICKNESS);
        samplePoint = divs3(samplePoint, 8.0f);
        /* getRelativeHeightInAtmosphere, :171-186 */
        float lenFromCam = length3(sub3(pos, ray.origin));
        v3 pointToEarthDir = normalize3(sub3(pos, earthCenter));
        float cosTheta = dot3(ray.direction, pointToEarthDir);
        float relativeHeight = fabsf(cosTheta * (lenFromCam - lengthToInner)) / ATMOSPHERE_THICKNESS;
        /* skewSamplePointWithWind, :489-497 */
        v3 skewed = add3(samplePoint, scale3(scale3(scale3(wind, relativeHeight), tun->cloud_top_offset), 0.009f));
        skewed = add3(skewed, scale3(scale3(add3(wind, V3(0.0f, 0.1f, 0.0f)), tun->cloud_speed), e->tm->time[1]));

        float baseDensity = sample_low_frequency(e, skewed, pos, relativeHeight) * baseDensityFactor;
        if (cnt) cnt->steps++;
        if (dbg) { dbg->steps++; dbg->jitter_hash = (dbg->jitter_hash ^ (uint32_t)_index) * 16777619u; }

        if (baseDensity > 0.0f) {
            if (cnt) cnt->steps_incloud++;
            float highFreqDensity = erode_high_frequency(e, baseDensity * 1.4f, skewed, relativeHeight);
            *accumDensity += highFreqDensity * 0.5f;

            float densityAlongLight = 0.0f;
            for (int i = 0; i < 6; ++i) {
                v3 lightPos = add3(pos, scale3(scale3(kernel[i], stepSize), (float)i));
                v3 sl = divs3(sub3(lightPos, V3(earthCenter.x, ATMOSPHERE_RADIUS_INNER - EARTH_RADIUS, earthCenter.z)), ATMOSPHERE_THICKNESS);
                float cur = sample_low_frequency(e, sl, sl, relativeHeight);
                if (cur > 0.0f) {
                    if (cnt) cnt->cone_hits++;
                    densityAlongLight += erode_high_frequency(e, 1.5f * cur, skewed, relativeHeight);
                }
            }
            float E = get_light_energy(relativeHeight, densityAlongLight, baseDensity, HG_light, cos_angle, 5.0f);
            transmittance = mixf(transmittance, E, 1.0f - *accumDensity);
            returnColor += transmittance;
        }
        if (*accumDensity >= 1.0f) {
            *accumDensity = 1.0f;
            if (cnt) cnt->early_exits++;
            break;
        }
    }
    return returnColor;
}

/* One invocation of cloudRayMarch.comp main(), :690-826, for pixel (px,py) with the given pixelID. */
static void cloud_pixel(const CloudEnv* e, int px, int py, int pixelID, float* hdr, float* mask, MtCounters* cnt, MtRayDebug* dbgbuf)
{
    const int W = e->W, H = e->H;
    float u = (float)px / (float)W;
    float v = (float)py / (float)H;
    v = 1.0f - v;
    v3 eyePos = V3(-e->cam->eye[0], -e->cam->eye[1], -e->cam->eye[2]);
    Ray ray = cast_ray(e->cam, e->tm, u, v, eyePos, pixelID, W, H, 0);

    MtRayDebug* dbg = dbgbuf ? &dbgbuf[(size_t)py * W + px] : NULL;
    if (dbg) {
        memset(dbg, 0, sizeof(*dbg));
        dbg->dir[0] = ray.direction.x; dbg->dir[1] = ray.direction.y; dbg->dir[2] = ray.direction.z;
    }
    if (cnt) cnt->rays++;

    const float sunIntensity = 0.780f;
    float _dot = dot3(V3(0.0f, 1.0f, 0.0f), ray.direction);
    const float bgMul = fmaxf(0.620f, _dot);
    const float cloudFadeOutPoint = 0.06f;
    const v3 skySun = V3(e->tun->sky_sun_location[0], e->tun->sky_sun_location[1], e->tun->sky_sun_location[2]);
    float* o = hdr + 4 * ((size_t)py * W + px);
    float* m = mask + 4 * ((size_t)py * W + px);

    if (_dot < 0.0f) {
        v3 colorNearHorizon = scale3(V3(0.0f, 0.16f, 0.51f), 0.4f);
        v3 color2 = scale3(V3(0.0f, 0.73f, 0.95f), 0.5f);
        float a = -ray.direction.y * 5.5f;
        o[0] = mixf(colorNearHorizon.x, color2.x, a);
        o[1] = mixf(colorNearHorizon.y, color2.y, a);
        o[2] = mixf(colorNearHorizon.z, color2.z, a);
        o[3] = 1.0f;
        m[0] = m[1] = m[2] = m[3] = 0.0f;
        if (dbg) dbg->branch = 0;
        return;
    }
    v3 bg = atmosphere_color(ray.direction, sub3(skySun, ray.origin), sunIntensity, skySun);
    bg = scale3(bg, bgMul);
    if (_dot < cloudFadeOutPoint) {
        o[0] = bg.x; o[1] = bg.y; o[2] = bg.z; o[3] = 1.0f;
        m[0] = m[1] = m[2] = m[3] = 0.0f;
        if (dbg) dbg->branch = 1;
        return;
    }

    v3 earthCenter = eyePos;
    earthCenter.y = -EARTH_RADIUS;
    Intersection in = ray_sphere(ray.origin, ray.direction, earthCenter, ATMOSPHERE_RADIUS_INNER);
    Intersection out = ray_sphere(ray.origin, ray.direction, earthCenter, ATMOSPHERE_RADIUS_OUTER);

    float accum = 0.0f;
    if (cnt) cnt->rays_marched++;
    if (dbg) { dbg->branch = 2; dbg->t_in = in.t; dbg->t_out = out.t; }
    float march = ray_march(e, ray, earthCenter, in.point, in.t, out.t, pixelID, &accum, cnt, dbg);
    if (dbg) dbg->accum = accum;
    float godAccum = accum;

    accum *= smoothstepf(0.0f, 1.0f, fminf(1.0f, remapf(ray.direction.y, cloudFadeOutPoint, 0.2f, 0.0f, 1.0f)));
    o[0] = mixf(bg.x, march, accum);
    o[1] = mixf(bg.y, march, accum);
    o[2] = mixf(bg.z, march, accum);
    o[3] = 1.0f;

    float grey = 25.0f * fminf(0.05f, 1.0f - godAccum);
    v4 enc = encode_float_rgba(grey);
    if (ray.direction.y < 0.05f) { /* dead in this branch (dir.y >= 0.06); kept for fidelity, :775-779 */
        float k = fmaxf(5.0f, ray.direction.y);
        enc.x *= k; enc.y *= k; enc.z *= k; enc.w *= k;
    }
    m[0] = enc.x; m[1] = enc.y; m[2] = enc.z; m[3] = enc.w;
}

static void counters_add(MtCounters* a, const MtCounters* b)
{
    a->rays += b->rays; a->rays_marched += b->rays_marched; a->steps += b->steps;
    a->steps_incloud += b->steps_incloud; a->cone_hits += b->cone_hits; a->early_exits += b->early_exits;
}

/* Reference grid for the cloud dispatch, Renderer.cpp:713-714 (integer division, then round up to 32). */
static void cloud_grid(int W, int H, int* tx, int* ty)
{
    *tx = (((W / 4) + 31) / 32) * 32;
    *ty = (((H / 4) + 31) / 32) * 32;
}

int mto_cloud(const MtCameraUBO* cam, const MtTimeUBO* tm, const MtTuning* tun, const MtoTextures* tex, int W, int H,
              int full, int row_begin, int row_end, int group_stride, float* hdr, float* mask, MtCounters* counters,
              MtRayDebug* debug)
{
    if (group_stride < 1) group_stride = 1;
    if (!cam || !tm || !tun || !tex || !hdr || !mask || W <= 0 || H <= 0) return 1;
    if (!tex->low || !tex->high || !tex->curl) return 1;
    CloudEnv e = { cam, tm, tun, tex, W, H };
    int tx, ty;
    cloud_grid(W, H, &tx, &ty);
    if (row_begin < 0) row_begin = 0;
    if (row_end > H || row_end <= 0) row_end = H;
    MtCounters total;
    memset(&total, 0, sizeof(total));
    const int id0 = tm->frameCountMod16;
    const int nid = full ? 16 : 1;
#pragma omp parallel
    {
        MtCounters local;
        memset(&local, 0, sizeof(local));
#pragma omp for schedule(dynamic, 1)
        for (int gy = 0; gy < ty; gy += group_stride) {
            for (int k = 0; k < nid; ++k) {
                int pixelID = full ? k : id0;
                int pX = pixelID / 4, pY = pixelID % 4; /* :697-698 */
                int py = gy * 4 + pY;
                if (py >= H || py < row_begin || py >= row_end) continue; /* imageStore outside the image is dropped */
                for (int gx = 0; gx < tx; ++gx) {
                    int px = gx * 4 + pX;
                    if (px >= W) continue;
                    cloud_pixel(&e, px, py, pixelID, hdr, mask, counters ? &local : NULL, debug);
                }
            }
        }
#pragma omp critical
        counters_add(&total, &local);
    }
    if (counters) counters_add(counters, &total);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* REPROJ  reprojection.comp:193-244                                                                 */
/* ------------------------------------------------------------------------------------------------ */
int mto_reproject(const MtCameraUBO* cam, const MtCameraUBO* camOld, const MtTimeUBO* tm, int W, int H, const float* prev,
                  float* cur, int32_t* taps)
{
    if (!cam || !camOld || !tm || !prev || !cur || W <= 0 || H <= 0) return 1;
    const int pixelID = tm->frameCountMod16;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            float u = (float)x / (float)W;
            float v = (float)y / (float)H;
            v3 eyePos = V3(-cam->eye[0], -cam->eye[1], -cam->eye[2]);
            Ray ray = cast_ray(cam, tm, u, v, eyePos, pixelID, W, H, 1);
            v3 earthCenter = eyePos;
            earthCenter.y = -EARTH_RADIUS;
            Intersection in = ray_sphere(ray.origin, ray.direction, earthCenter, ATMOSPHERE_RADIUS_INNER);

            v4 pc = { in.point.x, in.point.y, in.point.z, 1.0f };
            v4 q4 = mat4_mul_v4(camOld->view, pc);
            v3 q = normalize3(V3(q4.x, q4.y, q4.z));
            q = divs3(q, -q.z);
            float old_u = (q.x / cam->tanFovBy2[0]) * 0.5f + 0.5f;
            float old_v = (q.y / cam->tanFovBy2[1]) * 0.5f + 0.5f;
            float mx = old_u - u, my = old_v - v;
            float acc[4] = { 0, 0, 0, 0 };
            for (int i = 0; i < 10; ++i) {
                float f = (float)i / 10.0f;
                float bx = mx * f, by = my * f;
                float ix = rintf((old_u - bx) * (float)W);
                float iy = rintf((old_v - by) * (float)H);
                int cx = f2i(ix), cy = f2i(iy);
                cx = cx < 0 ? 0 : (cx > W - 1 ? W - 1 : cx);
                cy = cy < 0 ? 0 : (cy > H - 1 ? H - 1 : cy);
                const float* p = prev + 4 * ((size_t)cy * W + cx);
                acc[0] += p[0]; acc[1] += p[1]; acc[2] += p[2]; acc[3] += p[3];
                if (taps) taps[((size_t)y * W + x) * 10 + i] = cy * W + cx;
            }
            float* o = cur + 4 * ((size_t)y * W + x);
            o[0] = acc[0] / 10.0f; o[1] = acc[1] / 10.0f; o[2] = acc[2] / 10.0f; o[3] = acc[3] / 10.0f;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* GODRAY  postProcess_GodRays.frag:66-150.  Mask sampler: LINEAR, CLAMP_TO_BORDER, border (0,0,0,1). */
/* ------------------------------------------------------------------------------------------------ */
static inline void mask_texel(const float* mask, int W, int H, int x, int y, float out[4])
{
    if (x < 0 || y < 0 || x >= W || y >= H) { out[0] = out[1] = out[2] = 0.0f; out[3] = 1.0f; return; }
    const float* p = mask + 4 * ((size_t)y * W + x);
    out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = p[3];
}
static void mask_bilinear(const float* mask, int W, int H, float s, float t, float acc[4])
{
    float u = s * (float)W - 0.5f, v = t * (float)H - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float ax = u - fu, ay = v - fv;
    int x0 = f2i(fu), y0 = f2i(fv);
    float wx[2] = { 1.0f - ax, ax }, wy[2] = { 1.0f - ay, ay };
    int first = 1;
    for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 2; ++i) {
            float tx[4];
            mask_texel(mask, W, H, x0 + i, y0 + j, tx);
            float wgt = wx[i] * wy[j];
            for (int c = 0; c < 4; ++c) {
                if (first) acc[c] = wgt * tx[c];
                else acc[c] = fmaf(wgt, tx[c], acc[c]);
            }
            first = 0;
        }
}
static float mask_decode_bilinear(const float* mask, int W, int H, float s, float t)
{
    float acc[4];
    mask_bilinear(mask, W, H, s, t, acc);
    /* dot(x, 1/bitEnc) */
    const float d0 = 1.0f / 1.0f, d1 = 1.0f / 255.0f, d2 = 1.0f / 65025.0f, d3 = 1.0f / 16581375.0f;
    return ((acc[0] * d0 + acc[1] * d1) + acc[2] * d2) + acc[3] * d3;
}

int mto_godrays(const MtCameraUBO* cam, const MtSunAndSkyUBO* sky, int W, int H, const float* mask, float* hdr)
{
    if (!cam || !sky || !mask || !hdr || W <= 0 || H <= 0) return 1;
    const v3 sunLocation = V3(0.0f, 1.0f, 0.0f);
    v3 cam_to_sun = normalize3(sub3(sunLocation, V3(cam->eye[0], cam->eye[1], cam->eye[2])));
    v3 camForward = neg3(normalize3(V3(cam->view[0 * 4 + 2], cam->view[1 * 4 + 2], cam->view[2 * 4 + 2])));
    float blendFactor = dot3(cam_to_sun, camForward);
    if (blendFactor < 0.0f) return 0; /* every fragment returns without writing */

    float pv[16];
    mat4_mul_mat4(cam->proj, cam->view, pv);
    v4 sun4 = { 0.0f, 1.0f, 0.0f, 1.0f };
    v4 ndc = mat4_mul_v4(pv, sun4);
    float sunx = clampf((ndc.x + 1.0f) / 2.0f, 0.0f, 1.0f);
    float suny = clampf((ndc.y + 1.0f) / 2.0f, 0.0f, 1.0f);
    /* flag_sunOutsideCamFrustum is always false after the clamp: only the "normal" loop is live (:128-137) */

#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            float u = ((float)x + 0.5f) / (float)W;
            float v = ((float)y + 0.5f) / (float)H;
            float dux = ((u - sunx) / 100.0f) * 1.0f;
            float duy = ((v - suny) / 100.0f) * 1.0f;
            float acc[3] = { 0, 0, 0 };
            float decay = 1.0f;
            for (int i = 0; i < 100; ++i) {
                float a = mask_decode_bilinear(mask, W, H, u, v);
                for (int c = 0; c < 3; ++c) {
                    float sc = sky->lightColor[c] * a;
                    sc *= decay * 0.001f;
                    acc[c] += sc;
                }
                decay *= 1.0f;
                u -= dux;
                v -= duy;
            }
            float* o = hdr + 4 * ((size_t)y * W + x);
            o[0] = o[0] + (acc[0] * 1.0f) * blendFactor;
            o[1] = o[1] + (acc[1] * 1.0f) * blendFactor;
            o[2] = o[2] + (acc[2] * 1.0f) * blendFactor;
            o[3] = o[3] + 1.0f * blendFactor;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* TONEMAP  postProcess_ToneMap.frag:32-84                                                           */
/* ------------------------------------------------------------------------------------------------ */
static inline float uncharted2(float x)
{
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F;
}
uint32_t mto_wang_hash(uint32_t u, uint32_t v, uint32_t s)
{
    uint32_t seed = (u * 1664525u + v) + s;
    seed = (seed ^ 61u) ^ (seed >> 16u);
    seed *= 9u;
    seed = seed ^ (seed >> 4u);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15u);
    return seed;
}
static inline uint8_t to_unorm8(float v)
{
    if (v != v) return 0;
    v = clampf(v, 0.0f, 1.0f);
    return (uint8_t)rintf(v * 255.0f);
}

int mto_tonemap(const MtTimeUBO* tm, int W, int H, const float* hdr, uint8_t* ldr, float* ldr_f32)
{
    if (!tm || !hdr || (!ldr && !ldr_f32) || W <= 0 || H <= 0) return 1;
    const uint32_t s = f2u(tm->time[1]);
    const float whitemap = 1.0f / uncharted2(100.0f);
    const float invGamma = 1.0f / 2.2f;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const float* p = hdr + 4 * ((size_t)y * W + x);
            float noise = ((float)mto_wang_hash((uint32_t)x, (uint32_t)y, s) * (1.0f / 4294967296.0f)) * 0.01f;
            float o[3];
            for (int c = 0; c < 3; ++c) {
                float col = uncharted2(2.5f * p[c]);
                col *= whitemap;
                col = glsl_pow(col, invGamma);
                o[c] = col + noise;
            }
            size_t k = (size_t)y * W + x;
            if (ldr) {
                ldr[4 * k + 0] = to_unorm8(o[0]); ldr[4 * k + 1] = to_unorm8(o[1]);
                ldr[4 * k + 2] = to_unorm8(o[2]); ldr[4 * k + 3] = 255;
            }
            if (ldr_f32) {
                ldr_f32[4 * k + 0] = o[0]; ldr_f32[4 * k + 1] = o[1]; ldr_f32[4 * k + 2] = o[2]; ldr_f32[4 * k + 3] = 1.0f;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* TXAA  postProcess_TXAA.frag:171-270.  LDR images are RGBA8 UNORM.           */
/* The shader reads the 3x3 neighbourhood from, and writes its result to, the SAME storage image from   */
/* concurrently running fragments (a race); canonical: the neighbourhood is read from the tone-mapped    */
/* image as it was before the pass, the result goes to `out`.  imageLoad outside the image returns 0.   */
/* ------------------------------------------------------------------------------------------------ */
static inline void ldr_load(const uint8_t* img, int W, int H, int x, int y, float out[4])
{
    if (x < 0 || y < 0 || x >= W || y >= H) { out[0] = out[1] = out[2] = out[3] = 0.0f; return; }
    const uint8_t* p = img + 4 * ((size_t)y * W + x);
    for (int c = 0; c < 4; ++c) out[c] = (float)p[c] * (1.0f / 255.0f);
}
/* texture(prevFrameImage, uv): LINEAR, CLAMP_TO_BORDER with opaque black (Texture2D.cpp:75) */
static void ldr_sample_border(const uint8_t* img, int W, int H, float s, float t, float out[4])
{
    float u = s * (float)W - 0.5f, v = t * (float)H - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float ax = u - fu, ay = v - fv;
    int x0 = f2i(fu), y0 = f2i(fv);
    float wx[2] = { 1.0f - ax, ax }, wy[2] = { 1.0f - ay, ay };
    float acc[4] = { 0, 0, 0, 0 };
    int first = 1;
    for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 2; ++i) {
            float tx[4];
            int x = x0 + i, y = y0 + j;
            if (x < 0 || y < 0 || x >= W || y >= H) { tx[0] = tx[1] = tx[2] = 0.0f; tx[3] = 1.0f; }
            else ldr_load(img, W, H, x, y, tx);
            float wgt = wx[i] * wy[j];
            for (int c = 0; c < 4; ++c) {
                if (first) acc[c] = wgt * tx[c];
                else acc[c] = fmaf(wgt, tx[c], acc[c]);
            }
            first = 0;
        }
    for (int c = 0; c < 4; ++c) out[c] = acc[c];
}

int mto_txaa(const MtCameraUBO* cam, const MtCameraUBO* camOld, const MtTimeUBO* tm, int W, int H, const uint8_t* cur,
             const uint8_t* prev, uint8_t* out, float* out_f32)
{
    if (!cam || !camOld || !tm || !cur || !prev || (!out && !out_f32) || W <= 0 || H <= 0) return 1;
    const int pixelID = tm->frameCountMod16;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            float u = ((float)x + 0.5f) / (float)W;
            float v = ((float)y + 0.5f) / (float)H;
            v3 eyePos = V3(-cam->eye[0], -cam->eye[1], -cam->eye[2]);
            Ray ray = cast_ray(cam, tm, u, v, eyePos, pixelID, W, H, 0);
            v3 earthCenter = eyePos;
            earthCenter.y = -EARTH_RADIUS;
            Intersection in = ray_sphere(ray.origin, ray.direction, earthCenter, ATMOSPHERE_RADIUS_INNER);
            v4 pc = { in.point.x, in.point.y, in.point.z, 1.0f };
            v4 q4 = mat4_mul_v4(camOld->view, pc);
            v3 q = normalize3(V3(q4.x, q4.y, q4.z));
            q = divs3(q, -q.z);
            float old_u = (q.x / cam->tanFovBy2[0]) * 0.5f + 0.5f;
            float old_v = (q.y / cam->tanFovBy2[1]) * 0.5f + 0.5f;

            /* neighbourHoodClamping, :171-199; order: tl tc tr ml mc mr bl bc br */
            float n[9][4];
            int k = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) ldr_load(cur, W, H, x + dx, y + dy, n[k++]);
            float cmin[4], cmax[4], cavg[4];
            for (int c = 0; c < 4; ++c) {
                float mn = n[8][c], mx = n[8][c];
                for (int j = 7; j >= 0; --j) { mn = fminf(n[j][c], mn); mx = fmaxf(n[j][c], mx); }
                float sum = n[0][c];
                for (int j = 1; j < 9; ++j) sum += n[j][c];
                float avg = sum / 9.0f;
                float mn5 = fminf(n[1][c], fminf(n[3][c], fminf(n[4][c], fminf(n[5][c], n[7][c]))));
                float mx5 = fmaxf(n[1][c], fmaxf(n[3][c], fmaxf(n[4][c], fmaxf(n[5][c], n[7][c]))));
                float avg5 = ((((n[1][c] + n[3][c]) + n[4][c]) + n[5][c]) + n[7][c]) / 5.0f;
                cmin[c] = 0.5f * (mn + mn5);
                cmax[c] = 0.5f * (mx + mx5);
                cavg[c] = 0.5f * (avg + avg5);
            }
            const float* curr = n[4];
            float prevc[4];
            ldr_sample_border(prev, W, H, old_u, old_v, prevc);

            /* clip_aabb(cmin.xyz, cmax.xyz, clamp(cavg, cmin, cmax), prevColor), :150-169 */
            float pw = clampf(cavg[3], cmin[3], cmax[3]);
            float pclip[4], vclip[4], aunit[3];
            for (int c = 0; c < 3; ++c) {
                pclip[c] = 0.5f * (cmax[c] + cmin[c]);
                float e = 0.5f * (cmax[c] - cmin[c]) + 0.0000000001f; /* EPSILON */
                vclip[c] = prevc[c] - pclip[c];
                aunit[c] = fabsf(vclip[c] / e);
            }
            float ma = fmaxf(aunit[0], fmaxf(aunit[1], aunit[2]));
            pclip[3] = pw;
            vclip[3] = prevc[3] - pw;
            if (ma > 1.0f)
                for (int c = 0; c < 4; ++c) prevc[c] = pclip[c] + vclip[c] / ma;

            float lum0 = (curr[0] * 0.2125f + curr[1] * 0.7154f) + curr[2] * 0.0721f;
            float lum1 = (prevc[0] * 0.2125f + prevc[1] * 0.7154f) + prevc[2] * 0.0721f;
            float diff = fabsf(lum0 - lum1) / fmaxf(lum0, fmaxf(lum1, 0.2f));
            float wgt = 1.0f - diff;
            float kfb = mixf(0.0f, 0.5f, wgt * wgt);
            size_t idx = (size_t)y * W + x;
            for (int c = 0; c < 4; ++c) {
                float o = mixf(prevc[c], curr[c], kfb);
                if (out) out[4 * idx + c] = to_unorm8(o);
                if (out_f32) out_f32[4 * idx + c] = o;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* small exported helpers used by unit tests                                                         */
/* ------------------------------------------------------------------------------------------------ */
void mto_encode_float_rgba(float v, float out[4])
{
    v4 e = encode_float_rgba(v);
    out[0] = e.x; out[1] = e.y; out[2] = e.z; out[3] = e.w;
}
void mto_ray_sphere(const float ro[3], const float rd[3], const float c[3], float radius, float point[3], float* t, int* valid)
{
    Intersection is = ray_sphere(V3(ro[0], ro[1], ro[2]), V3(rd[0], rd[1], rd[2]), V3(c[0], c[1], c[2]), radius);
    point[0] = is.point.x; point[1] = is.point.y; point[2] = is.point.z;
    *t = is.t;
    *valid = is.valid;
}
/* the fixed-function pieces of the post passes, for oracle/glsl_rt.h (the reference's shaders compiled as C++) */
void mto_sample2d_f32_border(const float* img, int W, int H, float s, float t, float out[4]) { mask_bilinear(img, W, H, s, t, out); }
void mto_sample2d_unorm8_border(const uint8_t* img, int W, int H, float s, float t, float out[4]) { ldr_sample_border(img, W, H, s, t, out); }
void mto_load_unorm8(const uint8_t* img, int W, int H, int x, int y, float out[4]) { ldr_load(img, W, H, x, y, out); }
uint8_t mto_to_unorm8(float v) { return to_unorm8(v); }
float mto_density_height_gradient(float relativeHeight, float cloudType) { return density_height_gradient(relativeHeight, cloudType); }
void mto_cloud_grid(int W, int H, int* threads_x, int* threads_y) { cloud_grid(W, H, threads_x, threads_y); }
void mto_atmosphere_color(const float dir[3], const float sun_minus_origin[3], float sunIntensity, const float skySun[3], float out[3])
{
    v3 c = atmosphere_color(V3(dir[0], dir[1], dir[2]), V3(sun_minus_origin[0], sun_minus_origin[1], sun_minus_origin[2]), sunIntensity,
                            V3(skySun[0], skySun[1], skySun[2]));
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
