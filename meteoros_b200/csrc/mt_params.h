// mt_params.h -- kernel parameter blocks (passed by value; they live in the constant bank of each launch).
#pragma once

#include "../../include/meteoros_b200.h"
#if defined(MT_HOSTSIM)
struct float2 { float x, y; };
#else
#include <vector_types.h>
#endif
#include "mt_math.cuh"
#include "mt_tex.cuh"

// Ray tile of one warp / one 128-thread CTA in the full-quality Cloud launch (cloud_raymarch.cu).  0: 8x4 per warp, CTA
// 16x8; 1: 16x2, CTA 16x8; 2: 4x8, CTA 16x8; 3: 32x1, CTA 32x4 (rays of a row share dir.y, hence step count and size).
#ifndef MT_WARP_SHAPE
#define MT_WARP_SHAPE 1  /* measured at 4K: 16x2 5.07 ms, 8x4 5.12 ms, 32x1 5.15 ms, 4x8 5.29 ms */
#endif
#if MT_WARP_SHAPE == 3
#define MT_CTA_W 32
#define MT_CTA_H 4
#else
#define MT_CTA_W 16
#define MT_CTA_H 8
#endif

// Ray tile of one CTA of the fused 1-of-16 kernel (cloud_sixteenth_kernel): 2^MT_S16_LOG2W x (32 >> MT_S16_LOG2W) rays, four pixels apart
#ifndef MT_S16_LOG2W
#define MT_S16_LOG2W 3  /* 8x4 rays = 32x16 pixels */
#endif
#define MT_S16_TW (1 << MT_S16_LOG2W)
#define MT_S16_TH (32 >> MT_S16_LOG2W)

/* One cap on the march loop for every path -- the sequential kernels, the step-parallel 1-of-16 kernels and the oracle.  The
 * shader's own bound is maxSteps <= 60 (cloudRayMarch.comp:585), i.e. at most 61 iterations of `t += stepSize`; 59 is the most
 * any camera of the test suite produces.  The cap only guards degenerate shells and is never reached. */
#define MT_MAX_MARCH_ITERS 64
#define MT_STEP_SLICES MT_MAX_MARCH_ITERS  /* step-parallel paths: sample slots per ray */

struct F4 {  // 16-byte pixel; float4 on the device
    float x, y, z, w;
};

// Per-frame constants of the Preetham sky (cloudRayMarch.comp:401-467): everything that does not depend on the
// ray.  Continuous radiance terms only, so they are evaluated once on the host (mt_host_sky_const).
struct SkyConst {
    float sunDir[3];   // normalize(BACKGROUND_SKY_SUN_LOCATION - origin)
    float sunE;        // SUN_INTENSITY * calcSunIntensity()
    float betaR[3];    // calcSkyBetaR()
    float betaM[3];    // calcSkyBetaV()
    float invBeta[3];  // unused by the canonical path (kept for alignment)
    float yDotMix;     // clamp((1 - sunDir.y)^5, 0, 1)
};

// Per-frame constants of the march that feed discrete decisions, in the canonical operation order (cloudRayMarch.comp:114-132,
// 199-207, 585-624, 489-497).  Round 1 / 2 evaluated them in a one-thread kernel and every CTA staged them in shared memory;
// the compiler then kept the ones the march loop reads in registers and spilled them (six LDL per march step,
// profiles/r2_cloud_final.md).  Now the HOST evaluates them per dispatch (cloud_frame_setup is __host__ __device__: IEEE
// + - * / sqrt without contraction on both sides, the same bits -- the GPU parity tests hold the decisions they feed to the
// oracle's) and they travel in the kernel's parameter block: every use is a constant-bank operand, no register, no load, no
// setup launch.  Only the two jitter tables, which a lane indexes by its own pixel id / step, are staged in shared memory
// (a divergent constant-bank index would serialise).
struct MarchTabs {
    float rayJitter[8][2];              // getJitterOffset(id, dim): (halton_x / W, halton_y / H) for id/2 = 0..7
    float stepJitter[8][4];             // per-step direction offset (jx, (jx+jy)*1.18, jy, 0), j = halton / 75
};
struct MarchConst {
    f3 basisRight, basisUp, basisLook;  // castRay basis
    f3 eyePos;                          // -camera.eye
    f3 earthCenter;                     // (eye.x, -R, eye.z)
    f3 lightDir;                        // normalize(SUN_LOCATION - origin)
    f3 windSkew;                        // ((WIND_DIRECTION + (0,.1,0)) * CLOUD_SPEED) * time.y
    f3 coneStep[6];                     // noise_kernel[i] (unscaled)
    float covDen;                       // 1 - coverage: divisor of the coverage remap (its refined reciprocal is per ray: RaySetup.covRcp)
    float covScale;                     // coverage / (1 - coverage): the coverage remap of a light-cone sample as one multiplication
    MarchTabs tabs;
};
#define MT_MARCHTABS_WORDS (sizeof(MarchTabs) / 4)

// In-cloud compaction of the step-parallel march (cloud_raymarch.cu): measured, slower at 4K, off (profiles/r1_ab.md).
#ifndef MT_STEP_COMPACT
#define MT_STEP_COMPACT 0
#endif

struct RowTiles {  // which pixel rows this launch covers (multi-GPU row-tile shards); default = whole image
    int tile_rows;    // rows per tile (multiple of the block height)
    int tile_begin;   // first tile owned
    int tile_stride;  // distance between owned tiles
    int tile_count;   // number of owned tiles
    int heavy_first;  // launch order (mt_tile_order): the first `heavy_first` owned tiles lie above the horizon and march; they are
                      // issued from the horizon upwards -- most march steps first, zenith rows last -- so that the kernel's tail is
                      // made of its cheapest marching CTAs (matters when a GPU renders 1/8 of a frame), and the remaining tiles
                      // (sky band, ocean: ~100 instructions per CTA, but the same 32 bytes per pixel to store) are INTERLEAVED with
                      // them one for one, so that their stores -- half of the frame's bytes -- spread over the whole kernel instead of
                      // arriving as one burst at its end (a burst the NVLink gather of an 8-GPU frame then waits for); 0 = top-down
};

// j-th tile in launch order -> index among the owned tiles
MT_DEVICE int mt_tile_order(const RowTiles& r, int j)
{
    const int m = r.heavy_first, l = r.tile_count - m;  // marching tiles, light tiles
    if (m <= 0) return j;
    const int pairs = m < l ? m : l;
    if (j < 2 * pairs) return (j & 1) ? m + (j >> 1) : m - 1 - (j >> 1);
    const int k = j - 2 * pairs;                       // what is left of the longer list
    return m > l ? m - 1 - (pairs + k) : m + pairs + k;
}

struct CloudParams {
    CamU cam;
    TimeU tm;
    MtTuning tun;
    SkyConst sky;
    Tex3D low, high;
    Tex2D curl;
    Tex2D weather;         // sampled only when tun.use_weather
    MarchConst mc;         // per-frame constants, evaluated by the host (mt_context.cu, cloud_frame_setup)
    F4* hdr;
    F4* mask;
    int W, H;
    int tx, ty;  // threads of the reference dispatch (Renderer.cpp:713-716)
    int full;    // 0: one id (tm.frameCountMod16) -- 1: all sixteen
    int storage;     // MtStorage of the HDR / mask images (mt_pixel.cuh)
    int bulkStore;   // full-quality kernel: HDR pixels leave through shared memory + cp.async.bulk (mtSetCloudStoreMode)
    int hwCone;      // MT_FLAG_HW_CONE_FILTER: full-quality production launches take their light-cone samples through low.hwtex
    RowTiles rows;
    unsigned long long* counters;  // 6 x u64 or null
    MtRayDebug* debug;             // W*H records or null
    // step-parallel path of the 1-of-16 dispatch (cloud_raymarch.cu): per-ray records and per-(step, ray) samples
    void* rays;                    // RaySetup[tx*ty], tile-major
    float2* samples;               // [MT_STEP_SLICES][tx*ty] (inc, energy)
    int* ctaSteps;                 // per 128-ray CTA: the largest step count among its rays (0 = all horizon-culled)
    unsigned* items;               // compacted in-cloud (step, ray) pairs: step << 26 | ray, in no particular order
    unsigned* itemCount;           // how many of them (zeroed by cloud_rays_kernel, filled by cloud_base_kernel)
    int rayStride;                 // rays per sample slice = 128 * CTAs of the ray grid
    unsigned* tileDone;            // full-quality row-tile launches with forwarding (mtSetCloudForward): CTAs finished per tile
    float2* decoded;               // fused 1-of-16 kernel: non-null = also keep the god-ray pass's decoded pair image current (below)
    int decodedPitch;              //   its row pitch in elements (mt_godray_pitch)
};

// Per-frame values of the post passes.  Like MarchConst they are evaluated by the HOST per dispatch (reproject_frame, godray_frame,
// txaa_frame in post_core.cuh are __host__ __device__: the same IEEE operations, the same bits) and travel in the parameter
// block.  Rounds 1 and 2 had thread 0 of every CTA evaluate them into shared memory -- a ~300-instruction serial prologue,
// seven IEEE divisions deep, behind a barrier the other 255 threads waited at (4.6 barrier-stall cycles per issued instruction
// in reproject_kernel, profiles/r2_passes_1080p.md).
struct ReprojFrame {  // reprojection.comp:203-211: camera basis, ray origin, the unit-sphere origin of the inner-shell intersection
    RayBasis basis;   // (raySphereIntersection's rO, identical for every pixel) and its C term, this shader's Halton offset
    f3 eye, ec, o;
    float C;
    float jx, jy;
    float uMax, vMax;  // (dim - 1) / dim: old_uv inside [0, max] keeps every tap inside the image (reproject_taps, fast path)
};
struct GodRayFrame {  // per-frame values of postProcess_GodRays.frag:74-91
    float blend;      // dot(normalize(sun - eye), camForward); < 0 => the pass writes nothing
    float sunx, suny; // clamped screen-space sun position
};
struct TxaaFrame {  // per-frame: like ReprojFrame, but with the Cloud pass's Halton variant (postProcess_TXAA.frag:63-82)
    RayBasis basis;
    f3 eye, ec, o;
    float C;
    float jx, jy;
};
// uv table of a W x H context (built by the host at mtCreate / mtResize: IEEE divisions, the kernels' own operands):
//   [0, W) x / W    [W, W+H) y / H    [W+H, 2W+H) (x + .5) / W    [2W+H, 2W+2H) (y + .5) / H
// so that a pixel's uv is two loads instead of two IEEE divisions (or a per-CTA staging pass behind a barrier).

struct ReprojParams {
    CamU cam, camOld;
    TimeU tm;
    ReprojFrame frame;
    const float* uv;   // uv table (above)
    const F4* prev;
    F4* cur;
    int W, H;
    int storage;  // MtStorage of the HDR / mask images (mt_pixel.cuh)
    int* taps;  // debug: 10 per pixel, or null
    int skipId;   // >= 0: leave out the pixels the frame's 1-of-16 Cloud dispatch with this id writes (pixel % 4 == (id / 4, id % 4) inside
    int tx, ty;   //   its tx x ty thread grid): mtFrameEx runs that dispatch BESIDE this pass on a second stream; -1: every pixel
};

struct GodRayParams {
    CamU cam;
    GodRayFrame frame;
    const float* uv;    // uv table (above)
    float lightColor[3];
    const F4* mask;     // encoded god-ray mask (RGBA32F)
    float2* decoded;    // pitch x (H+2) pairs (d(x, y), d(x+1, y)): the mask decoded per texel, ringed by the sampler's border value
    int pitch;          // elements per row of `decoded`: a power of two >= W + 2 on the device (mt_godray_log2pitch), W + 2 otherwise
    int log2pitch;      // 0: pitch is not a power of two (images wider than 8190 pixels, host simulation)
    const float2* tapRow0;  // generic path: decoded + pitch + 1 - MT_FLOOR_MAGIC_BITS: image texel (0, 0), biased for post_core.cuh's magic floor
    const float2* tapRow1;  // ... one row further
    const char* tapBase;    // power-of-two pitch: image texel (0, 0) minus what the magic constant's bits add to the byte offset
    const char* tapBaseWide; // the same for a 64-bit byte offset (A/B build MT_GODRAY_WIDE)
    F4* hdr;
    int W, H;
    int storage;  // MtStorage of the HDR / mask images (mt_pixel.cuh)
    int decodedCurrent; // the decoded image already matches the mask (the 1-of-16 Cloud kernel kept it so): no decode launch
    uint32_t* ldr;      // non-null: the tone map is fused into this pass's store (mtFrameEx with both passes): packed RGBA8 out
    unsigned seed;      //   uint(time.y) of the tone map's dither
};

// Row pitch of the decoded god-ray mask: the smallest power of two >= W + 2, at least 2^10 -- the tap address is then
// (floor(y) << k) + floor(x), one LEA, and the second row a constant byte offset (post_core.cuh).  0 = too wide, generic addressing.
#ifndef MT_GODRAY_POW2
#define MT_GODRAY_POW2 1
#endif
static inline int mt_godray_log2pitch(int W)
{
    if (!MT_GODRAY_POW2) return 0;
    for (int k = 10; k <= 13; ++k)
        if ((1 << k) >= W + 2) return k;
    return 0;
}
static inline size_t mt_godray_pitch(int W) { const int k = mt_godray_log2pitch(W); return k ? (size_t)1 << k : (size_t)W + 2; }

struct TxaaParams {
    CamU cam, camOld;
    TimeU tm;
    TxaaFrame frame;
    const float* uv;       // uv table (above)
    const uint32_t* cur;   // tone-mapped LDR of this frame (RGBA8)
    const uint32_t* prev;  // presented LDR of the previous frame
    uint32_t* out;         // result (becomes this frame's LDR image)
    int W, H;
};

struct ToneMapParams {
    int storage;    // MtStorage of the HDR image (mt_pixel.cuh)
    const F4* hdr;
    uint32_t* ldr;  // packed RGBA8
    int W, H;
    unsigned seed;  // uint(time.y)
};
