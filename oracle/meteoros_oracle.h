/*
 * meteoros_oracle.h -- interface of the CPU parity oracle (TEST INFRASTRUCTURE, see meteoros_oracle.c).
 * Host pointers only; images are tightly packed RGBA32F [y][x][4]; textures RGBA8 [z][y][x][4].
 * Every function returns 0 on success, 1 on invalid arguments.
 */
#ifndef METEOROS_ORACLE_H
#define METEOROS_ORACLE_H

#include <stdint.h>
#include "../include/meteoros_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct MtoTextures {
    const uint8_t* low;     int32_t low_w, low_h, low_d;       /* cloudBaseShapeSampler        */
    const uint8_t* high;    int32_t high_w, high_h, high_d;    /* cloudDetailsHighFreqSampler  */
    const uint8_t* curl;    int32_t curl_w, curl_h;            /* curlNoiseSampler             */
    const uint8_t* weather; int32_t weather_w, weather_h;      /* weatherMapSampler (never sampled) */
} MtoTextures;

/* cloudRayMarch.comp main() over the reference grid (Renderer.cpp:713-716).  full == 0: one dispatch with
 * pixelID = tm->frameCountMod16; full != 0: the union of the 16 dispatches.  Only pixel rows in
 * [row_begin, row_end) are produced (row_end <= 0 means H), and of the 4-row groups only every group_stride-th
 * (1 = all) -- both used for bounded CPU-baseline samples spread evenly over the frame.
 * counters / debug may be NULL; debug holds W*H records.                                               */
int mto_cloud(const MtCameraUBO* cam, const MtTimeUBO* tm, const MtTuning* tun, const MtoTextures* tex, int W, int H,
              int full, int row_begin, int row_end, int group_stride, float* hdr, float* mask, MtCounters* counters,
              MtRayDebug* debug);
/* reprojection.comp main(); taps (optional) receives 10 clamped linear indices per pixel. */
int mto_reproject(const MtCameraUBO* cam, const MtCameraUBO* camOld, const MtTimeUBO* tm, int W, int H, const float* prev,
                  float* cur, int32_t* taps);
/* postProcess_GodRays.frag main(); hdr is read-modify-written. */
int mto_godrays(const MtCameraUBO* cam, const MtSunAndSkyUBO* sky, int W, int H, const float* mask, float* hdr);
/* postProcess_ToneMap.frag main(); ldr = RGBA8 UNORM (may be NULL), ldr_f32 = unquantised RGBA32F (may be NULL). */
int mto_tonemap(const MtTimeUBO* tm, int W, int H, const float* hdr, uint8_t* ldr, float* ldr_f32);

/* postProcess_TXAA.frag main() (the pass after tone map in the reference frame): cur = tone-mapped LDR of this frame,
 * prev = presented LDR of the previous frame; out (RGBA8) and/or out_f32 (unquantised) receive the result. */
int mto_txaa(const MtCameraUBO* cam, const MtCameraUBO* camOld, const MtTimeUBO* tm, int W, int H, const uint8_t* cur,
             const uint8_t* prev, uint8_t* out, float* out_f32);

/* OpenMP team size: set (n > 0) / query the threads the loops above really run on (bench.py's CPU timing legs). */
int mto_set_num_threads(int n);
int mto_num_threads(void);

/* unit-test hooks */
void mto_sample3d(const uint8_t* vol, int W, int H, int D, float s, float t, float r, float out[4]);
void mto_sample2d(const uint8_t* img, int W, int H, float s, float t, float out[4]);
uint32_t mto_wang_hash(uint32_t u, uint32_t v, uint32_t s);
void mto_encode_float_rgba(float v, float out[4]);
void mto_ray_sphere(const float ro[3], const float rd[3], const float c[3], float radius, float point[3], float* t, int* valid);
void mto_cloud_grid(int W, int H, int* threads_x, int* threads_y);
void mto_sample2d_f32_border(const float* img, int W, int H, float s, float t, float out[4]);      /* LINEAR, CLAMP_TO_BORDER (0,0,0,1) */
void mto_sample2d_unorm8_border(const uint8_t* img, int W, int H, float s, float t, float out[4]); /* same, RGBA8 UNORM image  */
void mto_load_unorm8(const uint8_t* img, int W, int H, int x, int y, float out[4]);               /* imageLoad on an rgba8 image */
uint8_t mto_to_unorm8(float v);                                                                    /* imageStore conversion      */
float mto_density_height_gradient(float relativeHeight, float cloudType);
void mto_atmosphere_color(const float dir[3], const float sun_minus_origin[3], float sunIntensity, const float skySun[3], float out[3]);

#ifdef __cplusplus
}
#endif
#endif
