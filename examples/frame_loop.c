/* frame_loop.c -- the reference's main loop (main.cpp:126-194) on top of the C ABI, in plain C.
 *
 *   gcc -std=c11 -Iinclude examples/frame_loop.c -Lmeteoros_b200 -lmeteoros_b200 -Wl,-rpath,$PWD/meteoros_b200 -o frame_loop
 *   ./frame_loop <CloudTextures dir> [frames] [width height] [out.rgba8]
 *
 * <CloudTextures dir> is either the reference's texture directory (TGA slices + PNGs, decoded with the library's own decoders)
 * or a directory holding the packed caches low.mtvol / high.mtvol / curl.mtvol / weather.mtvol (mtxSaveVolume); the last
 * presented LDR frame is written to [out.rgba8] (W*H*4 bytes) when given.
 *
 * Loads the four noise inputs with the library's own decoders (Sky::CreateCloudResources), creates the renderer
 * (Renderer::InitializeRenderer), then per frame: pan the camera by 0.25 degrees (main.cpp:76-77), update time / sky,
 * run REPROJ + CLOUD + TONEMAP + TXAA (Renderer::Frame) and copy camera -> cameraOld.  Without an sm_100 GPU mtCreate
 * fails and the program says so (there is no CPU fallback).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "meteoros_b200.h"

static int check(MtContext* ctx, MtStatus st, const char* what)
{
    if (st == MT_OK) return 0;
    fprintf(stderr, "%s: %s (%s)\n", what, mtStatusString(st), ctx ? mtGetLastError(ctx) : "no context");
    return 1;
}

int main(int argc, char** argv)
{
    const char* dir = argc > 1 ? argv[1] : "textures/CloudTextures";
    const int frames = argc > 2 ? atoi(argv[2]) : 16;
    const uint32_t W = argc > 4 ? (uint32_t)atoi(argv[3]) : 1920, H = argc > 4 ? (uint32_t)atoi(argv[4]) : 1080;
    printf("meteoros_b200 ABI %u\n", mtAbiVersion());

    /* ---- assets (Sky.cpp:25-58) ---- */
    char path[1024];
    uint8_t* low = malloc(128u * 128 * 128 * 4);
    uint8_t* high = malloc(32u * 32 * 32 * 4);
    uint8_t* curl = malloc(128u * 128 * 4);
    uint8_t* weather = malloc(512u * 512 * 4);
    uint32_t w = 0, h = 0, d = 0;
    snprintf(path, sizeof path, "%s/low.mtvol", dir);
    FILE* cache = fopen(path, "rb");
    if (cache) { /* packed caches written by mtxSaveVolume */
        fclose(cache);
        if (check(NULL, mtxLoadVolume(path, low, 128u * 128 * 128 * 4, &w, &h, &d), "low.mtvol") || w != 128 || h != 128 || d != 128) return 2;
        snprintf(path, sizeof path, "%s/high.mtvol", dir);
        if (check(NULL, mtxLoadVolume(path, high, 32u * 32 * 32 * 4, &w, &h, &d), "high.mtvol") || w != 32 || h != 32 || d != 32) return 2;
        snprintf(path, sizeof path, "%s/curl.mtvol", dir);
        if (check(NULL, mtxLoadVolume(path, curl, 128u * 128 * 4, &w, &h, &d), "curl.mtvol") || w != 128 || h != 128 || d != 1) return 2;
        snprintf(path, sizeof path, "%s/weather.mtvol", dir);
        if (check(NULL, mtxLoadVolume(path, weather, 512u * 512 * 4, &w, &h, &d), "weather.mtvol") || w != 512 || h != 512 || d != 1) return 2;
        printf("assets loaded from .mtvol caches\n");
    } else {
        snprintf(path, sizeof path, "%s/LowFrequency/", dir);
        if (check(NULL, mtxLoadVolumeFromSlices(path, "LowFrequency", ".tga", 128, 128, 128, low, 128u * 128 * 128 * 4), "low-frequency volume")) return 2;
        snprintf(path, sizeof path, "%s/HighFrequency/", dir);
        if (check(NULL, mtxLoadVolumeFromSlices(path, "HighFrequency", ".tga", 32, 32, 32, high, 32u * 32 * 32 * 4), "high-frequency volume")) return 2;
        snprintf(path, sizeof path, "%s/curlNoise.png", dir);
        if (check(NULL, mtxLoadImageFile(path, curl, 128u * 128 * 4, &w, &h), "curl noise")) return 2;
        snprintf(path, sizeof path, "%s/weatherMap.png", dir);
        if (check(NULL, mtxLoadImageFile(path, weather, 512u * 512 * 4, &w, &h), "weather map")) return 2;
        printf("assets decoded\n");
    }

    /* ---- renderer (Renderer.cpp:89-110) ---- */
    MtConfig cfg = { sizeof(MtConfig), W, H, 0, MT_STORAGE_F32, 0 };
    MtContext* ctx = NULL;
    MtStatus st = mtCreate(&cfg, &ctx);
    if (st != MT_OK) {
        fprintf(stderr, "mtCreate: %s -- an sm_100 (B200) device is required, there is no CPU path\n", mtStatusString(st));
        return 3;
    }
    if (check(ctx, mtUploadTexture3D(ctx, MT_TEX_LOW_FREQ, 128, 128, 128, low), "upload low")) return 4;
    if (check(ctx, mtUploadTexture3D(ctx, MT_TEX_HIGH_FREQ, 32, 32, 32, high), "upload high")) return 4;
    if (check(ctx, mtUploadTexture2D(ctx, MT_TEX_CURL, 128, 128, curl), "upload curl")) return 4;
    if (check(ctx, mtUploadTexture2D(ctx, MT_TEX_WEATHER, 512, 512, weather), "upload weather")) return 4;

    /* ---- scene state (main.cpp:157-160) ---- */
    MtxCamera cam;
    MtCameraUBO cam_old;
    MtTimeUBO time;
    mtxCameraInit(&cam, (int32_t)W, (int32_t)H, NULL, NULL, 45.0f, 0.1f, 1000.0f);
    mtxCameraUBO(&cam, &cam_old);
    mtxTimeInit(&time);

    /* ---- frame loop (main.cpp:172-194) ---- */
    uint8_t* ldr = malloc((size_t)W * H * 4);
    for (int f = 0; f < frames; ++f) {
        mtxCameraRotateAboutUp(&cam, 0.25f);
        if (check(ctx, mtxRunFrame(ctx, &cam, &cam_old, &time, 1.0f / 60.0f, MT_FRAME_TONEMAP | MT_FRAME_TXAA), "frame")) return 5;
    }
    if (check(ctx, mtReadImage(ctx, MT_IMAGE_LDR_PREV, ldr, (size_t)W * H * 4), "read back")) return 5;
    if (argc > 5) {
        FILE* out = fopen(argv[5], "wb");
        if (!out || fwrite(ldr, 1, (size_t)W * H * 4, out) != (size_t)W * H * 4) { fprintf(stderr, "cannot write %s\n", argv[5]); return 6; }
        fclose(out);
    }
    unsigned long long sum = 0;
    for (size_t i = 0; i < (size_t)W * H * 4; ++i) sum += ldr[i];
    printf("%d frames of %ux%u rendered; mean LDR value %.2f; %llu kernel launches\n", frames, W, H, (double)sum / ((double)W * H * 4),
           (unsigned long long)mtLaunchCount(ctx));
    mtDestroy(ctx);
    free(low); free(high); free(curl); free(weather); free(ldr);
    return 0;
}
