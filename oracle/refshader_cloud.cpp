// TEST INFRASTRUCTURE -- the reference's Cloud compute shader (cloudRayMarch.comp), compiled by g++ from its own text
// (oracle/glsl2cpp.py, oracle/glsl_rt.h) and dispatched the way Renderer.cpp:701-716 dispatches it.
//
// -DMTREF_WEATHER builds the same file around the shader with its dead weather-map block revived
// (glsl2cpp.py --revive-weather) and exports mtrefsh_cloud_weather: the pin for MtTuning.use_weather.
#include "glsl_rt.h"

namespace {
#ifdef MTREF_WEATHER
float mt_weather_scale = 1.0f;
#include "cloudRayMarch.weather.inc"
#else
#include "cloudRayMarch.comp.inc"
#endif
}
#undef E
#undef PI

#ifdef MTREF_WEATHER
#define MTREF_CLOUD_ENTRY mtrefsh_cloud_weather
#else
#define MTREF_CLOUD_ENTRY mtrefsh_cloud
#endif

extern "C" int MTREF_CLOUD_ENTRY(const void* camera152, const void* time76, const void* sky52, const uint8_t* low, int lw, int lh, int ld,
                             const uint8_t* high, int hw, int hh, int hd, const uint8_t* curl, int cw, int ch, const uint8_t* weather,
                             int ww, int wh, int W, int H, float* prev, float* hdr, float* mask, int group_stride, float weather_scale)
{
#ifdef MTREF_WEATHER
    mt_weather_scale = weather_scale;
#endif
    static_assert(sizeof(camera) == 152 && sizeof(sunAndSky) == 52, "uniform block layouts");
    memcpy(&camera, camera152, 152);
    const unsigned char* t = (const unsigned char*)time76;
    memcpy(&haltonSeq1, t, 16); memcpy(&haltonSeq2, t + 16, 16); memcpy(&haltonSeq3, t + 32, 16); memcpy(&haltonSeq4, t + 48, 16);
    memcpy(&time, t + 64, 8); memcpy(&frameCountMod16, t + 72, 4);
    memcpy(&sunAndSky, sky52, 52);
    cloudBaseShapeSampler = { low, lw, lh, ld };
    cloudDetailsHighFreqSampler = { high, hw, hh, hd };
    curlNoiseSampler = { curl, cw, ch, UNORM8_REPEAT };
    weatherMapSampler = { weather, ww, wh, UNORM8_REPEAT };
    currentFrameResultImage = { hdr, W, H, nullptr, nullptr };
    previousFrameResultImage = { prev, W, H, nullptr, nullptr };
    godRaysCreationDataImage = { mask, W, H, nullptr, nullptr };
    // Renderer.cpp:711-716: numBlocks = (std::ceil(window / 4) + 32 - 1) / 32 with an integer division inside ceil,
    // converted to uint32_t; 32 x 32 invocations per group (cloudRayMarch.comp:4-5)
    const uint32_t bx = (uint32_t)((std::ceil(W / 4) + 32 - 1) / 32), by = (uint32_t)((std::ceil(H / 4) + 32 - 1) / 32);
    // group_stride > 1 (bench.py's bounded CPU sample): only every group_stride-th row of invocations (= 4 pixel rows)
    const int64_t stride = group_stride > 1 ? group_stride : 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t gy = 0; gy < (int64_t)by * 32; gy += stride)
        for (uint32_t gx = 0; gx < bx * 32; ++gx) {
            gl_GlobalInvocationID = uvec3(gx, (uint32_t)gy, 0u);
            shader_main();
        }
    return 0;
}
