// TEST INFRASTRUCTURE -- the reference's Reprojection compute shader (reprojection.comp), compiled by g++ from its own
// text (oracle/glsl2cpp.py, oracle/glsl_rt.h) and dispatched the way Renderer.cpp:683-698 dispatches it.
#include "glsl_rt.h"

namespace {
#include "reprojection.comp.inc"
}

extern "C" int mtrefsh_reproject(const void* camera152, const void* cameraOld152, const void* time76, int W, int H, float* prev, float* cur)
{
    static_assert(sizeof(camera) == 152 && sizeof(cameraOld) == 152, "uniform block layouts");
    memcpy(&camera, camera152, 152);
    memcpy(&cameraOld, cameraOld152, 152);
    const unsigned char* t = (const unsigned char*)time76;
    memcpy(&haltonSeq1, t, 16); memcpy(&haltonSeq2, t + 16, 16); memcpy(&haltonSeq3, t + 32, 16); memcpy(&haltonSeq4, t + 48, 16);
    memcpy(&time, t + 64, 8); memcpy(&frameCountMod16, t + 72, 4);
    currentFrameResultImage = { cur, W, H, nullptr, nullptr };
    previousFrameResultImage = { prev, W, H, nullptr, nullptr };
    const uint32_t bx = ((uint32_t)W + 32 - 1) / 32, by = ((uint32_t)H + 32 - 1) / 32;  // Renderer.cpp:683-685
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t gy = 0; gy < (int64_t)by * 32; ++gy)
        for (uint32_t gx = 0; gx < bx * 32; ++gx) {
            gl_GlobalInvocationID = uvec3(gx, (uint32_t)gy, 0u);
            shader_main();
        }
    return 0;
}
