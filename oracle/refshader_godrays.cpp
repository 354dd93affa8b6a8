// TEST INFRASTRUCTURE -- the reference's god-ray fragment shader (postProcess_GodRays.frag), compiled by g++ from its
// own text and run once per pixel, as the full-screen triangle of Renderer.cpp:826-832 would.
#define MTREF_FRAGMENT_STAGE
#include "glsl_rt.h"

namespace {
#include "postProcess_GodRays.frag.inc"
}

extern "C" int mtrefsh_godrays(const void* camera152, const void* sky52, int W, int H, const float* mask, float* hdr)
{
    memcpy(&camera, camera152, 152);
    memcpy(&sunAndSky, sky52, 52);
    currentFrameResultImage = { hdr, W, H, nullptr, nullptr };
    godRayCreationDataSampler = { mask, W, H, F32_BORDER };
    MTREF_FOR_EACH_FRAGMENT(W, H, shader_main())
    return 0;
}
