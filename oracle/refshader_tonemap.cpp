// TEST INFRASTRUCTURE -- the reference's tone-map fragment shader (postProcess_ToneMap.frag), compiled by g++ from its
// own text and run once per pixel (Renderer.cpp:834-838).  The LDR image is RGBA8 UNORM.
#define MTREF_FRAGMENT_STAGE
#include "glsl_rt.h"

namespace {
#include "postProcess_ToneMap.frag.inc"
}

extern "C" int mtrefsh_tonemap(const void* time76, int W, int H, const float* hdr, uint8_t* ldr, float* ldr_f32)
{
    const unsigned char* t = (const unsigned char*)time76;
    memcpy(&haltonSeq1, t, 16); memcpy(&haltonSeq2, t + 16, 16); memcpy(&haltonSeq3, t + 32, 16); memcpy(&haltonSeq4, t + 48, 16);
    memcpy(&time, t + 64, 8); memcpy(&frameCountMod16, t + 72, 4);
    inputImageSampler = { hdr, W, H, F32_TEXEL };
    currentFrameResultImage = { ldr_f32, W, H, nullptr, ldr };
    MTREF_FOR_EACH_FRAGMENT(W, H, shader_main())
    return 0;
}
