// TEST INFRASTRUCTURE -- the reference's TXAA fragment shader (postProcess_TXAA.frag), compiled by g++ from its own
// text and run once per pixel (Renderer.cpp:840-846).
#define MTREF_FRAGMENT_STAGE
#include "glsl_rt.h"

namespace {
#include "postProcess_TXAA.frag.inc"
}

extern "C" int mtrefsh_txaa(const void* camera152, const void* cameraOld152, const void* time76, int W, int H, const uint8_t* cur,
                            const uint8_t* prev, uint8_t* out, float* out_f32)
{
    memcpy(&camera, camera152, 152);
    memcpy(&cameraOld, cameraOld152, 152);
    const unsigned char* t = (const unsigned char*)time76;
    memcpy(&haltonSeq1, t, 16); memcpy(&haltonSeq2, t + 16, 16); memcpy(&haltonSeq3, t + 32, 16); memcpy(&haltonSeq4, t + 48, 16);
    memcpy(&time, t + 64, 8); memcpy(&frameCountMod16, t + 72, 4);
    prevFrameImage = { prev, W, H, UNORM8_BORDER };
    currentFrameResultImage = { out_f32, W, H, cur, out };
    MTREF_FOR_EACH_FRAGMENT(W, H, shader_main())
    return 0;
}
