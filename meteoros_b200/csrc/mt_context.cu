// mt_context.cu -- host side of the C ABI (include/meteoros_b200.h): the part of Meteoros' Renderer that a CUDA
// replacement substitutes -- resource creation (Renderer.cpp:1428-1447), uniform / texture binding
// (Renderer.cpp:914-1165), the per-frame dispatch order and ping-pong (Renderer.cpp:122-192, 653-722, 823-846).
// One context = one device + one stream; calls execute in order on that stream.  No CPU fallback exists: every
// dispatch either launches the sm_100a kernels or returns an error.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "cloud_core.cuh"
#include "mt_host_consts.h"
#include "post_core.cuh"

#ifndef MT_FRAME_OVERLAP
#define MT_FRAME_OVERLAP 1  /* mtFrameEx: reprojection and the 1-of-16 Cloud dispatch side by side on two streams */
#endif
#ifndef MT_INCREMENTAL_DECODE
#define MT_INCREMENTAL_DECODE 1  /* the fused 1-of-16 Cloud kernel keeps the god-ray pass's decoded mask current (cloud_raymarch.cu) */
#endif
#include "mt_launch.h"

#define MT_FLAG_PASS_TIMING_INTERNAL MT_FLAG_PASS_TIMING
#define MT_USER_EVENTS 16

struct MtContext {
    int device = 0;
    int W = 0, H = 0;
    uint32_t storage = MT_STORAGE_F32;
    uint32_t flags = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;   // mtReadImageAsync
    cudaStream_t sideStream = nullptr;   // mtFrameEx: the 1-of-16 Cloud dispatch runs here, beside the reprojection on `stream`
    cudaEvent_t sideForkEv = nullptr, sideJoinEv = nullptr;
    int reprojSkipId = -1;               // set by mtFrameEx around its reprojection dispatch
    cudaStream_t fwdStream = nullptr;    // mtSetCloudForward: high-priority stream of tile_forward_kernel
    cudaEvent_t fwdArmEv = nullptr, fwdDoneEv = nullptr;
    void* forwardHdr = nullptr;          //   peer image the finished row tiles are pushed to (NULL = off)
    unsigned* tileDone = nullptr;        //   per-tile completion counters
    bool fwdBusy = false;
    bool fwdCheck = false;               //   a forwarder ran since the last mtSynchronize: read its overrun flag there
    int fwdTiles = 0;
    cudaEvent_t producedEv = nullptr;    // main stream -> copy stream
    struct PendingRead { const void* dev; cudaEvent_t done; bool active; } pending[4] = {};
    F4* hdr[2] = { nullptr, nullptr };
    int cur = 0;  // index of the image that currently plays "currentFrameResultImage"
    F4* mask = nullptr;
    float2* maskDecoded = nullptr; // pitch x (H+2) pairs: scratch of the god-ray pass
    float* uvTab = nullptr;        // per-size uv table of the post passes (mt_params.h)
    bool decodedCurrent = false;   // maskDecoded matches the mask: set by a god-ray dispatch, kept by the fused 1-of-16 Cloud kernel (which
                                   // updates the pairs it touches), cleared by every other writer of the mask
    bool maskShared = false;       // the mask's device pointer / IPC handle has been handed out: writes can no longer be tracked
    F4* maskStage = nullptr;       // device snapshot of the mask behind mtReadImageAsync (lazily allocated)
    float* greyStage = nullptr;    // the decoded one-float-per-pixel god-ray image behind mtReadGodRayGreyAsync (lazily allocated)
    uint32_t* ldr[2] = { nullptr, nullptr };  // ping-pong with the HDR images (same `cur`)
    uint32_t* ldrScratch = nullptr;           // TXAA output, swapped with ldr[cur] after the pass
    uint32_t* tex[4] = { nullptr, nullptr, nullptr, nullptr };
    void* quads[4] = { nullptr, nullptr, nullptr, nullptr };  // per-cell 2x2 texel quads (mt_tex.cuh), built at upload
    void* rfQuads = nullptr;      // low-frequency volume only: the quads in (r, F) form for the light-cone samples
    cudaArray_t hwArray = nullptr;          // MT_FLAG_HW_CONE_FILTER (mt_tex.cuh): the low-frequency volume as a CUDA array + texture object
    cudaTextureObject_t hwTex = 0;
    int texw[4] = { 0, 0, 0, 0 }, texh[4] = { 0, 0, 0, 0 }, texd[4] = { 0, 0, 0, 0 };
    uint32_t* occ = nullptr;      // empty-cell bitmap of the low-frequency volume (mt_tex.cuh)
    float occCoverage = -1.0f;    // coverage the bitmap was built for; < 0 = stale
    unsigned long long* counters = nullptr;
    void* rays = nullptr;         // step-parallel 1/16 path: RaySetup per ray (lazily allocated)
    float2* samples = nullptr;    //   and (inc, energy) per (step, ray)
    int* ctaSteps = nullptr;      //   and the per-CTA maximum step count
    unsigned* items = nullptr;    //   and the compacted in-cloud (step, ray) list + its length
    unsigned* itemCount = nullptr;
    MtRayDebug* debug = nullptr;  // lazily allocated W*H records
    int* taps = nullptr;          // lazily allocated W*H*10
    MtCameraUBO cam, camOld;
    MtTimeUBO tm;
    MtSunAndSkyUBO sky;
    MtTuning tun;
    int32_t key = 0;
    bool haveCam = false, haveCamOld = false, haveTime = false;
    int storeMode = 0;     // mtSetCloudStoreMode
    F4* outHdr = nullptr;  // mtSetCloudOutput overrides
    F4* outMask = nullptr;
    cudaEvent_t ev[MT_PASS_COUNT][2] = {};
    bool evValid[MT_PASS_COUNT] = {};
    cudaEvent_t userEv[MT_USER_EVENTS] = {};
    bool userEvValid[MT_USER_EVENTS] = {};
    uint64_t launches = 0;
    void* flushBuf = nullptr;
    size_t flushBytes = 0;
    std::string err;
};

static MtStatus fail(MtContext* c, MtStatus s, const std::string& msg)
{
    if (c) c->err = msg;
    return s;
}
static MtStatus cuda_fail(MtContext* c, cudaError_t e, const char* what)
{
    if (e == cudaErrorMemoryAllocation) return fail(c, MT_ERR_OOM, std::string(what) + ": " + cudaGetErrorString(e));
    return fail(c, MT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define MT_CUDA(c, call)                                    \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return cuda_fail((c), e__, #call); \
    } while (0)
#define MT_REQUIRE(c, cond, msg)                            \
    do {                                                    \
        if (!(cond)) return fail((c), MT_ERR_INVALID, (msg)); \
    } while (0)

static size_t hdr_pixel_bytes(uint32_t storage) { return storage == MT_STORAGE_F16 ? 8 : 16; }
static size_t pixel_bytes(const MtContext* c, MtImage w)
{
    return (w == MT_IMAGE_LDR || w == MT_IMAGE_LDR_PREV) ? 4 : hdr_pixel_bytes(c->storage);
}
static size_t image_bytes(const MtContext* c, MtImage w) { return (size_t)c->W * (size_t)c->H * pixel_bytes(c, w); }
static void* image_ptr(MtContext* c, MtImage w)
{
    switch (w) {
        case MT_IMAGE_CLOUD_CUR: return c->hdr[c->cur];
        case MT_IMAGE_CLOUD_PREV: return c->hdr[c->cur ^ 1];
        case MT_IMAGE_GODRAY_MASK: return c->mask;
        case MT_IMAGE_LDR: return c->ldr[c->cur];
        case MT_IMAGE_LDR_PREV: return c->ldr[c->cur ^ 1];
    }
    return nullptr;
}
// A pass is about to write device memory `dev`: make the main stream wait for an in-flight asynchronous read of it.
static void wait_pending_read(MtContext* c, const void* dev)
{
    for (auto& p : c->pending)
        if (p.active && p.dev == dev) {
            cudaStreamWaitEvent(c->stream, p.done, 0);
            p.active = false;
        }
}
static bool is_pow2(uint32_t v) { return v && !(v & (v - 1)); }

// The per-size device images of a context.  Allocated as a group so that mtCreate / mtResize are transactional: either
// every image of the new size exists, or nothing changed.
struct ImageSet {
    F4* hdr[2] = { nullptr, nullptr };
    F4* mask = nullptr;
    uint32_t* ldr[2] = { nullptr, nullptr };
    uint32_t* ldrScratch = nullptr;
    float2* maskDecoded = nullptr;
    float* uvTab = nullptr;  // mt_params.h: x / W, y / H, (x + .5) / W, (y + .5) / H
};
static void free_image_set(ImageSet& s)
{
    cudaFree(s.hdr[0]); cudaFree(s.hdr[1]); cudaFree(s.mask);
    cudaFree(s.ldr[0]); cudaFree(s.ldr[1]); cudaFree(s.ldrScratch); cudaFree(s.maskDecoded); cudaFree(s.uvTab);
    s = ImageSet();
}
static cudaError_t alloc_image_set(ImageSet& s, int W, int H, size_t hb, cudaStream_t stream)  // hb: bytes per HDR / mask pixel
{
    const size_t px = (size_t)W * (size_t)H;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&s.hdr[0], px * hb)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&s.hdr[1], px * hb)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&s.mask, px * hb)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&s.ldr[0], px * 4)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&s.ldr[1], px * 4)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&s.ldrScratch, px * 4)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&s.maskDecoded, mt_godray_pitch(W) * (size_t)(H + 2) * sizeof(float2))) != cudaSuccess) return e;
    {   // the uv table: the kernels' own operands, divided here (IEEE, as the shaders' in_uv / pixelPos arithmetic)
        std::vector<float> uv((size_t)2 * (W + H));
        for (int x = 0; x < W; ++x) { uv[x] = (float)x / (float)W; uv[(size_t)W + H + x] = ((float)x + 0.5f) / (float)W; }
        for (int y = 0; y < H; ++y) { uv[(size_t)W + y] = (float)y / (float)H; uv[(size_t)2 * W + H + y] = ((float)y + 0.5f) / (float)H; }
        if ((e = cudaMalloc((void**)&s.uvTab, uv.size() * sizeof(float))) != cudaSuccess) return e;
        if ((e = cudaMemcpy(s.uvTab, uv.data(), uv.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return e;  // synchronous: `uv` dies here
    }
    if ((e = cudaMemsetAsync(s.hdr[0], 0, px * hb, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s.hdr[1], 0, px * hb, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s.mask, 0, px * hb, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s.ldr[0], 0, px * 4, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s.ldr[1], 0, px * 4, stream)) != cudaSuccess) return e;
    return cudaMemsetAsync(s.ldrScratch, 0, px * 4, stream);
}

// Everything whose size or meaning depends on W x H: the images, the lazily allocated scratch of the step-parallel
// march / debug records, and all state that refers to the old images (output redirection, forwarding, pending reads).
static void free_images(MtContext* c)
{
    if (c->fwdStream) cudaStreamSynchronize(c->fwdStream);
    ImageSet old;
    old.hdr[0] = c->hdr[0]; old.hdr[1] = c->hdr[1]; old.mask = c->mask;
    old.ldr[0] = c->ldr[0]; old.ldr[1] = c->ldr[1]; old.ldrScratch = c->ldrScratch; old.maskDecoded = c->maskDecoded; old.uvTab = c->uvTab;
    free_image_set(old);
    c->ldr[0] = c->ldr[1] = c->ldrScratch = nullptr;
    c->hdr[0] = c->hdr[1] = c->mask = nullptr;
    c->maskDecoded = nullptr;
    c->uvTab = nullptr;
    cudaFree(c->debug); cudaFree(c->taps); cudaFree(c->rays); cudaFree(c->samples); cudaFree(c->ctaSteps);
    cudaFree(c->items); cudaFree(c->itemCount);
    cudaFree(c->tileDone);
    cudaFree(c->maskStage);
    cudaFree(c->greyStage);
    c->maskStage = nullptr;
    c->greyStage = nullptr;
    c->tileDone = nullptr; c->fwdBusy = false; c->fwdCheck = false; c->fwdTiles = 0;
    c->forwardHdr = nullptr;            // mapped for the old size: the peer must re-export, the caller re-arm (mtSetCloudForward)
    c->outHdr = c->outMask = nullptr;   // same for mtSetCloudOutput
    for (auto& p : c->pending) p.active = false;
    c->rays = nullptr; c->samples = nullptr; c->ctaSteps = nullptr; c->items = nullptr; c->itemCount = nullptr;
    c->debug = nullptr; c->taps = nullptr;
}
static void adopt_images(MtContext* c, const ImageSet& s)
{
    c->hdr[0] = s.hdr[0]; c->hdr[1] = s.hdr[1]; c->mask = s.mask;
    c->ldr[0] = s.ldr[0]; c->ldr[1] = s.ldr[1]; c->ldrScratch = s.ldrScratch; c->maskDecoded = s.maskDecoded; c->uvTab = s.uvTab;
    c->decodedCurrent = false;
    c->maskShared = false;
    c->cur = 0;
}
static MtStatus alloc_images(MtContext* c)
{
    ImageSet s;
    const cudaError_t e = alloc_image_set(s, c->W, c->H, hdr_pixel_bytes(c->storage), c->stream);
    if (e != cudaSuccess) {
        free_image_set(s);
        (void)cudaGetLastError();
        return cuda_fail(c, e, "image allocation");
    }
    adopt_images(c, s);
    return MT_OK;
}

// No exception crosses the C boundary (std::string / std::vector members can throw): every MtStatus entry point is a
// function-try-block ending in MT_NOTHROW.
#define MT_NOTHROW                                      \
    catch (const std::bad_alloc&) { return MT_ERR_OOM; } \
    catch (...) { return MT_ERR_INVALID; }

extern "C" {

uint32_t mtAbiVersion(void) { return MT_ABI_VERSION; }

const char* mtStatusString(MtStatus s)
{
    switch (s) {
        case MT_OK: return "MT_OK";
        case MT_ERR_INVALID: return "MT_ERR_INVALID";
        case MT_ERR_CUDA: return "MT_ERR_CUDA";
        case MT_ERR_OOM: return "MT_ERR_OOM";
        case MT_ERR_UNSUPPORTED_ARCH: return "MT_ERR_UNSUPPORTED_ARCH";
        case MT_ERR_NOT_READY: return "MT_ERR_NOT_READY";
    }
    return "MT_ERR_UNKNOWN";
}

void mtDefaultTuning(MtTuning* t)
{
    if (!t) return;
    t->coverage = 0.6f;
    t->sun_location[0] = 0.0f;
    t->sun_location[1] = MT_R_OUTER * 0.9f;
    t->sun_location[2] = -MT_R_OUTER * 0.9f;
    t->sky_sun_location[0] = 0.0f;
    t->sky_sun_location[1] = MT_EARTH_RADIUS * 2.0f;
    t->sky_sun_location[2] = -MT_EARTH_RADIUS * 10.0f;
    t->wind_direction[0] = 1.0f; t->wind_direction[1] = 0.0f; t->wind_direction[2] = 0.0f;
    t->cloud_speed = 0.080f;
    t->cloud_top_offset = 1.0f;
    t->base_density_factor = 0.380f;
    t->use_weather = 0u;
    t->weather_scale = 1.0f;
}

MtStatus mtCreate(const MtConfig* cfg, MtContext** out)
try {
    if (!cfg || !out) return MT_ERR_INVALID;
    *out = nullptr;
    if (cfg->struct_size != sizeof(MtConfig) || cfg->width == 0 || cfg->height == 0 || cfg->width > 32768 ||
        cfg->height > 32768 || cfg->storage > MT_STORAGE_F16)
        return MT_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) return MT_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return MT_ERR_CUDA;
    if (prop.major != 10) return MT_ERR_UNSUPPORTED_ARCH;  // the kernels exist as sm_100a SASS only
    MtContext* c = new (std::nothrow) MtContext();
    if (!c) return MT_ERR_OOM;
    c->device = cfg->device;
    c->W = (int)cfg->width;
    c->H = (int)cfg->height;
    c->storage = cfg->storage;
    c->flags = cfg->flags;
    memset(&c->cam, 0, sizeof(c->cam)); memset(&c->camOld, 0, sizeof(c->camOld));
    memset(&c->tm, 0, sizeof(c->tm)); memset(&c->sky, 0, sizeof(c->sky));
    mtDefaultTuning(&c->tun);
    MtStatus st = MT_OK;
    do {
        if (cudaSetDevice(c->device) != cudaSuccess) { st = MT_ERR_CUDA; break; }
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { st = MT_ERR_CUDA; break; }
        if (cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking) != cudaSuccess) { st = MT_ERR_CUDA; break; }
        if (cudaStreamCreateWithFlags(&c->sideStream, cudaStreamNonBlocking) != cudaSuccess) { st = MT_ERR_CUDA; break; }
        if (cudaEventCreateWithFlags(&c->sideForkEv, cudaEventDisableTiming) != cudaSuccess) { st = MT_ERR_CUDA; break; }
        if (cudaEventCreateWithFlags(&c->sideJoinEv, cudaEventDisableTiming) != cudaSuccess) { st = MT_ERR_CUDA; break; }
        if (cudaEventCreateWithFlags(&c->producedEv, cudaEventDisableTiming) != cudaSuccess) { st = MT_ERR_CUDA; break; }
        for (auto& p : c->pending)
            if (cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming) != cudaSuccess) st = MT_ERR_CUDA;
        if (st != MT_OK) break;
        for (int p = 0; p < MT_PASS_COUNT; ++p) {
            if (cudaEventCreate(&c->ev[p][0]) != cudaSuccess || cudaEventCreate(&c->ev[p][1]) != cudaSuccess) st = MT_ERR_CUDA;
        }
        if (st != MT_OK) break;
        if ((st = alloc_images(c)) != MT_OK) break;
        if (cudaMalloc((void**)&c->counters, 8 * sizeof(unsigned long long)) != cudaSuccess) { st = MT_ERR_OOM; break; }
        if (cudaMemsetAsync(c->counters, 0, 8 * sizeof(unsigned long long), c->stream) != cudaSuccess) { st = MT_ERR_CUDA; break; }
    } while (0);
    if (st != MT_OK) {
        mtDestroy(c);
        return st;
    }
    *out = c;
    return MT_OK;
} MT_NOTHROW

void mtDestroy(MtContext* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copyStream) { cudaStreamSynchronize(c->copyStream); cudaStreamDestroy(c->copyStream); }
    if (c->sideStream) { cudaStreamSynchronize(c->sideStream); cudaStreamDestroy(c->sideStream); }
    if (c->sideForkEv) cudaEventDestroy(c->sideForkEv);
    if (c->sideJoinEv) cudaEventDestroy(c->sideJoinEv);
    if (c->fwdStream) { cudaStreamSynchronize(c->fwdStream); cudaStreamDestroy(c->fwdStream); c->fwdStream = nullptr; }
    if (c->fwdArmEv) cudaEventDestroy(c->fwdArmEv);
    if (c->fwdDoneEv) cudaEventDestroy(c->fwdDoneEv);
    if (c->producedEv) cudaEventDestroy(c->producedEv);
    for (auto& p : c->pending)
        if (p.done) cudaEventDestroy(p.done);
    free_images(c);
    for (int i = 0; i < 4; ++i) { cudaFree(c->tex[i]); cudaFree(c->quads[i]); }
    cudaFree(c->rfQuads);
    if (c->hwTex) cudaDestroyTextureObject(c->hwTex);
    if (c->hwArray) cudaFreeArray(c->hwArray);
    cudaFree(c->occ);
    cudaFree(c->counters);
    cudaFree(c->flushBuf);
    for (int p = 0; p < MT_PASS_COUNT; ++p) {
        if (c->ev[p][0]) cudaEventDestroy(c->ev[p][0]);
        if (c->ev[p][1]) cudaEventDestroy(c->ev[p][1]);
    }
    for (int i = 0; i < MT_USER_EVENTS; ++i)
        if (c->userEv[i]) cudaEventDestroy(c->userEv[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* mtGetLastError(const MtContext* c) { return c ? c->err.c_str() : "null context"; }

MtStatus mtResize(MtContext* c, uint32_t w, uint32_t h)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, w > 0 && h > 0 && w <= 32768 && h <= 32768, "mtResize: bad size");
    MT_CUDA(c, cudaSetDevice(c->device));
    MT_CUDA(c, cudaStreamSynchronize(c->stream));
    MT_CUDA(c, cudaStreamSynchronize(c->copyStream));
    // Transactional: the images of the new size are allocated BEFORE anything of the old size is released.  On failure the
    // context keeps its old size, images and peer state and stays fully usable (the caller sees MT_ERR_OOM / MT_ERR_CUDA).
    ImageSet s;
    const cudaError_t e = alloc_image_set(s, (int)w, (int)h, hdr_pixel_bytes(c->storage), c->stream);
    if (e != cudaSuccess) {
        free_image_set(s);
        (void)cudaGetLastError();
        return cuda_fail(c, e, "mtResize: image allocation (context unchanged)");
    }
    // Success: drop everything tied to the old size.  Output redirection (mtSetCloudOutput), forwarding (mtSetCloudForward)
    // and exported IPC handles referred to images of the old size: they are cleared here, and peers must re-export /
    // re-open their handles and re-arm (include/meteoros_b200.h, mtResize).
    free_images(c);
    c->W = (int)w;
    c->H = (int)h;
    adopt_images(c, s);
    return MT_OK;
} MT_NOTHROW

MtStatus mtSetCamera(MtContext* c, const MtCameraUBO* u)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, u != nullptr, "mtSetCamera: null ubo");
    c->cam = *u;
    c->haveCam = true;
    return MT_OK;
} MT_NOTHROW
MtStatus mtSetCameraOld(MtContext* c, const MtCameraUBO* u)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, u != nullptr, "mtSetCameraOld: null ubo");
    c->camOld = *u;
    c->haveCamOld = true;
    return MT_OK;
} MT_NOTHROW
MtStatus mtSetTime(MtContext* c, const MtTimeUBO* u)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, u != nullptr, "mtSetTime: null ubo");
    MT_REQUIRE(c, u->frameCountMod16 >= 0 && u->frameCountMod16 < 16, "mtSetTime: frameCountMod16 outside 0..15");
    c->tm = *u;
    c->haveTime = true;
    return MT_OK;
} MT_NOTHROW
MtStatus mtSetSunAndSky(MtContext* c, const MtSunAndSkyUBO* u)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, u != nullptr, "mtSetSunAndSky: null ubo");
    c->sky = *u;
    return MT_OK;
} MT_NOTHROW
MtStatus mtSetKeyPressQuery(MtContext* c, int32_t k)
try {
    if (!c) return MT_ERR_INVALID;
    c->key = k;  // bound at set 5 of the cloud pipeline, never read by the shader (Renderer.cpp:706)
    return MT_OK;
} MT_NOTHROW
MtStatus mtSetTuning(MtContext* c, const MtTuning* t)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, t != nullptr, "mtSetTuning: null tuning");
    MT_REQUIRE(c, t->coverage >= 0.0f && t->coverage <= 0.91f, "mtSetTuning: coverage must be in [0, 0.91]");
    MT_REQUIRE(c, t->weather_scale == t->weather_scale, "mtSetTuning: weather_scale is NaN");
    c->tun = *t;
    return MT_OK;
} MT_NOTHROW

static MtStatus upload(MtContext* c, int slot, uint32_t w, uint32_t h, uint32_t d, const uint8_t* rgba8)
{
    MT_REQUIRE(c, rgba8 != nullptr, "texture upload: null data");
    MT_REQUIRE(c, is_pow2(w) && is_pow2(h) && is_pow2(d) && w <= 2048 && h <= 2048 && d <= 2048,
               "texture upload: extents must be powers of two <= 2048 (reference: 128^3, 32^3, 128^2, 512^2)");
    MT_CUDA(c, cudaSetDevice(c->device));
    size_t bytes = (size_t)w * h * d * 4;
    if (c->tex[slot]) {
        MT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->tex[slot]);
        c->tex[slot] = nullptr;
    }
    MT_CUDA(c, cudaMalloc((void**)&c->tex[slot], bytes));
    MT_CUDA(c, cudaMemcpyAsync(c->tex[slot], rgba8, bytes, cudaMemcpyHostToDevice, c->stream));
    cudaFree(c->quads[slot]);
    c->quads[slot] = nullptr;
    {
        MT_CUDA(c, cudaMalloc(&c->quads[slot], bytes * 4 * ((MT_TEX_BRICKS && d > 1) ? 2 : 1)));
        MT_CUDA(c, mt_launch_build_quads(c->tex[slot], (int)w, (int)h, (int)d, c->quads[slot], c->stream));
        c->launches += 1;
    }
    if (slot == MT_TEX_LOW_FREQ) {
        cudaFree(c->rfQuads);
        c->rfQuads = nullptr;
        if (!(c->flags & MT_FLAG_NO_CONE_RF)) {
            MT_CUDA(c, cudaMalloc(&c->rfQuads, bytes * 4 * (MT_RF_BRICKS ? 2 : 1)));
            MT_CUDA(c, mt_launch_build_rf_quads(c->tex[slot], (int)w, (int)h, (int)d, c->rfQuads, c->stream));
            c->launches += 1;
        }
    }
    if (slot == MT_TEX_LOW_FREQ) {
        if (c->hwTex) { cudaDestroyTextureObject(c->hwTex); c->hwTex = 0; }
        if (c->hwArray) { cudaFreeArray(c->hwArray); c->hwArray = nullptr; }
        if (c->flags & MT_FLAG_HW_CONE_FILTER) {
            cudaChannelFormatDesc fd = cudaCreateChannelDesc<uchar4>();
            MT_CUDA(c, cudaMalloc3DArray(&c->hwArray, &fd, make_cudaExtent(w, h, d)));
            cudaMemcpy3DParms cp = {};
            cp.srcPtr = make_cudaPitchedPtr((void*)rgba8, (size_t)w * 4, w, h);
            cp.dstArray = c->hwArray;
            cp.extent = make_cudaExtent(w, h, d);
            cp.kind = cudaMemcpyHostToDevice;
            MT_CUDA(c, cudaMemcpy3D(&cp));
            cudaResourceDesc rd = {};
            rd.resType = cudaResourceTypeArray;
            rd.res.array.array = c->hwArray;
            cudaTextureDesc td = {};
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;  // REPEAT (Texture3D.cpp:92-134)
            td.filterMode = cudaFilterModeLinear;
            td.readMode = cudaReadModeNormalizedFloat;
            td.normalizedCoords = 1;
            MT_CUDA(c, cudaCreateTextureObject(&c->hwTex, &rd, &td, nullptr));
        }
    }
    MT_CUDA(c, cudaStreamSynchronize(c->stream));  // the caller may free its buffer on return
    c->texw[slot] = (int)w; c->texh[slot] = (int)h; c->texd[slot] = (int)d;
    if (slot == MT_TEX_LOW_FREQ) {
        cudaFree(c->occ);
        c->occ = nullptr;
        c->occCoverage = -1.0f;
        if (w >= 32) MT_CUDA(c, cudaMalloc((void**)&c->occ, (size_t)(w / 32) * h * d * sizeof(uint32_t)));
    }
    return MT_OK;
}
MtStatus mtUploadTexture3D(MtContext* c, MtTextureSlot slot, uint32_t w, uint32_t h, uint32_t d, const uint8_t* rgba8)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, slot == MT_TEX_LOW_FREQ || slot == MT_TEX_HIGH_FREQ, "mtUploadTexture3D: slot is not a 3D texture");
    return upload(c, (int)slot, w, h, d, rgba8);
} MT_NOTHROW
MtStatus mtUploadTexture2D(MtContext* c, MtTextureSlot slot, uint32_t w, uint32_t h, const uint8_t* rgba8)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, slot == MT_TEX_CURL || slot == MT_TEX_WEATHER, "mtUploadTexture2D: slot is not a 2D texture");
    return upload(c, (int)slot, w, h, 1, rgba8);
} MT_NOTHROW

// ---- dispatch helpers ---------------------------------------------------------------------------------------------
static void copy_cam(CamU& d, const MtCameraUBO& s)
{
    static_assert(sizeof(CamU) == sizeof(MtCameraUBO), "CamU layout");
    memcpy(&d, &s, sizeof(CamU));
}
static void copy_time(TimeU& d, const MtTimeUBO& s)
{
    static_assert(sizeof(TimeU) == sizeof(MtTimeUBO), "TimeU layout");
    memcpy(&d, &s, sizeof(TimeU));
}
static void pass_begin(MtContext* c, MtPass p)
{
    if (c->flags & MT_FLAG_PASS_TIMING_INTERNAL) cudaEventRecord(c->ev[p][0], c->stream);
}
static void pass_end(MtContext* c, MtPass p)
{
    if (c->flags & MT_FLAG_PASS_TIMING_INTERNAL) {
        cudaEventRecord(c->ev[p][1], c->stream);
        c->evValid[p] = true;
    }
}

static MtStatus cloud_dispatch(MtContext* c, int full, const RowTiles* tiles, bool debug)
{
    if (!c->haveCam || !c->haveTime) return fail(c, MT_ERR_NOT_READY, "cloud dispatch: camera / time uniforms not set");
    if (!c->tex[MT_TEX_LOW_FREQ] || !c->tex[MT_TEX_HIGH_FREQ] || !c->tex[MT_TEX_CURL])
        return fail(c, MT_ERR_NOT_READY, "cloud dispatch: low-frequency, high-frequency and curl textures must be uploaded");
    MT_CUDA(c, cudaSetDevice(c->device));
    CloudParams P;
    memset(&P, 0, sizeof(P));
    copy_cam(P.cam, c->cam);
    copy_time(P.tm, c->tm);
    P.tun = c->tun;
    mt_host_sky_const(c->cam, c->tun, P.sky);
    P.low.quads = (const Quad*)c->quads[MT_TEX_LOW_FREQ];
    P.low.rfquads = (const Quad*)c->rfQuads;
    P.low.hwtex = (unsigned long long)c->hwTex;  // 0 unless MT_FLAG_HW_CONE_FILTER
    P.hwCone = (c->flags & MT_FLAG_HW_CONE_FILTER) && c->hwTex ? 1 : 0;
    P.high.quads = (const Quad*)c->quads[MT_TEX_HIGH_FREQ];
    P.curl.quads = (const Quad*)c->quads[MT_TEX_CURL];
    P.low.texels = c->tex[MT_TEX_LOW_FREQ];
    P.low.w = c->texw[MT_TEX_LOW_FREQ]; P.low.h = c->texh[MT_TEX_LOW_FREQ]; P.low.d = c->texd[MT_TEX_LOW_FREQ];
    P.high.texels = c->tex[MT_TEX_HIGH_FREQ];
    P.high.w = c->texw[MT_TEX_HIGH_FREQ]; P.high.h = c->texh[MT_TEX_HIGH_FREQ]; P.high.d = c->texd[MT_TEX_HIGH_FREQ];
    P.curl.texels = c->tex[MT_TEX_CURL];
    P.curl.w = c->texw[MT_TEX_CURL]; P.curl.h = c->texh[MT_TEX_CURL];
    if (c->tun.use_weather) {
        if (!c->tex[MT_TEX_WEATHER]) return fail(c, MT_ERR_NOT_READY, "cloud dispatch: use_weather needs the weather map uploaded");
        if (debug || (c->flags & MT_FLAG_COUNTERS))
            return fail(c, MT_ERR_INVALID, "cloud dispatch: counters / debug records are not available with use_weather");
        P.weather.texels = c->tex[MT_TEX_WEATHER];
        P.weather.quads = (const Quad*)c->quads[MT_TEX_WEATHER];
        P.weather.w = c->texw[MT_TEX_WEATHER]; P.weather.h = c->texh[MT_TEX_WEATHER];
    }
    P.low.occ = nullptr;
    P.high.occ = nullptr;
    if (c->occ) {  // (re)build the empty-cell bitmap when the volume or the coverage changed
        if (c->occCoverage != c->tun.coverage) {
            MT_CUDA(c, mt_launch_occupancy(P.low, c->occ, c->tun.coverage, c->stream));
            c->occCoverage = c->tun.coverage;
            c->launches += 1;
        }
        P.low.occ = c->occ;
    }
    // the per-frame constants of the march, evaluated here (cloud_core.cuh: __host__ __device__, IEEE operations without
    // contraction on both sides) and passed in the parameter block
    cloud_frame_setup(P.cam, P.tm, P.tun, P.mc);
    cloud_frame_jitter(P.tm, c->W, c->H, P.mc.tabs);
    P.hdr = c->outHdr ? c->outHdr : c->hdr[c->cur];
    P.mask = c->outMask ? c->outMask : c->mask;
    P.W = c->W; P.H = c->H;
    P.tx = (((c->W / 4) + 31) / 32) * 32;  // Renderer.cpp:713-714
    P.ty = (((c->H / 4) + 31) / 32) * 32;
    P.full = full;
    P.storage = (int)c->storage;
    P.bulkStore = c->storeMode == (int)MT_STORE_BULK;
    if (tiles) P.rows = *tiles;
    else if (full) {  // whole frame = every 8-row tile, so that the launch-order heuristic below applies to it as well
        P.rows.tile_rows = 8;
        P.rows.tile_begin = 0; P.rows.tile_stride = 1; P.rows.tile_count = (c->H + 7) / 8;
    }
    P.rows.heavy_first = 0;
    if (!full) {  // fused 1-of-16 kernel: its CTAs are 8x4-ray tiles; a row of tiles covers 16 pixel rows (cloud_sixteenth_kernel)
        P.rows.tile_rows = 4 * MT_S16_TH;
        P.rows.tile_begin = 0; P.rows.tile_stride = 1; P.rows.tile_count = P.ty / MT_S16_TH;
    }
    if (!(c->flags & MT_FLAG_TOP_DOWN)) {
        // first pixel row whose centre-column ray no longer marches (dir.y < 0.06, cloudRayMarch.comp:730): castRay of
        // camera.cpp's basis with nx = 0.  A heuristic for the launch order only -- any value renders the same image.
        const float* v = c->cam.view;
        const float upy = v[5], looky = -v[6], tany = c->cam.tanFovBy2[1];
        const float upl = sqrtf(v[1] * v[1] + v[5] * v[5] + v[9] * v[9]), lookl = sqrtf(v[2] * v[2] + v[6] * v[6] + v[10] * v[10]);
        int yh = c->H;
        for (int y = 0; y < c->H; ++y) {
            const float ny = (1.0f - (float)y / (float)c->H) * 2.0f - 1.0f;
            const float dx = 0.0f, dy = looky / (lookl > 0 ? lookl : 1.0f) + (upy / (upl > 0 ? upl : 1.0f)) * ny * tany;
            const float dz2 = 1.0f + (ny * tany) * (ny * tany);  // |look + up * ny * tany|^2 for an orthonormal basis
            (void)dx;
            if (dy / sqrtf(dz2) < 0.06f) { yh = y; break; }
        }
        int m = 0;
        for (int k = 0; k < P.rows.tile_count; ++k)
            if ((P.rows.tile_begin + k * P.rows.tile_stride) * P.rows.tile_rows < yh) m = k + 1;
        P.rows.heavy_first = m;
    }
    P.counters = (debug || (c->flags & MT_FLAG_COUNTERS)) ? c->counters : nullptr;
    P.debug = nullptr;
    if (debug) {
        if (!c->debug) MT_CUDA(c, cudaMalloc((void**)&c->debug, (size_t)c->W * c->H * sizeof(MtRayDebug)));
        MT_CUDA(c, cudaMemsetAsync(c->debug, 0, (size_t)c->W * c->H * sizeof(MtRayDebug), c->stream));
        P.debug = c->debug;
    }
    // the 1-of-16 dispatch runs step-parallel (cloud_raymarch.cu) unless counters / debug records are wanted
    const bool stepParallel = !full && !debug && !P.counters && !(c->flags & MT_FLAG_SEQUENTIAL_MARCH);
    const bool split = stepParallel && (c->flags & MT_FLAG_SPLIT_MARCH);  // the three-kernel form (global scratch); default: one fused kernel
    if (split) {  // lazily allocated, each pointer tested by itself: a failed allocation leaves no orphan behind
        const size_t nrays = (size_t)P.tx * (size_t)P.ty;
        if (!c->rays) MT_CUDA(c, cudaMalloc(&c->rays, nrays * 64));
        if (!c->samples) MT_CUDA(c, cudaMalloc((void**)&c->samples, nrays * MT_STEP_SLICES * sizeof(float2)));
        if (!c->ctaSteps) MT_CUDA(c, cudaMalloc((void**)&c->ctaSteps, (nrays / 128 + 1) * sizeof(int)));
#if MT_STEP_COMPACT
        if (!c->items) MT_CUDA(c, cudaMalloc((void**)&c->items, nrays * MT_STEP_SLICES * sizeof(unsigned)));   // only the compacting variant lists (step, ray) pairs
#endif
        if (!c->itemCount) MT_CUDA(c, cudaMalloc((void**)&c->itemCount, sizeof(unsigned)));
    }
    P.rays = c->rays;
    P.samples = c->samples;
    P.ctaSteps = c->ctaSteps;
    P.items = c->items;
    P.itemCount = c->itemCount;
    wait_pending_read(c, P.hdr);
    wait_pending_read(c, P.mask);
    pass_begin(c, MT_PASS_CLOUD);
    int n = 1;
    // gather by forwarding: full-quality row-tile launches keep their stores local and a side kernel pushes finished tiles
    const bool forward = full && tiles && c->forwardHdr && !debug && !P.counters && !c->outHdr;
    P.tileDone = nullptr;
    if (forward) {
        if (!c->tileDone) MT_CUDA(c, cudaMalloc((void**)&c->tileDone, ((size_t)c->H / 8 + 2) * sizeof(unsigned)));  // freed by mtResize
        if (c->fwdBusy) MT_CUDA(c, cudaStreamWaitEvent(c->stream, c->fwdDoneEv, 0));  // the previous frame's tiles have left
        MT_CUDA(c, cudaMemsetAsync(c->tileDone, 0, ((size_t)P.rows.tile_count + 1) * sizeof(unsigned), c->stream));
        MT_CUDA(c, cudaEventRecord(c->fwdArmEv, c->stream));
        P.tileDone = c->tileDone;
        n = 2;
    }
    // the decoded copy of the mask stays current only through the fused 1-of-16 kernel writing this context's own mask
    const bool keepsDecoded = stepParallel && !split && !c->outMask && !c->maskShared && c->decodedCurrent && MT_INCREMENTAL_DECODE;
    P.decoded = keepsDecoded ? c->maskDecoded : nullptr;
    P.decodedPitch = (int)mt_godray_pitch(c->W);
    c->decodedCurrent = keepsDecoded;
    if (split) MT_CUDA(c, mt_launch_cloud_sixteenth_split(P, c->stream, &n));
    else if (stepParallel) MT_CUDA(c, mt_launch_cloud_sixteenth_fused(P, c->stream));
    else MT_CUDA(c, mt_launch_cloud(P, c->stream));
    if (forward) {
        // Submitted AFTER the march kernel it waits on: streams can share a hardware work queue (CUDA_DEVICE_MAX_CONNECTIONS),
        // and a spinning kernel queued ahead of the kernel that feeds it would then never be fed.
        MT_CUDA(c, cudaStreamWaitEvent(c->fwdStream, c->fwdArmEv, 0));
        MT_CUDA(c, mt_launch_tile_forward(P.hdr, c->forwardHdr, c->W, c->H, (int)hdr_pixel_bytes(c->storage), P.rows, c->tileDone, 8, c->fwdStream));
        MT_CUDA(c, cudaEventRecord(c->fwdDoneEv, c->fwdStream));
        c->fwdBusy = true;
        c->fwdCheck = true;
        c->fwdTiles = P.rows.tile_count;
    }
    pass_end(c, MT_PASS_CLOUD);
    c->launches += (uint64_t)n;
    return MT_OK;
}

MtStatus mtDispatchCloud(MtContext* c)
try {
    if (!c) return MT_ERR_INVALID;
    return cloud_dispatch(c, 0, nullptr, false);
} MT_NOTHROW
MtStatus mtDispatchCloudFull(MtContext* c)
try {
    if (!c) return MT_ERR_INVALID;
    return cloud_dispatch(c, 1, nullptr, false);
} MT_NOTHROW
MtStatus mtDispatchCloudTiles(MtContext* c, uint32_t tile_rows, uint32_t tile_begin, uint32_t tile_end, uint32_t tile_stride)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, tile_rows >= 8 && tile_rows % 8 == 0, "mtDispatchCloudTiles: tile_rows must be a positive multiple of 8");
    MT_REQUIRE(c, tile_stride >= 1, "mtDispatchCloudTiles: tile_stride must be >= 1");
    uint32_t ntiles = ((uint32_t)c->H + tile_rows - 1) / tile_rows;
    if (tile_end > ntiles) tile_end = ntiles;
    RowTiles t;
    t.tile_rows = (int)tile_rows;
    t.tile_begin = (int)tile_begin;
    t.tile_stride = (int)tile_stride;
    t.tile_count = tile_begin < tile_end ? (int)((tile_end - tile_begin + tile_stride - 1) / tile_stride) : 0;
    if (t.tile_count == 0) return MT_OK;
    return cloud_dispatch(c, 1, &t, false);
} MT_NOTHROW
MtStatus mtDispatchCloudDebug(MtContext* c, int full, MtRayDebug* out, size_t out_bytes)
try {
    if (!c) return MT_ERR_INVALID;
    size_t need = (size_t)c->W * c->H * sizeof(MtRayDebug);
    MT_REQUIRE(c, out != nullptr && out_bytes >= need, "mtDispatchCloudDebug: output buffer too small");
    MtStatus st = cloud_dispatch(c, full ? 1 : 0, nullptr, true);
    if (st != MT_OK) return st;
    MT_CUDA(c, cudaMemcpyAsync(out, c->debug, need, cudaMemcpyDeviceToHost, c->stream));
    MT_CUDA(c, cudaStreamSynchronize(c->stream));
    return MT_OK;
} MT_NOTHROW

static MtStatus reproject_dispatch(MtContext* c, bool debug)
{
    if (!c->haveCam || !c->haveCamOld || !c->haveTime)
        return fail(c, MT_ERR_NOT_READY, "reprojection dispatch: camera, cameraOld and time uniforms must be set");
    MT_CUDA(c, cudaSetDevice(c->device));
    ReprojParams P;
    memset(&P, 0, sizeof(P));
    copy_cam(P.cam, c->cam);
    copy_cam(P.camOld, c->camOld);
    copy_time(P.tm, c->tm);
    P.prev = c->hdr[c->cur ^ 1];
    P.cur = c->hdr[c->cur];
    P.W = c->W; P.H = c->H;
    P.storage = (int)c->storage;
    P.frame = reproject_frame(P);
    P.uv = c->uvTab;
    P.skipId = debug ? -1 : c->reprojSkipId;
    P.tx = (((c->W / 4) + 31) / 32) * 32;  // the Cloud dispatch's thread grid (Renderer.cpp:713-714)
    P.ty = (((c->H / 4) + 31) / 32) * 32;
    P.taps = nullptr;
    if (debug) {
        if (!c->taps) MT_CUDA(c, cudaMalloc((void**)&c->taps, (size_t)c->W * c->H * 10 * sizeof(int)));
        P.taps = c->taps;
    }
    wait_pending_read(c, P.cur);
    pass_begin(c, MT_PASS_REPROJECT);
    MT_CUDA(c, mt_launch_reproject(P, c->stream));
    pass_end(c, MT_PASS_REPROJECT);
    c->launches += 1;
    return MT_OK;
}
MtStatus mtDispatchReprojection(MtContext* c)
try {
    if (!c) return MT_ERR_INVALID;
    return reproject_dispatch(c, false);
} MT_NOTHROW
MtStatus mtDispatchReprojectionDebug(MtContext* c, int32_t* taps, size_t taps_bytes)
try {
    if (!c) return MT_ERR_INVALID;
    size_t need = (size_t)c->W * c->H * 10 * sizeof(int32_t);
    MT_REQUIRE(c, taps != nullptr && taps_bytes >= need, "mtDispatchReprojectionDebug: output buffer too small");
    MtStatus st = reproject_dispatch(c, true);
    if (st != MT_OK) return st;
    MT_CUDA(c, cudaMemcpyAsync(taps, c->taps, need, cudaMemcpyDeviceToHost, c->stream));
    MT_CUDA(c, cudaStreamSynchronize(c->stream));
    return MT_OK;
} MT_NOTHROW

static unsigned tonemap_seed(const MtContext* c)
{
    const float ty = c->tm.time[1];  // uint(time.y): truncate, saturate, NaN -> 0
    return (ty != ty || ty <= 0.0f) ? 0u : (ty >= 4294967296.0f ? 0xffffffffu : (unsigned)ty);
}
// fuse_tonemap: the god-ray kernel also tone-maps the pixel it has just finished (mtFrameEx with both passes): one read of
// the HDR image less, one launch less.  The LDR bytes are those of mtDispatchToneMap run after mtDispatchGodRays.
static MtStatus godrays_dispatch(MtContext* c, bool fuse_tonemap)
{
    if (!c->haveCam) return fail(c, MT_ERR_NOT_READY, "god-ray dispatch: camera uniform not set");
    if (fuse_tonemap && !c->haveTime) return fail(c, MT_ERR_NOT_READY, "tone-map dispatch: time uniform not set");
    MT_CUDA(c, cudaSetDevice(c->device));
    GodRayParams P;
    memset(&P, 0, sizeof(P));
    copy_cam(P.cam, c->cam);
    P.lightColor[0] = c->sky.lightColor[0]; P.lightColor[1] = c->sky.lightColor[1]; P.lightColor[2] = c->sky.lightColor[2];
    P.mask = c->mask;
    P.decoded = c->maskDecoded;
    P.hdr = c->hdr[c->cur];
    P.W = c->W; P.H = c->H;
    P.storage = (int)c->storage;
    P.ldr = fuse_tonemap ? c->ldr[c->cur] : nullptr;
    P.seed = tonemap_seed(c);
    P.frame = godray_frame(P.cam);
    P.uv = c->uvTab;
    P.decodedCurrent = c->decodedCurrent && !c->maskShared;
    wait_pending_read(c, P.hdr);
    if (fuse_tonemap) wait_pending_read(c, P.ldr);
    pass_begin(c, MT_PASS_GODRAYS);
    MT_CUDA(c, mt_launch_godrays(P, c->stream));
    pass_end(c, MT_PASS_GODRAYS);
    if (fuse_tonemap) {  // mtLastPassMs(MT_PASS_TONEMAP) reads 0: its work is inside the god-ray pass
        pass_begin(c, MT_PASS_TONEMAP);
        pass_end(c, MT_PASS_TONEMAP);
    }
    c->launches += P.decodedCurrent ? 1 : 2;  // (mask_decode_kernel +) godrays_kernel
    c->decodedCurrent = true;
    return MT_OK;
}
MtStatus mtDispatchGodRays(MtContext* c)
try {
    if (!c) return MT_ERR_INVALID;
    return godrays_dispatch(c, false);
} MT_NOTHROW

MtStatus mtDispatchToneMap(MtContext* c)
try {
    if (!c) return MT_ERR_INVALID;
    if (!c->haveTime) return fail(c, MT_ERR_NOT_READY, "tone-map dispatch: time uniform not set");
    MT_CUDA(c, cudaSetDevice(c->device));
    ToneMapParams P;
    P.storage = (int)c->storage;
    P.hdr = c->hdr[c->cur];
    P.ldr = c->ldr[c->cur];
    P.W = c->W; P.H = c->H;
    P.seed = tonemap_seed(c);
    wait_pending_read(c, P.ldr);
    pass_begin(c, MT_PASS_TONEMAP);
    MT_CUDA(c, mt_launch_tonemap(P, c->stream));
    pass_end(c, MT_PASS_TONEMAP);
    c->launches += 1;
    return MT_OK;
} MT_NOTHROW

MtStatus mtDispatchTXAA(MtContext* c)
try {
    if (!c) return MT_ERR_INVALID;
    if (!c->haveCam || !c->haveCamOld || !c->haveTime)
        return fail(c, MT_ERR_NOT_READY, "TXAA dispatch: camera, cameraOld and time uniforms must be set");
    MT_CUDA(c, cudaSetDevice(c->device));
    TxaaParams P;
    memset(&P, 0, sizeof(P));
    copy_cam(P.cam, c->cam);
    copy_cam(P.camOld, c->camOld);
    copy_time(P.tm, c->tm);
    P.cur = c->ldr[c->cur];
    P.prev = c->ldr[c->cur ^ 1];
    P.out = c->ldrScratch;
    P.W = c->W; P.H = c->H;
    P.frame = txaa_frame(P);
    P.uv = c->uvTab;
    wait_pending_read(c, P.out);
    pass_begin(c, MT_PASS_TXAA);
    MT_CUDA(c, mt_launch_txaa(P, c->stream));
    pass_end(c, MT_PASS_TXAA);
    c->launches += 1;
    // the shader writes its result back into currentFrameResultImage: the output becomes this frame's LDR image
    uint32_t* t = c->ldr[c->cur];
    c->ldr[c->cur] = c->ldrScratch;
    c->ldrScratch = t;
    return MT_OK;
} MT_NOTHROW

MtStatus mtSwapPingPong(MtContext* c)
try {
    if (!c) return MT_ERR_INVALID;
    c->cur ^= 1;
    return MT_OK;
} MT_NOTHROW

MtStatus mtFrameEx(MtContext* c, uint32_t passes)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, !(passes & MT_FRAME_TXAA) || (passes & MT_FRAME_TONEMAP), "mtFrameEx: TXAA needs the tone-map pass");
    MtStatus st;
    // The frame's two compute passes touch disjoint pixels -- the Cloud dispatch writes the sixteenth of the image with this frame's
    // id (and reads no image), the reprojection fills the rest -- so they run SIDE BY SIDE on two streams: the reprojection is issue
    // bound, the step-parallel Cloud kernel latency bound with a third of its issue slots idle.  The reprojection leaves the Cloud
    // dispatch's pixels out (the shader writes them and the Cloud pass overwrites them: the same image).  Serial where a pass must be
    // timed or counted by itself, or where the dispatch is not the fused 1-of-16 kernel writing this context's own images.
    const bool overlap = MT_FRAME_OVERLAP && !(c->flags & (MT_FLAG_PASS_TIMING_INTERNAL | MT_FLAG_COUNTERS | MT_FLAG_SEQUENTIAL_MARCH | MT_FLAG_SPLIT_MARCH)) &&
                         !c->outHdr && !c->outMask && c->haveTime;
    if (overlap) {
        MT_CUDA(c, cudaSetDevice(c->device));
        MT_CUDA(c, cudaEventRecord(c->sideForkEv, c->stream));            // everything issued so far (the previous frame's passes) ...
        MT_CUDA(c, cudaStreamWaitEvent(c->sideStream, c->sideForkEv, 0)); // ... precedes the Cloud dispatch
        // the reprojection's launch first: its short CTAs take the machine, the Cloud kernel's long ones fill in as they retire
        // (the other order measures like the serial frame: 379.2 vs 379.9 us; this one 369.2 us at 1080p with TXAA)
        c->reprojSkipId = c->tm.frameCountMod16 & 15;
        st = reproject_dispatch(c, false);
        c->reprojSkipId = -1;
        if (st != MT_OK) return st;
        cudaStream_t mainStream = c->stream;
        c->stream = c->sideStream;
        st = cloud_dispatch(c, 0, nullptr, false);
        c->stream = mainStream;
        // joined whether or not the dispatch went through: whatever it did enqueue on the side stream precedes the caller's next call
        const cudaError_t je = cudaEventRecord(c->sideJoinEv, c->sideStream);
        const cudaError_t jw = je == cudaSuccess ? cudaStreamWaitEvent(c->stream, c->sideJoinEv, 0) : je;  // the later passes need both
        if (st != MT_OK) return st;
        MT_CUDA(c, jw);
    } else {
        if ((st = reproject_dispatch(c, false)) != MT_OK) return st;
        if ((st = cloud_dispatch(c, 0, nullptr, false)) != MT_OK) return st;
    }
    const bool fused = (passes & MT_FRAME_GODRAYS) && (passes & MT_FRAME_TONEMAP) && !(c->flags & MT_FLAG_NO_FUSED_TONEMAP);
    if ((passes & MT_FRAME_GODRAYS) && (st = godrays_dispatch(c, fused)) != MT_OK) return st;
    if ((passes & MT_FRAME_TONEMAP) && !fused && (st = mtDispatchToneMap(c)) != MT_OK) return st;
    if ((passes & MT_FRAME_TXAA) && (st = mtDispatchTXAA(c)) != MT_OK) return st;
    return mtSwapPingPong(c);
} MT_NOTHROW
MtStatus mtFrame(MtContext* c, int with_godrays)
try {
    return mtFrameEx(c, MT_FRAME_TONEMAP | (with_godrays ? MT_FRAME_GODRAYS : 0u));
} MT_NOTHROW

MtStatus mtSynchronize(MtContext* c)
try {
    if (!c) return MT_ERR_INVALID;
    MT_CUDA(c, cudaSetDevice(c->device));
    MT_CUDA(c, cudaStreamSynchronize(c->stream));
    MT_CUDA(c, cudaStreamSynchronize(c->copyStream));
    MT_CUDA(c, cudaStreamSynchronize(c->sideStream));
    if (c->fwdStream) MT_CUDA(c, cudaStreamSynchronize(c->fwdStream));
    c->fwdBusy = false;
    for (auto& p : c->pending) p.active = false;
    if (c->fwdCheck) {
        c->fwdCheck = false;
        unsigned overrun = 0;
        MT_CUDA(c, cudaMemcpy(&overrun, c->tileDone + c->fwdTiles, sizeof(unsigned), cudaMemcpyDeviceToHost));
        if (overrun) return fail(c, MT_ERR_CUDA, "mtSynchronize: the tile forwarder gave up waiting for the march kernel (tiles not delivered)");
    }
    return MT_OK;
} MT_NOTHROW

// ---- images -------------------------------------------------------------------------------------------------------
MtStatus mtImageBytes(const MtContext* c, MtImage which, size_t* bytes)
try {
    if (!c || !bytes || (int)which < 0 || (int)which > MT_IMAGE_LDR_PREV) return MT_ERR_INVALID;
    *bytes = image_bytes(c, which);
    return MT_OK;
} MT_NOTHROW
MtStatus mtReadImageRows(MtContext* c, MtImage which, uint32_t row_begin, uint32_t row_end, void* host, size_t bytes)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, (int)which >= 0 && (int)which <= MT_IMAGE_LDR_PREV && host != nullptr, "mtReadImageRows: bad arguments");
    MT_REQUIRE(c, row_begin <= row_end && row_end <= (uint32_t)c->H, "mtReadImageRows: bad row range");
    size_t pitch = (size_t)c->W * pixel_bytes(c, which);
    size_t need = pitch * (row_end - row_begin);
    MT_REQUIRE(c, bytes >= need, "mtReadImageRows: host buffer too small");
    MT_CUDA(c, cudaSetDevice(c->device));
    const char* src = (const char*)image_ptr(c, which) + pitch * row_begin;
    if (need) MT_CUDA(c, cudaMemcpyAsync(host, src, need, cudaMemcpyDeviceToHost, c->stream));
    MT_CUDA(c, cudaStreamSynchronize(c->stream));
    return MT_OK;
} MT_NOTHROW
MtStatus mtReadImage(MtContext* c, MtImage which, void* host, size_t bytes)
try {
    if (!c) return MT_ERR_INVALID;
    return mtReadImageRows(c, which, 0, (uint32_t)c->H, host, bytes);
} MT_NOTHROW
// shared tail of the asynchronous reads: `dev` (already valid on the main stream) -> host on the copy stream
static MtStatus async_read_tail(MtContext* c, const void* dev, void* host, size_t bytes)
{
    MtContext::PendingRead* slot = nullptr;
    for (auto& p : c->pending)
        if (p.active && p.dev == dev) slot = &p;     // a second read of the same image re-uses its slot
    for (auto& p : c->pending)
        if (!slot && !p.active) slot = &p;
    if (!slot) {                                     // all slots busy: retire the oldest by waiting for the copy stream
        MT_CUDA(c, cudaStreamSynchronize(c->copyStream));
        for (auto& p : c->pending) p.active = false;
        slot = &c->pending[0];
    }
    MT_CUDA(c, cudaEventRecord(c->producedEv, c->stream));
    MT_CUDA(c, cudaStreamWaitEvent(c->copyStream, c->producedEv, 0));
    MT_CUDA(c, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->copyStream));
    MT_CUDA(c, cudaEventRecord(slot->done, c->copyStream));
    slot->dev = dev;
    slot->active = true;
    return MT_OK;
}
MtStatus mtReadImageAsync(MtContext* c, MtImage which, void* host, size_t bytes)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, (int)which >= 0 && (int)which <= MT_IMAGE_LDR_PREV && host != nullptr, "mtReadImageAsync: bad arguments");
    MT_REQUIRE(c, bytes >= image_bytes(c, which), "mtReadImageAsync: host buffer too small");
    MT_CUDA(c, cudaSetDevice(c->device));
    const void* dev = image_ptr(c, which);
    if (which == MT_IMAGE_GODRAY_MASK) {
        // The god-ray mask is not ping-ponged (a 1-of-16 dispatch updates it in place), so a read-back straight from it would
        // make the next Cloud dispatch wait for PCIe.  Snapshot it device-to-device on the main stream (HBM speed) and read
        // the snapshot back on the copy stream instead.
        if (!c->maskStage) MT_CUDA(c, cudaMalloc((void**)&c->maskStage, image_bytes(c, which)));
        wait_pending_read(c, c->maskStage);  // the previous snapshot has left the device
        MT_CUDA(c, cudaMemcpyAsync(c->maskStage, dev, image_bytes(c, which), cudaMemcpyDeviceToDevice, c->stream));
        dev = c->maskStage;
    }
    return async_read_tail(c, dev, host, image_bytes(c, which));
} MT_NOTHROW
MtStatus mtReadGodRayGreyAsync(MtContext* c, float* host, size_t bytes)
try {
    if (!c) return MT_ERR_INVALID;
    const size_t need = (size_t)c->W * (size_t)c->H * sizeof(float);
    MT_REQUIRE(c, host != nullptr && bytes >= need, "mtReadGodRayGreyAsync: host buffer too small");
    MT_CUDA(c, cudaSetDevice(c->device));
    if (!c->greyStage) MT_CUDA(c, cudaMalloc((void**)&c->greyStage, need));
    wait_pending_read(c, c->greyStage);  // the previous decoded image has left the device
    GodRayParams P;
    memset(&P, 0, sizeof(P));
    P.mask = c->mask;
    P.W = c->W; P.H = c->H;
    P.storage = (int)c->storage;
    MT_CUDA(c, mt_launch_mask_grey(P, c->greyStage, c->stream));
    c->launches += 1;
    return async_read_tail(c, c->greyStage, host, need);
} MT_NOTHROW
MtStatus mtWaitReads(MtContext* c)
try {
    if (!c) return MT_ERR_INVALID;
    MT_CUDA(c, cudaSetDevice(c->device));
    MT_CUDA(c, cudaStreamSynchronize(c->copyStream));
    for (auto& p : c->pending) p.active = false;
    return MT_OK;
} MT_NOTHROW
MtStatus mtJoinCopies(MtContext* c)
try {
    if (!c) return MT_ERR_INVALID;
    MT_CUDA(c, cudaSetDevice(c->device));
    MT_CUDA(c, cudaEventRecord(c->producedEv, c->copyStream));
    MT_CUDA(c, cudaStreamWaitEvent(c->stream, c->producedEv, 0));
    if (c->fwdBusy) MT_CUDA(c, cudaStreamWaitEvent(c->stream, c->fwdDoneEv, 0));  // the tile forwarder of the last dispatch
    return MT_OK;
} MT_NOTHROW
MtStatus mtWriteImage(MtContext* c, MtImage which, const void* host, size_t bytes)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, (int)which >= 0 && (int)which <= MT_IMAGE_LDR_PREV && host != nullptr, "mtWriteImage: bad arguments");
    MT_REQUIRE(c, bytes == image_bytes(c, which), "mtWriteImage: size must equal the image size");
    MT_CUDA(c, cudaSetDevice(c->device));
    wait_pending_read(c, image_ptr(c, which));
    MT_CUDA(c, cudaMemcpyAsync(image_ptr(c, which), host, bytes, cudaMemcpyHostToDevice, c->stream));
    if (which == MT_IMAGE_GODRAY_MASK) c->decodedCurrent = false;
    return MT_OK;
} MT_NOTHROW
MtStatus mtClearImages(MtContext* c)
try {
    if (!c) return MT_ERR_INVALID;
    MT_CUDA(c, cudaSetDevice(c->device));
    for (auto& p : c->pending)
        if (p.active) wait_pending_read(c, p.dev);
    size_t px = (size_t)c->W * c->H;
    const size_t hb = hdr_pixel_bytes(c->storage);
    MT_CUDA(c, cudaMemsetAsync(c->hdr[0], 0, px * hb, c->stream));
    MT_CUDA(c, cudaMemsetAsync(c->hdr[1], 0, px * hb, c->stream));
    MT_CUDA(c, cudaMemsetAsync(c->mask, 0, px * hb, c->stream));
    c->decodedCurrent = false;
    MT_CUDA(c, cudaMemsetAsync(c->ldr[0], 0, px * 4, c->stream));
    MT_CUDA(c, cudaMemsetAsync(c->ldr[1], 0, px * 4, c->stream));
    MT_CUDA(c, cudaMemsetAsync(c->ldrScratch, 0, px * 4, c->stream));
    return MT_OK;
} MT_NOTHROW
MtStatus mtImageDevicePtr(MtContext* c, MtImage which, void** p)
try {
    if (!c || !p || (int)which < 0 || (int)which > MT_IMAGE_LDR_PREV) return MT_ERR_INVALID;
    *p = image_ptr(c, which);
    if (which == MT_IMAGE_GODRAY_MASK) c->maskShared = true;  // the caller may write it: the decoded copy is rebuilt by every god-ray dispatch from now on
    return MT_OK;
} MT_NOTHROW
MtStatus mtSetCloudOutput(MtContext* c, void* hdr, void* mask)
try {
    if (!c) return MT_ERR_INVALID;
    c->outHdr = (F4*)hdr;
    c->outMask = (F4*)mask;
    return MT_OK;
} MT_NOTHROW
MtStatus mtSetCloudStoreMode(MtContext* c, MtStoreMode mode)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, mode == MT_STORE_DIRECT || mode == MT_STORE_BULK, "mtSetCloudStoreMode: unknown mode");
    c->storeMode = (int)mode;
    return MT_OK;
} MT_NOTHROW
MtStatus mtSetCloudForward(MtContext* c, void* peer_hdr)
try {
    if (!c) return MT_ERR_INVALID;
    MT_CUDA(c, cudaSetDevice(c->device));
    MT_REQUIRE(c, !peer_hdr || c->storage != MT_STORAGE_F16 || (c->W % 2) == 0, "mtSetCloudForward: RGBA16F rows must be a multiple of 16 bytes (even width)");
    if (peer_hdr && !c->fwdStream) {
        int lo = 0, hi = 0;
        MT_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        MT_CUDA(c, cudaStreamCreateWithPriority(&c->fwdStream, cudaStreamNonBlocking, hi));  // its few CTAs go first when a slot frees
        MT_CUDA(c, cudaEventCreateWithFlags(&c->fwdArmEv, cudaEventDisableTiming));
        MT_CUDA(c, cudaEventCreateWithFlags(&c->fwdDoneEv, cudaEventDisableTiming));
    }
    if (c->fwdStream) MT_CUDA(c, cudaStreamSynchronize(c->fwdStream));
    c->fwdBusy = false;
    c->forwardHdr = peer_hdr;
    return MT_OK;
} MT_NOTHROW
MtStatus mtExportImageHandle(MtContext* c, MtImage which, uint8_t handle[64])
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, handle != nullptr && (int)which >= 0 && (int)which <= MT_IMAGE_LDR_PREV, "mtExportImageHandle: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    MT_CUDA(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    MT_CUDA(c, cudaIpcGetMemHandle(&h, image_ptr(c, which)));
    if (which == MT_IMAGE_GODRAY_MASK) c->maskShared = true;
    memcpy(handle, &h, 64);
    return MT_OK;
} MT_NOTHROW
MtStatus mtOpenPeerImage(MtContext* c, const uint8_t handle[64], void** p)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, handle != nullptr && p != nullptr, "mtOpenPeerImage: bad arguments");
    MT_CUDA(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    MT_CUDA(c, cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess));
    return MT_OK;
} MT_NOTHROW
MtStatus mtClosePeerImage(MtContext* c, void* p)
try {
    if (!c) return MT_ERR_INVALID;
    MT_CUDA(c, cudaSetDevice(c->device));
    MT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->outHdr == p) c->outHdr = nullptr;
    if (c->outMask == p) c->outMask = nullptr;
    MT_CUDA(c, cudaIpcCloseMemHandle(p));
    return MT_OK;
} MT_NOTHROW

MtStatus mtCopyTilesToPeer(MtContext* c, MtImage which, uint32_t tile_rows, uint32_t tile_begin, uint32_t tile_end,
                           uint32_t tile_stride, void* peer)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, peer != nullptr && (int)which >= 0 && (int)which <= MT_IMAGE_LDR_PREV, "mtCopyTilesToPeer: bad arguments");
    MT_REQUIRE(c, tile_rows >= 1 && tile_stride >= 1, "mtCopyTilesToPeer: bad tiling");
    MT_CUDA(c, cudaSetDevice(c->device));
    const size_t pitch = (size_t)c->W * pixel_bytes(c, which);
    const char* src = (const char*)image_ptr(c, which);
    const uint32_t ntiles = ((uint32_t)c->H + tile_rows - 1) / tile_rows;
    if (tile_end > ntiles) tile_end = ntiles;
    MT_CUDA(c, cudaEventRecord(c->producedEv, c->stream));
    MT_CUDA(c, cudaStreamWaitEvent(c->copyStream, c->producedEv, 0));
    for (uint32_t t = tile_begin; t < tile_end; t += tile_stride) {
        const size_t r0 = (size_t)t * tile_rows, r1 = (r0 + tile_rows < (size_t)c->H) ? r0 + tile_rows : (size_t)c->H;
        MT_CUDA(c, cudaMemcpyAsync((char*)peer + pitch * r0, src + pitch * r0, pitch * (r1 - r0), cudaMemcpyDeviceToDevice, c->copyStream));
    }
    return MT_OK;
} MT_NOTHROW

// ---- measurement --------------------------------------------------------------------------------------------------
MtStatus mtGetCounters(MtContext* c, MtCounters* out, int reset)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, out != nullptr, "mtGetCounters: null output");
    MT_CUDA(c, cudaSetDevice(c->device));
    unsigned long long v[8];
    MT_CUDA(c, cudaMemcpyAsync(v, c->counters, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    MT_CUDA(c, cudaStreamSynchronize(c->stream));
    out->rays = v[0]; out->rays_marched = v[1]; out->steps = v[2];
    out->steps_incloud = v[3]; out->cone_hits = v[4]; out->early_exits = v[5];
    if (reset) MT_CUDA(c, cudaMemsetAsync(c->counters, 0, sizeof(v), c->stream));
    return MT_OK;
} MT_NOTHROW
MtStatus mtLastPassMs(MtContext* c, MtPass pass, float* ms)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, ms != nullptr && (int)pass >= 0 && (int)pass < MT_PASS_COUNT, "mtLastPassMs: bad arguments");
    MT_REQUIRE(c, (c->flags & MT_FLAG_PASS_TIMING_INTERNAL) && c->evValid[pass], "mtLastPassMs: pass timing not enabled or pass never ran");
    MT_CUDA(c, cudaSetDevice(c->device));
    MT_CUDA(c, cudaEventSynchronize(c->ev[pass][1]));
    MT_CUDA(c, cudaEventElapsedTime(ms, c->ev[pass][0], c->ev[pass][1]));
    return MT_OK;
} MT_NOTHROW
MtStatus mtStreamHandle(MtContext* c, void** s)
try {
    if (!c || !s) return MT_ERR_INVALID;
    *s = (void*)c->stream;
    return MT_OK;
} MT_NOTHROW
MtStatus mtEventRecord(MtContext* c, uint32_t slot)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, slot < MT_USER_EVENTS, "mtEventRecord: slot out of range");
    MT_CUDA(c, cudaSetDevice(c->device));
    if (!c->userEv[slot]) MT_CUDA(c, cudaEventCreate(&c->userEv[slot]));
    MT_CUDA(c, cudaEventRecord(c->userEv[slot], c->stream));
    c->userEvValid[slot] = true;
    return MT_OK;
} MT_NOTHROW
MtStatus mtEventElapsedMs(MtContext* c, uint32_t from, uint32_t to, float* ms)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, ms != nullptr && from < MT_USER_EVENTS && to < MT_USER_EVENTS, "mtEventElapsedMs: bad arguments");
    MT_REQUIRE(c, c->userEvValid[from] && c->userEvValid[to], "mtEventElapsedMs: event slot never recorded");
    MT_CUDA(c, cudaSetDevice(c->device));
    MT_CUDA(c, cudaEventSynchronize(c->userEv[to]));
    MT_CUDA(c, cudaEventElapsedTime(ms, c->userEv[from], c->userEv[to]));
    return MT_OK;
} MT_NOTHROW
MtStatus mtFlushL2(MtContext* c, size_t bytes)
try {
    if (!c) return MT_ERR_INVALID;
    MT_CUDA(c, cudaSetDevice(c->device));
    const size_t unit = (size_t)256 << 20;
    size_t want = ((bytes ? bytes : unit) + unit - 1) / unit * unit;
    if (want > c->flushBytes) {
        MT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->flushBuf);
        c->flushBuf = nullptr;
        c->flushBytes = 0;
        MT_CUDA(c, cudaMalloc(&c->flushBuf, want));
        c->flushBytes = want;
    }
    MT_CUDA(c, cudaMemsetAsync(c->flushBuf, 0, want, c->stream));
    return MT_OK;
} MT_NOTHROW
MtStatus mtMeasureFp32Peak(MtContext* c, float* gflops)
try {
    if (!c) return MT_ERR_INVALID;
    MT_REQUIRE(c, gflops != nullptr, "mtMeasureFp32Peak: null output");
    MT_CUDA(c, cudaSetDevice(c->device));
    cudaDeviceProp prop;
    MT_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    float* sink = (float*)c->counters + 14;  // scratch word behind the six counters
    cudaEvent_t e0, e1;
    MT_CUDA(c, cudaEventCreate(&e0));
    MT_CUDA(c, cudaEventCreate(&e1));
    float best = 0.0f;
    for (int rep = 0; rep < 6; ++rep) {  // first reps double as warm-up
        cudaEventRecord(e0, c->stream);
        cudaError_t le = mt_launch_fma_probe(sink, blocks, iters, c->stream);
        cudaEventRecord(e1, c->stream);
        if (le != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) {
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            return cuda_fail(c, le != cudaSuccess ? le : cudaGetLastError(), "fma probe");
        }
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flop = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
        float g = (float)(flop / (ms * 1e-3) * 1e-9);
        if (rep >= 2 && g > best) best = g;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *gflops = best;
    return MT_OK;
} MT_NOTHROW
uint64_t mtLaunchCount(const MtContext* c) { return c ? c->launches : 0; }

}  // extern "C"
