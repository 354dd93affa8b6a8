/*
 * meteoros_b200.h -- C ABI of the B200-native cloud-rendering hot path.
 *
 * Drop-in boundary: the Vulkan compute/graphics dispatch that Meteoros records in
 *   Renderer::RecordComputeCommandBuffer   (src/CloudScapes/Renderer.cpp:653-722)
 *   Renderer::RecordGraphicsCommandBuffer  (src/CloudScapes/Renderer.cpp:723-856)
 * and submits from Renderer::Frame (Renderer.cpp:122-192).  Every entry point below
 * names the reference binding / call it replaces.  Plain pointers and sizes only; no
 * C++ or torch types cross this boundary.  All calls on one context are asynchronous
 * on that context's CUDA stream and execute in call order; mtSynchronize / mtReadImage
 * are the host sync points.  There is no CPU fallback: a context can only be created
 * on an sm_100 device (MT_ERR_UNSUPPORTED_ARCH otherwise).
 */
#ifndef METEOROS_B200_H
#define METEOROS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define MT_API
#else
#define MT_API __attribute__((visibility("default")))
#endif

#define MT_ABI_VERSION 1

typedef enum MtStatus {
    MT_OK = 0,
    MT_ERR_INVALID = 1,          /* bad argument / call order (reference: assert / validation layer) */
    MT_ERR_CUDA = 2,             /* a CUDA call failed (reference: throw std::runtime_error on VkResult != SUCCESS, Renderer.cpp:139-141) */
    MT_ERR_OOM = 3,
    MT_ERR_UNSUPPORTED_ARCH = 4, /* device is not sm_100; no fallback path exists */
    MT_ERR_NOT_READY = 5         /* a dispatch was issued before its textures were uploaded */
} MtStatus;

/* ---- uniform blocks: byte-identical to the reference's std140 UBOs -------------------------- */

/* CameraUBO: camera.h:12-18, cloudRayMarch.comp:15-21.  Column-major mat4 (glm).  152 bytes. */
typedef struct MtCameraUBO {
    float view[16];      /* @0   glm::lookAt RH                                        */
    float proj[16];      /* @64  glm::perspective RH, depth 0..1, proj[1][1] *= -1     */
    float eye[4];        /* @128 (eye, 1); shaders use -eye.xyz as the ray origin       */
    float tanFovBy2[2];  /* @144 (.x = aspect * .y, .y = |tan(fovy/2)|)                 */
} MtCameraUBO;

/* Time: Scene.h:12-21, cloudRayMarch.comp:23-31.  76 bytes. */
typedef struct MtTimeUBO {
    float haltonSeq1[4]; /* Halton base-3 index 1..4   (Scene.cpp:95-114) */
    float haltonSeq2[4]; /* index 5..8   */
    float haltonSeq3[4]; /* index 9..12  */
    float haltonSeq4[4]; /* index 13..16 */
    float time[2];       /* (delta t, total t) seconds */
    int32_t frameCountMod16;
} MtTimeUBO;

/* SunAndSky: Sky.h:9-15, cloudRayMarch.comp:33-39.  52 bytes.  Only lightColor is read (god rays). */
typedef struct MtSunAndSkyUBO {
    float sunLocation[4];
    float sunDirection[4];
    float lightColor[4];
    float sunIntensity;
} MtSunAndSkyUBO;

/*
 * Values the reference bakes into cloudRayMarch.comp as #defines / literals.  The defaults
 * (mtDefaultTuning) ARE those literals, so a context that never calls mtSetTuning renders the
 * reference's cloudscape.  Exposed so BASELINE config 5 can sweep sun elevation and coverage.
 */
typedef struct MtTuning {
    float coverage;            /* cloudRayMarch.comp:529  (0.6)                                   */
    float sun_location[3];     /* SUN_LOCATION, cloudRayMarch.comp:94  (0, .9*R_outer, -.9*R_outer) */
    float sky_sun_location[3]; /* BACKGROUND_SKY_SUN_LOCATION, cloudRayMarch.comp:95               */
    float wind_direction[3];   /* WIND_DIRECTION, cloudRayMarch.comp:89  (1,0,0)                   */
    float cloud_speed;         /* CLOUD_SPEED, cloudRayMarch.comp:90  (0.08)                       */
    float cloud_top_offset;    /* CLOUD_TOP_OFFSET, cloudRayMarch.comp:91  (1.0)                   */
    float base_density_factor; /* cloudRayMarch.comp:571  (0.38)                                   */
    /* Weather-map / cloud-type path.  The reference carries it as commented-out code
     * (cloudRayMarch.comp:515-525 + getDensityHeightGradientForPoint :475-487) and runs with it OFF; use_weather != 0
     * restores exactly that block: weather = texture(weatherMapSampler, unskewedSamplePoint.xz * weather_scale),
     * baseCloud *= heightGradient(relativeHeight, weather.g) * 0.5, coverage = weather.r (instead of `coverage`).   */
    uint32_t use_weather;      /* 0 (reference behaviour)                                          */
    float weather_scale;       /* 1.0 = the literal commented code (the map then repeats every unit of the sample point) */
} MtTuning;

typedef enum MtStorage {
    MT_STORAGE_F32 = 0,         /* HDR + mask images are RGBA32F (what the shaders declare, reprojection.comp:10-11) */
    MT_STORAGE_F16_EMULATE = 1, /* values are rounded through binary16 at every store, like the reference's
                                   R16G16B16A16_SFLOAT images (Renderer.cpp:1431-1440); memory stays RGBA32F */
    MT_STORAGE_F16 = 2          /* HDR + mask images ARE RGBA16F (8 bytes / pixel), the reference's actual format
                                   (Renderer.cpp:1431-1440): the same values as F16_EMULATE in half the HBM, NVLink and PCIe
                                   bytes.  mtReadImage / mtWriteImage / mtImageBytes then speak RGBA16F for the three float
                                   images; the width must be even for mtSetCloudForward. */
} MtStorage;

typedef struct MtConfig {
    uint32_t struct_size; /* = sizeof(MtConfig); ABI guard                       */
    uint32_t width;       /* window_width  (main.cpp:22-23)                      */
    uint32_t height;      /* window_height                                       */
    int32_t device;       /* CUDA device ordinal                                 */
    uint32_t storage;     /* MtStorage                                           */
    uint32_t flags;       /* MT_FLAG_*                                           */
} MtConfig;

#define MT_FLAG_COUNTERS 1u    /* cloud pass also accumulates MtCounters (slower; for work accounting) */
#define MT_FLAG_PASS_TIMING 2u /* bracket every pass with CUDA events so mtLastPassMs works               */
#define MT_FLAG_SEQUENTIAL_MARCH 4u /* 1-of-16 dispatch: one kernel, one thread per ray (default: step-parallel, one fused kernel) */
#define MT_FLAG_NO_FUSED_TONEMAP 64u /* mtFrame / mtFrameEx: run god rays and tone map as two passes (default: the god-ray kernel tone-maps
                                     * the pixel it has just finished; same bytes)                                               */
#define MT_FLAG_SPLIT_MARCH 32u     /* 1-of-16 dispatch: the three-kernel step-parallel form (rays / steps / fold through 512 B of
                                     * global scratch per ray; kept for A/B, profiles/r2_ab.md)                                   */
#define MT_FLAG_TOP_DOWN 8u         /* full-quality launches walk their row tiles top-down (default: from the horizon upwards, ocean last) */
#define MT_FLAG_NO_CONE_RF 16u      /* light-cone samples through the canonical four-channel filter instead of the (r, F) form of
                                     * the low-frequency volume (saves its 4 bytes/cell copy; bit-identical decisions either way,
                                     * radiance equal to rounding).  Must be set before the low-frequency texture is uploaded. */

#define MT_FLAG_HW_CONE_FILTER 128u /* OPT-IN, outside the parity bar: the six light-cone samples of an in-cloud step (production
                                     * kernels of mtDispatchCloud / mtDispatchCloudFull / mtDispatchCloudTiles / mtFrame; not the counting,
                                     * debug, sequential, split or weather forms) are filtered by the texture unit (a CUDA 3D
                                     * texture object over the low-frequency volume: LINEAR, REPEAT, as Texture3D.cpp:92-134 sets the
                                     * reference's sampler up) instead of the exact fp32 filter -- what a GPU running the reference does,
                                     * with the hardware's 8-bit filter weights.  ~29 % faster at 4K; the god-ray mask and alpha are
                                     * bit-identical to the default path (the cone samples never feed the accumulated density), the
                                     * HDR radiance differs by up to ~2e-3 relative on a few pixels per 4K frame (the default path:
                                     * ~1e-6).  Must be set before the low-frequency texture is uploaded.                            */

/* Texture slots = set 1 of the cloud pipeline (Renderer.cpp:1110-1114, cloudRayMarch.comp:9-12). */
typedef enum MtTextureSlot {
    MT_TEX_LOW_FREQ = 0,  /* sampler3D cloudBaseShapeSampler        128^3 RGBA8_UNORM (Sky.cpp:31-34) */
    MT_TEX_HIGH_FREQ = 1, /* sampler3D cloudDetailsHighFreqSampler  32^3  RGBA8_UNORM (Sky.cpp:40-43) */
    MT_TEX_CURL = 2,      /* sampler2D curlNoiseSampler             128^2 RGBA8_UNORM (Sky.cpp:47-50) */
    MT_TEX_WEATHER = 3    /* sampler2D weatherMapSampler            512^2 RGBA8_UNORM (Sky.cpp:54-57); bound, never sampled */
} MtTextureSlot;

/* Images = Renderer::CreateResources (Renderer.cpp:1428-1447).  CUR/PREV are the ping-pong ROLES at call time. */
typedef enum MtImage {
    MT_IMAGE_CLOUD_CUR = 0,   /* currentFrameResultImage   RGBA32F, W*H*16 bytes (RGBA16F, W*H*8, with MT_STORAGE_F16) */
    MT_IMAGE_CLOUD_PREV = 1,  /* previousFrameResultImage  RGBA32F               */
    MT_IMAGE_GODRAY_MASK = 2, /* godRaysCreationDataImage  RGBA32F               */
    MT_IMAGE_LDR = 3,         /* currentFrameTexture: tone-mapped (then TXAA'd) frame, W*H*4.  Stored here as RGBA8 UNORM -- what
                               * the shaders declare (`rgba8`, postProcess_ToneMap.frag:5); the reference CREATES the image as
                               * R8G8B8A8_SNORM (Renderer.cpp:1442-1445), a declared/actual mismatch that is deliberately not
                               * reproduced (INTEGRATION.md "Deviations").  The device pointer behind LDR / LDR_PREV changes
                               * after every mtDispatchTXAA / mtFrame: re-query mtImageDevicePtr per frame.              */
    MT_IMAGE_LDR_PREV = 4     /* previousFrameTexture: the LDR frame presented last frame (Renderer.cpp:1445)              */
} MtImage;

typedef enum MtPass {
    MT_PASS_REPROJECT = 0,
    MT_PASS_CLOUD = 1,
    MT_PASS_GODRAYS = 2,
    MT_PASS_TONEMAP = 3,
    MT_PASS_TXAA = 4,
    MT_PASS_COUNT = 5
} MtPass;

/* Work counters of the cloud pass (MT_FLAG_COUNTERS); same struct is filled by the CPU oracle. */
typedef struct MtCounters {
    uint64_t rays;          /* pixels processed by the cloud pass (incl. horizon-culled)      */
    uint64_t rays_marched;  /* rays that entered rayMarch                                      */
    uint64_t steps;         /* S: march iterations  (1 low-frequency fetch each)               */
    uint64_t steps_incloud; /* C: iterations with baseDensity > 0 (1 curl + 1 hi-freq + 6 low) */
    uint64_t cone_hits;     /* light-cone samples with density > 0                             */
    uint64_t early_exits;   /* rays that left the loop through accumDensity >= 1               */
} MtCounters;

/* Per-ray debug record (mtDispatchCloudDebug): every field must be bit-identical to the oracle's. */
typedef struct MtRayDebug {
    float dir[3];        /* ray direction from castRay                */
    float t_in, t_out;   /* inner / outer shell "t" (cloudRayMarch.comp:269 semantics) */
    float step_size;
    int32_t branch;      /* 0 ocean, 1 sky band, 2 marched            */
    int32_t steps;       /* loop iterations executed                  */
    uint32_t jitter_hash;/* FNV-1a over the per-step jitter indices   */
    float accum;         /* accumDensity when the loop ended          */
} MtRayDebug;

typedef struct MtContext MtContext;

/* ---- lifecycle ---------------------------------------------------------------------------- */
MT_API uint32_t mtAbiVersion(void);
MT_API const char* mtStatusString(MtStatus s);
MT_API void mtDefaultTuning(MtTuning* out);
/* Replaces Renderer::InitializeRenderer + CreateResources (Renderer.cpp:89-110, 1428-1447): allocates the two
 * ping-pong HDR images, the god-ray mask and the LDR image on `device`, zero-filled.                         */
MT_API MtStatus mtCreate(const MtConfig* cfg, MtContext** out);
MT_API void mtDestroy(MtContext* ctx);                 /* Renderer::~Renderer (Renderer.cpp:40-87) */
MT_API const char* mtGetLastError(const MtContext* ctx);/* text of the last non-OK status; "" if none */
/* Renderer::RecreateOnResize (Renderer.cpp:598-611).  Transactional: on failure (MT_ERR_OOM / MT_ERR_CUDA) the context keeps its
 * old size and images and stays usable.  On success images start as zeros, and everything that referred to the old images is
 * dropped: mtSetCloudOutput / mtSetCloudForward are cleared, device pointers from mtImageDevicePtr and handles from
 * mtExportImageHandle are stale -- peers must close, re-export and re-open them. */
MT_API MtStatus mtResize(MtContext* ctx, uint32_t width, uint32_t height);

/* ---- uniforms: copied at call time (reference: memcpy into the mapped UBO) ---------------- */
MT_API MtStatus mtSetCamera(MtContext* ctx, const MtCameraUBO* ubo);      /* Camera::CopyToGPUMemory, camera.cpp:50-53; set 2 / set 1 */
MT_API MtStatus mtSetCameraOld(MtContext* ctx, const MtCameraUBO* ubo);   /* cameraOld->CopyToGPUMemory, main.cpp:185-186             */
MT_API MtStatus mtSetTime(MtContext* ctx, const MtTimeUBO* ubo);          /* Scene::UpdateTime, Scene.cpp:65-85                       */
MT_API MtStatus mtSetSunAndSky(MtContext* ctx, const MtSunAndSkyUBO* ubo);/* Sky::UpdateSunAndSky, Sky.cpp:64-74                      */
MT_API MtStatus mtSetKeyPressQuery(MtContext* ctx, int32_t key_debug);    /* Scene::UpdateKeyPressQuery, Scene.cpp:144-147 (bound, unused) */
MT_API MtStatus mtSetTuning(MtContext* ctx, const MtTuning* tuning);      /* extension: shader #defines as data */

/* ---- textures: Sky::CreateCloudResources (Sky.cpp:25-58) ------------------------------------ */
/* rgba8 is tightly packed [z][y][x][4] (ImageLoadingUtility.cpp:87-98); sampler = LINEAR, REPEAT, normalized. */
MT_API MtStatus mtUploadTexture3D(MtContext* ctx, MtTextureSlot slot, uint32_t w, uint32_t h, uint32_t d, const uint8_t* rgba8);
MT_API MtStatus mtUploadTexture2D(MtContext* ctx, MtTextureSlot slot, uint32_t w, uint32_t h, const uint8_t* rgba8);

/* ---- dispatches ------------------------------------------------------------------------------ */
/* vkCmdDispatch of reprojectionPipeline, Renderer.cpp:683-698: PREV -> CUR, every pixel. */
MT_API MtStatus mtDispatchReprojection(MtContext* ctx);
/* vkCmdDispatch of cloudComputePipeline, Renderer.cpp:701-716: marches pixel (4*gx + id/4, 4*gy + id%4),
 * id = Time.frameCountMod16, over the reference's (over-provisioned, truncated) grid; writes CUR and the mask. */
MT_API MtStatus mtDispatchCloud(MtContext* ctx);
/* All 16 ids in one launch with the current camera/time: every pixel the 16 reference dispatches would write
 * (BASELINE config 3, "full quality, no reprojection").                                                       */
MT_API MtStatus mtDispatchCloudFull(MtContext* ctx);
/* Same as mtDispatchCloudFull restricted to 4-pixel-row tiles [tile_begin, tile_end) stepping by tile_stride,
 * each tile `tile_rows` pixel rows high (multiple of 4): the row-tile shard of one GPU (multi-GPU extension). */
MT_API MtStatus mtDispatchCloudTiles(MtContext* ctx, uint32_t tile_rows, uint32_t tile_begin, uint32_t tile_end, uint32_t tile_stride);
/* Debug variant: full==0 -> like mtDispatchCloud, else like mtDispatchCloudFull; additionally writes one
 * MtRayDebug per pixel (W*H records, untouched for unwritten pixels) to host memory `out`.  Synchronous.  */
MT_API MtStatus mtDispatchCloudDebug(MtContext* ctx, int full, MtRayDebug* out, size_t out_bytes);
/* vkCmdDraw(3) of postProcess_GodRays pipeline (Renderer.cpp:826-832; commented out in the reference frame). */
MT_API MtStatus mtDispatchGodRays(MtContext* ctx);
/* vkCmdDraw(3) of postProcess_ToneMap pipeline (Renderer.cpp:834-838): CUR -> LDR. */
MT_API MtStatus mtDispatchToneMap(MtContext* ctx);
/* vkCmdDraw(3) of postProcess_TXAA pipeline (Renderer.cpp:840-846): neighbourhood-clamped temporal blend of MT_IMAGE_LDR
 * with MT_IMAGE_LDR_PREV, result replaces MT_IMAGE_LDR (the image the reference presents).                      */
MT_API MtStatus mtDispatchTXAA(MtContext* ctx);
/* Debug variant of mtDispatchReprojection: also returns the 10 clamped tap indices (y*W+x) per pixel. Synchronous. */
MT_API MtStatus mtDispatchReprojectionDebug(MtContext* ctx, int32_t* taps, size_t taps_bytes);
/* One reference frame (Renderer::Frame, Renderer.cpp:122-192): REPROJ, CLOUD, [GODRAYS if with_godrays], TONEMAP,
 * then swap the ping-pong roles.  The caller copies camera -> cameraOld afterwards (main.cpp:185-186).         */
MT_API MtStatus mtFrame(MtContext* ctx, int with_godrays);
/* Same with a pass mask: REPROJ and CLOUD always run; MT_FRAME_GODRAYS / MT_FRAME_TONEMAP / MT_FRAME_TXAA select the post
 * passes.  The reference's live frame is MT_FRAME_TONEMAP | MT_FRAME_TXAA (god rays commented out, Renderer.cpp:826-846). */
#define MT_FRAME_GODRAYS 1u
#define MT_FRAME_TONEMAP 2u
#define MT_FRAME_TXAA 4u
MT_API MtStatus mtFrameEx(MtContext* ctx, uint32_t passes);
MT_API MtStatus mtSwapPingPong(MtContext* ctx);   /* swapPingPongBuffers = !swapPingPongBuffers, Renderer.cpp:191 */
MT_API MtStatus mtSynchronize(MtContext* ctx);    /* vkQueueWaitIdle */

/* ---- images ---------------------------------------------------------------------------------- */
MT_API MtStatus mtImageBytes(const MtContext* ctx, MtImage which, size_t* bytes);
MT_API MtStatus mtReadImage(MtContext* ctx, MtImage which, void* host, size_t bytes);        /* D2H + sync   */
MT_API MtStatus mtReadImageRows(MtContext* ctx, MtImage which, uint32_t row_begin, uint32_t row_end, void* host, size_t bytes);
/* Asynchronous read-back on the context's copy stream: starts when the work issued so far has produced the image, runs
 * concurrently with later dispatches, and any later pass that would overwrite that image waits for it.  `host` should be
 * pinned; it is valid after mtWaitReads (or mtSynchronize).  With the ping-pong swap this hides the read-back of frame k
 * behind the rendering of frame k+1.                                                                                  */
MT_API MtStatus mtReadImageAsync(MtContext* ctx, MtImage which, void* host, size_t bytes);
/* The god-ray image as one float per pixel (extension): the value the Cloud shader spreads over four channels with
 * EncodeFloatRGBA (cloudRayMarch.comp:106-112, 824-825), decoded on the device with the god-ray shader's own dot product
 * (postProcess_GodRays.frag:39-43) and read back like mtReadImageAsync -- W*H*4 bytes instead of W*H*16.  The god-ray image
 * is not ping-ponged, so both this and mtReadImageAsync(MT_IMAGE_GODRAY_MASK) copy from a device-side snapshot: the next
 * Cloud dispatch never waits for PCIe.                                                                                   */
MT_API MtStatus mtReadGodRayGreyAsync(MtContext* ctx, float* host, size_t bytes);
MT_API MtStatus mtWaitReads(MtContext* ctx);
/* Device-side join: work issued to the context after this call waits for every copy issued so far (no host sync). */
MT_API MtStatus mtJoinCopies(MtContext* ctx);
MT_API MtStatus mtWriteImage(MtContext* ctx, MtImage which, const void* host, size_t bytes); /* H2D, ordered */
MT_API MtStatus mtClearImages(MtContext* ctx);    /* zero all four images (reference: images are never cleared; zeros assumed) */
/* Zero-copy interop.  Asking for MT_IMAGE_GODRAY_MASK (here or through mtExportImageHandle) tells the library that the mask may be
 * written behind its back: from then on every god-ray dispatch rebuilds its decoded copy of the mask from the whole image, instead
 * of relying on the 1-of-16 Cloud kernel having kept it current. */
MT_API MtStatus mtImageDevicePtr(MtContext* ctx, MtImage which, void** dev_ptr);
/* Redirect the cloud pass's HDR / mask stores to caller-provided device memory (same layout and size as the
 * context's own images), e.g. a peer-mapped image on GPU 0 so that row tiles land there through NVLink stores
 * straight from the kernel epilogue.  NULL restores the context's own image.                                  */
MT_API MtStatus mtSetCloudOutput(MtContext* ctx, void* hdr_dev_ptr, void* mask_dev_ptr);
/* Gather by forwarding: with a peer image set here (mtOpenPeerImage pointer, same dimensions; NULL = off), every
 * mtDispatchCloudTiles keeps the march kernel's stores in this context's own HDR image and launches, on a high-priority
 * side stream, a small kernel that pushes each row tile into the peer image as soon as the march has finished it -- the
 * transfer overlaps the march tile by tile and no marching warp waits on NVLink.  mtJoinCopies orders the main stream
 * after it; mtSynchronize waits for it.  The god-ray mask stays local.  Ignored while mtSetCloudOutput is in effect.   */
MT_API MtStatus mtSetCloudForward(MtContext* ctx, void* peer_hdr_dev_ptr);
/* How the full-quality Cloud kernel stores its HDR pixels (extension; matters when mtSetCloudOutput points at a peer GPU):
 * MT_STORE_DIRECT  one 16-byte st.global per ray from the marching warp;
 * MT_STORE_BULK    a warp stages its 16x2 pixels in shared memory and two bulk asynchronous copies (cp.async.bulk, the TMA
 *                  engine) carry the two 256-byte row segments to the image -- the warp only waits until the engine has READ
 *                  the staging buffer, never for the remote write.  Same bytes either way.                                   */
typedef enum MtStoreMode { MT_STORE_DIRECT = 0, MT_STORE_BULK = 1 } MtStoreMode;
MT_API MtStatus mtSetCloudStoreMode(MtContext* ctx, MtStoreMode mode);
/* CUDA IPC plumbing for one-process-per-GPU sharding: export this context's image, map a peer's. 64-byte handles. */
MT_API MtStatus mtExportImageHandle(MtContext* ctx, MtImage which, uint8_t handle[64]);
MT_API MtStatus mtOpenPeerImage(MtContext* ctx, const uint8_t handle[64], void** dev_ptr);
MT_API MtStatus mtClosePeerImage(MtContext* ctx, void* dev_ptr);
/* Copy-engine variant of the gather: copies this context's rows of the row tiles [tile_begin, tile_end) step tile_stride
 * of image `which` into the same rows of a peer image (mtOpenPeerImage pointer, same dimensions) on the context's copy
 * stream, after the work issued so far; later dispatches overlap with the copies.  mtWaitReads / mtSynchronize wait.   */
MT_API MtStatus mtCopyTilesToPeer(MtContext* ctx, MtImage which, uint32_t tile_rows, uint32_t tile_begin, uint32_t tile_end,
                                  uint32_t tile_stride, void* peer_dev_ptr);

/* ---- uniform producers: the host-side state the reference keeps in Camera / Scene / Sky ------------------ */
/* Pure host functions (no context, no GPU): a C/C++ application links them instead of re-deriving glm's arithmetic.        */
typedef struct MtxCamera {       /* camera.h:20-68 */
    float eye[3], ref[3];        /* position, look-at point                                   */
    float forward[3], right[3], up[3];
    float fovy_deg, aspect, near_clip, far_clip;
    int32_t width, height;
} MtxCamera;
/* Camera::Camera + RecomputeAttributes (camera.cpp:3-17, 68-77).  Reference default: eye (0,0,2), ref (0,0,1), 45, .1, 1000. */
MT_API void mtxCameraInit(MtxCamera* cam, int32_t width, int32_t height, const float eye[3], const float ref[3], float fovy_deg,
                          float near_clip, float far_clip);
MT_API void mtxCameraRotateAboutUp(MtxCamera* cam, float deg);       /* camera.cpp:79-87  */
MT_API void mtxCameraRotateAboutRight(MtxCamera* cam, float deg);    /* camera.cpp:88-96  */
MT_API void mtxCameraTranslateAlongLook(MtxCamera* cam, float amt);  /* camera.cpp:98-104 */
MT_API void mtxCameraTranslateAlongRight(MtxCamera* cam, float amt); /* camera.cpp:105-111 */
MT_API void mtxCameraTranslateAlongUp(MtxCamera* cam, float amt);    /* camera.cpp:112-118 */
MT_API void mtxCameraUBO(const MtxCamera* cam, MtCameraUBO* out);    /* Camera::UpdateBuffer, camera.cpp:31-42 (glm lookAtRH / perspectiveRH_ZO) */
MT_API void mtxTimeInit(MtTimeUBO* t);                               /* Scene::InitializeTime, Scene.cpp:86-119: Halton base 3, frame 0 */
MT_API void mtxTimeUpdate(MtTimeUBO* t, float delta_seconds);        /* Scene::UpdateTime, Scene.cpp:65-85 with a caller-supplied delta   */
MT_API void mtxSunAndSky(MtSunAndSkyUBO* s);                         /* Sky::UpdateSunAndSky, Sky.cpp:64-74 */
/* One iteration of the reference main loop (main.cpp:172-194) for a context: uploads camera / cameraOld / time / sky,
 * runs mtFrameEx(passes), then copies camera into camera_old.                                                      */
MT_API MtStatus mtxRunFrame(MtContext* ctx, const MtxCamera* cam, MtCameraUBO* camera_old, MtTimeUBO* time, float delta_seconds,
                            uint32_t passes);

/* ---- asset pipeline: the reference's texture files -> RGBA8 arrays, without stb / PIL --------------------- */
/* TGA true-colour (raw / RLE, 24 / 32 bpp) and non-interlaced PNG (8 / 16 bit; 16-bit samples keep the high byte as stb_image
 * does).  Output is tightly packed RGBA8, rows top-down (what stbi_load(..., STBI_rgb_alpha) returns,
 * ImageLoadingUtility.cpp:91).  rgba8_out may be NULL to query the size.  MT_ERR_INVALID on any decode / IO failure.        */
MT_API MtStatus mtxDecodeImage(const uint8_t* file_bytes, size_t n, int is_png, uint8_t* rgba8_out, size_t out_bytes, uint32_t* w, uint32_t* h);
MT_API MtStatus mtxLoadImageFile(const char* path, uint8_t* rgba8_out, size_t out_bytes, uint32_t* w, uint32_t* h);
/* ImageLoadingUtility::create3DTextureFromMany2DTextures (ImageLoadingUtility.cpp:75-139): slice z = folder + base + "(z+1)" + ext. */
MT_API MtStatus mtxLoadVolumeFromSlices(const char* folder, const char* base_name, const char* extension, uint32_t w, uint32_t h, uint32_t d,
                                        uint8_t* rgba8_out, size_t out_bytes);
/* Packed cache of a decoded volume: "MTVOL001", u32 w, h, d, reserved, then raw RGBA8. */
MT_API MtStatus mtxSaveVolume(const char* path, uint32_t w, uint32_t h, uint32_t d, const uint8_t* rgba8);
MT_API MtStatus mtxLoadVolume(const char* path, uint8_t* rgba8_out, size_t out_bytes, uint32_t* w, uint32_t* h, uint32_t* d);

/* ---- measurement ----------------------------------------------------------------------------- */
MT_API MtStatus mtGetCounters(MtContext* ctx, MtCounters* out, int reset); /* sync; needs MT_FLAG_COUNTERS */
/* Device time (CUDA events on the context's stream) of the most recent dispatch of `pass`, in ms. Syncs. */
MT_API MtStatus mtLastPassMs(MtContext* ctx, MtPass pass, float* ms);
MT_API MtStatus mtStreamHandle(MtContext* ctx, void** cuda_stream);        /* cudaStream_t of the context */
/* User timing events on the context's stream (what bench.py brackets its timed region with): record into
 * slot 0..15, then read the device time between two recorded slots.  mtEventElapsedMs synchronises on `to`. */
MT_API MtStatus mtEventRecord(MtContext* ctx, uint32_t slot);
MT_API MtStatus mtEventElapsedMs(MtContext* ctx, uint32_t from, uint32_t to, float* ms);
/* Evict the L2 between timed iterations: overwrites an internal scratch buffer of `bytes` (rounded up to 256 MiB,
 * larger than the 126 MB L2) on the context's stream.  Not counted by mtLaunchCount.                            */
MT_API MtStatus mtFlushL2(MtContext* ctx, size_t bytes);
/* Micro-benchmark for the raymarch roofline: sustained FP32 FMA rate of the device in GFLOP/s (2 flop per FMA,
 * all SMs, register-resident independent chains).  The denominator for an issue-bound, non-tensor kernel.      */
MT_API MtStatus mtMeasureFp32Peak(MtContext* ctx, float* gflops);
/* Number of kernel launches issued by this context since creation (bench.py's gpu_launches). */
MT_API uint64_t mtLaunchCount(const MtContext* ctx);

#ifdef __cplusplus
}
#endif
#endif /* METEOROS_B200_H */
