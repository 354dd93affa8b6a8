// cloud_raymarch.cu -- the Cloud compute pass (cloudRayMarch.comp) as one sm_100a kernel.
//
// Launch shape: one thread per ray; a warp owns a 16x2 tile of rays (mt_params.h), a 128-thread CTA a 16x8 tile, so that the
// 32 rays of a warp stay inside the same few noise texels per step (pixel footprint << texel) and the HDR / mask
// stores of a warp are four 128-byte rows.  Below-horizon CTAs retire after ~100 instructions; the hardware CTA
// scheduler back-fills, so no persistent loop is needed (64 800 CTAs at 3840x2160).
//
// Compiled with -fmad=false: see mt_math.cuh.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "cloud_core.cuh"
#include "mt_launch.h"
#include "mt_pixel.cuh"
#include "post_core.cuh"

// The per-frame constants (MarchConst, mt_params.h) arrive in the parameter block: M is a reference into the constant bank.  The
// two jitter tables are indexed per lane, so every CTA stages them (48 words) in shared memory.  MT_MC_PARAM=0 is the A/B
// form: the whole block staged in shared memory, as rounds 1 and 2 did from a device buffer.
#ifndef MT_MC_PARAM
#define MT_MC_PARAM 1
#endif
#if MT_MC_PARAM
#define MT_MARCH_CONST(P)                                                                                                          \
    __shared__ MarchTabs J;                                                                                                        \
    if (threadIdx.x < MT_MARCHTABS_WORDS)                                                                                          \
        reinterpret_cast<float*>(&J)[threadIdx.x] = reinterpret_cast<const float*>(&(P).mc.tabs)[threadIdx.x];                    \
    __syncthreads();                                                                                                               \
    const MarchConst& M = (P).mc
#else
#define MT_MARCH_CONST(P)                                                                                                          \
    __shared__ MarchConst Ms;                                                                                                      \
    if (threadIdx.x < sizeof(MarchConst) / 4)                                                                                      \
        reinterpret_cast<float*>(&Ms)[threadIdx.x] = reinterpret_cast<const float*>(&(P).mc)[threadIdx.x];                        \
    __syncthreads();                                                                                                               \
    const MarchConst& M = Ms;                                                                                                      \
    const MarchTabs& J = Ms.tabs
#endif

// One thread per filter cell, a warp = the 32 cells of one bitmap word: each lane tests its cell's eight corner texels (coalesced
// along x, L1 / L2 hits), a ballot assembles the word.  Where the (r, F) bricks exist, bit 0 of each brick's first word repeats the
// cell's bit (the pipelined cone loop reads it from the brick it has loaded anyway); the other bits of that word are rf_pack of the
// cell's own texel, which the lane holds, so the word is stored whole -- no read-modify-write of a 64 MB volume.  (Rounds 1-2: one
// thread per 32 cells, 65 536 threads looping over 256 texel tests and 32 scattered read-modify-writes each: 71 us per coverage
// change, 9 % of the 256-view sweep.)
__global__ void __launch_bounds__(256) occupancy_build_kernel(Tex3D T, uint32_t* occ, float coverage, uint4* rfq)
{
    const unsigned cell = blockIdx.x * blockDim.x + threadIdx.x;            // (z*h + y)*w + x; w is a multiple of 32
    const unsigned ncells = (unsigned)T.w * (unsigned)T.h * (unsigned)T.d;
    if (cell >= ncells) return;                                              // whole warps: ncells is a multiple of 32
    const unsigned x = cell % (unsigned)T.w, y = (cell / (unsigned)T.w) % (unsigned)T.h, z = cell / ((unsigned)T.w * (unsigned)T.h);
    const unsigned x1 = (x + 1u) & (unsigned)(T.w - 1), y1 = (y + 1u) & (unsigned)(T.h - 1), z1 = (z + 1u) & (unsigned)(T.d - 1);
    const unsigned rows[4] = { (z * T.h + y) * T.w, (z * T.h + y1) * T.w, (z1 * T.h + y) * T.w, (z1 * T.h + y1) * T.w };
    const uint32_t own = __ldg(T.texels + rows[0] + x);
    bool any = occ_texel_may_be_cloud(own, coverage) || occ_texel_may_be_cloud(__ldg(T.texels + rows[0] + x1), coverage);
#pragma unroll
    for (int r = 1; r < 4; ++r)
        any = any || occ_texel_may_be_cloud(__ldg(T.texels + rows[r] + x), coverage) ||
              occ_texel_may_be_cloud(__ldg(T.texels + rows[r] + x1), coverage);
    const uint32_t bits = __ballot_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0) occ[cell >> 5] = bits;
#if MT_CONE_PIPE
    if (rfq) rfq[(MT_RF_BRICKS ? 2u : 1u) * cell].x = rf_pack(own) | (any ? 1u : 0u);  // = what build_rf_quads_kernel wrote, with this coverage's flag
#else
    (void)rfq;
#endif
}

// quads[(z*h + y)*w + x] = the 2x2 texels of filter cell (x, y) in slice z, REPEAT applied (d = 1 for 2D textures)
__global__ void __launch_bounds__(256) build_quads_kernel(const uint32_t* __restrict__ t, int w, int h, int d, uint4* __restrict__ q)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned n = (unsigned)w * (unsigned)h * (unsigned)d;
    if (i >= n) return;
    const unsigned x = i % (unsigned)w, y = (i / (unsigned)w) % (unsigned)h, z = i / ((unsigned)w * (unsigned)h);
    const unsigned x1 = (x + 1u) & (unsigned)(w - 1), y1 = (y + 1u) & (unsigned)(h - 1);
    const unsigned r0 = (z * h + y) * w, r1 = (z * h + y1) * w;
#if MT_TEX_BRICKS
    if (d > 1) {
        const unsigned z1 = (z + 1u) & (unsigned)(d - 1);
        const unsigned s0 = (z1 * h + y) * w, s1 = (z1 * h + y1) * w;
        q[2 * i] = make_uint4(t[r0 + x], t[r0 + x1], t[r1 + x], t[r1 + x1]);
        q[2 * i + 1] = make_uint4(t[s0 + x], t[s0 + x1], t[s1 + x], t[s1 + x1]);
        return;
    }
#endif
    q[i] = make_uint4(t[r0 + x], t[r0 + x1], t[r1 + x], t[r1 + x1]);
}
// the same quads of (r, F) words (mt_tex.cuh, rf_pack): the light-cone samples' form of the low-frequency volume
__global__ void __launch_bounds__(256) build_rf_quads_kernel(const uint32_t* __restrict__ t, int w, int h, int d, uint4* __restrict__ q)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned n = (unsigned)w * (unsigned)h * (unsigned)d;
    if (i >= n) return;
    const unsigned x = i % (unsigned)w, y = (i / (unsigned)w) % (unsigned)h, z = i / ((unsigned)w * (unsigned)h);
    const unsigned x1 = (x + 1u) & (unsigned)(w - 1), y1 = (y + 1u) & (unsigned)(h - 1);
    const unsigned r0 = (z * h + y) * w, r1 = (z * h + y1) * w;
#if MT_RF_BRICKS
    {   // the cell's two slices back to back: 32 bytes, one 256-bit load per light-cone sample
        const unsigned z1 = (z + 1u) & (unsigned)(d - 1);
        const unsigned s0 = (z1 * h + y) * w, s1 = (z1 * h + y1) * w;
        // bit 0 of the first word = "this cell may hold cloud" (set; occupancy_build_kernel clears it per coverage where it can)
        q[2 * i] = make_uint4(rf_pack(t[r0 + x]) | 1u, rf_pack(t[r0 + x1]), rf_pack(t[r1 + x]), rf_pack(t[r1 + x1]));
        q[2 * i + 1] = make_uint4(rf_pack(t[s0 + x]), rf_pack(t[s0 + x1]), rf_pack(t[s1 + x]), rf_pack(t[s1 + x1]));
        return;
    }
#endif
    // bit 0 of the first word = "this cell may hold cloud" (set; occupancy_build_kernel clears it per coverage where it can)
    q[i] = make_uint4(rf_pack(t[r0 + x]) | 1u, rf_pack(t[r0 + x1]), rf_pack(t[r1 + x]), rf_pack(t[r1 + x1]));
}
cudaError_t mt_launch_build_rf_quads(const uint32_t* texels, int w, int h, int d, void* quads, cudaStream_t stream)
{
    const unsigned n = (unsigned)w * (unsigned)h * (unsigned)d;
    build_rf_quads_kernel<<<(n + 255) / 256, 256, 0, stream>>>(texels, w, h, d, (uint4*)quads);
    return cudaGetLastError();
}

cudaError_t mt_launch_build_quads(const uint32_t* texels, int w, int h, int d, void* quads, cudaStream_t stream)
{
    const unsigned n = (unsigned)w * (unsigned)h * (unsigned)d;
    build_quads_kernel<<<(n + 255) / 256, 256, 0, stream>>>(texels, w, h, d, (uint4*)quads);
    return cudaGetLastError();
}

cudaError_t mt_launch_occupancy(const Tex3D& low, uint32_t* occ, float coverage, cudaStream_t stream)
{
    const unsigned ncells = (unsigned)low.w * (unsigned)low.h * (unsigned)low.d;  // w is a multiple of 32 (mtUploadTexture3D builds the bitmap only then)
    occupancy_build_kernel<<<(ncells + 255) / 256, 256, 0, stream>>>(low, occ, coverage, (uint4*)low.rfquads);
    return cudaGetLastError();
}


template <bool FULL, bool COUNT, bool DEBUG, bool WEATHER, bool STD, bool HWF = false>  // HWF: MT_FLAG_HW_CONE_FILTER (cloud_core.cuh, STD == 4)
#ifndef MT_CLOUD_MINBLOCKS
#define MT_CLOUD_MINBLOCKS 8  /* 64 registers/thread: 8 CTAs = 32 warps per SM (profiles/r1_cloud_ab.md) */
#endif
#ifndef MT_STREAM_STORES
#define MT_STREAM_STORES 1    /* full-quality kernel: HDR / mask stores carry the evict-first hint (mt_pixel.cuh) */
#endif
#ifndef MT_CONE_CACHE
#define MT_CONE_CACHE 1       /* per-ray light-cone offsets in shared memory (12 KB per CTA) */
#endif
__global__ void __launch_bounds__(128, MT_CLOUD_MINBLOCKS) cloud_raymarch_kernel(const __grid_constant__ CloudParams P)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int px, py, pixelID;
    int doneSlot = 0;  // FULL: index of this CTA's row tile within the launch (tile_forward_kernel's completion counters)
    bool valid;
    if (FULL) {
        // warp = MT_WARP_SHAPE ray tile (default 16x2: two 256-byte rows per float4 store), CTA = MT_CTA_W x MT_CTA_H rays
#if MT_WARP_SHAPE == 3   // 32x1 rays per warp, CTA = 1x4 warps (32x4 rays)
        const int lx = lane;
        const int ly = warp;
#elif MT_WARP_SHAPE == 1 // 16x2 rays per warp, CTA = 1x4 warps
        const int lx = lane & 15;
        const int ly = (warp << 1) + (lane >> 4);
#elif MT_WARP_SHAPE == 2  // 4x8 rays per warp, CTA = 4x1 warps
        const int lx = (warp << 2) + (lane & 3);
        const int ly = lane >> 2;
#else                     // 8x4 rays per warp, CTA = 2x2 warps
        const int lx = ((warp & 1) << 3) + (lane & 7);
        const int ly = ((warp >> 1) << 2) + (lane >> 3);
#endif
        const int bpt = P.rows.tile_rows / MT_CTA_H;            // CTAs per row tile, vertically
        // CTAs are dispatched in blockIdx order; mt_tile_order (mt_params.h) walks the marching tiles from the horizon upwards
        // and interleaves the ocean / sky-band tiles with them.
        const int by = (int)blockIdx.y;
        const int jtile = by / bpt;
        const int inTile = by - jtile * bpt;
        const int ltile = mt_tile_order(P.rows, jtile);
        doneSlot = ltile;
        const int tile = P.rows.tile_begin + ltile * P.rows.tile_stride;
        px = blockIdx.x * MT_CTA_W + lx;
        py = tile * P.rows.tile_rows + inTile * MT_CTA_H + ly;
        pixelID = ((px & 3) << 2) | (py & 3);                   // id = pX*4 + pY with (pX,pY) = (px%4, py%4)
        valid = px < P.W && py < P.H && (px >> 2) < P.tx && (py >> 2) < P.ty;
    } else {
        // 1-of-16 dispatch: only (W/4)x(H/4) rays, fewer warps than the GPU has slots, so every CTA is resident from
        // the start and nothing re-balances.  Warp q of CTA k therefore takes ray tile q*(NT/4)+k: one warp from each
        // horizontal quarter of the frame (sky .. ocean), which gives every CTA -- hence every SM -- the same mix of
        // marching and horizon-culled rays.
        const int tilesX = P.tx >> 3;                           // 8x4-ray tiles per row (tx is a multiple of 32)
        const int tileIdx = warp * (int)gridDim.x + (int)blockIdx.x;
        const int tyi = tileIdx / tilesX, txi = tileIdx - tyi * tilesX;
        const int gx = txi * 8 + (lane & 7), gy = tyi * 4 + (lane >> 3);
        pixelID = P.tm.frameCountMod16;
        px = gx * 4 + (pixelID >> 2);                           // pX = id/4
        py = gy * 4 + (pixelID & 3);                            // pY = id%4
        valid = gx < P.tx && gy < P.ty && px < P.W && py < P.H; // imageStore outside the image is dropped
    }

    MT_MARCH_CONST(P);

    // per-ray light-cone offsets (cloud_core.cuh, ConeOffsets): [sample][thread], written once per marching ray
#if MT_CONE_CACHE
    __shared__ __align__(16) F4 coneXYZ[6][128];
    F4* const cxyz = &coneXYZ[0][threadIdx.x];
#else
    F4* const cxyz = nullptr;
#endif
    // staging of the bulk-store epilogue (mtSetCloudStoreMode): one 16x2 pixel tile per warp, row-major = lane order
    __shared__ __align__(128) float4 outStage[4][32];
    RayCounters cnt = { 0u, 0u, 0u, 0u, 0u, 0u };
    const bool bulk = FULL && MT_WARP_SHAPE == 1 && P.bulkStore && __all_sync(0xffffffffu, valid);  // warp-uniform
    // MT_WARP_QUEUE (cloud_core.cuh): the production full-quality kernel marches a warp's rays together, in-cloud steps handed round
    constexpr bool QUEUED = MT_WARP_QUEUE && MT_CONE_CACHE && MT_CONE_PIPE && MT_PARK_BG && FULL && !COUNT && !DEBUG && !WEATHER && STD;
    F4 hdr, mask;
#if MT_WARP_QUEUE && MT_CONE_CACHE
    if constexpr (QUEUED) {
        __shared__ __align__(16) WarpQueue wq[4];
        cloud_ray_queued<2>(P, M, J, px, py, pixelID, valid, hdr, mask, &coneXYZ[0][warp << 5], 128, wq[warp]);
    }
#endif
    if (valid) {
        const size_t idx = (size_t)py * (size_t)P.W + (size_t)px;
        if constexpr (!QUEUED)
            cloud_ray<COUNT, DEBUG, WEATHER, (HWF ? 4 : STD ? (MT_CONE_PIPE ? 2 : 1) : 0)>(P, M, J, px, py, pixelID, hdr, mask, cnt, DEBUG ? (P.debug + idx) : nullptr, cxyz, 128);
        const float4 h4 = make_float4(hdr.x, hdr.y, hdr.z, hdr.w);
        if (bulk) {
            // The marching warp must not wait on a remote write: its 32 pixels go to shared memory, and lanes 0 and 16 each
            // hand one row segment (256 bytes, 128 in RGBA16F) to the bulk-copy engine, then wait only until the engine has
            // read the buffer.
            unsigned src;
            if (P.storage == MT_PX_F16) {
                reinterpret_cast<uint2*>(&outStage[warp][0])[lane] = px_pack_f16(h4);
                src = (unsigned)__cvta_generic_to_shared(reinterpret_cast<uint2*>(&outStage[warp][0]) + lane);
            } else {
                outStage[warp][lane] = P.storage == MT_PX_F16_EMULATE ? px_round_f16(h4) : h4;
                src = (unsigned)__cvta_generic_to_shared(&outStage[warp][lane]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
            __syncwarp();
            if ((lane & 15) == 0) {
                if (P.storage == MT_PX_F16)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(reinterpret_cast<uint2*>(P.hdr) + idx), "r"(src)
                                 : "memory");
                else
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 256;" ::"l"(reinterpret_cast<float4*>(P.hdr) + idx), "r"(src)
                                 : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else if (FULL && MT_STREAM_STORES) {
            px_store_streaming(P.hdr, idx, h4, P.storage);
        } else {
            px_store(P.hdr, idx, h4, P.storage);
        }
        if (FULL && MT_STREAM_STORES) px_store_streaming(P.mask, idx, make_float4(mask.x, mask.y, mask.z, mask.w), P.storage);
        else px_store(P.mask, idx, make_float4(mask.x, mask.y, mask.z, mask.w), P.storage);
    }
    if (FULL && !COUNT && !DEBUG && P.tileDone) {  // uniform: tell tile_forward_kernel that this CTA's pixels are in memory
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(P.tileDone + doneSlot, 1u);
    }
    if (COUNT) {
        unsigned v[6] = { cnt.rays, cnt.marched, cnt.steps, cnt.incloud, cnt.cone, cnt.early };
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            unsigned s = v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0 && s) atomicAdd(P.counters + k, (unsigned long long)s);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Step-parallel form of the 1-of-16 dispatch.  The reference's real-time frame marches only (W/4)x(H/4) rays: 4 050
// warps at 1080p for 148 SMs x 64 warp slots, each warp a ~50 000-instruction dependent chain -- the monolithic
// kernel is latency bound there (0.41 ms, IPC 0.23 per scheduler).  But a march step depends on the steps before it
// only through three running sums; the sample itself (position, densities, light cone, light energy) needs just t_k,
// and t_k = t_in + stepSize + ... + stepSize is k additions.  So: (1) one thread per ray does castRay / horizon
// branches / shells, files a 64-byte record and runs the loop's t sequence once, filing every t_k; (2) one thread per
// (ray, step) reads its t_k (the very value the sequential loop holds after k additions) and evaluates the sample --
// 64x more parallelism, the same arithmetic; (3) one thread per ray folds the
// samples in step order (the only sequential part, ~8 instructions per step) and composites.  Every value is produced
// by the same device functions in the same order as in the monolithic kernel, so the output is bit-identical.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sixteenth_pixel(const CloudParams& P, int& px, int& py, int& pixelID, bool& valid)
{
    const int tilesX = P.tx >> 3;  // 8x4-ray tiles; the record / sample index of a ray is its global thread id
    const int tileIdx = (int)blockIdx.x * 4 + (int)(threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int tyi = tileIdx / tilesX, txi = tileIdx - tyi * tilesX;
    const int gx = txi * 8 + (lane & 7), gy = tyi * 4 + (lane >> 3);
    pixelID = P.tm.frameCountMod16;
    px = gx * 4 + (pixelID >> 2);
    py = gy * 4 + (pixelID & 3);
    valid = gx < P.tx && gy < P.ty && px < P.W && py < P.H;
}

__device__ __forceinline__ void store_pixel(const CloudParams& P, size_t idx, F4 hdr, F4 mask)
{
    px_store(P.hdr, idx, make_float4(hdr.x, hdr.y, hdr.z, hdr.w), P.storage);
    px_store(P.mask, idx, make_float4(mask.x, mask.y, mask.z, mask.w), P.storage);
}
// The god-ray pass filters a DECODED copy of the mask (post_core.cuh: pairs (d(x, y), d(x+1, y)) with a one-texel ring), which
// mask_decode_kernel rebuilds from the whole mask -- 33 MB read for the sixteenth of it a frame's Cloud dispatch rewrote.  The
// fused 1-of-16 kernel therefore keeps that copy current itself: the texel it has just stored, decoded from the value AS STORED
// (rounded through binary16 where the images are), goes into its own pair and its left neighbour's.  The host tracks whether the
// copy is current (mt_context.cu) and skips the decode launch when it is.
__device__ __forceinline__ void store_pixel_decoded(const CloudParams& P, int px, int py, F4 hdr, F4 mask)
{
    store_pixel(P, (size_t)py * P.W + px, hdr, mask);
    if (P.decoded) {
        float4 m = make_float4(mask.x, mask.y, mask.z, mask.w);
        if (P.storage != MT_PX_F32) m = px_round_f16(m);
        F4 t;
        t.x = m.x; t.y = m.y; t.z = m.z; t.w = m.w;
        const float d = mask_texel_decode(t);
        float2* e = P.decoded + ((size_t)(py + 1) * (size_t)P.decodedPitch + (size_t)(px + 1));
        e->x = d;
        e[-1].y = d;
    }
}

__global__ void __launch_bounds__(128) cloud_rays_kernel(const __grid_constant__ CloudParams P)
{
    MT_MARCH_CONST(P);
    int px, py, pixelID;
    bool valid;
    sixteenth_pixel(P, px, py, pixelID, valid);
    __shared__ int ctaSteps;
    if (threadIdx.x == 0) ctaSteps = 0;
    __syncthreads();
    RaySetup* rec = reinterpret_cast<RaySetup*>(P.rays) + ((size_t)blockIdx.x * 128 + threadIdx.x);
    int n = 0;
    if (!valid) {
        rec->branch = -1;
    } else {
        F4 hdr, mask;
        mask.x = mask.y = mask.z = mask.w = 0.0f;
        const RaySetup R = cloud_ray_setup(P, M, J, px, py, pixelID, hdr);
        *rec = R;
        if (R.branch != 2) store_pixel(P, (size_t)py * P.W + px, hdr, mask);  // ocean / sky band: final
        else {
            // the march loop's own t sequence (t += stepSize, one rounding per step), filed per step: a sample's thread
            // reads its t_k instead of repeating k dependent additions
            float* tk = reinterpret_cast<float*>(P.samples + ((size_t)blockIdx.x * 128 + threadIdx.x));
            const size_t stride2 = 2 * (size_t)gridDim.x * 128;
            for (float t = R.t_in; t < R.t_out && n < MT_STEP_SLICES; t += R.stepSize) tk[(size_t)(n++) * stride2] = t;
        }
        rec->nsteps = n;
    }
    // the largest step count of the CTA's 128 rays: slices beyond it (and every slice of a horizon-culled CTA) have
    // nothing to do, and cloud_steps_kernel learns that from one load
    n = max(n, __shfl_xor_sync(0xffffffffu, n, 16)); n = max(n, __shfl_xor_sync(0xffffffffu, n, 8));
    n = max(n, __shfl_xor_sync(0xffffffffu, n, 4));  n = max(n, __shfl_xor_sync(0xffffffffu, n, 2));
    n = max(n, __shfl_xor_sync(0xffffffffu, n, 1));
    if ((threadIdx.x & 31) == 0) atomicMax(&ctaSteps, n);
    __syncthreads();
    if (threadIdx.x == 0) {
        P.ctaSteps[blockIdx.x] = ctaSteps;
        if (blockIdx.x == 0) *P.itemCount = 0u;  // the in-cloud list of this frame starts empty (cloud_base_kernel runs next)
    }
}

#ifndef MT_STEPS_MINBLOCKS
#define MT_STEPS_MINBLOCKS 12  /* 39 registers, no spills: 48 warps per SM; the step-parallel march is latency bound (profiles/r1_ab.md) */
#endif
template <bool WEATHER, bool STD>
__global__ void __launch_bounds__(128, MT_STEPS_MINBLOCKS) cloud_steps_kernel(const __grid_constant__ CloudParams P)
{
    const int k = blockIdx.y;
    if (k >= __ldg(P.ctaSteps + blockIdx.x)) return;  // whole CTA idle for this slice (uniform: taken by all 128 threads)
    MT_MARCH_CONST(P);
    const size_t ray = (size_t)blockIdx.x * 128 + threadIdx.x;
    float2* slot = P.samples + ((size_t)k * (size_t)P.rayStride + ray);
    const float t = *reinterpret_cast<const float*>(slot);  // t_k as the sequential loop rounds it (cloud_rays_kernel); plain load: the slot is overwritten below
    const RaySetup R = reinterpret_cast<const RaySetup*>(P.rays)[ray];
    if (R.branch != 2 || k >= R.nsteps) return;
    const int jidx = (P.tm.frameCountMod16 + mt_f2i(t)) & 15;
    RayCounters none = { 0u, 0u, 0u, 0u, 0u, 0u };
    const ConeOffsets noCache = { nullptr, 0 };  // one thread per (ray, step): nothing to share
    const StepSample S = cloud_step_sample<false, WEATHER, STD>(P, M, J, R, jidx, t, none, noCache);
    *slot = make_float2(S.inc, S.energy);
}

// The step-parallel march in two kernels (-DMT_STEP_COMPACT=1; OFF by default, see the measurement below): lanes of a warp are 32 neighbouring rays at the
// same step index, and at a cloud's edge only some of them are inside it -- in cloud_steps_kernel the others idle
// through the ~1 500-instruction lighting part (20.6 of 32 lanes active at 1080p).  So cloud_base_kernel evaluates only
// the sample point and its base density for every (step, ray), files the misses as (0, -1) and appends the hits to a
// list (one atomicAdd per warp); cloud_light_kernel then walks that list with every lane busy.  Which thread evaluates a
// sample does not change its arithmetic, and results are stored by (step, ray), so the output is bit-identical and
// independent of the list's order (the -m gpu suite passes with it).  Measured (profiles/r1_ab.md): 1080p 183.5 vs 189.4 us,
// but 4K 664.8 vs 627.7 us -- the step-parallel march is latency bound, not issue bound, and the list walk adds a
// dependent load chain (item -> ray record -> t -> sample point) while saving instructions that were not the limit.
#ifndef MT_STEP_COMPACT
#define MT_STEP_COMPACT 0
#endif
#define MT_ITEM_RAY_BITS 26

template <bool WEATHER, bool STD>
__global__ void __launch_bounds__(128, MT_CLOUD_MINBLOCKS) cloud_base_kernel(const __grid_constant__ CloudParams P)
{
    const int k = blockIdx.y;
    if (k >= __ldg(P.ctaSteps + blockIdx.x)) return;  // whole CTA idle for this slice (uniform: taken by all 128 threads)
    MT_MARCH_CONST(P);
    const size_t ray = (size_t)blockIdx.x * 128 + threadIdx.x;
    const RaySetup R = reinterpret_cast<const RaySetup*>(P.rays)[ray];
    float t = R.t_in;
    for (int i = 0; i < k; ++i) t += R.stepSize;  // the same k roundings the sequential loop performs
    const bool live = R.branch == 2 && t < R.t_out;
    bool hit = false;
    if (live) {
        const int jidx = (P.tm.frameCountMod16 + mt_f2i(t)) & 15;
        RayCounters none = { 0u, 0u, 0u, 0u, 0u, 0u };
        hit = cloud_step_base<false, WEATHER, STD>(P, M, J.stepJitter[jidx >> 1], R, t, none, ConeOffsets{ nullptr, 0 }).baseDensity > 0.0f;
        if (!hit) P.samples[(size_t)k * (size_t)P.rayStride + ray] = make_float2(0.0f, -1.0f);
    }
    const unsigned hits = __ballot_sync(0xffffffffu, hit);
    if (hits) {
        const int lane = threadIdx.x & 31, leader = __ffs(hits) - 1;
        unsigned at = 0u;
        if (lane == leader) at = atomicAdd(P.itemCount, (unsigned)__popc(hits));
        at = __shfl_sync(0xffffffffu, at, leader);
        if (hit) P.items[at + __popc(hits & ((1u << lane) - 1u))] = ((unsigned)k << MT_ITEM_RAY_BITS) | (unsigned)ray;
    }
}

template <bool WEATHER, bool STD>
__global__ void __launch_bounds__(128, MT_CLOUD_MINBLOCKS) cloud_light_kernel(const __grid_constant__ CloudParams P)
{
    MT_MARCH_CONST(P);
    const unsigned n = *reinterpret_cast<const volatile unsigned*>(P.itemCount);
    const unsigned step = gridDim.x * 128u;
    for (unsigned i = blockIdx.x * 128u + threadIdx.x; i < n; i += step) {
        const unsigned item = __ldg(P.items + i);
        const int k = (int)(item >> MT_ITEM_RAY_BITS);
        const size_t ray = item & ((1u << MT_ITEM_RAY_BITS) - 1u);
        const RaySetup R = reinterpret_cast<const RaySetup*>(P.rays)[ray];
        float t = R.t_in;
        for (int j = 0; j < k; ++j) t += R.stepSize;
        const int jidx = (P.tm.frameCountMod16 + mt_f2i(t)) & 15;
        RayCounters none = { 0u, 0u, 0u, 0u, 0u, 0u };
        const ConeOffsets noCache = { nullptr, 0 };
        const StepBase B = cloud_step_base<false, WEATHER, STD>(P, M, J.stepJitter[jidx >> 1], R, t, none, noCache);   // same arithmetic: B.baseDensity > 0 again
        const StepSample S = cloud_step_light<false, WEATHER, STD>(P, M, R, B, none, noCache);
        P.samples[(size_t)k * (size_t)P.rayStride + ray] = make_float2(S.inc, S.energy);
    }
}

__global__ void __launch_bounds__(128) cloud_fold_kernel(const __grid_constant__ CloudParams P)
{
    const size_t ray = (size_t)blockIdx.x * 128 + threadIdx.x;
    const RaySetup R = reinterpret_cast<const RaySetup*>(P.rays)[ray];
    if (R.branch != 2) return;
    int px, py, pixelID;
    bool valid;
    sixteenth_pixel(P, px, py, pixelID, valid);
    const size_t stride = (size_t)P.rayStride;
    const int n = R.nsteps;  // number of march iterations (cloud_rays_kernel ran the t sequence of the sequential loop)
    float accum = 0.0f, transmittance = 1.0f, color = 0.0f;
    bool stop = false;
    // the fold is the only sequential part; its loads do not depend on the running sums, so fetch eight steps at a time
    for (int k0 = 0; k0 < n && !stop; k0 += 8) {
        float2 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            v[j] = (k0 + j < n) ? __ldg(P.samples + (size_t)(k0 + j) * stride + ray) : make_float2(0.0f, -1.0f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (!stop && k0 + j < n) {
                StepSample S;
                S.inc = v[j].x; S.energy = v[j].y;
                stop = cloud_step_combine(S, accum, transmittance, color);
            }
        }
    }
    F4 hdr, mask;
    cloud_composite(R, accum, color, hdr, mask);
    store_pixel(P, (size_t)py * P.W + px, hdr, mask);
}

// the reference's texture extents (Sky.cpp:31-50): the STD kernels carry them as immediates and take the light-cone samples
// from the (r, F) quads (MT_FLAG_NO_CONE_RF or other extents: the generic kernels, canonical filter throughout)
// ... and floor their filter coordinates with the magic constant (mt_tex.cuh), exact for |u| < 2^22.  A conservative bound on
// every texture coordinate of the frame from its constants: ray length T (an eye inside the inner shell marches at most
// ~250 km, dir.y >= 0.06; otherwise eye altitude + two shell diameters), sample position <= 1.1 T / 12500, relative height
// <= T / 12500, wind skew = height * |wind| * |cloud_top_offset| * 0.009 + |windSkew|; times the largest extent (128), with a
// factor two to spare.  Frames outside it (an eye millions of metres up, a wind drift of tens of thousands of periods)
// run the generic kernels.
static bool mt_magic_floor_ok(const CloudParams& P)
{
    if (!MT_MAGIC_FLOOR) return true;
    const float ey = P.mc.eyePos.y;
    if (!(fabsf(ey) < 1e8f)) return false;  // also NaN
    const float T = (ey > -1000.0f && ey < 7000.0f) ? 3.0e5f : fabsf(ey) + 1.3e7f;
    const float wind = fabsf(P.tun.wind_direction[0]) + fabsf(P.tun.wind_direction[1]) + fabsf(P.tun.wind_direction[2]);
    const float drift = fmaxf(fabsf(P.mc.windSkew.x), fmaxf(fabsf(P.mc.windSkew.y), fabsf(P.mc.windSkew.z)));
    const float smax = 1.1f * T / 12500.0f + (T / 12500.0f) * wind * fabsf(P.tun.cloud_top_offset) * 0.009f + drift + 2.0f;
    return smax * 128.0f < 2097152.0f;  // 2^21; !(NaN < x)
}
static bool mt_std_dims(const CloudParams& P)
{
    return P.low.w == 128 && P.low.h == 128 && P.low.d == 128 && P.high.w == 32 && P.high.h == 32 && P.high.d == 32 &&
           P.curl.w == 128 && P.curl.h == 128 && (!MT_CONE_RF || P.low.rfquads != nullptr) &&  // STD also means: (r, F) quads exist
           mt_magic_floor_ok(P);
}

// ---------------------------------------------------------------------------------------------------------------------
// The 1-of-16 dispatch in ONE kernel (default).  The three-kernel form above pays for its parallelism with HBM: every
// (ray, step) sample travels through a [64][rays] scratch array (71 MB at 1080p, 54 MB of DRAM reads per frame for a pass
// that outputs 4 MB, profiles/r1_passes_1080p.md) and with two extra launches.  Here a CTA of eight warps owns one 8x4
// ray tile from start to finish and the samples never leave shared memory:
//   A  warp 0, one lane per ray: castRay / horizon branches / shells (cloud_ray_setup) and the march loop's own t
//      sequence, into shared memory (RaySetup[32], t[64][32]); ocean / sky-band pixels are stored right away;
//   B  all eight warps: warp w evaluates steps w, w+8, ... of the 32 rays (lane = ray), each sample from the filed t_k --
//      the value the sequential loop holds after k roundings -- into (inc, energy)[64][32];
//   C  warp 0 folds each ray's samples in step order, composites and stores.
// 26 KB of shared memory per CTA instead of 80 MB of global scratch; the same device functions in the same order as the
// sequential kernel, so the image is bit-identical (test_sixteenth_step_parallel_equals_sequential).  While warp 0 of one
// CTA is in A or C, the other resident CTAs of the SM (six) are in B.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef MT_S16_WARPS
#define MT_S16_WARPS 8
#endif
#ifndef MT_S16_CONE
#define MT_S16_CONE 3  /* 1080p: 139.7 (1), 135.6 us (3); 1: plain (r, F) cone loop; 2: the software-pipelined, unrolled one of the full-quality kernel; 3: plain loop, brick first (flag in the brick) */
#endif
#ifndef MT_S16_MINBLOCKS
#define MT_S16_MINBLOCKS 6
#endif
template <bool WEATHER, bool STD, bool HWF = false>  // HWF: MT_FLAG_HW_CONE_FILTER (cloud_core.cuh, STD == 5)
__global__ void __launch_bounds__(32 * MT_S16_WARPS, MT_S16_MINBLOCKS) cloud_sixteenth_kernel(const __grid_constant__ CloudParams P)
{
    __shared__ RaySetup rays[32];
    __shared__ float tk[MT_STEP_SLICES][32];
    __shared__ float2 smp[MT_STEP_SLICES][32];
    __shared__ f3 bgs[32];
    __shared__ int tileSteps;
    MT_MARCH_CONST(P);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the CTA's ray tile: 8x4 rays of the (tx, ty) grid; rows of tiles in mt_tile_order: the marching rows from the horizon
    // upwards first, interleaved with the ocean rows -- the kernel is about two waves of CTAs, so what runs last decides its tail
    const int tilesX = P.tx >> MT_S16_LOG2W;
    const int jrow = (int)blockIdx.x / tilesX, txi = (int)blockIdx.x - jrow * tilesX;
    const int tyi = mt_tile_order(P.rows, jrow);
    const int gx = txi * MT_S16_TW + (lane & (MT_S16_TW - 1)), gy = tyi * MT_S16_TH + (lane >> MT_S16_LOG2W);
    const int pixelID = P.tm.frameCountMod16;
    const int px = gx * 4 + (pixelID >> 2), py = gy * 4 + (pixelID & 3);
    const bool valid = gx < P.tx && gy < P.ty && px < P.W && py < P.H;
    if (warp == 0) {  // ---- A
        int n = 0;
        if (!valid) {
            rays[lane].branch = -1;
            rays[lane].nsteps = 0;
        } else {
            F4 hdr, mask;
            mask.x = mask.y = mask.z = mask.w = 0.0f;
            RaySetup R = cloud_ray_setup<false>(P, M, J, px, py, pixelID, hdr);  // geometry only: the sky is warp 1's
            if (R.branch == 0) store_pixel_decoded(P, px, py, hdr, mask);  // ocean: final
            else if (R.branch == 2)
                for (float t = R.t_in; t < R.t_out && n < MT_STEP_SLICES; t += R.stepSize) tk[n++][lane] = t;
            R.nsteps = n;
            rays[lane] = R;
        }
        n = max(n, __shfl_xor_sync(0xffffffffu, n, 16)); n = max(n, __shfl_xor_sync(0xffffffffu, n, 8));
        n = max(n, __shfl_xor_sync(0xffffffffu, n, 4));  n = max(n, __shfl_xor_sync(0xffffffffu, n, 2));
        n = max(n, __shfl_xor_sync(0xffffffffu, n, 1));
        if (lane == 0) tileSteps = n;
    } else if (warp == 1 && valid) {  // ---- A, beside warp 0: the background (Preetham sky, a dozen pow / exp) of the same 32 rays
        const f3 dir = cloud_ray_dir(P, M, J, px, py, pixelID);   // the same castRay: the same bits
        const float dotUp = (0.0f * dir.x + 1.0f * dir.y) + 0.0f * dir.z;
        if (!(dotUp < 0.0f)) {
            const f3 bg = cloud_ray_background(P, dir);
            if (dotUp < 0.06f) {  // sky band below the cloud fade-out: final (cloudRayMarch.comp:730-740)
                F4 hdr, mask;
                hdr.x = bg.x; hdr.y = bg.y; hdr.z = bg.z; hdr.w = 1.0f;
                mask.x = mask.y = mask.z = mask.w = 0.0f;
                store_pixel_decoded(P, px, py, hdr, mask);
            } else {
                bgs[lane] = bg;
            }
        }
    }
    __syncthreads();
    const int nmax = tileSteps;
    if (nmax == 0) return;  // horizon-culled tile (uniform)
    const RaySetup& R = rays[lane];
    {   // ---- B
        const int mine = R.branch == 2 ? R.nsteps : 0;
        const ConeOffsets noCache = { nullptr, 0 };  // a thread visits ~7 steps of its ray: not worth a cache
        RayCounters none = { 0u, 0u, 0u, 0u, 0u, 0u };
        for (int k = warp; k < nmax; k += MT_S16_WARPS) {
            if (k < mine) {
                const float t = tk[k][lane];
                const int jidx = (pixelID + mt_f2i(t)) & 15;
                const StepSample S = cloud_step_sample<false, WEATHER, (HWF ? 5 : STD ? MT_S16_CONE : 0)>(P, M, J, R, jidx, t, none, noCache);
                smp[k][lane] = make_float2(S.inc, S.energy);
            }
        }
    }
    __syncthreads();
    if (warp == 0 && R.branch == 2) {  // ---- C
        const int n = R.nsteps;
        float accum = 0.0f, transmittance = 1.0f, color = 0.0f;
        for (int k = 0; k < n; ++k) {
            const float2 v = smp[k][lane];
            StepSample S;
            S.inc = v.x; S.energy = v.y;
            if (cloud_step_combine(S, accum, transmittance, color)) break;
        }
        F4 hdr, mask;
        RaySetup Rc = R;
        Rc.bg = bgs[lane];
        cloud_composite(Rc, accum, color, hdr, mask);
        store_pixel_decoded(P, px, py, hdr, mask);
    }
}

cudaError_t mt_launch_cloud_sixteenth_fused(const CloudParams& P, cudaStream_t stream)
{
    const unsigned tiles = (unsigned)((P.tx / MT_S16_TW) * (P.ty / MT_S16_TH));  // tx, ty are multiples of 32
    if (P.tun.use_weather) cloud_sixteenth_kernel<true, false><<<tiles, 32 * MT_S16_WARPS, 0, stream>>>(P);
    else if (mt_std_dims(P) && P.hwCone && P.low.hwtex) cloud_sixteenth_kernel<false, true, true><<<tiles, 32 * MT_S16_WARPS, 0, stream>>>(P);  // opt-in
    else if (mt_std_dims(P)) cloud_sixteenth_kernel<false, true><<<tiles, 32 * MT_S16_WARPS, 0, stream>>>(P);
    else cloud_sixteenth_kernel<false, false><<<tiles, 32 * MT_S16_WARPS, 0, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t mt_launch_cloud_sixteenth_split(const CloudParams& P0, cudaStream_t stream, int* launches)
{
    CloudParams P = P0;
    const unsigned ctas = (unsigned)((P.tx / 8) * (P.ty / 4) / 4);  // tx, ty are multiples of 32
    P.rayStride = (int)(ctas * 128u);
    if ((size_t)P.rayStride > ((size_t)1 << MT_ITEM_RAY_BITS)) return cudaErrorInvalidValue;
    cloud_rays_kernel<<<ctas, 128, 0, stream>>>(P);
    const dim3 slices(ctas, MT_STEP_SLICES, 1);
    const bool std_dims = mt_std_dims(P);
#if MT_STEP_COMPACT
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const unsigned walkers = (unsigned)sms * MT_CLOUD_MINBLOCKS;   // one resident wave walks the whole list
    if (P.tun.use_weather) {
        cloud_base_kernel<true, false><<<slices, 128, 0, stream>>>(P);
        cloud_light_kernel<true, false><<<walkers, 128, 0, stream>>>(P);
    } else if (std_dims) {
        cloud_base_kernel<false, true><<<slices, 128, 0, stream>>>(P);
        cloud_light_kernel<false, true><<<walkers, 128, 0, stream>>>(P);
    } else {
        cloud_base_kernel<false, false><<<slices, 128, 0, stream>>>(P);
        cloud_light_kernel<false, false><<<walkers, 128, 0, stream>>>(P);
    }
    *launches = 4;
#else
    if (P.tun.use_weather) cloud_steps_kernel<true, false><<<slices, 128, 0, stream>>>(P);
    else if (std_dims) cloud_steps_kernel<false, true><<<slices, 128, 0, stream>>>(P);
    else cloud_steps_kernel<false, false><<<slices, 128, 0, stream>>>(P);
    *launches = 3;
#endif
    cloud_fold_kernel<<<ctas, 128, 0, stream>>>(P);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// Gather by forwarding (mtSetCloudForward): the march kernel keeps its stores local -- a marching warp never waits on
// NVLink -- and counts finished CTAs per row tile; this small kernel, launched on a high-priority stream beside it,
// waits for each of its tiles to complete and pushes it into the peer image with 16-byte loads / stores.  A handful of
// CTAs is enough: a rank moves ~66 MB of an 8K frame in the ~2.5 ms its share takes to march.
// ---------------------------------------------------------------------------------------------------------------------
//
// The wait is bounded (~seconds): should the march kernel never run beside this one, the kernel flags the overrun in
// the word after the last counter and leaves instead of hanging the device; mtSynchronize reports it.
#define MT_FORWARD_MAX_POLLS (1u << 24)
// W16 = 16-byte words per image row (W for RGBA32F, W / 2 for RGBA16F; W is even there: mtSetCloudForward checks)
__global__ void __launch_bounds__(256) tile_forward_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int W16, int H,
                                                           RowTiles rows, unsigned* tileDone, unsigned ctasPerTile)
{
    __shared__ int timedOut;
    for (int j = blockIdx.x; j < rows.tile_count; j += gridDim.x) {
        const int lt = mt_tile_order(rows, j);  // the order the march kernel issues its tiles in
        if (threadIdx.x == 0) {
            const volatile unsigned* done = tileDone + lt;
            unsigned polls = 0;
            while (*done < ctasPerTile && polls < MT_FORWARD_MAX_POLLS) { __nanosleep(256); ++polls; }
            timedOut = *done < ctasPerTile;
            if (timedOut) atomicExch(tileDone + rows.tile_count, 1u);
            __threadfence();
        }
        __syncthreads();
        if (timedOut) return;
        const size_t tile = (size_t)rows.tile_begin + (size_t)lt * (size_t)rows.tile_stride;
        const size_t r0 = tile * (size_t)rows.tile_rows;
        const size_t r1 = r0 + (size_t)rows.tile_rows < (size_t)H ? r0 + (size_t)rows.tile_rows : (size_t)H;
        const float4* s = src + r0 * (size_t)W16;
        float4* d = dst + r0 * (size_t)W16;
        const size_t n = (r1 - r0) * (size_t)W16;
        for (size_t i = threadIdx.x; i < n; i += 1024) {  // four independent 16-byte loads in flight per thread
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i + 256 * j < n) v[j] = __ldcg(s + i + 256 * j);  // L2: the pixels were written by other SMs
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i + 256 * j < n) d[i + 256 * j] = v[j];
        }
        __syncthreads();
    }
}

cudaError_t mt_launch_tile_forward(const void* src, void* dst, int W, int H, int bytesPerPixel, const RowTiles& rows, unsigned* tileDone,
                                   int ctas, cudaStream_t stream)
{
    if (rows.tile_count <= 0) return cudaSuccess;
    const unsigned ctasPerTile = (unsigned)((W + MT_CTA_W - 1) / MT_CTA_W) * (unsigned)(rows.tile_rows / MT_CTA_H);
    const int grid = ctas < rows.tile_count ? ctas : rows.tile_count;
    tile_forward_kernel<<<grid, 256, 0, stream>>>((const float4*)src, (float4*)dst, W * bytesPerPixel / 16, H, rows, tileDone, ctasPerTile);
    return cudaGetLastError();
}

cudaError_t mt_launch_cloud(const CloudParams& P, cudaStream_t stream)
{
    const int bpt = P.rows.tile_rows / MT_CTA_H;
    dim3 grid((unsigned)((P.W + MT_CTA_W - 1) / MT_CTA_W), (unsigned)(bpt * P.rows.tile_count), 1);
    if (!P.full) grid = dim3((unsigned)((P.tx / 8) * (P.ty / 4) / 4), 1, 1);  // tx, ty are multiples of 32
    dim3 block(128, 1, 1);
    if (grid.x == 0 || grid.y == 0) return cudaSuccess;
    const bool count = P.counters != nullptr, debug = P.debug != nullptr;
    const bool std_dims = mt_std_dims(P);
    if (P.tun.use_weather) {  // weather path: production variants only (mt_context.cu rejects counters / debug with it)
        if (P.full) cloud_raymarch_kernel<true, false, false, true, false><<<grid, block, 0, stream>>>(P);
        else cloud_raymarch_kernel<false, false, false, true, false><<<grid, block, 0, stream>>>(P);
    } else if (std_dims) {  // counting / debug variants follow the production kernel's path: same pixels
        if (P.full) {
            if (debug) cloud_raymarch_kernel<true, true, true, false, true><<<grid, block, 0, stream>>>(P);
            else if (count) cloud_raymarch_kernel<true, true, false, false, true><<<grid, block, 0, stream>>>(P);
            else if (P.hwCone && P.low.hwtex) cloud_raymarch_kernel<true, false, false, false, true, true><<<grid, block, 0, stream>>>(P);  // opt-in
            else cloud_raymarch_kernel<true, false, false, false, true><<<grid, block, 0, stream>>>(P);
        } else {
            if (debug) cloud_raymarch_kernel<false, true, true, false, true><<<grid, block, 0, stream>>>(P);
            else if (count) cloud_raymarch_kernel<false, true, false, false, true><<<grid, block, 0, stream>>>(P);
            else cloud_raymarch_kernel<false, false, false, false, true><<<grid, block, 0, stream>>>(P);
        }
    } else if (P.full) {
        if (debug) cloud_raymarch_kernel<true, true, true, false, false><<<grid, block, 0, stream>>>(P);
        else if (count) cloud_raymarch_kernel<true, true, false, false, false><<<grid, block, 0, stream>>>(P);
        else cloud_raymarch_kernel<true, false, false, false, false><<<grid, block, 0, stream>>>(P);
    } else {
        if (debug) cloud_raymarch_kernel<false, true, true, false, false><<<grid, block, 0, stream>>>(P);
        else if (count) cloud_raymarch_kernel<false, true, false, false, false><<<grid, block, 0, stream>>>(P);
        else cloud_raymarch_kernel<false, false, false, false, false><<<grid, block, 0, stream>>>(P);
    }
    return cudaGetLastError();
}
