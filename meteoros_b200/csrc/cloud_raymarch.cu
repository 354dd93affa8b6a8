// cloud_raymarch.cu -- the Cloud compute pass (cloudRayMarch.comp) as one sm_100a kernel.
//
// Launch shape: one thread per ray; a warp owns an 8x4 tile of rays, a 128-thread CTA a 16x8 tile, so that the
// 32 rays of a warp stay inside the same few noise texels per step (pixel footprint << texel) and the HDR / mask
// stores of a warp are four 128-byte rows.  Below-horizon CTAs retire after ~100 instructions; the hardware CTA
// scheduler back-fills, so no persistent loop is needed (64 800 CTAs at 3840x2160).
//
// Compiled with -fmad=false: see mt_math.cuh.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "cloud_core.cuh"
#include "mt_launch.h"

__global__ void cloud_setup_kernel(CamU cam, TimeU tm, MtTuning tun, int W, int H, MarchConst* out)
{
    MarchConst m;
    cloud_frame_setup(cam, tm, tun, m);
    cloud_frame_jitter(tm, W, H, m);
    *out = m;
}

// One thread per 32 cells (one output word): bit x of the word = any of the cell's eight corner texels may carry cloud.
__global__ void __launch_bounds__(256) occupancy_build_kernel(Tex3D T, uint32_t* occ, float coverage)
{
    const unsigned wpr = (unsigned)T.w >> 5;
    const unsigned nwords = wpr * (unsigned)T.h * (unsigned)T.d;
    const unsigned wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwords) return;
    const unsigned xw = wi % wpr, y = (wi / wpr) % (unsigned)T.h, z = wi / (wpr * (unsigned)T.h);
    const unsigned y1 = (y + 1u) & (unsigned)(T.h - 1), z1 = (z + 1u) & (unsigned)(T.d - 1);
    const unsigned rows[4] = { (z * T.h + y) * T.w, (z * T.h + y1) * T.w, (z1 * T.h + y) * T.w, (z1 * T.h + y1) * T.w };
    uint32_t bits = 0;
    for (unsigned k = 0; k < 32; ++k) {
        const unsigned x = xw * 32u + k, x1 = (x + 1u) & (unsigned)(T.w - 1);
        bool any = false;
#pragma unroll
        for (int r = 0; r < 4; ++r)
            any = any || occ_texel_may_be_cloud(__ldg(T.texels + rows[r] + x), coverage) ||
                  occ_texel_may_be_cloud(__ldg(T.texels + rows[r] + x1), coverage);
        bits |= (any ? 1u : 0u) << k;
    }
    occ[wi] = bits;
}

cudaError_t mt_launch_occupancy(const Tex3D& low, uint32_t* occ, float coverage, cudaStream_t stream)
{
    const unsigned nwords = (unsigned)(low.w >> 5) * (unsigned)low.h * (unsigned)low.d;
    occupancy_build_kernel<<<(nwords + 255) / 256, 256, 0, stream>>>(low, occ, coverage);
    return cudaGetLastError();
}

__device__ __forceinline__ float f16_round(float x) { return __half2float(__float2half_rn(x)); }

template <bool FULL, bool COUNT, bool DEBUG>
__global__ void __launch_bounds__(128) cloud_raymarch_kernel(const __grid_constant__ CloudParams P)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = ((warp & 1) << 3) + (lane & 7);
    const int ly = ((warp >> 1) << 2) + (lane >> 3);
    const int bpt = P.rows.tile_rows >> 3;                      // CTAs per tile, vertically
    const int ltile = blockIdx.y / bpt;
    const int tile = P.rows.tile_begin + ltile * P.rows.tile_stride;
    const int gx = blockIdx.x * 16 + lx;
    const int gy = tile * P.rows.tile_rows + (blockIdx.y - ltile * bpt) * 8 + ly;

    int px, py, pixelID;
    bool valid;
    if (FULL) {
        px = gx; py = gy;
        pixelID = ((px & 3) << 2) | (py & 3);                   // id = pX*4 + pY with (pX,pY) = (px%4, py%4)
        valid = px < P.W && py < P.H && (px >> 2) < P.tx && (py >> 2) < P.ty;
    } else {
        pixelID = P.tm.frameCountMod16;
        px = gx * 4 + (pixelID >> 2);                           // pX = id/4
        py = gy * 4 + (pixelID & 3);                            // pY = id%4
        valid = gx < P.tx && gy < P.ty && px < P.W && py < P.H; // imageStore outside the image is dropped
    }

    // stage the per-frame constants: read as warp-uniform (broadcast) or 8-way indexed LDS from here on
    __shared__ MarchConst M;
    if (threadIdx.x < MT_MARCHCONST_WORDS)
        reinterpret_cast<float*>(&M)[threadIdx.x] = __ldg(reinterpret_cast<const float*>(P.mc) + threadIdx.x);
    __syncthreads();

    RayCounters cnt = { 0u, 0u, 0u, 0u, 0u, 0u };
    if (valid) {
        F4 hdr, mask;
        const size_t idx = (size_t)py * (size_t)P.W + (size_t)px;
        cloud_ray<COUNT, DEBUG>(P, M, px, py, pixelID, hdr, mask, cnt, DEBUG ? (P.debug + idx) : nullptr);
        if (P.f16_emulate) {
            hdr.x = f16_round(hdr.x); hdr.y = f16_round(hdr.y); hdr.z = f16_round(hdr.z); hdr.w = f16_round(hdr.w);
            mask.x = f16_round(mask.x); mask.y = f16_round(mask.y); mask.z = f16_round(mask.z); mask.w = f16_round(mask.w);
        }
        reinterpret_cast<float4*>(P.hdr)[idx] = make_float4(hdr.x, hdr.y, hdr.z, hdr.w);
        reinterpret_cast<float4*>(P.mask)[idx] = make_float4(mask.x, mask.y, mask.z, mask.w);
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

cudaError_t mt_launch_cloud_setup(const CloudParams& P, MarchConst* out, cudaStream_t stream)
{
    cloud_setup_kernel<<<1, 1, 0, stream>>>(P.cam, P.tm, P.tun, P.W, P.H, out);
    return cudaGetLastError();
}

cudaError_t mt_launch_cloud(const CloudParams& P, cudaStream_t stream)
{
    const int bpt = P.rows.tile_rows / 8;
    int cols = P.full ? P.W : P.tx;
    dim3 grid((unsigned)((cols + 15) / 16), (unsigned)(bpt * P.rows.tile_count), 1);
    dim3 block(128, 1, 1);
    if (grid.x == 0 || grid.y == 0) return cudaSuccess;
    const bool count = P.counters != nullptr, debug = P.debug != nullptr;
    if (P.full) {
        if (debug) cloud_raymarch_kernel<true, true, true><<<grid, block, 0, stream>>>(P);
        else if (count) cloud_raymarch_kernel<true, true, false><<<grid, block, 0, stream>>>(P);
        else cloud_raymarch_kernel<true, false, false><<<grid, block, 0, stream>>>(P);
    } else {
        if (debug) cloud_raymarch_kernel<false, true, true><<<grid, block, 0, stream>>>(P);
        else if (count) cloud_raymarch_kernel<false, true, false><<<grid, block, 0, stream>>>(P);
        else cloud_raymarch_kernel<false, false, false><<<grid, block, 0, stream>>>(P);
    }
    return cudaGetLastError();
}
