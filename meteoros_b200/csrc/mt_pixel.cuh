// mt_pixel.cuh -- one HDR / god-ray-mask pixel in the context's storage format (MtStorage, include/meteoros_b200.h).
//   MT_STORAGE_F32          RGBA32F, 16 bytes: what the shaders declare (`rgba32f`, cloudRayMarch.comp:30-32)
//   MT_STORAGE_F16_EMULATE  RGBA32F holding binary16-representable values: every store rounds through binary16
//   MT_STORAGE_F16          RGBA16F, 8 bytes: what the reference actually allocates (VK_FORMAT_R16G16B16A16_SFLOAT,
//                           Renderer.cpp:1431-1440) -- half the HBM stream, the NVLink gather and the PCIe read-back
// F16 and F16_EMULATE hold the same values (the conversion back to fp32 is exact), so every pass computes the same bits
// from either; tests/test_gpu_parity.py::test_f16_storage holds them to that.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define MT_PX_F32 0
#define MT_PX_F16_EMULATE 1
#define MT_PX_F16 2

__device__ __forceinline__ uint2 px_pack_f16(float4 v)
{
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<const unsigned*>(&lo);
    r.y = *reinterpret_cast<const unsigned*>(&hi);
    return r;
}
__device__ __forceinline__ float4 px_unpack_f16(uint2 r)
{
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 px_round_f16(float4 v) { return px_unpack_f16(px_pack_f16(v)); }

// read-only image (never written by the running kernel): through the non-coherent path
__device__ __forceinline__ float4 px_load_ro(const void* img, size_t idx, int storage)
{
    if (storage == MT_PX_F16) return px_unpack_f16(__ldg(reinterpret_cast<const uint2*>(img) + idx));
    return __ldg(reinterpret_cast<const float4*>(img) + idx);
}
// image the running kernel also writes (each pixel by its own thread only)
__device__ __forceinline__ float4 px_load(const void* img, size_t idx, int storage)
{
    if (storage == MT_PX_F16) return px_unpack_f16(reinterpret_cast<const uint2*>(img)[idx]);
    return reinterpret_cast<const float4*>(img)[idx];
}
__device__ __forceinline__ void px_store(void* img, size_t idx, float4 v, int storage)
{
    if (storage == MT_PX_F16) {
        reinterpret_cast<uint2*>(img)[idx] = px_pack_f16(v);
    } else {
        if (storage == MT_PX_F16_EMULATE) v = px_round_f16(v);
        reinterpret_cast<float4*>(img)[idx] = v;
    }
}

// The same store with the evict-first ("streaming") hint: for the full-quality Cloud pass, whose 32 bytes per pixel (265 MB at
// 4K) are written once and not read again by the kernel -- routed through the L2 like ordinary stores they evict the 100 MB of
// noise copies every ray samples (ncu: 1.5 GB of DRAM reads per 4K launch with plain stores, profiles/r2_ab.md section 1).
__device__ __forceinline__ void px_store_streaming(void* img, size_t idx, float4 v, int storage)
{
    if (storage == MT_PX_F16) {
        const uint2 h = px_pack_f16(v);
        __stcs(reinterpret_cast<uint2*>(img) + idx, h);
    } else {
        if (storage == MT_PX_F16_EMULATE) v = px_round_f16(v);
        __stcs(reinterpret_cast<float4*>(img) + idx, v);
    }
}
