// mt_tex.cuh -- exact fp32 linear filtering of RGBA8_UNORM noise textures (sampler: LINEAR, REPEAT, normalized
// coordinates, one mip; Texture3D.cpp:92-134, Image.cpp:305-347).
//
// Canonical filter (identical, operation for operation, to oracle/meteoros_oracle.c tex3d_linear):
//   u = s*W - 0.5, i0 = floor(u), a = u - i0, i1 = i0 + 1 (both modulo W)           -- same for v, w
//   weight(k,j,i) = (wx_i * wy_j) * wz_k
//   channel = (w000*t000 (+fma) w001*t001 ... w111*t111) * (1/255)       texel order: z-major, x fastest
// The hardware filter (tex3D, cudaFilterModeLinear) quantises the weights to 8 fractional bits, which moves the
// density by up to ~2e-3 and with it every threshold the march branches on; it cannot meet the parity bar
// (DESIGN.md "Why not the hardware filter").  Texels are fetched through the read-only L1/tex data path.
// Texture extents must be powers of two (the reference's are 128^3, 32^3, 128^2, 512^2).
#pragma once

#include "mt_math.cuh"

struct Tex3D {
    const uint32_t* texels;  // packed RGBA8, little endian: r = bits 0..7
    int w, h, d;             // powers of two
};
struct Tex2D {
    const uint32_t* texels;
    int w, h;
};

struct LinAxis {
    int i0, i1;
    float w0, w1;
};

MT_DEVICE LinAxis lin_axis_repeat(float s, int n)
{
    LinAxis a;
    float u = s * (float)n - 0.5f;
    float fl = floorf(u);
    a.w1 = u - fl;
    a.w0 = 1.0f - a.w1;
    a.i0 = mt_f2i(fl) & (n - 1);
    a.i1 = (a.i0 + 1) & (n - 1);
    return a;
}

#define MT_B0(t) ((float)((t) & 0xffu))
#define MT_B1(t) ((float)(((t) >> 8) & 0xffu))
#define MT_B2(t) ((float)(((t) >> 16) & 0xffu))
#define MT_B3(t) ((float)((t) >> 24))

struct Rgba {
    float r, g, b, a;
};

MT_DEVICE Rgba tex3d_rgba(const Tex3D& T, float s, float t, float r)
{
    LinAxis X = lin_axis_repeat(s, T.w), Y = lin_axis_repeat(t, T.h), Z = lin_axis_repeat(r, T.d);
    const uint32_t* p00 = T.texels + (size_t)((Z.i0 * T.h + Y.i0) * T.w);
    const uint32_t* p01 = T.texels + (size_t)((Z.i0 * T.h + Y.i1) * T.w);
    const uint32_t* p10 = T.texels + (size_t)((Z.i1 * T.h + Y.i0) * T.w);
    const uint32_t* p11 = T.texels + (size_t)((Z.i1 * T.h + Y.i1) * T.w);
    uint32_t t000 = MT_LDG(p00 + X.i0), t001 = MT_LDG(p00 + X.i1);
    uint32_t t010 = MT_LDG(p01 + X.i0), t011 = MT_LDG(p01 + X.i1);
    uint32_t t100 = MT_LDG(p10 + X.i0), t101 = MT_LDG(p10 + X.i1);
    uint32_t t110 = MT_LDG(p11 + X.i0), t111 = MT_LDG(p11 + X.i1);
    float w00 = X.w0 * Y.w0, w01 = X.w1 * Y.w0, w10 = X.w0 * Y.w1, w11 = X.w1 * Y.w1;
    float w000 = w00 * Z.w0, w001 = w01 * Z.w0, w010 = w10 * Z.w0, w011 = w11 * Z.w0;
    float w100 = w00 * Z.w1, w101 = w01 * Z.w1, w110 = w10 * Z.w1, w111 = w11 * Z.w1;
    Rgba o;
#define MT_ACC(B)                                                                                                   \
    fmaf(w111, B(t111), fmaf(w110, B(t110), fmaf(w101, B(t101), fmaf(w100, B(t100),                                  \
         fmaf(w011, B(t011), fmaf(w010, B(t010), fmaf(w001, B(t001), w000 * B(t000))))))))
    const float inv255 = 1.0f / 255.0f;
    o.r = MT_ACC(MT_B0) * inv255;
    o.g = MT_ACC(MT_B1) * inv255;
    o.b = MT_ACC(MT_B2) * inv255;
    o.a = MT_ACC(MT_B3) * inv255;
#undef MT_ACC
    return o;
}

// Same filter, only the first three channels (the high-frequency volume's alpha is never read).
MT_DEVICE Rgba tex3d_rgb(const Tex3D& T, float s, float t, float r)
{
    LinAxis X = lin_axis_repeat(s, T.w), Y = lin_axis_repeat(t, T.h), Z = lin_axis_repeat(r, T.d);
    const uint32_t* p00 = T.texels + (size_t)((Z.i0 * T.h + Y.i0) * T.w);
    const uint32_t* p01 = T.texels + (size_t)((Z.i0 * T.h + Y.i1) * T.w);
    const uint32_t* p10 = T.texels + (size_t)((Z.i1 * T.h + Y.i0) * T.w);
    const uint32_t* p11 = T.texels + (size_t)((Z.i1 * T.h + Y.i1) * T.w);
    uint32_t t000 = MT_LDG(p00 + X.i0), t001 = MT_LDG(p00 + X.i1);
    uint32_t t010 = MT_LDG(p01 + X.i0), t011 = MT_LDG(p01 + X.i1);
    uint32_t t100 = MT_LDG(p10 + X.i0), t101 = MT_LDG(p10 + X.i1);
    uint32_t t110 = MT_LDG(p11 + X.i0), t111 = MT_LDG(p11 + X.i1);
    float w00 = X.w0 * Y.w0, w01 = X.w1 * Y.w0, w10 = X.w0 * Y.w1, w11 = X.w1 * Y.w1;
    float w000 = w00 * Z.w0, w001 = w01 * Z.w0, w010 = w10 * Z.w0, w011 = w11 * Z.w0;
    float w100 = w00 * Z.w1, w101 = w01 * Z.w1, w110 = w10 * Z.w1, w111 = w11 * Z.w1;
    Rgba o;
#define MT_ACC(B)                                                                                                   \
    fmaf(w111, B(t111), fmaf(w110, B(t110), fmaf(w101, B(t101), fmaf(w100, B(t100),                                  \
         fmaf(w011, B(t011), fmaf(w010, B(t010), fmaf(w001, B(t001), w000 * B(t000))))))))
    const float inv255 = 1.0f / 255.0f;
    o.r = MT_ACC(MT_B0) * inv255;
    o.g = MT_ACC(MT_B1) * inv255;
    o.b = MT_ACC(MT_B2) * inv255;
    o.a = 0.0f;
#undef MT_ACC
    return o;
}

// Bilinear, first two channels (curl noise: only .xy is read, cloudRayMarch.comp:545-546).
MT_DEVICE void tex2d_rg(const Tex2D& T, float s, float t, float& r, float& g)
{
    LinAxis X = lin_axis_repeat(s, T.w), Y = lin_axis_repeat(t, T.h);
    const uint32_t* p0 = T.texels + (size_t)(Y.i0 * T.w);
    const uint32_t* p1 = T.texels + (size_t)(Y.i1 * T.w);
    uint32_t t00 = MT_LDG(p0 + X.i0), t01 = MT_LDG(p0 + X.i1);
    uint32_t t10 = MT_LDG(p1 + X.i0), t11 = MT_LDG(p1 + X.i1);
    float w00 = X.w0 * Y.w0, w01 = X.w1 * Y.w0, w10 = X.w0 * Y.w1, w11 = X.w1 * Y.w1;
    const float inv255 = 1.0f / 255.0f;
    r = fmaf(w11, MT_B0(t11), fmaf(w10, MT_B0(t10), fmaf(w01, MT_B0(t01), w00 * MT_B0(t00)))) * inv255;
    g = fmaf(w11, MT_B1(t11), fmaf(w10, MT_B1(t10), fmaf(w01, MT_B1(t01), w00 * MT_B1(t00)))) * inv255;
}
