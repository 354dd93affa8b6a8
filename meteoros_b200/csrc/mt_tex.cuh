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

#include <string.h>

#include "mt_math.cuh"

// MT_TEX_QUADS: fetch from the per-cell quad copies (device default; measured 6.50 -> 5.77 ms at 4K, profiles/r1_ab.md).
// The host simulation reads the plain volumes: same texels, same arithmetic.
#ifndef MT_TEX_QUADS
#if defined(MT_HOSTSIM)
#define MT_TEX_QUADS 0
#else
#define MT_TEX_QUADS 1
#endif
#endif
#ifndef MT_TEX_BRICKS
#define MT_TEX_BRICKS 0
#endif
// MT_RF_BRICKS: the (r, F) copy of the low-frequency volume holds, per filter cell, the cell's 2x2x2 texels in 32 contiguous
// bytes (both slices' quads), fetched with ONE 256-bit load (LDG.E.256, new on sm_100) -- one sector and one LSU request per
// light-cone sample instead of two (64 MB instead of 32 MB; with the RGBA quads 97 MB of the 126 MB L2).
#ifndef MT_RF_BRICKS
#define MT_RF_BRICKS 1
#endif
// MT_FLAG_HW_CONE_FILTER (opt-in, per context): the light-cone samples of the full-quality kernel go through the texture unit's own
// trilinear filter (tex3D over a CUDA array of the low-frequency volume; 8-bit filter weights) instead of the exact fp32 filter.
// Never the default: radiance moves by up to ~1.6e-3 relative on a few pixels of a 4K frame (profiles/r2_ab.md, r2c_ab.md), beyond
// the 1e-3 bar; every decision-carrying value (march sample, erosion, accumulated density, mask, alpha) stays on the exact path.
#if defined(MT_HOSTSIM)
struct Quad { uint32_t x, y, z, w; };
#define MT_LDG_QUAD(p) (*(p))
#else
typedef uint4 Quad;
#define MT_LDG_QUAD(p) __ldg(p)
#endif

struct Tex3D {
    const uint32_t* texels;  // packed RGBA8, little endian: r = bits 0..7
    const Quad* quads;       // optional: per filter cell (x0,y0,z) its 2x2 texels {(x0,y0),(x1,y0),(x0,y1),(x1,y1)} (REPEAT
                             // applied), so that the eight corners of a trilinear fetch are TWO aligned 16-byte loads
                             // instead of eight scattered 4-byte loads (4x the memory: 32 MB for 128^3, L2 resident)
    const Quad* rfquads;     // optional (low-frequency volume): the same quads of (r, F) words, see rf_pack below
    unsigned long long hwtex; // MT_FLAG_HW_CONE_FILTER: cudaTextureObject_t over the same volume (LINEAR, WRAP, normalized float reads), else 0
    int w, h, d;             // powers of two
    const uint32_t* occ;     // optional: 1 bit per filter cell (x fastest, 32 cells per word), 0 = the cell's eight
                             // corner texels all have zero cloud density at the current coverage (see occ_* below)
};
struct Tex2D {
    const uint32_t* texels;
    const Quad* quads;  // optional, as above: one 16-byte load per bilinear fetch
    int w, h;
};

struct LinAxis {
    unsigned i0, i1;
    float w0, w1;
};

// index of filter cell (x0, y0, z0): addresses the quad copy and (>> 5, & 31) the empty-cell bitmap
MT_DEVICE unsigned tex_cell(const Tex3D& T, unsigned x0, unsigned y0, unsigned z0) { return (z0 * (unsigned)T.h + y0) * (unsigned)T.w + x0; }

// MAGIC (device, STD kernels): floor(u) without the conversion pipe.  u + 1.5 * 2^23 in round-down mode leaves floor(u) in the
// low mantissa bits (two's complement below the hidden bits, so `& (n - 1)` is the REPEAT wrap as before) and, minus the
// constant, as a float -- exactly floorf(u) for |u| < 2^22; x and y share ONE packed add.  The same floor, the same
// subtraction: bit-identical weights and indices.  The host selects these kernels only when the frame's constants keep every
// texture coordinate far inside that range (mt_std_dims, cloud_raymarch.cu).  F2I.FLOOR + I2FP per axis before: 2.5 % + 2.3 %
// of the march kernel's instructions, the former on the quarter-rate pipe (profiles/r2_cloud_final.md).
template <bool MAGIC = false>
MT_DEVICE LinAxis lin_axis_repeat(float s, int n)
{
    LinAxis a;
    float u = fmaf(s, (float)n, -0.5f);  // == s*n - 0.5 bit for bit: n is a power of two, the product is never rounded
#if !defined(MT_HOSTSIM)
    if (MAGIC) {
        float t;
        asm("add.rm.f32 %0, %1, %2;" : "=f"(t) : "f"(u), "f"(MT_FLOOR_MAGIC));
        a.w1 = u - (t - MT_FLOOR_MAGIC);
        a.w0 = 1.0f - a.w1;
        a.i0 = __float_as_uint(t) & (unsigned)(n - 1);
        a.i1 = (a.i0 + 1u) & (unsigned)(n - 1);
        return a;
    }
#endif
    int fi = mt_floor2i(u);   // F2I.FLOOR (XU) ...
    float fl = (float)fi;     // ... and I2FP back (ALU): == floorf(u) for |u| < 2^24, one XU op instead of two
    a.w1 = u - fl;
    a.w0 = 1.0f - a.w1;
    a.i0 = (unsigned)fi & (unsigned)(n - 1);
    a.i1 = (a.i0 + 1u) & (unsigned)(n - 1);
    return a;
}

// x and y axes of one sample computed as a pair (FMUL2 / FADD2), identical per-component arithmetic
template <bool MAGIC = false>
MT_DEVICE void lin_axes_xy(P2 st, int nx, int ny, LinAxis& X, LinAxis& Y)
{
    const P2 u = fma2(st, pk2((float)nx, (float)ny), bc2(-0.5f));  // exact product (power-of-two extent): == s*n - 0.5
#if !defined(MT_HOSTSIM)
    if (MAGIC) {
        P2 t;
        asm("add.rm.f32x2 %0, %1, %2;" : "=l"(t) : "l"(u), "l"(bc2(MT_FLOOR_MAGIC)));
        const P2 w1 = sub2(u, sub2(t, bc2(MT_FLOOR_MAGIC)));
        const P2 w0 = sub2(bc2(1.0f), w1);
        X.w1 = lo2(w1); X.w0 = lo2(w0);
        Y.w1 = hi2(w1); Y.w0 = hi2(w0);
        X.i0 = (unsigned)t & (unsigned)(nx - 1); X.i1 = (X.i0 + 1u) & (unsigned)(nx - 1);
        Y.i0 = (unsigned)(t >> 32) & (unsigned)(ny - 1); Y.i1 = (Y.i0 + 1u) & (unsigned)(ny - 1);
        return;
    }
#endif
    const float ux = lo2(u), uy = hi2(u);
    int fx = mt_floor2i(ux), fy = mt_floor2i(uy);
    P2 w1 = sub2(u, pk2((float)fx, (float)fy));
    P2 w0 = sub2(bc2(1.0f), w1);
    X.w1 = lo2(w1); X.w0 = lo2(w0);
    Y.w1 = hi2(w1); Y.w0 = hi2(w0);
    X.i0 = (unsigned)fx & (unsigned)(nx - 1); X.i1 = (X.i0 + 1u) & (unsigned)(nx - 1);
    Y.i0 = (unsigned)fy & (unsigned)(ny - 1); Y.i1 = (Y.i0 + 1u) & (unsigned)(ny - 1);
}


// Byte k of a packed texel as the float  c * 2^-133  -- WITHOUT an int->float conversion (I2F runs on the quarter-rate
// XU pipe and was 66 % of the kernel's critical pipe in the first profile, profiles/r1_cloud_v1.md).  Placing the
// byte at bits 16..23 of a zero word gives the bit pattern c << 16: for c < 128 a denormal, for c >= 128 exponent
// field 1 -- and binary32 is linear across that boundary, so the value is exactly c * 2^-133 for all 256 bytes.
// One PRMT on the ALU pipe.  The filter weights carry the compensating 2^120 (MT_WSCALE, folded into the two z
// weights) and the final 1/255 carries 2^13; powers of two commute with every rounding as long as nothing leaves
// the normal range (weights >= 2^-75, sums <= 255), so every result is bit-identical to the oracle's
// (w * float(c)) chain.  Needs denormal inputs honoured (nvcc default -ftz=false; never --use_fast_math).
#if defined(MT_HOSTSIM)
static inline float mt_bits_to_float(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
#define MT_B0(t) mt_bits_to_float(((t) & 0xffu) << 16)
#define MT_B1(t) mt_bits_to_float((((t) >> 8) & 0xffu) << 16)
#define MT_B2(t) mt_bits_to_float((t) & 0x00ff0000u)
#define MT_B3(t) mt_bits_to_float(((t) >> 24) << 16)
#else
#define MT_B0(t) __uint_as_float(__byte_perm((t), 0u, 0x4044))
#define MT_B1(t) __uint_as_float(__byte_perm((t), 0u, 0x4144))
#define MT_B2(t) __uint_as_float((t) & 0x00ff0000u)
#define MT_B3(t) __uint_as_float(__byte_perm((t), 0u, 0x4344))
#endif
#define MT_WSCALE 0x1p120f                          /* on the z weights (or y weights in 2D) */
#define MT_INV255 ((1.0f / 255.0f) * 8192.0f)       /* 2^13 / 255: undoes 2^-133 * 2^120      */

struct Rgba {
    float r, g, b, a;
};

// The eight filter weights (wx_i*wy_j)*wz_k as four pairs (i = 0,1 in the two halves): 6 packed multiplies.
struct Weights8 {
    P2 w00, w01, w10, w11;  // index = (k, j): pair over i
};
MT_DEVICE Weights8 filter_weights(const LinAxis& X, const LinAxis& Y, const LinAxis& Z)
{
    const P2 wx = pk2(X.w0, X.w1);
    const P2 xy0 = mul2(wx, bc2(Y.w0)), xy1 = mul2(wx, bc2(Y.w1));
    const float z0 = Z.w0 * MT_WSCALE, z1 = Z.w1 * MT_WSCALE;
    Weights8 w;
    w.w00 = mul2(xy0, bc2(z0)); w.w01 = mul2(xy1, bc2(z0));
    w.w10 = mul2(xy0, bc2(z1)); w.w11 = mul2(xy1, bc2(z1));
    return w;
}
#define MT_RG(t) pk2(MT_B0(t), MT_B1(t))
#define MT_BA(t) pk2(MT_B2(t), MT_B3(t))
// acc = w000*t000 (+fma) w001*t001 ... w111*t111 on a channel pair; texel order z-major, x fastest (as the oracle)
#define MT_ACC2(CH)                                                                                                    \
    fma2(bc2(hi2(w.w11)), CH(t111), fma2(bc2(lo2(w.w11)), CH(t110), fma2(bc2(hi2(w.w10)), CH(t101),                      \
    fma2(bc2(lo2(w.w10)), CH(t100), fma2(bc2(hi2(w.w01)), CH(t011), fma2(bc2(lo2(w.w01)), CH(t010),                      \
    fma2(bc2(hi2(w.w00)), CH(t001), mul2(bc2(lo2(w.w00)), CH(t000)))))))))

// the canonical four-channel filter of a cell whose two slices' quads are already in registers
MT_DEVICE Rgba tex3d_rgba_quads(const Quad& q0, const Quad& q1, const LinAxis& X, const LinAxis& Y, const LinAxis& Z)
{
    const uint32_t t000 = q0.x, t001 = q0.y, t010 = q0.z, t011 = q0.w, t100 = q1.x, t101 = q1.y, t110 = q1.z, t111 = q1.w;
    const Weights8 w = filter_weights(X, Y, Z);
    const P2 rg = mul2(MT_ACC2(MT_RG), bc2(MT_INV255));
    const P2 ba = mul2(MT_ACC2(MT_BA), bc2(MT_INV255));
    Rgba o;
    o.r = lo2(rg); o.g = hi2(rg); o.b = lo2(ba); o.a = hi2(ba);
    return o;
}

MT_DEVICE Rgba tex3d_rgba_axes(const Tex3D& T, const LinAxis& X, const LinAxis& Y, const LinAxis& Z, unsigned cell)
{
    // 32-bit unsigned texel offsets from one uniform base: no 64-bit pointer arithmetic per texel
    const unsigned W = (unsigned)T.w, H = (unsigned)T.h;
    uint32_t t000, t001, t010, t011, t100, t101, t110, t111;
#if MT_TEX_QUADS
    {
#if MT_TEX_BRICKS  // 3D: the cell's 2x2x2 texels are 32 contiguous bytes (8x the memory)
        const Quad* bp = T.quads + 2u * cell;
        const Quad q0 = MT_LDG_QUAD(bp);
        const Quad q1 = MT_LDG_QUAD(bp + 1);
#else
        // slice z1 = (z0 + 1) mod d: one slice further, wrapped by masking the cell index (w*h*d is a power of two)
        const Quad q0 = MT_LDG_QUAD(T.quads + cell);
        const Quad q1 = MT_LDG_QUAD(T.quads + ((cell + W * H) & (W * H * (unsigned)T.d - 1u)));
#endif
        t000 = q0.x; t001 = q0.y; t010 = q0.z; t011 = q0.w;
        t100 = q1.x; t101 = q1.y; t110 = q1.z; t111 = q1.w;
    }
#else
    {
        const unsigned r00 = (Z.i0 * H + Y.i0) * W, r01 = (Z.i0 * H + Y.i1) * W;
        const unsigned r10 = (Z.i1 * H + Y.i0) * W, r11 = (Z.i1 * H + Y.i1) * W;
        const uint32_t* __restrict__ tx = T.texels;
        t000 = MT_LDG(tx + (r00 + X.i0)); t001 = MT_LDG(tx + (r00 + X.i1));
        t010 = MT_LDG(tx + (r01 + X.i0)); t011 = MT_LDG(tx + (r01 + X.i1));
        t100 = MT_LDG(tx + (r10 + X.i0)); t101 = MT_LDG(tx + (r10 + X.i1));
        t110 = MT_LDG(tx + (r11 + X.i0)); t111 = MT_LDG(tx + (r11 + X.i1));
    }
#endif
    const Weights8 w = filter_weights(X, Y, Z);
    const P2 rg = mul2(MT_ACC2(MT_RG), bc2(MT_INV255));
    const P2 ba = mul2(MT_ACC2(MT_BA), bc2(MT_INV255));
    Rgba o;
    o.r = lo2(rg); o.g = hi2(rg); o.b = lo2(ba); o.a = hi2(ba);
    return o;
}

// ---- (r, F) form of the low-frequency volume, for the light-cone samples ---------------------------------------------
// A cone sample's density feeds radiance only (cloudRayMarch.comp:654-668): its VALUE may differ from the canonical one by
// rounding, its SIGN (density > 0 adds a term to the cone density, else nothing) may not.  fbm = .625 g + .25 b + .125 a is
// linear in the texel, so F = 5g + 2b + a (an integer 0..2040 = 8 * 255 * fbm) can be filtered as ONE channel beside r:
// 16 byte extractions and 8 packed multiply-adds per fetch instead of 32 and 16.  The filtered r is bit-identical to the
// canonical filter's (same chain); the filtered F differs from the canonical combination of three separately rounded
// channels by at most ~1.2e-6 (12 + 9 roundings of 2^-24), which moves the remapped base density by < 1.6e-6.  Callers
// compare that base density with the coverage threshold and fall back to the canonical four-channel filter inside a guard
// band (MT_RF_GUARD), so the sign decision is always the canonical one.
// Word layout: r in bits 24..31 (MT_B3 -> r * 2^-133), F in bits 13..23 (mask -> F * 2^-136): with the 2^120 of the z weights
// both sums come out scaled by 2^-13 resp. 2^-16, and F / 8 makes the final constant the same 2^13 / 255 for both halves.
#define MT_RF_GUARD 8e-6f
MT_DEVICE uint32_t rf_pack(uint32_t t)
{
    const uint32_t r = t & 0xffu, g = (t >> 8) & 0xffu, b = (t >> 16) & 0xffu, a = t >> 24;
    return (r << 24) | ((5u * g + 2u * b + a) << 13);
}
#if defined(MT_HOSTSIM)
#define MT_F13(t) mt_bits_to_float((t) & 0x00ffe000u)
#else
#define MT_F13(t) __uint_as_float((t) & 0x00ffe000u)
#endif
#define MT_RFP(t) pk2(MT_B3(t), MT_F13(t))
// MT_CONE_LERP: the (r, F) pair is radiance only and its decisions are guarded (cloud_core.cuh, cone_term_rf), so it is filtered
// as seven nested lerps a + f (b - a) on the raw denormal-scaled words: no weight products, 15 packed operations instead of 19
// and a dependent chain of three lerps instead of eight multiply-adds.  Differences and lerps are multiples of 2^-149 = 2^-16
// of one byte step, so each lerp is within 6e-8 of full scale of the exact one: < 2e-7 in (r, fbm) (measured 1.6e-7 against
// the real-number filter, tests/test_exact_tricks.py), inside what MT_RF_GUARD allows.  Every (r, F) fetch of every kernel goes
// through this one function, so the step-parallel and the sequential march stay bit-identical to each other.
#ifndef MT_CONE_LERP
#define MT_CONE_LERP 1
#endif
MT_DEVICE P2 rf_filter(uint32_t t000, uint32_t t001, uint32_t t010, uint32_t t011, uint32_t t100, uint32_t t101, uint32_t t110,
                       uint32_t t111, const LinAxis& X, const LinAxis& Y, const LinAxis& Z)
{
#if MT_CONE_LERP
    const P2 fx = bc2(X.w1), fy = bc2(Y.w1), fz = bc2(Z.w1);
    const P2 p000 = MT_RFP(t000), p010 = MT_RFP(t010), p100 = MT_RFP(t100), p110 = MT_RFP(t110);
    const P2 l00 = fma2(fx, sub2(MT_RFP(t001), p000), p000), l01 = fma2(fx, sub2(MT_RFP(t011), p010), p010);
    const P2 l10 = fma2(fx, sub2(MT_RFP(t101), p100), p100), l11 = fma2(fx, sub2(MT_RFP(t111), p110), p110);
    const P2 m0 = fma2(fy, sub2(l01, l00), l00), m1 = fma2(fy, sub2(l11, l10), l10);
    return mul2(fma2(fz, sub2(m1, m0), m0), bc2((float)(0x1p133 / 255.0)));  // undoes the 2^-133 of the byte placement (2^133 is no binary32)
#else
    const Weights8 w = filter_weights(X, Y, Z);
    return mul2(MT_ACC2(MT_RFP), bc2(MT_INV255));
#endif
}
// returns (r, fbm) of the filtered sample, both already divided by 255 (and F by 8)
MT_DEVICE P2 tex3d_rf_axes(const Tex3D& T, const LinAxis& X, const LinAxis& Y, const LinAxis& Z, unsigned cell)
{
    const unsigned W = (unsigned)T.w, H = (unsigned)T.h;
    uint32_t t000, t001, t010, t011, t100, t101, t110, t111;
#if MT_TEX_QUADS && MT_RF_BRICKS
    {
        (void)W; (void)H;
        asm("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
            : "=r"(t000), "=r"(t001), "=r"(t010), "=r"(t011), "=r"(t100), "=r"(t101), "=r"(t110), "=r"(t111)
            : "l"(T.rfquads + 2u * cell));
    }
#elif MT_TEX_QUADS
    {
        const Quad q0 = MT_LDG_QUAD(T.rfquads + cell);
        const Quad q1 = MT_LDG_QUAD(T.rfquads + ((cell + W * H) & (W * H * (unsigned)T.d - 1u)));
        t000 = q0.x; t001 = q0.y; t010 = q0.z; t011 = q0.w;
        t100 = q1.x; t101 = q1.y; t110 = q1.z; t111 = q1.w;
    }
#else
    {
        (void)cell;
        const unsigned r00 = (Z.i0 * H + Y.i0) * W, r01 = (Z.i0 * H + Y.i1) * W;
        const unsigned r10 = (Z.i1 * H + Y.i0) * W, r11 = (Z.i1 * H + Y.i1) * W;
        const uint32_t* __restrict__ tx = T.texels;
        t000 = rf_pack(MT_LDG(tx + (r00 + X.i0))); t001 = rf_pack(MT_LDG(tx + (r00 + X.i1)));
        t010 = rf_pack(MT_LDG(tx + (r01 + X.i0))); t011 = rf_pack(MT_LDG(tx + (r01 + X.i1)));
        t100 = rf_pack(MT_LDG(tx + (r10 + X.i0))); t101 = rf_pack(MT_LDG(tx + (r10 + X.i1)));
        t110 = rf_pack(MT_LDG(tx + (r11 + X.i0))); t111 = rf_pack(MT_LDG(tx + (r11 + X.i1)));
    }
#endif
    return rf_filter(t000, t001, t010, t011, t100, t101, t110, t111, X, Y, Z);
}

MT_DEVICE Rgba tex3d_rgba(const Tex3D& T, float s, float t, float r)
{
    LinAxis X = lin_axis_repeat(s, T.w), Y = lin_axis_repeat(t, T.h), Z = lin_axis_repeat(r, T.d);
    return tex3d_rgba_axes(T, X, Y, Z, tex_cell(T, X.i0, Y.i0, Z.i0));
}

// ---- empty-cell map of the low-frequency volume ----------------------------------------------------------------------
// A sample has cloud density > 0 iff remapClamped(r, fbm-.9, 1, 0, 1) > coverage (cloudRayMarch.comp:513-537), which
// in exact arithmetic is the LINEAR condition  L := r - (1-cov)*fbm > 1.9*cov - 0.9  on the filtered channels.  The
// filtered L is a convex combination of the eight corner texels' L_i, so if every corner has L_i <= threshold the
// sample's density is exactly 0 and the fetch + filter can be skipped without changing the result.  The test keeps
// a margin of 1e-4 (fp32 filter / remap error is < 2e-6), so a skipped sample is provably one the full evaluation
// would have returned +0 for.  One bit per cell, rebuilt whenever the texture or the coverage changes.
#define MT_OCC_MARGIN 1e-4f
MT_DEVICE bool occ_texel_may_be_cloud(uint32_t texel, float coverage)
{
    float r = (float)(texel & 0xffu), g = (float)((texel >> 8) & 0xffu), b = (float)((texel >> 16) & 0xffu), a = (float)(texel >> 24);
    float fbm = (g * 0.625f + b * 0.25f + a * 0.125f) * (1.0f / 255.0f);
    float L = r * (1.0f / 255.0f) - (1.0f - coverage) * fbm;
    return L > (1.9f * coverage - 0.9f) - MT_OCC_MARGIN;
}
// cell = (z0*h + y0)*w + x0: the bitmap is x fastest, 32 cells per word, and w is a multiple of 32 -- so the word is
// cell >> 5 and the bit cell & 31 (the shift instruction takes the low five bits by itself).  The same cell index
// addresses the quad copy, so the test costs a shift, a load and a bit test on top of the fetch's own addressing.
MT_DEVICE bool occ_cell_may_be_cloud(const Tex3D& T, unsigned cell)
{
    const uint32_t word = MT_LDG(T.occ + (cell >> 5));
    return (word >> (cell & 31u)) & 1u;
}

// Same filter, only the first three channels (the high-frequency volume's alpha is never read).
template <bool MAGIC = false>
MT_DEVICE Rgba tex3d_rgb(const Tex3D& T, float s, float t, float r)
{
    LinAxis X, Y, Z = lin_axis_repeat<MAGIC>(r, T.d);
    lin_axes_xy<MAGIC>(pk2(s, t), T.w, T.h, X, Y);
    const unsigned W = (unsigned)T.w, H = (unsigned)T.h;
    uint32_t t000, t001, t010, t011, t100, t101, t110, t111;
#if MT_TEX_QUADS
    {
#if MT_TEX_BRICKS  // 3D: the cell's 2x2x2 texels are 32 contiguous bytes (8x the memory)
        const Quad* bp = T.quads + 2u * ((Z.i0 * H + Y.i0) * W + X.i0);
        const Quad q0 = MT_LDG_QUAD(bp);
        const Quad q1 = MT_LDG_QUAD(bp + 1);
#else
        const unsigned cell = (Z.i0 * H + Y.i0) * W + X.i0;
        const Quad q0 = MT_LDG_QUAD(T.quads + cell);
        const Quad q1 = MT_LDG_QUAD(T.quads + ((cell + W * H) & (W * H * (unsigned)T.d - 1u)));
#endif
        t000 = q0.x; t001 = q0.y; t010 = q0.z; t011 = q0.w;
        t100 = q1.x; t101 = q1.y; t110 = q1.z; t111 = q1.w;
    }
#else
    {
        const unsigned r00 = (Z.i0 * H + Y.i0) * W, r01 = (Z.i0 * H + Y.i1) * W;
        const unsigned r10 = (Z.i1 * H + Y.i0) * W, r11 = (Z.i1 * H + Y.i1) * W;
        const uint32_t* __restrict__ tx = T.texels;
        t000 = MT_LDG(tx + (r00 + X.i0)); t001 = MT_LDG(tx + (r00 + X.i1));
        t010 = MT_LDG(tx + (r01 + X.i0)); t011 = MT_LDG(tx + (r01 + X.i1));
        t100 = MT_LDG(tx + (r10 + X.i0)); t101 = MT_LDG(tx + (r10 + X.i1));
        t110 = MT_LDG(tx + (r11 + X.i0)); t111 = MT_LDG(tx + (r11 + X.i1));
    }
#endif
    const Weights8 w = filter_weights(X, Y, Z);
    const P2 rg = mul2(MT_ACC2(MT_RG), bc2(MT_INV255));
    const float b = fmaf(hi2(w.w11), MT_B2(t111), fmaf(lo2(w.w11), MT_B2(t110), fmaf(hi2(w.w10), MT_B2(t101),
                    fmaf(lo2(w.w10), MT_B2(t100), fmaf(hi2(w.w01), MT_B2(t011), fmaf(lo2(w.w01), MT_B2(t010),
                    fmaf(hi2(w.w00), MT_B2(t001), lo2(w.w00) * MT_B2(t000)))))))) * MT_INV255;
    Rgba o;
    o.r = lo2(rg); o.g = hi2(rg); o.b = b; o.a = 0.0f;
    return o;
}

// Bilinear, first two channels (curl noise: only .xy is read, cloudRayMarch.comp:545-546).
template <bool MAGIC = false>
MT_DEVICE void tex2d_rg(const Tex2D& T, float s, float t, float& r, float& g)
{
    LinAxis X, Y;
    lin_axes_xy<MAGIC>(pk2(s, t), T.w, T.h, X, Y);
    const unsigned W = (unsigned)T.w;
    uint32_t t00, t01, t10, t11;
#if MT_TEX_QUADS
    {
        const Quad q = MT_LDG_QUAD(T.quads + (Y.i0 * W + X.i0));
        t00 = q.x; t01 = q.y; t10 = q.z; t11 = q.w;
    }
#else
    {
        const uint32_t* __restrict__ tx = T.texels;
        t00 = MT_LDG(tx + (Y.i0 * W + X.i0)); t01 = MT_LDG(tx + (Y.i0 * W + X.i1));
        t10 = MT_LDG(tx + (Y.i1 * W + X.i0)); t11 = MT_LDG(tx + (Y.i1 * W + X.i1));
    }
#endif
    // 2D: the oracle's weight is wx*wy, so the 2^120 rides on the y weight
    const P2 wx = pk2(X.w0, X.w1);
    const P2 w0 = mul2(wx, bc2(Y.w0 * MT_WSCALE)), w1 = mul2(wx, bc2(Y.w1 * MT_WSCALE));  // (w00, w01), (w10, w11)
    const P2 rg = mul2(fma2(bc2(hi2(w1)), MT_RG(t11), fma2(bc2(lo2(w1)), MT_RG(t10), fma2(bc2(hi2(w0)), MT_RG(t01),
                       mul2(bc2(lo2(w0)), MT_RG(t00))))), bc2(MT_INV255));
    r = lo2(rg);
    g = hi2(rg);
}
