// cloud_core.cuh -- one ray of the Cloud compute pass (cloudRayMarch.comp:690-826), restructured for the GPU:
//   * per-frame quantities (camera basis, light direction, cone kernel, wind drift, sky constants) are hoisted
//     out of the ray (MarchConst / SkyConst);
//   * the erosion term of the six light-cone samples re-uses the curl + high-frequency fetch of the march sample
//     (cloudRayMarch.comp:665 passes the march sample's coordinates), so an in-cloud step costs 1 + 1 + 1 + 6
//     filtered fetches instead of 1 + 2 + 6*(1+2);
//   * radiance is grey, so one scalar is carried instead of a vec3.
// None of this changes a rounding: the decision-carrying values (t sequence, jitter index, densities, accumulated
// density) are bit-identical to oracle/meteoros_oracle.c.
#pragma once

#include "mt_params.h"

// MT_CONE_RF: light-cone samples from the (r, F) form of the low-frequency volume (cone_density_rf)
#ifndef MT_CONE_RF
#define MT_CONE_RF 1
#endif

#ifndef MT_SPEC_QUADS_FULL
#define MT_SPEC_QUADS_FULL 0  /* A/B: the full-quality kernel also requests a march sample's quads together with its bitmap word */
#endif
struct RayCounters {
    unsigned rays, marched, steps, incloud, cone, early;
};

// cloud_frame_setup (host, per dispatch; hostsim): the per-frame MarchConst (cloudRayMarch.comp:199-207, 585-624, 489-497).
MT_HD void cloud_frame_setup(const CamU& cam, const TimeU& tm, const MtTuning& tun, MarchConst& m)
{
    RayBasis b = ray_basis(cam);
    m.basisRight = b.right;
    m.basisUp = b.up;
    m.basisLook = b.look;
    m.eyePos = mk3(-cam.eye[0], -cam.eye[1], -cam.eye[2]);
    m.earthCenter = mk3(m.eyePos.x, -MT_EARTH_RADIUS, m.eyePos.z);
    f3 sun = mk3(tun.sun_location[0], tun.sun_location[1], tun.sun_location[2]);
    f3 l = norm3(sun - m.eyePos);
    m.lightDir = l;
    float a0 = fabsf(l.x), a1 = fabsf(l.y), a2 = fabsf(l.z);
    f3 mc;
    if (a0 > a1 && a0 > a2) mc = mk3(a0, 0.0f, 0.0f);
    else if (a1 > a0 && a1 > a2) mc = mk3(0.0f, a1, 0.0f);
    else mc = mk3(0.0f, 0.0f, a2);
    f3 zc = cross3(l, mc);
    f3 xc = cross3(zc, l);
    const float K[6][3] = { { 0.1f, 0.25f, -0.15f }, { 0.2f, 0.5f, 0.2f },  { -0.2f, 0.1f, -0.1f },
                            { -0.05f, 0.75f, 0.05f }, { -0.1f, 1.0f, 0.0f }, { 0.0f, 3.0f, 0.0f } };
    for (int i = 0; i < 6; ++i) m.coneStep[i] = (xc * K[i][0] + l * K[i][1]) + zc * K[i][2];
    f3 wind = mk3(tun.wind_direction[0], tun.wind_direction[1], tun.wind_direction[2]);
    m.windSkew = ((wind + mk3(0.0f, 0.1f, 0.0f)) * tun.cloud_speed) * tm.time[1];
    m.covDen = 1.0f - tun.coverage;
    m.covScale = tun.coverage / m.covDen;  // coverage in [0, 0.91]
}

// The two Halton look-ups of the shader (getJitterOffset, cloudRayMarch.comp:114-132) have only eight distinct
// results per frame each; tabulate them so that the march loop does no division and no divergent constant fetch.
MT_HD void cloud_frame_jitter(const TimeU& tm, int W, int H, MarchTabs& m)
{
    for (int hj = 0; hj < 8; ++hj) {
        int hx = hj < 4 ? hj : hj + 4;  // haltonSeq1/2 for index < 4, haltonSeq3/4 otherwise
        float x = tm.halton[hx], y = tm.halton[hx + 4];
        m.rayJitter[hj][0] = x / (float)W;
        m.rayJitter[hj][1] = y / (float)H;
        float sx = x / 75.0f, sy = y / 75.0f;
        m.stepJitter[hj][0] = sx;
        m.stepJitter[hj][1] = (sx + sy) * 1.180f;
        m.stepJitter[hj][2] = sy;
        m.stepJitter[hj][3] = 0.0f;
    }
}

MT_DEVICE float hg_phase(float cosa, float g)
{
    float num = 1.0f - g * g;
    float den = MT_POWF((1.0f + g * g) - (2.0f * g) * cosa, 1.5f);
    return (num / den) * 0.07957747154594767f;
}

// getAtmosphereColorPhysical (cloudRayMarch.comp:427-467) with the ray-independent part in SkyConst.
MT_DEVICE f3 sky_color(const SkyConst& S, f3 dir)
{
    const float PI_F = 3.14159265f;
    // cos(acos(x)) == x up to 1 ulp: the shader's cos(zenith) is evaluated as x itself (radiance-only term; this
    // also keeps cosf's Payne-Hanek slow path, and its local-memory scratch, out of the kernel)
    const float cosZenith = fmaxf(0.0f, dir.y);
    float zenith = acosf(cosZenith);
    float inverse = 1.0f / (cosZenith + 0.15f * MT_POWF(93.885f - ((zenith * 180.0f) / PI_F), -1.253f));
    float sR = 8.4E3f * inverse;
    float sM = 1.25E3f * inverse;
    float fex[3], col[3];
    f3 sunDir = mk3(S.sunDir[0], S.sunDir[1], S.sunDir[2]);
    float cosTheta = dot3(sunDir, dir);
    float rc = cosTheta * 0.5f + 0.5f;
    float rPhase = 0.05968310365946075f * (1.0f + rc * rc);
    float mPhase = hg_phase(cosTheta, 0.8f);
    const float SUN_ANGULAR_COS = 0.999956676946448443553574619906976478926848692873900859324f;
    float sunDisk = smoothstep1(SUN_ANGULAR_COS, SUN_ANGULAR_COS + 0.00002f, cosTheta);
    const float add[3] = { 0.0f, 0.0003f, 0.00075f };
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        fex[c] = MT_EXPF((-S.betaR[c]) * sR + S.betaM[c] * sM);
        float betas = (S.betaR[c] * rPhase + S.betaM[c] * mPhase) / (S.betaR[c] + S.betaM[c]);
        float lin = MT_POWF((S.sunE * betas) * (1.0f - fex[c]), 1.5f);
        lin *= mix1(1.0f, MT_POWF((S.sunE * betas) * fex[c], 0.5f), S.yDotMix);
        float l0 = 0.1f * fex[c];
        l0 += ((S.sunE * 15000.0f) * fex[c]) * sunDisk;
        col[c] = (lin + l0) * 0.04f + add[c];
    }
    return mk3(col[0], col[1], col[2]);
}

// getDensityHeightGradientForPoint (cloudRayMarch.comp:475-487); only the weather path calls it.
MT_DEVICE float height_gradient(float h, float cloudType)
{
    h = sat1(h);
    const float stratocumulus = fmaxf(0.0f, remap1(h, 0.0f, 0.25f, 0.0f, 1.0f) * remap1(h, 0.3f, 0.65f, 1.0f, 0.0f));
    const float stratus = fmaxf(0.0f, remap1(h, 0.0f, 0.1f, 0.0f, 1.0f) * remap1(h, 0.2f, 0.3f, 1.0f, 0.0f));
    const float a = mix1(stratus, stratocumulus, sat1(cloudType * 2.0f));
    const float b = mix1(stratocumulus, stratus, sat1((cloudType - 0.5f) * 2.0f));
    return mix1(a, b, cloudType);
}

// sampleLowFrequency (cloudRayMarch.comp:499-540): base cloud density with coverage applied.
// WEATHER = the shader's commented-out block :515-525 restored (MtTuning.use_weather): the weather map is sampled at
// unskewedSamplePoint.xz, the base cloud is scaled by the height gradient of its cloud type and its red channel
// replaces the constant coverage -- so neither the empty-cell bitmap (built for one coverage) nor the "nice" division
// (1 - coverage may be 0) applies on that path.
// STD != 0: the textures have the reference's extents (low 128^3, high 32^3, curl 128^2: Sky.cpp:31-50), which the host checks
// per dispatch: the extents become immediates (no constant-bank loads, shifts instead of multiplies in the addressing).
// STD == 2 additionally selects the software-pipelined, unrolled light-cone loop (the one-thread-per-ray kernels).
// the STD kernels floor their filter coordinates with the magic constant (mt_tex.cuh); the host launches them only inside its range
#define MT_STD_MAGIC(STD) ((STD) != 0 && MT_MAGIC_FLOOR != 0)
template <int STD>
MT_DEVICE Tex3D std_low(const Tex3D& t)
{
    Tex3D r = t;
    if (STD) r.w = r.h = r.d = 128;
    return r;
}
template <int STD>
MT_DEVICE Tex3D std_high(const Tex3D& t)
{
    Tex3D r = t;
    if (STD) r.w = r.h = r.d = 32;
    return r;
}
template <int STD>
MT_DEVICE Tex2D std_curl(const Tex2D& t)
{
    Tex2D r = t;
    if (STD) r.w = r.h = 128;
    return r;
}

template <bool WEATHER, int STD>
MT_DEVICE float low_freq_density(const CloudParams& P, const MarchConst& M, float covRcp, float coverage, P2 pxy, float pz, float ux, float uz, float relH)
{
    const Tex3D low = std_low<STD>(P.low);
    LinAxis X, Y, Z = lin_axis_repeat<MT_STD_MAGIC(STD)>(pz, low.d);
    lin_axes_xy<MT_STD_MAGIC(STD)>(pxy, low.w, low.h, X, Y);
    const unsigned cell = tex_cell(low, X.i0, Y.i0, Z.i0);
    // provably empty filter cell: the result is exactly +0.  A warp whose lanes all sit in empty cells skips the
    // whole fetch + filter (SIMT: the branch is free when nobody takes it).
#if MT_TEX_QUADS && !MT_TEX_BRICKS && !defined(MT_HOSTSIM)
    Rgba n;
    if ((STD == 3 || STD == 5 || (STD == 2 && MT_SPEC_QUADS_FULL)) && !WEATHER && low.occ) {
        // latency-bound callers (the step-parallel 1-of-16 kernel): the cell's quads are requested TOGETHER with its bitmap word
        // instead of after it -- one memory round trip per march sample instead of two, at the price of 32 unused bytes for an
        // empty cell
        const uint32_t word = MT_LDG(low.occ + (cell >> 5));
        const Quad q0 = MT_LDG_QUAD(low.quads + cell);
        const Quad q1 = MT_LDG_QUAD(low.quads + ((cell + 128u * 128u) & (128u * 128u * 128u - 1u)));
        if (!((word >> (cell & 31u)) & 1u)) return 0.0f;
        n = tex3d_rgba_quads(q0, q1, X, Y, Z);
    } else {
        if (!WEATHER && low.occ && !occ_cell_may_be_cloud(low, cell)) return 0.0f;
        n = tex3d_rgba_axes(low, X, Y, Z, cell);
    }
#else
    if (!WEATHER && low.occ && !occ_cell_may_be_cloud(low, cell)) return 0.0f;
    Rgba n = tex3d_rgba_axes(low, X, Y, Z, cell);
#endif
    float fbm = sat1((n.g * 0.625f + n.b * 0.25f) + n.a * 0.125f);
    float omin = fbm - 0.9f;
    float base = sat1(div_nice(n.r - omin, 1.0f - omin));  // remapClamped(r, fbm-.9, 1, 0, 1); denominator in [0.9, 1.9]
    if (WEATHER) {
        float wr, wg;
        tex2d_rg(P.weather, ux * P.tun.weather_scale, uz * P.tun.weather_scale, wr, wg);
        base *= height_gradient(relH, wg) * 0.5f;
        coverage = wr;
        const float v = clamp1(base, coverage, 1.0f);  // remapClampedBeforeAndAfter(base, cov, 1, 0, 1) * cov, IEEE division
        return sat1((v - coverage) / (1.0f - coverage)) * coverage;
    }
    // remapClampedBeforeAndAfter(base, cov, 1, 0, 1) * cov.  base <= cov clamps to cov and yields exactly +0.
    if (!(base > coverage)) return 0.0f;
    float b = sat1(div_nice_r(base - coverage, M.covDen, covRcp));  // coverage in [0, 0.91] (mtSetTuning); reciprocal prepared per ray
    return b * coverage;
}

// sampleLowFrequency for a LIGHT-CONE sample (cloudRayMarch.comp:658-665): the same function of the same sample point, but
// its result only feeds the cone density (radiance), so it is evaluated from the (r, F) form of the volume (mt_tex.cuh);
// whenever the remapped base density comes within MT_RF_GUARD of the coverage threshold the canonical four-channel
// evaluation decides instead.  Sign decisions (and with them MtCounters.cone_hits) are exactly the canonical ones; values
// differ by rounding (< 4e-6 relative in the density, < 1e-5 in a pixel against the 1e-3 bar).
// density of a cone sample from its filtered (r, fbm) pair; `exact` re-evaluates through the canonical filter inside the guard band
template <int STD>
MT_DEVICE float cone_density_rf(const CloudParams& P, const MarchConst& M, float covRcp, float coverage, P2 rf, const LinAxis& X, const LinAxis& Y,
                                const LinAxis& Z, unsigned cell)
{
    float fbm = sat1(hi2(rf));  // F sits three bits lower than r in its word: the common 2^13 / 255 already divides it by 8
    float omin = fbm - 0.9f;
    float base = sat1(div_nice(lo2(rf) - omin, 1.0f - omin));
    if (fabsf(base - coverage) <= MT_RF_GUARD) {  // too close to call from the one-channel fbm: canonical evaluation
        const Rgba n = tex3d_rgba_axes(std_low<STD>(P.low), X, Y, Z, cell);
        fbm = sat1((n.g * 0.625f + n.b * 0.25f) + n.a * 0.125f);
        omin = fbm - 0.9f;
        base = sat1(div_nice(n.r - omin, 1.0f - omin));
    }
    if (!(base > coverage)) return 0.0f;
    return sat1(div_nice_r(base - coverage, M.covDen, covRcp)) * coverage;
}
MT_DEVICE float erode(float base, float edge) { return div_nice(base - edge, 1.0f - edge); }  // remap(base, edge, 1, 0, 1); edge in [0, 0.005]
// The whole contribution of one light-cone sample to the cone density, erode(1.5 * density, edge) or 0 (cloudRayMarch.comp:
// 658-667), from its filtered (r, fbm) pair.  Radiance only, so outside the guard band nothing here needs the canonical
// roundings: the threshold test base > coverage is taken in its division-free form q > coverage * d (q = r - (fbm - .9),
// d = 1.9 - fbm > 0), the base density is q * rcp(d) with the SFU reciprocal, the coverage remap is one multiplication by
// coverage / (1 - coverage) (no clamp: base <= 1) and the erosion remap one FMA-style pair with the step's reciprocal of
// 1 - edge -- about 15 instructions instead of 29.  Inside the guard band (|q - coverage * d| <= MT_RF_GUARD * d, i.e. the base
// density within 8e-6 of the threshold, four times what the (r, F) filter and this arithmetic can be off by) the canonical
// four-channel evaluation decides AND supplies the value.  `hit` = the canonical `density > 0`.
template <int STD>
MT_DEVICE float cone_term_rf(const CloudParams& P, const MarchConst& M, float covRcp, float coverage, P2 rf, const LinAxis& X, const LinAxis& Y,
                             const LinAxis& Z, unsigned cell, float edge, float edgeRcp, bool& hit)
{
    const float fbm = sat1(hi2(rf));
    const float omin = fbm - 0.9f;
    const float d = 1.0f - omin, q = lo2(rf) - omin;
    const float diff = fmaf(-coverage, d, q);
    if (fabsf(diff) <= MT_RF_GUARD * d) {  // too close to call: the canonical evaluation, values included
        const float dens = cone_density_rf<STD>(P, M, covRcp, coverage, rf, X, Y, Z, cell);
        hit = dens > 0.0f;
        return hit ? erode(1.5f * dens, edge) : 0.0f;
    }
    hit = false;
    if (!(diff > 0.0f)) return 0.0f;
#if defined(MT_HOSTSIM)
    const float base = fminf(q / d, 1.0f);
#else
    float rd;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rd) : "f"(d));
    const float base = fminf(q * rd, 1.0f);
#endif
    const float dens = (base - coverage) * M.covScale;
    hit = dens > 0.0f;  // coverage == 0 scales every density to 0: no hit, as in the shader
    return hit ? fmaf(1.5f, dens, -edge) * edgeRcp : 0.0f;
}

// The part of erodeCloudWithHighFrequency (cloudRayMarch.comp:542-563) that depends only on the march sample:
// returns high_freq_modifier * 0.005, the lower edge of the final remap.
template <bool MAGIC>
MT_DEVICE float erosion_edge(const Tex2D& curl, const Tex3D& high, f3 p, float h)
{
    float cr, cg;
    tex2d_rg<MAGIC>(curl, p.x, p.y, cr, cg);
    float px = p.x + (cr * (1.0f - h)) * 0.5f;
    float py = p.y + (cg * (1.0f - h)) * 0.5f;
    Rgba n = tex3d_rgb<MAGIC>(high, px, py, p.z);
    float fbm = (n.r * 0.625f + n.g * 0.25f) + n.b * 0.125f;
    float m = sat1(mix1(fbm, 1.0f - fbm, sat1(h * 2.0f)));
    return m * 0.005f;
}

// GetLightEnergy (cloudRayMarch.comp:331-388), live branch only.
// Radiance only (no decision reads it): the three remaps divide by constants (1 - 0.7, 0.85 - 0.3, 0.34 - 0.07), written as
// multiplications by the reciprocal -- one instruction instead of the ~9 of an IEEE division, one ulp apart at most.
MT_DEVICE float remap_rcp(float v, float omin, float rcpRange, float nmin, float nmax) { return nmin + (((v - omin) * rcpRange) * (nmax - nmin)); }
// The ray's share of it: att * p * phase * 5 with att = max(remap(cosa, .7, 1, p, p * .25), p) = p * max(1 - .75 (cosa - .7) / .3, 1)
// (p > 0), so everything that depends on the ray alone -- the attenuation's angle factor, the phase function, the constant 5 -- is ONE
// per-ray factor (evaluated once per ray, parked beside the cone offsets) and a step's energy is scale * p^2 * depth * vert: eight
// instructions and one shared-memory load fewer per in-cloud step, the same real number (radiance only: equal to rounding).
MT_DEVICE float light_scale(float phase, float cosa)
{
    return (fmaxf(1.0f - 0.75f * ((cosa - 0.7f) * (1.0f / (1.0f - 0.7f))), 1.0f) * phase) * 5.0f;
}
MT_DEVICE float light_energy(float h, float dl, float ds, float scale)
{
    // p^2 = exp(-2 dl) in one ex2; the two remaps are affine in h: one multiply-add each (the same real numbers, radiance only)
    const float kd = (0.125f * (1.0f / (0.85f - 0.3f))) * (2.0f - 0.5f), cd = 0.5f - (0.3f * (1.0f / (0.85f - 0.3f))) * (2.0f - 0.5f);
    const float kv = (1.5f * (1.0f / (0.34f - 0.07f))) * (1.0f - 0.1f), cv = 0.1f - (0.07f * (1.0f / (0.34f - 0.07f))) * (1.0f - 0.1f);
    float pp = MT_EXP_NEG2(dl);
    float depth = 0.05f + MT_POWF(ds, clamp1(fmaf(h, kd, cd), 0.5f, 2.0f));
    float vert = MT_POWF(clamp1(fmaf(h, kv, cv), 0.1f, 1.0f), 0.8f);
    return (pp * (depth * vert)) * scale;
}

MT_DEVICE void encode_mask(float v, F4& o)
{
    float e0 = v, e1 = 255.0f * v, e2 = 65025.0f * v, e3 = 16581375.0f * v;
    e0 = e0 - floorf(e0); e1 = e1 - floorf(e1); e2 = e2 - floorf(e2); e3 = e3 - floorf(e3);
    const float k = 1.0f / 255.0f;
    o.x = e0 - e1 * k; o.y = e1 - e2 * k; o.z = e2 - e3 * k; o.w = e3 - e3 * 0.0f;
}


// Everything about one ray that does not change along the march (castRay, the horizon branches, the two shell
// intersections, the phase function).  64 bytes: also the record the step-parallel path keeps per ray.
struct RaySetup {
    f3 dir;
    float t_in, t_out, stepSize;
    float lenToInner, cosAngle, phase;  // phase: light_scale(HGModified, cosAngle), the per-ray factor of GetLightEnergy
    f3 bg;           // Preetham sky * max(.62, dir.y)
    int branch;      // 0 ocean, 1 sky band, 2 march
    int nsteps;      // step-parallel path only: iterations of the march loop (cloud_rays_kernel fills it in)
    float covRcp;    // refined reciprocal of 1 - coverage (nice_rcp: the device's own MUFU.RCP + Newton step, hence per ray, not per frame)
    int pad;
};

static_assert(sizeof(RaySetup) == 64, "RaySetup is the 64-byte per-ray record of the step-parallel path");

// One march step's contribution, independent of the steps before it.
struct StepSample {
    float inc;     // highFreqDensity * 0.5  (added to accumDensity)
    float energy;  // GetLightEnergy(...)    (mixed into the transmittance); < 0 marks "baseDensity <= 0"
};

// The ray direction of castRay for pixel (px, py) -- the first lines of cloud_ray_setup, for callers that need the
// direction (and from it the background) without the rest.
MT_DEVICE f3 cloud_ray_dir(const CloudParams& P, const MarchConst& M, const MarchTabs& J, int px, int py, int pixelID)
{
    float u = (float)px / (float)P.W;
    float v = 1.0f - (float)py / (float)P.H;
    const float jx = J.rayJitter[pixelID >> 1][0], jy = J.rayJitter[pixelID >> 1][1];
    RayBasis B;
    B.right = M.basisRight; B.up = M.basisUp; B.look = M.basisLook;
    return cast_ray_dir(P.cam, B, M.eyePos, u, v, jx, jy);
}
// background of a ray that is not ocean: Preetham sky * max(.62, dir.y) (cloudRayMarch.comp:731-733)
MT_DEVICE f3 cloud_ray_background(const CloudParams& P, f3 dir)
{
    const float dotUp = (0.0f * dir.x + 1.0f * dir.y) + 0.0f * dir.z;
    return sky_color(P.sky, dir) * fmaxf(0.620f, dotUp);
}

// castRay + branches (cloudRayMarch.comp:690-753).  For branch 0/1 `hdr` is final; mask is zero.
// WITH_BG = false leaves the background (R.bg, and hdr of a sky-band pixel) to the caller: the fused 1-of-16 kernel
// evaluates the Preetham sky in a second warp beside the geometry.
template <bool WITH_BG = true>
MT_DEVICE RaySetup cloud_ray_setup(const CloudParams& P, const MarchConst& M, const MarchTabs& J, int px, int py, int pixelID, F4& hdr)
{
    RaySetup R;
    float u = (float)px / (float)P.W;
    float v = 1.0f - (float)py / (float)P.H;
    const float jx = J.rayJitter[pixelID >> 1][0], jy = J.rayJitter[pixelID >> 1][1];
    RayBasis B;
    B.right = M.basisRight; B.up = M.basisUp; B.look = M.basisLook;
    const f3 origin = M.eyePos;
    const f3 dir = cast_ray_dir(P.cam, B, origin, u, v, jx, jy);
    R.dir = dir;
    R.t_in = R.t_out = R.stepSize = R.lenToInner = R.cosAngle = R.phase = 0.0f;
    R.bg = mk3(0.0f, 0.0f, 0.0f);
    R.nsteps = 0;
    R.covRcp = 0.0f;
    R.pad = 0;
    const float dotUp = (0.0f * dir.x + 1.0f * dir.y) + 0.0f * dir.z;
    hdr.w = 1.0f;
    if (dotUp < 0.0f) {  // ocean (:718-729)
        float a = -dir.y * 5.5f;
        hdr.x = mix1(0.0f * 0.4f, 0.0f * 0.5f, a);
        hdr.y = mix1(0.16f * 0.4f, 0.73f * 0.5f, a);
        hdr.z = mix1(0.51f * 0.4f, 0.95f * 0.5f, a);
        R.branch = 0;
        return R;
    }
    if (WITH_BG) R.bg = sky_color(P.sky, dir) * fmaxf(0.620f, dotUp);
    if (dotUp < 0.06f) {  // sky band below the cloud fade-out (:730-740)
        hdr.x = R.bg.x; hdr.y = R.bg.y; hdr.z = R.bg.z;
        R.branch = 1;
        return R;
    }
    R.branch = 2;
    R.covRcp = nice_rcp(M.covDen);
    // shells (:750-753) and the per-ray constants of rayMarch (:567-590)
    const f3 ec = M.earthCenter;
    ShellHit hin = ray_shell(origin, dir, ec, MT_R_INNER);
    ShellHit hout = ray_shell(origin, dir, ec, MT_R_OUTER);
    const float maxSteps = floorf(mix1(35.0f, 60.0f, 1.0f - dotUp));
    R.t_in = hin.t;
    R.t_out = hout.t;
    R.stepSize = (hout.t - hin.t) / maxSteps;
    R.cosAngle = dot3(norm3(dir), M.lightDir);
    // the ray's light-energy scale (light_scale): the phase function times the attenuation's angle factor times 5
    R.phase = light_scale(fmaxf(hg_phase(R.cosAngle, 0.6f), 0.7f * hg_phase(R.cosAngle, 0.99f - 0.1f)), R.cosAngle);
    R.lenToInner = len3(hin.point - origin);
    return R;
}

#ifndef MT_PACK_ITERS
#define MT_PACK_ITERS 1
#endif
#ifndef MT_PARK_BG
#if defined(MT_HOSTSIM)
#define MT_PARK_BG 0
#else
#define MT_PARK_BG 1
#endif
#endif
#ifndef MT_CONE_SKIP0
#define MT_CONE_SKIP0 0  /* 4K: 3.621 ms without, 3.633 ms with: not worth the special case */
#endif
struct ConeOffsets {
    const F4* xyz;    // this thread's first offset (x, y, z, -); the offset of sample i is xyz[i * stride]   (null: compute per step)
    int stride;
};

// The sample point of one march iteration and its base density (cloudRayMarch.comp:629-640): everything a step needs
// before it knows whether it is inside a cloud.
struct StepBase {
    f3 pos, skew;
    float h, baseDensity;
};

// MT_BASE_PACKED: the x, y components of the per-step vector arithmetic as fp32x2 pairs (one issue slot for two IEEE operations;
// the kernel is issue bound).  Every operation is the same IEEE operation on the same operands -- a product that feeds an
// addition is added with scalar FADDs, never mul2 -> add2 (mt_math.cuh) -- so the result is bit-identical to the scalar form.
#ifndef MT_BASE_PACKED
#define MT_BASE_PACKED 1
#endif
template <bool COUNT, bool WEATHER, int STD>
MT_DEVICE StepBase cloud_step_base(const CloudParams& P, const MarchConst& M, const float* sj, const RaySetup& R, float t, RayCounters& cnt,
                                    const ConeOffsets& CO)
{
    StepBase B;
    // parked per-ray values (cloud_ray): an LDS instead of a register the compiler would spill to local memory
    const float lenToInner = (MT_PARK_BG && CO.xyz) ? CO.xyz[5 * CO.stride].w : R.lenToInner;
    const f3 origin = M.eyePos, ec = M.earthCenter, dir = R.dir;
    const f3 relOrigin = mk3(ec.x, MT_R_INNER - MT_EARTH_RADIUS, ec.z);
    const f3 wind = mk3(P.tun.wind_direction[0], P.tun.wind_direction[1], P.tun.wind_direction[2]);
#if MT_BASE_PACKED
    // pos = origin + (dir + jitter) * t
    const P2 jxy = add2(pk2(dir.x, dir.y), pk2(sj[0], sj[1]));
    const float jz = dir.z + sj[2];
    const P2 mxy = mul2(jxy, bc2(t));
    const f3 pos = mk3(origin.x + lo2(mxy), origin.y + hi2(mxy), origin.z + jz * t);
    const P2 pxy = pk2(pos.x, pos.y);
    // sp = ((pos - relOrigin) / 12500) / 8
    const P2 spxy = mul2(div_thickness2(sub2(pxy, pk2(relOrigin.x, relOrigin.y))), bc2(0.125f));
    const float spz = div_thickness(pos.z - relOrigin.z) * 0.125f;
    // getRelativeHeightInAtmosphere (:171-186)
    const P2 cxy = sub2(pxy, pk2(origin.x, origin.y));                // pos - origin
    const float cz = pos.z - origin.z;
    const P2 c2 = mul2(cxy, cxy);
    const float lenFromCam = sqrt_nice((lo2(c2) + hi2(c2)) + cz * cz);  // 2e4 .. 3e5 m
    const P2 exy = sub2(pxy, pk2(ec.x, ec.y));                        // pos - earth centre, ~6.4e6 m
    const float ez = pos.z - ec.z;
    const P2 e2 = mul2(exy, exy);
    const float rinv = div_nice(1.0f, sqrt_nice((lo2(e2) + hi2(e2)) + ez * ez));
    const P2 nxy = mul2(mul2(exy, bc2(rinv)), pk2(dir.x, dir.y));     // dir * normalize(pos - ec), x and y
    const float cosTheta = (lo2(nxy) + hi2(nxy)) + dir.z * (ez * rinv);
    const float h = div_thickness(fabsf(cosTheta * (lenFromCam - lenToInner)));
    // skewSamplePointWithWind (:489-497): (sp + ((wind * h) * offset) * 0.009) + windSkew
    const P2 wxy = mul2(mul2(mul2(pk2(wind.x, wind.y), bc2(h)), bc2(P.tun.cloud_top_offset)), bc2(0.009f));
    const float wz = ((wind.z * h) * P.tun.cloud_top_offset) * 0.009f;
    const P2 skxy = add2(pk2(lo2(spxy) + lo2(wxy), hi2(spxy) + hi2(wxy)), pk2(M.windSkew.x, M.windSkew.y));
    const f3 skew = mk3(lo2(skxy), hi2(skxy), (spz + wz) + M.windSkew.z);
    B.pos = pos; B.skew = skew; B.h = h;
    B.baseDensity = low_freq_density<WEATHER, STD>(P, M, R.covRcp, P.tun.coverage, skxy, skew.z, pos.x, pos.z, h) * P.tun.base_density_factor;
#else
    f3 jdir = dir + mk3(sj[0], sj[1], sj[2]);
    f3 pos = origin + jdir * t;
    f3 rp = pos - relOrigin;
    f3 sp = mk3(div_thickness(rp.x) * 0.125f, div_thickness(rp.y) * 0.125f, div_thickness(rp.z) * 0.125f);  // /12500, /8
    // getRelativeHeightInAtmosphere (:171-186)
    float lenFromCam = len3_nice(pos - origin);          // 2e4 .. 3e5 m
    float cosTheta = dot3(dir, norm3_nice(pos - ec));    // ~6.4e6 m
    float h = div_thickness(fabsf(cosTheta * (lenFromCam - lenToInner)));
    // skewSamplePointWithWind (:489-497)
    f3 skew = (sp + ((wind * h) * P.tun.cloud_top_offset) * 0.009f) + M.windSkew;
    B.pos = pos; B.skew = skew; B.h = h;
    B.baseDensity = low_freq_density<WEATHER, STD>(P, M, R.covRcp, P.tun.coverage, pk2(skew.x, skew.y), skew.z, pos.x, pos.z, h) * P.tun.base_density_factor;
#endif
    if (COUNT) cnt.steps++;
    return B;
}

// The in-cloud part of a march iteration (cloudRayMarch.comp:642-672): erosion, the six light-cone samples, light
// energy.  Call only when B.baseDensity > 0.
// The six light-cone offsets of one ray, (stepSize * noise_kernel[i]) * i, are the same at every step: the one-thread-per-ray
// kernel computes them once per ray into shared memory ((x, y, z, -) per sample, [i][thread]: conflict-free) and each in-cloud
// step reads them back -- one 16-byte load instead of a float conversion and four multiplies per cone sample, the same values.

// MT_CONE_PIPE: software-pipelined light-cone loop (device only, one-thread-per-ray kernels).  A cone sample is a dependent
// chain: position -> cell -> bitmap word -> brick load (L2 latency: the rays of an SM walk 64 MB) -> filter -> two divisions.
// One thread per ray has no second sample in flight, so each of the six samples of a step pays both memory round trips in
// full (3.1 long-scoreboard stall cycles per issued instruction, profiles/r2_cloud_final.md).  Here sample i+1's brick is
// requested -- one unconditional 256-bit load into eight registers -- BEFORE sample i is filtered, and the empty-cell test
// moves behind the load: bit 0 of the brick's first word carries the cell's flag (occupancy_build_kernel writes it), so the
// bitmap load and its dependent round trip disappear.  Same arithmetic in the same order: bit-identical.
// (Round 2 first built this with cp.async into shared staging slots: 4.69 / 4.61 ms against 4.48 ms for the plain loop --
// LDGSTS + LDS per sample cost more LSU work than the hidden latency bought; profiles/r2_ab.md.)
#ifndef MT_CONE_PIPE
#define MT_CONE_PIPE 1
#endif

#ifndef MT_BRICK_EVICT_LAST
#define MT_BRICK_EVICT_LAST 1
#endif
#if !defined(MT_HOSTSIM)
struct Brick {
    uint32_t t[8];  // t000 t001 t010 t011 t100 t101 t110 t111
};
__device__ __forceinline__ Brick ldg_brick(const Quad* bricks, unsigned cell, unsigned slice, unsigned wrap)
{
    Brick b;
#if !MT_RF_BRICKS  // quad layout: the two slices' quads are two 16-byte loads (32 MB instead of 64 MB of (r, F) data)
    const Quad q0 = __ldg(bricks + cell), q1 = __ldg(bricks + ((cell + slice) & wrap));
    b.t[0] = q0.x; b.t[1] = q0.y; b.t[2] = q0.z; b.t[3] = q0.w;
    b.t[4] = q1.x; b.t[5] = q1.y; b.t[6] = q1.z; b.t[7] = q1.w;
    return b;
#endif
    (void)slice; (void)wrap;
#if MT_BRICK_EVICT_LAST   // the bricks are what every ray samples six times per in-cloud step: last to leave the L2
    asm volatile("ld.global.nc.L2::evict_last.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
#else
    asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
#endif
                 : "=r"(b.t[0]), "=r"(b.t[1]), "=r"(b.t[2]), "=r"(b.t[3]), "=r"(b.t[4]), "=r"(b.t[5]), "=r"(b.t[6]), "=r"(b.t[7])
                 : "l"(reinterpret_cast<const char*>(bricks) + (size_t)cell * 32u));  // one IMAD.WIDE (2u * cell would be a separate 32-bit add)
    return b;
}
#endif

MT_DEVICE void cone_offset(const MarchConst& M, float stepSize, int i, P2& xy, float& z)
{
    const float fi = (float)i;
    const f3 cs = M.coneStep[i];
    xy = mul2(mul2(pk2(cs.x, cs.y), bc2(stepSize)), bc2(fi));
    z = (cs.z * stepSize) * fi;
}

// position of cone sample i in texture space and its filter axes / cell
template <int STD>
MT_DEVICE void cone_axes(const CloudParams& P, const MarchConst& M, const ConeOffsets& CO, f3 pos, f3 relOrigin, float stepSize, int i,
                         P2& sxy, float& sz, LinAxis& X, LinAxis& Y, LinAxis& Z, unsigned& cell)
{
    // lightPos = pos + (stepSize * noise_kernel[i]) * i ; sample = (lightPos - relOrigin) / 12500  -- x,y as a pair
    P2 off;
    float offz;
    P2 lxy;
    float lz;
    if (MT_CONE_SKIP0 && CO.xyz && i == 0) {
        // sample 0's offset is (stepSize * kernel[0]) * 0 = +-0 and pos + (+-0) == pos: no load, no add (a -0 component of pos would
        // come out as +0 in the canonical form; it is subtracted from or divided into the same value either way)
        lxy = sub2(pk2(pos.x, pos.y), pk2(relOrigin.x, relOrigin.y));
        lz = pos.z - relOrigin.z;
    } else {
        if (CO.xyz) {  // ONE 16-byte shared load (the cache is 16-byte aligned: cloud_raymarch.cu)
#if defined(MT_HOSTSIM)
            const F4 o = CO.xyz[i * CO.stride];
#else
            const float4 o = *reinterpret_cast<const float4*>(CO.xyz + i * CO.stride);
#endif
            off = pk2(o.x, o.y);
            offz = o.z;
        } else cone_offset(M, stepSize, i, off, offz);
        // the offset products are separate values (memory or scalar adds): a mul2 feeding an add2 would be contracted into
        // an FFMA2 (mt_math.cuh)
        lxy = sub2(CO.xyz ? add2(pk2(pos.x, pos.y), off) : pk2(pos.x + lo2(off), pos.y + hi2(off)), pk2(relOrigin.x, relOrigin.y));
        lz = (pos.z + offz) - relOrigin.z;
    }
    sxy = div_thickness2(lxy);
    sz = div_thickness(lz);
    const Tex3D low = std_low<STD>(P.low);
    Z = lin_axis_repeat<MT_STD_MAGIC(STD)>(sz, low.d);
    lin_axes_xy<MT_STD_MAGIC(STD)>(sxy, low.w, low.h, X, Y);
    cell = tex_cell(low, X.i0, Y.i0, Z.i0);
}

template <bool COUNT, bool WEATHER, int STD>
MT_DEVICE StepSample cloud_step_light(const CloudParams& P, const MarchConst& M, const RaySetup& R, const StepBase& B, RayCounters& cnt,
                                      const ConeOffsets& CO)
{
    StepSample S;
    const f3 ec = M.earthCenter, pos = B.pos, skew = B.skew;
    const f3 relOrigin = mk3(ec.x, MT_R_INNER - MT_EARTH_RADIUS, ec.z);
    const float coverage = P.tun.coverage, h = B.h, baseDensity = B.baseDensity;
    if (COUNT) cnt.incloud++;
    float dl = 0.0f;
#if !defined(MT_HOSTSIM)
    if (STD == 4 || STD == 5) {  // 4: the one-thread-per-ray kernel, 5: the step-parallel 1-of-16 kernel (whose march sample is STD == 3's)
        // MT_FLAG_HW_CONE_FILTER (opt-in; mt_tex.cuh): the texture unit fetches AND filters the six light-cone samples -- one TEX
        // instruction in place of cell index, bitmap flag, 256-bit brick load, sixteen field extractions and seven packed lerps.  The
        // samples feed radiance only, and this mode makes no exactness claim for it: the position is scaled by the rounded reciprocal
        // of the layer thickness (one multiplication instead of the three-instruction exact division), the term is the guard-less
        // form of cone_term_rf.  Erosion (S.inc, which feeds the accumulated density) stays on the exact path above it.
        const float edge = erosion_edge<MT_STD_MAGIC(STD)>(std_curl<STD>(P.curl), std_high<STD>(P.high), skew, h);
        S.inc = erode(baseDensity * 1.4f, edge) * 0.5f;
        const float edgeRcp = nice_rcp(1.0f - edge);
        const cudaTextureObject_t tex = (cudaTextureObject_t)P.low.hwtex;
        const float invT = 1.0f / MT_THICKNESS;
#ifndef MT_HW_BATCH
#define MT_HW_BATCH 3  /* fetches requested before the first one is consumed (registers: 4 per fetch in flight) */
#endif
#pragma unroll
        for (int i0 = 0; i0 < 6; i0 += MT_HW_BATCH) {
            float4 n[MT_HW_BATCH];
#pragma unroll
            for (int j = 0; j < MT_HW_BATCH; ++j) {
                const int i = i0 + j;
                P2 off;
                float offz;
                if (CO.xyz) {
                    const float4 o = *reinterpret_cast<const float4*>(CO.xyz + i * CO.stride);
                    off = pk2(o.x, o.y);
                    offz = o.z;
                } else cone_offset(M, R.stepSize, i, off, offz);
                const P2 lxy = sub2(pk2(pos.x + lo2(off), pos.y + hi2(off)), pk2(relOrigin.x, relOrigin.y));
                const float lz = (pos.z + offz) - relOrigin.z;
                n[j] = tex3D<float4>(tex, lo2(lxy) * invT, hi2(lxy) * invT, lz * invT);
            }
#pragma unroll
            for (int j = 0; j < MT_HW_BATCH; ++j) {
                const float fbm = sat1(fmaf(n[j].y, 0.625f, fmaf(n[j].z, 0.25f, n[j].w * 0.125f)));
                const float omin = fbm - 0.9f;
                const float d = 1.0f - omin, q = n[j].x - omin;
                if (q - coverage * d > 0.0f) {
                    float rd;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rd) : "f"(d));
                    const float dens = (fminf(q * rd, 1.0f) - coverage) * M.covScale;
                    if (dens > 0.0f) dl += fmaf(1.5f, dens, -edge) * edgeRcp;
                }
            }
        }
        const bool parkedH = MT_PARK_BG && CO.xyz;
        S.energy = light_energy(h, dl, baseDensity, parkedH ? CO.xyz[3 * CO.stride].w : R.phase);
        return S;
    }
#endif
#if !defined(MT_HOSTSIM) && MT_CONE_PIPE
    if (STD == 2 && MT_CONE_RF && !WEATHER) {  // STD kernels are only launched with the (r, F) bricks present (mt_std_dims)
        const Tex3D low = std_low<STD>(P.low);
        P2 sxy;
        float sz;
        LinAxis X, Y, Z;
        unsigned cell;
        // sample 0's brick is requested before the erosion fetches in the source; ptxas schedules the load (and a prefetch in its
        // place) behind them all the same, so 3.7 % of the kernel's stall samples still sit on the first brick's flag test
        cone_axes<STD>(P, M, CO, pos, relOrigin, R.stepSize, 0, sxy, sz, X, Y, Z, cell);
        const unsigned wrap = (unsigned)low.w * (unsigned)low.h * (unsigned)low.d - 1u, slice = (unsigned)low.w * (unsigned)low.h;
        Brick cur = ldg_brick(low.rfquads, cell, slice, wrap);
        const float edge = erosion_edge<MT_STD_MAGIC(STD)>(std_curl<STD>(P.curl), std_high<STD>(P.high), skew, h);
        S.inc = erode(baseDensity * 1.4f, edge) * 0.5f;
        const float edgeRcp = nice_rcp(1.0f - edge);  // the step's erosion divisor, shared by the six cone samples (radiance only)
#ifndef MT_CONE_UNROLL
#define MT_CONE_UNROLL 6  /* 4K: 4.204 (1), 4.144 (2), 4.137 (3), 4.043 ms (6): no loop-carried register rotation, constant offsets */
#endif
#define MT_PRAGMA_(x) _Pragma(#x)
#define MT_UNROLL_(n) MT_PRAGMA_(unroll n)
        MT_UNROLL_(MT_CONE_UNROLL)
        for (int i = 0; i < 6; ++i) {
            LinAxis Xn = X, Yn = Y, Zn = Z;
            unsigned celln = cell;
            Brick nxt = cur;
            if (i < 5) {  // request sample i+1's brick before sample i is filtered
                cone_axes<STD>(P, M, CO, pos, relOrigin, R.stepSize, i + 1, sxy, sz, Xn, Yn, Zn, celln);
                nxt = ldg_brick(low.rfquads, celln, slice, wrap);
            }
            if (cur.t[0] & 1u) {  // the cell may hold cloud (flag written with the empty-cell bitmap); else the density is exactly +0
                const uint32_t t000 = cur.t[0], t001 = cur.t[1], t010 = cur.t[2], t011 = cur.t[3], t100 = cur.t[4], t101 = cur.t[5],
                               t110 = cur.t[6], t111 = cur.t[7];
#ifdef MT_EMU_WBITS  /* probe build (tools/probes/emu_weight_bits.py, profiles/r2c_ab.md): filter fractions rounded to MT_EMU_WBITS bits, as a texture unit would */
                LinAxis Xq = X, Yq = Y, Zq = Z;
                const float qs = (float)(1 << MT_EMU_WBITS);
                Xq.w1 = rintf(X.w1 * qs) / qs; Yq.w1 = rintf(Y.w1 * qs) / qs; Zq.w1 = rintf(Z.w1 * qs) / qs;
                const P2 rf = rf_filter(t000, t001, t010, t011, t100, t101, t110, t111, Xq, Yq, Zq);
#else
                const P2 rf = rf_filter(t000, t001, t010, t011, t100, t101, t110, t111, X, Y, Z);
#endif
                bool hit;
                const float term = cone_term_rf<STD>(P, M, R.covRcp, coverage, rf, X, Y, Z, cell, edge, edgeRcp, hit);
                if (COUNT && hit) cnt.cone++;
                dl += term;  // + 0 where the sample holds no cloud: the same sum
            }
            X = Xn; Y = Yn; Z = Zn; cell = celln; cur = nxt;
        }
        S.energy = light_energy(h, dl, baseDensity, (MT_PARK_BG && CO.xyz) ? CO.xyz[3 * CO.stride].w : R.phase);
        return S;
    }
#endif
    const float edge = erosion_edge<MT_STD_MAGIC(STD)>(std_curl<STD>(P.curl), std_high<STD>(P.high), skew, h);
    S.inc = erode(baseDensity * 1.4f, edge) * 0.5f;
    const float edgeRcp = nice_rcp(1.0f - edge);  // the step's erosion divisor, shared by the six cone samples (radiance only)
    {
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {  // one copy of the filter in the instruction stream (I-cache)
            P2 sxy;
            float sz;
            LinAxis X, Y, Z;
            unsigned cell;
            float cur;
            if (STD && MT_CONE_RF && !WEATHER) {  // the plain (r, F) loop: the same per-sample term as the pipelined one
                cone_axes<STD>(P, M, CO, pos, relOrigin, R.stepSize, i, sxy, sz, X, Y, Z, cell);
                const Tex3D low = std_low<STD>(P.low);
#if !defined(MT_HOSTSIM) && MT_CONE_PIPE && MT_RF_BRICKS
                if (STD == 3) {
                    // brick first: the cell's empty flag is bit 0 of the brick's first word (occupancy_build_kernel), so a sample
                    // is ONE memory round trip (256-bit load, then flag test and filter) instead of bitmap word -> brick.  The
                    // step-parallel 1-of-16 kernel is latency bound on exactly that chain (profiles/r2_passes_1080p.md: 12 % of
                    // its stall samples wait for the brick, 8 % for the bitmap word before it).
                    const Brick b = ldg_brick(low.rfquads, cell, 0u, 0u);
                    if (b.t[0] & 1u) {
                        bool hit;
                        dl += cone_term_rf<STD>(P, M, R.covRcp, coverage, rf_filter(b.t[0], b.t[1], b.t[2], b.t[3], b.t[4], b.t[5], b.t[6], b.t[7], X, Y, Z),
                                                X, Y, Z, cell, edge, edgeRcp, hit);
                        if (COUNT && hit) cnt.cone++;
                    }
                    continue;
                }
#endif
                if (!low.occ || occ_cell_may_be_cloud(low, cell)) {
                    bool hit;
                    dl += cone_term_rf<STD>(P, M, R.covRcp, coverage, tex3d_rf_axes(low, X, Y, Z, cell), X, Y, Z, cell, edge, edgeRcp, hit);
                    if (COUNT && hit) cnt.cone++;
                }
                continue;
            }
            if (STD && MT_CONE_RF && !WEATHER) {
                cone_axes<STD>(P, M, CO, pos, relOrigin, R.stepSize, i, sxy, sz, X, Y, Z, cell);
                const Tex3D low = std_low<STD>(P.low);
                cur = (low.occ && !occ_cell_may_be_cloud(low, cell)) ? 0.0f
                      : cone_density_rf<STD>(P, M, R.covRcp, coverage, tex3d_rf_axes(low, X, Y, Z, cell), X, Y, Z, cell);
            } else {
                cone_axes<STD>(P, M, CO, pos, relOrigin, R.stepSize, i, sxy, sz, X, Y, Z, cell);
                cur = low_freq_density<WEATHER, STD>(P, M, R.covRcp, coverage, sxy, sz, lo2(sxy), sz, h);
            }
            if (cur > 0.0f) {
                if (COUNT) cnt.cone++;
                dl += erode(1.5f * cur, edge);
            }
        }
    }
    const bool parked = MT_PARK_BG && CO.xyz;
    S.energy = light_energy(h, dl, baseDensity, parked ? CO.xyz[3 * CO.stride].w : R.phase);
    return S;
}

// One iteration of the march loop (cloudRayMarch.comp:629-678) at parameter t, without the running sums.
template <bool COUNT, bool WEATHER, int STD>
MT_DEVICE StepSample cloud_step_sample_sj(const CloudParams& P, const MarchConst& M, const float* sj, const RaySetup& R, float t,
                                          RayCounters& cnt, const ConeOffsets& CO)
{
    const StepBase B = cloud_step_base<COUNT, WEATHER, STD>(P, M, sj, R, t, cnt, CO);
    if (B.baseDensity > 0.0f) return cloud_step_light<COUNT, WEATHER, STD>(P, M, R, B, cnt, CO);
    StepSample S;
    S.inc = 0.0f;
    S.energy = -1.0f;
    return S;
}

// jidx = int(mod(float(pixelID + int(t)), 16.0)) selects the step's jitter (cloudRayMarch.comp:629-631): entry jidx / 2 of the table
template <bool COUNT, bool WEATHER, int STD>
MT_DEVICE StepSample cloud_step_sample(const CloudParams& P, const MarchConst& M, const MarchTabs& J, const RaySetup& R, int jidx, float t,
                                       RayCounters& cnt, const ConeOffsets& CO)
{
    return cloud_step_sample_sj<COUNT, WEATHER, STD>(P, M, J.stepJitter[jidx >> 1], R, t, cnt, CO);
}

// Running sums of the march (cloudRayMarch.comp:648, 674-684).  Returns true when the loop must stop.
MT_DEVICE bool cloud_step_combine(const StepSample& S, float& accum, float& transmittance, float& color)
{
    if (S.energy >= 0.0f) {
        accum += S.inc;
        // mix(transmittance, energy, 1 - accum) in its lerp form, a + t (b - a): radiance only (accum, which decides, is untouched)
        transmittance = fmaf(1.0f - accum, S.energy - transmittance, transmittance);
        color += transmittance;
    }
    if (accum >= 1.0f) {
        accum = 1.0f;
        return true;
    }
    return false;
}

// Composite + god-ray mask (cloudRayMarch.comp:759-779).
MT_DEVICE void cloud_composite(const RaySetup& R, float accum, float color, F4& hdr, F4& mask)
{
    float fade = smoothstep1(0.0f, 1.0f, fminf(1.0f, remap1(R.dir.y, 0.06f, 0.2f, 0.0f, 1.0f)));
    float a = accum * fade;
    hdr.x = mix1(R.bg.x, color, a);
    hdr.y = mix1(R.bg.y, color, a);
    hdr.z = mix1(R.bg.z, color, a);
    hdr.w = 1.0f;
    encode_mask(25.0f * fminf(0.05f, 1.0f - accum), mask);
    if (R.dir.y < 0.05f) {  // unreachable here (dir.y >= 0.06) but part of the shader (:775-779)
        float k = fmaxf(5.0f, R.dir.y);
        mask.x *= k; mask.y *= k; mask.z *= k; mask.w *= k;
    }
}

// One invocation of main(): setup, the sequential march, composite.
template <bool COUNT, bool DEBUG, bool WEATHER, int STD>
MT_DEVICE void cloud_ray(const CloudParams& P, const MarchConst& M, const MarchTabs& J, int px, int py, int pixelID, F4& hdr, F4& mask,
                         RayCounters& cnt, MtRayDebug* dbg, F4* coneXYZ, int coneStride)
{
    mask.x = mask.y = mask.z = mask.w = 0.0f;
    const RaySetup R = cloud_ray_setup(P, M, J, px, py, pixelID, hdr);
    if (COUNT) cnt.rays++;
    if (DEBUG) {
        dbg->dir[0] = R.dir.x; dbg->dir[1] = R.dir.y; dbg->dir[2] = R.dir.z;
        dbg->t_in = dbg->t_out = dbg->step_size = dbg->accum = 0.0f;
        dbg->branch = R.branch; dbg->steps = 0; dbg->jitter_hash = 0u;
    }
    if (R.branch != 2) return;

    ConeOffsets CO;
    CO.xyz = coneXYZ; CO.stride = coneStride;
    if (coneXYZ) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            P2 oxy;
            F4 o;
            cone_offset(M, R.stepSize, i, oxy, o.z);
            o.x = lo2(oxy); o.y = hi2(oxy); o.w = 0.0f;
#if MT_PARK_BG
            // the spare components carry what the ray needs once per step or only after the march (background colour, phase,
            // cos of the sun angle, distance to the inner shell): an LDS where the compiler would otherwise spill a register to
            // local memory -- the loop runs at the register limit (64)
            o.w = i == 0 ? R.bg.x : i == 1 ? R.bg.y : i == 2 ? R.bg.z : i == 3 ? R.phase : i == 4 ? R.cosAngle : R.lenToInner;
#endif
            coneXYZ[i * coneStride] = o;
        }
    }
    float accum = 0.0f, transmittance = 1.0f, color = 0.0f;
    unsigned jhash = 2166136261u;
    int iters = 0;
    if (COUNT) cnt.marched++;
    if (DEBUG) { dbg->t_in = R.t_in; dbg->t_out = R.t_out; dbg->step_size = R.stepSize; }
#if !defined(MT_HOSTSIM)
    if (MT_PACK_ITERS && !DEBUG) {
        // ONE register for the iteration count and the pixel id: st = iters << 8 | pixelID << 3.  The cap is st < 64 << 8, and the
        // step's jitter entry, byte offset ((pixelID + int(t)) & 15) / 2 * 16 = ((pixelID << 3) + (int(t) << 3)) & 0x70, is one
        // shift-add and one mask on st -- before, the pixel id was spilled and every step began with its reload from local memory.
        unsigned st = (unsigned)pixelID << 3;
        for (float t = R.t_in; t < R.t_out && st < ((unsigned)MT_MAX_MARCH_ITERS << 8); t += R.stepSize, st += 256u) {
            const unsigned off = (st + ((unsigned)mt_f2i(t) << 3)) & 0x70u;
            const float* sj = reinterpret_cast<const float*>(reinterpret_cast<const char*>(&J.stepJitter[0][0]) + off);
            const StepSample S = cloud_step_sample_sj<COUNT, WEATHER, STD>(P, M, sj, R, t, cnt, CO);
            if (cloud_step_combine(S, accum, transmittance, color)) {
                if (COUNT) cnt.early++;
                break;
            }
        }
    } else
#endif
    for (float t = R.t_in; t < R.t_out && iters < MT_MAX_MARCH_ITERS; t += R.stepSize, ++iters) {
        const int jidx = (pixelID + mt_f2i(t)) & 15;  // int(mod(float(pixelID + int(t)), 16.0)), argument >= 0
        if (DEBUG) jhash = (jhash ^ (unsigned)jidx) * 16777619u;
        const StepSample S = cloud_step_sample<COUNT, WEATHER, STD>(P, M, J, R, jidx, t, cnt, CO);
        if (cloud_step_combine(S, accum, transmittance, color)) {
            if (COUNT) cnt.early++;
            ++iters;
            break;
        }
    }
    if (DEBUG) { dbg->steps = iters; dbg->jitter_hash = jhash; dbg->accum = accum; }
#if MT_PARK_BG
    if (coneXYZ) {
        RaySetup Rc;
        Rc.dir = R.dir;
        Rc.bg = mk3(coneXYZ[0].w, coneXYZ[coneStride].w, coneXYZ[2 * coneStride].w);
        cloud_composite(Rc, accum, color, hdr, mask);
        return;
    }
#endif
    cloud_composite(R, accum, color, hdr, mask);
}

// ---------------------------------------------------------------------------------------------------------------------
// MT_WARP_QUEUE: the march of one warp's 32 rays with the in-cloud part of the steps handed round the warp.
//
// A march step's contribution (StepSample) does not depend on the steps before it; only the running sums do.  In the plain loop
// (cloud_ray) the in-cloud part of a step -- erosion, six light-cone samples, light energy: two thirds of the kernel's
// instructions -- runs with the 26 of 32 lanes whose ray happens to be inside a cloud at that step
// (profiles/r2b_cloud_final.md).  Here every lane first takes its ray's NEXT step as far as the base density; rays inside a cloud
// append (sample point, height, base density) to the warp's ring of pending items in shared memory and march on.  Whenever 32
// items are pending (or a ray has four outstanding, or nothing is left to march) lane j evaluates item j -- whoever's ray it
// belongs to, with that ray's cached cone offsets -- and writes (inc, energy) back into the item's slot; the owners then fold
// their results in step order.  The arithmetic of every (ray, step) is the plain loop's, so the image is bit-identical; steps
// taken past a ray's early exit (accumulated density >= 1) are computed and dropped.
// ---------------------------------------------------------------------------------------------------------------------
#if !defined(MT_HOSTSIM)
#ifndef MT_WARP_QUEUE
#define MT_WARP_QUEUE 0
#endif
#define MT_WQ_SLOTS 64
struct WarpQueue {
    float4 a[MT_WQ_SLOTS];    // pos.xyz, h          -- after evaluation: (inc, energy, -, -)
    float4 b[MT_WQ_SLOTS];    // skew.xyz, baseDensity
    unsigned owner[MT_WQ_SLOTS];
};

template <int STD>
__device__ __forceinline__ void cloud_ray_queued(const CloudParams& P, const MarchConst& M, const MarchTabs& J, int px, int py, int pixelID,
                                                 bool valid, F4& hdr, F4& mask, F4* coneWarp, int coneStride, WarpQueue& Q)
{
    const unsigned FULLM = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    RayCounters cnt = { 0u, 0u, 0u, 0u, 0u, 0u };
    mask.x = mask.y = mask.z = mask.w = 0.0f;
    RaySetup R;
    R.branch = 0;
    R.t_in = R.t_out = R.stepSize = 0.0f;
    R.dir = mk3(0.0f, 0.0f, 0.0f);
    if (valid) R = cloud_ray_setup(P, M, J, px, py, pixelID, hdr);
    const bool marching = valid && R.branch == 2;
    if (!__any_sync(FULLM, marching)) return;  // warp-uniform

    ConeOffsets CO;
    CO.xyz = coneWarp + lane;
    CO.stride = coneStride;
    if (marching) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            P2 oxy;
            F4 o;
            cone_offset(M, R.stepSize, i, oxy, o.z);
            o.x = lo2(oxy); o.y = hi2(oxy);
            o.w = i == 0 ? R.bg.x : i == 1 ? R.bg.y : i == 2 ? R.bg.z : i == 3 ? R.phase : i == 4 ? R.cosAngle : R.lenToInner;
            coneWarp[lane + i * coneStride] = o;
        }
    }
    __syncwarp();
    RaySetup Rt;  // what cloud_step_light reads of a ray beside its parked values: the reciprocal of 1 - coverage (one value per frame)
    Rt.covRcp = nice_rcp(M.covDen);
    Rt.stepSize = 0.0f; Rt.phase = 0.0f; Rt.cosAngle = 0.0f;

    float accum = 0.0f, transmittance = 1.0f, color = 0.0f;
    unsigned st = (unsigned)pixelID << 3;
    float t = R.t_in;
    bool active = marching && t < R.t_out;
    bool done = false;
    unsigned pend = 0u;            // up to four outstanding items of this ray, oldest in the low byte: slot + 1
    unsigned head = 0u, count = 0u;  // warp-uniform: the ring holds slots head .. head + count - 1
    for (;;) {
        bool hit = false;
        StepBase B;
        if (active) {
            const unsigned off = (st + ((unsigned)mt_f2i(t) << 3)) & 0x70u;
            const float* sj = reinterpret_cast<const float*>(reinterpret_cast<const char*>(&J.stepJitter[0][0]) + off);
            B = cloud_step_base<false, false, STD>(P, M, sj, R, t, cnt, CO);
            hit = B.baseDensity > 0.0f;
            t += R.stepSize;
            st += 256u;
            active = t < R.t_out && st < ((unsigned)MT_MAX_MARCH_ITERS << 8);
        }
        const unsigned hm = __ballot_sync(FULLM, hit);
        if (hit) {
            const unsigned slot = (head + count + __popc(hm & ((1u << lane) - 1u))) & (MT_WQ_SLOTS - 1u);
            Q.a[slot] = make_float4(B.pos.x, B.pos.y, B.pos.z, B.h);
            Q.b[slot] = make_float4(B.skew.x, B.skew.y, B.skew.z, B.baseDensity);
            Q.owner[slot] = lane;
            // append to the ray's outstanding list: first free byte
            const unsigned sh = pend == 0u ? 0u : pend < 0x100u ? 8u : pend < 0x10000u ? 16u : 24u;
            pend |= (slot + 1u) << sh;
        }
        count += __popc(hm);
        bool anyActive = __any_sync(FULLM, active);
        __syncwarp();
        while (count >= 32u || (count > 0u && (!anyActive || __any_sync(FULLM, (pend >> 24) != 0u)))) {
            const unsigned n = count < 32u ? count : 32u;
            if (lane < n) {
                const unsigned slot = (head + lane) & (MT_WQ_SLOTS - 1u);
                const float4 a = Q.a[slot], b = Q.b[slot];
                StepBase I;
                I.pos = mk3(a.x, a.y, a.z); I.h = a.w;
                I.skew = mk3(b.x, b.y, b.z); I.baseDensity = b.w;
                ConeOffsets CI;
                CI.xyz = coneWarp + Q.owner[slot];
                CI.stride = coneStride;
                const StepSample S = cloud_step_light<false, false, STD>(P, M, Rt, I, cnt, CI);
                *reinterpret_cast<float2*>(&Q.a[slot]) = make_float2(S.inc, S.energy);
            }
            __syncwarp();
            while ((pend & 0xffu) != 0u && ((((pend & 0xffu) - 1u) - head) & (MT_WQ_SLOTS - 1u)) < n) {
                const float2 r = *reinterpret_cast<const float2*>(&Q.a[(pend & 0xffu) - 1u]);
                pend >>= 8;
                if (!done) {
                    StepSample S;
                    S.inc = r.x; S.energy = r.y;
                    if (cloud_step_combine(S, accum, transmittance, color)) {
                        done = true;
                        active = false;
                    }
                }
            }
            __syncwarp();
            head = (head + n) & (MT_WQ_SLOTS - 1u);
            count -= n;
            anyActive = __any_sync(FULLM, active);
        }
        if (!anyActive && count == 0u) break;
    }
    if (marching) {
        RaySetup Rc;
        Rc.dir = R.dir;
        Rc.bg = mk3(CO.xyz[0].w, CO.xyz[coneStride].w, CO.xyz[2 * coneStride].w);
        cloud_composite(Rc, accum, color, hdr, mask);
    }
}
#endif
