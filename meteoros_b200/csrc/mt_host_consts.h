// mt_host_consts.h -- ray-independent part of the Preetham sky (cloudRayMarch.comp:406-434, 452-456), evaluated
// once per dispatch on the host in fp32.  Pure radiance (never a branch condition), so host libm is fine.
#pragma once

#include <math.h>

#include "mt_params.h"

static inline void mt_host_sky_const(const MtCameraUBO& cam, const MtTuning& tun, SkyConst& S)
{
    const float E_F = 2.718281828459f;
    const float sun[3] = { tun.sky_sun_location[0], tun.sky_sun_location[1], tun.sky_sun_location[2] };
    // sunDir = normalize(BACKGROUND_SKY_SUN_LOCATION - ray.origin), ray.origin = -camera.eye
    float d[3] = { sun[0] - (-cam.eye[0]), sun[1] - (-cam.eye[1]), sun[2] - (-cam.eye[2]) };
    float r = 1.0f / sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    for (int i = 0; i < 3; ++i) S.sunDir[i] = d[i] * r;
    // calcSunIntensity()
    float rs = 1.0f / sqrtf((sun[0] * sun[0] + sun[1] * sun[1]) + sun[2] * sun[2]);
    float zc = fminf(fmaxf(sun[1] * rs, -1.0f), 1.0f);
    float sunI = 1000.0f * fmaxf(0.0f, 1.0f - powf(E_F, -((1.6110731557f - acosf(zc)) / 1.5f)));
    S.sunE = 0.780f * sunI;
    // calcSkyBetaR() / calcSkyBetaV()
    float sunFade = 1.0f - fminf(fmaxf(1.0f - expf(sun[1] / 450000.0f), 0.0f), 1.0f);
    const float rayleighTotal[3] = { 5.804542996261093E-6f, 1.3562911419845635E-5f, 3.0265902468824876E-5f };
    const float mieConst[3] = { 1.839991851443397f, 2.779802391966052f, 4.079047954386109f };
    float c = (0.2f * 10.0f) * 10E-18f;
    for (int i = 0; i < 3; ++i) {
        S.betaR[i] = rayleighTotal[i] * ((2.0f - 1.0f) + sunFade);
        S.betaM[i] = (mieConst[i] * (0.434f * c)) * 0.005f;
        S.invBeta[i] = 0.0f;
    }
    float yDot = 1.0f - S.sunDir[1];
    yDot *= ((yDot * yDot) * yDot) * yDot;
    S.yDotMix = fminf(fmaxf(yDot, 0.0f), 1.0f);
}
