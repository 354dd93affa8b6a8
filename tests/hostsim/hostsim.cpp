// hostsim.cpp -- TEST TOOL, never part of the product.  Compiles the per-pixel device functions of the CUDA
// kernels (meteoros_b200/csrc/*_core.cuh) with g++ (MT_HOSTSIM) so that the restructured arithmetic of the
// kernels -- hoisted frame constants, shared erosion fetch, scalar radiance, specialised filters -- can be checked
// bit-for-bit against the independent oracle on a machine without a GPU.  With -ffp-contract=off the host
// evaluates the same IEEE operations the device does under -fmad=false; only exp/pow/acos/cos differ on the GPU.
// libmeteoros_b200.so does not contain, link or load any of this.
#define MT_HOSTSIM 1
#include <algorithm>
#include <cstring>
#include <vector>
using std::max;
using std::min;

#include "../../meteoros_b200/csrc/cloud_core.cuh"
#include "../../meteoros_b200/csrc/mt_host_consts.h"
#include "../../meteoros_b200/csrc/post_core.cuh"

extern "C" {

int hs_cloud(const MtCameraUBO* cam, const MtTimeUBO* tm, const MtTuning* tun, const uint8_t* low, int lw, int lh, int ld,
             const uint8_t* high, int hw, int hh, int hd, const uint8_t* curl, int cw, int ch, const uint8_t* weather, int ww, int wh,
             int W, int H, int full,
             float* hdr, float* mask, unsigned long long* counters, MtRayDebug* debug)
{
    CloudParams P;
    memset(&P, 0, sizeof(P));
    memcpy(&P.cam, cam, sizeof(CamU));
    memcpy(&P.tm, tm, sizeof(TimeU));
    P.tun = *tun;
    mt_host_sky_const(*cam, *tun, P.sky);
    P.low.texels = (const uint32_t*)low; P.low.w = lw; P.low.h = lh; P.low.d = ld;
    P.low.rfquads = (const Quad*)P.low.texels;  // non-null = light-cone samples take the (r, F) path; the host build packs (r, F) per texel on the fly
    P.high.texels = (const uint32_t*)high; P.high.w = hw; P.high.h = hh; P.high.d = hd;
    P.curl.texels = (const uint32_t*)curl; P.curl.w = cw; P.curl.h = ch;
    P.weather.texels = (const uint32_t*)weather; P.weather.w = ww; P.weather.h = wh;
    // empty-cell bitmap, built exactly like occupancy_build_kernel does
    std::vector<uint32_t> occ;
    if (lw >= 32) {
        const unsigned wpr = (unsigned)lw >> 5;
        occ.assign((size_t)wpr * lh * ld, 0u);
        for (unsigned z = 0; z < (unsigned)ld; ++z)
            for (unsigned y = 0; y < (unsigned)lh; ++y)
                for (unsigned x = 0; x < (unsigned)lw; ++x) {
                    bool any = false;
                    for (unsigned dz = 0; dz < 2; ++dz)
                        for (unsigned dy = 0; dy < 2; ++dy)
                            for (unsigned dx = 0; dx < 2; ++dx) {
                                unsigned xx = (x + dx) & (lw - 1), yy = (y + dy) & (lh - 1), zz = (z + dz) & (ld - 1);
                                any = any || occ_texel_may_be_cloud(P.low.texels[(zz * lh + yy) * lw + xx], tun->coverage);
                            }
                    if (any) occ[(z * lh + y) * wpr + (x >> 5)] |= 1u << (x & 31u);
                }
        P.low.occ = occ.data();
    }
    P.W = W; P.H = H;
    P.tx = (((W / 4) + 31) / 32) * 32;
    P.ty = (((H / 4) + 31) / 32) * 32;
    P.full = full == 1;
    const bool std_dims = lw == 128 && lh == 128 && ld == 128 && hw == 32 && hh == 32 && hd == 32 && cw == 128 && ch == 128;
    MarchConst M;
    cloud_frame_setup(P.cam, P.tm, P.tun, M);
    cloud_frame_jitter(P.tm, W, H, M.tabs);
    const MarchTabs& J = M.tabs;
    RayCounters cnt = { 0, 0, 0, 0, 0, 0 };
    unsigned long long tot[6] = { 0, 0, 0, 0, 0, 0 };
    const int gw = full == 1 ? W : P.tx, gh = full == 1 ? H : P.ty;  // full: 0 = 1-of-16, 1 = all pixels, 2 = 1-of-16 step-parallel
    for (int gy = 0; gy < gh; ++gy)
        for (int gx = 0; gx < gw; ++gx) {
            int px, py, id;
            bool valid;
            if (full == 1) {
                px = gx; py = gy;
                id = ((px & 3) << 2) | (py & 3);
                valid = px < W && py < H && (px >> 2) < P.tx && (py >> 2) < P.ty;
            } else {
                id = P.tm.frameCountMod16;
                px = gx * 4 + (id >> 2);
                py = gy * 4 + (id & 3);
                valid = px < W && py < H;
            }
            if (!valid) continue;
            F4 h, m;
            size_t idx = (size_t)py * W + px;
            MtRayDebug scratch;
            F4 coneXYZ[6];  // the kernel's per-ray light-cone offset cache (shared memory there), stride 1 here
            memset(&cnt, 0, sizeof(cnt));
            if (full == 2) {
                // The step-parallel decomposition of the 1-of-16 dispatch (cloud_rays_kernel / cloud_steps_kernel /
                // cloud_fold_kernel), with the same device functions: the ray record and its filed t sequence, every
                // (ray, step) sample evaluated on its own -- last step first, to make the independence explicit -- and
                // the fold in step order.
                m.x = m.y = m.z = m.w = 0.0f;
                RaySetup R = cloud_ray_setup(P, M, J, px, py, id, h);
                if (R.branch == 2) {
                    float tk[MT_STEP_SLICES];
                    int n = 0;
                    for (float t = R.t_in; t < R.t_out && n < MT_STEP_SLICES; t += R.stepSize) tk[n++] = t;
                    R.nsteps = n;
                    StepSample S[MT_STEP_SLICES];
                    RayCounters none = { 0, 0, 0, 0, 0, 0 };
                    const ConeOffsets noCache = { nullptr, 0 };
                    for (int k = n - 1; k >= 0; --k) {
                        const int jidx = (P.tm.frameCountMod16 + mt_f2i(tk[k])) & 15;
                        S[k] = tun->use_weather ? cloud_step_sample<false, true, 0>(P, M, J, R, jidx, tk[k], none, noCache)
                                 : std_dims     ? cloud_step_sample<false, false, 1>(P, M, J, R, jidx, tk[k], none, noCache)   // as cloud_steps_kernel<false, true>
                                                : cloud_step_sample<false, false, 0>(P, M, J, R, jidx, tk[k], none, noCache);
                    }
                    float accum = 0.0f, transmittance = 1.0f, color = 0.0f;
                    for (int k = 0; k < R.nsteps; ++k)
                        if (cloud_step_combine(S[k], accum, transmittance, color)) break;
                    cloud_composite(R, accum, color, h, m);
                }
                memcpy(hdr + 4 * idx, &h, 16);
                memcpy(mask + 4 * idx, &m, 16);
                continue;
            }
            if (tun->use_weather) cloud_ray<true, true, true, 0>(P, M, J, px, py, id, h, m, cnt, debug ? debug + idx : &scratch, coneXYZ, 1);
            else if (std_dims)  // STD: extents as immediates, light-cone samples from the (r, F) form
                cloud_ray<true, true, false, 1>(P, M, J, px, py, id, h, m, cnt, debug ? debug + idx : &scratch, coneXYZ, 1);
            else cloud_ray<true, true, false, 0>(P, M, J, px, py, id, h, m, cnt, debug ? debug + idx : &scratch, nullptr, 0);
            tot[0] += cnt.rays; tot[1] += cnt.marched; tot[2] += cnt.steps; tot[3] += cnt.incloud; tot[4] += cnt.cone; tot[5] += cnt.early;
            memcpy(hdr + 4 * idx, &h, 16);
            memcpy(mask + 4 * idx, &m, 16);
        }
    if (counters) for (int k = 0; k < 6; ++k) counters[k] += tot[k];
    return 0;
}

// launch order of the row tiles (mt_params.h): j-th tile issued -> index among the owned tiles
int hs_tile_order(int heavy_first, int tile_count, int j)
{
    RowTiles r;
    memset(&r, 0, sizeof(r));
    r.heavy_first = heavy_first;
    r.tile_count = tile_count;
    return mt_tile_order(r, j);
}

int hs_reproject(const MtCameraUBO* cam, const MtCameraUBO* camOld, const MtTimeUBO* tm, int W, int H, const float* prev,
                 float* cur, int* taps_out)
{
    ReprojParams P;
    memset(&P, 0, sizeof(P));
    memcpy(&P.cam, cam, sizeof(CamU));
    memcpy(&P.camOld, camOld, sizeof(CamU));
    memcpy(&P.tm, tm, sizeof(TimeU));
    P.W = W; P.H = H;
    ReprojFrame F = reproject_frame(P);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int taps[10];
            reproject_taps(P, F, (float)x / (float)W, (float)y / (float)H, taps);
            P2 axy = pk2(0.0f, 0.0f), azw = pk2(0.0f, 0.0f);
            for (int i = 0; i < 10; ++i) {
                const float* p = prev + 4 * (size_t)taps[i];
                axy = add2(axy, pk2(p[0], p[1]));
                azw = add2(azw, pk2(p[2], p[3]));
                if (taps_out) taps_out[((size_t)y * W + x) * 10 + i] = taps[i];
            }
            axy = MT_DIV_CONST2(axy, 10.0f);
            azw = MT_DIV_CONST2(azw, 10.0f);
            float* o = cur + 4 * ((size_t)y * W + x);
            o[0] = lo2(axy); o[1] = hi2(axy); o[2] = lo2(azw); o[3] = hi2(azw);
        }
    return 0;
}

int hs_godrays(const MtCameraUBO* cam, const float* lightColor, int W, int H, const float* mask, float* hdr)
{
    GodRayParams P;
    memset(&P, 0, sizeof(P));
    memcpy(&P.cam, cam, sizeof(CamU));
    P.lightColor[0] = lightColor[0]; P.lightColor[1] = lightColor[1]; P.lightColor[2] = lightColor[2];
    P.mask = (const F4*)mask;
    P.W = W; P.H = H;
    std::vector<float> dec1((size_t)(W + 2) * (H + 2), MT_MASK_BORDER_DECODED);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) dec1[(size_t)(y + 1) * (W + 2) + (x + 1)] = mask_texel_decode(P.mask[(size_t)y * W + x]);
    std::vector<float2> dec((size_t)(W + 2) * (H + 2));   // pairs (d(x, y), d(x+1, y)), as mask_decode_kernel writes them
    for (int y = 0; y < H + 2; ++y)
        for (int x = 0; x < W + 2; ++x) {
            const size_t i = (size_t)y * (W + 2) + x;
            dec[i].x = dec1[i];
            dec[i].y = x + 1 < W + 2 ? dec1[i + 1] : MT_MASK_BORDER_DECODED;
        }
    P.decoded = dec.data();
    P.pitch = W + 2;
    GodRayFrame G = godray_frame(P.cam);
    if (G.blend < 0.0f) return 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            F4 g = godray_pixel(P, G, x, y);
            float* o = hdr + 4 * ((size_t)y * W + x);
            o[0] += g.x; o[1] += g.y; o[2] += g.z; o[3] += g.w;
        }
    return 0;
}

int hs_tonemap(const MtTimeUBO* tm, int W, int H, const float* hdr, uint32_t* ldr)
{
    ToneMapParams P;
    P.storage = 0;
    P.hdr = (const F4*)hdr;
    P.ldr = ldr;
    P.W = W; P.H = H;
    P.seed = mt_f2u(tm->time[1]);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            F4 in;
            memcpy(&in, hdr + 4 * ((size_t)y * W + x), 16);
            ldr[(size_t)y * W + x] = tonemap_pixel(P, in, x, y);
        }
    return 0;
}

int hs_txaa(const MtCameraUBO* cam, const MtCameraUBO* camOld, const MtTimeUBO* tm, int W, int H, const uint32_t* cur,
            const uint32_t* prev, uint32_t* out)
{
    TxaaParams P;
    memset(&P, 0, sizeof(P));
    memcpy(&P.cam, cam, sizeof(CamU));
    memcpy(&P.camOld, camOld, sizeof(CamU));
    memcpy(&P.tm, tm, sizeof(TimeU));
    P.cur = cur; P.prev = prev; P.out = out;
    P.W = W; P.H = H;
    TxaaFrame F = txaa_frame(P);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) out[(size_t)y * W + x] = txaa_pixel(P, F, x, y, ((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
    return 0;
}

}  // extern "C"
