// warp_stats.cpp -- MEASUREMENT TOOL (host), never part of the product.  Replays the full-quality Cloud pass warp by warp
// (16x2 ray tiles, lock-step march iterations, exactly the launch shape of cloud_raymarch_kernel<FULL>) with the kernels'
// own device functions compiled for the host, and counts how many lanes are busy in each phase of a march iteration:
// base sample, in-cloud part (erosion + lighting), the six cone samples, and of those the ones that survive the
// empty-cell test.  The numbers decide between lane-redistribution schemes before GPU time is spent (profiles/r2_*).
#define MT_HOSTSIM 1
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>
using std::max;
using std::min;
#include "../../meteoros_b200/csrc/cloud_core.cuh"
#include "../../meteoros_b200/csrc/mt_host_consts.h"

struct Stats {
    unsigned long long warps, warps_marching, warp_iters, lane_iters, warp_iters_hit, lane_hits;
    unsigned long long cone_warp_iters[6], cone_lanes[6], cone_nonempty_warp[6], cone_nonempty_lanes[6], cone_hit_lanes[6];
    unsigned long long hist_hits[33];       // warp iterations by number of in-cloud lanes
    unsigned long long early_lane_idle;     // lane-iterations spent idle after an early exit while the warp still marches
    unsigned long long tail_lane_idle;      // ... idle because the ray ran out of steps / was culled
    unsigned long long nonempty_total_hist[193]; // per warp iteration with hits: number of non-empty cone samples (0..192)
};

extern "C" int ws_run(const MtCameraUBO* cam, const MtTimeUBO* tm, const MtTuning* tun, const uint8_t* low, int lw, int lh, int ld,
                      const uint8_t* high, int hw, int hh, int hd, const uint8_t* curl, int cw, int ch, int W, int H,
                      int tile_w, int tile_h, int row_stride, Stats* out)
{
    CloudParams P;
    memset(&P, 0, sizeof(P));
    memcpy(&P.cam, cam, sizeof(CamU));
    memcpy(&P.tm, tm, sizeof(TimeU));
    P.tun = *tun;
    mt_host_sky_const(*cam, *tun, P.sky);
    P.low.texels = (const uint32_t*)low; P.low.w = lw; P.low.h = lh; P.low.d = ld;
    P.high.texels = (const uint32_t*)high; P.high.w = hw; P.high.h = hh; P.high.d = hd;
    P.curl.texels = (const uint32_t*)curl; P.curl.w = cw; P.curl.h = ch;
    std::vector<uint32_t> occ;
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
    P.W = W; P.H = H;
    P.tx = (((W / 4) + 31) / 32) * 32;
    P.ty = (((H / 4) + 31) / 32) * 32;
    P.full = 1;
    MarchConst M;
    cloud_frame_setup(P.cam, P.tm, P.tun, M);
    cloud_frame_jitter(P.tm, W, H, M.tabs);
    const MarchTabs& J = M.tabs;
    Stats S;
    memset(&S, 0, sizeof(S));
    const int lanes = tile_w * tile_h;  // 32
#pragma omp parallel
    {
        Stats L;
        memset(&L, 0, sizeof(L));
#pragma omp for schedule(dynamic, 1)
        for (int ty0 = 0; ty0 < H; ty0 += tile_h * row_stride) {
            for (int tx0 = 0; tx0 < W; tx0 += tile_w) {
                RaySetup R[32];
                float t[32], accum[32], tr[32], col[32];
                bool live[32], early[32];
                int nlive = 0;
                for (int l = 0; l < lanes; ++l) {
                    int px = tx0 + l % tile_w, py = ty0 + l / tile_w;
                    live[l] = false; early[l] = false;
                    if (px >= W || py >= H) continue;
                    F4 hdr;
                    int id = ((px & 3) << 2) | (py & 3);
                    R[l] = cloud_ray_setup(P, M, J, px, py, id, hdr);
                    if (R[l].branch != 2) continue;
                    t[l] = R[l].t_in; accum[l] = 0; tr[l] = 1; col[l] = 0;
                    live[l] = t[l] < R[l].t_out;
                    nlive += live[l];
                }
                L.warps++;
                if (!nlive) continue;
                L.warps_marching++;
                int iters = 0;
                for (;;) {
                    int a = 0;
                    for (int l = 0; l < lanes; ++l) a += live[l];
                    if (!a || iters >= MT_MAX_MARCH_ITERS) break;
                    L.warp_iters++;
                    L.lane_iters += a;
                    for (int l = 0; l < lanes; ++l) if (!live[l]) { if (early[l]) L.early_lane_idle++; else L.tail_lane_idle++; }
                    int hits = 0, cone_l[6] = {0,0,0,0,0,0}, cone_ne[6] = {0,0,0,0,0,0}, cone_h[6] = {0,0,0,0,0,0};
                    for (int l = 0; l < lanes; ++l) {
                        if (!live[l]) continue;
                        int px = tx0 + l % tile_w, py = ty0 + l / tile_w;
                        int id = ((px & 3) << 2) | (py & 3);
                        const int jidx = (id + mt_f2i(t[l])) & 15;
                        RayCounters none = {0,0,0,0,0,0};
                        StepBase B = cloud_step_base<false, false, false>(P, M, J.stepJitter[jidx >> 1], R[l], t[l], none, ConeOffsets{ nullptr, 0 });
                        StepSample smp; smp.inc = 0; smp.energy = -1;
                        if (B.baseDensity > 0.0f) {
                            hits++;
                            // replicate cloud_step_light with per-sample accounting
                            const f3 ec = M.earthCenter, pos = B.pos;
                            const f3 relOrigin = mk3(ec.x, MT_R_INNER - MT_EARTH_RADIUS, ec.z);
                            for (int i = 0; i < 6; ++i) {
                                const float fi = (float)i;
                                const f3 cs = M.coneStep[i];
                                float lx = (pos.x + (cs.x * R[l].stepSize) * fi) - relOrigin.x;
                                float ly = (pos.y + (cs.y * R[l].stepSize) * fi) - relOrigin.y;
                                float lz = (pos.z + (cs.z * R[l].stepSize) * fi) - relOrigin.z;
                                float sx = div_thickness(lx), sy = div_thickness(ly), sz = div_thickness(lz);
                                LinAxis X = lin_axis_repeat(sx, lw), Y = lin_axis_repeat(sy, lh), Z = lin_axis_repeat(sz, ld);
                                cone_l[i]++;
                                if (occ_cell_may_be_cloud(P.low, tex_cell(P.low, X.i0, Y.i0, Z.i0))) {
                                    cone_ne[i]++;
                                    float cur = low_freq_density<false, false>(P, M, R[l].covRcp, P.tun.coverage, pk2(sx, sy), sz, sx, sz, B.h);
                                    if (cur > 0.0f) cone_h[i]++;
                                }
                            }
                            { const ConeOffsets noCache = { nullptr, 0 }; smp = cloud_step_light<false, false, false>(P, M, R[l], B, none, noCache); }
                        }
                        if (cloud_step_combine(smp, accum[l], tr[l], col[l])) { live[l] = false; early[l] = true; }
                        else {
                            t[l] += R[l].stepSize;
                            if (!(t[l] < R[l].t_out)) live[l] = false;
                        }
                    }
                    L.hist_hits[hits]++;
                    if (hits) {
                        L.warp_iters_hit++;
                        L.lane_hits += hits;
                        int ne = 0;
                        for (int i = 0; i < 6; ++i) {
                            L.cone_warp_iters[i]++;
                            L.cone_lanes[i] += cone_l[i];
                            if (cone_ne[i]) L.cone_nonempty_warp[i]++;
                            L.cone_nonempty_lanes[i] += cone_ne[i];
                            L.cone_hit_lanes[i] += cone_h[i];
                            ne += cone_ne[i];
                        }
                        L.nonempty_total_hist[ne]++;
                    }
                    ++iters;
                }
            }
        }
#pragma omp critical
        {
            unsigned long long* d = (unsigned long long*)&S;
            const unsigned long long* s = (const unsigned long long*)&L;
            for (size_t k = 0; k < sizeof(Stats) / 8; ++k) d[k] += s[k];
        }
    }
    *out = S;
    return 0;
}
