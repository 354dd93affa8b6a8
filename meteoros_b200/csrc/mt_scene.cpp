// mt_scene.cpp -- the uniform producers of the reference application as plain host functions:
// Camera (camera.cpp), Scene's Time block (Scene.cpp:65-138) and Sky's SunAndSky block (Sky.cpp:64-74).  The vector
// arithmetic follows glm 0.9.9.0 in binary32 (normalize = v * inversesqrt(dot), lookAtRH, perspectiveRH_ZO, rotate),
// compiled with -ffp-contract=off so the UBO bytes do not depend on the host's FMA support.
#include <cmath>
#include <cstring>

#include "../../include/meteoros_b200.h"

namespace {

struct V3 {
    float x, y, z;
};
inline V3 v3(const float* p) { return V3{ p[0], p[1], p[2] }; }
inline void put(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
inline V3 operator+(V3 a, V3 b) { return V3{ a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator-(V3 a, V3 b) { return V3{ a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator*(V3 a, float s) { return V3{ a.x * s, a.y * s, a.z * s }; }
inline float dot(V3 a, V3 b)  // glm::dot: tmp = a * b; tmp.x + tmp.y + tmp.z
{
    float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return (tx + ty) + tz;
}
inline V3 normalize(V3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline V3 cross(V3 a, V3 b) { return V3{ a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y }; }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

void recompute(MtxCamera* c)  // Camera::RecomputeAttributes
{
    V3 f = normalize(v3(c->ref) - v3(c->eye));
    V3 r = normalize(cross(f, V3{ 0.0f, 1.0f, 0.0f }));
    V3 u = cross(r, f);
    put(c->forward, f);
    put(c->right, r);
    put(c->up, u);
    c->aspect = (float)c->width / (float)c->height;
}

void rotate_about(MtxCamera* c, float deg, V3 axis_in)  // glm::rotate(mat4(1), radians(deg), axis) applied to ref - eye
{
    const float a = radians(deg);
    const float cs = std::cos(a), sn = std::sin(a);
    const V3 ax = normalize(axis_in);
    const V3 t = ax * (1.0f - cs);
    float rot[3][3];  // [col][row]
    rot[0][0] = cs + t.x * ax.x; rot[0][1] = t.x * ax.y + sn * ax.z; rot[0][2] = t.x * ax.z - sn * ax.y;
    rot[1][0] = t.y * ax.x - sn * ax.z; rot[1][1] = cs + t.y * ax.y; rot[1][2] = t.y * ax.z + sn * ax.x;
    rot[2][0] = t.z * ax.x + sn * ax.y; rot[2][1] = t.z * ax.y - sn * ax.x; rot[2][2] = cs + t.z * ax.z;
    const V3 r = v3(c->ref) - v3(c->eye);
    float o[3];
    for (int row = 0; row < 3; ++row) o[row] = (rot[0][row] * r.x + rot[1][row] * r.y) + rot[2][row] * r.z;
    put(c->ref, V3{ o[0], o[1], o[2] } + v3(c->eye));
    recompute(c);
}

void translate(MtxCamera* c, V3 d)
{
    put(c->eye, v3(c->eye) + d);
    put(c->ref, v3(c->ref) + d);
    recompute(c);
}

float halton_at(int index, int base)  // Scene::HaltonSequenceAt, Scene.cpp:125-138
{
    float f = 1.0f, r = 0.0f;
    while (index > 0) {
        f = f / (float)base;
        r += f * (float)(index % base);
        index = (int)std::floor((double)(index / base));
    }
    return r;
}

}  // namespace

extern "C" {

void mtxCameraInit(MtxCamera* cam, int32_t width, int32_t height, const float eye[3], const float ref[3], float fovy_deg,
                   float near_clip, float far_clip)
{
    if (!cam) return;
    std::memset(cam, 0, sizeof(*cam));
    cam->width = width;
    cam->height = height;
    const float de[3] = { 0.0f, 0.0f, 2.0f }, dr[3] = { 0.0f, 0.0f, 1.0f };  // main.cpp:157-158
    std::memcpy(cam->eye, eye ? eye : de, sizeof(de));
    std::memcpy(cam->ref, ref ? ref : dr, sizeof(dr));
    cam->fovy_deg = fovy_deg;
    cam->near_clip = near_clip;
    cam->far_clip = far_clip;
    recompute(cam);
}
void mtxCameraRotateAboutUp(MtxCamera* cam, float deg) { if (cam) rotate_about(cam, deg, v3(cam->up)); }
void mtxCameraRotateAboutRight(MtxCamera* cam, float deg) { if (cam) rotate_about(cam, deg, v3(cam->right)); }
void mtxCameraTranslateAlongLook(MtxCamera* cam, float amt) { if (cam) translate(cam, v3(cam->forward) * amt); }
void mtxCameraTranslateAlongRight(MtxCamera* cam, float amt) { if (cam) translate(cam, v3(cam->right) * amt); }
void mtxCameraTranslateAlongUp(MtxCamera* cam, float amt) { if (cam) translate(cam, v3(cam->up) * amt); }

void mtxCameraUBO(const MtxCamera* cam, MtCameraUBO* out)
{
    if (!cam || !out) return;
    std::memset(out, 0, sizeof(*out));
    // glm::lookAtRH(eye, ref, up), column-major m[c*4 + r]
    const V3 eye = v3(cam->eye);
    const V3 f = normalize(v3(cam->ref) - eye);
    const V3 s = normalize(cross(f, v3(cam->up)));
    const V3 u = cross(s, f);
    float* v = out->view;
    v[0] = s.x; v[4] = s.y; v[8] = s.z;
    v[1] = u.x; v[5] = u.y; v[9] = u.z;
    v[2] = -f.x; v[6] = -f.y; v[10] = -f.z;
    v[12] = -dot(s, eye); v[13] = -dot(u, eye); v[14] = dot(f, eye);
    v[15] = 1.0f;
    // glm::perspectiveRH_ZO(radians(fovy), width / (float)height, near, far); proj[1][1] *= -1
    const float aspect = (float)cam->width / (float)cam->height;
    const float tanHalf = std::tan(radians(cam->fovy_deg) / 2.0f);
    float* p = out->proj;
    p[0] = 1.0f / (aspect * tanHalf);
    p[5] = (1.0f / tanHalf) * -1.0f;
    p[10] = cam->far_clip / (cam->near_clip - cam->far_clip);
    p[11] = -1.0f;
    p[14] = -(cam->far_clip * cam->near_clip) / (cam->far_clip - cam->near_clip);
    out->eye[0] = eye.x; out->eye[1] = eye.y; out->eye[2] = eye.z; out->eye[3] = 1.0f;
    // std::abs(std::tan(fovy*0.5 * (PI / 180.0))) in double with PI = 3.14159 (camera.h:10, camera.cpp:40)
    out->tanFovBy2[1] = (float)std::fabs(std::tan((double)cam->fovy_deg * 0.5 * (3.14159 / 180.0)));
    out->tanFovBy2[0] = cam->aspect * out->tanFovBy2[1];
}

void mtxTimeInit(MtTimeUBO* t)
{
    if (!t) return;
    std::memset(t, 0, sizeof(*t));
    float h[16];
    for (int i = 0; i < 16; ++i) h[i] = halton_at(i + 1, 3);  // base 3 for all sixteen (Scene.cpp:95-114)
    std::memcpy(t->haltonSeq1, h + 0, 16);
    std::memcpy(t->haltonSeq2, h + 4, 16);
    std::memcpy(t->haltonSeq3, h + 8, 16);
    std::memcpy(t->haltonSeq4, h + 12, 16);
    t->frameCountMod16 = 0;
}
void mtxTimeUpdate(MtTimeUBO* t, float dt)
{
    if (!t) return;
    t->time[0] = dt;
    t->time[1] += dt;
    t->frameCountMod16 = (t->frameCountMod16 + 1) % 16;
}
void mtxSunAndSky(MtSunAndSkyUBO* s)
{
    if (!s) return;
    const float loc[4] = { 0.0f, 1.0f, 0.0f, 0.0f }, dir[4] = { 1.0f, 1.0f, 1.0f, 0.0f }, col[4] = { 1.0f, 1.0f, 0.57f, 1.0f };
    std::memcpy(s->sunLocation, loc, 16);
    std::memcpy(s->sunDirection, dir, 16);
    std::memcpy(s->lightColor, col, 16);
    s->sunIntensity = 5.0f;
}

MtStatus mtxRunFrame(MtContext* ctx, const MtxCamera* cam, MtCameraUBO* camera_old, MtTimeUBO* time, float dt, uint32_t passes)
{
    if (!ctx || !cam || !camera_old || !time) return MT_ERR_INVALID;
    MtCameraUBO cur;
    MtSunAndSkyUBO sky;
    mtxCameraUBO(cam, &cur);
    mtxTimeUpdate(time, dt);   // scene->UpdateTime()
    mtxSunAndSky(&sky);        // sky->UpdateSunAndSky()
    MtStatus st;
    if ((st = mtSetCamera(ctx, &cur)) != MT_OK) return st;
    if ((st = mtSetCameraOld(ctx, camera_old)) != MT_OK) return st;
    if ((st = mtSetTime(ctx, time)) != MT_OK) return st;
    if ((st = mtSetSunAndSky(ctx, &sky)) != MT_OK) return st;
    if ((st = mtFrameEx(ctx, passes)) != MT_OK) return st;  // renderer->Frame()
    *camera_old = cur;                                       // cameraOld->UpdateBuffer(camera)
    return MT_OK;
}

}  // extern "C"
