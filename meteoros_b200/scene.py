"""Uniform producers: the host-side state Meteoros keeps in Camera / Scene / Sky.

Mirrors (all arithmetic in float32, same operation order as glm 0.9.9.0):
  Camera::UpdateBuffer / RecomputeAttributes / RotateAboutUp / RotateAboutRight   camera.cpp:31-42, 68-96
  glm::lookAtRH, glm::perspectiveRH_ZO, glm::rotate        external/glm/glm/gtc/matrix_transform.inl:754, 327
  Scene::InitializeTime / UpdateTime / HaltonSequenceAt    Scene.cpp:86-119, 65-85, 125-138
  Sky::UpdateSunAndSky                                     Sky.cpp:64-74
  frame loop order                                         main.cpp:172-194

The structs are numpy structured dtypes whose byte layout equals MtCameraUBO / MtTimeUBO / MtSunAndSkyUBO
(include/meteoros_b200.h) = the reference's std140 UBOs (camera.h:12-18, Scene.h:12-21, Sky.h:9-15).
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32

CAMERA_DTYPE = np.dtype(
    [("view", "<f4", (16,)), ("proj", "<f4", (16,)), ("eye", "<f4", (4,)), ("tanFovBy2", "<f4", (2,))]
)
TIME_DTYPE = np.dtype(
    [
        ("haltonSeq1", "<f4", (4,)),
        ("haltonSeq2", "<f4", (4,)),
        ("haltonSeq3", "<f4", (4,)),
        ("haltonSeq4", "<f4", (4,)),
        ("time", "<f4", (2,)),
        ("frameCountMod16", "<i4"),
    ]
)
SUNSKY_DTYPE = np.dtype(
    [("sunLocation", "<f4", (4,)), ("sunDirection", "<f4", (4,)), ("lightColor", "<f4", (4,)), ("sunIntensity", "<f4")]
)
TUNING_DTYPE = np.dtype(
    [
        ("coverage", "<f4"),
        ("sun_location", "<f4", (3,)),
        ("sky_sun_location", "<f4", (3,)),
        ("wind_direction", "<f4", (3,)),
        ("cloud_speed", "<f4"),
        ("cloud_top_offset", "<f4"),
        ("base_density_factor", "<f4"),
        ("use_weather", "<u4"),
        ("weather_scale", "<f4"),
    ]
)
assert CAMERA_DTYPE.itemsize == 152 and TIME_DTYPE.itemsize == 76 and SUNSKY_DTYPE.itemsize == 52
assert TUNING_DTYPE.itemsize == 60

EARTH_RADIUS = 6371000.0
ATMOSPHERE_RADIUS_OUTER = EARTH_RADIUS + 20000.0


def default_tuning() -> np.ndarray:
    """The literals of cloudRayMarch.comp (:89-95, :529, :571) as an MtTuning record."""
    t = np.zeros((), TUNING_DTYPE)
    t["coverage"] = 0.6
    t["sun_location"] = (0.0, f32(ATMOSPHERE_RADIUS_OUTER) * f32(0.9), -f32(ATMOSPHERE_RADIUS_OUTER) * f32(0.9))
    t["sky_sun_location"] = (0.0, f32(EARTH_RADIUS) * f32(2.0), -f32(EARTH_RADIUS) * f32(10.0))
    t["wind_direction"] = (1.0, 0.0, 0.0)
    t["cloud_speed"] = 0.080
    t["cloud_top_offset"] = 1.0
    t["base_density_factor"] = 0.380
    t["use_weather"] = 0
    t["weather_scale"] = 1.0
    return t


def sun_on_elevation_circle(elevation_deg: float) -> tuple[float, float, float]:
    """Sun position for BASELINE config 5: the circle through the reference SUN_LOCATION (45 degrees) in the
    y-z plane, radius 0.9 * R_outer * sqrt(2)."""
    r = 0.9 * ATMOSPHERE_RADIUS_OUTER * math.sqrt(2.0)
    e = math.radians(elevation_deg)
    return (0.0, r * math.sin(e), -r * math.cos(e))


# ---------------------------------------------------------------------------------------------
# float32 vector helpers in glm's operation order
# ---------------------------------------------------------------------------------------------
def _v(x, y, z):
    return np.array([x, y, z], dtype=f32)


def _dot(a, b):
    # glm::dot(vec3): tmp = a*b; tmp.x + tmp.y + tmp.z
    t = a * b
    return f32(f32(t[0] + t[1]) + t[2])


def _normalize(a):
    # glm::normalize = v * inversesqrt(dot(v, v)); inversesqrt = 1 / sqrt
    return (a * f32(f32(1.0) / np.sqrt(_dot(a, a), dtype=f32))).astype(f32)


def _cross(a, b):
    return _v(
        f32(a[1] * b[2]) - f32(b[1] * a[2]),
        f32(a[2] * b[0]) - f32(b[2] * a[0]),
        f32(a[0] * b[1]) - f32(b[0] * a[1]),
    )


def look_at_rh(eye, center, up) -> np.ndarray:
    f = _normalize((center - eye).astype(f32))
    s = _normalize(_cross(f, up))
    u = _cross(s, f)
    m = np.eye(4, dtype=f32)  # m[c][r]
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0] = -_dot(s, eye)
    m[3][1] = -_dot(u, eye)
    m[3][2] = _dot(f, eye)
    return m


def perspective_rh_zo(fovy, aspect, z_near, z_far) -> np.ndarray:
    fovy, aspect, z_near, z_far = f32(fovy), f32(aspect), f32(z_near), f32(z_far)
    tan_half = f32(math.tan(float(fovy / f32(2))))  # std::tan(float)
    m = np.zeros((4, 4), dtype=f32)
    m[0][0] = f32(1) / (aspect * tan_half)
    m[1][1] = f32(1) / tan_half
    m[2][2] = z_far / (z_near - z_far)
    m[2][3] = -f32(1)
    m[3][2] = -(z_far * z_near) / (z_far - z_near)
    return m


def rotate(angle_rad, axis) -> np.ndarray:
    """Upper 3x3 of glm::rotate(mat4(1), angle, axis), indexed [col][row]."""
    a = f32(angle_rad)
    c, s = f32(math.cos(float(a))), f32(math.sin(float(a)))
    ax = _normalize(axis)
    t = (f32(f32(1) - c) * ax).astype(f32)
    r = np.zeros((3, 3), dtype=f32)
    r[0][0] = c + t[0] * ax[0]
    r[0][1] = t[0] * ax[1] + s * ax[2]
    r[0][2] = t[0] * ax[2] - s * ax[1]
    r[1][0] = t[1] * ax[0] - s * ax[2]
    r[1][1] = c + t[1] * ax[1]
    r[1][2] = t[1] * ax[2] + s * ax[0]
    r[2][0] = t[2] * ax[0] + s * ax[1]
    r[2][1] = t[2] * ax[1] - s * ax[0]
    r[2][2] = c + t[2] * ax[2]
    return r


def _radians(deg):
    return f32(f32(deg) * f32(0.01745329251994329576923690768489))


class Camera:
    """camera.h / camera.cpp.  Default = main.cpp:157-158: eye (0,0,2) -> ref (0,0,1), fovy 45, near .1, far 1000."""

    def __init__(self, width, height, eye=(0.0, 0.0, 2.0), ref=(0.0, 0.0, 1.0), fovy=45.0, near=0.1, far=1000.0):
        self.width, self.height = int(width), int(height)
        self.eye = _v(*eye)
        self.ref = _v(*ref)
        self.fovy = f32(fovy)
        self.near, self.far = f32(near), f32(far)
        self.world_up = _v(0, 1, 0)
        self.recompute_attributes()

    def recompute_attributes(self):  # camera.cpp:68-77
        self.forward = _normalize((self.ref - self.eye).astype(f32))
        self.right = _normalize(_cross(self.forward, self.world_up))
        self.up = _cross(self.right, self.forward)
        self.aspect = f32(self.width) / f32(self.height)

    def _rotate_about(self, deg, axis):  # camera.cpp:79-96
        rot = rotate(_radians(deg), axis)
        r = (self.ref - self.eye).astype(f32)
        # vec3(rotation * vec4(ref, 1)): column-major product, translation column is zero
        out = np.zeros(3, dtype=f32)
        for row in range(3):
            out[row] = f32(f32(f32(rot[0][row] * r[0]) + f32(rot[1][row] * r[1])) + f32(rot[2][row] * r[2]))
        self.ref = (out + self.eye).astype(f32)
        self.recompute_attributes()

    def rotate_about_up(self, deg):
        self._rotate_about(deg, self.up)

    def rotate_about_right(self, deg):
        self._rotate_about(deg, self.right)

    def _translate(self, axis, amt):  # camera.cpp:98-118
        t = (axis * f32(amt)).astype(f32)
        self.eye = (self.eye + t).astype(f32)
        self.ref = (self.ref + t).astype(f32)
        self.recompute_attributes()

    def translate_along_look(self, amt):
        self._translate(self.forward, amt)

    def translate_along_right(self, amt):
        self._translate(self.right, amt)

    def translate_along_up(self, amt):
        self._translate(self.up, amt)

    def ubo(self) -> np.ndarray:  # Camera::UpdateBuffer, camera.cpp:31-42
        u = np.zeros((), CAMERA_DTYPE)
        view = look_at_rh(self.eye, self.ref, self.up)
        proj = perspective_rh_zo(_radians(self.fovy), f32(self.width) / f32(self.height), self.near, self.far)
        proj[1][1] *= f32(-1)
        u["view"] = view.reshape(16)
        u["proj"] = proj.reshape(16)
        u["eye"] = (self.eye[0], self.eye[1], self.eye[2], 1.0)
        # std::abs(std::tan(fovy*0.5 * (PI / 180.0))) in double with PI = 3.14159 (camera.h:10), stored as float
        tan_y = abs(math.tan(float(self.fovy) * 0.5 * (3.14159 / 180.0)))
        u["tanFovBy2"][1] = f32(tan_y)
        u["tanFovBy2"][0] = self.aspect * f32(tan_y)
        return u


def halton_sequence_at(index: int, base: int) -> np.float32:
    """Scene::HaltonSequenceAt, Scene.cpp:125-138 (float accumulation)."""
    f, r = f32(1.0), f32(0.0)
    while index > 0:
        f = f32(f / f32(base))
        r = f32(r + f32(f * f32(index % base)))
        index = int(math.floor(index / base))
    return r


class Scene:
    """Scene.h / Scene.cpp: the Time uniform.  frameCount starts at 0 and is incremented BEFORE the first frame."""

    def __init__(self):
        self.time = np.zeros((), TIME_DTYPE)
        h = [halton_sequence_at(i, 3) for i in range(1, 17)]  # Scene.cpp:95-114: base 3 for all sixteen
        self.time["haltonSeq1"] = h[0:4]
        self.time["haltonSeq2"] = h[4:8]
        self.time["haltonSeq3"] = h[8:12]
        self.time["haltonSeq4"] = h[12:16]
        self.time["frameCountMod16"] = 0

    def update_time(self, dt: float):  # Scene.cpp:65-85 with a caller-supplied delta instead of the wall clock
        self.time["time"][0] = f32(dt)
        self.time["time"][1] = f32(self.time["time"][1] + f32(dt))
        self.time["frameCountMod16"] = (int(self.time["frameCountMod16"]) + 1) % 16

    def ubo(self) -> np.ndarray:
        return self.time.copy()


class Sky:
    """Sky.h / Sky.cpp: the SunAndSky uniform (Sky.cpp:64-74)."""

    def __init__(self):
        self.s = np.zeros((), SUNSKY_DTYPE)
        self.update_sun_and_sky()

    def update_sun_and_sky(self):
        self.s["sunLocation"] = (0.0, 1.0, 0.0, 0.0)
        self.s["sunDirection"] = (1.0, 1.0, 1.0, 0.0)
        self.s["lightColor"] = (1.0, 1.0, 0.57, 1.0)
        self.s["sunIntensity"] = 5.0

    def ubo(self) -> np.ndarray:
        return self.s.copy()


def cloud_dispatch_threads(width: int, height: int) -> tuple[int, int]:
    """Threads launched by the reference cloud dispatch (Renderer.cpp:713-716): workgroups of 32 over W/4, H/4
    with the integer truncation of `window_width / 4` kept."""
    return (((width // 4) + 31) // 32) * 32, (((height // 4) + 31) // 32) * 32


def cloud_pixels_written(width: int, height: int, frame_id: int) -> np.ndarray:
    """Boolean H x W mask of the pixels one cloud dispatch writes (cloudRayMarch.comp:695-704)."""
    tx, ty = cloud_dispatch_threads(width, height)
    px, py = frame_id // 4, frame_id % 4
    m = np.zeros((height, width), dtype=bool)
    xs = np.arange(tx) * 4 + px
    ys = np.arange(ty) * 4 + py
    xs, ys = xs[xs < width], ys[ys < height]
    m[np.ix_(ys, xs)] = True
    return m
