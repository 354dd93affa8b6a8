"""CPU tests of the oracle against hand-derived known answers (the reference ships no golden vectors:
SURVEY.md section 4 -- these pins are what stands in for them)."""
import hashlib
import math
import re
from fractions import Fraction
from pathlib import Path

import numpy as np
import pytest

from conftest import default_scene
from meteoros_b200 import scene, textures

ROOT = Path(__file__).resolve().parents[1]


def test_noise_fixture_matches_reference_fingerprints(noise):
    # SHA-256 of the decoded reference textures
    for k, v in noise.items():
        assert hashlib.sha256(v.tobytes()).hexdigest() == textures.SHA256[k]
    assert tuple(noise["low"][0, 0, 0]) == (144, 172, 215, 198)  # first texel of LowFrequency(1).tga
    assert noise["high"][..., 3].max() == 0  # hi-freq alpha is all zero


def test_halton_table_base3():
    # Scene.cpp:95-114 uses base 3 for all sixteen values; exact radical inverses:
    def radical_inverse(i, b):
        f, r = Fraction(1), Fraction(0)
        while i:
            f /= b
            r += f * (i % b)
            i //= b
        return r

    sc = scene.Scene()
    got = np.concatenate([sc.time[f"haltonSeq{k}"] for k in range(1, 5)])
    want = [float(radical_inverse(i, 3)) for i in range(1, 17)]
    assert np.allclose(got, want, rtol=0, atol=1e-7)
    assert got[0] == np.float32(1 / 3) and got[3] == np.float32(4 / 9)
    assert int(sc.time["frameCountMod16"]) == 0
    sc.update_time(1 / 60)
    assert int(sc.time["frameCountMod16"]) == 1  # first rendered frame has id 1
    for _ in range(15):
        sc.update_time(1 / 60)
    assert int(sc.time["frameCountMod16"]) == 0


def test_ubo_layouts_match_header():
    hdr = (ROOT / "include" / "meteoros_b200.h").read_text()
    assert "float tanFovBy2[2]" in hdr and "int32_t frameCountMod16" in hdr
    assert scene.CAMERA_DTYPE.fields["proj"][1] == 64 and scene.CAMERA_DTYPE.fields["eye"][1] == 128
    assert scene.CAMERA_DTYPE.fields["tanFovBy2"][1] == 144 and scene.CAMERA_DTYPE.itemsize == 152
    assert scene.TIME_DTYPE.fields["time"][1] == 64 and scene.TIME_DTYPE.fields["frameCountMod16"][1] == 72
    assert scene.TIME_DTYPE.itemsize == 76
    assert scene.SUNSKY_DTYPE.fields["lightColor"][1] == 32 and scene.SUNSKY_DTYPE.itemsize == 52


def test_default_camera_ubo():
    cam, _, sky, tun = default_scene(1920, 1080)
    view = cam["view"].reshape(4, 4)  # [col][row]
    # eye (0,0,2) looking down -z: identity rotation, translation (0,0,-2)
    assert np.array_equal(np.abs(view[:3, :3]), np.eye(3, dtype=np.float32))
    assert tuple(view[3]) == (0.0, 0.0, -2.0, 1.0)
    proj = cam["proj"].reshape(4, 4)
    assert proj[1][1] < 0 and proj[2][3] == -1.0  # y flip, RH
    t = math.tan(math.radians(22.5))
    assert abs(proj[0][0] - 1 / (16 / 9 * t)) < 1e-6 and abs(proj[1][1] + 1 / t) < 1e-6
    assert abs(proj[2][2] - 1000 / (0.1 - 1000)) < 1e-6  # zero-to-one depth
    # tanFovBy2 uses PI = 3.14159 in double (camera.h:10, camera.cpp:40)
    assert cam["tanFovBy2"][1] == np.float32(abs(math.tan(45 * 0.5 * (3.14159 / 180.0))))
    assert cam["tanFovBy2"][0] == np.float32(np.float32(1920) / np.float32(1080)) * cam["tanFovBy2"][1]
    assert tuple(sky["lightColor"]) == (1.0, 1.0, np.float32(0.57), 1.0) and sky["sunIntensity"] == 5.0
    assert tuple(tun["sun_location"]) == (0.0, 5751900.0, -5751900.0)
    assert tuple(tun["sky_sun_location"]) == (0.0, 12742000.0, -63710000.0)


def test_pan_rotation_is_quarter_degree():
    cam = scene.Camera(1920, 1080)
    f0 = cam.forward.copy()
    cam.rotate_about_up(0.25)
    ang = math.degrees(math.asin(float(np.linalg.norm(np.cross(f0.astype(np.float64), cam.forward.astype(np.float64))))))
    assert abs(ang - 0.25) < 1e-4
    assert abs(float(np.linalg.norm(cam.ref - cam.eye)) - 1.0) < 1e-6


@pytest.mark.parametrize("w,h", [(1920, 1080), (1284, 720), (130, 70), (64, 36), (3840, 2160), (7, 5)])
def test_cloud_grid_and_pixel_selection(oracle_mod, w, h):
    tx, ty = oracle_mod.cloud_grid(w, h)
    assert (tx, ty) == scene.cloud_dispatch_threads(w, h)
    assert tx % 32 == 0 and ty % 32 == 0 and tx >= 32
    # Renderer.cpp:713: ceil() is applied to an already truncated integer division
    assert tx == ((w // 4 + 31) // 32) * 32
    union = np.zeros((h, w), bool)
    for fid in range(16):
        m = scene.cloud_pixels_written(w, h, fid)
        assert not (union & m).any()  # the sixteen ids never overlap
        union |= m
        ys, xs = np.nonzero(m)
        if len(xs):
            assert set(xs % 4) == {fid // 4} and set(ys % 4) == {fid % 4}  # (pX, pY) = (id/4, id%4), not Bayer
    if w == 130:  # 130/4 = 32 -> exactly one workgroup -> columns 128, 129 are never marched
        assert union[:, :128].all() and not union[:, 128:].any()
    elif w % 4 == 0 or (w // 4) % 32 != 0:
        assert union.all()


def test_oracle_writes_exactly_the_selected_pixels(oracle_mod, noise):
    w, h = 52, 30
    for fid in (0, 5, 15):
        cam, tm, _, tun = default_scene(w, h, frame_id=fid)
        hdr = np.full((h, w, 4), -7.0, np.float32)
        mask = np.full((h, w, 4), -7.0, np.float32)
        oracle_mod.cloud(cam, tm, tun, noise, w, h, full=False, hdr=hdr, mask=mask)
        written = hdr[..., 3] != -7.0
        assert np.array_equal(written, scene.cloud_pixels_written(w, h, fid))
        assert np.array_equal(mask[..., 0] != -7.0, written)
        assert np.all(hdr[written][:, 3] == 1.0)


def test_full_dispatch_equals_sixteen_single_dispatches(oracle_mod, noise):
    w, h = 64, 36
    cam, tm, _, tun = default_scene(w, h)
    full = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True)
    hdr = np.zeros((h, w, 4), np.float32)
    mask = np.zeros((h, w, 4), np.float32)
    for fid in range(16):
        tm["frameCountMod16"] = fid
        oracle_mod.cloud(cam, tm, tun, noise, w, h, full=False, hdr=hdr, mask=mask)
    assert np.array_equal(full["hdr"], hdr) and np.array_equal(full["mask"], mask)


def test_ray_sphere_known_answers(oracle_mod):
    R = 6378500.0
    c = (0.0, -6371000.0, -2.0)
    # straight up from the eye: the inner shell is 7500 m above the ground plane
    p, t, valid = oracle_mod.ray_sphere((0.0, 0.0, -2.0), (0.0, 1.0, 0.0), c, R)
    assert valid and abs(p[1] - 7500.0) < 1.0 and abs(p[0]) < 1e-3
    # the quirk (cloudRayMarch.comp:238-239,269): t = |p_world - rO_unit|, rO_unit = (0, 6371000/6378500, 0)
    ro_unit = np.array([0.0, 6371000.0 / R, 0.0])
    assert abs(float(t) - np.linalg.norm(p.astype(np.float64) - ro_unit)) < 0.5
    # a ray that misses (origin far outside, pointing away) is invalid with point = 0, t = 0
    p, t, valid = oracle_mod.ray_sphere((0.0, 1e8, 0.0), (0.0, 1.0, 0.0), c, R)
    assert not valid and not p.any() and t == 0.0
    # horizontal ray from inside: distance to the shell = sqrt(R^2 - r0^2)
    p, t, valid = oracle_mod.ray_sphere((0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, -6371000.0, 0.0), R)
    assert valid and abs(p[0] - math.sqrt(R * R - 6371000.0**2)) / p[0] < 2e-3


def test_sampler_texel_centres_wrap_and_bruteforce(oracle_mod, noise):
    vol = noise["high"]  # 32^3
    n = 32
    # texel centre (i+.5)/n returns exactly the texel
    for (x, y, z) in [(0, 0, 0), (31, 31, 31), (5, 17, 9)]:
        got = oracle_mod.sample3d(vol, (x + 0.5) / n, (y + 0.5) / n, (z + 0.5) / n)
        assert np.allclose(got, vol[z, y, x].astype(np.float64) / 255.0, atol=1e-7)
    # REPEAT: coordinates one period apart give the same bits; negative coordinates wrap
    a = oracle_mod.sample3d(vol, 0.3, 0.6, 0.9)
    assert np.allclose(oracle_mod.sample3d(vol, 1.3, -0.4, 2.9), a, atol=2e-6)
    # s = 0 blends texel n-1 and texel 0 equally
    got = oracle_mod.sample3d(vol, 0.0, 0.5 / n, 0.5 / n)
    want = (vol[0, 0, n - 1].astype(np.float64) + vol[0, 0, 0]) / 2 / 255
    assert np.allclose(got, want, atol=1e-6)
    # brute force in float64 at random coordinates
    rng = np.random.default_rng(7)
    for s, t, r in rng.random((50, 3)) * 3 - 1:
        u, v, w = s * n - 0.5, t * n - 0.5, r * n - 0.5
        i, j, k = math.floor(u), math.floor(v), math.floor(w)
        fx, fy, fz = u - i, v - j, w - k
        acc = np.zeros(4)
        for dz, wz in ((0, 1 - fz), (1, fz)):
            for dy, wy in ((0, 1 - fy), (1, fy)):
                for dx, wx in ((0, 1 - fx), (1, fx)):
                    acc += wx * wy * wz * vol[(k + dz) % n, (j + dy) % n, (i + dx) % n]
        assert np.allclose(oracle_mod.sample3d(vol, np.float32(s), np.float32(t), np.float32(r)), acc / 255, atol=2e-5)
    img = noise["curl"]
    got = oracle_mod.sample2d(img, 10.5 / 128, 77.5 / 128)
    assert np.allclose(got, img[77, 10].astype(np.float64) / 255, atol=1e-7)


def test_encode_float_rgba_roundtrip(oracle_mod):
    dec = 1.0 / np.array([1.0, 255.0, 65025.0, 16581375.0])
    for v in (0.0, 0.1, 0.25, 0.5, 0.999, 1.0, 1.25):
        e = oracle_mod.encode_float_rgba(v)
        # the encoding keeps fract(v): v = 1.25 (accumulated density <= 0.95) decodes to 0.25
        assert abs(float(e.astype(np.float64) @ dec) - (v - math.floor(v))) < 1e-6
        assert (e > -1e-6).all() and (e < 1.0).all()


def test_wang_hash_integer_exact(oracle_mod):
    def wang(u, v, s):
        m = 0xFFFFFFFF
        seed = ((u * 1664525 + v) + s) & m
        seed = ((seed ^ 61) ^ (seed >> 16)) & m
        seed = (seed * 9) & m
        seed = seed ^ (seed >> 4)
        seed = (seed * 0x27D4EB2D) & m
        return seed ^ (seed >> 15)

    for u, v, s in [(0, 0, 0), (1, 2, 3), (1919, 1079, 123), (3839, 2159, 4000000000), (7, 0, 0xFFFFFFFF)]:
        assert oracle_mod.wang_hash(u, v, s) == wang(u, v, s)
    assert oracle_mod.wang_hash(0, 0, 0) == wang(0, 0, 0) != 0


def test_tonemap_known_values(oracle_mod):
    def u2(x):
        A, B, C, D, E, F = 0.15, 0.5, 0.1, 0.2, 0.02, 0.3
        return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F

    _, tm, _, _ = default_scene(8, 4, total_time=5.9)
    hdr = np.zeros((4, 8, 4), np.float32)
    hdr[..., :3] = np.linspace(0.01, 3.0, 32, dtype=np.float32).reshape(4, 8, 1)
    ldr, f = oracle_mod.tonemap(tm, hdr, want_f32=True)
    for y in range(4):
        for x in range(8):
            noise = oracle_mod.wang_hash(x, y, 5) / 2**32 * 0.01  # uint(5.9) = 5
            want = (u2(2.5 * float(hdr[y, x, 0])) / u2(100.0)) ** (1 / 2.2) + noise
            assert abs(f[y, x, 0] - want) < 2e-6
            assert abs(int(ldr[y, x, 1]) - round(min(max(want, 0), 1) * 255)) <= 1
    assert (ldr[..., 3] == 255).all()


def test_sky_colour_against_float64_restatement(oracle_mod):
    """Independent float64 evaluation of the Preetham block (cloudRayMarch.comp:401-467)."""
    sun = np.array([0.0, 12742000.0, -63710000.0])
    origin = np.array([0.0, 0.0, -2.0])

    def sky(d):
        E = 2.718281828459
        sd = (sun - origin) / np.linalg.norm(sun - origin)
        zc = sun[1] / np.linalg.norm(sun)
        sunE = 0.780 * 1000.0 * max(0.0, 1.0 - E ** (-((1.6110731557 - math.acos(zc)) / 1.5)))
        fade = 1.0 - min(max(1.0 - math.exp(sun[1] / 450000.0), 0.0), 1.0)
        bR = np.array([5.804542996261093e-6, 1.3562911419845635e-5, 3.0265902468824876e-5]) * (1.0 + fade)
        bM = 0.434 * (2.0 * 1e-17) * np.array([1.839991851443397, 2.779802391966052, 4.079047954386109]) * 0.005
        zen = math.acos(max(0.0, d[1]))
        inv = 1.0 / (math.cos(zen) + 0.15 * (93.885 - zen * 180.0 / 3.14159265) ** -1.253)
        fex = np.exp(-bR * 8.4e3 * inv + bM * 1.25e3 * inv)
        ct = float(sd @ d)
        rph = 0.05968310365946075 * (1.0 + (ct * 0.5 + 0.5) ** 2)
        g = 0.8
        mph = (1 - g * g) / (1 + g * g - 2 * g * ct) ** 1.5 * 0.07957747154594767
        betas = (bR * rph + bM * mph) / (bR + bM)
        lin = (sunE * betas * (1 - fex)) ** 1.5
        yd = min(max((1 - sd[1]) ** 5, 0), 1)
        lin = lin * ((1 - yd) + (sunE * betas * fex) ** 0.5 * yd)
        l0 = 0.1 * fex  # sun disk term is 0 away from the sun
        return (lin + l0) * 0.04 + np.array([0.0, 0.0003, 0.00075])

    for d in ([0.0, 0.3, -0.95], [0.5, 0.06, -0.86], [-0.2, 0.9, 0.3]):
        d = np.array(d) / np.linalg.norm(d)
        got = oracle_mod.atmosphere_color(d.astype(np.float32), (sun - origin).astype(np.float32), 0.780, sun.astype(np.float32))
        assert np.allclose(got, sky(d), rtol=2e-4), (got, sky(d))


def test_default_frame_statistics(oracle_mod, noise):
    """Work profile of the default cloudscape at 480x270: SURVEY.md section 8a quotes 49 % / 8 % / 43 % for the
    three branches, ~54 steps per marching ray, shell distances 19-109 km / 51-251 km, steps of 0.64-2.4 km."""
    w, h = 480, 270
    cam, tm, _, tun = default_scene(w, h)
    r = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True, counters=True, debug=True)
    d, c = r["debug"], r["counters"]
    frac = [(d["branch"] == b).mean() for b in range(3)]
    assert abs(frac[0] - 0.49) < 0.02 and abs(frac[1] - 0.08) < 0.02 and abs(frac[2] - 0.43) < 0.02
    m = d["branch"] == 2
    assert c["rays"] == w * h and c["rays_marched"] == m.sum()
    assert 50 < c["steps"] / c["rays_marched"] < 58
    assert 19e3 < d["t_in"][m].min() < 20e3 and 105e3 < d["t_in"][m].max() < 112e3
    assert 51e3 < d["t_out"][m].min() < 52e3 and 245e3 < d["t_out"][m].max() < 255e3
    assert 600 < d["step_size"][m].min() and d["step_size"][m].max() < 2500
    full_len = m & (d["accum"] < 1.0)  # rays that did not leave through accumDensity >= 1
    assert d["steps"][full_len].min() >= 35 and d["steps"][m].max() <= 60
    assert c["early_exits"] == (m & (d["accum"] >= 1.0)).sum() > 0
    assert np.isfinite(r["hdr"]).all() and (r["hdr"][..., 3] == 1.0).all()
    # mask of non-marching pixels is zero; of marching pixels encodes 25*min(.05, 1-accum)
    assert not r["mask"][~m].any()
    dec = r["mask"].astype(np.float64) @ (1.0 / np.array([1.0, 255.0, 65025.0, 16581375.0]))
    v = 25.0 * np.minimum(0.05, 1.0 - d["accum"].astype(np.float64))
    assert np.allclose(dec[m], (v - np.floor(v))[m], atol=2e-6)


def test_golden_frame_regression(oracle_mod, noise):
    """tests/golden/cloud_64x36.npz was minted by tests/golden/make_goldens.py from the REFERENCE'S OWN cloud shader
    (compiled from its text, oracle/refshaders.py) where /root/reference exists; here, on any machine, the oracle must
    reproduce it (the per-ray records in the file are the oracle's own)."""
    g = np.load(ROOT / "tests" / "golden" / "cloud_64x36.npz")
    w, h = 64, 36
    cam, tm, _, tun = default_scene(w, h, frame_id=int(g["frame_id"]), total_time=float(g["total_time"]), yaw=float(g["yaw"]))
    r = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True, debug=True)
    assert np.array_equal(r["debug"]["steps"], g["steps"]) and np.array_equal(r["debug"]["jitter_hash"], g["jitter_hash"])
    assert np.array_equal(r["mask"], g["mask"])
    assert np.allclose(r["hdr"], g["hdr"], rtol=1e-5, atol=1e-7)  # libm exp/pow may differ in the last ulp across hosts


def test_golden_sequence_regression(oracle_mod, noise):
    """Four frames of the full loop (main.cpp:172-194 order) against tests/golden/sequence_96x54.npz (reference shaders)."""
    g = np.load(ROOT / "tests" / "golden" / "sequence_96x54.npz")
    w, h = 96, 54
    cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
    tun = scene.default_tuning()
    img = [np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)]
    mask = np.zeros((h, w, 4), np.float32)
    cur, cam_old = 0, cam.ubo()
    for k in range(4):
        cam.rotate_about_up(0.25)
        sc.update_time(1 / 60)
        c, t = cam.ubo(), sc.ubo()
        assert int(t["frameCountMod16"]) == k + 1
        img[cur] = oracle_mod.reproject(c, cam_old, t, img[cur ^ 1])
        oracle_mod.cloud(c, t, tun, noise, w, h, full=False, hdr=img[cur], mask=mask)
        img[cur] = oracle_mod.godrays(c, sky.ubo(), mask, img[cur])
        ldr = oracle_mod.tonemap(t, img[cur])
        assert np.allclose(img[cur], g["hdr"][k], rtol=1e-5, atol=1e-7)
        assert np.abs(ldr.astype(int) - g["ldr"][k].astype(int)).max() <= 1
        cur ^= 1
        cam_old = c


def test_golden_live_sequence_all_five_shaders(oracle_mod, noise):
    """Sixteen frames of REPROJ, CLOUD, GODRAYS, TONEMAP, TXAA chained, against tests/golden/live_sequence_96x54.npz,
    which the reference's five shaders wrote (tests/golden/make_goldens.py).  Every pixel id 1..15, 0 is exercised."""
    g = np.load(ROOT / "tests" / "golden" / "live_sequence_96x54.npz")
    assert "reference shaders" in str(g["source"])
    w, h = 96, 54
    cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
    tun = scene.default_tuning()
    img = [np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)]
    mask = np.zeros((h, w, 4), np.float32)
    hist = np.zeros((h, w, 4), np.uint8)
    cur, cam_old = 0, cam.ubo()
    for k in range(16):
        cam.rotate_about_up(0.25)
        sc.update_time(1 / 60)
        c, t = cam.ubo(), sc.ubo()
        img[cur] = oracle_mod.reproject(c, cam_old, t, img[cur ^ 1])
        oracle_mod.cloud(c, t, tun, noise, w, h, full=False, hdr=img[cur], mask=mask)
        img[cur] = oracle_mod.godrays(c, sky.ubo(), mask, img[cur])
        ldr = oracle_mod.tonemap(t, img[cur])
        hist = oracle_mod.txaa(c, cam_old, t, ldr, hist)
        if k == 0:
            assert np.allclose(img[cur], g["hdr_first"], rtol=1e-5, atol=1e-7)
        # libm exp / pow may differ in the last ulp across hosts: one LSB of slack on the 8-bit images
        assert np.abs(ldr.astype(int) - g["ldr"][k].astype(int)).max() <= 1
        assert np.abs(hist.astype(int) - g["txaa"][k].astype(int)).max() <= 1
        cur ^= 1
        cam_old = c
    assert np.allclose(img[cur ^ 1], g["hdr_last"], rtol=1e-5, atol=1e-7)
    assert np.abs(mask - g["mask_last"]).max() <= 1.0 / 255.0


def test_density_height_gradient_known_answers(oracle_mod):
    """getDensityHeightGradientForPoint (cloudRayMarch.comp:475-487), used only by the opt-in weather path."""
    def remap(v, a, b, c, d):
        return c + (v - a) / (b - a) * (d - c)

    def ref(h, ct):
        h = min(max(h, 0.0), 1.0)
        sc = max(0.0, remap(h, 0.0, 0.25, 0.0, 1.0) * remap(h, 0.3, 0.65, 1.0, 0.0))
        st = max(0.0, remap(h, 0.0, 0.1, 0.0, 1.0) * remap(h, 0.2, 0.3, 1.0, 0.0))
        a = st + (sc - st) * min(max(ct * 2.0, 0.0), 1.0)
        b = sc + (st - sc) * min(max((ct - 0.5) * 2.0, 0.0), 1.0)
        return a + (b - a) * ct

    g = oracle_mod.density_height_gradient
    assert g(0.0, 0.0) == 0.0 and g(0.0, 1.0) == 0.0            # nothing at the shell's floor
    assert g(0.9, 0.3) == 0.0 and g(2.0, 0.5) == 0.0            # above every profile / clamped height
    assert g(0.1, 0.0) == pytest.approx(2.0, abs=1e-6)          # remap() is unclamped: the falling ramp reads 2 at h = 0.1
    rng = np.random.default_rng(5)
    for h, ct in rng.random((200, 2)):
        assert g(h, ct) == pytest.approx(ref(h, ct), abs=2e-6)


def test_golden_fullsize_digests_1080p(oracle_mod, noise):
    """The oracle against the digests the reference's Cloud shader left of its 1920x1080 frame (tests/golden/
    cloud_fullsize_digests.npz, written by make_goldens.py where /root/reference exists): mask and alpha rows CRC-equal,
    sampled colour and row sums equal bit for bit (the oracle is an op-for-op restatement)."""
    import pathlib
    import zlib

    from conftest import default_scene

    w, h = 1920, 1080
    g = np.load(pathlib.Path(__file__).parent / "golden" / "cloud_fullsize_digests.npz")
    cam, tm, _, tun = default_scene(w, h)
    r = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True)
    sub = int(g[f"sub_{w}x{h}"])
    assert np.array_equal(np.array([zlib.crc32(r["mask"][y].tobytes()) for y in range(h)], np.uint32), g[f"mask_crc_{w}x{h}"])
    assert np.array_equal(np.array([zlib.crc32(np.ascontiguousarray(r["hdr"][y, :, 3]).tobytes()) for y in range(h)], np.uint32),
                          g[f"alpha_crc_{w}x{h}"])
    assert np.array_equal(r["hdr"][::sub, ::sub, :3], g[f"hdr_sub_{w}x{h}"])
    assert np.array_equal(r["hdr"][..., :3].astype(np.float64).sum(axis=1), g[f"hdr_rowsum_{w}x{h}"])
