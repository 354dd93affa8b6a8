"""The kernels' per-pixel cores (meteoros_b200/csrc/*_core.cuh), compiled for the host, against the oracle.

The CUDA kernels restructure the shader (hoisted frame constants, one shared erosion fetch for the six cone
samples, scalar radiance, channel-specialised filters).  This checks on CPU that none of that changes a single
bit of any decision-carrying value; the -m gpu tests then check the same through the real kernels."""
import numpy as np
import pytest

from conftest import default_scene
from tests_hostsim_loader import hostsim  # noqa: F401  (see tests_hostsim_loader.py)


@pytest.mark.parametrize("w,h,full,fid,yaw,pitch,t", [
    (96, 54, True, 1, 0.0, 0.0, 0.016),
    (130, 70, True, 7, 12.0, 3.0, 41.5),      # grid quirk: columns 128,129 never marched
    (200, 112, False, 14, -30.0, 8.0, 3.25),
    (64, 36, False, 0, 0.0, -4.0, 0.0),
])
def test_cloud_core_bit_identical(oracle_mod, noise, hostsim, w, h, full, fid, yaw, pitch, t):
    cam, tm, _, tun = default_scene(w, h, frame_id=fid, total_time=t, yaw=yaw, pitch=pitch)
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=full, counters=True, debug=True)
    hdr, mask, cnt, dbg = hostsim.cloud(cam, tm, tun, noise, w, h, full, oracle_mod.RAY_DEBUG_DTYPE)
    assert cnt == ref["counters"]
    for f in oracle_mod.RAY_DEBUG_DTYPE.names:
        assert np.array_equal(dbg[f], ref["debug"][f]), f
    assert np.array_equal(mask, ref["mask"])
    # radiance-only terms may be evaluated differently (cos(acos(x)) := x, SFU exp/pow on the GPU): ulp-level only
    assert np.allclose(hdr, ref["hdr"], rtol=2e-6, atol=0)


def test_cloud_core_tuning_sweep(oracle_mod, noise, hostsim):
    from meteoros_b200 import scene

    w, h = 80, 45
    cam, tm, _, tun = default_scene(w, h, frame_id=5, total_time=10.0)
    for cov, elev in ((0.3, 5.0), (0.9, 85.0), (0.45, 30.0)):
        tun["coverage"] = cov
        tun["sun_location"] = scene.sun_on_elevation_circle(elev)
        ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True, debug=True)
        hdr, mask, _, dbg = hostsim.cloud(cam, tm, tun, noise, w, h, True, oracle_mod.RAY_DEBUG_DTYPE)
        assert np.array_equal(dbg["accum"], ref["debug"]["accum"])
        assert np.allclose(hdr, ref["hdr"], rtol=2e-6, atol=0) and np.array_equal(mask, ref["mask"])


def test_post_cores_bit_identical(oracle_mod, hostsim):
    from meteoros_b200 import scene

    w, h = 160, 90
    rng = np.random.default_rng(3)
    prev = rng.random((h, w, 4), dtype=np.float32)
    cam = scene.Camera(w, h)
    old = cam.ubo()
    cam.rotate_about_up(0.25)
    cam.rotate_about_right(-0.5)
    new = cam.ubo()
    sc = scene.Scene()
    for fid in (1, 8, 15):
        sc.time["frameCountMod16"] = fid
        cur, taps = oracle_mod.reproject(new, old, sc.ubo(), prev, taps=True)
        cur2, taps2 = hostsim.reproject(new, old, sc.ubo(), prev)
        assert np.array_equal(taps, taps2) and np.array_equal(cur, cur2)
    mask = rng.random((h, w, 4), dtype=np.float32)
    hdr = rng.random((h, w, 4), dtype=np.float32)
    sky = scene.Sky().ubo()
    for cu in (old, new):
        want = oracle_mod.godrays(cu, sky, mask, hdr)
        got = hostsim.godrays(cu, sky["lightColor"][:3], mask, hdr)
        # the kernel decodes each mask texel once and filters the scalar (linear ops reordered): rounding-level only
        assert np.allclose(got, want, rtol=1e-6, atol=0)
        assert (np.abs((got - hdr.astype(np.float64)) - (want - hdr.astype(np.float64))) <= 2e-5 * np.abs(want - hdr) + 2.5e-7 * np.abs(hdr)).all()
    sc.time["time"][1] = 77.7
    hdr[0, 0] = 0.0
    hdr[0, 1] = (1e4, 1e-8, -1.0, 1.0)
    assert np.array_equal(oracle_mod.tonemap(sc.ubo(), hdr * 4), hostsim.tonemap(sc.ubo(), hdr * 4))


def test_txaa_core_bit_identical(oracle_mod, hostsim):
    from meteoros_b200 import scene

    w, h = 150, 84
    rng = np.random.default_rng(9)
    cur = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    prev = np.roll(cur, 2, axis=1)
    prev[::7] = rng.integers(0, 256, prev[::7].shape, dtype=np.uint8)
    cam = scene.Camera(w, h)
    old = cam.ubo()
    cam.rotate_about_up(0.25)
    sc = scene.Scene()
    for fid in (0, 5, 15):
        sc.time["frameCountMod16"] = fid
        assert np.array_equal(oracle_mod.txaa(cam.ubo(), old, sc.ubo(), cur, prev), hostsim.txaa(cam.ubo(), old, sc.ubo(), cur, prev))


def test_txaa_properties(oracle_mod):
    """Static camera, identical smooth history: the blend returns the current frame; a history far outside the
    neighbourhood box is clipped to it (postProcess_TXAA.frag:150-169)."""
    from meteoros_b200 import scene

    w, h = 64, 36
    cam = scene.Camera(w, h).ubo()
    sc = scene.Scene()
    sc.update_time(1 / 60)
    smooth = np.tile(np.linspace(40, 210, w).astype(np.uint8)[None, :, None], (h, 1, 4))
    out = oracle_mod.txaa(cam, cam, sc.ubo(), smooth, smooth)
    assert np.abs(out.astype(int) - smooth.astype(int))[2:-2, 2:-2].max() <= 1
    flat = np.full((h, w, 4), 100, np.uint8)
    wild = np.full((h, w, 4), 250, np.uint8)
    out = oracle_mod.txaa(cam, cam, sc.ubo(), flat, wild)
    assert np.abs(out.astype(int)[3:-3, 3:-3, :3] - 100).max() <= 1  # history clipped onto the (degenerate) box


def test_static_camera_reprojects_onto_itself(oracle_mod):
    """reprojection.comp does not flip v (:200-201) yet lands on (nearly) the same pixel when the camera is static:
    the ten taps stay within two texels of the pixel."""
    from meteoros_b200 import scene

    w, h = 128, 72
    cam = scene.Camera(w, h).ubo()
    sc = scene.Scene()
    sc.update_time(1 / 60)
    _, taps = oracle_mod.reproject(cam, cam, sc.ubo(), np.zeros((h, w, 4), np.float32), taps=True)
    ys, xs = np.divmod(taps, w)
    dy = np.abs(ys - np.arange(h)[:, None, None])
    dx = np.abs(xs - np.arange(w)[None, :, None])
    assert dx.max() <= 2 and dy.max() <= 2
    assert (taps >= 0).all() and (taps < w * h).all()


@pytest.mark.parametrize("scale", [1.0e-4, 3.7e-5])
def test_cloud_core_weather_path_bit_identical(oracle_mod, noise, hostsim, scale):
    """MtTuning.use_weather: cloudRayMarch.comp:515-525 (the commented-out weather block) restored, same bit-exact bar."""
    w, h = 96, 54
    cam, tm, _, tun = default_scene(w, h, frame_id=3, total_time=7.5, yaw=20.0, pitch=2.0)
    tun["use_weather"], tun["weather_scale"] = 1, scale
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True, counters=True, debug=True)
    hdr, mask, cnt, dbg = hostsim.cloud(cam, tm, tun, noise, w, h, True, oracle_mod.RAY_DEBUG_DTYPE)
    assert cnt == ref["counters"]
    for f in oracle_mod.RAY_DEBUG_DTYPE.names:
        assert np.array_equal(dbg[f], ref["debug"][f]), f
    assert np.array_equal(mask, ref["mask"])
    assert np.allclose(hdr, ref["hdr"], rtol=2e-6, atol=0)
    # the path is live: it changes which rays find cloud
    tun["use_weather"] = 0
    base = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True, debug=True)
    assert not np.array_equal(base["debug"]["accum"], ref["debug"]["accum"])
    assert (ref["debug"]["accum"] > 0).mean() > 0.02


def test_randomised_scenes_kernel_cores(oracle_mod, noise, hostsim):
    """Sixty random cameras -- eye inside, between and above the cloud shells, any heading, fov 20..100 degrees -- frame
    ids, times and window sizes: the kernels' per-pixel cores against the oracle (which tests/test_reference_shaders.py holds
    to the reference's shader text on the same kind of sweep)."""
    from meteoros_b200 import scene

    rng = np.random.default_rng(77)
    marched = 0
    for trial in range(60):
        w, h = int(rng.integers(9, 70)), int(rng.integers(9, 50))
        ey = float(-rng.choice([0.0, 10.0, 5e3, 7.4e3, 7.6e3, 9e3, 1.9e4, 2.1e4, 3e4]))
        eye = (float(rng.uniform(-3e3, 3e3)), ey, float(rng.uniform(-3e3, 3e3)))
        cam = scene.Camera(w, h, eye=eye, ref=(eye[0], eye[1], eye[2] - 1.0), fovy=float(rng.uniform(20.0, 100.0)))
        cam.rotate_about_up(float(rng.uniform(-180.0, 180.0)))
        cam.rotate_about_right(float(rng.uniform(-60.0, 85.0)))
        old = cam.ubo()
        cam.rotate_about_up(float(rng.uniform(-3.0, 3.0)))
        cam.translate_along_look(float(rng.uniform(-50.0, 50.0)))
        new = cam.ubo()
        sc, tun = scene.Scene(), scene.default_tuning()
        sc.time["time"] = (0.016, float(rng.uniform(0.0, 500.0)))
        sc.time["frameCountMod16"] = int(rng.integers(0, 16))
        tm = sc.ubo()
        ref = oracle_mod.cloud(new, tm, tun, noise, w, h, full=True, debug=True, counters=True)
        hdr, mask, cnt, dbg = hostsim.cloud(new, tm, tun, noise, w, h, True, oracle_mod.RAY_DEBUG_DTYPE)
        marched += cnt["rays_marched"]
        assert cnt == ref["counters"], trial
        assert np.array_equal(mask, ref["mask"], equal_nan=True) and np.array_equal(dbg["accum"], ref["debug"]["accum"], equal_nan=True), trial
        assert np.allclose(hdr, ref["hdr"], rtol=2e-6, atol=0, equal_nan=True), trial
        prev = rng.random((h, w, 4), dtype=np.float32)
        assert np.array_equal(hostsim.reproject(new, old, tm, prev)[0], oracle_mod.reproject(new, old, tm, prev), equal_nan=True), trial
        ldr = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        hist = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        assert np.array_equal(hostsim.txaa(new, old, tm, ldr, hist), oracle_mod.txaa(new, old, tm, ldr, hist)), trial
    assert marched > 10000


@pytest.mark.parametrize("w,h,fid,yaw,pitch,weather", [(200, 112, 14, -30.0, 8.0, False), (130, 70, 7, 12.0, 3.0, False),
                                                     (96, 54, 3, 20.0, 2.0, True), (64, 36, 0, 0.0, -4.0, False)])
def test_step_parallel_decomposition_equals_sequential_march(oracle_mod, noise, hostsim, w, h, fid, yaw, pitch, weather):
    """The 1-of-16 dispatch runs as rays -> independent (ray, step) samples -> fold on the GPU (cloud_raymarch.cu).  The
    same decomposition with the same device functions on the host -- samples evaluated last step first -- gives the
    sequential march's bytes (on the GPU: test_sixteenth_step_parallel_equals_sequential)."""
    cam, tm, _, tun = default_scene(w, h, frame_id=fid, total_time=9.0, yaw=yaw, pitch=pitch)
    if weather:
        tun["use_weather"], tun["weather_scale"] = 1, 1.0e-4
    seq_hdr, seq_mask, _, _ = hostsim.cloud(cam, tm, tun, noise, w, h, 0, oracle_mod.RAY_DEBUG_DTYPE)
    par_hdr, par_mask, _, _ = hostsim.cloud(cam, tm, tun, noise, w, h, 2, oracle_mod.RAY_DEBUG_DTYPE)
    assert np.array_equal(seq_hdr, par_hdr) and np.array_equal(seq_mask, par_mask)
    assert seq_mask.any()


def test_tile_launch_order_is_a_permutation_that_interleaves_light_tiles(hostsim):
    """mt_tile_order (mt_params.h): every owned tile is issued exactly once; the marching tiles go from the horizon upwards
    (index heavy_first-1 down to 0) and the light tiles (sky band, ocean) are slotted in one for one, so that neither the
    heaviest CTAs nor the bulk of the stores end up in the kernel's tail."""
    lib = hostsim.lib()
    for n in list(range(1, 40)) + [135, 270, 540]:
        for m in sorted({0, 1, n // 3, n // 2, n - 1, n}):
            if not (0 <= m <= n):
                continue
            order = [lib.hs_tile_order(m, n, j) for j in range(n)]
            assert sorted(order) == list(range(n)), (m, n)
            heavy = [t for t in order if t < m]
            assert heavy == list(range(m - 1, -1, -1))                  # horizon first, zenith last
            light = [t for t in order if t >= m]
            assert light == list(range(m, n))                           # sky band, then ocean, top-down
            if m and n - m:
                first = order[: 2 * min(m, n - m)]
                assert all((t < m) == (k % 2 == 0) for k, t in enumerate(first))  # strictly alternating while both last
