"""Parity of the sm_100a kernels against the CPU oracle, through the C ABI (needs a B200: run with -m gpu).

Bars (BASELINE.json north_star): pixel selection and reprojection tap indices bit-exact; HDR radiance max relative
error <= 1e-3 and PSNR >= 50 dB; god-ray mask within 1/255.  The kernels are designed for more than that -- every
decision-carrying value (ray, shell distances, step count, jitter sequence, accumulated density, mask) bit-identical,
radiance differing only through the SFU exp/pow -- and the tests assert the stronger property too.
"""
import numpy as np
import pytest

from conftest import default_scene, psnr, rel_err

pytestmark = pytest.mark.gpu

HDR_MAX_REL = 1e-3   # north_star tolerance
HDR_MIN_PSNR = 50.0  # dB
MASK_TOL = 1.0 / 255.0


@pytest.fixture(scope="module")
def api():
    from meteoros_b200 import api as _api

    return _api


def make_renderer(api, noise, w, h, **kw):
    r = api.CloudRenderer(w, h, **kw)
    r.upload_noise(noise)
    return r


def check_hdr(got, want, where=None):
    if where is not None:
        got, want = got[where], want[where]
    e = rel_err(got[..., :3], want[..., :3])
    assert e.max() <= HDR_MAX_REL, f"max rel err {e.max():.3e}"
    assert psnr(got[..., :3], want[..., :3]) >= HDR_MIN_PSNR
    return float(e.max())


@pytest.mark.parametrize("w,h,fid,yaw,pitch,t", [
    (240, 136, 1, 0.0, 0.0, 0.016),
    (240, 136, 6, 25.0, 5.0, 12.5),
    (130, 70, 15, -10.0, -3.0, 100.0),   # trailing columns 128,129 are never marched (Renderer.cpp:713)
    (1284, 720, 9, 0.0, 0.0, 0.5),       # the reference's own default window (main.cpp:22-23)
])
def test_cloud_sixteenth_dispatch(api, oracle_mod, noise, w, h, fid, yaw, pitch, t):
    cam, tm, _, tun = default_scene(w, h, frame_id=fid, total_time=t, yaw=yaw, pitch=pitch)
    sentinel = np.full((h, w, 4), -7.0, np.float32)
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=False, hdr=sentinel.copy(), mask=sentinel.copy())
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.write_image(api.IMAGE_CLOUD_CUR, sentinel)
        r.write_image(api.IMAGE_GODRAY_MASK, sentinel)
        r.dispatch_cloud()
        hdr = r.read_image(api.IMAGE_CLOUD_CUR)
        mask = r.read_image(api.IMAGE_GODRAY_MASK)
    written = hdr[..., 3] != -7.0
    assert np.array_equal(written, ref["hdr"][..., 3] != -7.0)      # pixel selection: bit-exact
    assert np.array_equal(hdr[~written], sentinel[~written])        # nothing else touched
    check_hdr(hdr, ref["hdr"], written)
    assert np.abs(mask - ref["mask"]).max() <= MASK_TOL
    assert np.array_equal(mask, ref["mask"])                        # designed to be exact


@pytest.mark.parametrize("w,h,yaw,pitch,t", [(480, 270, 0.0, 0.0, 0.016), (322, 182, 40.0, 10.0, 7.0)])
def test_cloud_full_dispatch_and_debug_records(api, oracle_mod, noise, w, h, yaw, pitch, t):
    cam, tm, _, tun = default_scene(w, h, frame_id=3, total_time=t, yaw=yaw, pitch=pitch)
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True, counters=True, debug=True)
    with make_renderer(api, noise, w, h, flags=api.FLAG_COUNTERS) as r:
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.dispatch_cloud_full()
        hdr = r.read_image(api.IMAGE_CLOUD_CUR)
        mask = r.read_image(api.IMAGE_GODRAY_MASK)
        cnt = r.counters()
        dbg = r.dispatch_cloud_debug(True)
        hdr2 = r.read_image(api.IMAGE_CLOUD_CUR)
    assert cnt == ref["counters"]
    for f in oracle_mod.RAY_DEBUG_DTYPE.names:  # ray, shells, step count, jitter sequence, accumulated density
        assert np.array_equal(dbg[f], ref["debug"][f]), f
    assert np.array_equal(mask, ref["mask"])
    e = check_hdr(hdr, ref["hdr"])
    assert e < 1e-4  # SFU exp/pow only
    assert np.array_equal(hdr, hdr2)  # debug / counter variants compute the same pixels


@pytest.mark.parametrize("w,h", [(1920, 1080), (130, 70), (500, 281)])
def test_sixteenth_step_parallel_equals_sequential(api, noise, w, h):
    """The 1-of-16 dispatch runs as rays -> (ray, step) samples -> fold by default; MT_FLAG_SEQUENTIAL_MARCH keeps the
    one-thread-per-ray kernel.  Same device functions, same order: the images must be bit-identical."""
    sentinel = np.full((h, w, 4), -7.0, np.float32)
    out = {}
    for flags in (0, api.FLAG_SEQUENTIAL_MARCH):
        with make_renderer(api, noise, w, h, flags=flags) as r:
            imgs = []
            for fid in (0, 7, 13):
                cam, tm, _, tun = default_scene(w, h, frame_id=fid, total_time=3.0 + fid, yaw=2.0 * fid)
                r.set_camera(cam); r.set_time(tm)
                r.write_image(api.IMAGE_CLOUD_CUR, sentinel)
                r.write_image(api.IMAGE_GODRAY_MASK, sentinel)
                r.dispatch_cloud()
                imgs.append((r.read_image(api.IMAGE_CLOUD_CUR), r.read_image(api.IMAGE_GODRAY_MASK)))
            out[flags] = imgs
    for (h0, m0), (h1, m1) in zip(out[0], out[api.FLAG_SEQUENTIAL_MARCH]):
        assert np.array_equal(h0, h1) and np.array_equal(m0, m1)


def test_randomised_scenes(api, oracle_mod, noise):
    """Twelve random scenes (camera position / orientation / field of view, time, coverage, sun, frame id), both
    dispatch modes: decisions bit-exact, radiance within tolerance -- the parity does not hinge on the default view."""
    from meteoros_b200 import scene

    rng = np.random.default_rng(2024)
    w, h = 112, 63
    with make_renderer(api, noise, w, h) as r:
        for k in range(12):
            eye = (float(rng.uniform(-500, 500)), float(rng.uniform(-3000, 50)), float(rng.uniform(-500, 500)))
            cam = scene.Camera(w, h, eye=eye, ref=(eye[0], eye[1], eye[2] - 1.0), fovy=float(rng.uniform(25, 80)))
            cam.rotate_about_up(float(rng.uniform(-180, 180)))
            cam.rotate_about_right(float(rng.uniform(-10, 60)))
            sc = scene.Scene()
            sc.time["time"] = (0.016, float(rng.uniform(0, 500)))
            sc.time["frameCountMod16"] = int(rng.integers(0, 16))
            tun = scene.default_tuning()
            tun["coverage"] = float(rng.uniform(0.3, 0.85))
            tun["sun_location"] = scene.sun_on_elevation_circle(float(rng.uniform(5, 85)))
            tun["cloud_speed"] = float(rng.uniform(0.0, 0.2))
            c, t = cam.ubo(), sc.ubo()
            r.set_camera(c); r.set_time(t); r.set_tuning(tun)
            ref = oracle_mod.cloud(c, t, tun, noise, w, h, full=True, debug=True)
            dbg = r.dispatch_cloud_debug(True)
            for f in oracle_mod.RAY_DEBUG_DTYPE.names:
                assert np.array_equal(dbg[f], ref["debug"][f]), (k, f)
            assert np.array_equal(r.read_image(api.IMAGE_GODRAY_MASK), ref["mask"]), k
            check_hdr(r.read_image(api.IMAGE_CLOUD_CUR), ref["hdr"])
            sentinel = np.full((h, w, 4), -3.0, np.float32)
            ref16 = oracle_mod.cloud(c, t, tun, noise, w, h, full=False, hdr=sentinel.copy(), mask=sentinel.copy())
            r.write_image(api.IMAGE_CLOUD_CUR, sentinel)
            r.write_image(api.IMAGE_GODRAY_MASK, sentinel)
            r.dispatch_cloud()  # step-parallel path
            got = r.read_image(api.IMAGE_CLOUD_CUR)
            wr = got[..., 3] != -3.0
            assert np.array_equal(wr, ref16["hdr"][..., 3] != -3.0)
            assert np.array_equal(r.read_image(api.IMAGE_GODRAY_MASK), ref16["mask"]), k
            check_hdr(got, ref16["hdr"], wr)


@pytest.mark.parametrize("eye_y", [-7400.0, -9000.0, -25000.0])
def test_degenerate_cameras(api, oracle_mod, noise, eye_y):
    """Cameras the reference never meets (ray origin = -eye just under the cloud base, inside the cloud layer, above the
    outer shell): shell hits become invalid (t = 0, point = 0) and the march starts at the eye.  The kernels must still
    agree with the oracle, NaNs included."""
    from meteoros_b200 import scene

    w, h = 96, 54
    cam = scene.Camera(w, h, eye=(0.0, eye_y, 2.0), ref=(0.0, eye_y, 1.0))
    cam.rotate_about_right(15.0)
    sc = scene.Scene()
    sc.update_time(0.5)
    tun = scene.default_tuning()
    c, t = cam.ubo(), sc.ubo()
    ref = oracle_mod.cloud(c, t, tun, noise, w, h, full=True, debug=True)
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(c); r.set_time(t)
        dbg = r.dispatch_cloud_debug(True)
        hdr = r.read_image(api.IMAGE_CLOUD_CUR)
        mask = r.read_image(api.IMAGE_GODRAY_MASK)
    for f in ("branch", "steps", "jitter_hash", "t_in", "t_out", "step_size"):
        assert np.array_equal(dbg[f], ref["debug"][f], equal_nan=(dbg[f].dtype.kind == "f")), f
    assert np.array_equal(np.isnan(hdr), np.isnan(ref["hdr"]))
    ok = ~np.isnan(ref["hdr"])
    assert np.allclose(hdr[ok], ref["hdr"][ok], rtol=1e-3, atol=1e-6)
    assert np.array_equal(np.isnan(mask), np.isnan(ref["mask"]))
    assert np.array_equal(mask[~np.isnan(mask)], ref["mask"][~np.isnan(mask)])


def test_cloud_tuning_sweep(api, oracle_mod, noise):
    from meteoros_b200 import scene

    w, h = 160, 90
    cam, tm, _, tun = default_scene(w, h, frame_id=11, total_time=33.0)
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(cam); r.set_time(tm)
        for cov, elev in ((0.3, 5.0), (0.9, 85.0), (0.5, 45.0)):
            tun["coverage"] = cov
            tun["sun_location"] = scene.sun_on_elevation_circle(elev)
            r.set_tuning(tun)
            r.dispatch_cloud_full()
            hdr = r.read_image(api.IMAGE_CLOUD_CUR)
            mask = r.read_image(api.IMAGE_GODRAY_MASK)
            ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True)
            check_hdr(hdr, ref["hdr"])
            assert np.array_equal(mask, ref["mask"])


@pytest.mark.parametrize("scale", [1.0e-4, 3.7e-5])
def test_cloud_weather_path(api, oracle_mod, noise, scale):
    """MtTuning.use_weather (cloudRayMarch.comp:515-525 restored): per-sample coverage and height gradient from the
    weather map.  Same bars; both dispatch shapes (one launch per ray / step-parallel 1-of-16) go through it."""
    w, h = 322, 182
    cam, tm, _, tun = default_scene(w, h, frame_id=4, total_time=21.0, yaw=15.0, pitch=4.0)
    tun["use_weather"], tun["weather_scale"] = 1, scale
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True, debug=True)
    assert (ref["debug"]["accum"] > 0).mean() > 0.02
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.dispatch_cloud_full()
        hdr = r.read_image(api.IMAGE_CLOUD_CUR)
        mask = r.read_image(api.IMAGE_GODRAY_MASK)
        check_hdr(hdr, ref["hdr"])
        assert np.array_equal(mask, ref["mask"])
        assert np.array_equal(hdr[..., 3], ref["hdr"][..., 3])       # accumulated density: bit-exact
        # 1-of-16: step-parallel and sequential marches agree bit for bit, and with the oracle's selected pixels
        r.clear_images(); r.dispatch_cloud()
        step_par = r.read_image(api.IMAGE_CLOUD_CUR)
        sel = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=False, hdr=np.zeros((h, w, 4), np.float32))
        assert np.array_equal(step_par[..., 3], sel["hdr"][..., 3])
        check_hdr(step_par, sel["hdr"])
    with make_renderer(api, noise, w, h, flags=api.FLAG_SEQUENTIAL_MARCH) as r:
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.dispatch_cloud()
        assert np.array_equal(r.read_image(api.IMAGE_CLOUD_CUR), step_par)
        with pytest.raises(api.MeteorosError):
            r.dispatch_cloud_debug(False) # debug records / counters are not built for the weather variant
    with api.CloudRenderer(w, h) as r:    # weather requested but never uploaded
        r.upload_texture_3d(api.TEX_LOW_FREQ, noise["low"]); r.upload_texture_3d(api.TEX_HIGH_FREQ, noise["high"])
        r.upload_texture_2d(api.TEX_CURL, noise["curl"])
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        with pytest.raises(api.MeteorosError) as e:
            r.dispatch_cloud()
        assert e.value.status == 5 and "weather" in str(e.value)


def test_cloud_row_tiles_are_bit_identical_to_one_launch(api, noise):
    """Multi-GPU sharding must not change arithmetic: N tile launches == one launch (run here on one device)."""
    w, h = 320, 200  # 200 rows: the last 32-row tile is partial
    cam, tm, _, tun = default_scene(w, h, yaw=15.0)
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(cam); r.set_time(tm)
        r.dispatch_cloud_full()
        full_hdr = r.read_image(api.IMAGE_CLOUD_CUR)
        full_mask = r.read_image(api.IMAGE_GODRAY_MASK)
        for world in (2, 3, 8):
            r.clear_images()
            n = (h + 31) // 32
            for rank in range(world):
                r.dispatch_cloud_tiles(32, rank, n, world)
            assert np.array_equal(r.read_image(api.IMAGE_CLOUD_CUR), full_hdr)
            assert np.array_equal(r.read_image(api.IMAGE_GODRAY_MASK), full_mask)
        # a single rank's shard leaves the other tiles untouched
        r.clear_images()
        r.dispatch_cloud_tiles(32, 1, n, 2)
        part = r.read_image(api.IMAGE_CLOUD_CUR)
        for t in range(n):
            rows = slice(t * 32, min(h, (t + 1) * 32))
            if t % 2 == 1:
                assert np.array_equal(part[rows], full_hdr[rows])
            else:
                assert not part[rows].any()


def test_cloud_output_redirect_and_ipc_roundtrip(api, noise):
    """mtSetCloudOutput: the kernel stores into caller-provided device memory (the peer-mapped image on GPU 0 in a
    multi-GPU run; here a second context's image on the same device)."""
    w, h = 128, 72
    cam, tm, _, tun = default_scene(w, h)
    with make_renderer(api, noise, w, h) as a, make_renderer(api, noise, w, h) as b:
        for r in (a, b):
            r.set_camera(cam); r.set_time(tm)
        a.dispatch_cloud_full()
        want = a.read_image(api.IMAGE_CLOUD_CUR)
        b.set_cloud_output(a.image_device_ptr(api.IMAGE_CLOUD_PREV), None)
        b.dispatch_cloud_full()
        b.synchronize()
        assert np.array_equal(a.read_image(api.IMAGE_CLOUD_PREV), want)
        assert not b.read_image(api.IMAGE_CLOUD_CUR).any()
        assert len(a.export_image_handle(api.IMAGE_CLOUD_CUR)) == 64


@pytest.mark.parametrize("w,h", [(480, 270), (322, 182), (1920, 1080)])
def test_cloud_bulk_store_mode_is_bit_identical(api, noise, w, h):
    """mtSetCloudStoreMode(MT_STORE_BULK): a warp's pixels leave through shared memory and cp.async.bulk copies -- into the
    context's own image and into a "peer" image (a second context on this device) -- and are the bytes of the direct stores,
    also where a tile is partial (322 = 20 * 16 + 2 columns, 182 = 22 * 8 + 6 rows) and for row-tile shards."""
    cam, tm, _, tun = default_scene(w, h, frame_id=5, total_time=2.0, yaw=-8.0)
    with make_renderer(api, noise, w, h) as a, make_renderer(api, noise, w, h) as b:
        for r in (a, b):
            r.set_camera(cam); r.set_time(tm)
        a.dispatch_cloud_full()
        want, want_mask = a.read_image(api.IMAGE_CLOUD_CUR), a.read_image(api.IMAGE_GODRAY_MASK)
        b.set_cloud_store_mode(api.STORE_BULK)
        b.dispatch_cloud_full()
        assert np.array_equal(b.read_image(api.IMAGE_CLOUD_CUR), want)
        assert np.array_equal(b.read_image(api.IMAGE_GODRAY_MASK), want_mask)
        b.clear_images()
        b.set_cloud_output(a.image_device_ptr(api.IMAGE_CLOUD_PREV), None)
        n = (h + 7) // 8
        for rank in range(3):
            b.dispatch_cloud_tiles(8, rank, n, 3)
        b.synchronize()
        assert np.array_equal(a.read_image(api.IMAGE_CLOUD_PREV), want)
        assert not b.read_image(api.IMAGE_CLOUD_CUR).any()
        b.set_cloud_output(None, None)
        b.set_cloud_store_mode(api.STORE_DIRECT)


@pytest.mark.parametrize("w,h,tile_rows", [(480, 270, 8), (322, 182, 16), (1920, 1080, 8)])
def test_cloud_forward_pushes_finished_tiles(api, noise, w, h, tile_rows):
    """mtSetCloudForward: the march kernel stores locally and counts finished CTAs per row tile; the side kernel pushes each
    finished tile to the "peer" image (here a second context's image on the same device).  Three frames in a row, then
    every other tile only: the destination receives exactly the dispatched tiles, bit-identical to one full launch."""
    from meteoros_b200 import sharding

    cam, tm, _, tun = default_scene(w, h, frame_id=2, total_time=4.0, yaw=5.0)
    n = sharding.num_tiles(h, tile_rows)
    with make_renderer(api, noise, w, h) as a, make_renderer(api, noise, w, h) as b:
        for r in (a, b):
            r.set_camera(cam); r.set_time(tm)
        a.dispatch_cloud_full()
        want = a.read_image(api.IMAGE_CLOUD_CUR)
        b.set_cloud_forward(a.image_device_ptr(api.IMAGE_CLOUD_PREV))
        for _ in range(3):
            b.dispatch_cloud_tiles(tile_rows, 0, n, 1)
        b.join_copies()
        b.synchronize()
        assert np.array_equal(a.read_image(api.IMAGE_CLOUD_PREV), want)
        assert np.array_equal(b.read_image(api.IMAGE_CLOUD_CUR), want)      # the stores stayed local as well
        a.clear_images(); b.clear_images()
        b.dispatch_cloud_tiles(tile_rows, 1, n, 2)                          # odd tiles only
        b.synchronize()
        got = a.read_image(api.IMAGE_CLOUD_PREV)
        for t in range(n):
            r0, r1 = sharding.rows_of_tile(h, tile_rows, t)
            assert np.array_equal(got[r0:r1], want[r0:r1]) if t % 2 else not got[r0:r1].any()
        b.set_cloud_forward(None)
        a.clear_images()
        b.dispatch_cloud_tiles(tile_rows, 0, n, 1)
        b.synchronize()
        assert not a.read_image(api.IMAGE_CLOUD_PREV).any()                 # off again: nothing leaves the context


def test_reprojection_indices_and_image(api, oracle_mod):
    from meteoros_b200 import scene

    w, h = 400, 226
    rng = np.random.default_rng(5)
    prev = rng.random((h, w, 4), dtype=np.float32)
    cam = scene.Camera(w, h)
    old = cam.ubo()
    cam.rotate_about_up(0.25)
    cam.rotate_about_right(0.25)
    new = cam.ubo()
    sc = scene.Scene()
    with api.CloudRenderer(w, h) as r:
        for fid in (1, 10):
            sc.time["frameCountMod16"] = fid
            r.set_camera(new); r.set_camera_old(old); r.set_time(sc.ubo())
            r.write_image(api.IMAGE_CLOUD_PREV, prev)
            taps = r.dispatch_reprojection_debug()
            cur = r.read_image(api.IMAGE_CLOUD_CUR)
            ref, ref_taps = oracle_mod.reproject(new, old, sc.ubo(), prev, taps=True)
            assert np.array_equal(taps, ref_taps)   # reprojection indices: bit-exact
            assert np.array_equal(cur, ref)         # ten exact adds and one divide
            r.dispatch_reprojection()
            assert np.array_equal(r.read_image(api.IMAGE_CLOUD_CUR), ref)


def test_godrays_and_tonemap(api, oracle_mod, noise):
    w, h = 256, 144
    cam, tm, sky, tun = default_scene(w, h, total_time=9.75)
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True)
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(cam); r.set_time(tm); r.set_sun_and_sky(sky)
        # feed the oracle's images so each pass is judged on identical inputs
        r.write_image(api.IMAGE_CLOUD_CUR, ref["hdr"])
        r.write_image(api.IMAGE_GODRAY_MASK, ref["mask"])
        r.dispatch_god_rays()
        got = r.read_image(api.IMAGE_CLOUD_CUR)
        want = oracle_mod.godrays(cam, sky, ref["mask"], ref["hdr"])
        assert (want != ref["hdr"]).any()
        # decode-then-filter vs the shader's filter-then-decode: rounding-level difference on the added term only
        assert rel_err(got, want).max() <= 1e-6
        base = ref["hdr"].astype(np.float64)
        added_err = np.abs((got - base) - (want - base))
        assert (added_err <= 2e-5 * np.abs(want - base) + 2.5e-7 * np.abs(base)).all()  # + one ulp of the HDR value
        r.dispatch_tone_map()
        ldr = r.read_image(api.IMAGE_LDR)
        want_ldr = oracle_mod.tonemap(tm, got)
        d = np.abs(ldr.astype(np.int32) - want_ldr.astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 0.01  # SFU pow can move a value across a rounding boundary
        assert (ldr[..., 3] == 255).all()
    # sun behind the camera: the god-ray pass writes nothing
    from meteoros_b200 import scene

    c2 = scene.Camera(w, h)
    c2.rotate_about_right(-80.0)
    with api.CloudRenderer(w, h) as r:
        r.set_camera(c2.ubo()); r.set_sun_and_sky(sky)
        r.write_image(api.IMAGE_CLOUD_CUR, ref["hdr"])
        r.write_image(api.IMAGE_GODRAY_MASK, ref["mask"])
        r.dispatch_god_rays()
        out = r.read_image(api.IMAGE_CLOUD_CUR)
        assert np.array_equal(out, oracle_mod.godrays(c2.ubo(), sky, ref["mask"], ref["hdr"]))
        assert np.array_equal(out, ref["hdr"])


@pytest.mark.parametrize("w,h", [(160, 90), (33, 17)])
def test_godrays_sun_position_sweep(api, oracle_mod, w, h):
    """The taps walk from the pixel centre to the clamped sun position and are not clamped themselves (they provably stay
    inside the decoded image's one-texel ring): sun inside the frame, on every edge and corner, behind the camera."""
    from meteoros_b200 import scene

    rng = np.random.default_rng(w)
    hdr = rng.random((h, w, 4), dtype=np.float32)
    mask = rng.random((h, w, 4), dtype=np.float32)      # every texel non-trivial, incl. the ones next to the ring
    sky = scene.Sky().ubo()
    lit = 0
    with api.CloudRenderer(w, h) as r:
        r.set_sun_and_sky(sky)
        for yaw in (0.0, 35.0, -35.0, 90.0, 179.0):
            for pitch in (0.0, 30.0, 60.0, 89.0, -30.0):
                c = scene.Camera(w, h)
                c.rotate_about_up(yaw)
                if pitch:
                    c.rotate_about_right(pitch)
                cam = c.ubo()
                r.set_camera(cam)
                r.write_image(api.IMAGE_CLOUD_CUR, hdr)
                r.write_image(api.IMAGE_GODRAY_MASK, mask)
                r.dispatch_god_rays()
                got = r.read_image(api.IMAGE_CLOUD_CUR)
                want = oracle_mod.godrays(cam, sky, mask, hdr)
                lit += int((want != hdr).any())
                added_err = np.abs((got - hdr.astype(np.float64)) - (want - hdr.astype(np.float64)))
                assert (added_err <= 2e-5 * np.abs(want - hdr) + 2.5e-7 * np.abs(hdr)).all(), (yaw, pitch)
    assert 5 <= lit < 25        # both outcomes of the blend < 0 test occurred


def test_sixteen_frame_pan_sequence(api, oracle_mod, noise):
    """BASELINE config 2 at reduced size: 16 frames, 0.25 deg/frame pan, REPROJ + CLOUD + GODRAYS + TONEMAP + swap."""
    from meteoros_b200 import scene

    w, h = 192, 108
    cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
    tun = scene.default_tuning()
    img = [np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)]
    mask = np.zeros((h, w, 4), np.float32)
    cur = 0
    cam_old = cam.ubo()
    worst = 0.0
    with make_renderer(api, noise, w, h) as r:
        r.set_sun_and_sky(sky.ubo())
        for frame in range(16):
            cam.rotate_about_up(0.25)
            sc.update_time(1 / 60)
            c, t = cam.ubo(), sc.ubo()
            # oracle frame
            img[cur] = oracle_mod.reproject(c, cam_old, t, img[cur ^ 1])
            oracle_mod.cloud(c, t, tun, noise, w, h, full=False, hdr=img[cur], mask=mask)
            img[cur] = oracle_mod.godrays(c, sky.ubo(), mask, img[cur])
            ldr_ref = oracle_mod.tonemap(t, img[cur])
            # CUDA frame through mtFrame
            r.set_camera(c); r.set_camera_old(cam_old); r.set_time(t)
            r.frame(with_godrays=True)
            got = r.read_image(api.IMAGE_CLOUD_PREV)  # roles swapped at the end of the frame
            worst = max(worst, check_hdr(got, img[cur]))
            assert np.abs(r.read_image(api.IMAGE_GODRAY_MASK) - mask).max() <= MASK_TOL
            d = np.abs(r.read_image(api.IMAGE_LDR_PREV).astype(np.int32) - ldr_ref.astype(np.int32))  # LDR swaps too
            assert d.max() <= 1
            cur ^= 1
            cam_old = c
    assert worst < 1e-3
    g = np.load(__import__("pathlib").Path(__file__).parent / "golden" / "sequence_96x54.npz")
    assert g["ldr"].shape == (4, 54, 96, 4)  # the committed sequence golden is checked on CPU (test_golden_sequence)


def test_txaa_pass_and_reference_live_frame(api, oracle_mod, noise):
    """TXAA alone on identical inputs, then the reference's live frame REPROJ + CLOUD + TONEMAP + TXAA
    (god rays off, Renderer.cpp:826-846) over eight frames against the oracle."""
    from meteoros_b200 import scene

    w, h = 208, 117
    rng = np.random.default_rng(2)
    cur_ldr = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    prev_ldr = np.roll(cur_ldr, 1, axis=0)
    cam = scene.Camera(w, h)
    old = cam.ubo()
    cam.rotate_about_up(0.25)
    sc = scene.Scene()
    sc.update_time(1 / 60)
    with api.CloudRenderer(w, h) as r:
        r.set_camera(cam.ubo()); r.set_camera_old(old); r.set_time(sc.ubo())
        r.write_image(api.IMAGE_LDR, cur_ldr)
        r.write_image(api.IMAGE_LDR_PREV, prev_ldr)
        r.dispatch_txaa()
        got = r.read_image(api.IMAGE_LDR)
        assert np.array_equal(r.read_image(api.IMAGE_LDR_PREV), prev_ldr)
    want = oracle_mod.txaa(cam.ubo(), old, sc.ubo(), cur_ldr, prev_ldr)
    assert np.array_equal(got, want)  # no transcendental in this pass

    cam, sc = scene.Camera(w, h), scene.Scene()
    tun = scene.default_tuning()
    img = [np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)]
    ldr = [np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 4), np.uint8)]
    mask = np.zeros((h, w, 4), np.float32)
    cur, cam_old = 0, cam.ubo()
    with make_renderer(api, noise, w, h) as r:
        for frame in range(8):
            cam.rotate_about_up(0.25)
            sc.update_time(1 / 60)
            c, t = cam.ubo(), sc.ubo()
            img[cur] = oracle_mod.reproject(c, cam_old, t, img[cur ^ 1])
            oracle_mod.cloud(c, t, tun, noise, w, h, full=False, hdr=img[cur], mask=mask)
            ldr[cur] = oracle_mod.txaa(c, cam_old, t, oracle_mod.tonemap(t, img[cur]), ldr[cur ^ 1])
            r.set_camera(c); r.set_camera_old(cam_old); r.set_time(t)
            r.frame(with_godrays=False, with_txaa=True)
            got = r.read_image(api.IMAGE_LDR_PREV)  # roles swapped at the end of the frame
            d = np.abs(got.astype(np.int32) - ldr[cur].astype(np.int32))
            assert d.max() <= 2 and (d > 0).mean() < 0.02  # SFU pow in the tone map can move a value by one LSB
            cur ^= 1
            cam_old = c


def test_reference_shader_golden_live_sequence(api, noise):
    """The CUDA frame (mtFrameEx: REPROJ, CLOUD, GODRAYS, TONEMAP, TXAA, swap) against tests/golden/live_sequence_96x54.npz
    -- sixteen frames written by the REFERENCE'S OWN five shaders (compiled from their text where /root/reference
    exists; tests/golden/make_goldens.py).  No oracle involved: this is kernel vs reference shader output."""
    from pathlib import Path

    from meteoros_b200 import scene

    g = np.load(Path(__file__).parent / "golden" / "live_sequence_96x54.npz")
    assert "reference shaders" in str(g["source"])
    w, h = 96, 54
    cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
    cam_old = cam.ubo()
    with make_renderer(api, noise, w, h) as r:
        r.set_sun_and_sky(sky.ubo())
        for k in range(16):
            cam.rotate_about_up(0.25)
            sc.update_time(1 / 60)
            c, t = cam.ubo(), sc.ubo()
            r.set_camera(c); r.set_camera_old(cam_old); r.set_time(t)
            r.frame(with_godrays=True, with_txaa=True)
            hdr = r.read_image(api.IMAGE_CLOUD_PREV)            # roles swapped at the end of the frame
            aa = r.read_image(api.IMAGE_LDR_PREV)
            d = np.abs(aa.astype(np.int32) - g["txaa"][k].astype(np.int32))
            assert d.max() <= 2 and (d > 0).mean() < 0.02        # SFU exp / pow can move an 8-bit value by one LSB
            if k == 0:
                check_hdr(hdr, g["hdr_first"])
            cam_old = c
        check_hdr(hdr, g["hdr_last"])
        assert np.abs(r.read_image(api.IMAGE_GODRAY_MASK) - g["mask_last"]).max() <= MASK_TOL


def test_cxx_frame_driver_matches_python_driven_frames(api, noise):
    """mtxRunFrame (the reference main loop in C++) against the same frames driven from Python."""
    import ctypes as C

    from meteoros_b200 import _lib, scene

    lib = _lib.load()
    w, h = 160, 90
    passes = api.FRAME_TONEMAP | api.FRAME_TXAA
    with make_renderer(api, noise, w, h) as a, make_renderer(api, noise, w, h) as b:
        cam = _lib.MtxCamera()
        lib.mtxCameraInit(C.byref(cam), w, h, None, None, 45.0, 0.1, 1000.0)
        cam_old = np.zeros((), scene.CAMERA_DTYPE)
        lib.mtxCameraUBO(C.byref(cam), cam_old.ctypes.data)
        tm = np.zeros((), scene.TIME_DTYPE)
        lib.mtxTimeInit(tm.ctypes.data)
        py_old = cam_old.copy()
        py_tm = scene.Scene()
        for _ in range(5):
            lib.mtxCameraRotateAboutUp(C.byref(cam), 0.25)
            a._check(lib.mtxRunFrame(a._h, C.byref(cam), cam_old.ctypes.data, tm.ctypes.data, C.c_float(1 / 60), passes), "mtxRunFrame")
            cur = np.zeros((), scene.CAMERA_DTYPE)
            lib.mtxCameraUBO(C.byref(cam), cur.ctypes.data)
            py_tm.update_time(1 / 60)
            b.set_camera(cur); b.set_camera_old(py_old); b.set_time(py_tm.ubo()); b.set_sun_and_sky(scene.Sky().ubo())
            b.frame(with_godrays=False, with_txaa=True)
            py_old = cur
            assert cam_old.tobytes() == cur.tobytes() and tm.tobytes() == py_tm.ubo().tobytes()
            assert np.array_equal(a.read_image(api.IMAGE_LDR_PREV), b.read_image(api.IMAGE_LDR_PREV))
            assert np.array_equal(a.read_image(api.IMAGE_CLOUD_PREV), b.read_image(api.IMAGE_CLOUD_PREV))


def test_plain_c_host_runs_the_reference_frame_loop(api, noise, tmp_path):
    """examples/frame_loop.c -- the reference's main() on the C ABI, plain C11 -- built with gcc and RUN on the GPU: assets from
    .mtvol caches (written here from the committed noise fixture with mtxSaveVolume), 16 frames of REPROJ + CLOUD + TONEMAP +
    TXAA through mtxRunFrame, last presented frame dumped; it must equal the same 16 frames driven through the Python binding."""
    import ctypes as C
    import shutil
    import subprocess
    from pathlib import Path

    from meteoros_b200 import _lib, scene

    root = Path(__file__).resolve().parents[1]
    lib = _lib.load()
    for name, key in (("low", "low"), ("high", "high"), ("curl", "curl"), ("weather", "weather")):
        v = np.ascontiguousarray(noise[key])
        dims = v.shape[:3] if v.ndim == 4 else (1,) + v.shape[:2]
        assert lib.mtxSaveVolume(str(tmp_path / f"{name}.mtvol").encode(), dims[2], dims[1], dims[0], v.ctypes.data) == 0
    cc = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else shutil.which("gcc")
    exe, out = tmp_path / "frame_loop", tmp_path / "last.rgba8"
    r = subprocess.run([cc, "-std=c11", "-Wall", "-Werror", f"-I{root / 'include'}", str(root / "examples" / "frame_loop.c"),
                        f"-L{_lib.LIB_PATH.parent}", "-lmeteoros_b200", f"-Wl,-rpath,{_lib.LIB_PATH.parent}", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    w, h, frames = 480, 270, 16
    run = subprocess.run([str(exe), str(tmp_path), str(frames), str(w), str(h), str(out)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "assets loaded from .mtvol caches" in run.stdout and f"{frames} frames of {w}x{h} rendered" in run.stdout
    got = np.fromfile(out, np.uint8).reshape(h, w, 4)
    cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
    with make_renderer(api, noise, w, h) as b:
        old = cam.ubo()
        for _ in range(frames):
            cam.rotate_about_up(0.25)
            sc.update_time(1 / 60)
            b.set_camera(cam.ubo()); b.set_camera_old(old); b.set_time(sc.ubo()); b.set_sun_and_sky(sky.ubo())
            b.frame(with_godrays=False, with_txaa=True)
            old = cam.ubo()
        want = b.read_image(api.IMAGE_LDR_PREV)
    # the C++ producers (glm's sinf / cosf) and the Python mirror (one rounding of double sin / cos) agree to 4e-6 in the
    # camera, which can move a dithered 8-bit value by one step on a few pixels (test_cxx_frame_driver holds the C++ path exactly)
    dlt = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert dlt.max() <= 2 and (dlt > 0).mean() < 0.02
    assert got[..., :3].any() and (got[..., 3] > 0).all()


def test_hw_cone_filter_opt_in_mode(api, oracle_mod, noise):
    """MT_FLAG_HW_CONE_FILTER (opt-in, NOT the parity path): the six light-cone samples of a full-quality dispatch through a CUDA 3D
    texture object (LINEAR / REPEAT, the sampler state of Texture3D.cpp:92-134) -- the texture unit's 8-bit filter weights
    instead of the exact fp32 filter.  What the mode promises, and what this test holds it to: every decision-carrying value is
    still the oracle's bit for bit (the cone samples never feed the accumulated density: god-ray mask and alpha array_equal),
    the radiance stays within 5e-3 relative of the oracle's with PSNR >= 50 dB and all but a vanishing share of the pixels inside
    the 1e-3 bar of the default path -- for the full-quality dispatch and for the step-parallel 1-of-16 dispatch alike; the counting
    kernel does not take the mode (same bytes as without it); row-tile launches give the same image as one launch."""
    w, h = 1284, 720
    cam, tm, _, tun = default_scene(w, h, frame_id=3, total_time=2.5, yaw=12.0, pitch=4.0)
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True)
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.dispatch_cloud_full()
        exact = r.read_image(api.IMAGE_CLOUD_CUR)
        r.clear_images()
        r.dispatch_cloud()
        sixteenth = r.read_image(api.IMAGE_CLOUD_CUR)
        sixteenth_mask = r.read_image(api.IMAGE_GODRAY_MASK)
    with make_renderer(api, noise, w, h, flags=api.FLAG_HW_CONE_FILTER) as r:
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.dispatch_cloud_full()
        hdr = r.read_image(api.IMAGE_CLOUD_CUR)
        mask = r.read_image(api.IMAGE_GODRAY_MASK)
        r.clear_images()
        r.dispatch_cloud_tiles(8, 0, (h + 7) // 8, 2)
        r.dispatch_cloud_tiles(8, 1, (h + 7) // 8, 2)
        assert np.array_equal(r.read_image(api.IMAGE_CLOUD_CUR), hdr)
        r.clear_images()
        r.dispatch_cloud()  # the step-parallel 1-of-16 kernel takes the mode too: same decisions, radiance to the same bar
        hw16 = r.read_image(api.IMAGE_CLOUD_CUR)
        assert np.array_equal(r.read_image(api.IMAGE_GODRAY_MASK), sixteenth_mask) and np.array_equal(hw16[..., 3], sixteenth[..., 3])
        e16 = rel_err(hw16[..., :3], sixteenth[..., :3])
        assert (hw16 != sixteenth).any() and e16.max() <= 5e-3 and int((e16.max(axis=-1) > HDR_MAX_REL).sum()) <= w * h * 1e-4
    assert np.array_equal(mask, ref["mask"])
    assert np.array_equal(hdr[..., 3], ref["hdr"][..., 3])
    e = rel_err(hdr[..., :3], ref["hdr"][..., :3])
    over = int((e.max(axis=-1) > HDR_MAX_REL).sum())
    changed = int((hdr[..., :3] != exact[..., :3]).any(axis=-1).sum())
    print(f"hw cone filter: max rel err {e.max():.2e}, {over} of {w * h} pixels beyond {HDR_MAX_REL:g}, {changed} differ from the exact path, "
          f"PSNR {psnr(hdr[..., :3], ref['hdr'][..., :3]):.1f} dB")
    assert changed > 0, "the mode did not engage"
    assert e.max() <= 5e-3 and over <= w * h * 1e-4
    assert psnr(hdr[..., :3], ref["hdr"][..., :3]) >= HDR_MIN_PSNR
    with make_renderer(api, noise, w, h, flags=api.FLAG_HW_CONE_FILTER | api.FLAG_COUNTERS) as r:  # counting launches: the exact kernel
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.dispatch_cloud_full()
        assert np.array_equal(r.read_image(api.IMAGE_CLOUD_CUR), exact)


def test_f16_storage_emulation(api, oracle_mod, noise):
    w, h = 128, 72
    cam, tm, _, tun = default_scene(w, h)
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True)
    with make_renderer(api, noise, w, h, storage=api.STORAGE_F16_EMULATE) as r:
        r.set_camera(cam); r.set_time(tm)
        r.dispatch_cloud_full()
        hdr = r.read_image(api.IMAGE_CLOUD_CUR)
    want = ref["hdr"].astype(np.float16).astype(np.float32)  # R16G16B16A16_SFLOAT store (Renderer.cpp:1431)
    finite = np.isfinite(want)
    assert np.array_equal(hdr.astype(np.float16).astype(np.float32), hdr)  # every stored value is a binary16 value
    assert rel_err(hdr[finite], want[finite]).max() <= 2e-3               # at most one binary16 ulp apart


def test_f16_storage(api, oracle_mod, noise):
    """MT_STORAGE_F16: the HDR / mask images are RGBA16F in memory (the reference's VK_FORMAT_R16G16B16A16_SFLOAT,
    Renderer.cpp:1431-1440), 8 bytes per pixel through every pass, over mtRead/WriteImage, the bulk-store epilogue and the
    tile forwarder.  Bars: (1) the same values as MT_STORAGE_F16_EMULATE, bit for bit, through a 6-frame sequence of all
    passes; (2) the Cloud pass within one binary16 ulp of the oracle's frame rounded to binary16, mask likewise."""
    from meteoros_b200 import scene

    w, h = 320, 184
    cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
    frames = {}
    for storage in (api.STORAGE_F16_EMULATE, api.STORAGE_F16):
        cam, sc = scene.Camera(w, h), scene.Scene()
        out = []
        with make_renderer(api, noise, w, h, storage=storage) as r:
            assert r.read_image(api.IMAGE_CLOUD_CUR).nbytes == w * h * (8 if storage == api.STORAGE_F16 else 16)
            r.set_sun_and_sky(sky.ubo())
            old = cam.ubo()
            for f in range(6):
                cam.rotate_about_up(0.25)
                sc.update_time(1 / 60)
                r.set_camera(cam.ubo()); r.set_camera_old(old); r.set_time(sc.ubo())
                r.frame(with_godrays=True, with_txaa=True)
                old = cam.ubo()
                out.append((r.read_image(api.IMAGE_CLOUD_PREV).astype(np.float32), r.read_image(api.IMAGE_GODRAY_MASK).astype(np.float32),
                            r.read_image(api.IMAGE_LDR_PREV)))
            # full-quality dispatch: direct stores, the bulk-store epilogue, and the forwarder into a second context's image
            r.dispatch_cloud_full()
            full = r.read_image(api.IMAGE_CLOUD_CUR)
            r.set_cloud_store_mode(api.STORE_BULK)
            r.clear_images()
            r.dispatch_cloud_full()
            assert np.array_equal(r.read_image(api.IMAGE_CLOUD_CUR), full)
            r.set_cloud_store_mode(api.STORE_DIRECT)
            with make_renderer(api, noise, w, h, storage=storage) as peer:
                r.set_cloud_forward(peer.image_device_ptr(api.IMAGE_CLOUD_CUR))
                r.dispatch_cloud_tiles(8, 0, (h + 7) // 8, 1)
                r.join_copies(); r.synchronize()
                assert np.array_equal(peer.read_image(api.IMAGE_CLOUD_CUR), full)
                r.set_cloud_forward(None)
            out.append((full.astype(np.float32),))
        frames[storage] = out
    for a, b in zip(frames[api.STORAGE_F16_EMULATE], frames[api.STORAGE_F16]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    # against the oracle: its Cloud frame rounded to binary16, at most one binary16 ulp (2^-10 relative) apart
    tun = scene.default_tuning()
    ref = oracle_mod.cloud(cam.ubo(), sc.ubo(), tun, noise, w, h, full=True)
    with make_renderer(api, noise, w, h, storage=api.STORAGE_F16) as r:
        r.set_camera(cam.ubo()); r.set_time(sc.ubo())
        r.dispatch_cloud_full()
        hdr, mask = r.read_image(api.IMAGE_CLOUD_CUR), r.read_image(api.IMAGE_GODRAY_MASK)
    assert hdr.dtype == np.float16
    want = ref["hdr"].astype(np.float16).astype(np.float32)
    assert rel_err(hdr.astype(np.float32), want).max() <= 2.0 ** -10
    assert (hdr.astype(np.float32) != want).mean() < 1e-3          # in fact equal except where the fp32 values straddle a tie
    assert np.array_equal(mask, ref["mask"].astype(np.float16))      # the mask is bit-exact before rounding, hence after


def test_godray_grey_readback(api, oracle_mod, noise):
    """mtReadGodRayGreyAsync: the god-ray image as one float per pixel = the shader's own decode of the four encoded channels
    (postProcess_GodRays.frag:39-43), bit for bit, for both storage formats; a later Cloud dispatch does not disturb a read
    that is still in flight (it copies from a snapshot)."""
    w, h = 256, 144
    cam, tm, _, tun = default_scene(w, h, frame_id=4, total_time=6.0)
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True)["mask"]

    def decode(m):  # ((x * 1 + y / 255) + z / 65025) + w / 16581375, every operation in binary32
        m = m.astype(np.float32)
        k = [np.float32(1.0), np.float32(1.0) / np.float32(255.0), np.float32(1.0) / np.float32(65025.0), np.float32(1.0) / np.float32(16581375.0)]
        return ((m[..., 0] * k[0] + m[..., 1] * k[1]) + m[..., 2] * k[2]) + m[..., 3] * k[3]

    for storage in (api.STORAGE_F32, api.STORAGE_F16):
        with make_renderer(api, noise, w, h, storage=storage) as r:
            r.set_camera(cam); r.set_time(tm)
            r.dispatch_cloud_full()
            grey = r.read_godray_grey()
            want = decode(ref if storage == api.STORAGE_F32 else ref.astype(np.float16))
            assert grey.dtype == np.float32 and grey.shape == (h, w)
            assert np.array_equal(grey, want)
            assert np.array_equal(decode(r.read_image(api.IMAGE_GODRAY_MASK)), grey)
            out = np.empty((h, w), np.float32)
            r.read_godray_grey_async(out.ctypes.data, out.nbytes)
            r.clear_images()              # overwrites the mask while the read may still be in flight
            r.wait_reads()
            assert np.array_equal(out, want)


def test_async_readback_overlaps_next_frame(api, noise):
    """mtReadImageAsync: frame k is copied out on the copy stream while frame k+1 renders into the other ping-pong
    image; a third frame that re-uses the first image must wait for its copy."""
    import torch

    w, h = 512, 288
    views = [default_scene(w, h, yaw=10.0 * k) for k in range(3)]
    with make_renderer(api, noise, w, h) as r:
        want = []
        for cam, tm, _, _ in views:
            r.set_camera(cam); r.set_time(tm)
            r.dispatch_cloud_full()
            want.append(r.read_image(api.IMAGE_CLOUD_CUR))
        assert not np.array_equal(want[0], want[1])
        bufs = [torch.empty(w * h * 16, dtype=torch.uint8, pin_memory=True) for _ in range(3)]
        for k, (cam, tm, _, _) in enumerate(views):
            r.set_camera(cam); r.set_time(tm)
            r.dispatch_cloud_full()
            r.swap_ping_pong()
            r.read_image_async(api.IMAGE_CLOUD_PREV, bufs[k].data_ptr(), w * h * 16)
        r.wait_reads()
        for k in range(3):
            got = bufs[k].numpy().view(np.float32).reshape(h, w, 4)
            assert np.array_equal(got, want[k])


def test_error_paths(api, noise):
    with api.CloudRenderer(64, 36) as r:
        with pytest.raises(api.MeteorosError) as e:
            r.dispatch_cloud()            # no uniforms yet
        assert e.value.status == 5
        cam, tm, _, tun = default_scene(64, 36)
        r.set_camera(cam); r.set_time(tm)
        with pytest.raises(api.MeteorosError) as e:
            r.dispatch_cloud()            # no textures yet
        assert e.value.status == 5 and "textures" in str(e.value)
        with pytest.raises(api.MeteorosError):
            r.upload_texture_3d(api.TEX_LOW_FREQ, np.zeros((3, 3, 3, 4), np.uint8))  # not a power of two
        with pytest.raises(api.MeteorosError):
            r.upload_texture_3d(api.TEX_CURL, np.zeros((4, 4, 4, 4), np.uint8))      # wrong slot kind
        bad = tm.copy()
        bad["frameCountMod16"] = 16
        with pytest.raises(api.MeteorosError):
            r.set_time(bad)
        with pytest.raises(api.MeteorosError):
            r.dispatch_cloud_tiles(12, 0, 4, 1)
        r.upload_noise(noise)
        r.dispatch_cloud()
        r.synchronize()
        assert r.launch_count() >= 1


def test_resize(api, oracle_mod, noise):
    cam, tm, _, tun = default_scene(96, 54)
    with make_renderer(api, noise, 64, 36) as r:
        r.resize(96, 54)
        r.set_camera(cam); r.set_time(tm)
        r.dispatch_cloud_full()
        hdr = r.read_image(api.IMAGE_CLOUD_CUR)
    check_hdr(hdr, oracle_mod.cloud(cam, tm, tun, noise, 96, 54, full=True)["hdr"])


def test_full_size_properties_4k(api, noise):
    """BASELINE config 3 size (3840x2160), checked through size-independent properties: determinism, the union of
    the sixteen 1/16 dispatches equals the full dispatch bit for bit, row shards equal the full frame, counters obey
    their identities, and a band of rows matches the oracle."""
    w, h = 3840, 2160
    cam, tm, _, tun = default_scene(w, h)
    with make_renderer(api, noise, w, h, flags=api.FLAG_COUNTERS) as r:
        r.set_camera(cam); r.set_time(tm)
        r.dispatch_cloud_full()
        a = r.read_image(api.IMAGE_CLOUD_CUR)
        am = r.read_image(api.IMAGE_GODRAY_MASK)
        c = r.counters()
        assert c["rays"] == w * h
        assert c["steps"] >= 35 * (c["rays_marched"] - c["early_exits"]) and c["steps"] <= 60 * c["rays_marched"]
        assert c["steps_incloud"] <= c["steps"] and c["cone_hits"] <= 6 * c["steps_incloud"]
        assert 0.40 < c["rays_marched"] / c["rays"] < 0.46
        assert np.isfinite(a).all() and (a[..., 3] == 1.0).all()
        r.clear_images()
        r.dispatch_cloud_full()
        assert np.array_equal(r.read_image(api.IMAGE_CLOUD_CUR), a)  # deterministic
        r.clear_images()
        t = tm.copy()
        for fid in range(16):
            t["frameCountMod16"] = fid
            r.set_time(t)
            r.dispatch_cloud()
        assert np.array_equal(r.read_image(api.IMAGE_CLOUD_CUR), a)
        assert np.array_equal(r.read_image(api.IMAGE_GODRAY_MASK), am)
        r.set_time(tm)
        r.clear_images()
        n = (h + 31) // 32
        for rank in range(8):
            r.dispatch_cloud_tiles(32, rank, n, 8)
        assert np.array_equal(r.read_image(api.IMAGE_CLOUD_CUR), a)
    import oracle

    rows = (700, 716)  # a band in the marched half
    ref = oracle.cloud(cam, tm, tun, noise, w, h, full=True, rows=rows)
    check_hdr(a[rows[0]:rows[1]], ref["hdr"][rows[0]:rows[1]])
    assert np.array_equal(am[rows[0]:rows[1]], ref["mask"][rows[0]:rows[1]])


def test_committed_golden_frame(api, noise):  # the golden was written by the reference's own cloud shader (make_goldens.py)
    """The CUDA path against the committed golden (tests/golden/cloud_64x36.npz: written by the reference's own cloud shader compiled
    from its text, tests/golden/make_goldens.py, which refuses to write where the oracle disagrees with it)."""
    import pathlib

    g = np.load(pathlib.Path(__file__).parent / "golden" / "cloud_64x36.npz")
    w, h = 64, 36
    cam, tm, _, tun = default_scene(w, h, frame_id=int(g["frame_id"]), total_time=float(g["total_time"]), yaw=float(g["yaw"]))
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(cam); r.set_time(tm)
        dbg = r.dispatch_cloud_debug(True)
        hdr = r.read_image(api.IMAGE_CLOUD_CUR)
        mask = r.read_image(api.IMAGE_GODRAY_MASK)
    assert np.array_equal(dbg["steps"], g["steps"]) and np.array_equal(dbg["jitter_hash"], g["jitter_hash"])
    assert np.array_equal(dbg["accum"], g["accum"]) and np.array_equal(mask, g["mask"])
    check_hdr(hdr, g["hdr"])


def test_multi_gpu_sharded_frame_is_bit_identical():
    """Two or more B200s: the row-tile sharded frame gathered on rank 0 (peer stores and copy-engine pushes) equals the
    single-GPU frame bit for bit (tools/check_sharded.py under torchrun).  Skipped on a single-GPU box."""
    import subprocess
    import sys
    from pathlib import Path

    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
                        "--master-addr", "127.0.0.1", "--master-port", "29877", str(root / "tools" / "check_sharded.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    # peer_store, bulk_store and copy with and without the mask, forward without: seven gathered frames, all identical
    assert r.stdout.count("bit-identical to single-GPU: True") == 7 and "single-GPU: False" not in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("case", ["default", "origin_above_inner_shell", "wide_fov", "quarter_turn", "old_eye_elsewhere"])
def test_post_passes_fast_and_generic_paths(api, oracle_mod, case):
    """The reprojection / TXAA kernels exist twice: fast-path-only division / square root for frames whose constants the host has
    checked (mt_post_nice_ok) and the IEEE sequences for everything else; inside both, a tap takes the clamp-free path only while
    old_uv is inside the image.  Cameras on either side of every one of those conditions: tap indices, reprojected image and
    TXAA bytes stay bit-identical to the oracle."""
    from meteoros_b200 import scene

    w, h = 208, 117
    kw, yaw, pitch = {}, 0.25, 0.25
    old_kw = None
    if case == "origin_above_inner_shell":
        kw = dict(eye=(0.0, -9000.0, 2.0), ref=(0.0, -9000.0, 1.0))
    elif case == "wide_fov":
        kw = dict(fovy=179.92)                      # tan(fov / 2) > 1e3
    elif case == "quarter_turn":
        yaw, pitch = 88.0, 20.0                     # -q.z crosses zero inside the frame: taps far outside the image, per-pixel IEEE path
    elif case == "old_eye_elsewhere":
        old_kw = dict(eye=(300.0, -7450.0, 2.0), ref=(300.0, -7450.0, 1.0))   # previous eye within 100 m of the inner shell
    cam = scene.Camera(w, h, **kw)
    old = scene.Camera(w, h, **(old_kw or kw)).ubo()
    cam.rotate_about_up(yaw)
    cam.rotate_about_right(pitch)
    new = cam.ubo()
    sc = scene.Scene()
    sc.update_time(1 / 60)
    rng = np.random.default_rng(17)
    prev = rng.random((h, w, 4), dtype=np.float32)
    cur_ldr = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    prev_ldr = np.roll(cur_ldr, 3, axis=1)
    with api.CloudRenderer(w, h) as r:
        for fid in (3, 12):
            sc.time["frameCountMod16"] = fid
            tm = sc.ubo()
            r.set_camera(new); r.set_camera_old(old); r.set_time(tm)
            r.write_image(api.IMAGE_CLOUD_PREV, prev)
            taps = r.dispatch_reprojection_debug()
            ref, ref_taps = oracle_mod.reproject(new, old, tm, prev, taps=True)
            assert np.array_equal(taps, ref_taps), (case, fid)
            assert np.array_equal(r.read_image(api.IMAGE_CLOUD_CUR), ref)
            r.write_image(api.IMAGE_LDR, cur_ldr)
            r.write_image(api.IMAGE_LDR_PREV, prev_ldr)
            r.dispatch_txaa()
            assert np.array_equal(r.read_image(api.IMAGE_LDR), oracle_mod.txaa(new, old, tm, cur_ldr, prev_ldr)), (case, fid)


def test_cloud_generic_kernels_when_texture_coordinates_are_large(api, oracle_mod, noise):
    """The STD march kernels floor their filter coordinates with the magic constant (exact for |u| < 2^22); a frame whose constants
    do not bound the coordinates (here: a wind drift of ~2e4 texture periods) must run the generic kernels -- and still match."""
    from meteoros_b200 import scene

    w, h = 160, 90
    cam, tm, _, tun = default_scene(w, h, frame_id=5, total_time=100.0, yaw=15.0)
    tun["cloud_speed"] = 200.0
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True, counters=True)
    with make_renderer(api, noise, w, h, flags=api.FLAG_COUNTERS) as r:
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.dispatch_cloud_full()
        assert r.counters() == ref["counters"]
        assert np.array_equal(r.read_image(api.IMAGE_GODRAY_MASK), ref["mask"])
        check_hdr(r.read_image(api.IMAGE_CLOUD_CUR), ref["hdr"])
    sentinel = np.full((h, w, 4), -3.0, np.float32)
    ref16 = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=False, hdr=sentinel.copy(), mask=sentinel.copy())
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.write_image(api.IMAGE_CLOUD_CUR, sentinel)
        r.write_image(api.IMAGE_GODRAY_MASK, sentinel)
        r.dispatch_cloud()
        got = r.read_image(api.IMAGE_CLOUD_CUR)
        wr = got[..., 3] != -3.0
        assert np.array_equal(wr, ref16["hdr"][..., 3] != -3.0)
        assert np.array_equal(r.read_image(api.IMAGE_GODRAY_MASK), ref16["mask"])
        check_hdr(got, ref16["hdr"], wr)


@pytest.mark.parametrize("storage", ["f32", "f16"])
def test_incremental_mask_decode_equals_full_decode(api, noise, storage):
    """The fused 1-of-16 Cloud kernel keeps the god-ray pass's decoded copy of the mask current (it rewrites the pairs of the texels
    it stores), so a frame's god-ray dispatch skips mask_decode_kernel.  A context whose mask pointer has been handed out decodes
    the whole mask every time.  Twelve frames of the reference loop -- with a host write into the mask and a full-quality dispatch
    in between, both of which must invalidate the copy -- have to come out bit-identical from both."""
    from meteoros_b200 import scene

    w, h = 256, 144
    st = api.STORAGE_F16 if storage == "f16" else api.STORAGE_F32
    outs, launches = [], []
    for shared in (False, True):
        cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
        frames = []
        with api.CloudRenderer(w, h, storage=st) as r:
            r.upload_noise(noise)
            r.set_sun_and_sky(sky.ubo())
            if shared:
                assert r.image_device_ptr(api.IMAGE_GODRAY_MASK) != 0   # from here on the library cannot track writes to the mask
            old = cam.ubo()
            n0 = r.launch_count()
            for f in range(12):
                cam.rotate_about_up(0.5)
                sc.update_time(1 / 60)
                r.set_camera(cam.ubo()); r.set_camera_old(old); r.set_time(sc.ubo())
                if f == 5:     # a host write into the mask: the decoded copy is stale
                    m = r.read_image(api.IMAGE_GODRAY_MASK)
                    m[h // 3: h // 2] *= 0.5
                    r.write_image(api.IMAGE_GODRAY_MASK, m)
                if f == 8:     # a full-quality dispatch rewrites every mask texel without touching the copy
                    r.dispatch_cloud_full()
                r.frame(True, False)
                frames.append((r.read_image(api.IMAGE_CLOUD_PREV), r.read_image(api.IMAGE_LDR_PREV), r.read_image(api.IMAGE_GODRAY_MASK)))
                old = cam.ubo()
            launches.append(r.launch_count() - n0)
        outs.append(frames)
    for (h0, l0, m0), (h1, l1, m1) in zip(*outs):
        assert np.array_equal(m0, m1) and np.array_equal(h0, h1, equal_nan=True) and np.array_equal(l0, l1)
    assert launches[0] < launches[1]   # the tracked context skipped decode launches (9 of 12 frames)
