"""Full-size parity: every BASELINE.json configuration at ITS stated size, CUDA (through the C ABI) against the CPU oracle.

  config 1  whole 1920x1080 frame: all sixteen 1-of-16 dispatches (pixel selection per id, union) and the full dispatch
  config 2  the 16-frame 0.25 deg/frame pan at 1920x1080 through mtFrameEx (REPROJ + CLOUD 1/16 + GODRAYS + TONEMAP + swap)
  config 3  the WHOLE 3840x2160 full-quality frame
  config 4  a 7680x4320 frame rendered as eight cyclic row-tile shards, against the oracle on bands spread over the marched half
  config 5  eight views of the 256-view sun-elevation x coverage sweep at 1920x1080, incl. the extreme coverages 0.3 and 0.9

Bars (north_star): pixel selection / mask / alpha bit-exact, HDR max relative error <= 1e-3 and PSNR >= 50 dB.  The oracle
renders ~1.5 Mrays/s on the GPU box's cores: the whole file is about a minute of CPU time.  Grid: Renderer.cpp:713-716;
stores: cloudRayMarch.comp:690-826.
"""
import numpy as np
import pytest

from conftest import default_scene, psnr, rel_err

pytestmark = pytest.mark.gpu

HDR_MAX_REL = 1e-3
HDR_MIN_PSNR = 50.0


@pytest.fixture(scope="module")
def api():
    from meteoros_b200 import api as _api

    return _api


def make_renderer(api, noise, w, h, **kw):
    r = api.CloudRenderer(w, h, **kw)
    r.upload_noise(noise)
    return r


def check_frame(hdr, mask, ref, what):
    """mask + alpha bit-exact, radiance within the north_star bar; returns the worst relative error."""
    assert np.array_equal(mask, ref["mask"]), f"{what}: god-ray mask differs from the oracle"
    assert np.array_equal(hdr[..., 3], ref["hdr"][..., 3]), f"{what}: alpha differs from the oracle"
    e = rel_err(hdr[..., :3], ref["hdr"][..., :3])
    assert e.max() <= HDR_MAX_REL, f"{what}: HDR max rel err {e.max():.3e} at {np.unravel_index(e.argmax(), e.shape)}"
    p = psnr(hdr[..., :3], ref["hdr"][..., :3])
    assert p >= HDR_MIN_PSNR, f"{what}: PSNR {p:.1f} dB"
    return float(e.max())


def test_config1_whole_1080p_frame_all_sixteen_ids_and_full_dispatch(api, oracle_mod, noise):
    from meteoros_b200 import scene

    w, h = 1920, 1080
    cam, tm, _, tun = default_scene(w, h, frame_id=1, total_time=0.016)
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True)
    sentinel = np.full((h, w, 4), -7.0, np.float32)
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(cam); r.set_tuning(tun)
        r.set_time(tm)
        r.dispatch_cloud_full()
        check_frame(r.read_image(api.IMAGE_CLOUD_CUR), r.read_image(api.IMAGE_GODRAY_MASK), ref, "1080p full dispatch")
        # the sixteen 1-of-16 dispatches of the reference's real-time mode: each writes exactly its own pixels ...
        r.write_image(api.IMAGE_CLOUD_CUR, sentinel)
        r.write_image(api.IMAGE_GODRAY_MASK, sentinel)
        seen = np.zeros((h, w), bool)
        t = tm.copy()
        for fid in range(16):
            t["frameCountMod16"] = fid
            r.set_time(t)
            r.dispatch_cloud()
            now = r.read_image(api.IMAGE_CLOUD_CUR)[..., 3] != -7.0
            assert np.array_equal(now & ~seen, scene.cloud_pixels_written(w, h, fid)), f"pixel selection of id {fid}"
            seen = now
        # ... and their union is the full-quality frame (the jitter index depends on the id, not on the dispatch)
        check_frame(r.read_image(api.IMAGE_CLOUD_CUR), r.read_image(api.IMAGE_GODRAY_MASK), ref, "1080p union of 16 ids")
        # one id against the oracle's own 1-of-16 dispatch, untouched pixels included
        t["frameCountMod16"] = 11
        one = oracle_mod.cloud(cam, t, tun, noise, w, h, full=False, hdr=sentinel.copy(), mask=sentinel.copy())
        r.set_time(t)
        r.write_image(api.IMAGE_CLOUD_CUR, sentinel)
        r.write_image(api.IMAGE_GODRAY_MASK, sentinel)
        r.dispatch_cloud()
        check_frame(r.read_image(api.IMAGE_CLOUD_CUR), r.read_image(api.IMAGE_GODRAY_MASK), one, "1080p id 11")


def test_config2_sixteen_frame_pan_1080p_through_mtFrameEx(api, oracle_mod, noise):
    from meteoros_b200 import scene

    w, h = 1920, 1080
    cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
    tun = scene.default_tuning()
    img = [np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)]
    mask = np.zeros((h, w, 4), np.float32)
    cur, cam_old, worst, worst_ldr = 0, cam.ubo(), 0.0, 0
    with make_renderer(api, noise, w, h) as r:
        r.set_sun_and_sky(sky.ubo()); r.set_tuning(tun)
        for frame in range(16):
            cam.rotate_about_up(0.25)
            sc.update_time(1 / 60)
            c, t = cam.ubo(), sc.ubo()
            img[cur], taps = oracle_mod.reproject(c, cam_old, t, img[cur ^ 1], taps=True)
            oracle_mod.cloud(c, t, tun, noise, w, h, full=False, hdr=img[cur], mask=mask)
            img[cur] = oracle_mod.godrays(c, sky.ubo(), mask, img[cur])
            ldr_ref = oracle_mod.tonemap(t, img[cur])
            r.set_camera(c); r.set_camera_old(cam_old); r.set_time(t)
            if frame in (3, 12):  # reprojection tap indices at full size: bit-exact
                assert np.array_equal(r.dispatch_reprojection_debug(), taps), f"frame {frame}: reprojection taps"
            r.frame(with_godrays=True)
            got = r.read_image(api.IMAGE_CLOUD_PREV)  # roles swapped at the end of the frame
            e = rel_err(got[..., :3], img[cur][..., :3])
            assert e.max() <= HDR_MAX_REL, f"frame {frame}: HDR max rel err {e.max():.3e}"
            assert psnr(got[..., :3], img[cur][..., :3]) >= HDR_MIN_PSNR
            worst = max(worst, float(e.max()))
            assert np.array_equal(r.read_image(api.IMAGE_GODRAY_MASK), mask), f"frame {frame}: mask"
            d = np.abs(r.read_image(api.IMAGE_LDR_PREV).astype(np.int32) - ldr_ref.astype(np.int32))
            assert d.max() <= 1, f"frame {frame}: LDR differs by {d.max()}"
            worst_ldr = max(worst_ldr, int(d.max()))
            cur ^= 1
            cam_old = c
    print(f"1080p pan: worst HDR rel err {worst:.2e}, worst LDR difference {worst_ldr} LSB")


def test_config3_whole_4k_frame(api, oracle_mod, noise):
    w, h = 3840, 2160
    cam, tm, _, tun = default_scene(w, h)
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True, counters=True)
    with make_renderer(api, noise, w, h, flags=api.FLAG_COUNTERS) as r:
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.dispatch_cloud_full()
        hdr = r.read_image(api.IMAGE_CLOUD_CUR)
        mask = r.read_image(api.IMAGE_GODRAY_MASK)
        cnt = r.counters()
    assert cnt == ref["counters"]  # rays, marched rays, steps, in-cloud steps, cone hits, early exits: the same work
    e = check_frame(hdr, mask, ref, "4K full dispatch")
    print(f"4K frame: HDR max rel err {e:.2e}, PSNR {psnr(hdr[..., :3], ref['hdr'][..., :3]):.1f} dB, counters equal")
    with make_renderer(api, noise, w, h) as r:  # the production (non-counting) kernel writes the same image
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.dispatch_cloud_full()
        assert np.array_equal(r.read_image(api.IMAGE_CLOUD_CUR), hdr)
        assert np.array_equal(r.read_image(api.IMAGE_GODRAY_MASK), mask)


def test_config4_8k_frame_as_eight_row_tile_shards(api, oracle_mod, noise):
    from meteoros_b200 import sharding

    w, h, world, tile_rows = 7680, 4320, 8, 8
    cam, tm, _, tun = default_scene(w, h)
    n = sharding.num_tiles(h, tile_rows)
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        for rank in range(world):  # the launches the eight ranks of bench.py --workload frame8k issue
            tiles = sharding.tiles_of_rank(h, tile_rows, world, rank)
            r.dispatch_cloud_tiles(tile_rows, tiles.start, n, world)
        hdr = r.read_image(api.IMAGE_CLOUD_CUR)
        mask = r.read_image(api.IMAGE_GODRAY_MASK)
        r.clear_images()
        r.dispatch_cloud_full()
        assert np.array_equal(r.read_image(api.IMAGE_CLOUD_CUR), hdr)  # shards == one launch, whole frame
        assert np.array_equal(r.read_image(api.IMAGE_GODRAY_MASK), mask)
    for r0 in (4, 700, 1420, 1900, 2030, 3000):  # zenith .. just above the horizon band .. sky band .. ocean; bands straddle tiles of different ranks
        r1 = r0 + 20
        ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True, rows=(r0, r1))
        sub = {"hdr": ref["hdr"][r0:r1], "mask": ref["mask"][r0:r1]}
        check_frame(hdr[r0:r1], mask[r0:r1], sub, f"8K rows {r0}..{r1}")


@pytest.mark.parametrize("view", [0, 15, 96, 127, 128, 143, 240, 255])
def test_config5_sweep_views_1080p(api, oracle_mod, noise, view):
    """views 0/15: coverage 0.3 at sun elevation 5 / 85 deg; 240/255: coverage 0.9; the others in between."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import bench

    w, h = 1920, 1080
    cam, tm, _, tun = bench.scene_for_view(view, w, h, sweep=True)
    ref = oracle_mod.cloud(cam, tm, tun, noise, w, h, full=True)
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.dispatch_cloud_full()
        e = check_frame(r.read_image(api.IMAGE_CLOUD_CUR), r.read_image(api.IMAGE_GODRAY_MASK), ref, f"view {view}")
    print(f"view {view}: coverage {float(tun['coverage']):.2f}, HDR max rel err {e:.2e}")


@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 2160)])
def test_reference_shader_digests_at_full_size(api, noise, w, h):
    """The CUDA frame against the REFERENCE'S OWN Cloud shader at BASELINE's sizes, without the oracle in between:
    tests/golden/cloud_fullsize_digests.npz was written by cloudRayMarch.comp compiled from its text (make_goldens.py; the
    full frames are 66 / 265 MB, so what is committed is a CRC-32 of every pixel row of the god-ray mask and of alpha -- the
    bit-exact quantities -- every 8th / 16th pixel of the colour, and float64 row sums of the colour)."""
    import pathlib
    import zlib

    g = np.load(pathlib.Path(__file__).parent / "golden" / "cloud_fullsize_digests.npz")
    sub = int(g[f"sub_{w}x{h}"])
    cam, tm, _, tun = default_scene(w, h)
    with make_renderer(api, noise, w, h) as r:
        r.set_camera(cam); r.set_time(tm); r.set_tuning(tun)
        r.dispatch_cloud_full()
        hdr, mask = r.read_image(api.IMAGE_CLOUD_CUR), r.read_image(api.IMAGE_GODRAY_MASK)
    mask_crc = np.array([zlib.crc32(mask[y].tobytes()) for y in range(h)], np.uint32)
    alpha_crc = np.array([zlib.crc32(np.ascontiguousarray(hdr[y, :, 3]).tobytes()) for y in range(h)], np.uint32)
    assert np.array_equal(mask_crc, g[f"mask_crc_{w}x{h}"]), "god-ray mask: some pixel row differs from the reference shader's"
    assert np.array_equal(alpha_crc, g[f"alpha_crc_{w}x{h}"])
    want = g[f"hdr_sub_{w}x{h}"]
    e = rel_err(hdr[::sub, ::sub, :3], want)
    assert e.max() <= HDR_MAX_REL and psnr(hdr[::sub, ::sub, :3], want) >= HDR_MIN_PSNR
    rows = hdr[..., :3].astype(np.float64).sum(axis=1)
    assert np.allclose(rows, g[f"hdr_rowsum_{w}x{h}"], rtol=1e-5, atol=1e-7)   # every pixel enters a row sum
    print(f"{w}x{h}: mask + alpha rows CRC-equal to the reference shader, sampled HDR max rel err {e.max():.2e}")
