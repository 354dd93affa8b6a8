#!/usr/bin/env python
"""Small frame through every pass, for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402

from meteoros_b200 import api, scene, textures  # noqa: E402

w, h = 130, 70  # odd size: over-provisioned grid, partial tiles, dropped stores
cam, sc, sky = scene.Camera(w, h), scene.Scene(), scene.Sky()
with api.CloudRenderer(w, h, flags=api.FLAG_COUNTERS) as r:
    r.upload_noise(textures.load_noise())
    r.set_sun_and_sky(sky.ubo())
    old = cam.ubo()
    for f in range(3):
        cam.rotate_about_up(0.25)
        sc.update_time(1 / 60)
        r.set_camera(cam.ubo()); r.set_camera_old(old); r.set_time(sc.ubo())
        r.frame(True, True)
        old = cam.ubo()
    r.dispatch_cloud_full()
    r.dispatch_cloud_tiles(8, 1, 9, 2)
    r.dispatch_cloud_debug(False)
    r.dispatch_reprojection_debug()
    print("counters", r.counters())
with api.CloudRenderer(w, h) as r:  # step-parallel 1/16 path
    r.upload_noise(textures.load_noise())
    r.set_camera(cam.ubo()); r.set_camera_old(old); r.set_time(sc.ubo())
    r.dispatch_cloud()
    print("mean", float(np.nanmean(r.read_image(api.IMAGE_CLOUD_CUR))))
# god rays walk from every pixel to the clamped sun position without clamping the taps (post_core.cuh::mask_decode):
# sweep the sun over the frame's interior, edges and corners, and past them
for wh in ((130, 70), (33, 17), (64, 36)):
    with api.CloudRenderer(*wh) as r:
        r.upload_noise(textures.load_noise())
        r.set_sun_and_sky(sky.ubo())
        for yaw in (0.0, 35.0, -35.0, 90.0, 179.0):
            for pitch in (0.0, 30.0, 60.0, 89.0, -30.0):
                c = scene.Camera(*wh)
                c.rotate_about_up(yaw)
                if pitch:
                    c.rotate_about_right(pitch)
                r.set_camera(c.ubo()); r.set_camera_old(c.ubo()); r.set_time(sc.ubo())
                r.frame(True, True)
        r.synchronize()
tun = scene.default_tuning()
tun["use_weather"], tun["weather_scale"] = 1, 1.0e-4
with api.CloudRenderer(w, h) as r:  # weather variants of the march kernels
    r.upload_noise(textures.load_noise())
    r.set_camera(cam.ubo()); r.set_camera_old(old); r.set_time(sc.ubo()); r.set_tuning(tun)
    r.dispatch_cloud()
    r.dispatch_cloud_full()
    print("weather mean", float(np.nanmean(r.read_image(api.IMAGE_CLOUD_CUR))))
# round 2: RGBA16F storage through every pass, the bulk-store epilogue, the forwarder, the one-float god-ray read-back, the
# three-kernel 1-of-16 form (MT_FLAG_SPLIT_MARCH) and the canonical cone filter (MT_FLAG_NO_CONE_RF)
for storage in (api.STORAGE_F16, api.STORAGE_F16_EMULATE):
    with api.CloudRenderer(w, h, storage=storage) as r, api.CloudRenderer(w, h, storage=storage) as peer:
        r.upload_noise(textures.load_noise())
        r.set_sun_and_sky(sky.ubo())
        for f in range(2):
            cam.rotate_about_up(0.25)
            sc.update_time(1 / 60)
            r.set_camera(cam.ubo()); r.set_camera_old(old); r.set_time(sc.ubo())
            r.frame(True, True)
            old = cam.ubo()
        r.set_cloud_store_mode(api.STORE_BULK)
        r.dispatch_cloud_full()
        r.set_cloud_output(peer.image_device_ptr(api.IMAGE_CLOUD_CUR), None)
        r.dispatch_cloud_tiles(8, 0, 9, 1)
        r.set_cloud_output(None, None)
        r.set_cloud_store_mode(api.STORE_DIRECT)
        if w % 2 == 0 or storage != api.STORAGE_F16:
            r.set_cloud_forward(peer.image_device_ptr(api.IMAGE_CLOUD_PREV))
            r.dispatch_cloud_tiles(8, 0, 9, 1)
            r.join_copies(); r.synchronize()
            r.set_cloud_forward(None)
        print("grey mean", float(r.read_godray_grey().mean()), "storage", storage)
for flags in (api.FLAG_SPLIT_MARCH, api.FLAG_NO_CONE_RF, api.FLAG_TOP_DOWN | api.FLAG_NO_FUSED_TONEMAP, api.FLAG_HW_CONE_FILTER):
    with api.CloudRenderer(w, h, flags=flags) as r:
        r.upload_noise(textures.load_noise())
        r.set_sun_and_sky(sky.ubo())
        r.set_camera(cam.ubo()); r.set_camera_old(old); r.set_time(sc.ubo())
        r.frame(True, True)
        r.dispatch_cloud_full()
        print("flags", flags, "mean", float(np.nanmean(r.read_image(api.IMAGE_CLOUD_CUR))))
# round 2, second session: the two-stream frame with the Cloud kernel keeping the decoded mask current (no timing / counting flag),
# a host write into the mask in between; the generic march kernels (texture coordinates beyond the magic floor's range); the IEEE
# reprojection / TXAA kernels (ray origin above the inner shell) and taps far outside the image (a quarter turn between frames)
for size in ((130, 70), (33, 17), (64, 36)):
    ww, hh = size
    cam2, sc2 = scene.Camera(ww, hh), scene.Scene()
    with api.CloudRenderer(ww, hh) as r:
        r.upload_noise(textures.load_noise())
        r.set_sun_and_sky(sky.ubo())
        old2 = cam2.ubo()
        for f in range(5):
            cam2.rotate_about_up(0.25 if f != 3 else 88.0)
            sc2.update_time(1 / 60)
            r.set_camera(cam2.ubo()); r.set_camera_old(old2); r.set_time(sc2.ubo())
            if f == 2:
                r.write_image(api.IMAGE_GODRAY_MASK, r.read_image(api.IMAGE_GODRAY_MASK) * 0.5)
            r.frame(True, True)
            old2 = cam2.ubo()
        tun = scene.default_tuning()
        tun["cloud_speed"] = 200.0
        sc2.time["time"] = (0.016, 100.0)
        r.set_time(sc2.ubo()); r.set_tuning(tun)
        r.frame(True, True)
        r.dispatch_cloud_full()
        print("two-stream frames", size, "mean", float(np.nanmean(r.read_image(api.IMAGE_CLOUD_CUR))))
    high = scene.Camera(ww, hh, eye=(0.0, -9000.0, 2.0), ref=(0.0, -9000.0, 1.0))
    with api.CloudRenderer(ww, hh) as r:
        r.upload_noise(textures.load_noise())
        r.set_sun_and_sky(sky.ubo())
        o = high.ubo()
        high.rotate_about_up(0.25)
        r.set_camera(high.ubo()); r.set_camera_old(o); r.set_time(sc2.ubo())
        r.frame(True, True)
        print("eye above the inner shell", size, "mean", float(np.nanmean(r.read_image(api.IMAGE_CLOUD_PREV))))
