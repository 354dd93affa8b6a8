"""The C++ uniform producers exported by the library (mtx*) against the independent Python mirror of
Camera / Scene / Sky in meteoros_b200/scene.py -- two restatements of glm's arithmetic that must agree."""
import ctypes as C

import numpy as np

from meteoros_b200 import _lib, scene


def _ubo(lib, cam):
    u = np.zeros((), scene.CAMERA_DTYPE)
    lib.mtxCameraUBO(C.byref(cam), u.ctypes.data)
    return u


def _close(a, b, ulps=2):
    a = np.asarray(a, np.float32).ravel()
    b = np.asarray(b, np.float32).ravel()
    return bool(np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)) + 1e-45))


def test_default_camera_and_pan_match_python_mirror():
    lib = _lib.load()
    for w, h in ((1920, 1080), (1284, 720), (3840, 2160)):
        cam = _lib.MtxCamera()
        lib.mtxCameraInit(C.byref(cam), w, h, None, None, 45.0, 0.1, 1000.0)
        py = scene.Camera(w, h)
        a, b = _ubo(lib, cam), py.ubo()
        assert a["view"].tobytes() == b["view"].tobytes() and a["eye"].tobytes() == b["eye"].tobytes()
        assert a["tanFovBy2"].tobytes() == b["tanFovBy2"].tobytes()  # double tan with PI = 3.14159, camera.cpp:40
        assert _close(a["proj"], b["proj"])                           # tanf vs tan(double) rounded: <= 1-2 ulp
        for k in range(16):  # main.cpp:76-77: 0.25 degrees per frame; sinf/cosf vs double sin/cos rounded
            lib.mtxCameraRotateAboutUp(C.byref(cam), 0.25)
            py.rotate_about_up(0.25)
            if k % 5 == 0:
                lib.mtxCameraRotateAboutRight(C.byref(cam), -0.25)
                py.rotate_about_right(-0.25)
        a, b = _ubo(lib, cam), py.ubo()
        assert np.allclose(a["view"], b["view"], rtol=0, atol=5e-6) and np.allclose(a["proj"], b["proj"], rtol=1e-6)
        assert np.allclose(np.array(cam.forward[:]), py.forward, atol=2e-6)
        lib.mtxCameraTranslateAlongLook(C.byref(cam), 3.0)
        assert abs(np.linalg.norm(np.array(cam.eye[:]) - py.eye) - 3.0) < 1e-5


def test_time_and_sky_match_python_mirror():
    lib = _lib.load()
    t = np.zeros((), scene.TIME_DTYPE)
    lib.mtxTimeInit(t.ctypes.data)
    py = scene.Scene()
    assert t.tobytes() == py.ubo().tobytes()  # Halton base 3 table, frame 0
    for _ in range(20):
        lib.mtxTimeUpdate(t.ctypes.data, C.c_float(1.0 / 60.0))
        py.update_time(1.0 / 60.0)
        assert t.tobytes() == py.ubo().tobytes()
    assert int(t["frameCountMod16"]) == 4
    s = np.zeros((), scene.SUNSKY_DTYPE)
    lib.mtxSunAndSky(s.ctypes.data)
    assert s.tobytes() == scene.Sky().ubo().tobytes()


def test_run_frame_rejects_null_arguments():
    lib = _lib.load()
    cam = _lib.MtxCamera()
    assert lib.mtxRunFrame(None, C.byref(cam), None, None, C.c_float(0.016), 2) == 1
