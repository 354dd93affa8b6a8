"""Uniform producers against the REFERENCE'S OWN code.

tests/golden/ref_uniforms.npz was minted by tests/golden/make_ref_uniforms.py from camera.cpp / Scene.cpp / Sky.cpp
compiled unmodified (vendored glm, a stand-in Vulkan header; oracle/Makefile target `ref`).  This is the one part of the
hot path's inputs where the reference itself runs here, so the bar is byte equality for the C++ producers the library
ships (mt_scene.cpp, mtx*), and a float-rounding bound for the Python mirror (double sin/cos rounded once vs sinf/cosf)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

from meteoros_b200 import _lib, scene

GOLDEN = Path(__file__).parent / "golden"
sys.path.insert(0, str(GOLDEN))
import make_ref_uniforms as ref  # noqa: E402

FIX = np.load(GOLDEN / "ref_uniforms.npz")
MTX = {"rotate_up": "mtxCameraRotateAboutUp", "rotate_right": "mtxCameraRotateAboutRight", "along_look": "mtxCameraTranslateAlongLook",
       "along_right": "mtxCameraTranslateAlongRight", "along_up": "mtxCameraTranslateAlongUp"}
PY = {"rotate_up": "rotate_about_up", "rotate_right": "rotate_about_right", "along_look": "translate_along_look",
      "along_right": "translate_along_right", "along_up": "translate_along_up"}


@pytest.mark.parametrize("case", ref.CASES, ids=[c[0] for c in ref.CASES])
def test_cxx_camera_is_byte_identical_to_reference_camera(case):
    name, w, h, eye, center, fovy, near, far, ops = case
    want = FIX["camera_" + name]
    lib = _lib.load()
    cam = _lib.MtxCamera()
    lib.mtxCameraInit(C.byref(cam), w, h, (C.c_float * 3)(*eye), (C.c_float * 3)(*center), C.c_float(fovy), C.c_float(near), C.c_float(far))
    for k in range(len(ops) + 1):
        if k:
            getattr(lib, MTX[ops[k - 1][0]])(C.byref(cam), C.c_float(ops[k - 1][1]))
        u = np.zeros((), scene.CAMERA_DTYPE)
        lib.mtxCameraUBO(C.byref(cam), u.ctypes.data)
        assert u.tobytes() == want[k].tobytes(), f"{name}: state {k}"


@pytest.mark.parametrize("case", ref.CASES, ids=[c[0] for c in ref.CASES])
def test_python_camera_mirror_tracks_reference_camera(case):
    name, w, h, eye, center, fovy, near, far, ops = case
    want = FIX["camera_" + name].view(np.float32)
    cam = scene.Camera(w, h, eye=eye, ref=center, fovy=fovy, near=near, far=far)
    if fovy == 45.0:
        assert cam.ubo().tobytes() == want[0].tobytes()        # the reference's own camera, before any sin/cos: identical
    for k, (op, arg) in enumerate([(None, 0.0)] + list(ops)):
        if op:
            getattr(cam, PY[op])(arg)
        got = np.frombuffer(cam.ubo().tobytes(), np.float32)
        scale = np.maximum(1.0, np.abs(want[k]))
        assert np.all(np.abs(got - want[k]) <= 4e-6 * scale), f"{name}: state {k}"


def test_time_sky_and_halton_are_byte_identical_to_reference():
    lib = _lib.load()
    t = np.zeros((), scene.TIME_DTYPE)
    py = scene.Scene()
    lib.mtxTimeInit(t.ctypes.data)
    for k in range(21):
        if k:
            lib.mtxTimeUpdate(t.ctypes.data, C.c_float(1.0 / 60.0))
            py.update_time(1.0 / 60.0)
        for blk in (t, py.ubo()):
            b = np.frombuffer(blk.tobytes(), np.uint8).copy()
            b[64:72] = 0                                       # wall-clock fields are not in the fixture
            assert b.tobytes() == FIX["time"][k].tobytes(), k
    s = np.zeros((), scene.SUNSKY_DTYPE)
    lib.mtxSunAndSky(s.ctypes.data)
    assert s.tobytes() == FIX["sun_and_sky"].tobytes() == scene.Sky().ubo().tobytes()
    for j, base in enumerate((2, 3, 5)):
        for i in range(65):
            assert scene.halton_sequence_at(i, base) == FIX["halton"][j][i]


@pytest.mark.skipif(not Path("/root/reference/src/CloudScapes/camera.cpp").exists(), reason="reference tree not present (GPU box)")
def test_fixture_regenerates_from_the_reference_sources():
    live = ref.collect(ref.build())
    assert set(live) == set(FIX.files)
    for k in live:
        assert np.array_equal(live[k], FIX[k]), k


@pytest.mark.skipif(not Path("/root/reference/src/CloudScapes/ImageLoadingUtility.cpp").exists(), reason="reference tree not present (GPU box)")
def test_reference_texture_loader_runs_here_and_agrees_with_ours():
    """ImageLoadingUtility.cpp + the vendored stb_image.h, compiled unmodified (oracle/_ref), load the four cloud
    textures exactly as Sky::CreateCloudResources does; the bytes it would upload equal (a) the committed fixture the
    GPU tests run on, (b) the library's own decoders (mt_assets.cpp), (c) the SHA-256 values of SURVEY.md appendix A."""
    import hashlib

    from meteoros_b200 import textures

    lib_ref = ref.build()
    ours = _lib.load()
    tex = Path("/root/reference/src/CloudScapes/textures/CloudTextures")
    fixture = textures.load_noise()
    for key, base, n in (("low", "LowFrequency", 128), ("high", "HighFrequency", 32)):
        folder = (str(tex / base) + "/").encode()
        got = np.zeros((n, n, n, 4), np.uint8)
        assert lib_ref.mtref_load_volume(folder, base.encode(), b".tga", n, n, n, got.ctypes.data_as(C.c_void_p)) == 0
        mine = np.zeros_like(got)
        assert ours.mtxLoadVolumeFromSlices(folder, base.encode(), b".tga", n, n, n, mine.ctypes.data, mine.nbytes) == 0
        assert np.array_equal(got, fixture[key]) and np.array_equal(got, mine)
        assert hashlib.sha256(got.tobytes()).hexdigest() == textures.SHA256[key]
    for key, name, n in (("curl", "curlNoise.png", 128), ("weather", "weatherMap.png", 512)):
        path = str(tex / name).encode()
        got = np.zeros((n, n, 4), np.uint8)
        assert lib_ref.mtref_load_image(path, n, n, got.ctypes.data_as(C.c_void_p)) == 0
        mine = np.zeros_like(got)
        w, h = C.c_uint32(), C.c_uint32()
        assert ours.mtxLoadImageFile(path, mine.ctypes.data, mine.nbytes, C.byref(w), C.byref(h)) == 0
        assert np.array_equal(got, fixture[key]) and np.array_equal(got, mine)
        assert hashlib.sha256(got.tobytes()).hexdigest() == textures.SHA256[key]
    assert lib_ref.mtref_load_image(b"/nonexistent.png", 4, 4, np.zeros(64, np.uint8).ctypes.data_as(C.c_void_p)) == -1
