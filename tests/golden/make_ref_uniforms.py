#!/usr/bin/env python
"""Mint tests/golden/ref_uniforms.npz from the REFERENCE'S OWN uniform producers.

Runs only where /root/reference exists: `make -C oracle ref` compiles camera.cpp / Scene.cpp / Sky.cpp unmodified (with
the reference's vendored glm) into oracle/_ref/libmeteoros_ref.so; this script drives them through the control calls
of main.cpp:60-110 and the frame loop main.cpp:172-194 and records the uniform blocks they memcpy into their mapped
buffers.  The fixture travels; the reference does not.  tests/test_reference_inputs.py checks mt_scene.cpp (mtx*) and
meteoros_b200/scene.py against it, and re-derives it live when the reference is present."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
OPS = {"rotate_up": 0, "rotate_right": 1, "along_look": 2, "along_right": 3, "along_up": 4}

# (name, width, height, eye, ref, fovy, near, far, [(op, arg), ...])
CASES = [
    ("pan_1080p", 1920, 1080, (0, 0, 2), (0, 0, 1), 45.0, 0.1, 1000.0, [("rotate_up", 0.25)] * 16),     # BASELINE config 2
    ("default_720p", 1284, 720, (0, 0, 2), (0, 0, 1), 45.0, 0.1, 1000.0, []),                         # main.cpp:22-23, 157-158
    ("default_4k", 3840, 2160, (0, 0, 2), (0, 0, 1), 45.0, 0.1, 1000.0, []),
    ("default_8k", 7680, 4320, (0, 0, 2), (0, 0, 1), 45.0, 0.1, 1000.0, []),
    ("odd_130x70", 130, 70, (0, 0, 2), (0, 0, 1), 45.0, 0.1, 1000.0, [("rotate_right", -3.0), ("rotate_up", 12.0)]),
    ("keys_mixed", 1920, 1080, (0, 0, 2), (0, 0, 1), 45.0, 0.1, 1000.0,
     [("rotate_up", 0.25), ("rotate_up", 0.25), ("rotate_right", -0.25), ("along_look", 0.5), ("along_right", -0.5),
      ("rotate_up", -0.25), ("along_up", 0.5), ("rotate_right", 0.25), ("along_look", -0.5), ("rotate_up", 0.25)] * 3),
    ("off_axis", 1000, 700, (3.5, -2.0, 10.0), (1.0, 4.0, -6.0), 60.0, 0.5, 5000.0,
     [("rotate_up", 7.5), ("rotate_right", 2.5), ("along_look", 25.0), ("rotate_up", -15.0)]),
]


def build() -> C.CDLL:
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "ref"], check=True)
    lib = C.CDLL(str(ROOT / "oracle" / "_ref" / "libmeteoros_ref.so"))
    lib.mtref_halton.restype = C.c_float
    return lib


def camera_case(lib, case) -> np.ndarray:
    _, w, h, eye, ref, fovy, near, far, ops = case
    out = np.zeros((len(ops) + 1, 152), np.uint8)
    op = (C.c_int * max(1, len(ops)))(*[OPS[o] for o, _ in ops])
    arg = (C.c_float * max(1, len(ops)))(*[a for _, a in ops])
    rc = lib.mtref_camera((C.c_float * 3)(*eye), (C.c_float * 3)(*ref), w, h, C.c_float(fovy), C.c_float(near), C.c_float(far),
                          len(ops), op, arg, out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def collect(lib) -> dict:
    data = {}
    for case in CASES:
        data["camera_" + case[0]] = camera_case(lib, case)
    t = np.zeros((21, 76), np.uint8)
    assert lib.mtref_time(20, t.ctypes.data_as(C.c_void_p)) == 0
    t[:, 64:72] = 0                      # delta / total time come from the wall clock: not part of the fixture
    data["time"] = t
    s = np.zeros(52, np.uint8)
    assert lib.mtref_sun_and_sky(s.ctypes.data_as(C.c_void_p)) == 0
    data["sun_and_sky"] = s
    data["halton"] = np.array([[lib.mtref_halton(i, b) for i in range(0, 65)] for b in (2, 3, 5)], np.float32)
    return data


if __name__ == "__main__":
    d = collect(build())
    np.savez_compressed(ROOT / "tests" / "golden" / "ref_uniforms.npz", **d)
    print({k: v.shape for k, v in d.items()})
