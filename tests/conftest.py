import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def noise():
    from meteoros_b200 import textures

    return textures.load_noise()


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle

    oracle.lib()
    return oracle


def default_scene(width, height, frame_id=1, total_time=0.016, yaw=0.0, pitch=0.0):
    """Default cloudscape of the reference (main.cpp:157-158) with a chosen frame id / time / camera rotation."""
    from meteoros_b200 import scene

    cam = scene.Camera(width, height)
    if yaw:
        cam.rotate_about_up(yaw)
    if pitch:
        cam.rotate_about_right(pitch)
    sc = scene.Scene()
    sc.time["time"] = (0.016, total_time)
    sc.time["frameCountMod16"] = frame_id
    return cam.ubo(), sc.ubo(), scene.Sky().ubo(), scene.default_tuning()


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def psnr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    mse = np.mean((a - b) ** 2)
    peak = max(float(np.max(np.abs(b))), 1e-12)
    return float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)
