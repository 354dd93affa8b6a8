"""Noise inputs of the cloud pass (Sky::CreateCloudResources, Sky.cpp:25-58).

`load_noise()` returns the four RGBA8 arrays the reference uploads, in the reference's memory order
(ImageLoadingUtility.cpp:87-98: volume[z][y][x][rgba], slice z = file "Name(z+1).tga", rows top-down):
  low     (128,128,128,4)  cloudBaseShapeSampler
  high    (32,32,32,4)     cloudDetailsHighFreqSampler (alpha == 0)
  curl    (128,128,4)      curlNoiseSampler
  weather (512,512,4)      weatherMapSampler (bound, never sampled)

They come from tests/golden/noise_volumes.npz, the decoded copy of the reference's texture files made by
tools/make_noise_fixture.py (the GPU box has no /root/reference); SHA-256 of each array is checked against
SURVEY.md appendix A.  `synthetic_noise()` makes deterministic random volumes of the same shapes for stress runs.
"""
from __future__ import annotations

import hashlib
from pathlib import Path

import numpy as np

FIXTURE = Path(__file__).resolve().parents[1] / "tests" / "golden" / "noise_volumes.npz"

SHA256 = {
    "low": "44448f940ff2f3ba3698ce7ab31e4a91de41915cf2a574bf3f67a1f902530867",
    "high": "bd87fefa78192ef26d1914b96bb44a8dea2cd7247dbe588d5cb8d7976ef8fb5c",
    "curl": "21cc9bcdbe4c90f018aa3687741a8fc3d162ea1d49a2923ff719cae3b6012d9a",
    "weather": "a425eef74edfe98bdb435e0cd0714c9a8f0b7dafa278dfdf8b5811baa4ae62e2",
}

_cache: dict[str, np.ndarray] | None = None


def load_noise(verify: bool = True) -> dict[str, np.ndarray]:
    global _cache
    if _cache is None:
        if not FIXTURE.exists():
            raise FileNotFoundError(f"{FIXTURE} missing; run tools/make_noise_fixture.py where /root/reference exists")
        with np.load(FIXTURE) as z:
            data = {k: np.ascontiguousarray(z[k]) for k in ("low", "high", "curl", "weather")}
        if verify:
            for k, v in data.items():
                h = hashlib.sha256(v.tobytes()).hexdigest()
                if h != SHA256[k]:
                    raise ValueError(f"noise fixture '{k}' has sha256 {h}, expected {SHA256[k]}")
        _cache = data
    return _cache


def synthetic_noise(seed: int = 0) -> dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    return {
        "low": rng.integers(0, 256, (128, 128, 128, 4), dtype=np.uint8),
        "high": rng.integers(0, 256, (32, 32, 32, 4), dtype=np.uint8),
        "curl": rng.integers(0, 256, (128, 128, 4), dtype=np.uint8),
        "weather": rng.integers(0, 256, (512, 512, 4), dtype=np.uint8),
    }


def load_noise_from_reference_tree(textures_dir) -> dict[str, np.ndarray]:
    """Decode the reference's own texture files with the library's C++ decoders (mtx*, no PIL): what a C++ host does in
    place of Sky::CreateCloudResources.  `textures_dir` = .../src/CloudScapes/textures/CloudTextures."""
    import ctypes as C

    from . import _lib

    lib = _lib.load()
    d = str(Path(textures_dir)) + "/"
    low = np.zeros((128, 128, 128, 4), np.uint8)
    high = np.zeros((32, 32, 32, 4), np.uint8)
    for arr, sub, base, n in ((low, "LowFrequency/", "LowFrequency", 128), (high, "HighFrequency/", "HighFrequency", 32)):
        st = lib.mtxLoadVolumeFromSlices((d + sub).encode(), base.encode(), b".tga", n, n, n, arr.ctypes.data, arr.nbytes)
        if st != 0:
            raise OSError(f"mtxLoadVolumeFromSlices failed for {d + sub}")
    out = {"low": low, "high": high}
    for key, name, n in (("curl", "curlNoise.png", 128), ("weather", "weatherMap.png", 512)):
        arr = np.zeros((n, n, 4), np.uint8)
        w, h = C.c_uint32(), C.c_uint32()
        st = lib.mtxLoadImageFile((d + name).encode(), arr.ctypes.data, arr.nbytes, C.byref(w), C.byref(h))
        if st != 0 or (w.value, h.value) != (n, n):
            raise OSError(f"mtxLoadImageFile failed for {d + name}")
        out[key] = arr
    return out
