#!/usr/bin/env python
"""Decode the reference's shipped noise inputs into one compressed fixture.

Runs only in the build container (needs /root/reference).  The GPU box has no
/root/reference, so the decoded RGBA8 volumes travel as tests/golden/noise_volumes.npz.

Inputs ("Input fixtures that DO exist"):
  textures/CloudTextures/LowFrequency/LowFrequency(1..128).tga   128x128 RGBA  -> low  [z][y][x][c]
  textures/CloudTextures/HighFrequency/HighFrequency(1..32).tga  32x32 RGBA    -> high [z][y][x][c]
  textures/CloudTextures/curlNoise.png                           128x128 RGBA  -> curl [y][x][c]
  textures/CloudTextures/weatherMap.png                          512x512 16-bit RGBA, 8-bit = value >> 8
Layout follows the reference loader (ImageLoadingUtility.cpp:75-139: slice z = file index z+1,
rows top-down as stb_image decodes them); sampler state is LINEAR/REPEAT (Texture3D.cpp:92-134).
The textures are (c) 2017 Aman Sachan, MIT licence (see /root/reference/LICENSE).
"""
import hashlib
import sys
from pathlib import Path

import numpy as np
from PIL import Image

REF = Path("/root/reference/src/CloudScapes/textures/CloudTextures")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "noise_volumes.npz"

EXPECT = {
    "low": "44448f940ff2f3ba3698ce7ab31e4a91de41915cf2a574bf3f67a1f902530867",
    "high": "bd87fefa78192ef26d1914b96bb44a8dea2cd7247dbe588d5cb8d7976ef8fb5c",
    "curl": "21cc9bcdbe4c90f018aa3687741a8fc3d162ea1d49a2923ff719cae3b6012d9a",
    "weather": "a425eef74edfe98bdb435e0cd0714c9a8f0b7dafa278dfdf8b5811baa4ae62e2",
}


def slices(folder, base, n):
    vol = []
    for z in range(n):
        im = Image.open(REF / folder / f"{base}({z + 1}).tga").convert("RGBA")
        vol.append(np.asarray(im, dtype=np.uint8))
    return np.ascontiguousarray(np.stack(vol, axis=0))


def weather():
    im = Image.open(REF / "weatherMap.png")
    a = np.asarray(im)
    if a.dtype == np.uint16:  # stb_image 16 -> 8 bit conversion is a plain >> 8
        a = (a >> 8).astype(np.uint8)
    if a.ndim == 2:
        a = np.stack([a] * 3 + [np.full_like(a, 255)], -1)
    if a.shape[-1] == 3:
        a = np.concatenate([a, np.full(a.shape[:2] + (1,), 255, np.uint8)], -1)
    return np.ascontiguousarray(a.astype(np.uint8))


def main():
    data = {
        "low": slices("LowFrequency", "LowFrequency", 128),
        "high": slices("HighFrequency", "HighFrequency", 32),
        "curl": np.ascontiguousarray(np.asarray(Image.open(REF / "curlNoise.png").convert("RGBA"), dtype=np.uint8)),
        "weather": weather(),
    }
    ok = True
    for k, v in data.items():
        h = hashlib.sha256(v.tobytes()).hexdigest()
        flag = "ok" if h == EXPECT[k] else "MISMATCH"
        ok &= h == EXPECT[k]
        print(f"{k:8s} {v.shape!s:22s} sha256={h} {flag}")
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, OUT.stat().st_size, "bytes")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
