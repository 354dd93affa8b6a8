"""Asset pipeline: the library's own TGA / PNG decoders (mt_assets.cpp, no stb, no PIL)."""
import ctypes as C
import hashlib
import io
from pathlib import Path

import numpy as np
import pytest

from meteoros_b200 import _lib, textures

REF_TEXTURES = Path("/root/reference/src/CloudScapes/textures/CloudTextures")


def _decode(lib, data: bytes, is_png: bool):
    w, h = C.c_uint32(), C.c_uint32()
    assert lib.mtxDecodeImage(data, len(data), int(is_png), None, 0, C.byref(w), C.byref(h)) == 0
    out = np.zeros((h.value, w.value, 4), np.uint8)
    assert lib.mtxDecodeImage(data, len(data), int(is_png), out.ctypes.data, out.nbytes, C.byref(w), C.byref(h)) == 0
    return out


@pytest.mark.skipif(not REF_TEXTURES.exists(), reason="reference tree not present on this machine")
def test_reference_textures_decode_to_the_survey_fingerprints():
    """The strongest external pin available for this path: decoding the reference's shipped files with our decoders gives
    the SHA-256 values SURVEY.md appendix A recorded from stb_image / PIL."""
    got = textures.load_noise_from_reference_tree(REF_TEXTURES)
    for k, v in got.items():
        assert hashlib.sha256(v.tobytes()).hexdigest() == textures.SHA256[k], k
    fixture = textures.load_noise()
    for k in got:
        assert np.array_equal(got[k], fixture[k])


def test_png_and_tga_decoders_against_pil(tmp_path):
    from PIL import Image

    lib = _lib.load()
    rng = np.random.default_rng(4)
    rgba = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    rgba[:10] = rgba[0, 0]  # long runs for the RLE / LZ77 paths
    for mode, arr in (("RGBA", rgba), ("RGB", rgba[..., :3]), ("L", rgba[..., 0]), ("LA", rgba[..., [0, 3]])):
        buf = io.BytesIO()
        Image.fromarray(arr.squeeze(), mode).save(buf, format="PNG", compress_level=6)
        want = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGBA"))
        assert np.array_equal(_decode(lib, buf.getvalue(), True), want), mode
    # 16-bit grey: the high byte survives (stb_image: v >> 8)
    g16 = rng.integers(0, 65536, (9, 11), dtype=np.uint16)
    buf = io.BytesIO()
    Image.fromarray(g16).save(buf, format="PNG")  # uint16 -> mode I;16
    got = _decode(lib, buf.getvalue(), True)
    assert np.array_equal(got[..., 0], (g16 >> 8).astype(np.uint8)) and (got[..., 3] == 255).all()
    # stored (uncompressed) deflate blocks
    buf = io.BytesIO()
    Image.fromarray(rgba, "RGBA").save(buf, format="PNG", compress_level=0)
    assert np.array_equal(_decode(lib, buf.getvalue(), True), rgba)
    # TGA: PIL writes bottom-up or top-down, raw or RLE
    for kw in ({}, {"compression": "tga_rle"}, {"orientation": 1}, {"compression": "tga_rle", "orientation": 1}):
        for mode, arr in (("RGBA", rgba), ("RGB", rgba[..., :3])):
            buf = io.BytesIO()
            Image.fromarray(arr, mode).save(buf, format="TGA", **kw)
            want = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGBA"))
            assert np.array_equal(_decode(lib, buf.getvalue(), False), want), (kw, mode)
    # garbage is rejected, not crashed on
    w, h = C.c_uint32(), C.c_uint32()
    assert lib.mtxDecodeImage(b"not an image at all", 19, 1, None, 0, C.byref(w), C.byref(h)) == 1
    assert lib.mtxDecodeImage(buf.getvalue()[:40], 40, 0, None, 0, C.byref(w), C.byref(h)) == 1


def test_volume_cache_roundtrip(tmp_path):
    lib = _lib.load()
    vol = np.random.default_rng(1).integers(0, 256, (8, 16, 32, 4), dtype=np.uint8)
    path = str(tmp_path / "low.mtvol").encode()
    assert lib.mtxSaveVolume(path, 32, 16, 8, vol.ctypes.data) == 0
    w, h, d = C.c_uint32(), C.c_uint32(), C.c_uint32()
    assert lib.mtxLoadVolume(path, None, 0, C.byref(w), C.byref(h), C.byref(d)) == 0
    assert (w.value, h.value, d.value) == (32, 16, 8)
    back = np.zeros_like(vol)
    assert lib.mtxLoadVolume(path, back.ctypes.data, back.nbytes, C.byref(w), C.byref(h), C.byref(d)) == 0
    assert np.array_equal(back, vol)
    assert lib.mtxLoadVolume(str(tmp_path / "missing.mtvol").encode(), None, 0, C.byref(w), C.byref(h), C.byref(d)) == 1


def test_decoders_reject_malformed_files():
    """Corrupt, truncated and hostile files (size fields of 2^31, pixel data missing, garbage deflate streams) must come
    back as an error status -- never an exception across the C boundary, a crash or an unbounded allocation.
    tools/fuzz_assets.cpp is the long-running ASan / UBSan version of this loop."""
    from PIL import Image

    lib = _lib.load()
    rng = np.random.default_rng(1)

    def enc(arr, fmt, **kw):
        b = io.BytesIO()
        Image.fromarray(arr).save(b, fmt, **kw)
        return b.getvalue()

    seeds = [(enc(rng.integers(0, 256, (9, 17, 4), dtype=np.uint8), "PNG"), 1), (enc(rng.integers(0, 256, (8, 8, 3), dtype=np.uint8), "PNG"), 1),
             (enc(rng.integers(0, 256, (7, 5), dtype=np.uint8), "PNG"), 1),
             (enc(rng.integers(0, 256, (16, 16, 4), dtype=np.uint8), "TGA", compression="tga_rle"), 0),
             (enc(rng.integers(0, 256, (5, 9, 4), dtype=np.uint8), "TGA"), 0)]

    def decode(data, is_png):
        w, h = C.c_uint32(), C.c_uint32()
        st = lib.mtxDecodeImage(data, len(data), is_png, None, 0, C.byref(w), C.byref(h))
        if st != 0:
            return st
        assert w.value <= 16384 and h.value <= 16384
        out = np.zeros((h.value, w.value, 4), np.uint8)
        return lib.mtxDecodeImage(data, len(data), is_png, out.ctypes.data, out.nbytes, C.byref(w), C.byref(h))

    # the case that used to abort the process: a PNG / TGA header announcing a 2^31-pixel-wide image
    png = bytearray(seeds[0][0]); png[16:20] = (0x7F, 0xFF, 0xFF, 0xFF)
    tga = bytearray(seeds[4][0]); tga[12:16] = (0xFF, 0xFF, 0xFF, 0xFF)
    assert decode(bytes(png), 1) != 0 and decode(bytes(tga), 0) != 0
    ok = 0
    for it in range(4000):
        data, is_png = seeds[it % len(seeds)]
        b = bytearray(data)
        mode = it % 4
        if mode == 0:
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif mode == 1:
            b = b[:int(rng.integers(0, len(b)))]
        elif mode == 2:
            i = int(rng.integers(0, len(b)))
            b[i:i] = bytes(rng.integers(0, 256, int(rng.integers(1, 9)), dtype=np.uint8))
        else:
            i = int(rng.integers(0, max(1, len(b) - 4)))
            b[i:i + 4] = int(rng.integers(0, 2**32)).to_bytes(4, "big")
        ok += decode(bytes(b), is_png) == 0
    assert 0 < ok < 4000          # some mutations only touch pixel data; most must be rejected cleanly
