"""Host-side mirror of the reference's dispatch surface, on top of the C ABI.

`CloudRenderer` plays the role of Meteoros' `Renderer` for the hot path (Renderer.cpp:122-192, 653-722, 823-846):
it owns the images, takes the uniforms the reference keeps in Camera / Scene / Sky, and issues the passes in the
reference's order.  Method names follow the reference (`frame`, `dispatch_*`, `swap_ping_pong`).  All work is
done by libmeteoros_b200.so; nothing here computes pixels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .scene import CAMERA_DTYPE, SUNSKY_DTYPE, TIME_DTYPE, TUNING_DTYPE

MT_OK = 0
STORAGE_F32, STORAGE_F16_EMULATE, STORAGE_F16 = 0, 1, 2
FLAG_COUNTERS, FLAG_PASS_TIMING, FLAG_SEQUENTIAL_MARCH, FLAG_TOP_DOWN, FLAG_NO_CONE_RF, FLAG_SPLIT_MARCH, FLAG_NO_FUSED_TONEMAP = 1, 2, 4, 8, 16, 32, 64
FLAG_HW_CONE_FILTER = 128  # opt-in, outside the parity bar: light-cone samples of full-quality dispatches through the texture unit (meteoros_b200.h)
TEX_LOW_FREQ, TEX_HIGH_FREQ, TEX_CURL, TEX_WEATHER = 0, 1, 2, 3
IMAGE_CLOUD_CUR, IMAGE_CLOUD_PREV, IMAGE_GODRAY_MASK, IMAGE_LDR, IMAGE_LDR_PREV = 0, 1, 2, 3, 4
PASS_REPROJECT, PASS_CLOUD, PASS_GODRAYS, PASS_TONEMAP, PASS_TXAA = 0, 1, 2, 3, 4
FRAME_GODRAYS, FRAME_TONEMAP, FRAME_TXAA = 1, 2, 4
STORE_DIRECT, STORE_BULK = 0, 1

RAY_DEBUG_DTYPE = np.dtype(
    [
        ("dir", "<f4", (3,)),
        ("t_in", "<f4"),
        ("t_out", "<f4"),
        ("step_size", "<f4"),
        ("branch", "<i4"),
        ("steps", "<i4"),
        ("jitter_hash", "<u4"),
        ("accum", "<f4"),
    ]
)
assert RAY_DEBUG_DTYPE.itemsize == 40


class MeteorosError(RuntimeError):
    """Raised for any non-MT_OK status (the reference throws std::runtime_error, Renderer.cpp:139-141)."""

    def __init__(self, status: int, what: str, detail: str):
        self.status = status
        super().__init__(f"{what}: {detail}")


def _as_bytes(rec: np.ndarray, dtype: np.dtype) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(rec, dtype=dtype))
    if a.nbytes != dtype.itemsize:
        raise ValueError(f"expected one {dtype} record")
    return a


class CloudRenderer:
    def __init__(self, width: int, height: int, device: int = 0, storage: int = STORAGE_F32, flags: int = 0):
        self._lib = _lib.load()
        self.width, self.height = int(width), int(height)
        self.storage = int(storage)
        cfg = _lib.MtConfig(C.sizeof(_lib.MtConfig), self.width, self.height, int(device), int(storage), int(flags))
        h = C.c_void_p()
        st = self._lib.mtCreate(C.byref(cfg), C.byref(h))
        if st != MT_OK:
            raise MeteorosError(st, "mtCreate", self._lib.mtStatusString(st).decode())
        self._h = h

    # ---- plumbing -------------------------------------------------------------------------------------------
    def _check(self, st: int, what: str):
        if st != MT_OK:
            raise MeteorosError(st, what, f"{self._lib.mtStatusString(st).decode()}: {self._lib.mtGetLastError(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mtDestroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- uniforms -------------------------------------------------------------------------------------------
    def set_camera(self, ubo):
        a = _as_bytes(ubo, CAMERA_DTYPE)
        self._check(self._lib.mtSetCamera(self._h, a.ctypes.data), "mtSetCamera")

    def set_camera_old(self, ubo):
        a = _as_bytes(ubo, CAMERA_DTYPE)
        self._check(self._lib.mtSetCameraOld(self._h, a.ctypes.data), "mtSetCameraOld")

    def set_time(self, ubo):
        a = _as_bytes(ubo, TIME_DTYPE)
        self._check(self._lib.mtSetTime(self._h, a.ctypes.data), "mtSetTime")

    def set_sun_and_sky(self, ubo):
        a = _as_bytes(ubo, SUNSKY_DTYPE)
        self._check(self._lib.mtSetSunAndSky(self._h, a.ctypes.data), "mtSetSunAndSky")

    def resize(self, width: int, height: int):
        """mtResize: new images of the new size (transactional: on failure the context keeps the old ones); every image starts cleared,
        device pointers and exported handles of the old images are invalid afterwards."""
        self._check(self._lib.mtResize(self._h, int(width), int(height)), "mtResize")
        self.width, self.height = int(width), int(height)

    def set_key_press_query(self, key_debug: int):
        self._check(self._lib.mtSetKeyPressQuery(self._h, int(key_debug)), "mtSetKeyPressQuery")

    def set_tuning(self, tuning):
        a = _as_bytes(tuning, TUNING_DTYPE)
        self._check(self._lib.mtSetTuning(self._h, a.ctypes.data), "mtSetTuning")

    def default_tuning(self) -> np.ndarray:
        t = np.zeros((), TUNING_DTYPE)
        self._lib.mtDefaultTuning(t.ctypes.data)
        return t

    # ---- textures -------------------------------------------------------------------------------------------
    def upload_texture_3d(self, slot: int, vol: np.ndarray):
        v = np.ascontiguousarray(vol, dtype=np.uint8)
        if v.ndim != 4 or v.shape[3] != 4:
            raise ValueError("3D texture must be [d][h][w][4] uint8")
        d, h, w, _ = v.shape
        self._check(self._lib.mtUploadTexture3D(self._h, slot, w, h, d, v.ctypes.data), "mtUploadTexture3D")

    def upload_texture_2d(self, slot: int, img: np.ndarray):
        v = np.ascontiguousarray(img, dtype=np.uint8)
        if v.ndim != 3 or v.shape[2] != 4:
            raise ValueError("2D texture must be [h][w][4] uint8")
        h, w, _ = v.shape
        self._check(self._lib.mtUploadTexture2D(self._h, slot, w, h, v.ctypes.data), "mtUploadTexture2D")

    def upload_noise(self, noise: dict):
        """Sky::CreateCloudResources (Sky.cpp:25-58)."""
        self.upload_texture_3d(TEX_LOW_FREQ, noise["low"])
        self.upload_texture_3d(TEX_HIGH_FREQ, noise["high"])
        self.upload_texture_2d(TEX_CURL, noise["curl"])
        if "weather" in noise:
            self.upload_texture_2d(TEX_WEATHER, noise["weather"])

    # ---- dispatches -----------------------------------------------------------------------------------------
    def dispatch_reprojection(self):
        self._check(self._lib.mtDispatchReprojection(self._h), "mtDispatchReprojection")

    def dispatch_cloud(self):
        self._check(self._lib.mtDispatchCloud(self._h), "mtDispatchCloud")

    def dispatch_cloud_full(self):
        self._check(self._lib.mtDispatchCloudFull(self._h), "mtDispatchCloudFull")

    def dispatch_cloud_tiles(self, tile_rows: int, tile_begin: int, tile_end: int, tile_stride: int):
        self._check(self._lib.mtDispatchCloudTiles(self._h, tile_rows, tile_begin, tile_end, tile_stride), "mtDispatchCloudTiles")

    def dispatch_cloud_debug(self, full: bool) -> np.ndarray:
        out = np.zeros((self.height, self.width), RAY_DEBUG_DTYPE)
        self._check(self._lib.mtDispatchCloudDebug(self._h, int(bool(full)), out.ctypes.data, out.nbytes), "mtDispatchCloudDebug")
        return out

    def dispatch_god_rays(self):
        self._check(self._lib.mtDispatchGodRays(self._h), "mtDispatchGodRays")

    def dispatch_tone_map(self):
        self._check(self._lib.mtDispatchToneMap(self._h), "mtDispatchToneMap")

    def dispatch_reprojection_debug(self) -> np.ndarray:
        taps = np.zeros((self.height, self.width, 10), np.int32)
        self._check(self._lib.mtDispatchReprojectionDebug(self._h, taps.ctypes.data, taps.nbytes), "mtDispatchReprojectionDebug")
        return taps

    def dispatch_txaa(self):
        self._check(self._lib.mtDispatchTXAA(self._h), "mtDispatchTXAA")

    def frame(self, with_godrays: bool = False, with_txaa: bool = False):
        """Renderer::Frame: REPROJ, CLOUD, [GODRAYS], TONEMAP, [TXAA], swap."""
        passes = FRAME_TONEMAP | (FRAME_GODRAYS if with_godrays else 0) | (FRAME_TXAA if with_txaa else 0)
        self._check(self._lib.mtFrameEx(self._h, passes), "mtFrameEx")

    def swap_ping_pong(self):
        self._check(self._lib.mtSwapPingPong(self._h), "mtSwapPingPong")

    def synchronize(self):
        self._check(self._lib.mtSynchronize(self._h), "mtSynchronize")

    # ---- images ---------------------------------------------------------------------------------------------
    def _image_shape(self, which: int):
        if which in (IMAGE_LDR, IMAGE_LDR_PREV):
            return (self.height, self.width, 4), np.uint8
        return (self.height, self.width, 4), (np.float16 if self.storage == STORAGE_F16 else np.float32)

    def read_image(self, which: int, out: np.ndarray | None = None) -> np.ndarray:
        shape, dt = self._image_shape(which)
        if out is None:
            out = np.empty(shape, dt)
        assert out.dtype == dt and out.flags.c_contiguous and out.size == np.prod(shape)
        self._check(self._lib.mtReadImage(self._h, which, out.ctypes.data, out.nbytes), "mtReadImage")
        return out

    def read_image_rows(self, which: int, row_begin: int, row_end: int) -> np.ndarray:
        _, dt = self._image_shape(which)
        out = np.empty((row_end - row_begin, self.width, 4), dt)
        self._check(self._lib.mtReadImageRows(self._h, which, row_begin, row_end, out.ctypes.data, out.nbytes), "mtReadImageRows")
        return out

    def write_image(self, which: int, data: np.ndarray):
        shape, dt = self._image_shape(which)
        a = np.ascontiguousarray(data, dtype=dt)
        if a.shape != shape:
            raise ValueError(f"image must have shape {shape}")
        self._check(self._lib.mtWriteImage(self._h, which, a.ctypes.data, a.nbytes), "mtWriteImage")
        self.synchronize()  # `a` may be a temporary

    def write_image_async(self, which: int, host_ptr: int, nbytes: int):
        """H2D from caller-owned (ideally pinned) memory; ordered on the context's stream, no sync."""
        self._check(self._lib.mtWriteImage(self._h, which, C.c_void_p(host_ptr), nbytes), "mtWriteImage")

    def read_image_into(self, which: int, host_ptr: int, nbytes: int):
        self._check(self._lib.mtReadImage(self._h, which, C.c_void_p(host_ptr), nbytes), "mtReadImage")

    def read_image_async(self, which: int, host_ptr: int, nbytes: int):
        """D2H on the copy stream, overlapping later dispatches; valid after wait_reads()."""
        self._check(self._lib.mtReadImageAsync(self._h, which, C.c_void_p(host_ptr), nbytes), "mtReadImageAsync")

    def read_godray_grey_async(self, host_ptr: int, nbytes: int):
        """The god-ray image as one float32 per pixel (decoded on the device), D2H on the copy stream; valid after wait_reads()."""
        self._check(self._lib.mtReadGodRayGreyAsync(self._h, C.c_void_p(host_ptr), nbytes), "mtReadGodRayGreyAsync")

    def read_godray_grey(self) -> np.ndarray:
        out = np.empty((self.height, self.width), np.float32)
        self.read_godray_grey_async(out.ctypes.data, out.nbytes)
        self.wait_reads()
        return out

    def wait_reads(self):
        self._check(self._lib.mtWaitReads(self._h), "mtWaitReads")

    def join_copies(self):
        self._check(self._lib.mtJoinCopies(self._h), "mtJoinCopies")

    def clear_images(self):
        self._check(self._lib.mtClearImages(self._h), "mtClearImages")

    def image_device_ptr(self, which: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.mtImageDevicePtr(self._h, which, C.byref(p)), "mtImageDevicePtr")
        return int(p.value)

    def set_cloud_output(self, hdr_ptr: int | None, mask_ptr: int | None):
        self._check(self._lib.mtSetCloudOutput(self._h, C.c_void_p(hdr_ptr or 0), C.c_void_p(mask_ptr or 0)), "mtSetCloudOutput")


    def set_cloud_forward(self, hdr_ptr):
        """Gather by forwarding: finished row tiles of dispatch_cloud_tiles are pushed to this peer image by a side kernel."""
        self._check(self._lib.mtSetCloudForward(self._h, C.c_void_p(hdr_ptr or 0)), "mtSetCloudForward")
    def set_cloud_store_mode(self, mode: int):
        """0 = direct 16-byte stores, 1 = staged bulk asynchronous copies (see the header)."""
        self._check(self._lib.mtSetCloudStoreMode(self._h, int(mode)), "mtSetCloudStoreMode")

    def export_image_handle(self, which: int) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._check(self._lib.mtExportImageHandle(self._h, which, buf), "mtExportImageHandle")
        return bytes(buf)

    def open_peer_image(self, handle: bytes) -> int:
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        self._check(self._lib.mtOpenPeerImage(self._h, buf, C.byref(p)), "mtOpenPeerImage")
        return int(p.value)

    def close_peer_image(self, ptr: int):
        self._check(self._lib.mtClosePeerImage(self._h, C.c_void_p(ptr)), "mtClosePeerImage")

    def copy_tiles_to_peer(self, which: int, tile_rows: int, tile_begin: int, tile_end: int, tile_stride: int, peer_ptr: int):
        self._check(self._lib.mtCopyTilesToPeer(self._h, which, tile_rows, tile_begin, tile_end, tile_stride, C.c_void_p(peer_ptr)),
                    "mtCopyTilesToPeer")

    # ---- measurement ----------------------------------------------------------------------------------------
    def counters(self, reset: bool = True) -> dict:
        c = _lib.MtCounters()
        self._check(self._lib.mtGetCounters(self._h, C.byref(c), int(reset)), "mtGetCounters")
        return c.as_dict()

    def last_pass_ms(self, which: int) -> float:
        ms = C.c_float()
        self._check(self._lib.mtLastPassMs(self._h, which, C.byref(ms)), "mtLastPassMs")
        return float(ms.value)

    def event_record(self, slot: int):
        self._check(self._lib.mtEventRecord(self._h, slot), "mtEventRecord")

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._check(self._lib.mtEventElapsedMs(self._h, a, b, C.byref(ms)), "mtEventElapsedMs")
        return float(ms.value)

    def stream_handle(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.mtStreamHandle(self._h, C.byref(p)), "mtStreamHandle")
        return int(p.value or 0)

    def flush_l2(self, nbytes: int = 0):
        self._check(self._lib.mtFlushL2(self._h, nbytes), "mtFlushL2")

    def measure_fp32_peak_gflops(self) -> float:
        g = C.c_float()
        self._check(self._lib.mtMeasureFp32Peak(self._h, C.byref(g)), "mtMeasureFp32Peak")
        return float(g.value)

    def launch_count(self) -> int:
        return int(self._lib.mtLaunchCount(self._h))
