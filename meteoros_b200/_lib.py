"""ctypes loader for libmeteoros_b200.so -- the C ABI declared in include/meteoros_b200.h.

The library is built in-tree (meteoros_b200/libmeteoros_b200.so) by `make -C meteoros_b200/csrc` /
`__graft_entry__.build()`.  There is no fallback: if it is missing or does not load, importing the
dispatch API raises.  Every prototype below mirrors one declaration of the header; tests/test_abi.py checks that
the set of names here equals the set of MT_API declarations in the header and the exported dynamic symbols.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# METEOROS_B200_LIB: developer override used by tools/ab_bench.sh to A/B kernel variants; normal use loads the in-tree build
LIB_PATH = Path(os.environ.get("METEOROS_B200_LIB") or (Path(__file__).resolve().parent / "libmeteoros_b200.so"))

c_void_pp = C.POINTER(C.c_void_p)


class MtConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("device", C.c_int32),
        ("storage", C.c_uint32),
        ("flags", C.c_uint32),
    ]


class MtCounters(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64),
        ("rays_marched", C.c_uint64),
        ("steps", C.c_uint64),
        ("steps_incloud", C.c_uint64),
        ("cone_hits", C.c_uint64),
        ("early_exits", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class MtxCamera(C.Structure):
    _fields_ = [
        ("eye", C.c_float * 3), ("ref", C.c_float * 3), ("forward", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3),
        ("fovy_deg", C.c_float), ("aspect", C.c_float), ("near_clip", C.c_float), ("far_clip", C.c_float),
        ("width", C.c_int32), ("height", C.c_int32),
    ]


# name -> (restype, argtypes); the single source of truth for the Python side of the ABI
PROTOTYPES = {
    "mtAbiVersion": (C.c_uint32, []),
    "mtStatusString": (C.c_char_p, [C.c_int]),
    "mtDefaultTuning": (None, [C.c_void_p]),
    "mtCreate": (C.c_int, [C.POINTER(MtConfig), c_void_pp]),
    "mtDestroy": (None, [C.c_void_p]),
    "mtGetLastError": (C.c_char_p, [C.c_void_p]),
    "mtResize": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "mtSetCamera": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mtSetCameraOld": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mtSetTime": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mtSetSunAndSky": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mtSetKeyPressQuery": (C.c_int, [C.c_void_p, C.c_int32]),
    "mtSetTuning": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mtUploadTexture3D": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "mtUploadTexture2D": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]),
    "mtDispatchReprojection": (C.c_int, [C.c_void_p]),
    "mtDispatchCloud": (C.c_int, [C.c_void_p]),
    "mtDispatchCloudFull": (C.c_int, [C.c_void_p]),
    "mtDispatchCloudTiles": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "mtDispatchCloudDebug": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "mtDispatchGodRays": (C.c_int, [C.c_void_p]),
    "mtDispatchToneMap": (C.c_int, [C.c_void_p]),
    "mtDispatchTXAA": (C.c_int, [C.c_void_p]),
    "mtFrameEx": (C.c_int, [C.c_void_p, C.c_uint32]),
    "mtDispatchReprojectionDebug": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mtFrame": (C.c_int, [C.c_void_p, C.c_int]),
    "mtSwapPingPong": (C.c_int, [C.c_void_p]),
    "mtSynchronize": (C.c_int, [C.c_void_p]),
    "mtImageBytes": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]),
    "mtReadImage": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "mtReadImageRows": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]),
    "mtReadImageAsync": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "mtReadGodRayGreyAsync": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mtWaitReads": (C.c_int, [C.c_void_p]),
    "mtJoinCopies": (C.c_int, [C.c_void_p]),
    "mtWriteImage": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "mtClearImages": (C.c_int, [C.c_void_p]),
    "mtImageDevicePtr": (C.c_int, [C.c_void_p, C.c_int, c_void_pp]),
    "mtSetCloudOutput": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mtSetCloudForward": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mtSetCloudStoreMode": (C.c_int, [C.c_void_p, C.c_int]),
    "mtExportImageHandle": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "mtOpenPeerImage": (C.c_int, [C.c_void_p, C.c_void_p, c_void_pp]),
    "mtClosePeerImage": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mtxCameraInit": (None, [C.POINTER(MtxCamera), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float]),
    "mtxCameraRotateAboutUp": (None, [C.POINTER(MtxCamera), C.c_float]),
    "mtxCameraRotateAboutRight": (None, [C.POINTER(MtxCamera), C.c_float]),
    "mtxCameraTranslateAlongLook": (None, [C.POINTER(MtxCamera), C.c_float]),
    "mtxCameraTranslateAlongRight": (None, [C.POINTER(MtxCamera), C.c_float]),
    "mtxCameraTranslateAlongUp": (None, [C.POINTER(MtxCamera), C.c_float]),
    "mtxCameraUBO": (None, [C.POINTER(MtxCamera), C.c_void_p]),
    "mtxTimeInit": (None, [C.c_void_p]),
    "mtxTimeUpdate": (None, [C.c_void_p, C.c_float]),
    "mtxSunAndSky": (None, [C.c_void_p]),
    "mtxRunFrame": (C.c_int, [C.c_void_p, C.POINTER(MtxCamera), C.c_void_p, C.c_void_p, C.c_float, C.c_uint32]),
    "mtxDecodeImage": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "mtxLoadImageFile": (C.c_int, [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "mtxLoadVolumeFromSlices": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]),
    "mtxSaveVolume": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "mtxLoadVolume": (C.c_int, [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "mtCopyTilesToPeer": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "mtGetCounters": (C.c_int, [C.c_void_p, C.POINTER(MtCounters), C.c_int]),
    "mtLastPassMs": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "mtStreamHandle": (C.c_int, [C.c_void_p, c_void_pp]),
    "mtEventRecord": (C.c_int, [C.c_void_p, C.c_uint32]),
    "mtEventElapsedMs": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]),
    "mtFlushL2": (C.c_int, [C.c_void_p, C.c_size_t]),
    "mtMeasureFp32Peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "mtLaunchCount": (C.c_uint64, [C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library (no GPU needed to load it; creating a context needs one)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} not built -- run `make -C meteoros_b200/csrc` (or __graft_entry__.build()). "
                "meteoros_b200 has no CPU fallback."
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
