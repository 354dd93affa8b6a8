"""The C-ABI library loads without a GPU and exports exactly what include/meteoros_b200.h declares."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from meteoros_b200 import _lib, scene

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "meteoros_b200.h"


def declared_symbols():
    text = HEADER.read_text()
    return set(re.findall(r"MT_API\s+[\w\s\*]+?\b(mtx?[A-Z]\w+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = declared_symbols()
    assert len(declared) >= 40
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert declared <= exported
    # nothing but the ABI leaks out of the library (-fvisibility=hidden)
    assert {s for s in exported if s.startswith("mt")} == declared


def test_abi_basics_without_gpu():
    lib = _lib.load()
    assert lib.mtAbiVersion() == 1
    assert lib.mtStatusString(0) == b"MT_OK" and lib.mtStatusString(4) == b"MT_ERR_UNSUPPORTED_ARCH"
    t = np.zeros((), scene.TUNING_DTYPE)
    lib.mtDefaultTuning(t.ctypes.data)
    assert t.tobytes() == scene.default_tuning().tobytes()  # header defaults == the shader literals == Python mirror
    assert C.sizeof(_lib.MtConfig) == 24 and C.sizeof(_lib.MtCounters) == 48
    # argument validation happens before any CUDA call
    h = C.c_void_p()
    assert lib.mtCreate(None, C.byref(h)) == 1
    bad = _lib.MtConfig(0, 64, 36, 0, 0, 0)
    assert lib.mtCreate(C.byref(bad), C.byref(h)) == 1 and not h.value
    assert lib.mtDispatchCloud(None) == 1 and lib.mtFrame(None, 0) == 1
    assert lib.mtGetLastError(None) == b"null context"


def test_no_cpu_fallback_in_product():
    """The product package never touches the oracle or the host simulation."""
    for p in (ROOT / "meteoros_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp"):
            txt = p.read_text()
            assert "import oracle" not in txt and "from oracle" not in txt, p
            assert "libmeteoros_oracle" not in txt and "libhostsim" not in txt, p
    deps = subprocess.run(["ldd", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle" not in deps and "hostsim" not in deps


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from meteoros_b200.api import CloudRenderer, MeteorosError

    with pytest.raises(MeteorosError):
        CloudRenderer(64, 36)


def test_header_is_plain_c_and_the_example_links(tmp_path):
    """include/meteoros_b200.h must be consumable from C (the reference-side binding is C/C++): compile and link
    examples/frame_loop.c with gcc -std=c11 and run it without a GPU -- it must fail loudly at mtCreate, not fall back."""
    import shutil

    cc = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else shutil.which("gcc")
    exe = tmp_path / "frame_loop"
    r = subprocess.run([cc, "-std=c11", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "frame_loop.c"),
                        f"-L{_lib.LIB_PATH.parent}", "-lmeteoros_b200", f"-Wl,-rpath,{_lib.LIB_PATH.parent}", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref = Path("/root/reference/src/CloudScapes/textures/CloudTextures")
    import torch

    if ref.exists() and not torch.cuda.is_available():
        run = subprocess.run([str(exe), str(ref), "2", "64", "36"], capture_output=True, text=True)
        assert "assets decoded" in run.stdout            # the C++ decoders ran on the reference's files
        assert run.returncode == 3 and "no CPU path" in run.stderr


def test_python_binding_constants_match_the_header():
    """The MT_FLAG_* values the ctypes binding passes are the header's (a flag added on one side only would silently select nothing)."""
    import re

    from meteoros_b200 import api

    text = (ROOT / "include" / "meteoros_b200.h").read_text()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+MT_FLAG_(\w+)\s+(\d+)u", text)}
    assert flags, "no MT_FLAG_* definitions found in the header"
    for name, value in flags.items():
        if name == "PASS_TIMING_INTERNAL":
            continue
        assert getattr(api, "FLAG_" + name) == value, f"MT_FLAG_{name}"
    assert len(set(flags.values())) == len(flags), "two flags share a bit"
