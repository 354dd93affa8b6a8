"""pytest fixture giving access to the host build of the kernel cores (tests/hostsim)."""
import importlib.util
import sys
from pathlib import Path

import pytest

_spec = importlib.util.spec_from_file_location("mt_hostsim", Path(__file__).resolve().parent / "hostsim" / "__init__.py")
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mt_hostsim"] = _mod
_spec.loader.exec_module(_mod)


@pytest.fixture(scope="session")
def hostsim():
    _mod.lib()
    return _mod
