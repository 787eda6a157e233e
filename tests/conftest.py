"""Shared pytest setup.

* registers the ``gpu`` marker; GPU tests call the CUDA library through the C ABI
  and are skipped only when the box has no CUDA device (a missing or unloadable
  library is a FAILURE, never a skip);
* builds libb2exec.so once per session (nvcc cross-compiles without a GPU).
"""

import ctypes
import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))
os.environ.setdefault("HOME", "/root")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (runs the CUDA path through the C ABI)")


@pytest.fixture(scope="session", autouse=True)
def built_library():
    from covalent_ssh_plugin_b200.build import build_all

    return build_all()


def _cuda_device_count() -> int:
    from covalent_ssh_plugin_b200 import _b2lib as L

    n = ctypes.c_int(0)
    rc = L.load().b2_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0


@pytest.fixture(scope="session")
def gpu_count(built_library):
    return _cuda_device_count()


@pytest.fixture(scope="module")
def raw_device(gpu_count):
    """A bare b2_init'ed device 0 for kernel-level tests (no Python engine)."""
    if gpu_count < 1:
        pytest.skip("no CUDA device on this box")
    from covalent_ssh_plugin_b200 import _b2lib as L

    arena = 3 << 30
    L.call("b2_init", 0, arena, 64 << 20, 64 << 20, 1 << 16)
    yield {"dev": 0, "arena": arena, "L": L}
    L.call("b2_shutdown", 0)
