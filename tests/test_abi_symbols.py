"""The C-ABI library loads without a GPU and exports every entry point that
include/b2exec.h declares (no compute calls here)."""

import ctypes
import re
from pathlib import Path

from covalent_ssh_plugin_b200 import _b2lib as L

HEADER = Path(__file__).resolve().parents[1] / "include" / "b2exec.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    names = declared_functions()
    for must in ("b2_init", "b2_shutdown", "b2_pack", "b2_unpack", "b2_move_spans", "b2_register",
                 "b2_enqueue", "b2_launch", "b2_poll", "b2_wait", "b2_peer_copy", "b2_peer_pack",
                 "b2_ipc_export", "b2_ipc_open", "b2_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(built_library):
    lib = ctypes.CDLL(str(built_library))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in b2exec.h but not exported"


def test_binding_covers_every_declared_symbol(built_library):
    assert set(declared_functions()) == set(L.EXPORTED_SYMBOLS)
    lib = L.load()
    assert lib.b2_abi_version() == L.ABI_VERSION


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(L.Task) == 64 and ctypes.sizeof(L.Span) == 24
    import numpy as np

    assert np.dtype(L.TASK_DTYPE).itemsize == 64 and np.dtype(L.SPAN_DTYPE).itemsize == 24
    assert [n for n, _ in L.TASK_DTYPE] == [f[0] for f in L.Task._fields_]


def test_calls_fail_loudly_without_a_device(built_library, gpu_count):
    """No CPU implementation hides behind the ABI: on a box without a GPU every
    device call reports an error instead of computing anything."""
    import pytest

    if gpu_count:
        pytest.skip("this box has a GPU")
    with pytest.raises(L.B2Error):
        L.call("b2_init", 0, 1 << 20, 1 << 20, 1 << 20, 1024)
    with pytest.raises(L.B2Error):
        L.call("b2_sync", 0)
