"""The device 'connection' can be closed and reopened (own module: needs device 0 free)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_shutdown_and_reinit(gpu_count):
    """The 'connection' can be closed and reopened (ssh.py:585-587 / 263-268)."""
    if gpu_count < 1:
        pytest.skip("no CUDA device on this box")
    from covalent_ssh_plugin_b200 import _b2lib as L

    if L.load().b2_is_initialized(0):
        pytest.skip("device 0 is held by another fixture")
    for _ in range(3):
        L.call("b2_init", 0, 64 << 20, 1 << 20, 1 << 20, 1024)
        x = np.arange(1024, dtype=np.float32)
        L.call("b2_write", 0, 0, x.ctypes.data, x.nbytes)
        y = np.empty_like(x)
        L.call("b2_read", 0, 0, y.ctypes.data, y.nbytes)
        assert np.array_equal(x, y)
        L.call("b2_shutdown", 0)
    with pytest.raises(L.B2Error):
        L.call("b2_sync", 0)
