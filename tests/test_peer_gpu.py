"""K4 on two GPUs of one process: the span mover writing straight into a peer
arena over NVLink (b2_peer_pack) and the copy-engine path (b2_peer_copy), both
checked byte for byte.  Skipped when the box has fewer than two GPUs."""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def two_devices(gpu_count):
    if gpu_count < 2:
        pytest.skip("needs two GPUs")
    from covalent_ssh_plugin_b200 import _b2lib as L

    arena = 1 << 30
    started = []
    for d in (0, 1):
        if not L.load().b2_is_initialized(d):
            L.call("b2_init", d, arena, 16 << 20, 16 << 20, 1 << 12)
            started.append(d)
    yield L, arena
    for d in started:
        L.call("b2_shutdown", d)


def test_peer_pack_and_peer_copy(two_devices):
    L, arena = two_devices
    aid = C.c_uint32()
    L.call("b2_peer_enable", 0, 1, C.byref(aid))
    assert aid.value >= 1
    rng = np.random.default_rng(3)
    n, sz = 500, 65536
    data = rng.integers(0, 256, n * sz, dtype=np.uint8)
    L.call("b2_write", 0, 0, data.ctypes.data, data.size)
    perm = rng.permutation(n)
    spans = np.zeros((n, 3), dtype=np.uint64)
    spans[:, 0] = perm * sz                 # scattered on GPU 0
    spans[:, 1] = (1 << 20) + np.arange(n) * sz   # contiguous on GPU 1
    spans[:, 2] = sz
    L.call("b2_peer_pack", 0, aid.value, spans.ctypes.data, n)
    L.call("b2_sync", 0)
    got = np.empty(n * sz, dtype=np.uint8)
    L.call("b2_read", 1, 1 << 20, got.ctypes.data, got.size)
    assert np.array_equal(got.reshape(n, sz), data.reshape(n, sz)[perm])

    L.call("b2_peer_copy", 0, aid.value, 0, 512 << 20, n * sz)
    L.call("b2_sync", 0)
    L.call("b2_read", 1, 512 << 20, got.ctypes.data, got.size)
    assert np.array_equal(got, data)


def test_worker_reads_operand_from_peer_arena(two_devices):
    """Fused route + compute: GPU 0's worker runs scale_add on an operand that
    lives in GPU 1's arena (peer loads over NVLink), result local."""
    L, arena = two_devices
    aid = C.c_uint32()
    L.call("b2_peer_enable", 0, 1, C.byref(aid))
    L.call("b2_register", 0, 2, L.K_SCALE_ADD, 2.0, 1.0, 0, 0)
    x = np.random.default_rng(9).standard_normal(1 << 18, dtype=np.float32)
    L.call("b2_write", 1, 64 << 20, x.ctypes.data, x.nbytes)
    t = np.zeros(1, dtype=L.TASK_DTYPE)
    t["fn_id"], t["arg0"], t["out"], t["count"] = 2, 64 << 20, 128 << 20, x.size
    t["flags"] = aid.value << L.F_ARENA0_SHIFT
    L.call("b2_enqueue", 0, t.ctypes.data, 1, None, 0, None)
    ticket = C.c_uint64()
    L.call("b2_launch", 0, 0, C.byref(ticket))
    L.call("b2_wait", 0, ticket.value)
    out = np.empty_like(x)
    L.call("b2_read", 0, 128 << 20, out.ctypes.data, out.nbytes)
    assert np.array_equal(out, np.float32(2.0) * x + np.float32(1.0))
