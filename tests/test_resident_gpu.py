"""GPU tests of the resident worker (b2_resident_enable): the kernel stays on the GPU across
batches, batches arrive through a descriptor ring + doorbell written by stream memory operations,
completion is a word the streams wait on.  Same results as the launch-per-batch path, bit for bit."""

import ctypes as C
import time

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]

GiB, MiB = 1 << 30, 1 << 20


@pytest.fixture(scope="module")
def res_device(gpu_count):
    if gpu_count < 1:
        pytest.skip("no CUDA device on this box")
    from covalent_ssh_plugin_b200 import _b2lib as L
    from covalent_ssh_plugin_b200.engine import shutdown_all

    shutdown_all()
    L.call("b2_init", 0, 3 * GiB, 64 * MiB, 64 * MiB, 1 << 16)
    L.call("b2_register", 0, 21, L.K_SCALE_ADD, 2.0, 1.0, 0, 0)
    L.call("b2_register", 0, 22, L.K_COPY, 0.0, 0.0, 0, 0)
    L.call("b2_register", 0, 23, L.K_IADD, 0.0, 0.0, 0, 0)
    L.call("b2_register", 0, 24, L.K_STACK_SUM, 0.0, 0.0, 0, 0)
    L.call("b2_resident_enable", 0, 300)
    yield {"dev": 0, "L": L}
    L.call("b2_shutdown", 0)


def _stats(L, dev):
    s = L.Stats()
    L.call("b2_get_stats", dev, C.byref(s))
    return s


def _host_out(L, dev):
    ptr, size = C.c_void_p(), C.c_uint64()
    L.call("b2_host_buffer", dev, 1, C.byref(ptr), C.byref(size))
    return np.frombuffer((C.c_uint8 * size.value).from_address(ptr.value), dtype=np.uint8)


def _fanout(L, n, elems, in0, out0):
    slot = (elems * 4 + 255) & ~255
    t = np.zeros(n, dtype=L.TASK_DTYPE)
    t["fn_id"], t["count"] = 21, elems
    t["arg0"] = in0 + np.arange(n, dtype=np.uint64) * slot
    t["out"] = out0 + np.arange(n, dtype=np.uint64) * slot
    return t, slot


def test_back_to_back_batches_without_launches(res_device):
    """60 batches of 700 electrons published back to back: one (or very few) kernel instances serve
    all of them, every result is the model's."""
    L, dev = res_device["L"], res_device["dev"]
    n, elems = 700, 16384
    tasks, slot = _fanout(L, n, elems, 256 * MiB, 1 * GiB)
    src = np.random.default_rng(0).standard_normal(n * slot // 4, dtype=np.float32)
    L.call("b2_write", dev, 256 * MiB, src.ctypes.data, src.nbytes)
    s0 = _stats(L, dev)
    t = C.c_uint64()
    h_out = _host_out(L, dev)
    for rep in range(60):
        L.call("b2_enqueue", dev, tasks.ctypes.data, n, None, 0, None)
        L.call("b2_launch", dev, 4096 * (rep % 8), C.byref(t))
    L.call("b2_wait", dev, t.value)
    s1 = _stats(L, dev)
    assert s1.resident_batches - s0.resident_batches == 60
    assert 1 <= s1.resident_instances - s0.resident_instances <= 6
    assert (h_out[4096 * 3:4096 * 3 + 4 * n].view(np.uint32) == L.TS_OK).all()
    got = np.empty(n * slot // 4, dtype=np.float32)
    L.call("b2_read", dev, 1 * GiB, got.ctypes.data, got.nbytes)
    for i in (0, n // 2, n - 1):
        o = i * slot // 4
        assert np.array_equal(got[o:o + elems], np.float32(2.0) * src[o:o + elems] + np.float32(1.0)), i


def test_reused_table_slots_never_go_stale(res_device):
    """The eight task-table slots are reused cyclically and — for batches that queue nothing else — are
    guarded by the batch's sequence number instead of an event.  Every batch here carries different
    records at addresses an earlier batch also used: a slot handed out too early, or a stale cached
    line, would show up as a wrong sum."""
    L, dev = res_device["L"], res_device["dev"]
    n, reps = 777, 48
    base = 2 * GiB
    t = C.c_uint64()
    tickets = []
    for rep in range(reps):
        tasks = np.zeros(n, dtype=L.TASK_DTYPE)
        tasks["fn_id"], tasks["flags"], tasks["count"] = 23, L.F_IMM0 | L.F_IMM1, 1
        tasks["arg0"], tasks["arg1"] = np.arange(n) * 3 + rep, 1000 * rep
        tasks["out"] = base + 8 * (rep * n + np.arange(n))
        L.call("b2_enqueue", dev, tasks.ctypes.data, n, None, 0, None)
        L.call("b2_launch", dev, L.NO_STATUS_COPY, C.byref(t))
        tickets.append(t.value)
    for tk in tickets[::7] + [tickets[-1]]:
        L.call("b2_wait", dev, tk)
    assert L.load().b2_poll(dev, tickets[0]) == 0
    got = np.empty(reps * n, dtype=np.int64)
    L.call("b2_read", dev, base, got.ctypes.data, got.nbytes)
    want = np.concatenate([np.arange(n) * 3 + rep + 1000 * rep for rep in range(reps)])
    assert np.array_equal(got, want)


def test_dependencies_inside_a_resident_batch(res_device):
    """A 300-long chain of integer adds plus a fan-out/fan-in diamond, released on the device through
    the batch's own queue cells — in resident mode those are per-batch state in the control block."""
    L, dev = res_device["L"], res_device["dev"]
    n = 300
    base = 2 * GiB
    for rep in range(4):
        tasks = np.zeros(n, dtype=L.TASK_DTYPE)
        tasks["fn_id"], tasks["flags"], tasks["count"] = 23, L.F_IMM0 | L.F_IMM1, 1
        tasks["arg0"], tasks["arg1"] = np.arange(n) + rep, 5
        tasks["out"] = base + 8 * np.arange(n)
        order = np.random.default_rng(rep).permutation(n)       # the chain visits the table in scrambled order
        succ = np.zeros(n - 1, dtype=np.uint32)
        for k in range(n - 1):
            a, b = order[k], order[k + 1]
            tasks["succ_begin"][a], tasks["succ_count"][a] = k, 1
            succ[k] = b
            tasks["deps"][b] = 1
        L.call("b2_fill", dev, base, 0, 8 * n)
        L.call("b2_enqueue", dev, tasks.ctypes.data, n, succ.ctypes.data, n - 1, None)
        t = C.c_uint64()
        L.call("b2_launch", dev, 0, C.byref(t))
        L.call("b2_wait", dev, t.value)
        h_out = _host_out(L, dev)
        assert (h_out[:4 * n].view(np.uint32) == L.TS_OK).all(), rep
        got = np.empty(n, dtype=np.int64)
        L.call("b2_read", dev, base, got.ctypes.data, 8 * n)
        assert (got == np.arange(n) + rep + 5).all(), rep


def test_instance_leaves_when_idle_and_a_new_one_starts(res_device):
    L, dev = res_device["L"], res_device["dev"]
    tasks, slot = _fanout(L, 64, 4096, 256 * MiB, 1 * GiB)
    t = C.c_uint64()
    s0 = _stats(L, dev)
    for _ in range(3):
        L.call("b2_enqueue", dev, tasks.ctypes.data, len(tasks), None, 0, None)
        L.call("b2_launch", dev, 0, C.byref(t))
        L.call("b2_wait", dev, t.value)
        time.sleep(0.05)                       # >> the 300 us idle timeout: the instance has left
    s1 = _stats(L, dev)
    assert s1.resident_batches - s0.resident_batches == 3
    assert s1.resident_instances - s0.resident_instances == 3
    # while no instance is resident, an ordinary kernel of the library runs at once
    spans = np.array([[256 * MiB, 2 * GiB, 1 * MiB]], dtype=np.uint64)
    t0 = time.perf_counter()
    L.call("b2_move_spans", dev, spans.ctypes.data, 1)
    L.call("b2_sync", dev)
    assert time.perf_counter() - t0 < 0.5


def test_span_mover_between_resident_batches(res_device):
    """pack -> resident batch -> unpack, all stream-ordered: the mover asks the resident instance to
    leave (it finishes what was submitted first); the batch after it starts a new instance."""
    L, dev = res_device["L"], res_device["dev"]
    n, elems = 500, 4096
    payload = elems * 4
    rng = np.random.default_rng(3)
    src = rng.standard_normal(n * elems, dtype=np.float32)
    scat0, packed, result, scat1 = 256 * MiB, 512 * MiB, 768 * MiB, 1 * GiB
    perm = rng.permutation(n).astype(np.uint64)
    L.call("b2_write", dev, scat0, src.ctypes.data, src.nbytes)
    offs = scat0 + perm * payload
    sizes = np.full(n, payload, dtype=np.uint64)
    L.call("b2_pack", dev, offs.ctypes.data, sizes.ctypes.data, n, packed, 256, None, None)
    tasks = np.zeros(n, dtype=L.TASK_DTYPE)
    tasks["fn_id"], tasks["count"] = 21, elems
    tasks["arg0"] = packed + np.arange(n, dtype=np.uint64) * payload
    tasks["out"] = result + np.arange(n, dtype=np.uint64) * payload
    t = C.c_uint64()
    L.call("b2_enqueue", dev, tasks.ctypes.data, n, None, 0, None)
    L.call("b2_launch", dev, 0, C.byref(t))
    dsts = scat1 + perm * payload
    L.call("b2_unpack", dev, result, dsts.ctypes.data, sizes.ctypes.data, n, 256, None, None)
    got = np.empty(n * elems, dtype=np.float32)
    L.call("b2_read", dev, scat1, got.ctypes.data, got.nbytes)
    assert np.array_equal(got, np.float32(2.0) * src + np.float32(1.0))


def test_submit_pipeline_in_resident_mode(res_device):
    """b2_submit with host operands: the payload rides the upload stream in front of the table and
    the doorbell; results and statuses come back through the egress stream."""
    L, dev = res_device["L"], res_device["dev"]
    h_out = _host_out(L, dev)
    n, elems = 300, 16384
    nbytes = elems * 4
    slot = (nbytes + 255) & ~255
    for rep in range(5):
        xs = [np.random.default_rng([rep, i]).standard_normal(elems, dtype=np.float32) for i in range(n)]
        jobs = np.zeros(n, dtype=L.JOB_DTYPE)
        tasks = np.zeros(n, dtype=L.TASK_DTYPE)
        for i, x in enumerate(xs):
            jobs[i] = (x.ctypes.data, i * slot, nbytes)
            tasks[i]["fn_id"], tasks[i]["arg0"], tasks[i]["out"], tasks[i]["count"] = 21, i * slot, GiB + i * slot, elems
        b = L.Batch()
        b.jobs, b.tasks, b.n_jobs, b.n_tasks = jobs.ctypes.data, tasks.ctypes.data, n, n
        b.flags = L.BATCH_INDEPENDENT
        b.in_off, b.in_bytes = 0, n * slot
        b.status_host_off = n * slot
        b.out_dev_off, b.out_host_off, b.out_bytes = GiB, 0, n * slot
        L.call("b2_submit", dev, C.byref(b))
        L.call("b2_wait", dev, b.ticket)
        assert (h_out[n * slot:n * slot + 4 * n].view(np.uint32) == L.TS_OK).all(), rep
        for i in (0, n // 3, n - 1):
            assert np.array_equal(h_out[i * slot:i * slot + nbytes].view(np.float32),
                                  np.float32(2.0) * xs[i] + np.float32(1.0)), (rep, i)


def test_disable_returns_to_one_launch_per_batch(res_device):
    L, dev = res_device["L"], res_device["dev"]
    L.call("b2_resident_disable", dev)
    tasks, slot = _fanout(L, 100, 1024, 256 * MiB, 1 * GiB)
    s0 = _stats(L, dev)
    t = C.c_uint64()
    L.call("b2_enqueue", dev, tasks.ctypes.data, len(tasks), None, 0, None)
    L.call("b2_launch", dev, 0, C.byref(t))
    L.call("b2_wait", dev, t.value)
    s1 = _stats(L, dev)
    assert s1.resident_batches == s0.resident_batches and s1.launches_worker == s0.launches_worker + 1
    L.call("b2_resident_enable", dev, 0)
