"""GPU tests of the round-2 ABI entry points, through the C ABI: the asynchronous batch
pipeline (b2_submit: host copy pool -> H2D -> worker -> D2H -> ticket), tombstones (b2_cancel /
B2_F_CANCELLED) and the source hash the library was built from."""

import asyncio
import ctypes as C

import numpy as np
import pytest

from oracle import device_model as M

pytestmark = pytest.mark.gpu


def test_library_was_built_from_this_source_tree(raw_device):
    """every number of round 1 came from a stale binary.  The library that runs on
    the GPU box must report the fingerprint of the csrc/ tree that travelled with it."""
    from covalent_ssh_plugin_b200 import build

    L = raw_device["L"]
    assert L.source_hash() == build.fingerprint()[:16], "libb2exec.so is not built from csrc/ as it is now"


def _batch(L, jobs, tasks, in_off, in_bytes, status_off, out_dev, out_host, out_bytes, independent=True):
    b = L.Batch()
    b.jobs = jobs.ctypes.data if len(jobs) else None
    b.tasks = tasks.ctypes.data
    b.n_jobs, b.n_tasks = len(jobs), len(tasks)
    b.flags = L.BATCH_INDEPENDENT if independent else 0
    b.in_off, b.in_bytes = in_off, in_bytes
    b.status_host_off = status_off
    b.out_dev_off, b.out_host_off, b.out_bytes = out_dev, out_host, out_bytes
    return b


def _host_views(L, dev):
    views = []
    for which in (0, 1):
        ptr, size = C.c_void_p(), C.c_uint64()
        L.call("b2_host_buffer", dev, which, C.byref(ptr), C.byref(size))
        views.append(np.frombuffer((C.c_uint8 * size.value).from_address(ptr.value), dtype=np.uint8))
    return views


@pytest.mark.parametrize("n,elems", [(1, 16), (257, 1000), (900, 16384), (5, (3 << 20) // 4 + 5)])
def test_submit_pipeline_matches_the_model(raw_device, n, elems):
    """Operands live in ordinary host arrays; one b2_submit moves them (copy pool, non-temporal
    stores, pieces of 256 KiB), uploads, runs scale_add on every one, downloads results and
    statuses.  Bit-exact against numpy, including sizes that are no multiple of the piece size."""
    L, dev = raw_device["L"], raw_device["dev"]
    h_in, h_out = _host_views(L, dev)
    L.call("b2_register", dev, 21, L.K_SCALE_ADD, 2.0, 1.0, 0, 0)
    rng = np.random.default_rng(n)
    xs = [rng.standard_normal(elems, dtype=np.float32) for _ in range(n)]
    nbytes = elems * 4
    slot = (nbytes + 255) & ~255
    if n * slot > min(h_in.size, h_out.size) - 65536:
        pytest.skip("staging buffers of the test fixture are too small for this case")
    out_dev0 = 1 << 30                                   # results land in the heap half of the arena
    jobs = np.zeros(n, dtype=L.JOB_DTYPE)
    tasks = np.zeros(n, dtype=L.TASK_DTYPE)
    for i, x in enumerate(xs):
        jobs[i] = (x.ctypes.data, i * slot, nbytes)
        tasks[i]["fn_id"], tasks[i]["arg0"], tasks[i]["out"], tasks[i]["count"] = 21, i * slot, out_dev0 + i * slot, elems
    status_off = n * slot
    h_out[:n * slot + 4 * n] = 0
    b = _batch(L, jobs, tasks, 0, n * slot, status_off, out_dev0, 0, n * slot)
    L.call("b2_submit", dev, C.byref(b))
    assert b.ticket > 0
    L.call("b2_wait", dev, b.ticket)
    assert L.load().b2_poll(dev, b.ticket) == 0
    st = h_out[status_off:status_off + 4 * n].view(np.uint32)
    assert (st == L.TS_OK).all()
    for i in (0, n // 2, n - 1):
        got = h_out[i * slot:i * slot + nbytes].view(np.float32)
        assert np.array_equal(got, np.float32(2.0) * xs[i] + np.float32(1.0)), i


def test_submissions_keep_program_order_with_other_calls(raw_device):
    """A b2_submit followed at once by stream-ordered calls (b2_read is one) must see the
    submission's effects: every ordered entry point first drains the submission thread."""
    L, dev = raw_device["L"], raw_device["dev"]
    L.call("b2_register", dev, 22, L.K_COPY, 0.0, 0.0, 0, 0)
    for round_ in range(20):
        x = np.full(4096, round_, dtype=np.float32)
        jobs = np.zeros(1, dtype=L.JOB_DTYPE)
        jobs[0] = (x.ctypes.data, 0, x.nbytes)
        tasks = np.zeros(1, dtype=L.TASK_DTYPE)
        tasks[0]["fn_id"], tasks[0]["arg0"], tasks[0]["out"], tasks[0]["count"] = 22, 0, 1 << 30, x.nbytes
        b = _batch(L, jobs, tasks, 0, x.nbytes, 1 << 20, 1 << 30, 0, 0)
        L.call("b2_submit", dev, C.byref(b))
        back = np.empty_like(x)
        L.call("b2_read", dev, 1 << 30, back.ctypes.data, back.nbytes)      # no explicit wait in between
        assert np.array_equal(back, x), round_
        L.call("b2_wait", dev, b.ticket)


def test_submit_reports_pipeline_failures_on_the_ticket(raw_device):
    L, dev = raw_device["L"], raw_device["dev"]
    x = np.zeros(64, dtype=np.float32)
    jobs = np.zeros(1, dtype=L.JOB_DTYPE)
    jobs[0] = (x.ctypes.data, (1 << 40), x.nbytes)             # far outside the ingress buffer
    tasks = np.zeros(1, dtype=L.TASK_DTYPE)
    tasks[0]["fn_id"], tasks[0]["out"], tasks[0]["count"] = 22, 1 << 30, 16
    b = _batch(L, jobs, tasks, 0, 256, 1 << 20, 1 << 30, 0, 0)
    L.call("b2_submit", dev, C.byref(b))
    with pytest.raises(L.B2Error, match="outside ingress"):
        L.call("b2_wait", dev, b.ticket)
    assert L.load().b2_poll(dev, b.ticket) < 0
    # the device is still usable afterwards
    t = C.c_uint64()
    L.call("b2_mark", dev, C.byref(t))
    L.call("b2_wait", dev, t.value)


def test_tombstones_in_the_table_skip_electrons_but_release_successors(raw_device):
    """B2_F_CANCELLED: the worker reports CANCELLED, does not touch the output, and still releases
    the electron's successors (a cancelled node must not wedge the launch)."""
    L, dev = raw_device["L"], raw_device["dev"]
    L.call("b2_register", dev, 23, L.K_IADD, 0.0, 0.0, 0, 0)
    n = 3000
    base = 1 << 30
    L.call("b2_fill", dev, base, 0xEE, 8 * n)
    tasks = np.zeros(n, dtype=L.TASK_DTYPE)
    tasks["fn_id"] = 23
    tasks["flags"] = L.F_IMM0 | L.F_IMM1
    tasks["arg0"] = np.arange(n)
    tasks["arg1"] = 7
    tasks["out"] = base + 8 * np.arange(n)
    tasks["count"] = 1
    dead = np.arange(n) % 5 == 2
    tasks["flags"][dead] |= L.F_CANCELLED
    # a chain at the end: task n-2 (cancelled) -> task n-1 must still run
    dead[n - 2] = True
    tasks["flags"][n - 2] |= L.F_CANCELLED
    tasks["succ_begin"][n - 2], tasks["succ_count"][n - 2] = 0, 1
    tasks["deps"][n - 1] = 1
    tasks["flags"][n - 1] &= ~np.uint32(L.F_CANCELLED)
    dead[n - 1] = False
    succ = np.array([n - 1], dtype=np.uint32)
    L.call("b2_enqueue", dev, tasks.ctypes.data, n, succ.ctypes.data, 1, None)
    t = C.c_uint64()
    L.call("b2_launch", dev, 0, C.byref(t))
    L.call("b2_wait", dev, t.value)
    _, h_out = _host_views(L, dev)
    st = h_out[:4 * n].view(np.uint32)
    assert (st[dead] == L.TS_CANCELLED).all() and (st[~dead] == L.TS_OK).all()
    got = np.empty(n, dtype=np.int64)
    L.call("b2_read", dev, base, got.ctypes.data, 8 * n)
    assert (got[~dead] == np.arange(n)[~dead] + 7).all()
    assert (got[dead].view(np.uint64) == 0xEEEEEEEEEEEEEEEE).all()         # untouched


def test_b2_cancel_before_the_worker_reaches_the_batch(raw_device):
    """b2_cancel on a ticket whose batch is still queued / tabled is exact: a long COPY batch keeps
    the device busy while a second submission waits behind it and gets tombstoned."""
    L, dev = raw_device["L"], raw_device["dev"]
    L.call("b2_register", dev, 22, L.K_COPY, 0.0, 0.0, 0, 0)
    L.call("b2_register", dev, 23, L.K_IADD, 0.0, 0.0, 0, 0)
    # 1. something long: 2000 x 1 MiB device-to-device copies (~1 ms of HBM time), resident operands
    big = np.zeros(2000, dtype=L.TASK_DTYPE)
    big["fn_id"], big["count"] = 22, 1 << 20
    big["arg0"] = (1 << 30) + (np.arange(2000) % 256) * (1 << 20)
    big["out"] = (1 << 30) + (256 << 20) + (np.arange(2000) % 256) * (1 << 20)
    b1 = _batch(L, np.zeros(0, dtype=L.JOB_DTYPE), big, 0, 0, 1 << 20, 0, 0, 0)
    # 2. the batch to cancel from
    n = 64
    tasks = np.zeros(n, dtype=L.TASK_DTYPE)
    tasks["fn_id"], tasks["flags"], tasks["count"] = 23, L.F_IMM0 | L.F_IMM1, 1
    tasks["arg0"], tasks["arg1"] = np.arange(n), 1
    tasks["out"] = (1 << 30) + (600 << 20) + 8 * np.arange(n)
    status_off = 2 << 20
    b2 = _batch(L, np.zeros(0, dtype=L.JOB_DTYPE), tasks, 0, 0, status_off, 0, 0, 0)
    L.call("b2_submit", dev, C.byref(b1))
    L.call("b2_submit", dev, C.byref(b2))
    L.call("b2_cancel", dev, b2.ticket, 10, 20)            # electrons 10..29
    L.call("b2_wait", dev, b2.ticket)
    _, h_out = _host_views(L, dev)
    st = h_out[status_off:status_off + 4 * n].view(np.uint32)
    cancelled = st == L.TS_CANCELLED
    # exact when the tombstone beat the worker; never anything but OK / CANCELLED, and only in range
    assert not cancelled[:10].any() and not cancelled[30:].any()
    assert ((st == L.TS_OK) | cancelled).all()
    assert cancelled[10:30].all(), "b2_cancel issued right after b2_submit, behind ~1 ms of work, must be exact"


def _model_scale_add(x):
    return np.float32(2.0) * x + np.float32(1.0)


@pytest.mark.parametrize("elems,n", [(64, 60000), (256, 40000), (1024, 30000), (3000, 4000), (16384, 1500), (70000, 300)])
def test_both_worker_flavours_are_bit_exact(raw_device, elems, n):
    """The host picks the streaming worker's flavour from the average operand size (12 x 4 KiB stages up to
    8 KiB, with the cooperative producer warp and one worker warp per electron; 3 x 16 KiB stages above).
    Same electrons, same bytes out, whichever runs — including ragged sizes that straddle tile boundaries,
    and launches big enough for the widest claims (8 tickets)."""
    L, dev = raw_device["L"], raw_device["dev"]
    L.call("b2_register", dev, 21, L.K_SCALE_ADD, 2.0, 1.0, 0, 0)
    rng = np.random.default_rng(elems)
    sizes = np.maximum(1, elems + rng.integers(-elems // 8, elems // 8 + 1, n)).astype(np.int64)
    slots = ((sizes * 4 + 255) & ~255).astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(slots)[:-1]])
    total = int(slots.sum())
    assert 2 * total + (64 << 20) < raw_device["arena"]
    src = rng.standard_normal(total // 4, dtype=np.float32)
    base_in, base_out = 128 << 20, (128 << 20) + total
    L.call("b2_write", dev, base_in, src.ctypes.data, total)
    L.call("b2_fill", dev, base_out, 0, total)
    tasks = np.zeros(n, dtype=L.TASK_DTYPE)
    tasks["fn_id"], tasks["count"] = 21, sizes
    tasks["arg0"], tasks["out"] = base_in + offs, base_out + offs
    L.call("b2_enqueue", dev, tasks.ctypes.data, n, None, 0, None)
    t = C.c_uint64()
    L.call("b2_launch", dev, 0, C.byref(t))
    L.call("b2_wait", dev, t.value)
    _, h_out = _host_views(L, dev)
    assert (h_out[:4 * n].view(np.uint32) == L.TS_OK).all()
    got = np.empty(total // 4, dtype=np.float32)
    L.call("b2_read", dev, base_out, got.ctypes.data, total)
    for i in list(range(0, n, max(1, n // 64))) + [n - 1]:
        o, k = int(offs[i]) // 4, int(sizes[i])
        assert np.array_equal(got[o:o + k], _model_scale_add(src[o:o + k])), (i, k)
        assert not got[o + k:o + int(slots[i]) // 4].any(), "bytes beyond the electron's result were written"


def test_resident_table_relaunches_with_dependencies(raw_device):
    """b2_table_create uploads a table once; every b2_table_launch re-arms the dependency counters
    on the device.  A diamond (root -> 64 fan-out -> reduce behind a gate) launched five times over
    changing inputs gives the model's answer every time."""
    L, dev = raw_device["L"], raw_device["dev"]
    L.call("b2_register", dev, 21, L.K_SCALE_ADD, 2.0, 1.0, 0, 0)
    L.call("b2_register", dev, 24, L.K_STACK_SUM, 0.0, 0.0, 0, 0)
    fan, elems = 64, 4096
    row = elems * 4
    x_off, mid_off, out_off = 256 << 20, (256 << 20) + (1 << 20), (256 << 20) + (64 << 20)
    tasks = np.zeros(fan + 2, dtype=L.TASK_DTYPE)
    # 0: root  y = 2x+1 ; 1..fan: z_i = 2y+1 into consecutive rows ; fan+1: stack_sum of the rows
    tasks[0]["fn_id"], tasks[0]["arg0"], tasks[0]["out"], tasks[0]["count"] = 21, x_off, x_off + row, elems
    tasks[0]["succ_begin"], tasks[0]["succ_count"] = 0, fan
    for i in range(1, fan + 1):
        tasks[i]["fn_id"], tasks[i]["arg0"], tasks[i]["out"], tasks[i]["count"] = 21, x_off + row, mid_off + (i - 1) * row, elems
        tasks[i]["deps"], tasks[i]["succ_begin"], tasks[i]["succ_count"] = 1, fan, 1
    r = fan + 1
    tasks[r]["fn_id"], tasks[r]["arg0"], tasks[r]["arg1"], tasks[r]["out"] = 24, mid_off, row, out_off
    tasks[r]["count"], tasks[r]["nrows"], tasks[r]["deps"] = elems, fan, fan
    succ = np.array(list(range(1, fan + 1)) + [r], dtype=np.uint32)
    tid = C.c_uint32()
    L.call("b2_table_create", dev, tasks.ctypes.data, len(tasks), succ.ctypes.data, len(succ), C.byref(tid))
    t = C.c_uint64()
    for rep in range(5):
        x = np.random.default_rng(rep).random(elems, dtype=np.float32)
        L.call("b2_write", dev, x_off, x.ctypes.data, row)
        L.call("b2_table_launch", dev, tid.value, 0, C.byref(t))
        L.call("b2_wait", dev, t.value)
        _, h_out = _host_views(L, dev)
        assert (h_out[:4 * (fan + 2)].view(np.uint32) == L.TS_OK).all(), rep
        got = np.empty(elems, dtype=np.float32)
        L.call("b2_read", dev, out_off, got.ctypes.data, row)
        z = _model_scale_add(_model_scale_add(x))
        want = z.copy()
        for _ in range(fan - 1):
            want = want + z                                  # row order, fp32 — as the callback adds
        assert np.array_equal(got, want), rep
    L.call("b2_table_destroy", dev, tid.value)
    with pytest.raises(L.B2Error):
        L.call("b2_table_launch", dev, tid.value, 0, C.byref(t))
