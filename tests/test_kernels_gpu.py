"""GPU parity tests for the kernels, through the C ABI, against the numpy model
in oracle/device_model.py (bit-exact for byte/integer work and for elementwise
fp32; 1e-6 relative for the tree-ordered sum)."""

import ctypes as C

import numpy as np
import pytest

from oracle import device_model as M

pytestmark = pytest.mark.gpu


def _write(L, dev, off, arr):
    a = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    L.call("b2_write", dev, off, a.ctypes.data, a.size)


def _read(L, dev, off, n):
    out = np.empty(n, dtype=np.uint8)
    L.call("b2_read", dev, off, out.ctypes.data, n)
    return out


def _spans(rows):
    arr = np.array(rows, dtype=np.uint64).reshape(-1, 3)
    return arr


def _move(L, dev, rows):
    arr = _spans(rows)
    L.call("b2_move_spans", dev, arr.ctypes.data, len(arr))
    L.call("b2_sync", dev)


EDGE_SIZES = [0, 1, 15, 16, 17, 31, 255, 256, 4095, 4096, 16383, 16384, 16385, 65536, 65537,
              (1 << 20) + 3]


def test_move_spans_ragged_edges(raw_device):
    L, dev = raw_device["L"], raw_device["dev"]
    rng = np.random.default_rng(0)
    region = 8 << 20
    src = rng.integers(0, 256, region, dtype=np.uint8)
    model = np.zeros(2 * region, dtype=np.uint8)
    model[:region] = src
    _write(L, dev, 0, model)
    rows, s, d = [], 0, region
    for n in EDGE_SIZES:
        rows.append((s, d, n))
        s += (n + 15) & ~15
        d += (n + 255) & ~255
    _move(L, dev, rows)
    M.move_spans(model, rows)
    got = _read(L, dev, 0, 2 * region)
    assert np.array_equal(got, model)


@pytest.mark.parametrize("src_mis,dst_mis", [(0, 0), (3, 3), (8, 8), (1, 5), (0, 7), (4, 0)])
def test_move_spans_misaligned(raw_device, src_mis, dst_mis):
    L, dev = raw_device["L"], raw_device["dev"]
    rng = np.random.default_rng(src_mis * 16 + dst_mis)
    region = 1 << 20
    model = np.zeros(2 * region, dtype=np.uint8)
    model[:region] = rng.integers(0, 256, region, dtype=np.uint8)
    _write(L, dev, 0, model)
    rows = []
    s, d = 64 + src_mis, region + 128 + dst_mis
    for n in [1, 7, 16, 33, 1000, 16384, 40000, 70001]:
        rows.append((s, d, n))
        s += ((n + 63) & ~63) + 0
        d += ((n + 63) & ~63) + 64
    _move(L, dev, rows)
    M.move_spans(model, rows)
    assert np.array_equal(_read(L, dev, 0, 2 * region), model)


def test_pack_then_unpack_round_trip_random(raw_device):
    L, dev = raw_device["L"], raw_device["dev"]
    rng = np.random.default_rng(7)
    n = 3000
    sizes = rng.integers(0, 40000, n).astype(np.uint64)
    sizes[rng.integers(0, n, 50)] = 0                       # empty payloads
    slot = 40960
    src_offs = (np.arange(n, dtype=np.uint64) * slot)[rng.permutation(n)]
    heap = n * slot
    data = rng.integers(0, 256, heap, dtype=np.uint8)
    model = np.zeros(3 * heap, dtype=np.uint8)
    model[:heap] = data
    _write(L, dev, 0, model)
    offs = np.zeros(n, dtype=np.uint64)
    total = C.c_uint64()
    L.call("b2_pack", dev, src_offs.ctypes.data, sizes.ctypes.data, n, heap, 16, offs.ctypes.data,
           C.byref(total))
    want_offs = M.pack(model, src_offs, sizes, heap, 16)
    assert list(offs) == want_offs
    assert total.value == M.pack_layout(sizes, heap, 16)[1]
    # scatter back to fresh slots in a third region
    dst_offs = (2 * heap + np.arange(n, dtype=np.uint64) * slot)[rng.permutation(n)]
    L.call("b2_unpack", dev, heap, dst_offs.ctypes.data, sizes.ctypes.data, n, 16, None, None)
    M.unpack(model, heap, dst_offs, sizes, 16)
    L.call("b2_sync", dev)
    got = _read(L, dev, 0, 3 * heap)
    assert np.array_equal(got, model)
    # property: every payload survived gather + scatter
    for i in range(0, n, 97):
        a = data[int(src_offs[i]):int(src_offs[i] + sizes[i])]
        b = got[int(dst_offs[i]):int(dst_offs[i] + sizes[i])]
        assert np.array_equal(a, b)


def test_pack_uniform_full_size_round_trip(raw_device):
    """BASELINE config 3 shape: 10 000 payloads of 64 KiB; checked by round trip."""
    L, dev = raw_device["L"], raw_device["dev"]
    n, sz = 10000, 65536
    rng = np.random.default_rng(11)
    data = rng.integers(0, 2 ** 32, n * sz // 4, dtype=np.uint32).view(np.uint8)
    _write(L, dev, 0, data)
    perm = rng.permutation(n).astype(np.uint64)
    src_offs = perm * sz
    sizes = np.full(n, sz, dtype=np.uint64)
    base = n * sz
    L.call("b2_pack", dev, src_offs.ctypes.data, sizes.ctypes.data, n, base, 16, None, None)
    L.call("b2_sync", dev)
    packed = _read(L, dev, base, n * sz)
    assert np.array_equal(packed.reshape(n, sz), data.reshape(n, sz)[perm.astype(np.int64)])
    L.call("b2_unpack", dev, base, src_offs.ctypes.data, sizes.ctypes.data, n, 16, None, None)
    L.call("b2_fill", dev, 0, 0, 16)  # touch: ordering is by stream
    L.call("b2_sync", dev)
    back = _read(L, dev, 16, n * sz - 16)
    assert np.array_equal(back, data[16:])


def test_move_spans_rejects_out_of_range(raw_device):
    L, dev = raw_device["L"], raw_device["dev"]
    arr = _spans([(0, raw_device["arena"] - 8, 64)])
    with pytest.raises(L.B2Error) as err:
        L.call("b2_move_spans", dev, arr.ctypes.data, 1)
    assert err.value.code == L.B2_E_ARG


# ------------------------------------------------------------------ worker ---
FN = {
    1: dict(kind=M.K_COPY), 2: dict(kind=M.K_SCALE_ADD, alpha=2.0, beta=1.0), 3: dict(kind=M.K_ADD),
    4: dict(kind=M.K_MUL), 5: dict(kind=M.K_AXPY, alpha=0.5), 6: dict(kind=M.K_SUM),
    7: dict(kind=M.K_STACK_SUM), 8: dict(kind=M.K_IADD), 9: dict(kind=M.K_IMUL),
    10: dict(kind=M.K_ISCALE_ADD, ialpha=2, ibeta=1), 11: dict(kind=M.K_SCALE_ADD, alpha=-1.5, beta=0.25),
}


@pytest.fixture(scope="module")
def registered(raw_device):
    L, dev = raw_device["L"], raw_device["dev"]
    for fid, e in FN.items():
        L.call("b2_register", dev, fid, e["kind"], e.get("alpha", 0.0), e.get("beta", 0.0),
               e.get("ialpha", 0), e.get("ibeta", 0))
    return raw_device


def _task(fn_id, arg0=0, arg1=0, out=0, count=0, nrows=0, deps=0, succ_begin=0, succ_count=0, flags=0):
    return dict(fn_id=fn_id, flags=flags, arg0=arg0 & (2 ** 64 - 1), arg1=arg1 & (2 ** 64 - 1), out=out,
                count=count, nrows=nrows, deps=deps, succ_begin=succ_begin, succ_count=succ_count, tag=0)


def _run(L, dev, tasks, succ=()):
    arr = np.zeros(len(tasks), dtype=L.TASK_DTYPE)
    for i, t in enumerate(tasks):
        for k, v in t.items():
            arr[i][k] = v
    s = np.array(list(succ), dtype=np.uint32)
    L.call("b2_enqueue", dev, arr.ctypes.data, len(arr), s.ctypes.data if len(s) else None, len(s), None)
    ticket = C.c_uint64()
    L.call("b2_launch", dev, 0, C.byref(ticket))
    L.call("b2_wait", dev, ticket.value)
    ptr, size = C.c_void_p(), C.c_uint64()
    L.call("b2_host_buffer", dev, 1, C.byref(ptr), C.byref(size))
    st = np.frombuffer((C.c_uint8 * (4 * len(tasks))).from_address(ptr.value), dtype=np.uint32).copy()
    return st


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 255, 1024, 1027, 16384, 262144 + 1])
def test_elementwise_callbacks_bit_exact(registered, n):
    L, dev = registered["L"], registered["dev"]
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n, dtype=np.float32)
    y = rng.standard_normal(n, dtype=np.float32)
    stride = (4 * n + 255) & ~255
    model = np.zeros(8 * stride + 256, dtype=np.uint8)
    model[0:4 * n] = x.view(np.uint8)
    model[stride:stride + 4 * n] = y.view(np.uint8)
    _write(L, dev, 0, model)
    tasks = [
        _task(1, 0, 0, 2 * stride, 4 * n),              # copy (count in bytes)
        _task(2, 0, 0, 3 * stride, n),
        _task(3, 0, stride, 4 * stride, n),
        _task(4, 0, stride, 5 * stride, n),
        _task(5, 0, stride, 6 * stride, n),
        _task(11, stride, 0, 7 * stride, n),
    ]
    st = _run(L, dev, tasks)
    want = M.run_tasks(model, FN, tasks, [])
    assert list(st) == want == [M.TS_OK] * len(tasks)
    got = _read(L, dev, 0, model.size)
    assert np.array_equal(got, model)


@pytest.mark.parametrize("n", [1, 7, 1000, 16384, 1 << 20])
def test_sum_callback_tolerance(registered, n):
    L, dev = registered["L"], registered["dev"]
    x = np.random.default_rng(n).random(n, dtype=np.float32)
    _write(L, dev, 0, x)
    out = (4 * n + 255) & ~255
    st = _run(L, dev, [_task(6, 0, 0, out, n)])
    assert list(st) == [M.TS_OK]
    got = _read(L, dev, out, 4).view(np.float32)[0]
    ref = x.sum()                                           # numpy pairwise fp32, what exec.py would produce
    assert abs(float(got) - float(ref)) <= 1e-6 * abs(float(ref))


@pytest.mark.parametrize("rows,n", [(1, 5), (3, 5), (17, 1000), (64, 4096), (5, 4099)])
def test_stack_sum_row_order_bit_exact(registered, rows, n):
    L, dev = registered["L"], registered["dev"]
    rng = np.random.default_rng(rows * 100003 + n)
    stride = (4 * n + 15) & ~15
    model = np.zeros(rows * stride + 4 * n + 512, dtype=np.uint8)
    xs = []
    for r in range(rows):
        v = rng.random(n, dtype=np.float32)
        xs.append(v)
        model[r * stride:r * stride + 4 * n] = v.view(np.uint8)
    _write(L, dev, 0, model)
    out = (rows * stride + 255) & ~255
    tasks = [_task(7, 0, stride, out, n, nrows=rows)]
    st = _run(L, dev, tasks)
    assert list(st) == [M.TS_OK]
    got = _read(L, dev, out, 4 * n).view(np.float32)
    assert np.array_equal(got, np.sum(np.stack(xs), axis=0))  # what the electron's Python body returns


def test_int_callbacks_and_overflow(registered):
    L, dev = registered["L"], registered["dev"]
    big = (1 << 63) - 1
    cases = [(8, 3, 4), (8, -5, 5), (8, big, 0), (8, big, 1), (8, -big - 1, -1),
             (9, 7, 6), (9, 1 << 31, (1 << 31) - 1), (9, 1 << 32, 1 << 32), (9, -big - 1, -1), (9, -3, 1 << 40),
             (10, 9999, 0), (10, -17, 0), (10, 1 << 62, 0), (10, (1 << 62) - 1, 0)]
    tasks = [_task(f, a, b, 4096 + 16 * i, 1, flags=3) for i, (f, a, b) in enumerate(cases)]
    st = _run(L, dev, tasks)
    raw = _read(L, dev, 4096, 16 * len(cases))
    for i, (f, a, b) in enumerate(cases):
        exact = {8: a + b, 9: a * b, 10: 2 * a + 1}[f]
        fits = -(1 << 63) <= exact <= big and (f != 10 or -(1 << 63) <= 2 * a <= big)
        assert st[i] == (M.TS_OK if fits else M.TS_OVERFLOW), (f, a, b)
        if fits:
            assert int(raw[16 * i:16 * i + 8].view(np.int64)[0]) == exact


def test_status_bad_fn_and_bad_arg(registered):
    L, dev = registered["L"], registered["dev"]
    tasks = [_task(500, 0, 0, 1024, 4),                       # nothing registered at 500
             _task(2, 8, 0, 1024, 4),                         # misaligned operand
             _task(2, registered["arena"] - 16, 0, 1024, 64),  # runs past the arena
             _task(2, 0, 0, 2048, 4)]
    st = _run(L, dev, tasks)
    assert list(st) == [M.TS_BAD_FN, M.TS_BAD_ARG, M.TS_BAD_ARG, M.TS_OK]


def test_many_independent_electrons(registered):
    """10 000 int electrons through the queue (BASELINE config 2)."""
    L, dev = registered["L"], registered["dev"]
    n = 10000
    tasks = [_task(8, i, 2 * i + 1, 1 << 20 | (16 * i), 1, flags=3) for i in range(n)]
    st = _run(L, dev, tasks)
    assert np.all(st == M.TS_OK)
    raw = _read(L, dev, 1 << 20, 16 * n).view(np.int64)[::2]
    assert np.array_equal(raw, 3 * np.arange(n, dtype=np.int64) + 1)


def test_dependency_chain_and_diamond(registered):
    """Successor release through the MPMC queue: a 200-long chain (each electron
    reads its predecessor's output) and a fan-out -> gate -> column-split reduce."""
    L, dev = registered["L"], registered["dev"]
    n = 4096
    x = np.random.default_rng(5).standard_normal(n, dtype=np.float32)
    slot = 4 * n
    chain = 200
    model = np.zeros((chain + 1) * slot + 64 * slot, dtype=np.uint8)
    model[:slot] = x.view(np.uint8)
    _write(L, dev, 0, model)
    tasks, succ = [], []
    for i in range(chain):
        tasks.append(_task(2, i * slot, 0, (i + 1) * slot, n, deps=0 if i == 0 else 1,
                           succ_begin=len(succ), succ_count=1 if i + 1 < chain else 0))
        if i + 1 < chain:
            succ.append(i + 1)
    st = _run(L, dev, tasks, succ)
    want = M.run_tasks(model, FN, tasks, succ)
    assert list(st) == want
    assert np.array_equal(_read(L, dev, 0, (chain + 1) * slot), model[:(chain + 1) * slot])

    # diamond: 48 fan-out electrons -> gate (fn 0) -> 4 column slices of a stack-sum
    fan, cols = 48, 4
    base = 0
    rows_off = (chain + 2) * slot
    tasks, succ = [], []
    gate = fan
    for i in range(fan):
        tasks.append(_task(11 if i % 2 else 2, base, 0, rows_off + i * slot, n, succ_begin=len(succ), succ_count=1))
        succ.append(gate)
    tasks.append(_task(0, deps=fan, succ_begin=len(succ), succ_count=cols))
    succ += [gate + 1 + c for c in range(cols)]
    out_off = rows_off + fan * slot
    w = n // cols
    for c in range(cols):
        tasks.append(_task(7, rows_off + 4 * c * w, slot, out_off + 4 * c * w, w, nrows=fan, deps=1))
    st = _run(L, dev, tasks, succ)
    want = M.run_tasks(model, FN, tasks, succ)
    assert list(st) == want == [M.TS_OK] * len(tasks)
    got = _read(L, dev, 0, model.size)
    assert np.array_equal(got, model)


@pytest.mark.parametrize("seed,n,width", [(1, 500, 64), (2, 3000, 256), (3, 3000, 4), (4, 20000, 16)])
def test_random_dags_match_the_model(registered, seed, n, width):
    """Stress of the MPMC release logic: random DAGs (each electron reads up to two
    earlier results), ready and dependent electrons interleaved in the table so that
    the ready-prefix shortcut, the cursor-publishing prologue and device-side
    successor release are all exercised under contention.  Every valid execution
    order gives the same bytes, so the arena must equal the model's."""
    L, dev = registered["L"], registered["dev"]
    rng = np.random.default_rng(seed)
    slot = (4 * width + 255) & ~255
    n_inputs = 8
    model = np.zeros((n_inputs + n) * slot + 256, dtype=np.uint8)
    for i in range(n_inputs):
        model[i * slot:i * slot + 4 * width] = rng.random(width, dtype=np.float32).view(np.uint8)
    _write(L, dev, 0, model)
    tasks, succ_lists = [], [[] for _ in range(n)]
    for i in range(n):
        out = (n_inputs + i) * slot
        n_pred = 0 if i < 3 else int(rng.integers(0, 3))
        preds = sorted(set(int(p) for p in rng.integers(max(0, i - 50), i, n_pred))) if n_pred else []
        srcs = [(n_inputs + p) * slot for p in preds] + [int(rng.integers(0, n_inputs)) * slot for _ in range(2 - len(preds))]
        if rng.random() < 0.5:
            t = _task(3 if rng.random() < 0.5 else 4, srcs[0], srcs[1], out, width, deps=len(preds))   # add / mul
        else:
            t = _task(11, srcs[0], 0, out, width, deps=len(preds))                                     # -1.5x + 0.25
            preds = preds[:1] if preds and srcs[0] == (n_inputs + preds[0]) * slot else []
            t["deps"] = len(preds)
        tasks.append(t)
        for p in preds:
            succ_lists[p].append(i)
    succ = []
    for i, t in enumerate(tasks):
        t["succ_begin"], t["succ_count"] = len(succ), len(succ_lists[i])
        succ += succ_lists[i]
    st = _run(L, dev, tasks, succ)
    want = M.run_tasks(model, FN, tasks, succ)
    assert list(st) == want and all(s == M.TS_OK for s in want)
    got = _read(L, dev, 0, model.size)
    assert np.array_equal(got, model)


def test_abi_housekeeping(registered):
    """Lifecycle / error-code behaviour of the C ABI that the engine relies on."""
    L, dev = registered["L"], registered["dev"]
    info = L.DeviceInfo()
    L.call("b2_get_info", dev, C.byref(info))
    assert info.cc_major == 10 and info.sm_count >= 100 and info.abi_version == L.ABI_VERSION
    assert info.arena_bytes == registered["arena"] and b"B200" in info.name
    with pytest.raises(L.B2Error) as e:                       # double init
        L.call("b2_init", dev, 1 << 20, 1 << 20, 1 << 20, 1024)
    assert e.value.code == L.B2_E_STATE
    with pytest.raises(L.B2Error) as e:                       # not a power of two
        L.call("b2_init", 9, 1 << 20, 0, 0, 1000)
    assert e.value.code == L.B2_E_ARG
    # more electrons than the queue can hold in one launch
    too_many = np.zeros((1 << 16) + 1, dtype=L.TASK_DTYPE)
    with pytest.raises(L.B2Error) as e:
        L.call("b2_enqueue", dev, too_many.ctypes.data, len(too_many), None, 0, None)
    assert e.value.code == L.B2_E_NOMEM
    # successor range outside the supplied list
    bad = np.zeros(1, dtype=L.TASK_DTYPE)
    bad["succ_begin"], bad["succ_count"] = 0, 3
    with pytest.raises(L.B2Error) as e:
        L.call("b2_enqueue", dev, bad.ctypes.data, 1, None, 0, None)
    assert e.value.code == L.B2_E_ARG
    assert L.load().b2_poll(dev, 10 ** 9) == L.B2_E_ARG                  # unknown ticket
    # an empty launch is a valid ticket; fill + events + stats
    t = C.c_uint64()
    L.call("b2_launch", dev, 0, C.byref(t))
    L.call("b2_wait", dev, t.value)
    assert L.load().b2_poll(dev, t.value) == 0
    s0, s1 = L.Stats(), L.Stats()
    L.call("b2_get_stats", dev, C.byref(s0))
    L.call("b2_fill", dev, 4096, 0xAB, 1000)
    e0, e1 = C.c_uint32(), C.c_uint32()
    L.call("b2_event_create", dev, C.byref(e0))
    L.call("b2_event_create", dev, C.byref(e1))
    L.call("b2_event_record", dev, e0.value)
    L.call("b2_flush_l2", dev, 1 << 20, 256 << 20)
    spans = _spans([(4096, 8192, 1000)])
    L.call("b2_move_spans", dev, spans.ctypes.data, 1)
    L.call("b2_event_record", dev, e1.value)
    ms = C.c_float()
    L.call("b2_event_elapsed_ms", dev, e0.value, e1.value, C.byref(ms))
    assert ms.value > 0
    assert np.all(_read(L, dev, 8192, 1000) == 0xAB)
    L.call("b2_get_stats", dev, C.byref(s1))
    assert s1.launches_pack == s0.launches_pack + 1 and s1.bytes_packed == s0.bytes_packed + 1000
