"""Round-2 host-side regressions (CPU only, numpy model of the device behind the C ABI):

* the build fingerprint covers every csrc/ file and the library reports the hash it was built from;
* stack_sum staging budget, cancellation isolation inside a batch, batching under
  back-pressure, an unloadable library is never retried as "host unreachable";
* cancel(): tombstones before the batch ships, raises TaskCancelledError, leaves batch-mates alone;
* the fused body of run() is taken only while no helper is patched, and behaves like the
  stepwise body (same results, same failure decisions).
"""

from __future__ import annotations

import asyncio
import shutil
from pathlib import Path
from unittest.mock import patch

import numpy as np
import pytest

from covalent_ssh_plugin_b200 import SSHExecutor, _b2lib, build, electrons
from covalent_ssh_plugin_b200 import engine as engine_mod
from covalent_ssh_plugin_b200 import ssh as impl
from tests.fake_b2lib import FakeLib

SMALL = dict(arena_bytes=48 << 20, staging_bytes=16 << 20, n_segments=4, max_batch=16, queue_cap=64)


@pytest.fixture()
def fake_lib():
    impl._inflight.clear()
    impl._live.clear()
    engine_mod.shutdown_all()
    saved = _b2lib._lib
    lib = FakeLib(n_devices=2)
    _b2lib._lib = lib
    yield lib
    engine_mod.shutdown_all()
    _b2lib._lib = saved


def make_executor(tmp_path, **kw):
    key = tmp_path / "id_stub"
    key.touch()
    return SSHExecutor(username="u", hostname=kw.pop("hostname", "cuda:0"), ssh_key_file=str(key),
                       cache_dir=str(tmp_path), **kw)


# ------------------------------------------------------------------ build ------
def test_fingerprint_covers_every_source_file(tmp_path):
    """the shipped .so was stale because two headers were missing from a hand-kept
    list.  Any byte of any file under csrc/ (and the public header) must change the fingerprint."""
    csrc = tmp_path / "csrc"
    shutil.copytree(build.CSRC, csrc)
    base = build.fingerprint(csrc)
    names = sorted(p.name for p in build.source_files(csrc))
    assert {"b2exec.cu", "b2_worker_tma.cuh", "b2_spans_tma.cuh", "b2_worker.cuh", "b2_spans.cuh",
            "b2_device.cuh", "b2_hostcopy.cpp", "b2_hostcopy.h"} <= set(names)
    for name in names:
        path = csrc / name
        original = path.read_bytes()
        path.write_bytes(original + b"\n// touched\n")
        assert build.fingerprint(csrc) != base, f"{name} is not part of the fingerprint"
        path.write_bytes(original)
    assert build.fingerprint(csrc) == base
    (csrc / "brand_new.cuh").write_text("// a header nobody listed\n")
    assert build.fingerprint(csrc) != base


def test_is_fresh_goes_false_when_a_source_byte_changes(tmp_path, built_library):
    csrc = tmp_path / "csrc"
    shutil.copytree(build.CSRC, csrc)
    assert build.is_fresh()                                   # conftest built the real tree
    assert build.is_fresh(csrc=csrc)                          # identical copy
    with open(csrc / "b2_worker_tma.cuh", "ab") as fh:
        fh.write(b"\n")
    assert not build.is_fresh(csrc=csrc)


def test_library_reports_the_source_hash_it_was_built_from(built_library):
    saved, _b2lib._lib = _b2lib._lib, None
    try:
        assert _b2lib.source_hash() == build.fingerprint()[:16] == build.built_source_hash()
    finally:
        _b2lib._lib = saved


# ------------------------------------------------------------------ ADVICE -----
def test_stack_sum_budget_matches_its_layout(fake_lib, tmp_path):
    """[K, n] stack_sum rows are re-strided to 16 B; the budget must use
    the same formula or a batch writes past its pinned segment."""
    eng = engine_mod.get_engine(0, arena_bytes=24 << 20, staging_bytes=4 << 20, n_segments=4, max_batch=1024,
                                queue_cap=1024)
    ex = make_executor(tmp_path)
    stacks = [np.random.default_rng(i).random((300, 3), dtype=np.float32) for i in range(700)]

    async def go():
        return await asyncio.gather(*[
            ex.run(electrons.stack_sum, [stacks[i]], {}, {"dispatch_id": "ss", "node_id": i}) for i in range(700)])

    out = asyncio.run(go())
    for i in (0, 1, 350, 699):
        assert np.array_equal(out[i], np.sum(np.stack(list(stacks[i])), axis=0))
    enc = electrons.encode(electrons.stack_sum, [stacks[0]], {})
    assert eng._need_in(enc) == 300 * 16 + (-(300 * 16) % 256)
    assert sorted(eng._free_segments) == [0, 1, 2, 3]


def test_cancelling_one_run_leaves_its_batch_mates_alone(fake_lib, tmp_path):
    """awaiting one shared future meant that a cancelled or timed-out run()
    cancelled the whole batch."""
    engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    xs = [np.full(64, i, dtype=np.float32) for i in range(10)]

    async def go():
        tasks = [asyncio.ensure_future(ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": "c", "node_id": i}))
                 for i in range(10)]
        await asyncio.sleep(0)                    # every run() has staged and is waiting for the batch
        tasks[4].cancel()
        return await asyncio.gather(*tasks, return_exceptions=True)

    out = asyncio.run(go())
    assert isinstance(out[4], asyncio.CancelledError)
    for i in range(10):
        if i != 4:
            assert np.array_equal(out[i], np.float32(2.0) * xs[i] + np.float32(1.0)), i
    eng = engine_mod.existing_engine(0)
    assert sorted(eng._free_segments) == [0, 1, 2, 3] and not impl._live


def test_stepwise_body_also_isolates_cancellation(fake_lib, tmp_path):
    engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)

    class Patched(SSHExecutor):           # an overridden helper switches run() to the stepwise body
        async def get_status(self, conn, remote_result_file):
            return await SSHExecutor.get_status(self, conn, remote_result_file)

    key = tmp_path / "id_stub"
    ex = Patched(username="u", hostname="cuda:0", ssh_key_file=str(key), cache_dir=str(tmp_path))
    assert not ex._helpers_untouched()

    async def go():
        tasks = [asyncio.ensure_future(ex.run(electrons.add_int, [i, 1], {}, {"dispatch_id": "c2", "node_id": i}))
                 for i in range(6)]
        for _ in range(4):
            await asyncio.sleep(0)
        tasks[2].cancel()
        return await asyncio.gather(*tasks, return_exceptions=True)

    out = asyncio.run(go())
    assert isinstance(out[2], asyncio.CancelledError)
    assert [o for i, o in enumerate(out) if i != 2] == [1, 2, 4, 5, 6]
    assert not impl._inflight and not impl._live


def test_back_pressure_keeps_batches_full(fake_lib, tmp_path):
    """one waiter woken per released segment shipped ~1-electron batches."""
    fake_lib.delay_polls = 2
    eng = engine_mod.get_engine(0, arena_bytes=64 << 20, staging_bytes=16 << 20, n_segments=4, max_batch=4096,
                                queue_cap=4096)
    ex = make_executor(tmp_path)
    xs = [np.random.default_rng(i).standard_normal(16384, dtype=np.float32) for i in range(600)]

    async def go():
        return await asyncio.gather(*[
            ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": "bp", "node_id": i}) for i in range(600)])

    out = asyncio.run(go())
    for i in range(0, 600, 37):
        assert np.array_equal(out[i], np.float32(2.0) * xs[i] + np.float32(1.0))
    per_segment = (4 << 20) // ((64 << 10) + 256)                # ~63 electrons fit one 4 MiB segment
    ideal = -(-600 // per_segment)
    assert eng.counters["batches"] <= 3 * ideal, (eng.counters["batches"], ideal)
    assert sorted(eng._free_segments) == [0, 1, 2, 3]


def test_unloadable_library_is_library_missing_not_oserror(tmp_path):
    """ctypes.CDLL raising OSError was retried five times as "could not connect"
    and, with run_local_on_ssh_fail, silently ran on the CPU."""
    bogus = tmp_path / "libb2exec.so"
    bogus.write_bytes(b"not an ELF file")
    saved = _b2lib._lib
    _b2lib._lib = None
    try:
        with patch.dict("os.environ", {"B2EXEC_LIB": str(bogus)}):
            with pytest.raises(_b2lib.LibraryMissing):
                _b2lib.load()
    finally:
        _b2lib._lib = saved


def test_key_file_cache_is_per_check(tmp_path):
    """a directory passes os.path.exists but must not pass the isfile check."""
    import os

    d = tmp_path / "a_directory"
    d.mkdir()
    impl._key_seen.clear()
    impl._key_is_file.clear()
    assert impl._key_file_present(str(d), os.path.exists)
    assert not impl._key_file_present(str(d), os.path.isfile)
    assert str(d) not in impl._key_is_file


def test_single_element_stack_rows_are_left_to_the_host():
    with pytest.raises(electrons.NotDeviceEligible):
        electrons.encode(electrons.stack_sum, [np.ones((300, 1), dtype=np.float32)], {})
    electrons.encode(electrons.stack_sum, [np.ones((4, 1), dtype=np.float32)], {})


# ------------------------------------------------------------------ cancel -----
def test_cancel_one_electron_before_its_batch_ships(fake_lib, tmp_path):
    engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    xs = [np.full(32, i, dtype=np.float32) for i in range(8)]

    async def go():
        tasks = [asyncio.ensure_future(ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": "k", "node_id": i}))
                 for i in range(8)]
        await asyncio.sleep(0)                    # staged, batch still open (the idle flush has not run)
        eng = engine_mod.existing_engine(0)
        assert eng._open is not None and not eng._open.flushed
        table = eng._open.table
        assert await ex.cancel({"dispatch_id": "k", "node_id": 5}) is True
        assert int(table["flags"][5]) & _b2lib.F_CANCELLED and not int(table["flags"][4]) & _b2lib.F_CANCELLED
        assert await ex.cancel({"dispatch_id": "k", "node_id": 5}) is False          # already gone
        return await asyncio.gather(*tasks, return_exceptions=True)

    out = asyncio.run(go())
    assert isinstance(out[5], impl.TaskCancelledError)
    for i in range(8):
        if i != 5:
            assert np.array_equal(out[i], np.float32(2.0) * xs[i] + np.float32(1.0))
    eng = engine_mod.existing_engine(0)
    del out                                        # results are views of the pinned egress segment ...
    import gc

    gc.collect()                                   # ... and the stored exception's traceback keeps a cycle alive
    assert sorted(eng._free_segments) == [0, 1, 2, 3] and sorted(eng._free_out) == [0, 1, 2, 3]
    assert not impl._live


def test_cancel_everything_of_one_executor_only(fake_lib, tmp_path):
    engine_mod.get_engine(0, **SMALL)
    engine_mod.get_engine(1, **SMALL)
    a, b = make_executor(tmp_path, hostname="cuda:0"), make_executor(tmp_path, hostname="cuda:1")

    async def go():
        ta = [asyncio.ensure_future(a.run(electrons.add_int, [i, 1], {}, {"dispatch_id": "a", "node_id": i}))
              for i in range(5)]
        tb = [asyncio.ensure_future(b.run(electrons.add_int, [i, 2], {}, {"dispatch_id": "b", "node_id": i}))
              for i in range(5)]
        await asyncio.sleep(0)
        assert await a.cancel() is True
        return await asyncio.gather(*ta, *tb, return_exceptions=True)

    out = asyncio.run(go())
    assert all(isinstance(o, impl.TaskCancelledError) for o in out[:5])
    assert out[5:] == [2, 3, 4, 5, 6]


def test_cancel_after_submission_goes_through_b2_cancel(fake_lib, tmp_path):
    import threading

    gate = threading.Event()
    real_wait = fake_lib.b2_wait
    fake_lib.b2_wait = lambda dev, ticket: (gate.wait(10), real_wait(dev, ticket))[1]   # the "GPU" is busy
    fake_lib.delay_polls = 1 << 30
    engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)

    async def go():
        t = asyncio.ensure_future(ex.run(electrons.add_int, [1, 2], {}, {"dispatch_id": "late", "node_id": 0}))
        for _ in range(6):
            await asyncio.sleep(0)                # the idle flush has submitted the batch
        eng = engine_mod.existing_engine(0)
        assert eng._open is None and eng._inflight and eng._inflight[0].flushed
        assert await ex.cancel({"dispatch_id": "late", "node_id": 0}) is True
        out = await asyncio.gather(t, return_exceptions=True)
        gate.set()
        return out

    out = asyncio.run(go())
    assert isinstance(out[0], impl.TaskCancelledError)
    assert fake_lib.cancels and fake_lib.cancels[0][2:] == (0, 1)
    gate.set()


def test_device_model_reports_tombstones(fake_lib):
    from oracle import device_model as M

    arena = np.zeros(4096, dtype=np.uint8)
    tasks = [dict(fn_id=1, flags=M.F_IMM0 | M.F_IMM1 | (M.F_CANCELLED if i == 1 else 0), arg0=i, arg1=1, out=256 * i,
                  count=1, nrows=0, deps=0, succ_begin=0, succ_count=0, tag=0) for i in range(3)]
    st = M.run_tasks(arena, {1: dict(kind=M.K_IADD)}, tasks, [])
    assert st == [M.TS_OK, M.TS_CANCELLED, M.TS_OK]


# ------------------------------------------------------------- fused run() -----
def test_fused_body_is_taken_only_while_nothing_is_patched(fake_lib, tmp_path):
    engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    assert ex._helpers_untouched()
    asyncio.run(ex.run(electrons.add_int, [1, 2], {}, {"dispatch_id": "f", "node_id": 0}))   # opens the connection
    seen = []
    real = SSHExecutor._write_function_files

    def spy(self, *a, **k):
        seen.append(a[0])
        return real(self, *a, **k)

    assert asyncio.run(ex.run(electrons.add_int, [2, 3], {}, {"dispatch_id": "f", "node_id": 1})) == 5
    with patch.object(SSHExecutor, "_write_function_files", spy):
        assert not ex._helpers_untouched()
        assert asyncio.run(ex.run(electrons.add_int, [3, 4], {}, {"dispatch_id": "f", "node_id": 2})) == 7
    assert seen == ["f_2"]                         # only the patched run went through the helpers
    with patch.object(ex, "cleanup", side_effect=SSHExecutor.cleanup.__get__(ex)):
        assert not ex._helpers_untouched()         # instance-level patches count too
    assert ex._helpers_untouched()


def test_fused_and_stepwise_bodies_agree(fake_lib, tmp_path):
    """Same electrons through both bodies: identical results, types, shapes, orders, and the
    same failure decision for an overflowing int electron."""
    engine_mod.get_engine(0, **SMALL)
    fused = make_executor(tmp_path)

    class Stepwise(SSHExecutor):
        async def _poll_task(self, conn, remote_result_file, retries=5):
            return await SSHExecutor._poll_task(self, conn, remote_result_file, retries)

    step = Stepwise(username="u", hostname="cuda:0", ssh_key_file=str(tmp_path / "id_stub"), cache_dir=str(tmp_path))
    rng = np.random.default_rng(7)
    m = rng.standard_normal((5, 7), dtype=np.float32)
    calls = [
        (electrons.add_int, [3, 4]), (electrons.mul_int, [-5, 6]), (electrons.affine_int, [10]),
        (electrons.scale_add, [rng.standard_normal(1000, dtype=np.float32)]),
        (electrons.identity, [m]), (electrons.plus_one, [np.asfortranarray(m)]), (electrons.scale_add, [m[:, ::2]]),
        (electrons.vec_add, [m, m * 2]), (electrons.vec_mul, [m[0], m[1]]), (electrons.axpy_half, [m, m]),
        (electrons.vec_sum, [rng.random(333, dtype=np.float32)]),
        (electrons.scale_add, [np.zeros(0, dtype=np.float32)]),
        (electrons.stack_sum, [[m, m, m]]),
    ]

    async def through(ex, tag):
        return await asyncio.gather(*[ex.run(fn, args, {}, {"dispatch_id": tag, "node_id": i})
                                      for i, (fn, args) in enumerate(calls)])

    a, b = asyncio.run(through(fused, "fu")), asyncio.run(through(step, "st"))
    for (fn, args), x, y in zip(calls, a, b):
        want = fn(*args)
        assert type(x) is type(y) is type(want), fn
        if isinstance(want, np.ndarray):
            assert x.shape == y.shape == want.shape and x.dtype == y.dtype == want.dtype
            assert x.flags.c_contiguous == y.flags.c_contiguous
            assert np.array_equal(x, y) and np.array_equal(x, want)
        elif isinstance(want, np.floating):
            assert abs(float(x) - float(want)) <= 1e-6 * abs(float(want)) and x == y
        else:
            assert x == y == want
    big = (1 << 63) - 1
    for ex in (fused, step):
        with pytest.raises(RuntimeError, match="does not fit int64"):
            asyncio.run(ex.run(electrons.add_int, [big, 1], {}, {"dispatch_id": "ov", "node_id": 0}))
    local = make_executor(tmp_path, run_local_on_ssh_fail=True)
    assert asyncio.run(local.run(electrons.add_int, [big, 1], {}, {"dispatch_id": "ov", "node_id": 1})) == big + 1


def test_missing_key_file_raises_the_same_way_in_both_bodies(fake_lib, tmp_path):
    engine_mod.get_engine(0, **SMALL)
    ex = SSHExecutor(username="u", hostname="cuda:0", ssh_key_file=str(tmp_path / "nope"), cache_dir=str(tmp_path))
    impl._key_seen.clear()
    impl._key_is_file.clear()
    with pytest.raises(RuntimeError, match="does not exist"):
        asyncio.run(ex.run(electrons.add_int, [1, 2], {}, {"dispatch_id": "nk", "node_id": 0}))


# ------------------------------------------------------------- host pool -------
def _shape_and_cwd(x):
    import os

    return x * 2, os.getcwd()


def _boom(msg):
    raise KeyError(msg)


def _exit_hard():
    import os

    os._exit(3)


def test_host_pool_honours_python_path_and_conda_env(fake_lib, tmp_path):
    """SURVEY.md §8f-2 / ssh.py:377-380, 508-524: with the pool switched on, an executor that names
    its own interpreter gets warm workers started as `<python_path> hostworker.py` (behind
    `conda activate <env>` when asked), verified the way the reference verifies the remote side.
    Large arrays cross /dev/shm as protocol-5 out-of-band buffers, both ways."""
    import os
    import sys

    import cloudpickle

    import tests.test_round2_host as me
    from covalent_ssh_plugin_b200 import hostpool

    cloudpickle.register_pickle_by_value(me)
    engine_mod.get_engine(0, **SMALL)
    wd = tmp_path / "wd"
    with patch.dict(os.environ, {"B2_HOST_POOL": "1", "B2_HOST_POOL_WORKERS": "2"}):
        try:
            ex = make_executor(tmp_path, python_path=sys.executable, remote_workdir=str(wd))
            big = np.random.default_rng(0).standard_normal(1 << 18, dtype=np.float32)        # 1 MiB: out of band
            async def go():
                return await asyncio.gather(
                    ex.run(_shape_and_cwd, [big], {}, {"dispatch_id": "pp", "node_id": 0}),
                    ex.run(_shape_and_cwd, [np.arange(3.0)], {}, {"dispatch_id": "pp", "node_id": 1}),
                    ex.run(_boom, ["k"], {}, {"dispatch_id": "pp", "node_id": 2}), return_exceptions=True)

            (r0, cwd0), (r1, _), err = asyncio.run(go())
            assert np.array_equal(r0, big * 2) and cwd0 == str(wd)
            assert np.array_equal(r1, np.arange(3.0) * 2)
            assert isinstance(err, KeyError) and err.args == ("k",)
            assert not [p for p in os.listdir("/dev/shm") if p.startswith("b2pool-")], "shared-memory files leaked"
            with pytest.raises(RuntimeError, match="worker process died"):
                asyncio.run(ex.run(_exit_hard, [], {}, {"dispatch_id": "pp", "node_id": 3}))
            # an interpreter that is not Python 3 / does not exist: the reference's message, the failure decision
            bad = make_executor(tmp_path, python_path=str(tmp_path / "no-such-python"))
            with pytest.raises(RuntimeError, match="No Python 3 installation found"):
                asyncio.run(bad.run(_boom, ["x"], {}, {"dispatch_id": "pp", "node_id": 4}))
            # a conda environment that does not exist (this image has no conda at all)
            env = make_executor(tmp_path, python_path=sys.executable, conda_env="no_such_env")
            with pytest.raises(RuntimeError) as ei:
                asyncio.run(env.run(_boom, ["x"], {}, {"dispatch_id": "pp", "node_id": 5}))
            assert "conda" in str(ei.value).lower()
            local = make_executor(tmp_path, python_path=sys.executable, conda_env="no_such_env", run_local_on_ssh_fail=True)
            with pytest.raises(KeyError):
                asyncio.run(local.run(_boom, ["x"], {}, {"dispatch_id": "pp", "node_id": 6}))       # ran in process
            # device electrons never go to the pool, whatever the interpreter settings
            assert asyncio.run(ex.run(electrons.add_int, [2, 3], {}, {"dispatch_id": "pp", "node_id": 7})) == 5
        finally:
            hostpool.shutdown()


def test_a_loop_closed_under_pending_runs_does_not_take_the_engine_down(fake_lib, tmp_path):
    """The completion thread serves every later batch: waking the waiters of a loop that was closed
    with runs still pending must not end it."""
    import threading
    import time

    engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    gate = threading.Event()
    real_wait = fake_lib.b2_wait

    def slow_wait(dev, ticket):
        gate.wait(5)
        return real_wait(dev, ticket)

    fake_lib.b2_wait = slow_wait
    x = np.arange(64, dtype=np.float32)
    loop = asyncio.new_event_loop()
    tasks = [loop.create_task(ex.run(electrons.scale_add, [x], {}, {"dispatch_id": "a", "node_id": i}))
             for i in range(3)]
    loop.run_until_complete(asyncio.sleep(0.05))       # staged and shipped; the completion thread sits in b2_wait
    assert not any(t.done() for t in tasks)
    loop.close()                                       # abnormal: pending runs abandoned with their loop
    gate.set()
    time.sleep(0.2)
    fake_lib.b2_wait = real_wait
    assert engine_mod.get_engine(0)._waiter.is_alive()
    out = asyncio.run(ex.run(electrons.scale_add, [x], {}, {"dispatch_id": "b", "node_id": 0}))
    assert np.array_equal(out, np.float32(2) * x + np.float32(1))
    del tasks


def test_back_pressure_keeps_the_fused_path(fake_lib, tmp_path, monkeypatch):
    """With every staging segment in flight a fused run() parks and retries the one-call staging; it
    does not drop to the step-by-step body (eight awaited helpers per electron), and a release wakes
    a batch worth of parked electrons, not one and not all."""
    import time

    fake_lib.delay_polls = 10 ** 9                             # batches finish through b2_wait only ...
    real_wait = fake_lib.b2_wait

    def slow_wait(dev, ticket):                                # ... and each takes a while
        time.sleep(0.004)
        return real_wait(dev, ticket)

    fake_lib.b2_wait = slow_wait
    eng = engine_mod.get_engine(0, **SMALL)                    # 4 x 4 MiB segments, 16 electrons per batch
    ex = make_executor(tmp_path)
    general = []
    real_stage = eng.stage

    async def counted_stage(enc):
        general.append(enc)
        return await real_stage(enc)

    monkeypatch.setattr(eng, "stage", counted_stage)
    parked_peak = [0]
    real_park = eng.segment_wait

    def counted_wait(loop):
        fut = real_park(loop)
        parked_peak[0] = max(parked_peak[0], len(eng._segment_waiters))
        return fut

    monkeypatch.setattr(eng, "segment_wait", counted_wait)
    n = 400
    xs = [np.random.default_rng(i).standard_normal(16384, dtype=np.float32) for i in range(n)]

    async def go():
        return await asyncio.gather(*[
            ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": "bp2", "node_id": i}) for i in range(n)])

    asyncio.run(ex.run(electrons.add_int, [1, 2], {}, {"dispatch_id": "bp2", "node_id": "open"}))   # opens the connection
    general.clear()
    base = eng.counters["batches"]
    out = asyncio.run(go())
    for i in range(n):
        assert np.array_equal(out[i], np.float32(2.0) * xs[i] + np.float32(1.0))
    assert not general, "electrons left the fused path under back-pressure"
    assert parked_peak[0] >= n - 4 * 16                        # it really was back-pressured
    assert eng.counters["batches"] - base <= 2 * (n // 16)     # batches stay (nearly) full
    del out
    import gc

    gc.collect()
    assert sorted(eng._free_segments) == [0, 1, 2, 3] and not eng._segment_waiters


def test_cancelling_parked_electrons_never_strands_the_others(fake_lib, tmp_path):
    """Electrons parked for a staging segment that are cancelled — also just after having been woken — hand
    the wake-up on: everybody else still runs, and nothing stays claimed."""
    import time

    fake_lib.delay_polls = 10 ** 9
    real_wait = fake_lib.b2_wait

    def slow_wait(dev, ticket):
        time.sleep(0.003)
        return real_wait(dev, ticket)

    fake_lib.b2_wait = slow_wait
    eng = engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    asyncio.run(ex.run(electrons.add_int, [1, 2], {}, {"dispatch_id": "cp", "node_id": "open"}))
    n = 300
    xs = [np.random.default_rng(i).standard_normal(16384, dtype=np.float32) for i in range(n)]
    doomed = set(range(70, 300, 3))                           # all of them beyond the first 64: parked

    async def go():
        loop = asyncio.get_running_loop()
        tasks = [loop.create_task(ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": "cp", "node_id": i}))
                 for i in range(n)]
        await asyncio.sleep(0)                                # everybody staged or parked
        assert len(eng._segment_waiters) >= n - 64
        for i in sorted(doomed)[::2]:
            tasks[i].cancel()                                 # cancelled while parked
        # the rest of the doomed ones are cancelled the moment a release has woken them
        real_wake = eng._wake_waiters

        def wake_then_cancel():
            real_wake()
            for i in sorted(doomed)[1::2]:
                if not tasks[i].done():
                    tasks[i].cancel()

        eng._wake_waiters = wake_then_cancel
        try:
            return await asyncio.wait_for(asyncio.gather(*tasks, return_exceptions=True), timeout=60)
        finally:
            del eng._wake_waiters

    out = asyncio.run(go())
    for i in range(n):
        if i in doomed:
            assert isinstance(out[i], asyncio.CancelledError)
        else:
            assert np.array_equal(out[i], np.float32(2.0) * xs[i] + np.float32(1.0)), i
    del out
    import gc

    gc.collect()
    assert sorted(eng._free_segments) == [0, 1, 2, 3] and sorted(eng._free_out) == [0, 1, 2, 3]
    assert not any(not f.done() for _, f in eng._segment_waiters) and not impl._live


def test_a_woken_then_cancelled_electron_hands_its_wakeup_on(fake_lib, tmp_path):
    """One pair comes back, one parked electron is woken for it and cancelled before it runs: the next
    parked electron must get the pair (nothing else will ever wake it)."""
    eng = engine_mod.get_engine(0, **{**SMALL, "max_batch": 1})         # one electron per batch: one wake-up per pair
    ex = make_executor(tmp_path)
    asyncio.run(ex.run(electrons.add_int, [1, 2], {}, {"dispatch_id": "hw", "node_id": "open"}))
    with eng._seg_lock:
        ins, outs = list(eng._free_segments), list(eng._free_out)
        eng._free_segments.clear()
        eng._free_out.clear()
    xs = [np.full(256, i, dtype=np.float32) for i in range(3)]

    async def go():
        loop = asyncio.get_running_loop()
        tasks = [loop.create_task(ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": "hw", "node_id": i}))
                 for i in range(3)]
        await asyncio.sleep(0)
        assert len(eng._segment_waiters) == 3
        with eng._seg_lock:
            eng._free_segments.append(ins.pop())
            eng._free_out.append(outs.pop())
        eng._wake_waiters()                                   # wakes exactly one (budget = 1 pair x max_batch 1)
        assert sum(f.done() for _, f in eng._segment_waiters) == 0 and len(eng._segment_waiters) == 2
        tasks[0].cancel()                                     # ... which is cancelled before it gets to run
        done = await asyncio.wait_for(asyncio.gather(*tasks, return_exceptions=True), timeout=20)
        return done

    try:
        out = asyncio.run(go())
    finally:
        with eng._seg_lock:
            eng._free_segments.extend(ins)
            eng._free_out.extend(outs)
    assert isinstance(out[0], asyncio.CancelledError)
    assert np.array_equal(out[1], np.float32(2.0) * xs[1] + np.float32(1.0))
    assert np.array_equal(out[2], np.float32(2.0) * xs[2] + np.float32(1.0))
