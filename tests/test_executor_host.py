"""Host logic of the executor, on CPU.

Part 1 re-expresses the reference's seven unit tests
(/root/reference/tests/ssh_test.py) against the new SSHExecutor with plain
``asyncio.run`` + ``unittest.mock`` (pytest-asyncio / pytest-mock are not in this
image).  Part 2 drives the engine through a numpy test double of the C ABI
(tests/fake_b2lib.py) to cover batching, segment recycling, completion and the
failure paths without a GPU.
"""

import asyncio
import hashlib
import json
import os
import tempfile
from pathlib import Path
from unittest.mock import AsyncMock, MagicMock, mock_open, patch

import numpy as np
import pytest

from covalent_ssh_plugin import SSHExecutor
from covalent_ssh_plugin_b200 import _b2lib, electrons, engine as engine_mod
from tests.fake_b2lib import FakeLib
from tests.golden import cases

config_data = {
    "dispatcher.cache_dir": "cache_dir",
    "executors.ssh.username": "centos",
    "executors.ssh.hostname": "12.12.12.12",
    "executors.ssh.ssh_key_file": "~/.ssh/id_rsa",
    "executors.ssh.remote_cache": "/home/centos",
    "executors.ssh.python_path": "python3.8",
    "executors.ssh.conda_env": "py-3.8",
    "executors.ssh.remote_workdir": "covalent-workdir",
    "executors.ssh.create_unique_workdir": False,
}


def get_config_mock(key):
    return config_data[key]


CFG = "covalent_ssh_plugin.ssh.get_config"     # the reference tests' patch target still works
META = {"dispatch_id": -1, "node_id": -1}


# ===================================================================== part 1 ===
def test_init(tmp_path):
    key_path = tmp_path / "key_file"
    key_path.touch()
    with patch(CFG, side_effect=get_config_mock):
        executor = SSHExecutor(username="user", hostname="host", ssh_key_file=str(key_path),
                               create_unique_workdir=True)
    assert executor.username == "user"
    assert executor.hostname == "host"
    assert executor.ssh_key_file == str(key_path)
    assert executor.remote_cache == config_data["executors.ssh.remote_cache"]
    assert executor.python_path == config_data["executors.ssh.python_path"]
    assert executor.run_local_on_ssh_fail is False
    assert executor.remote_workdir == config_data["executors.ssh.remote_workdir"]
    assert executor.create_unique_workdir is True
    assert executor.do_cleanup is True
    # every attribute is plain data (executors are shipped as JSON by Covalent)
    json.dumps(executor.__dict__)


def test_constructor_signature_matches_reference():
    import inspect

    params = list(inspect.signature(SSHExecutor.__init__).parameters.items())[1:]
    assert [(n, p.default) for n, p in params] == [
        ("username", inspect._empty), ("hostname", inspect._empty), ("ssh_key_file", None),
        ("cache_dir", None), ("python_path", ""), ("conda_env", None), ("remote_cache", ""),
        ("run_local_on_ssh_fail", False), ("remote_workdir", ""), ("create_unique_workdir", None),
        ("poll_freq", 15), ("do_cleanup", True), ("retry_connect", True),
        ("max_connection_attempts", 5), ("retry_wait_time", 5)]
    import covalent_ssh_plugin.ssh as mod

    assert mod.EXECUTOR_PLUGIN_NAME == "SSHExecutor"
    assert set(mod._EXECUTOR_PLUGIN_DEFAULTS) == {
        "username", "hostname", "ssh_key_file", "cache_dir", "python_path", "conda_env", "remote_cache",
        "run_local_on_ssh_fail", "remote_workdir", "create_unique_workdir"}


def test_on_ssh_fail(tmp_path):
    def simple_task(x):
        return x ** 2

    key = tmp_path / "k"
    key.touch()
    with patch(CFG, side_effect=get_config_mock), \
            patch("covalent_ssh_plugin.SSHExecutor._validate_credentials", return_value=True):
        executor = SSHExecutor(username="user", hostname="host", ssh_key_file=str(key),
                               run_local_on_ssh_fail=True, retry_connect=False)
        result = asyncio.run(executor.run(function=simple_task, args=[5], kwargs={}, task_metadata=META))
        assert result == 25
        executor.run_local_on_ssh_fail = False
        with pytest.raises(RuntimeError):
            asyncio.run(executor.run(function=simple_task, args=[5], kwargs={}, task_metadata=META))


def test_nonzero_exit_status(tmp_path):
    class _FakeConn:
        async def run(self, *_):
            return MagicMock(stdout="Python 3.8", stderr="")

        def close(self):
            pass

        async def wait_closed(self):
            return True

    class _FakeResultFailed:
        stderr = "Fake error message"
        exit_status = 1

    key = tmp_path / "k"
    key.touch()
    base = "covalent_ssh_plugin.ssh.SSHExecutor."
    with patch(CFG, side_effect=get_config_mock), \
            patch(base + "_validate_credentials", return_value=True), \
            patch(base + "_client_connect", return_value=(True, _FakeConn())), \
            patch(base + "_write_function_files", return_value=[""] * 5), \
            patch(base + "_upload_task", return_value=True), \
            patch(base + "submit_task", return_value=_FakeResultFailed()):
        executor = SSHExecutor(username="user", hostname="host", ssh_key_file=str(key),
                               run_local_on_ssh_fail=False, retry_connect=False)
        executor.conda_env = None
        with pytest.raises(RuntimeError, match="Fake error message"):
            asyncio.run(executor.run(function=lambda: "dummy", args=[5], kwargs={}, task_metadata=META))


def test_client_connect():
    with patch(CFG, side_effect=get_config_mock):
        executor = SSHExecutor(username="user", hostname="cuda:0", ssh_key_file="non-existent_key",
                               retry_connect=False)
    with pytest.raises(RuntimeError):
        asyncio.run(executor._client_connect())
    with patch("os.path.exists", return_value=True), \
            patch("covalent_ssh_plugin.ssh._open_engine", AsyncMock()):
        connected, _ = asyncio.run(executor._client_connect())
    assert connected is True


def test_client_connect_retry_attempts(tmp_path):
    state = {"err_counter": -1, "succeed_after": 0}

    async def _mock_open(*args, **kwargs):
        if state["err_counter"] < 0:
            return "immediate_connection_object"
        if state["err_counter"] > state["succeed_after"] - 1:
            return "eventual_connection_object"
        err = ConnectionRefusedError("refused") if state["err_counter"] % 2 == 0 else OSError("unreachable")
        state["err_counter"] += 1
        raise err

    key = tmp_path / "k"
    key.touch()
    with patch(CFG, side_effect=get_config_mock), patch("covalent_ssh_plugin.ssh._open_engine", side_effect=_mock_open):
        executor = SSHExecutor(username="user", hostname="cuda:0", ssh_key_file=str(key), retry_connect=False)
        executor.retry_wait_time = 0.01
        assert asyncio.run(executor._client_connect()) == (True, "immediate_connection_object")

        state.update(err_counter=0, succeed_after=3)
        executor.retry_connect = True
        assert asyncio.run(executor._client_connect()) == (True, "eventual_connection_object")

        state.update(err_counter=0)
        executor.retry_connect = False
        with pytest.raises(ConnectionRefusedError):
            asyncio.run(executor._client_connect())

        state.update(err_counter=0, succeed_after=executor.max_connection_attempts + 1)
        executor.retry_connect = True
        assert asyncio.run(executor._client_connect()) == (False, None)


def test_current_remote_workdir():
    mock_conn = MagicMock()
    mock_conn.wait_closed = AsyncMock(return_value=True)

    async def mock_submit_task(conn, file):
        return MagicMock(stderr="", exit_status=0)

    base = "covalent_ssh_plugin.ssh.SSHExecutor."
    with patch(CFG, side_effect=get_config_mock), \
            patch(base + "_client_connect", return_value=(True, mock_conn)), \
            patch(base + "_validate_credentials", return_value=True), \
            patch(base + "_upload_task"), \
            patch(base + "submit_task", side_effect=mock_submit_task), \
            patch(base + "_poll_task", return_value=True), \
            patch(base + "query_result", return_value=(5, None)):
        with tempfile.TemporaryDirectory() as tmp_dir:
            executor = SSHExecutor(username="user", hostname="host", ssh_key_file="key_file",
                                   remote_workdir=tmp_dir, create_unique_workdir=True, do_cleanup=False)
        executor.conda_env = None

        def simple_task(x):
            return x

        dispatch_id, node_id = "asdf", 1
        expected = os.path.join(tmp_dir, dispatch_id, f"node_{node_id}")
        with patch.object(SSHExecutor, "_write_function_files", return_value=("a", "b", "c", "d", "e")) as wff:
            out = asyncio.run(executor.run(simple_task, [5], {}, {"dispatch_id": dispatch_id, "node_id": node_id}))
            wff.assert_called_with(f"{dispatch_id}_{node_id}", simple_task, [5], {}, expected)
        assert out == 5


def test_file_writes():
    with patch(CFG, side_effect=get_config_mock):
        executor = SSHExecutor(username="user", hostname="host", ssh_key_file="key_file")

    def simple_task(x):
        return x

    operation_id = "dispatchid_taskid"
    with patch("builtins.open", new_callable=mock_open()):
        names = executor._write_function_files(operation_id, simple_task, [5], {})
    function_file, script_file, remote_function_file, remote_script_file, remote_result_file = names
    assert script_file == os.path.join(executor.cache_dir, f"exec_{operation_id}.py")
    assert remote_script_file == os.path.join(executor.remote_cache, f"exec_{operation_id}.py")
    assert function_file == os.path.join(executor.cache_dir, f"function_{operation_id}.pkl")
    assert remote_function_file == os.path.join(executor.remote_cache, f"function_{operation_id}.pkl")
    assert remote_result_file == os.path.join(executor.remote_cache, f"result_{operation_id}.pkl")


def test_cancel_with_nothing_in_flight_reports_false():
    """The reference raises NotImplementedError here (ssh.py:460-464); this executor can
    cancel (tests below), and with nothing in flight there is nothing to do."""
    with patch(CFG, side_effect=get_config_mock):
        executor = SSHExecutor(username="user", hostname="host", ssh_key_file="key_file")
    assert asyncio.run(executor.cancel()) is False
    assert asyncio.run(executor.cancel({"dispatch_id": "d", "node_id": 1})) is False


def test_hostname_resolution():
    import socket

    from covalent_ssh_plugin_b200.ssh import resolve_device

    assert resolve_device("cuda:3") == 3 and resolve_device("GPU:1") == 1 and resolve_device("b200:7") == 7
    assert resolve_device("gpu2") == 2
    with patch.dict(os.environ, {"LOCAL_RANK": "5"}):
        assert resolve_device("localhost") == 5
    for bad in ("host", "12.12.12.12", "cuda:x", "node17"):
        with pytest.raises(socket.gaierror):
            resolve_device(bad)


# ===================================================================== part 2 ===
GOLDEN = json.loads((Path(__file__).parent / "golden" / "golden_v1.json").read_text())["cases"]


@pytest.fixture()
def fake_lib():
    from covalent_ssh_plugin_b200 import ssh as impl

    impl._inflight.clear()          # part-1 tests call _write_function_files on its own
    engine_mod.shutdown_all()
    saved = _b2lib._lib
    lib = FakeLib(n_devices=2)
    _b2lib._lib = lib
    yield lib
    engine_mod.shutdown_all()
    _b2lib._lib = saved


SMALL = dict(arena_bytes=48 << 20, staging_bytes=16 << 20, n_segments=4, max_batch=16, queue_cap=64)


def make_executor(tmp_path, **kw):
    key = tmp_path / "id_stub"
    key.touch()
    return SSHExecutor(username="u", hostname=kw.pop("hostname", "cuda:0"), ssh_key_file=str(key),
                       cache_dir=str(tmp_path), **kw)


def golden_ok(cid, value):
    g = GOLDEN[cid]["result"]
    if g["type"] == "int":
        return type(value) is int and value == int(g["value"])
    if g["type"] == "numpy_scalar":
        ref = np.frombuffer(bytes.fromhex(g["hex"]), dtype=g["dtype"])[0]
        return isinstance(value, np.generic) and abs(float(value) - float(ref)) <= 1e-6 * abs(float(ref))
    return (isinstance(value, np.ndarray) and str(value.dtype) == g["dtype"] and list(value.shape) == g["shape"]
            and bool(value.flags.c_contiguous) == g["c_contiguous"]
            and hashlib.sha256(value.tobytes(order="A")).hexdigest() == g["sha256"])


def test_golden_cases_through_host_logic(fake_lib, tmp_path):
    """Encoding, staging offsets, task records, decode: the full host path on the
    numpy model of the device, against what the real reference returned."""
    engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    todo = [c for c in cases.CASES if c[1] not in ("raise_value_error", "plain_python")
            and "overflow" not in c[0] and "beyond" not in c[0] and c[0] != "int_mul_min_neg1"]

    async def go():
        return await asyncio.gather(*[
            ex.run(cases.ELECTRONS[e], cases.build_args(r), {}, {"dispatch_id": "g", "node_id": i})
            for i, (_, e, r) in enumerate(todo)])

    for (cid, _, _), v in zip(todo, asyncio.run(go())):
        assert golden_ok(cid, v), cid
    eng = engine_mod.existing_engine(0)
    assert eng.counters["electrons"] == len(todo) and eng.counters["batches"] >= len(todo) // 16
    assert sorted(eng._free_segments) == [0, 1, 2, 3]            # every segment came back
    from covalent_ssh_plugin_b200 import ssh as impl

    assert not impl._inflight


def test_more_electrons_than_segments_can_hold(fake_lib, tmp_path):
    """Back-pressure: 400 x 64 KiB with 4 x 4 MiB segments and slow completion."""
    fake_lib.delay_polls = 3
    engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    xs = [np.random.default_rng(i).standard_normal(16384, dtype=np.float32) for i in range(400)]

    async def go():
        return await asyncio.gather(*[
            ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": "bp", "node_id": i}) for i in range(400)])

    out = asyncio.run(go())
    for i in range(400):
        assert np.array_equal(out[i], np.float32(2.0) * xs[i] + np.float32(1.0))
    assert sorted(engine_mod.existing_engine(0)._free_segments) == [0, 1, 2, 3]


def test_sequential_runs_and_two_devices(fake_lib, tmp_path):
    engine_mod.get_engine(0, **SMALL)
    engine_mod.get_engine(1, **SMALL)
    a, b = make_executor(tmp_path, hostname="cuda:0"), make_executor(tmp_path, hostname="gpu:1")
    for i in range(5):
        assert asyncio.run(a.run(electrons.add_int, [i, 1], {}, {"dispatch_id": "s", "node_id": i})) == i + 1
        assert asyncio.run(b.run(electrons.mul_int, [i, 3], {}, {"dispatch_id": "s", "node_id": i})) == 3 * i
    assert engine_mod.existing_engine(0).counters["electrons"] == 5
    assert engine_mod.existing_engine(1).counters["electrons"] == 5


def test_overflow_and_unregistered_take_the_failure_path(fake_lib, tmp_path):
    engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    big = (1 << 63) - 1
    with pytest.raises(RuntimeError, match="int64"):
        asyncio.run(ex.run(electrons.add_int, [big, 1], {}, META))
    with pytest.raises(RuntimeError, match="not a registered device electron"):
        asyncio.run(ex.run(cases.plain_python, ["a", "b"], {}, META))
    ex.run_local_on_ssh_fail = True
    assert asyncio.run(ex.run(electrons.add_int, [big, 1], {}, META)) == big + 1
    assert asyncio.run(ex.run(cases.plain_python, ["a", "b"], {}, META)) == "a, b"
    with pytest.raises(ValueError):                                  # the electron's own exception surfaces
        asyncio.run(ex.run(cases.raise_value_error, ["boom", 7], {}, META))
    assert sorted(engine_mod.existing_engine(0)._free_segments) == [0, 1, 2, 3]


def test_no_cleanup_still_releases_staging(fake_lib, tmp_path):
    engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path, do_cleanup=False)
    for i in range(10):
        assert asyncio.run(ex.run(electrons.affine_int, [i], {}, {"dispatch_id": "n", "node_id": i})) == 2 * i + 1
    assert sorted(engine_mod.existing_engine(0)._free_segments) == [0, 1, 2, 3]


def test_device_init_retry_then_success(fake_lib, tmp_path):
    fake_lib.fail_init = 2
    ex = make_executor(tmp_path, retry_wait_time=0.01)
    with patch.dict(engine_mod.DEFAULTS, SMALL):
        assert asyncio.run(ex.run(electrons.add_int, [1, 2], {}, META)) == 3
    assert [c for c in fake_lib.calls if c[0] == "init"] == [("init", 0)] * 3


def test_device_init_failure_without_retry_propagates(fake_lib, tmp_path):
    fake_lib.fail_init = 1
    ex = make_executor(tmp_path, retry_connect=False)
    with patch.dict(engine_mod.DEFAULTS, SMALL), pytest.raises(ConnectionRefusedError):
        asyncio.run(ex.run(electrons.add_int, [1, 2], {}, META))


def test_missing_gpu_is_a_failed_connection(fake_lib, tmp_path):
    ex = make_executor(tmp_path, hostname="cuda:7", retry_wait_time=0.01, max_connection_attempts=2)
    with patch.dict(engine_mod.DEFAULTS, SMALL), pytest.raises(RuntimeError, match="Could not connect"):
        asyncio.run(ex.run(electrons.add_int, [1, 2], {}, META))


def test_missing_library_is_never_swallowed(tmp_path):
    """No .so -> LibraryMissing, even with run_local_on_ssh_fail=True."""
    engine_mod.shutdown_all()
    saved = _b2lib._lib
    _b2lib._lib = None
    try:
        with patch.dict(os.environ, {"B2EXEC_LIB": str(tmp_path / "nope.so")}):
            ex = make_executor(tmp_path, run_local_on_ssh_fail=True)
            with pytest.raises(_b2lib.LibraryMissing):
                asyncio.run(ex.run(electrons.add_int, [1, 2], {}, META))
    finally:
        _b2lib._lib = saved


# ------------------------------------------------------------ zero-copy results ---
def test_results_alias_pinned_egress_and_segments_recycle(fake_lib, tmp_path):
    eng = engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    xs = [np.random.default_rng(i).standard_normal(2048, dtype=np.float32) for i in range(40)]

    async def go():
        return await asyncio.gather(*[
            ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": "z", "node_id": i}) for i in range(40)])

    out = asyncio.run(go())
    lo, hi = eng._h_out_addr, eng._h_out_addr + eng.staging_out
    aliased = [lo <= a.ctypes.data < hi for a in out]
    assert any(aliased) and eng.counters["zero_copy_batches"] >= 1
    for i, a in enumerate(out):
        assert a.flags.writeable and np.array_equal(a, np.float32(2.0) * xs[i] + np.float32(1.0))
    # ingress segments are back already; egress segments come back as the arrays die
    assert sorted(eng._free_segments) == [0, 1, 2, 3]
    assert len(eng._free_out) < 4 and eng._leases_out >= 1
    keep = out[0]
    expect = keep.copy()
    del out, a
    assert len(eng._free_out) == 3 and eng._leases_out == 1       # one batch still aliased by `keep`
    # new work must not overwrite memory a live result still points into
    asyncio.run(go())
    assert np.array_equal(keep, expect)
    del keep
    assert sorted(eng._free_out) == [0, 1, 2, 3] and eng._leases_out == 0


def test_copy_out_when_egress_pool_is_tight(fake_lib, tmp_path):
    """Callers that hoard results must not starve the pinned pool: with little
    room left the engine copies results out instead of leasing segments."""
    eng = engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    hoard = []
    for r in range(12):      # 12 batches >> 4 egress segments
        x = np.full(1024, r, dtype=np.float32)
        hoard.append(asyncio.run(ex.run(electrons.plus_one, [x], {}, {"dispatch_id": "h", "node_id": r})))
    assert all(np.array_equal(h, np.full(1024, r + 1, dtype=np.float32)) for r, h in enumerate(hoard))
    assert len(eng._free_out) >= 1                       # never exhausted
    lo, hi = eng._h_out_addr, eng._h_out_addr + eng.staging_out
    assert sum(lo <= h.ctypes.data < hi for h in hoard) <= 3


def test_close_with_live_results_detaches_the_pinned_buffer(fake_lib, tmp_path):
    eng = engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    x = np.arange(4096, dtype=np.float32)
    r = asyncio.run(ex.run(electrons.scale_add, [x], {}, {"dispatch_id": "c", "node_id": 0}))
    engine_mod.shutdown_all()
    assert getattr(fake_lib, "detached", None), "egress buffer must be handed over, not freed"
    assert np.array_equal(r, np.float32(2.0) * x + np.float32(1.0))


def test_zero_copy_can_be_disabled(fake_lib, tmp_path):
    eng = engine_mod.get_engine(0, zero_copy=False, **SMALL)
    ex = make_executor(tmp_path)
    x = np.arange(4096, dtype=np.float32)
    r = asyncio.run(ex.run(electrons.scale_add, [x], {}, {"dispatch_id": "c", "node_id": 0}))
    assert r.flags.owndata and sorted(eng._free_out) == [0, 1, 2, 3]


def test_device_resident_operand_through_run(fake_lib, tmp_path):
    """A shared dependency uploaded once (engine.put / broadcast_shared) is an
    ordinary electron argument afterwards: nothing is staged for it."""
    eng = engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    shared = np.random.default_rng(1).standard_normal(3000, dtype=np.float32)
    da = eng.put(shared)
    xs = [np.random.default_rng(10 + i).standard_normal(3000, dtype=np.float32) for i in range(20)]
    before = eng.counters["bytes_in"]

    async def go():
        return await asyncio.gather(*[
            ex.run(electrons.vec_add, [xs[i], da], {}, {"dispatch_id": "sh", "node_id": i}) for i in range(20)])

    out = asyncio.run(go())
    for i in range(20):
        assert np.array_equal(out[i], xs[i] + shared)
    assert eng.counters["bytes_in"] - before == 20 * 12032          # only the 20 private operands were staged
    da.free()
    with pytest.raises(RuntimeError, match="freed"):
        asyncio.run(ex.run(electrons.vec_add, [xs[0], da], {}, META))


def test_broadcast_shared_single_process(fake_lib):
    from covalent_ssh_plugin_b200 import distributed as D

    eng = engine_mod.get_engine(0, **SMALL)
    x = np.arange(12, dtype=np.float32).reshape(3, 4)
    da = D.broadcast_shared(eng, x)
    assert da.shape == (3, 4) and np.array_equal(da.numpy(), x)


# ------------------------------------------------------------------ host pool ---
def _where_am_i():
    import os
    return os.getcwd()


def _die():
    import os
    os._exit(7)


def test_host_pool_is_opt_in_and_mirrors_exec_py(fake_lib, tmp_path):
    """Unregistered electrons: refused by default; with the pool switched on they run
    in warm worker processes with exec.py's semantics (workdir, in-band exception,
    crash = non-zero exit).  Device electrons never go there."""
    import cloudpickle

    from covalent_ssh_plugin_b200 import hostpool
    import tests.test_executor_host as me

    cloudpickle.register_pickle_by_value(me)
    cloudpickle.register_pickle_by_value(cases)
    engine_mod.get_engine(0, **SMALL)
    wd = tmp_path / "wd"
    ex = make_executor(tmp_path, remote_workdir=str(wd), create_unique_workdir=True)
    with pytest.raises(RuntimeError, match="not a registered device electron"):
        asyncio.run(ex.run(cases.plain_python, ["a", "b"], {}, {"dispatch_id": "hp", "node_id": 0}))
    with patch.dict(os.environ, {"B2_HOST_POOL": "1", "B2_HOST_POOL_WORKERS": "2"}):
        try:
            before = hostpool.stats["submitted"]
            assert asyncio.run(ex.run(cases.plain_python, ["a", "b"], {}, {"dispatch_id": "hp", "node_id": 1})) == "a, b"
            assert asyncio.run(ex.run(_where_am_i, [], {}, {"dispatch_id": "hp", "node_id": 2})) == str(wd / "hp" / "node_2")
            with pytest.raises(ValueError) as ei:
                asyncio.run(ex.run(cases.raise_value_error, ["boom", 7], {}, {"dispatch_id": "hp", "node_id": 3}))
            assert ei.value.args == ("boom", 7)
            with pytest.raises(RuntimeError, match="worker process died"):
                asyncio.run(ex.run(_die, [], {}, {"dispatch_id": "hp", "node_id": 4}))
            # the pool recovers, and device electrons still run on the device
            assert asyncio.run(ex.run(cases.plain_python, ["c", "d"], {}, {"dispatch_id": "hp", "node_id": 5})) == "c, d"
            n_dev = engine_mod.existing_engine(0).counters["electrons"]
            assert asyncio.run(ex.run(electrons.add_int, [2, 3], {}, {"dispatch_id": "hp", "node_id": 6})) == 5
            assert engine_mod.existing_engine(0).counters["electrons"] == n_dev + 1
            assert hostpool.stats["submitted"] - before == 5
        finally:
            hostpool.shutdown()
    assert not __import__("covalent_ssh_plugin_b200.ssh", fromlist=["_inflight"])._inflight


def test_large_operands_are_staged_by_threads(fake_lib, tmp_path):
    """Operands travel as copy jobs (source address, ingress offset, bytes) that the library's
    copy threads run when the batch is submitted — never inline on the event-loop thread — and
    results stay correct."""
    eng = engine_mod.get_engine(0, parallel_stage_bytes=64 << 10, **SMALL)
    ex = make_executor(tmp_path)
    n = 60
    xs = [np.random.default_rng(i).standard_normal(20000 + 31 * i, dtype=np.float32) for i in range(n)]   # ~80 KiB each
    ys = [np.random.default_rng(100 + i).standard_normal(20000 + 31 * i, dtype=np.float32) for i in range(n)]
    small = [np.random.default_rng(200 + i).standard_normal(100, dtype=np.float32) for i in range(n)]

    async def go():
        big = [ex.run(electrons.vec_mul, [xs[i], ys[i]], {}, {"dispatch_id": "big", "node_id": i}) for i in range(n)]
        tiny = [ex.run(electrons.scale_add, [small[i]], {}, {"dispatch_id": "tiny", "node_id": i}) for i in range(n)]
        return await asyncio.gather(*big, *tiny)

    out = asyncio.run(go())
    for i in range(n):
        assert np.array_equal(out[i], xs[i] * ys[i])
        assert np.array_equal(out[n + i], np.float32(2.0) * small[i] + np.float32(1.0))
    assert fake_lib.jobs_run >= n                # (electrons that had to wait for a segment copy small operands inline)
    del out
    assert sorted(eng._free_segments) == [0, 1, 2, 3]


def test_artifact_compatibility_mode(fake_lib, tmp_path):
    """do_cleanup=False + B2_EMIT_ARTIFACTS=1 leaves the reference's artefacts — local AND
    remote-side names (pinned by the reference's test_file_writes), its pickle formats, and a
    run script formatted from the template exactly as ssh.py:164-169 formats exec.py — and the
    emitted script really reproduces the result from the remote-side files."""
    import subprocess
    import sys

    import cloudpickle

    engine_mod.get_engine(0, **SMALL)
    cache, remote = tmp_path / "cache", tmp_path / "remote"
    cache.mkdir()
    key = tmp_path / "id_stub"
    key.touch()
    ex = SSHExecutor(username="u", hostname="cuda:0", ssh_key_file=str(key), cache_dir=str(cache),
                     remote_cache=str(remote), remote_workdir=str(tmp_path / "work"), do_cleanup=False)
    x = np.arange(64, dtype=np.float32)
    with patch.dict(os.environ, {"B2_EMIT_ARTIFACTS": "1"}):
        y = asyncio.run(ex.run(electrons.scale_add, [x], {}, {"dispatch_id": "art", "node_id": 3}))
    want = ["exec_art_3.py", "function_art_3.pkl", "result_art_3.pkl"]
    assert sorted(p.name for p in cache.iterdir()) == want
    assert sorted(p.name for p in remote.iterdir()) == want
    fn, args, kwargs = cloudpickle.load(open(cache / "function_art_3.pkl", "rb"))
    assert fn is electrons.scale_add and np.array_equal(args[0], x) and kwargs == {}
    res, exc = cloudpickle.load(open(remote / "result_art_3.pkl", "rb"))
    assert exc is None and np.array_equal(res, y)
    script = (remote / "exec_art_3.py").read_text()
    assert str(remote / "result_art_3.pkl") in script and str(remote / "function_art_3.pkl") in script
    assert str(tmp_path / "work") in script and "{" not in script.split("import os")[1]
    # the script is the real thing: running it rewrites result_<id>.pkl from function_<id>.pkl
    (remote / "result_art_3.pkl").unlink()
    env = dict(os.environ, PYTHONPATH=str(Path(__file__).resolve().parents[1]))
    assert subprocess.run([sys.executable, str(remote / "exec_art_3.py")], env=env).returncode == 0
    res2, exc2 = cloudpickle.load(open(remote / "result_art_3.pkl", "rb"))
    assert exc2 is None and np.array_equal(res2, y) and (tmp_path / "work").is_dir()
    # default: nothing is written
    asyncio.run(ex.run(electrons.scale_add, [x], {}, {"dispatch_id": "art", "node_id": 4}))
    assert not [p for p in list(cache.iterdir()) + list(remote.iterdir()) if "art_4" in p.name]


def test_large_elementwise_electrons_are_sliced(fake_lib, tmp_path):
    """A 5 MiB electron goes to the device as 1 MiB slices (so that a handful of huge
    electrons still occupy every CTA); results and failure reporting are unchanged."""
    eng = engine_mod.get_engine(0, arena_bytes=160 << 20, staging_bytes=64 << 20, n_segments=4, max_batch=64,
                                queue_cap=256)
    ex = make_executor(tmp_path)
    n = (5 << 20) // 4 + 37                                   # not a multiple of anything convenient
    x = np.random.default_rng(0).standard_normal(n, dtype=np.float32)
    y = np.random.default_rng(1).standard_normal(n, dtype=np.float32)
    before = eng.counters["electrons"]
    out = asyncio.run(ex.run(electrons.axpy_half, [x, y], {}, {"dispatch_id": "big", "node_id": 0}))
    assert np.array_equal(out, np.float32(0.5) * x + y)
    assert eng.counters["electrons"] - before >= 5            # task records, one per slice
    z = asyncio.run(ex.run(electrons.identity, [x], {}, {"dispatch_id": "big", "node_id": 1}))
    assert np.array_equal(z, x)


def test_launch_failure_reaches_every_electron_of_the_batch(fake_lib, tmp_path):
    """A CUDA failure while shipping a batch is raised from every run() of that batch (it is
    not turned into a silent local run), the staging segments come back, and the engine keeps
    working afterwards."""
    eng = engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path, run_local_on_ssh_fail=True)     # must NOT mask a device failure
    fake_lib.fail_launch = 1

    async def go():
        return await asyncio.gather(*[
            ex.run(electrons.add_int, [i, 1], {}, {"dispatch_id": "boom", "node_id": i}) for i in range(10)],
            return_exceptions=True)

    out = asyncio.run(go())
    assert all(isinstance(o, _b2lib.B2Error) for o in out)
    assert sorted(eng._free_segments) == [0, 1, 2, 3] and sorted(eng._free_out) == [0, 1, 2, 3]
    assert asyncio.run(ex.run(electrons.add_int, [2, 3], {}, {"dispatch_id": "after", "node_id": 0})) == 5


# ------------------------------------------------------- calls as Covalent core makes them ---
class FakeTransportable:
    """Duck-type of covalent's TransportableObject (serialises on construction)."""

    def __init__(self, obj):
        import cloudpickle

        self._blob = cloudpickle.dumps(obj)

    def get_deserialized(self):
        import cloudpickle

        return cloudpickle.loads(self._blob)


def _covalent_wrapper(user_fn, call_before, call_after, *args, **kwargs):
    fn = user_fn.get_deserialized()
    out = fn(*[a.get_deserialized() for a in args], **{k: v.get_deserialized() for k, v in kwargs.items()})
    return FakeTransportable(out)


def test_covalent_style_wrapped_call_runs_on_the_device(fake_lib, tmp_path):
    """functools.partial(wrapper, TransportableObject(fn), [], []) with TransportableObject
    arguments (how Covalent core calls run) is unwrapped, runs on the device, and the result
    comes back wrapped the same way."""
    import functools

    eng = engine_mod.get_engine(0, **SMALL)
    ex = make_executor(tmp_path)
    x = np.arange(300, dtype=np.float32)
    call = functools.partial(_covalent_wrapper, FakeTransportable(electrons.scale_add), [], [])
    before = eng.counters["electrons"]
    out = asyncio.run(ex.run(call, [FakeTransportable(x)], {}, {"dispatch_id": "cov", "node_id": 0}))
    assert isinstance(out, FakeTransportable)
    assert np.array_equal(out.get_deserialized(), np.float32(2.0) * x + np.float32(1.0))
    assert eng.counters["electrons"] == before + 1
    # with call_before hooks the electron needs a Python host: failure path (or local run if opted in)
    hooked = functools.partial(_covalent_wrapper, FakeTransportable(electrons.scale_add), ["pip install x"], [])
    with pytest.raises(RuntimeError):
        asyncio.run(ex.run(hooked, [FakeTransportable(x)], {}, {"dispatch_id": "cov", "node_id": 1}))
    ex.run_local_on_ssh_fail = True
    out = asyncio.run(ex.run(hooked, [FakeTransportable(x)], {}, {"dispatch_id": "cov", "node_id": 2}))
    assert np.array_equal(out.get_deserialized(), np.float32(2.0) * x + np.float32(1.0))
