"""GPU parity of the whole hot path through the reference-facing API
(SSHExecutor.run on cuda:0 -> engine -> C ABI -> kernels) against the golden
vectors the unmodified reference produced (tests/golden/golden_v1.json)."""

import asyncio
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from tests.golden import cases

pytestmark = pytest.mark.gpu

GOLDEN = json.loads((Path(__file__).parent / "golden" / "golden_v1.json").read_text())["cases"]


@pytest.fixture(scope="module")
def executor(gpu_count, tmp_path_factory):
    if gpu_count < 1:
        pytest.skip("no CUDA device on this box")
    from covalent_ssh_plugin_b200 import SSHExecutor
    from covalent_ssh_plugin_b200.engine import shutdown_all

    tmp = tmp_path_factory.mktemp("exec")
    key = tmp / "id_stub"
    key.touch()
    yield SSHExecutor(username="user", hostname="cuda:0", ssh_key_file=str(key), cache_dir=str(tmp))
    shutdown_all()


def check_against_golden(cid, value):
    g = GOLDEN[cid]["result"]
    if g["type"] == "int":
        assert type(value) is int and value == int(g["value"]), cid
    elif g["type"] == "numpy_scalar":
        assert isinstance(value, np.generic) and str(value.dtype) == g["dtype"], cid
        ref = np.frombuffer(bytes.fromhex(g["hex"]), dtype=g["dtype"])[0]
        assert abs(float(value) - float(ref)) <= 1e-6 * max(abs(float(ref)), 1e-30), (cid, value, ref)
    elif g["type"] == "ndarray":
        assert isinstance(value, np.ndarray) and str(value.dtype) == g["dtype"], cid
        assert list(value.shape) == g["shape"], cid
        assert bool(value.flags.c_contiguous) == g["c_contiguous"], cid
        assert bool(value.flags.f_contiguous) == g["f_contiguous"], cid
        assert hashlib.sha256(value.tobytes(order="A")).hexdigest() == g["sha256"], cid
    else:
        assert value == g["value"], cid


DEVICE_CASES = [c for c in cases.CASES
                if c[1] not in ("raise_value_error", "plain_python")
                and "overflow" not in c[0] and "beyond" not in c[0] and c[0] != "int_mul_min_neg1"]
FAIL_PATH_CASES = [c for c in cases.CASES
                   if "overflow" in c[0] or "beyond" in c[0] or c[0] == "int_mul_min_neg1"]


def test_golden_cases_on_device(executor):
    from covalent_ssh_plugin_b200.engine import existing_engine

    async def go():
        coros = [executor.run(cases.ELECTRONS[e], cases.build_args(r), {}, {"dispatch_id": "g", "node_id": i})
                 for i, (_, e, r) in enumerate(DEVICE_CASES)]
        return await asyncio.gather(*coros)

    before = existing_engine(0).stats().electrons if existing_engine(0) else 0
    results = asyncio.run(go())
    for (cid, _, _), value in zip(DEVICE_CASES, results):
        check_against_golden(cid, value)
    assert existing_engine(0).stats().electrons - before == len(DEVICE_CASES)  # all ran on the GPU


def test_overflow_takes_the_reference_failure_path(executor):
    """ints that do not fit int64: RuntimeError by default, exact Python result
    when the user opted into run_local_on_ssh_fail (ssh.py:181-208)."""
    for cid, e, r in FAIL_PATH_CASES:
        with pytest.raises(RuntimeError):
            asyncio.run(executor.run(cases.ELECTRONS[e], cases.build_args(r), {}, {"dispatch_id": "o", "node_id": 0}))
    executor.run_local_on_ssh_fail = True
    try:
        for cid, e, r in FAIL_PATH_CASES:
            v = asyncio.run(executor.run(cases.ELECTRONS[e], cases.build_args(r), {}, {"dispatch_id": "o", "node_id": 1}))
            check_against_golden(cid, v)
    finally:
        executor.run_local_on_ssh_fail = False


def test_unregistered_electron_is_not_silently_run(executor):
    with pytest.raises(RuntimeError, match="not a registered device electron"):
        asyncio.run(executor.run(cases.plain_python, ["Hello", "World"], {}, {"dispatch_id": "u", "node_id": 0}))


def test_fan_out_10k_int_electrons(executor):
    """BASELINE config 2: 10 000 scalar-int electrons, bit-exact."""
    from covalent_ssh_plugin_b200 import electrons

    n = 10000

    async def go():
        return await asyncio.gather(*[
            executor.run(electrons.add_int, [i, 2 * i + 1], {}, {"dispatch_id": "c2", "node_id": i})
            for i in range(n)])

    out = asyncio.run(go())
    assert all(type(v) is int for v in out)
    assert out == [3 * i + 1 for i in range(n)]


def test_fan_out_64k_arrays(executor):
    """BASELINE config 3 at full size: 10 000 x 64 KiB fp32 in/out, every result bit-exact."""
    from covalent_ssh_plugin_b200 import electrons

    n = 10000
    xs = [np.random.default_rng([0, i]).standard_normal(16384, dtype=np.float32) for i in range(n)]

    async def go():
        return await asyncio.gather(*[
            executor.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": "c3", "node_id": i})
            for i in range(n)])

    out = asyncio.run(go())
    for i in range(n):
        assert np.array_equal(out[i], np.float32(2.0) * xs[i] + np.float32(1.0))


def test_same_seeded_inputs_through_oracle_port_and_gpu(executor, tmp_path):
    """The CUDA path and the oracle (the CPU restatement of the reference's per-electron path,
    oracle/reference_port.py) on the same seeded inputs: bit-exact for integer and elementwise
    fp32 electrons, <= 1e-6 relative for the tree-ordered sum."""
    from covalent_ssh_plugin_b200 import electrons as E
    from oracle import reference_port as P

    repo = str(Path(__file__).resolve().parents[1])
    rng = np.random.default_rng(20260921)
    x = rng.standard_normal(70001, dtype=np.float32)
    y = rng.standard_normal(70001, dtype=np.float32)
    z = rng.random((129, 257), dtype=np.float32)
    calls = [(E.add_int, [int(rng.integers(-2 ** 62, 2 ** 62)), int(rng.integers(-2 ** 61, 2 ** 61))]),
             (E.mul_int, [int(rng.integers(-2 ** 31, 2 ** 31)), int(rng.integers(-2 ** 31, 2 ** 31))]),
             (E.affine_int, [int(rng.integers(-2 ** 61, 2 ** 61))]),
             (E.scale_add, [x]), (E.vec_add, [x, y]), (E.vec_mul, [x, y]), (E.axpy_half, [x, y]),
             (E.identity, [z]), (E.plus_one, [np.asfortranarray(z)]), (E.stack_sum, [[x, y, x, y, x]]),
             (E.vec_sum, [z.reshape(-1)])]
    s = P.PortSettings(str(tmp_path), pythonpath=repo)

    async def oracle_all():
        out = []
        for lo in range(0, len(calls), 6):
            out += await asyncio.gather(*[P.run_electron(s, fn, list(a), {}, {"dispatch_id": "o", "node_id": lo + j})
                                          for j, (fn, a) in enumerate(calls[lo:lo + 6])])
        return out

    async def gpu_all():
        return await asyncio.gather(*[executor.run(fn, list(a), {}, {"dispatch_id": "d", "node_id": i})
                                      for i, (fn, a) in enumerate(calls)])

    want = asyncio.run(oracle_all())
    got = asyncio.run(gpu_all())
    for (fn, _), w, g in zip(calls, want, got):
        assert type(w) is type(g), fn
        if isinstance(w, np.ndarray):
            assert w.dtype == g.dtype and w.shape == g.shape
            assert w.flags.c_contiguous == g.flags.c_contiguous and w.flags.f_contiguous == g.flags.f_contiguous
            assert w.tobytes(order="A") == g.tobytes(order="A"), fn
        elif isinstance(w, np.floating):
            assert abs(float(w) - float(g)) <= 1e-6 * abs(float(w)), fn          # tolerance stated in north_star
        else:
            assert w == g, fn


def test_executor_cancel_on_the_gpu(executor, gpu_count, tmp_path):
    """SSHExecutor.cancel through the whole stack on cuda:0: tombstoned electrons raise
    TaskCancelledError, their batch-mates return exact results."""
    if gpu_count < 1:
        pytest.skip("no CUDA device on this box")
    from covalent_ssh_plugin_b200 import electrons
    from covalent_ssh_plugin_b200.ssh import TaskCancelledError

    ex = executor
    xs = [np.random.default_rng(i).standard_normal(16384, dtype=np.float32) for i in range(200)]

    async def go():
        await ex.run(electrons.add_int, [1, 1], {}, {"dispatch_id": "warm", "node_id": 0})    # opens the engine
        tasks = [asyncio.ensure_future(ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": "gc", "node_id": i}))
                 for i in range(200)]
        await asyncio.sleep(0)
        for i in range(0, 200, 7):
            assert await ex.cancel({"dispatch_id": "gc", "node_id": i})
        return await asyncio.gather(*tasks, return_exceptions=True)

    out = asyncio.run(go())
    for i in range(200):
        if i % 7 == 0:
            assert isinstance(out[i], TaskCancelledError), i
        else:
            assert np.array_equal(out[i], np.float32(2.0) * xs[i] + np.float32(1.0)), i
