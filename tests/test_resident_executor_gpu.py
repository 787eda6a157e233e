"""The whole hot path — SSHExecutor.run -> engine -> b2_submit -> worker — with the RESIDENT worker
switched on (engine option ``resident=True`` / B2_RESIDENT=1): same golden vectors, same bytes."""

import asyncio
import ctypes as C

import numpy as np
import pytest

from tests.golden import cases
from tests.test_executor_gpu import DEVICE_CASES, check_against_golden

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]


@pytest.fixture(scope="module")
def resident_executor(gpu_count, tmp_path_factory):
    if gpu_count < 1:
        pytest.skip("no CUDA device on this box")
    from covalent_ssh_plugin_b200 import SSHExecutor
    from covalent_ssh_plugin_b200.engine import get_engine, shutdown_all

    shutdown_all()
    eng = get_engine(0, resident=True, arena_bytes=3 << 30, staging_bytes=512 << 20, n_segments=16, max_batch=512)
    assert eng.resident
    tmp = tmp_path_factory.mktemp("resexec")
    key = tmp / "id_stub"
    key.touch()
    yield SSHExecutor(username="user", hostname="cuda:0", ssh_key_file=str(key), cache_dir=str(tmp)), eng
    shutdown_all()


def test_golden_cases_through_the_resident_worker(resident_executor):
    ex, eng = resident_executor

    async def go():
        return await asyncio.gather(*[
            ex.run(cases.ELECTRONS[e], cases.build_args(r), {}, {"dispatch_id": "rg", "node_id": i})
            for i, (_, e, r) in enumerate(DEVICE_CASES)])

    before = eng.stats()
    for (cid, _, _), v in zip(DEVICE_CASES, asyncio.run(go())):
        check_against_golden(cid, v)
    after = eng.stats()
    assert after.resident_batches > before.resident_batches, "the batches did not go through the resident worker"
    assert after.resident_instances >= 1


def test_fan_out_of_5000_electrons_through_the_resident_worker(resident_executor):
    from covalent_ssh_plugin_b200 import electrons

    ex, eng = resident_executor
    xs = [np.random.default_rng(i).standard_normal(16384, dtype=np.float32) for i in range(5000)]

    async def go(tag):
        return await asyncio.gather(*[
            ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": tag, "node_id": i}) for i in range(len(xs))])

    before = eng.stats()
    for rep in range(3):
        out = asyncio.run(go(f"rf{rep}"))
        for i in range(0, len(xs), 97):
            assert np.array_equal(out[i], np.float32(2.0) * xs[i] + np.float32(1.0)), (rep, i)
        del out
    after = eng.stats()
    n_batches = after.resident_batches - before.resident_batches
    assert n_batches >= 3 * (5000 // 512)
    assert after.launches_worker - before.launches_worker == after.resident_instances - before.resident_instances, \
        "in resident mode the only worker launches are the resident instances themselves"
