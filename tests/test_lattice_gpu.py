"""BASELINE config 4 on the GPU: a 1 000-node diamond (root -> 998 fan-out ->
reduce) with 1 MiB fp32 operands, compiled to one device DAG (dependencies
released on the device through the MPMC queue) and compared bit for bit with
in-process execution of the same electron bodies — which is what the reference
would compute electron by electron (exec.py:37)."""

import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mk_executor(gpu_count, tmp_path_factory):
    if gpu_count < 1:
        pytest.skip("no CUDA device on this box")
    from covalent_ssh_plugin_b200 import SSHExecutor
    from covalent_ssh_plugin_b200.engine import get_engine, shutdown_all

    shutdown_all()
    for d in range(min(gpu_count, 8)):
        get_engine(d, arena_bytes=6 << 30, staging_bytes=256 << 20, n_segments=4, max_batch=1024)
    tmp = tmp_path_factory.mktemp("lat")
    key = tmp / "id_stub"
    key.touch()
    yield lambda host: SSHExecutor(username="u", hostname=host, ssh_key_file=str(key), cache_dir=str(tmp))
    shutdown_all()


def diamond(ct, E, ex_of, n_fan, n, seed):
    root = ct.electron(E.plus_one, executor=ex_of(0))
    mul = [ct.electron(E.vec_mul, executor=ex_of(i)) for i in range(n_fan)]
    reduce_ = ct.electron(E.stack_sum, executor=ex_of(0))
    rng = np.random.default_rng(seed)
    x0 = rng.random(n, dtype=np.float32)
    ws = [rng.random(n, dtype=np.float32) for _ in range(n_fan)]

    @ct.lattice
    def wf(x, weights):
        r = root(x)
        return reduce_([mul[i](r, weights[i]) for i in range(n_fan)])

    def host(x, weights):
        r = E.plus_one(x)
        return E.stack_sum([E.vec_mul(r, w) for w in weights])

    return wf, host, x0, ws


def test_diamond_1000_nodes_1mib_one_gpu(mk_executor):
    from covalent_ssh_plugin_b200 import electrons as E, lattice as ct
    from covalent_ssh_plugin_b200.engine import existing_engine

    ex = mk_executor("cuda:0")
    wf, host, x0, ws = diamond(ct, E, lambda i: ex, 998, 262144, seed=4)
    t0 = time.perf_counter()
    r = ct.get_result(ct.dispatch(wf)(x0, ws))
    dt = time.perf_counter() - t0
    assert r.status == ct.status.COMPLETED, r.error
    want = host(x0, ws)
    assert r.result.dtype == want.dtype and r.result.shape == want.shape
    assert np.array_equal(r.result, want)
    assert r.stats["device_batches"] == 1 and r.stats["device_electrons"] == 1000
    assert r.stats["host_electrons"] == 0 and "demoted" not in r.stats
    assert existing_engine(0).heap.bytes_in_use() == 0
    print(f"\n1000-node diamond, 1 MiB operands, one GPU: {dt * 1e3:.1f} ms end to end "
          f"(998 MiB of inputs uploaded, one DAG launch of {r.stats['device_tasks']} task records: "
          f"{r.stats['dag_seconds'] * 1e3:.2f} ms incl. table build and wait)")


def test_int_chain_and_mixed_graph(mk_executor):
    from covalent_ssh_plugin_b200 import electrons as E, lattice as ct

    ex = mk_executor("cuda:0")
    add = ct.electron(E.add_int, executor=ex)
    scale = ct.electron(E.scale_add, executor=ex)
    total = ct.electron(E.vec_sum, executor=ex)

    @ct.electron
    def make(n):
        return np.linspace(0, 1, n, dtype=np.float32)

    @ct.lattice
    def wf(a, b, n):
        return add(add(a, b), 5), total(scale(scale(make(n))))

    r = ct.get_result(ct.dispatch(wf)(3, 4, 100000))
    assert r.status == ct.status.COMPLETED, r.error
    x = np.linspace(0, 1, 100000, dtype=np.float32)
    ref = E.vec_sum(E.scale_add(E.scale_add(x)))
    assert r.result[0] == 12 and abs(float(r.result[1]) - float(ref)) <= 1e-6 * abs(float(ref))


def test_diamond_across_two_gpus(mk_executor, gpu_count):
    """Dependency edges across GPUs: odd fan-out electrons run on GPU 1 and read the
    root's output out of GPU 0's arena; the reducer gathers GPU 1's rows over NVLink."""
    if gpu_count < 2:
        pytest.skip("needs two GPUs")
    from covalent_ssh_plugin_b200 import electrons as E, lattice as ct

    exs = [mk_executor("cuda:0"), mk_executor("cuda:1")]
    wf, host, x0, ws = diamond(ct, E, lambda i: exs[i % 2], 200, 262144, seed=6)
    r = ct.get_result(ct.dispatch(wf)(x0, ws))
    assert r.status == ct.status.COMPLETED, r.error
    assert np.array_equal(r.result, host(x0, ws))
    assert "demoted" not in r.stats


def test_diamond_1000_nodes_on_every_gpu(mk_executor, gpu_count):
    """BASELINE config 4 as written: 1 000 nodes, 1 MiB operands, fan-out placed round-robin over
    all GPUs of the box (7/8 of the edges cross GPUs at 8), reducer gathers the rows over NVLink."""
    if gpu_count < 4:
        pytest.skip("needs at least four GPUs")
    from covalent_ssh_plugin_b200 import electrons as E, lattice as ct

    g = min(gpu_count, 8)
    exs = [mk_executor(f"cuda:{d}") for d in range(g)]
    wf, host, x0, ws = diamond(ct, E, lambda i: exs[i % g], 998, 262144, seed=8)
    t0 = time.perf_counter()
    r = ct.get_result(ct.dispatch(wf)(x0, ws))
    dt = time.perf_counter() - t0
    assert r.status == ct.status.COMPLETED, r.error
    assert np.array_equal(r.result, host(x0, ws))
    assert "demoted" not in r.stats and r.stats["device_electrons"] == 1000
    print(f"\n1000-node diamond over {g} GPUs: {dt * 1e3:.1f} ms end to end, {r.stats}")
