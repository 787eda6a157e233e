"""The mini lattice driver on CPU (numpy test double of the C ABI): the
reference's functional-test workflows run unchanged, and graphs of registered
electrons compile to one device DAG whose results equal plain in-process
execution of the same Python bodies."""

import numpy as np
import pytest

from covalent_ssh_plugin_b200 import _b2lib, electrons as E, engine as engine_mod, lattice as ct
from covalent_ssh_plugin import SSHExecutor
from tests.fake_b2lib import FakeLib

SMALL = dict(arena_bytes=96 << 20, staging_bytes=16 << 20, n_segments=4, max_batch=64, queue_cap=4096)


@pytest.fixture()
def gpu(tmp_path):
    engine_mod.shutdown_all()
    saved = _b2lib._lib
    _b2lib._lib = FakeLib(n_devices=4)
    for d in range(4):
        engine_mod.get_engine(d, **SMALL)
    key = tmp_path / "k"
    key.touch()
    mk = lambda host: SSHExecutor(username="u", hostname=host, ssh_key_file=str(key), cache_dir=str(tmp_path))  # noqa: E731
    yield mk
    engine_mod.shutdown_all()
    _b2lib._lib = saved


def test_basic_workflow_like_the_reference():
    """basic_workflow_test.py:7-29 with local electrons (no server needed)."""
    @ct.electron
    def join_words(a, b):
        return ", ".join([a, b])

    @ct.electron
    def excitement(a):
        return f"{a}!"

    @ct.lattice
    def basic_workflow(a, b):
        return excitement(join_words(a, b))

    result = ct.get_result(ct.dispatch(basic_workflow)("Hello", "World"), wait=True)
    assert str(result.status) == str(ct.status.COMPLETED) and result.result == "Hello, World!"


def test_failing_electron_fails_the_lattice(gpu):
    """basic_workflow_test.py:32-49: an electron that raises -> FAILED."""
    @ct.electron(executor=gpu("cuda:0"))
    def join_words(a, b):
        raise Exception(f"{', '.join([a, b])} -- but something went wrong!")

    @ct.lattice
    def wf(a, b):
        return join_words(a, b)

    result = ct.get_result(ct.dispatch(wf)("Hello", "World"))
    assert result.status == ct.status.FAILED and isinstance(result.error, RuntimeError)


def test_int_chain_goes_through_the_device(gpu):
    """BASELINE config 1: add(3,4) then add(prev,5)."""
    add = ct.electron(E.add_int, executor=gpu("cuda:0"))

    @ct.lattice
    def wf(a, b):
        return add(add(a, b), 5)

    r = ct.get_result(ct.dispatch(wf)(3, 4))
    assert r.status == ct.status.COMPLETED and r.result == 12 and type(r.result) is int
    assert r.stats["device_electrons"] == 2 and r.stats["host_electrons"] == 0


def _diamond(ex, n_fan, n, seed=0):
    root = ct.electron(E.plus_one, executor=ex(0))
    scale = [ct.electron(E.scale_add if i % 2 else E.plus_one, executor=ex(i)) for i in range(n_fan)]
    mul = [ct.electron(E.vec_mul, executor=ex(i)) for i in range(n_fan)]
    reduce_ = ct.electron(E.stack_sum, executor=ex(0))
    rng = np.random.default_rng(seed)
    x0 = rng.random(n, dtype=np.float32)
    ws = [rng.random(n, dtype=np.float32) for _ in range(n_fan)]

    @ct.lattice
    def wf(x, weights):
        r = root(x)
        return reduce_([mul[i](scale[i](r), weights[i]) for i in range(n_fan)])

    def host(x, weights):
        r = E.plus_one(x)
        return E.stack_sum([E.vec_mul((E.scale_add if i % 2 else E.plus_one)(r), weights[i]) for i in range(n_fan)])

    return wf, host, x0, ws


def test_diamond_compiles_to_one_device_dag(gpu):
    ex0 = gpu("cuda:0")
    wf, host, x0, ws = _diamond(lambda i: ex0, 24, 5000)
    r = ct.get_result(ct.dispatch(wf)(x0, ws))
    assert r.status == ct.status.COMPLETED, r.error
    assert np.array_equal(r.result, host(x0, ws))                     # bit-exact vs the Python bodies
    assert r.stats["device_batches"] == 1 and r.stats["device_electrons"] == 2 * 24 + 2
    assert r.stats["host_electrons"] == 0 and "demoted" not in r.stats
    assert engine_mod.existing_engine(0).heap.bytes_in_use() == 0     # every intermediate was freed


def test_wide_fan_in_is_split_by_column(gpu, monkeypatch):
    monkeypatch.setattr(ct, "SPLIT_MIN_BYTES", 1 << 16)
    monkeypatch.setattr(ct, "SLICE_ELEMS", 1024)
    ex0 = gpu("cuda:0")
    wf, host, x0, ws = _diamond(lambda i: ex0, 16, 5000, seed=3)
    r = ct.get_result(ct.dispatch(wf)(x0, ws))
    assert r.status == ct.status.COMPLETED, r.error
    assert np.array_equal(r.result, host(x0, ws))
    assert engine_mod.existing_engine(0).counters["electrons"] > 2 * 16 + 2   # gate + slices + join


def test_diamond_across_two_gpus_runs_level_by_level(gpu):
    exs = [gpu("cuda:0"), gpu("cuda:1")]
    wf, host, x0, ws = _diamond(lambda i: exs[i % 2], 10, 3000, seed=5)
    r = ct.get_result(ct.dispatch(wf)(x0, ws))
    assert r.status == ct.status.COMPLETED, r.error
    assert np.array_equal(r.result, host(x0, ws))
    # root (level 0), scale (1), mul (2), row copies from the other GPU (3), reduce (4)
    assert "demoted" not in r.stats and r.stats["device_batches"] >= 5


def test_mixed_host_and_device_electrons(gpu):
    ex0 = gpu("cuda:0")
    dev_scale = ct.electron(E.scale_add, executor=ex0)
    dev_sum = ct.electron(E.vec_sum, executor=ex0)

    @ct.electron
    def make(n):
        return np.arange(n, dtype=np.float32)

    @ct.electron
    def describe(total, arr):
        return f"{float(total):.1f}/{arr.shape[0]}"

    @ct.lattice
    def wf(n):
        a = dev_scale(make(n))
        return describe(dev_sum(a), a)

    r = ct.get_result(ct.dispatch(wf)(100))
    assert r.status == ct.status.COMPLETED, r.error
    assert r.result == f"{float((2 * np.arange(100, dtype=np.float32) + 1).sum()):.1f}/100"
    assert r.stats["host_electrons"] == 2 and r.stats["device_electrons"] == 2


def test_overflow_in_a_dag_is_demoted_to_the_executor_decision(gpu):
    big = (1 << 63) - 1
    ex = gpu("cuda:0")
    add = ct.electron(E.add_int, executor=ex)

    @ct.lattice
    def wf(a):
        return add(a, 1)

    r = ct.get_result(ct.dispatch(wf)(big))
    assert r.status == ct.status.FAILED and isinstance(r.error, RuntimeError)
    ex.run_local_on_ssh_fail = True
    r = ct.get_result(ct.dispatch(wf)(big))
    assert r.status == ct.status.COMPLETED and r.result == big + 1


def test_bulk_upload_and_download_use_unpack_and_pack(gpu):
    """Lattice inputs cross PCIe as one copy per staging segment + one unpack launch (K2);
    a lattice that returns many arrays gets them back through one pack launch + one D2H (K1)."""
    ex0 = gpu("cuda:0")
    scale = ct.electron(E.scale_add, executor=ex0)
    mul = ct.electron(E.vec_mul, executor=ex0)
    rng = np.random.default_rng(8)
    xs = [rng.random(700 + 13 * i, dtype=np.float32) for i in range(30)]
    ws = [rng.random(700 + 13 * i, dtype=np.float32) for i in range(30)]

    @ct.lattice
    def wf(a, b):
        return [mul(scale(a[i]), b[i]) for i in range(30)]

    eng = engine_mod.existing_engine(0)
    before = eng.stats().launches_pack
    r = ct.get_result(ct.dispatch(wf)(xs, ws))
    assert r.status == ct.status.COMPLETED, r.error
    for i in range(30):
        assert np.array_equal(r.result[i], E.vec_mul(E.scale_add(xs[i]), ws[i]))
    assert eng.stats().launches_pack - before == 2          # one unpack (60 inputs) + one pack (30 outputs)
    assert eng.heap.bytes_in_use() == 0 and sorted(eng._free_out) == [0, 1, 2, 3]


def test_put_many_get_many_chunking(gpu):
    """More bytes than one staging segment: several chunks, same values back."""
    gpu("cuda:0")
    eng = engine_mod.existing_engine(0)
    rng = np.random.default_rng(9)
    arrs = [rng.random((300 + i, 500), dtype=np.float32) for i in range(12)]      # ~7.3 MiB > 4 MiB segments
    arrs.append(np.asfortranarray(rng.random((40, 30), dtype=np.float32)))
    arrs.append(np.zeros(0, dtype=np.float32))
    das = eng.put_many(arrs)
    back = eng.get_many(das)
    for a, b, da in zip(arrs, back, das):
        assert b.shape == a.shape and b.flags.f_contiguous == a.flags.f_contiguous and np.array_equal(a, b)
        da.free()
    assert eng.heap.bytes_in_use() == 0


def test_diamond_across_four_gpus(gpu):
    """BASELINE config 4 shape: fan-out placed round-robin on every GPU, reducer on GPU 0 gathers
    the rows of the other three over (modelled) peer arenas."""
    exs = [gpu(f"cuda:{d}") for d in range(4)]
    wf, host, x0, ws = _diamond(lambda i: exs[i % 4], 21, 2500, seed=11)
    r = ct.get_result(ct.dispatch(wf)(x0, ws))
    assert r.status == ct.status.COMPLETED, r.error
    assert np.array_equal(r.result, host(x0, ws))
    assert "demoted" not in r.stats and r.stats["device_electrons"] == 2 * 21 + 2
    for d in range(4):
        assert engine_mod.existing_engine(d).heap.bytes_in_use() == 0


@pytest.fixture()
def ssh_alias(gpu, tmp_path, monkeypatch):
    """``executor="ssh"`` resolves through the user's covalent.conf, as Covalent's executor manager does."""
    from covalent_ssh_plugin_b200 import _covalent_compat, hostpool

    key = tmp_path / "k"
    (tmp_path / "covalent.conf").write_text(
        '[dispatcher]\ncache_dir = "%s"\n[executors.ssh]\nusername = "u"\nhostname = "cuda:0"\n'
        'ssh_key_file = "%s"\nhost_pool = true\n' % (tmp_path, key))
    monkeypatch.setenv("COVALENT_CONFIG_DIR", str(tmp_path))
    monkeypatch.setattr(_covalent_compat, "_loaded", None)
    yield
    monkeypatch.setattr(_covalent_compat, "_loaded", None)
    hostpool.shutdown()


def test_reference_basic_workflow_with_the_ssh_alias(ssh_alias):
    """basic_workflow_test.py:7-29, body unchanged: ``executor="ssh"`` on an electron that is plain Python
    (not a registered device electron) runs in the host pool."""
    @ct.electron(executor="ssh")
    def join_words(a, b):
        return ", ".join([a, b])

    @ct.electron
    def excitement(a):
        return f"{a}!"

    @ct.lattice
    def basic_workflow(a, b):
        phrase = join_words(a, b)
        return excitement(phrase)

    dispatch_id = ct.dispatch(basic_workflow)("Hello", "World")
    result = ct.get_result(dispatch_id=dispatch_id, wait=True)
    assert str(result.status) == str(ct.status.COMPLETED), result.error
    assert result.result == "Hello, World!"


def test_reference_svm_workflow_unchanged(ssh_alias):
    """svm_workflow.py:1-46, body unchanged (DepsPip included): the SVC is trained by the ssh executor's
    host pool and travels back as a pickled object."""
    datasets = pytest.importorskip("sklearn.datasets")
    svm = pytest.importorskip("sklearn.svm")
    from numpy.random import permutation

    deps_pip = ct.DepsPip(packages=["numpy==1.23.2", "scikit-learn==1.1.2"])

    @ct.electron
    def load_data():
        iris = datasets.load_iris()
        perm = permutation(iris.target.size)
        iris.data = iris.data[perm]
        iris.target = iris.target[perm]
        return iris.data, iris.target

    @ct.electron(executor="ssh", deps_pip=deps_pip)
    def train_svm(data, C, gamma):
        X, y = data
        clf = svm.SVC(C=C, gamma=gamma)
        clf.fit(X[90:], y[90:])
        return clf

    @ct.electron
    def score_svm(data, clf):
        X_test, y_test = data
        return clf.score(X_test[:90], y_test[:90])

    @ct.lattice
    def run_experiment(C=1.0, gamma=0.7):
        data = load_data()
        clf = train_svm(data=data, C=C, gamma=gamma)
        score = score_svm(data=data, clf=clf)
        return score

    dispatchable_func = ct.dispatch(run_experiment)
    dispatch_id = dispatchable_func(C=1.0, gamma=0.7)
    result = ct.get_result(dispatch_id=dispatch_id, wait=True)
    assert str(result.status) == str(ct.status.COMPLETED), result.error
    assert 0.5 < result.result <= 1.0


def test_readme_usage_with_an_executor_object(gpu, tmp_path):
    """README.md:52-65 of the reference: ``ct.executor.SSHExecutor(...)`` handed to an electron."""
    key = tmp_path / "k"
    executor = ct.executor.SSHExecutor(username="user", hostname="cuda:1", ssh_key_file=str(key),
                                       cache_dir=str(tmp_path))
    assert isinstance(executor, SSHExecutor)

    @ct.electron(executor=executor)
    def my_custom_task(x, y):
        return x + y

    @ct.lattice
    def wf(x, y):
        return my_custom_task(x, y)

    r = ct.get_result(ct.dispatch(wf)(3, 4))
    # a plain-Python electron on a GPU "host" without the host pool: the reference's failure decision
    assert r.status == ct.status.FAILED and isinstance(r.error, RuntimeError)

    @ct.lattice
    def wf2(x, y):
        return ct.electron(E.add_int, executor=executor)(x, y)

    r = ct.get_result(ct.dispatch(wf2)(3, 4))
    assert r.status == ct.status.COMPLETED and r.result == 7 and type(r.result) is int
