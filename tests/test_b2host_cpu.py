"""The CPython staging helper (hostext/b2host.c -> _b2host.so) on its own: what it accepts, what it
leaves to the general Python path, what it writes.  CPU only — the helper never touches CUDA."""

import numpy as np
import pytest

from covalent_ssh_plugin_b200 import _b2lib as L
from covalent_ssh_plugin_b200.engine import IDX_BITS

MASK = (1 << IDX_BITS) - 1
_b2host = None


@pytest.fixture(scope="module", autouse=True)
def helper(built_library):
    """The helper is imported only after the session fixture has built it (a fresh clone has no .so)."""
    global _b2host
    from covalent_ssh_plugin_b200.engine import _load_b2host

    _b2host = _load_b2host()
    return _b2host


@pytest.fixture()
def stager():
    st = _b2host.Stager(np.ndarray)
    jobs = np.zeros(16, dtype=L.JOB_DTYPE)
    tasks = np.zeros(8, dtype=L.TASK_DTYPE)
    st.begin(jobs.ctypes.data, len(jobs), tasks.ctypes.data, len(tasks), 1 << 20, 1 << 20, 5 << 20, 1 << 20)
    return st, jobs, tasks


def test_record_sizes_match_the_header():
    assert _b2host.TASK_BYTES == 64 and _b2host.JOB_BYTES == 24 and _b2host.IDX_BITS == IDX_BITS
    assert np.dtype(L.TASK_DTYPE).itemsize == 64 and np.dtype(L.JOB_DTYPE).itemsize == 24


def test_stage1_writes_the_task_and_the_copy_job(stager):
    st, jobs, tasks = stager
    x = np.arange(1000, dtype=np.float32)
    r = st.stage1(7, L.K_SCALE_ADD, x)
    assert r >= 0 and (r & MASK) == 0 and (r >> IDX_BITS) == 0
    assert st.state() == (1, 1, 4096, 4096)                       # 4000 B rounded up to 256
    assert tuple(jobs[0]) == (x.ctypes.data, 1 << 20, 4000)
    t = tasks[0]
    assert (t["fn_id"], t["flags"], t["arg0"], t["out"], t["count"]) == (7, 0, 1 << 20, 5 << 20, 1000)
    y = np.ones((3, 5), dtype=np.float32)
    r = st.stage1(7, L.K_COPY, y)                                  # COPY counts bytes
    assert (r & MASK) == 1 and (r >> IDX_BITS) == 4096 and tasks[1]["count"] == 60
    r = st.stage1(9, L.K_SUM, x)                                   # SUM: 4 result bytes in a 256 B slot
    assert (r >> IDX_BITS) == 4096 + 256 and st.state()[3] == 4096 + 256 + 256


@pytest.mark.parametrize("bad", [
    np.arange(10, dtype=np.float64), np.arange(10, dtype=np.float32)[::2], np.asfortranarray(np.ones((3, 4), np.float32)),
    np.float32(1.0), np.array(1.0, dtype=np.float32), [1.0, 2.0], None, np.arange(10, dtype=np.float32).view(np.ndarray)[::-1],
    np.ones(3, dtype=">f4"),
])
def test_anything_but_a_plain_c_contiguous_float32_ndarray_goes_to_the_general_path(stager, bad):
    st, _, _ = stager
    assert st.stage1(1, L.K_SCALE_ADD, bad) == _b2host.GENERAL
    assert st.state() == (0, 0, 0, 0)                              # nothing was reserved


def test_ndarray_subclasses_and_large_electrons_are_general(stager):
    st, _, _ = stager

    class Sub(np.ndarray):
        pass

    assert st.stage1(1, L.K_SCALE_ADD, np.ones(4, np.float32).view(Sub)) == _b2host.GENERAL
    assert st.stage1(1, L.K_SCALE_ADD, np.ones((2 << 20) // 4, np.float32)) == _b2host.GENERAL   # sliced by the engine
    assert st.stage1(1, L.K_STACK_SUM, np.ones(4, np.float32)) == _b2host.GENERAL                # not a one-array kind


def test_empty_arrays_take_a_slot_but_no_copy_job(stager):
    st, jobs, tasks = stager
    r = st.stage1(3, L.K_SCALE_ADD, np.zeros(0, dtype=np.float32))
    assert r >= 0 and st.state() == (1, 0, 0, 256) and tasks[0]["count"] == 0


def test_stage2_needs_identical_shapes(stager):
    st, jobs, tasks = stager
    a, b = np.ones((4, 8), np.float32), np.full((4, 8), 2, np.float32)
    r = st.stage2(5, a, b)
    assert r >= 0 and st.state() == (1, 2, 512, 256)
    assert tuple(jobs[0]) == (a.ctypes.data, 1 << 20, 128) and tuple(jobs[1]) == (b.ctypes.data, (1 << 20) + 256, 128)
    assert (tasks[0]["arg0"], tasks[0]["arg1"], tasks[0]["count"]) == (1 << 20, (1 << 20) + 256, 32)
    assert st.stage2(5, a, b.reshape(8, 4)) == _b2host.GENERAL      # same bytes, different shape
    assert st.stage2(5, a, b[:2]) == _b2host.GENERAL
    assert st.stage2(5, a, b.astype(np.float64)) == _b2host.GENERAL


def test_stage_int_takes_exact_python_ints_within_int64(stager):
    st, _, tasks = stager
    r = st.stage_int(2, -5, (1 << 63) - 1)
    assert r >= 0 and tasks[0]["flags"] == (L.F_IMM0 | L.F_IMM1)
    assert int(np.int64(np.uint64(tasks[0]["arg0"]))) == -5 and int(tasks[0]["arg1"]) == (1 << 63) - 1
    for bad in ((1 << 63, 1), (1, -(1 << 63) - 1), (True, 1), (np.int64(3), 1), (1.0, 2)):
        assert st.stage_int(2, *bad) == _b2host.GENERAL


def test_full_is_reported_for_every_capacity(stager):
    st, jobs, tasks = stager
    x = np.ones(16, np.float32)
    for _ in range(8):
        assert st.stage1(1, L.K_SCALE_ADD, x) >= 0
    assert st.stage1(1, L.K_SCALE_ADD, x) == _b2host.FULL            # task capacity (8)
    st.begin(jobs.ctypes.data, 16, tasks.ctypes.data, 8, 0, 512, 0, 1 << 20)
    assert st.stage1(1, L.K_SCALE_ADD, np.ones(64, np.float32)) >= 0
    assert st.stage1(1, L.K_SCALE_ADD, np.ones(65, np.float32)) == _b2host.FULL     # ingress bytes (512)
    st.begin(jobs.ctypes.data, 16, tasks.ctypes.data, 8, 0, 1 << 20, 0, 256)
    assert st.stage_int(1, 1, 2) >= 0 and st.stage_int(1, 1, 2) == _b2host.FULL       # egress bytes (256)
    st.begin(jobs.ctypes.data, 1, tasks.ctypes.data, 8, 0, 1 << 20, 0, 1 << 20)
    assert st.stage2(1, x, x) == _b2host.FULL                         # two copy jobs do not fit one slot
    st.end()
    assert st.stage1(1, L.K_SCALE_ADD, x) == _b2host.FULL             # no open batch


def test_state_round_trip_and_bounds(stager):
    st, _, _ = stager
    st.set_state(3, 2, 1024, 2048)
    assert st.state() == (3, 2, 1024, 2048)
    with pytest.raises(ValueError):
        st.set_state(9, 0, 0, 0)                                       # beyond the task capacity
    with pytest.raises(ValueError):
        st.set_state(0, 0, (1 << 20) + 1, 0)
    assert st.add_job(1234, 5678, 9) == 0 and st.state()[1] == 3
    with pytest.raises(TypeError):
        st.stage1(1, 2)
