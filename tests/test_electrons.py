"""Host-side encoding rules of registered electrons (electrons.py)."""

import functools

import numpy as np
import pytest

from covalent_ssh_plugin_b200 import electrons as E


def test_lookup_sees_through_partial_and_wrapped():
    assert E.lookup(E.scale_add).kind == E.K_SCALE_ADD
    assert E.lookup(functools.partial(E.scale_add)) is E.lookup(E.scale_add)

    @functools.wraps(E.vec_add)
    def wrapper(*a, **k):
        return E.vec_add(*a, **k)

    assert E.lookup(wrapper) is E.lookup(E.vec_add)
    assert E.lookup(len) is None
    assert E.lookup(functools.partial(E.vec_add, 1)) is None      # bound args change the call


def test_registration_is_idempotent_and_validates():
    a = E.device_electron("scale_add", alpha=3.0, beta=0.5, name="t.same")(lambda x: x)
    b = E.device_electron("scale_add", alpha=3.0, beta=0.5, name="t.same")(lambda x: x)
    assert E.lookup(a) is E.lookup(b)
    with pytest.raises(ValueError):
        E.device_electron("scale_add", alpha=0.1)          # 0.1 is not an fp32 value
    with pytest.raises(ValueError):
        E.device_electron("iscale_add", ialpha=1 << 70)
    with pytest.raises(KeyError):
        E.device_electron("no_such_kind")


def test_encode_int_rules():
    t = E.encode(E.add_int, [3, 4], {})
    assert t.imm == (3, 4) and t.payloads == () and t.result_meta == ("int",) and t.out_bytes == 8
    assert E.encode(E.add_int, [], {"a": 1, "b": 2}).imm == (1, 2)
    for bad in ([1 << 63, 0], [True, 1], [1.0, 2], [np.int64(1), 2], [1], [1, 2, 3]):
        with pytest.raises(E.NotDeviceEligible):
            E.encode(E.add_int, bad, {})


def test_encode_array_rules():
    x = np.arange(12, dtype=np.float32).reshape(3, 4)
    t = E.encode(E.scale_add, [x], {})
    assert t.count == 12 and t.out_bytes == 48 and t.result_meta == ("array", np.float32, (3, 4), "C")
    f = np.asfortranarray(x)
    tf = E.encode(E.scale_add, [f], {})
    assert tf.result_meta[3] == "F" and bytes(tf.payloads[0]) == f.tobytes(order="F")
    v = x[::2, ::2]
    tv = E.encode(E.scale_add, [v], {})
    assert tv.result_meta[2] == (2, 2) and bytes(tv.payloads[0]) == np.ascontiguousarray(v).tobytes()
    for bad in (x.astype(np.float64), [1.0, 2.0], x.tolist(), np.float32(1.0)):
        with pytest.raises(E.NotDeviceEligible):
            E.encode(E.scale_add, [bad], {})
    with pytest.raises(E.NotDeviceEligible):
        E.encode(E.vec_add, [x, x[:2]], {})
    with pytest.raises(E.NotDeviceEligible):
        E.encode(E.vec_add, [x, f], {})
    with pytest.raises(E.NotDeviceEligible):
        E.encode(lambda z: z, [x], {})


def test_decode_types_and_no_aliasing():
    raw = np.arange(64, dtype=np.uint8)
    a = E.decode(("array", np.float32, (2, 4), "C"), raw)
    assert a.shape == (2, 4) and a.flags.c_contiguous and a.tobytes() == raw[:32].tobytes()
    raw[:] = 0
    assert a.tobytes() != raw[:32].tobytes()
    f = E.decode(("array", np.float32, (2, 4), "F"), np.arange(64, dtype=np.uint8))
    assert f.flags.f_contiguous and f.tobytes(order="F") == bytes(range(32))
    s = E.decode(("f32",), np.frombuffer(np.float32(1.5).tobytes() + bytes(12), dtype=np.uint8))
    assert type(s) is np.float32 and s == 1.5
    i = E.decode(("int",), np.frombuffer(np.int64(-7).tobytes() + bytes(8), dtype=np.uint8))
    assert type(i) is int and i == -7


def test_stack_sum_encoding():
    xs = [np.full(5, r, dtype=np.float32) for r in range(3)]
    t = E.encode(E.stack_sum, [xs], {})
    assert t.nrows == 3 and t.count == 5 and len(t.payloads) == 3 and t.out_bytes == 20
    t2 = E.encode(E.stack_sum, [np.stack(xs)], {})
    assert t2.nrows == 3 and t2.count == 5 and len(t2.payloads) == 1
    with pytest.raises(E.NotDeviceEligible):
        E.encode(E.stack_sum, [[xs[0], xs[1][:3]]], {})
