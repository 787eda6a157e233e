"""Registered electrons: the host side of the on-device callback table.

The reference ships an arbitrary Python callable to the remote host by
cloudpickling ``(fn, args, kwargs)`` (/root/reference/covalent_ssh_plugin/
ssh.py:147-150) and runs it in a fresh interpreter (exec.py:29-46).  A GPU
cannot run arbitrary Python, so electrons that should execute on the B200 are
*registered*: ``@device_electron(kind, ...)`` binds a Python function (whose
body stays the host-side definition of the electron's semantics, and is what
the oracle executes) to one of the built-in device callbacks
(include/b2exec.h ``enum b2_kind``) plus its registered constants.

``encode`` turns one call ``fn(*args, **kwargs)`` into the binary task record
the worker kernel consumes — the replacement of the pickle at ssh.py:150 — and
``decode`` rebuilds the Python result object from the result bytes — the
replacement of ``pickle.load`` at ssh.py:455-456 — with the same type, dtype,
shape and memory order the reference round trip would have produced.

This module imports neither CUDA nor the C library: it is importable inside
the oracle's subprocesses.
"""

from __future__ import annotations

import functools
import inspect
import threading
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

# Mirror of enum b2_kind (include/b2exec.h); kept literal so that importing this
# module never loads the CUDA library.
K_NONE, K_COPY, K_SCALE_ADD, K_ADD, K_MUL, K_AXPY, K_SUM, K_STACK_SUM, K_IADD, K_IMUL, K_ISCALE_ADD = range(11)

_ARRAY1 = (K_COPY, K_SCALE_ADD, K_SUM)
_ARRAY2 = (K_ADD, K_MUL, K_AXPY)
_INT2 = (K_IADD, K_IMUL)
_INT1 = (K_ISCALE_ADD,)

KIND_NAMES = {
    "copy": K_COPY, "identity": K_COPY, "scale_add": K_SCALE_ADD, "add": K_ADD, "mul": K_MUL,
    "axpy": K_AXPY, "sum": K_SUM, "stack_sum": K_STACK_SUM, "iadd": K_IADD, "imul": K_IMUL,
    "iscale_add": K_ISCALE_ADD,
}

ATTR = "__b200_electron__"
INT64_MIN, INT64_MAX = -(1 << 63), (1 << 63) - 1
MAX_FN = 1024


class NotDeviceEligible(Exception):
    """This call cannot run on the device (unregistered fn or unsupported args)."""


@dataclass(frozen=True)
class DeviceElectron:
    """One callback-table entry: what ``b2_register`` receives."""

    fn_id: int
    name: str
    kind: int
    alpha: float = 0.0
    beta: float = 0.0
    ialpha: int = 0
    ibeta: int = 0


class EncodedTask:
    """One electron in wire form, before arena offsets are assigned."""

    __slots__ = ("spec", "payloads", "imm", "count", "nrows", "out_bytes", "result_meta")

    def __init__(self, spec, payloads, imm, count, nrows, out_bytes, result_meta):
        self.spec = spec                  # DeviceElectron
        self.payloads = payloads          # contiguous uint8 views (or DeviceRef) to stage, in operand order
        self.imm = imm                    # immediates (int electrons)
        self.count = count                # elements per operand (bytes for K_COPY)
        self.nrows = nrows
        self.out_bytes = out_bytes
        self.result_meta = result_meta    # ("array", dtype, shape, order) | ("f32",) | ("int",)


_lock = threading.Lock()
_registry: Dict[int, DeviceElectron] = {}     # fn_id -> spec
_next_fn_id = 1                               # 0 is the gate electron


def registry_snapshot() -> List[DeviceElectron]:
    with _lock:
        return [_registry[k] for k in sorted(_registry)]


def registry_size() -> int:
    return len(_registry)


def _new_spec(name: str, kind: int, alpha, beta, ialpha, ibeta) -> DeviceElectron:
    global _next_fn_id
    with _lock:
        # identical registrations share a table entry (re-decorating in tests/REPLs)
        for spec in _registry.values():
            if (spec.name, spec.kind, spec.alpha, spec.beta, spec.ialpha, spec.ibeta) == (
                name, kind, float(alpha), float(beta), int(ialpha), int(ibeta)):
                return spec
        if _next_fn_id >= MAX_FN:
            raise RuntimeError(f"callback table full ({MAX_FN} entries)")
        spec = DeviceElectron(_next_fn_id, name, kind, float(alpha), float(beta), int(ialpha), int(ibeta))
        _registry[spec.fn_id] = spec
        _next_fn_id += 1
        return spec


def device_electron(kind, *, alpha: float = 0.0, beta: float = 0.0, ialpha: int = 0,
                    ibeta: int = 0, name: Optional[str] = None) -> Callable[[Callable], Callable]:
    """Decorator: register ``fn`` as a device electron of the given kind.

    The function object itself is returned (not a wrapper), tagged with its
    callback-table entry, so it still pickles, still runs on the host, and
    ``@ct.electron`` can be stacked on top.
    """
    k = KIND_NAMES[kind] if isinstance(kind, str) else int(kind)
    if k <= K_NONE or k > K_ISCALE_ADD:
        raise ValueError(f"unknown device callback kind {kind!r}")
    if not (INT64_MIN <= ialpha <= INT64_MAX and INT64_MIN <= ibeta <= INT64_MAX):
        raise ValueError("ialpha/ibeta must fit int64")
    if k in (K_SCALE_ADD, K_AXPY):
        # the device computes in fp32; refuse constants fp32 cannot hold exactly
        for v in (alpha, beta):
            if float(np.float32(v)) != float(v):
                raise ValueError(f"constant {v!r} is not exactly representable in float32")

    def wrap(fn: Callable) -> Callable:
        spec = _new_spec(name or getattr(fn, "__qualname__", repr(fn)), k, alpha, beta, ialpha, ibeta)
        setattr(fn, ATTR, spec)
        return fn

    return wrap


def lookup(fn: Callable) -> Optional[DeviceElectron]:
    """Callback-table entry of ``fn`` (seeing through functools.partial with no
    bound arguments and through ``__wrapped__`` chains), else None."""
    for _ in range(8):
        spec = getattr(fn, ATTR, None)
        if spec is not None:
            return spec
        if isinstance(fn, functools.partial) and not fn.args and not fn.keywords:
            fn = fn.func
            continue
        nxt = getattr(fn, "__wrapped__", None)
        if nxt is None:
            return None
        fn = nxt
    return None


@functools.lru_cache(maxsize=4096)
def _signature(fn: Callable) -> inspect.Signature:
    return inspect.signature(fn)


def _positional(fn: Callable, args, kwargs) -> tuple:
    if not kwargs:
        return tuple(args)
    try:
        bound = _signature(fn).bind(*args, **kwargs)
    except TypeError as err:
        raise NotDeviceEligible(f"arguments do not bind: {err}") from None
    if bound.kwargs:
        raise NotDeviceEligible("keyword-only / **kwargs arguments are not device-eligible")
    return tuple(bound.args)


class DeviceRef:
    """An operand that already lives in a GPU arena (engine.DeviceArray): nothing is
    staged for it, the task record points at it."""

    __slots__ = ("ordinal", "offset", "size")

    def __init__(self, ordinal: int, offset: int, size: int):
        self.ordinal, self.offset, self.size = ordinal, offset, size


def _as_operand(x: Any) -> Tuple[Any, tuple, str]:
    """(flat uint8 view of the raw buffer — or a DeviceRef —, shape, order) of a
    float32 ndarray / device-resident array."""
    if getattr(x, "__b200_device_array__", False):
        if not x.alive:
            raise NotDeviceEligible("device array has been freed")
        return DeviceRef(x.ordinal, x.offset, x.nbytes), tuple(x.shape), x.order
    if type(x) is not np.ndarray or x.dtype != np.float32:
        raise NotDeviceEligible("array operands must be float32 numpy.ndarray")
    if x.ndim == 0:
        raise NotDeviceEligible("0-d arrays behave like numpy scalars on the host")
    if x.flags.c_contiguous:
        order = "C"
    elif x.flags.f_contiguous:
        order = "F"
    else:
        # cloudpickle ships non-contiguous views as contiguous copies too
        x = np.ascontiguousarray(x)
        order = "C"
    flat = x.reshape(-1, order="A" if order == "F" else "C")
    return flat.view(np.uint8), x.shape, order


def encode(fn: Callable, args, kwargs) -> EncodedTask:
    """Binary form of one call.  Raises NotDeviceEligible when the electron (or
    these particular arguments) must take the reference's failure path."""
    spec = lookup(fn)
    if spec is None:
        raise NotDeviceEligible(
            f"{getattr(fn, '__qualname__', fn)!r} is not a registered device electron")
    pos = _positional(fn, args, kwargs)
    k = spec.kind
    if k in _ARRAY1:
        if len(pos) != 1:
            raise NotDeviceEligible(f"{spec.name} takes one array operand, got {len(pos)}")
        raw, shape, order = _as_operand(pos[0])
        n = raw.size // 4
        if k == K_COPY:
            return EncodedTask(spec, (raw,), (0, 0), raw.size, 0, raw.size,
                               ("array", np.float32, shape, order))
        if k == K_SUM:
            return EncodedTask(spec, (raw,), (0, 0), n, 0, 4, ("f32",))
        return EncodedTask(spec, (raw,), (0, 0), n, 0, raw.size, ("array", np.float32, shape, order))
    if k in _ARRAY2:
        if len(pos) != 2:
            raise NotDeviceEligible(f"{spec.name} takes two array operands, got {len(pos)}")
        r0, s0, o0 = _as_operand(pos[0])
        r1, s1, o1 = _as_operand(pos[1])
        if s0 != s1 or o0 != o1:
            raise NotDeviceEligible("binary electrons need operands of identical shape and order")
        return EncodedTask(spec, (r0, r1), (0, 0), r0.size // 4, 0, r0.size,
                           ("array", np.float32, s0, o0))
    if k == K_STACK_SUM:
        if len(pos) != 1:
            raise NotDeviceEligible(f"{spec.name} takes one sequence of arrays")
        seq = pos[0]
        if type(seq) is np.ndarray:
            if seq.ndim < 2 or seq.dtype != np.float32 or not seq.flags.c_contiguous:
                raise NotDeviceEligible("stack_sum needs a C-contiguous float32 [K, ...] array")
            raws = (seq.reshape(-1).view(np.uint8),)
            nrows, shape = seq.shape[0], seq.shape[1:]
        else:
            if not isinstance(seq, (list, tuple)) or not seq:
                raise NotDeviceEligible("stack_sum needs a non-empty list of arrays")
            ops = [_as_operand(a) for a in seq]
            if any(type(o[0]) is DeviceRef for o in ops):
                raise NotDeviceEligible("device-resident rows go through the lattice compiler, not run()")
            if any(o[1] != ops[0][1] or o[2] != "C" for o in ops):
                raise NotDeviceEligible("stack_sum rows must share one C-ordered shape")
            raws = tuple(o[0] for o in ops)
            nrows, shape = len(ops), ops[0][1]
        row_bytes = int(np.prod(shape, dtype=np.int64)) * 4
        if row_bytes == 4 and nrows > 8:
            # np.sum(np.stack(xs), axis=0) over [K, 1] reduces a contiguous axis: numpy adds
            # pairwise there, the device adds in row order — the low bits can differ
            raise NotDeviceEligible("stack_sum rows of a single element reduce pairwise on the host")
        return EncodedTask(spec, raws, (0, 0), row_bytes // 4, nrows, row_bytes,
                           ("array", np.float32, tuple(shape), "C"))
    if k in _INT2 or k in _INT1:
        want = 2 if k in _INT2 else 1
        if len(pos) != want:
            raise NotDeviceEligible(f"{spec.name} takes {want} integer operand(s), got {len(pos)}")
        for v in pos:
            if type(v) is not int:
                raise NotDeviceEligible("integer electrons take plain Python ints")
            if not INT64_MIN <= v <= INT64_MAX:
                raise NotDeviceEligible("integer operand outside int64")
        imm = (pos[0], pos[1] if want == 2 else 0)
        return EncodedTask(spec, (), imm, 1, 0, 8, ("int",))
    raise NotDeviceEligible(f"kind {k} has no encoder")


def plan_call(spec: DeviceElectron, operands: Sequence) -> dict:
    """Shape inference for one call whose operands are known only by metadata —
    what the lattice compiler needs to lay a whole DAG out in the arena before
    any electron has run.  ``operands`` holds, per positional argument,
    ``("array", shape, order)``, ``("int", value)`` or a list of array metas.
    Returns count / nrows / out_bytes / result_meta, or raises NotDeviceEligible.
    """
    k = spec.kind

    def arr(m):
        if not (isinstance(m, tuple) and m and m[0] == "array"):
            raise NotDeviceEligible(f"{spec.name}: operand is not a float32 array")
        if len(m[1]) == 0:
            raise NotDeviceEligible("0-d arrays behave like numpy scalars on the host")
        return tuple(m[1]), m[2]

    if k in _ARRAY1:
        if len(operands) != 1:
            raise NotDeviceEligible(f"{spec.name} takes one array operand, got {len(operands)}")
        shape, order = arr(operands[0])
        n = int(np.prod(shape, dtype=np.int64))
        if k == K_SUM:
            return dict(count=n, nrows=0, out_bytes=4, result_meta=("f32",))
        return dict(count=4 * n if k == K_COPY else n, nrows=0, out_bytes=4 * n,
                    result_meta=("array", np.float32, shape, order))
    if k in _ARRAY2:
        if len(operands) != 2:
            raise NotDeviceEligible(f"{spec.name} takes two array operands, got {len(operands)}")
        (s0, o0), (s1, o1) = arr(operands[0]), arr(operands[1])
        if s0 != s1 or o0 != o1:
            raise NotDeviceEligible("binary electrons need operands of identical shape and order")
        n = int(np.prod(s0, dtype=np.int64))
        return dict(count=n, nrows=0, out_bytes=4 * n, result_meta=("array", np.float32, s0, o0))
    if k == K_STACK_SUM:
        if len(operands) != 1 or not isinstance(operands[0], list) or not operands[0]:
            raise NotDeviceEligible("stack_sum needs a non-empty list of arrays")
        metas = [arr(m) for m in operands[0]]
        if any(m != metas[0] or m[1] != "C" for m in metas):
            raise NotDeviceEligible("stack_sum rows must share one C-ordered shape")
        n = int(np.prod(metas[0][0], dtype=np.int64))
        return dict(count=n, nrows=len(metas), out_bytes=4 * n,
                    result_meta=("array", np.float32, metas[0][0], "C"))
    if k in _INT2 or k in _INT1:
        want = 2 if k in _INT2 else 1
        if len(operands) != want or any(not (isinstance(m, tuple) and m[0] == "int") for m in operands):
            raise NotDeviceEligible(f"{spec.name} takes {want} integer operand(s)")
        for m in operands:
            if type(m[1]) is not int or not INT64_MIN <= m[1] <= INT64_MAX:
                raise NotDeviceEligible("integer operand outside int64")
        return dict(count=1, nrows=0, out_bytes=8, result_meta=("int",))
    raise NotDeviceEligible(f"kind {k} has no planner")


def decode(meta: tuple, raw: np.ndarray) -> Any:
    """Result object from the result bytes (a uint8 view that may be reused by
    the caller afterwards: the returned object never aliases it)."""
    tag = meta[0]
    if tag == "array":
        _, dtype, shape, order = meta
        out = np.empty(shape, dtype=dtype, order=order)
        flat = out.reshape(-1, order="A" if order == "F" else "C")
        flat.view(np.uint8)[:] = raw[: flat.nbytes]
        return out
    if tag == "f32":
        return np.float32(raw[:4].view(np.float32)[0])
    if tag == "int":
        return int(raw[:8].view(np.int64)[0])
    raise ValueError(f"unknown result meta {meta!r}")


# ---------------------------------------------------------------------------
# A small standard library of registered electrons, used by the tests, the
# benchmark and the golden-vector generator.  The Python bodies ARE the
# definition of each electron (the oracle runs them under numpy on the CPU).

@device_electron("iadd", name="b200.add_int")
def add_int(a, b):
    return a + b


@device_electron("imul", name="b200.mul_int")
def mul_int(a, b):
    return a * b


@device_electron("iscale_add", ialpha=2, ibeta=1, name="b200.affine_int")
def affine_int(i):
    return 2 * i + 1


@device_electron("copy", name="b200.identity")
def identity(x):
    return x


@device_electron("scale_add", alpha=2.0, beta=1.0, name="b200.scale_add")
def scale_add(x):
    return np.float32(2.0) * x + np.float32(1.0)


@device_electron("scale_add", alpha=1.0, beta=1.0, name="b200.plus_one")
def plus_one(x):
    return x + np.float32(1.0)


@device_electron("add", name="b200.vec_add")
def vec_add(x, y):
    return x + y


@device_electron("mul", name="b200.vec_mul")
def vec_mul(x, y):
    return x * y


@device_electron("axpy", alpha=0.5, name="b200.axpy_half")
def axpy_half(x, y):
    return np.float32(0.5) * x + y


@device_electron("sum", name="b200.vec_sum")
def vec_sum(x):
    return x.sum()


@device_electron("stack_sum", name="b200.stack_sum")
def stack_sum(xs):
    return np.sum(np.stack(xs), axis=0)
