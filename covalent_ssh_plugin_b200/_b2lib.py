"""ctypes binding of libb2exec.so (C ABI declared in include/b2exec.h).

There is no CPU implementation behind this module: if the CUDA library is not
built, or a call fails, it raises.  ctypes releases the GIL around every call.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libb2exec.so"

# ---- enums mirrored from include/b2exec.h ----------------------------------
B2_OK = 0
B2_E_CUDA, B2_E_ARG, B2_E_STATE, B2_E_NOMEM, B2_E_NOTREADY, B2_E_UNSUPPORTED = -1, -2, -3, -4, -5, -6

TS_PENDING, TS_OK, TS_OVERFLOW, TS_BAD_FN, TS_BAD_ARG, TS_CANCELLED = 0, 1, 2, 3, 4, 5

K_NONE, K_COPY, K_SCALE_ADD, K_ADD, K_MUL, K_AXPY, K_SUM, K_STACK_SUM, K_IADD, K_IMUL, K_ISCALE_ADD = range(11)

F_IMM0, F_IMM1, F_CANCELLED = 0x1, 0x2, 0x4
BATCH_INDEPENDENT = 0x1
F_ARENA0_SHIFT, F_ARENA1_SHIFT, F_ARENAO_SHIFT = 8, 12, 16
NO_STATUS_COPY = 0xFFFFFFFFFFFFFFFF
MAX_FN = 1024
MAX_ARENAS = 16
ABI_VERSION = 2


class B2Error(RuntimeError):
    """A libb2exec call returned a negative b2_err code."""

    def __init__(self, code: int, message: str, call: str):
        super().__init__(f"{call} failed ({code}): {message}")
        self.code = code
        self.call = call


class LibraryMissing(RuntimeError):
    """libb2exec.so has not been built (run `python -m covalent_ssh_plugin_b200.build`)."""


class Task(C.Structure):
    """struct b2_task (64 bytes)."""

    _fields_ = [
        ("fn_id", C.c_uint32),
        ("flags", C.c_uint32),
        ("arg0", C.c_uint64),
        ("arg1", C.c_uint64),
        ("out", C.c_uint64),
        ("count", C.c_uint64),
        ("nrows", C.c_uint32),
        ("deps", C.c_uint32),
        ("succ_begin", C.c_uint32),
        ("succ_count", C.c_uint32),
        ("tag", C.c_uint64),
    ]


class Span(C.Structure):
    """struct b2_span (24 bytes)."""

    _fields_ = [("src_off", C.c_uint64), ("dst_off", C.c_uint64), ("bytes", C.c_uint64)]


class CopyJob(C.Structure):
    """struct b2_copy_job (24 bytes)."""

    _fields_ = [("src", C.c_uint64), ("dst_off", C.c_uint64), ("bytes", C.c_uint64)]


class Batch(C.Structure):
    """struct b2_batch: one whole batch for b2_submit."""

    _fields_ = [
        ("jobs", C.c_void_p),
        ("tasks", C.c_void_p),
        ("n_jobs", C.c_uint32),
        ("n_tasks", C.c_uint32),
        ("flags", C.c_uint32),
        ("reserved", C.c_uint32),
        ("in_off", C.c_uint64),
        ("in_bytes", C.c_uint64),
        ("status_host_off", C.c_uint64),
        ("out_dev_off", C.c_uint64),
        ("out_host_off", C.c_uint64),
        ("out_bytes", C.c_uint64),
        ("ticket", C.c_uint64),
    ]


class DeviceInfo(C.Structure):
    _fields_ = [
        ("cc_major", C.c_int32),
        ("cc_minor", C.c_int32),
        ("sm_count", C.c_int32),
        ("l2_bytes_mb", C.c_int32),
        ("total_mem", C.c_uint64),
        ("arena_bytes", C.c_uint64),
        ("staging_in_bytes", C.c_uint64),
        ("staging_out_bytes", C.c_uint64),
        ("queue_cap", C.c_uint32),
        ("abi_version", C.c_uint32),
        ("name", C.c_char * 64),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("launches_pack", C.c_uint64),
        ("launches_worker", C.c_uint64),
        ("launches_other", C.c_uint64),
        ("electrons", C.c_uint64),
        ("bytes_packed", C.c_uint64),
        ("bytes_h2d", C.c_uint64),
        ("bytes_d2h", C.c_uint64),
        ("resident_instances", C.c_uint64),
        ("resident_batches", C.c_uint64),
    ]


assert C.sizeof(Task) == 64 and C.sizeof(Span) == 24 and C.sizeof(CopyJob) == 24 and C.sizeof(Batch) == 88

# numpy dtypes with the same layouts (vectorised table building)
TASK_DTYPE = [
    ("fn_id", "<u4"), ("flags", "<u4"), ("arg0", "<u8"), ("arg1", "<u8"), ("out", "<u8"),
    ("count", "<u8"), ("nrows", "<u4"), ("deps", "<u4"), ("succ_begin", "<u4"),
    ("succ_count", "<u4"), ("tag", "<u8"),
]
SPAN_DTYPE = [("src_off", "<u8"), ("dst_off", "<u8"), ("bytes", "<u8")]
JOB_DTYPE = [("src", "<u8"), ("dst_off", "<u8"), ("bytes", "<u8")]

_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)

# name -> (argtypes, returns_status)
_SIGNATURES = {
    "b2_last_error": ([], False),
    "b2_abi_version": ([], False),
    "b2_source_hash": ([], False),
    "b2_device_count": ([C.POINTER(C.c_int)], True),
    "b2_init": ([C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32], True),
    "b2_shutdown": ([C.c_int], True),
    "b2_is_initialized": ([C.c_int], False),
    "b2_get_info": ([C.c_int, C.POINTER(DeviceInfo)], True),
    "b2_get_stats": ([C.c_int, C.POINTER(Stats)], True),
    "b2_host_buffer": ([C.c_int, C.c_int, C.POINTER(C.c_void_p), _u64p], True),
    "b2_arena_base": ([C.c_int, C.c_uint32, C.POINTER(C.c_void_p), _u64p], True),
    "b2_detach_host_buffer": ([C.c_int, C.c_int], True),
    "b2_h2d": ([C.c_int, C.c_uint64, C.c_uint64, C.c_uint64], True),
    "b2_d2h": ([C.c_int, C.c_uint64, C.c_uint64, C.c_uint64], True),
    "b2_write": ([C.c_int, C.c_uint64, C.c_void_p, C.c_uint64], True),
    "b2_read": ([C.c_int, C.c_uint64, C.c_void_p, C.c_uint64], True),
    "b2_fill": ([C.c_int, C.c_uint64, C.c_int, C.c_uint64], True),
    "b2_move_spans": ([C.c_int, C.c_void_p, C.c_uint32], True),
    "b2_pack": ([C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p, _u64p], True),
    "b2_unpack": ([C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, _u64p], True),
    "b2_register": ([C.c_int, C.c_uint32, C.c_int, C.c_double, C.c_double, C.c_int64, C.c_int64], True),
    "b2_enqueue": ([C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, _u32p], True),
    "b2_launch": ([C.c_int, C.c_uint64, _u64p], True),
    "b2_submit": ([C.c_int, C.POINTER(Batch)], True),
    "b2_stage_copy": ([C.c_int, C.c_void_p, C.c_uint32], True),
    "b2_cancel": ([C.c_int, C.c_uint64, C.c_uint32, C.c_uint32], True),
    "b2_table_create": ([C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, _u32p], True),
    "b2_table_launch": ([C.c_int, C.c_uint32, C.c_uint64, _u64p], True),
    "b2_table_destroy": ([C.c_int, C.c_uint32], True),
    "b2_resident_enable": ([C.c_int, C.c_uint32], True),
    "b2_resident_disable": ([C.c_int], True),
    "b2_mark": ([C.c_int, _u64p], True),
    "b2_poll": ([C.c_int, C.c_uint64], False),
    "b2_wait": ([C.c_int, C.c_uint64], True),
    "b2_sync": ([C.c_int], True),
    "b2_peer_enable": ([C.c_int, C.c_int, _u32p], True),
    "b2_ipc_export": ([C.c_int, C.c_void_p], True),
    "b2_ipc_open": ([C.c_int, C.c_void_p, C.c_uint64, _u32p], True),
    "b2_peer_copy": ([C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64], True),
    "b2_peer_pack": ([C.c_int, C.c_uint32, C.c_void_p, C.c_uint32], True),
    "b2_event_create": ([C.c_int, _u32p], True),
    "b2_event_record": ([C.c_int, C.c_uint32], True),
    "b2_event_elapsed_ms": ([C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)], True),
    "b2_flush_l2": ([C.c_int, C.c_uint64, C.c_uint64], True),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def library_path() -> Path:
    return Path(os.environ.get("B2EXEC_LIB", str(LIB_PATH)))


def load():
    """Load libb2exec.so once; raises LibraryMissing if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise LibraryMissing(
            f"{path} not found: build it with `python -m covalent_ssh_plugin_b200.build` "
            "(there is no CPU implementation to fall back to)"
        )
    try:
        lib = C.CDLL(str(path))
    except OSError as err:          # never let the connect-retry loop treat this as "host unreachable"
        raise LibraryMissing(f"{path} could not be loaded: {err}") from err
    for name, (argtypes, _) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = C.c_char_p if name in ("b2_last_error", "b2_source_hash") else C.c_int
    got = lib.b2_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"libb2exec ABI version {got} != binding version {ABI_VERSION}")
    _lib = lib
    return lib


def source_hash() -> str:
    """Which source tree the loaded library was built from (build.py fingerprint, 16 hex digits)."""
    return (load().b2_source_hash() or b"").decode("ascii", "replace")


def last_error() -> str:
    return (load().b2_last_error() or b"").decode("utf-8", "replace")


def call(name: str, *args) -> int:
    """Call a status-returning entry point; raise B2Error on a negative code."""
    rc = getattr(load(), name)(*args)
    if rc < 0:
        raise B2Error(rc, last_error(), name)
    return rc
