"""Warm worker processes for electrons that cannot run on a GPU (arbitrary Python).

OPT-IN (config key ``executors.ssh.host_pool`` or ``B2_HOST_POOL=1``): by default
an unregistered electron takes the reference's failure decision
(``_on_ssh_fail``) and nothing ever runs on the CPU behind the user's back.

When enabled, this is the reference's remote side without the SSH transport and
without a fresh interpreter per electron (/root/reference/covalent_ssh_plugin/
exec.py:29-46: unpickle ``(fn, args, kwargs)``, create and enter the workdir,
call, always return to the previous directory, ship ``(result, exception)``
back): the same semantics, in a pool of long-lived processes that have already
paid the 70-450 ms interpreter/numpy start the reference pays per electron
(SURVEY.md §6).  It is NOT a fallback for device electrons: those never come
here, and a device failure is still reported as a failure.
"""

from __future__ import annotations

import asyncio
import os
from concurrent.futures import ProcessPoolExecutor
from concurrent.futures.process import BrokenProcessPool
from multiprocessing import get_context
from typing import Any, Callable, Optional, Tuple

import cloudpickle

stats = {"submitted": 0, "crashed": 0}
_pool: Optional[ProcessPoolExecutor] = None
_atexit_done = False


def enabled(get_config) -> bool:
    env = os.environ.get("B2_HOST_POOL", "").lower()
    if env in ("1", "true", "yes"):
        return True
    if env in ("0", "false", "no"):
        return False
    try:
        return bool(get_config("executors.ssh.host_pool"))
    except Exception:
        return False


def _remote_call(blob: bytes, workdir: str) -> bytes:
    """Runs inside a worker process (exec.py:29-46 restated)."""
    fn, args, kwargs = cloudpickle.loads(blob)
    result, exception = None, None
    here = os.getcwd()
    try:
        os.makedirs(workdir or ".", exist_ok=True)
        os.chdir(workdir or ".")
        result = fn(*args, **kwargs)
    except Exception as err:          # BaseException (SystemExit, ...) kills the worker, like a non-zero exit
        exception = err
    finally:
        os.chdir(here)
    return cloudpickle.dumps((result, exception))


def _get_pool() -> ProcessPoolExecutor:
    global _pool
    if _pool is None:
        n = int(os.environ.get("B2_HOST_POOL_WORKERS", "0")) or min(8, len(os.sched_getaffinity(0)))
        _pool = ProcessPoolExecutor(max_workers=n, mp_context=get_context("spawn"))
        global _atexit_done
        if not _atexit_done:
            import atexit

            atexit.register(shutdown)
            _atexit_done = True
    return _pool


async def call(fn: Callable, args, kwargs, workdir: str) -> Tuple[int, str, Any, Optional[Exception]]:
    """(exit_status, stderr, result, exception) of one electron run in the pool."""
    global _pool
    stats["submitted"] += 1
    try:
        blob = cloudpickle.dumps((fn, list(args), dict(kwargs)))
    except Exception as err:
        return 1, f"electron is not picklable: {err}", None, None
    pool = _get_pool()
    try:
        out = await asyncio.wrap_future(pool.submit(_remote_call, blob, workdir))
    except BrokenProcessPool:
        stats["crashed"] += 1
        _discard(pool)                 # a fresh pool for the next electron
        _pool = None
        return 1, "worker process died while running the electron", None, None
    except Exception as err:           # result could not be pickled, etc.
        return 1, f"{type(err).__name__}: {err}", None, None
    result, exception = cloudpickle.loads(out)
    return 0, "", result, exception


def _discard(pool: ProcessPoolExecutor) -> None:
    """Stop a pool and make sure none of its workers outlives it."""
    procs = list(getattr(pool, "_processes", {}).values())
    pool.shutdown(wait=False, cancel_futures=True)
    for p in procs:
        try:
            p.terminate()
        except Exception:
            pass
    for p in procs:
        try:
            p.join(2)
        except Exception:
            pass


def shutdown() -> None:
    global _pool
    if _pool is not None:
        _discard(_pool)
        _pool = None
