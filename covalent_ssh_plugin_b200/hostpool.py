"""Warm worker processes for electrons that cannot run on a GPU (arbitrary Python).

Two kinds of worker, same semantics:
* default interpreter (``python_path`` left at "python", no ``conda_env``): a ProcessPoolExecutor
  of this interpreter;
* the executor names its own interpreter (``python_path`` / ``conda_env``, ssh.py:377-380 and
  508-524 in the reference): ``hostworker.py`` started with the reference's command line
  (``conda activate <env> && <python_path> ...``), verified the way the reference verifies the
  remote side (conda env must exist, ``--version`` must say Python 3), kept warm, and fed
  protocol-5 pickles whose large buffers travel through /dev/shm instead of through the pipe.

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


def _default_interpreter(python_path: str, conda_env: str) -> bool:
    return not conda_env and python_path in ("", "python")


async def call(fn: Callable, args, kwargs, workdir: str, python_path: str = "python",
               conda_env: str = "") -> Tuple[int, str, Any, Optional[Exception]]:
    """(exit_status, stderr, result, exception) of one electron run in the pool."""
    global _pool
    stats["submitted"] += 1
    if not _default_interpreter(python_path, conda_env or ""):
        return await _external_call(fn, args, kwargs, workdir, python_path or "python", conda_env or "")
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


# ---------------------------------------------------------------------------------------------
# workers under the executor's own interpreter (python_path / conda_env)
class _External:
    """One warm ``hostworker.py`` process.  Plain pipes and a thread lock — not asyncio streams —
    so that a worker outlives the event loop it was first used from (tests call asyncio.run per
    electron; Covalent's dispatcher keeps one loop)."""

    def __init__(self, proc):
        import threading

        self.proc = proc
        self.lock = threading.Lock()
        self.busy = 0
        self.serial = 0

    def alive(self) -> bool:
        return self.proc.poll() is None

    def roundtrip(self, frame):
        from . import hostworker

        with self.lock:
            hostworker.send(self.proc.stdin, frame)
            return hostworker.recv(self.proc.stdout)


_external: dict = {}          # (python_path, conda_env) -> list of _External
_verified: dict = {}          # (python_path, conda_env) -> "" | failure message


def _conda_prefix(conda_env: str) -> str:
    return f'eval "$(conda shell.bash hook)" && conda activate {conda_env} && ' if conda_env else ""


async def _shell(cmd: str):
    proc = await asyncio.create_subprocess_shell(cmd, stdout=asyncio.subprocess.PIPE, stderr=asyncio.subprocess.PIPE)
    out, err = await proc.communicate()
    return out.decode(errors="replace"), err.decode(errors="replace"), proc.returncode


async def _verify(python_path: str, conda_env: str) -> str:
    """The reference's two remote checks (ssh.py:508-524), once per interpreter: the conda
    environment must be listed, and ``<python_path> --version`` must mention Python 3."""
    key = (python_path, conda_env)
    if key in _verified:
        return _verified[key]
    message = ""
    if conda_env:
        _, err, code = await _shell(f'eval "$(conda shell.bash hook)" && conda env list | grep {conda_env}')
        if code != 0:
            message = err.strip() or f"No conda environment named {conda_env} found on remote machine."
    if not message:
        out, _, _ = await _shell(f"{_conda_prefix(conda_env)}{python_path} --version")
        if "3" not in out.strip():                  # the reference's test, on stdout only (ssh.py:521-522)
            message = "No Python 3 installation found on remote machine localhost"
    _verified[key] = message
    return message


def _spawn(python_path: str, conda_env: str) -> _External:
    import subprocess
    from pathlib import Path

    from . import hostworker

    worker = Path(__file__).with_name("hostworker.py")
    cmd = f"{_conda_prefix(conda_env)}exec {python_path} {worker}"
    proc = subprocess.Popen(cmd, shell=True, executable="/bin/bash", stdin=subprocess.PIPE, stdout=subprocess.PIPE)
    w = _External(proc)
    try:
        kind, problem = hostworker.recv(proc.stdout)
    except EOFError:
        kind, problem = "eof", "the interpreter exited at once"
    if kind != "hello" or problem:
        proc.kill()
        raise RuntimeError(f"host worker under '{python_path}' could not start: {problem or kind}")
    return w


async def _external_call(fn, args, kwargs, workdir: str, python_path: str, conda_env: str):
    from . import hostworker

    problem = await _verify(python_path, conda_env)
    if problem:
        return 1, problem, None, None
    loop = asyncio.get_running_loop()
    workers = _external.setdefault((python_path, conda_env), [])
    workers[:] = [x for x in workers if x.alive()]
    limit = int(os.environ.get("B2_HOST_POOL_WORKERS", "0")) or min(8, len(os.sched_getaffinity(0)))
    w = next((x for x in workers if x.busy == 0), None)
    if w is None and len(workers) < limit:
        try:
            w = await loop.run_in_executor(None, _spawn, python_path, conda_env)
        except Exception as err:
            return 1, str(err), None, None
        workers.append(w)
        _register_atexit()
    if w is None:
        w = min(workers, key=lambda x: x.busy)
    w.busy += 1
    shm_path = ""
    try:
        w.serial += 1
        shm_dir = "/dev/shm" if os.path.isdir("/dev/shm") else None
        try:
            if shm_dir:
                blob, shm_path, ranges = hostworker.pack(cloudpickle, (fn, list(args), dict(kwargs)), shm_dir,
                                                         f"q{id(w) & 0xffff}-{w.serial}")
            else:
                blob, shm_path, ranges = cloudpickle.dumps((fn, list(args), dict(kwargs))), "", []
        except Exception as err:
            return 1, f"electron is not picklable: {err}", None, None
        try:
            rblob, rshm, rranges = await loop.run_in_executor(None, w.roundtrip, (blob, shm_path, ranges, workdir))
        except (EOFError, BrokenPipeError, ConnectionResetError, OSError):
            stats["crashed"] += 1
            if w in workers:
                workers.remove(w)
            w.proc.kill()
            if shm_path and os.path.exists(shm_path):
                os.unlink(shm_path)
            return 1, "worker process died while running the electron", None, None
        (result, exception), _ = hostworker.unpack(cloudpickle, rblob, rshm, rranges)
    finally:
        w.busy -= 1
    return 0, "", result, exception


def _register_atexit() -> None:
    global _atexit_done
    if not _atexit_done:
        import atexit

        atexit.register(shutdown)
        _atexit_done = True


def shutdown() -> None:
    global _pool
    if _pool is not None:
        _discard(_pool)
        _pool = None
    for workers in _external.values():
        for w in workers:
            try:
                w.proc.kill()
                w.proc.wait(2)
                w.proc.stdin.close()
                w.proc.stdout.close()
            except Exception:
                pass
    _external.clear()
    _verified.clear()
