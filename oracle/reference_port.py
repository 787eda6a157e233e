"""TEST INFRASTRUCTURE — CPU restatement of the reference's per-electron path.

This is the ORACLE for the hot path and the CPU baseline that bench.py times
(`--impl reference`, `cpu_baseline`).  It is never imported by the product
(covalent_ssh_plugin_b200/); only tests/, __graft_entry__.smoke() and bench.py's
reference/cpu_baseline legs use it.

It restates, step for step, what /root/reference/covalent_ssh_plugin/ssh.py ::
SSHExecutor.run does for one electron, with the same LOCAL transport that
oracle/ref_harness.py substitutes for asyncssh (shell commands run locally,
"scp" is a file copy — no TCP, no SSH handshake, no encryption; this favours the
reference).  /root/reference does not exist on the GPU box, so this port is what
travels; it is pinned against the real reference by
tests/test_oracle_vs_reference.py (run in the build container) and against the
committed golden vectors by tests/test_oracle_golden.py (run everywhere).

Parity status: PINNED — every golden case in tests/golden/golden_v1.json was
produced by the unmodified reference (tests/golden/make_golden.py) and this
port reproduces each of them bit for bit.

Step map (reference line numbers):
  op id + workdir naming                ssh.py:482-491
  key file must exist                   ssh.py:317-335, 224-227
  "connect"                             ssh.py:237-282   (local: nothing to do)
  `<python> --version` must contain 3   ssh.py:521-524
  `mkdir -p <remote_cache>`             ssh.py:528-532
  pickle (fn, args, kwargs) to file     ssh.py:147-150   cloudpickle
  write the per-task run script         ssh.py:160-171   (+ exec.py:12-46 semantics)
  two uploads                           ssh.py:360-361
  `<python> <script>`                   ssh.py:377-383
  non-zero exit -> failure path         ssh.py:553-557
  `ls <result>` equals the path         ssh.py:401-406, 426-432
  download + unpickle (result, exc)     ssh.py:451-456
  three local + three remote deletes    ssh.py:310-315
  re-raise the electron's exception     ssh.py:581-583
  failure path: local run or error      ssh.py:202-208
"""

from __future__ import annotations

import asyncio
import os
import shutil
import sys
import tempfile
from pathlib import Path
from typing import Any, Callable, Dict, Optional

import cloudpickle

# The remote side (exec.py:12-46), restated.  The three @...@ markers are
# substituted per task, the way the reference formats its template
# (ssh.py:164-169).  Semantics kept: cloudpickle-or-exit(1) with the import
# error shipped in-band; run inside the workdir (created on demand) and always
# return to the starting directory; an electron that raises exits 0 and ships
# its exception as the second tuple element.
_REMOTE_SCRIPT = '''\
import os, sys
RESULT = os.path.realpath(@RESULT@)
TASK = os.path.realpath(@TASK@)
WORKDIR = os.path.realpath(@WORKDIR@)
try:
    import cloudpickle as codec
except Exception as import_error:
    import pickle as codec
    with open(RESULT, "wb") as sink:
        codec.dump((None, import_error), sink)
    sys.exit(1)
with open(TASK, "rb") as source:
    electron, positional, named = codec.load(source)
outcome, failure = None, None
start_dir = os.getcwd()
try:
    os.makedirs(WORKDIR, exist_ok=True)
    os.chdir(WORKDIR)
    outcome = electron(*positional, **named)
except Exception as raised:
    failure = raised
finally:
    os.chdir(start_dir)
with open(RESULT, "wb") as sink:
    codec.dump((outcome, failure), sink)
'''


class PortSettings:
    """The executor attributes the path reads (ssh.py:94-124), as plain data."""

    def __init__(self, workroot: str, python_path: str = sys.executable,
                 run_local_on_ssh_fail: bool = False, create_unique_workdir: bool = False,
                 do_cleanup: bool = True, poll_freq: float = 1.0, pythonpath: Optional[str] = None):
        root = Path(workroot)
        self.cache_dir = str(root / "cache")
        self.remote_cache = str(root / "remote")
        self.remote_workdir = str(root / "work")
        for d in (self.cache_dir, self.remote_cache, self.remote_workdir):
            os.makedirs(d, exist_ok=True)
        self.ssh_key_file = str(root / "id_stub")
        Path(self.ssh_key_file).touch()
        self.python_path = python_path
        self.run_local_on_ssh_fail = run_local_on_ssh_fail
        self.create_unique_workdir = create_unique_workdir
        self.do_cleanup = do_cleanup
        self.poll_freq = poll_freq
        self.hostname = "localhost"
        self.username = "user"
        self.env = dict(os.environ)
        if pythonpath:
            self.env["PYTHONPATH"] = pythonpath + os.pathsep + self.env.get("PYTHONPATH", "")


async def _shell(cmd: str, env: dict):
    """Local stand-in for ``conn.run`` (one process spawn per call)."""
    proc = await asyncio.create_subprocess_shell(
        cmd, stdout=asyncio.subprocess.PIPE, stderr=asyncio.subprocess.PIPE, env=env)
    out, err = await proc.communicate()
    return out.decode(), err.decode(), proc.returncode


def _fail(s: PortSettings, fn, args, kwargs, message: str):
    """ssh.py:202-208."""
    if s.run_local_on_ssh_fail:
        return fn(*args, **kwargs)
    raise RuntimeError(message)


async def run_electron(s: PortSettings, fn: Callable, args: list, kwargs: dict,
                       task_metadata: Dict[str, Any]) -> Any:
    """One electron, the reference's way (ssh.py:466-591)."""
    dispatch_id, node_id = task_metadata["dispatch_id"], task_metadata["node_id"]
    op = f"{dispatch_id}_{node_id}"
    workdir = s.remote_workdir
    if s.create_unique_workdir:
        workdir = os.path.join(workdir, dispatch_id, f"node_{node_id}")

    if not Path(s.ssh_key_file).is_file():
        raise RuntimeError(f"SSH key file {s.ssh_key_file} does not exist.")

    out, _, _ = await _shell(f"{s.python_path} --version", s.env)
    if "3" not in out.strip():
        return _fail(s, fn, args, kwargs, f"No Python 3 installation found on remote machine {s.hostname}")
    await _shell(f"mkdir -p {s.remote_cache}", s.env)

    task_local = os.path.join(s.cache_dir, f"function_{op}.pkl")
    script_local = os.path.join(s.cache_dir, f"exec_{op}.py")
    result_local = os.path.join(s.cache_dir, f"result_{op}.pkl")
    task_remote = os.path.join(s.remote_cache, f"function_{op}.pkl")
    script_remote = os.path.join(s.remote_cache, f"exec_{op}.py")
    result_remote = os.path.join(s.remote_cache, f"result_{op}.pkl")

    with open(task_local, "wb") as fh:
        cloudpickle.dump((fn, args, kwargs), fh)
    script = (_REMOTE_SCRIPT.replace("@RESULT@", repr(result_remote))
              .replace("@TASK@", repr(task_remote)).replace("@WORKDIR@", repr(workdir)))
    with open(script_local, "w", encoding="utf-8") as fh:
        fh.write(script)

    shutil.copyfile(task_local, task_remote)        # upload 1
    shutil.copyfile(script_local, script_remote)    # upload 2

    _, err, code = await _shell(f"{s.python_path} {script_remote}", s.env)
    if code != 0:
        return _fail(s, fn, args, kwargs, err.strip() or f"Task exited with nonzero exit status {code}.")

    tries = 5
    while True:
        listing, _, _ = await _shell(f"ls {result_remote}", s.env)
        if listing.strip() == result_remote:
            break
        if tries == 1:
            return _fail(s, fn, args, kwargs,
                         f"Result file {result_remote} on remote host {s.hostname} was not found")
        await asyncio.sleep(s.poll_freq)
        tries -= 1

    shutil.copyfile(result_remote, result_local)    # download
    with open(result_local, "rb") as fh:
        result, exception = cloudpickle.load(fh)

    if s.do_cleanup:
        for p in (task_local, script_local, result_local):
            os.remove(p)
        for p in (task_remote, script_remote, result_remote):
            await _shell(f"rm {p}", s.env)

    if exception is not None:
        raise exception
    return result


def run_one(fn: Callable, args: list, kwargs: Optional[dict] = None, workroot: Optional[str] = None,
            node_id: int = 0, **settings) -> Any:
    """Synchronous convenience wrapper around ``run_electron``."""
    own = workroot is None
    root = workroot or tempfile.mkdtemp(prefix="b2-oracle-")
    try:
        s = PortSettings(root, **settings)
        return asyncio.run(run_electron(s, fn, list(args), dict(kwargs or {}),
                                        {"dispatch_id": "oracle", "node_id": node_id}))
    finally:
        if own:
            shutil.rmtree(root, ignore_errors=True)


async def run_many(s: PortSettings, calls, concurrency: int) -> list:
    """Electrons in waves of ``concurrency`` coroutines, the way Covalent awaits
    independent electrons; used for the timed CPU baseline."""
    results = []
    for lo in range(0, len(calls), concurrency):
        wave = calls[lo:lo + concurrency]
        results += await asyncio.gather(*[
            run_electron(s, fn, list(a), dict(k), {"dispatch_id": "oracle", "node_id": lo + j})
            for j, (fn, a, k) in enumerate(wave)])
    return results


def in_process_floor(fn: Callable, args: list, kwargs: dict) -> Any:
    """Best-case CPU restatement (BASELINE.md §3 "T-floor"): the same pickle
    round trips with no files and no subprocess."""
    f, a, k = cloudpickle.loads(cloudpickle.dumps((fn, args, kwargs)))
    try:
        res, exc = f(*a, **k), None
    except Exception as err:  # exec.py:39-40
        res, exc = None, err
    res, exc = cloudpickle.loads(cloudpickle.dumps((res, exc)))
    if exc is not None:
        raise exc
    return res
