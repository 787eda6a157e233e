"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference plugin and runs it.

Usable where the reference is present: /root/reference (the build container) or the copy that
oracle/make_ref.py pip-installed into git-ignored oracle/_ref/ (which travels to the GPU box).
Nothing on the GPU box reads /root/reference.  It is used to (a) generate the golden
vectors under tests/golden/ (tests/golden/make_golden.py), (b) validate the
restated port in oracle/reference_port.py (tests/test_oracle_vs_reference.py) and
(c) time the reference's own code as the CPU baseline (bench.py, kind "reference").

The reference's two third-party imports that are absent here are stubbed:

* ``covalent`` — only ``Result`` (type hint), ``logger.app_log``,
  ``get_config`` and the ``RemoteExecutor`` base are touched
  (/root/reference/covalent_ssh_plugin/ssh.py:29-32).
* ``asyncssh`` — replaced by a LOCAL transport: ``connect`` returns an object
  whose ``run(cmd)`` executes the command with the local shell and whose "scp"
  is a file copy.  No TCP, no key exchange, no encryption, no remote ``scp -t``
  processes: every byte of the reference's control flow, file naming,
  templating and pickle formats is kept; only the SSH byte transport is
  substituted, which strictly favours the reference in timings.

The reference's ``ssh.py`` and ``exec.py`` are imported / formatted from where
they lie; no reference source is copied into this repository.
"""

from __future__ import annotations

import asyncio
import importlib.util
import logging
import os
import shutil
import sys
import types
from pathlib import Path
from typing import Any, Dict, Optional

def _reference_root() -> Path:
    """$B2_REFERENCE_ROOT, else the source tree (build container), else the installed copy."""
    if os.environ.get("B2_REFERENCE_ROOT"):
        return Path(os.environ["B2_REFERENCE_ROOT"])
    for cand in (Path("/root/reference"), Path(__file__).resolve().parent / "_ref"):
        if (cand / "covalent_ssh_plugin" / "ssh.py").is_file():
            return cand
    return Path("/root/reference")


REFERENCE_ROOT = _reference_root()
_PLUGIN_FILE = REFERENCE_ROOT / "covalent_ssh_plugin" / "ssh.py"

_module = None
_config: Dict[str, Any] = {}
_remote_env: Dict[str, Optional[dict]] = {"env": None}


def available() -> bool:
    return _PLUGIN_FILE.is_file()


# ------------------------------------------------------------ local "ssh" ---
class _Completed:
    def __init__(self, stdout: str, stderr: str, code: int):
        self.stdout = stdout
        self.stderr = stderr
        self.exit_status = code
        self.returncode = code


class _LocalConnection:
    """Runs commands with the local shell; stands in for SSHClientConnection."""

    def __init__(self, env: Optional[dict] = None):
        self._env = env
        self.commands = []

    async def run(self, cmd: str, *_, **__):
        self.commands.append(cmd)
        proc = await asyncio.create_subprocess_shell(
            cmd, stdout=asyncio.subprocess.PIPE, stderr=asyncio.subprocess.PIPE, env=self._env)
        out, err = await proc.communicate()
        return _Completed(out.decode(), err.decode(), proc.returncode)

    def close(self):
        return None

    async def wait_closed(self):
        return True


class _ConnectionLost(Exception):
    pass


def _install_stubs() -> None:
    def mod(name: str) -> types.ModuleType:
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    cov = mod("covalent")
    rm = mod("covalent._results_manager")
    rm.Result = type("Result", (), {})
    shared = mod("covalent._shared_files")
    logger = mod("covalent._shared_files.logger")
    logger.app_log = logging.getLogger("reference.covalent_ssh_plugin")
    logger.log_stack_info = False
    shared.logger = logger
    config = mod("covalent._shared_files.config")
    config.get_config = lambda key: _config[key]
    executor = mod("covalent.executor")
    plugins = mod("covalent.executor.executor_plugins")
    remote = mod("covalent.executor.executor_plugins.remote_executor")

    class RemoteExecutor:
        def __init__(self, poll_freq=15, remote_cache="", credentials_file="", **_):
            self.poll_freq = poll_freq
            self.remote_cache = remote_cache
            self.credentials_file = credentials_file

    remote.RemoteExecutor = RemoteExecutor
    cov._results_manager, cov._shared_files, cov.executor = rm, shared, executor
    executor.executor_plugins = plugins
    plugins.remote_executor = remote

    ssh = mod("asyncssh")

    async def connect(host, username=None, client_keys=None, known_hosts=None, **_):
        return _LocalConnection(_remote_env["env"])

    async def scp(src, dst, **_):
        s = src[1] if isinstance(src, tuple) else src
        d = dst[1] if isinstance(dst, tuple) else dst
        shutil.copyfile(s, d)

    ssh.connect = connect
    ssh.scp = scp
    ssh.ConnectionLost = _ConnectionLost
    ssh.SSHClientConnection = _LocalConnection
    ssh.SSHCompletedProcess = _Completed


def load_reference(cache_dir: str, remote_env: Optional[dict] = None):
    """Import the reference's ssh.py under the stubs (once per process)."""
    global _module
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _config.update({
        "dispatcher.cache_dir": cache_dir,
        "executors.ssh.username": "",
        "executors.ssh.hostname": "",
        "executors.ssh.ssh_key_file": "",
        "executors.ssh.remote_cache": "",
        "executors.ssh.python_path": "",
        "executors.ssh.conda_env": "",
        "executors.ssh.remote_workdir": "covalent-workdir",
        "executors.ssh.create_unique_workdir": False,
    })
    _remote_env["env"] = remote_env
    if _module is None:
        os.environ.setdefault("HOME", "/root")
        _install_stubs()
        spec = importlib.util.spec_from_file_location("reference_covalent_ssh_plugin_ssh", _PLUGIN_FILE)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _module = m
    return _module


def make_executor(workroot: str, pythonpath: Optional[str] = None, **overrides):
    """A reference SSHExecutor wired to the local transport.  ``workroot`` gets
    cache/ (local cache_dir), remote/ (remote_cache) and work/ (remote_workdir)."""
    root = Path(workroot)
    for sub in ("cache", "remote", "work"):
        (root / sub).mkdir(parents=True, exist_ok=True)
    key = root / "id_stub"
    key.touch()
    env = dict(os.environ)
    if pythonpath:
        env["PYTHONPATH"] = pythonpath + os.pathsep + env.get("PYTHONPATH", "")
    m = load_reference(str(root / "cache"), env)
    kwargs = dict(
        username="user", hostname="localhost", ssh_key_file=str(key), cache_dir=str(root / "cache"),
        python_path=sys.executable, remote_cache=str(root / "remote"),
        remote_workdir=str(root / "work"), poll_freq=1, retry_connect=False,
    )
    kwargs.update(overrides)
    return m.SSHExecutor(**kwargs)


def reference_run(executor, fn, args, kwargs, dispatch_id="golden", node_id=0):
    """One electron through the unmodified reference ``run``."""
    meta = {"dispatch_id": dispatch_id, "node_id": node_id}
    return asyncio.run(executor.run(fn, list(args), dict(kwargs), meta))
