"""Stand-ins for the three things the plugin imports from the Covalent SDK
(/root/reference/covalent_ssh_plugin/ssh.py:29-32), used only when ``covalent``
itself is not installed: a logger, ``get_config`` and the ``RemoteExecutor`` base.

``get_config`` answers from (1) a TOML file named by ``$COVALENT_CONFIG_DIR`` /
``~/.config/covalent/covalent.conf`` if one exists, else (2) the same defaults
Covalent would have merged from ``_EXECUTOR_PLUGIN_DEFAULTS``.
"""

from __future__ import annotations

import logging
import os
from pathlib import Path
from typing import Any

app_log = logging.getLogger("covalent_ssh_plugin_b200")
log_stack_info = False

_BUILTIN = {
    "dispatcher": {"cache_dir": os.path.join(os.environ.get("HOME", "/tmp"), ".cache/covalent")},
    "executors": {
        "ssh": {
            "username": "",
            "hostname": "",
            "ssh_key_file": os.path.join(os.environ.get("HOME", "/tmp"), ".ssh/id_rsa"),
            "python_path": "python",
            "conda_env": "",
            "remote_cache": ".cache/covalent",
            "run_local_on_ssh_fail": False,
            "remote_workdir": "covalent-workdir",
            "create_unique_workdir": False,
        }
    },
}

_loaded = None


def _config() -> dict:
    global _loaded
    if _loaded is None:
        cfg_dir = os.environ.get("COVALENT_CONFIG_DIR") or os.path.join(
            os.environ.get("HOME", "/tmp"), ".config/covalent")
        path = Path(cfg_dir) / "covalent.conf"
        data = {}
        if path.is_file():
            try:
                import tomllib             # Python >= 3.11
            except ImportError:            # 3.9 / 3.10 (pyproject: requires-python >= 3.9)
                import tomli as tomllib
            with open(path, "rb") as fh:
                data = tomllib.load(fh)
        _loaded = data
    return _loaded


def get_config(key: str) -> Any:
    """Dotted-key lookup, user file first, built-in defaults second."""
    for root in (_config(), _BUILTIN):
        node = root
        try:
            for part in key.split("."):
                node = node[part]
            return node
        except (KeyError, TypeError):
            continue
    raise KeyError(key)


class RemoteExecutor:
    """The slice of covalent's RemoteExecutor the plugin relies on: it stores
    ``poll_freq`` / ``remote_cache`` and offers no-op ``setup`` / ``teardown``."""

    def __init__(self, poll_freq: int = 15, remote_cache: str = "", credentials_file: str = "",
                 **kwargs) -> None:
        self.poll_freq = poll_freq
        self.remote_cache = remote_cache
        self.credentials_file = credentials_file

    async def setup(self, task_metadata: dict) -> None:
        return None

    async def teardown(self, task_metadata: dict) -> None:
        return None
