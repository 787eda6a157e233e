"""Import-path compatibility: ``from covalent_ssh_plugin import SSHExecutor``
(/root/reference/covalent_ssh_plugin/__init__.py:17) and the entry-point module
``covalent_ssh_plugin.ssh`` (reference setup.py:36) resolve to the B200
implementation in ``covalent_ssh_plugin_b200``."""

from covalent_ssh_plugin_b200.ssh import SSHExecutor  # noqa: F401
