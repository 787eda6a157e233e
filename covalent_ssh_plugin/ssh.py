"""``covalent_ssh_plugin.ssh`` IS ``covalent_ssh_plugin_b200.ssh``: the module
object is aliased (not re-exported) so that attribute patches such as
``covalent_ssh_plugin.ssh.get_config`` reach the implementation."""

import sys

import covalent_ssh_plugin_b200.ssh as _impl

sys.modules[__name__] = _impl
