"""B200-native executor behind the covalent-ssh-plugin surface.

``from covalent_ssh_plugin_b200 import SSHExecutor, device_electron``

Importing the package does not load CUDA; the C library (libb2exec.so) is
loaded when the first executor connects to a GPU, and there is no CPU
implementation to fall back to if it is missing.
"""

from .electrons import device_electron, DeviceElectron, NotDeviceEligible  # noqa: F401
from .ssh import SSHExecutor, TaskCancelledError, EXECUTOR_PLUGIN_NAME, _EXECUTOR_PLUGIN_DEFAULTS  # noqa: F401

__all__ = ["SSHExecutor", "device_electron", "DeviceElectron", "NotDeviceEligible", "TaskCancelledError",
           "EXECUTOR_PLUGIN_NAME"]
__version__ = "0.1.0"
