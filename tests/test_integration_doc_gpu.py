"""The ctypes binding shown in INTEGRATION.md §2 is executed as written (only the library
path is substituted), so the document cannot drift from the C ABI."""

import re
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu


def test_integration_md_binding_runs(gpu_count, built_library):
    if gpu_count < 1:
        pytest.skip("no CUDA device on this box")
    from covalent_ssh_plugin_b200 import _b2lib as L

    if L.load().b2_is_initialized(0):
        pytest.skip("device 0 is held by another fixture")
    text = (Path(__file__).resolve().parents[1] / "INTEGRATION.md").read_text()
    block = re.search(r"```python\n(# covalent_ssh_plugin/_b2\.py.*?)```", text, flags=re.S).group(1)
    code = block.replace('C.CDLL("libb2exec.so")', f'C.CDLL("{built_library}")')
    exec(compile(code, "INTEGRATION.md", "exec"), {})
