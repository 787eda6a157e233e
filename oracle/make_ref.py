"""TEST INFRASTRUCTURE — recipe that installs the UNMODIFIED reference into oracle/_ref/.

    python oracle/make_ref.py            # build container only: needs /root/reference

``oracle/_ref/`` is git-ignored (no reference source ever enters this repository's history)
but NOT gpurun-ignored, so the installed package travels to the GPU box with the snapshot,
exactly like this repo's own built ``.so`` files.  There ``bench.py --impl reference`` and the
``cpu_baseline`` leg time the reference's own ``SSHExecutor.run`` + ``exec.py``
(through oracle/ref_harness.py: stub ``covalent`` modules, local-transport ``asyncssh``)
instead of the restated port.

How: the same offline install the base contract prescribes for ``baseline/_ref`` —
``pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of
/root/reference>`` (the copy lives under /tmp because /root/reference is read-only and
setuptools writes build artefacts next to setup.py; ``--no-deps`` because neither ``covalent``
nor ``asyncssh`` is installable here — the harness stubs exactly the names ssh.py:27-32 imports).
Called by ``__graft_entry__.build()`` whenever /root/reference exists.
"""

from __future__ import annotations

import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

REFERENCE = Path("/root/reference")
TARGET = Path(__file__).resolve().parent / "_ref"
MARKER = TARGET / "covalent_ssh_plugin" / "ssh.py"


def installed() -> bool:
    return MARKER.is_file()


def make_ref(force: bool = False) -> str:
    """Returns "installed", "present" or the one-line reason it could not be done."""
    if installed() and not force:
        return "present"
    if not (REFERENCE / "setup.py").is_file():
        return f"{REFERENCE} is not available on this machine"
    with tempfile.TemporaryDirectory(prefix="b2-ref-src-") as tmp:
        src = Path(tmp) / "reference"
        shutil.copytree(REFERENCE, src)
        if TARGET.exists():
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", str(TARGET), str(src)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0 or not installed():
            return f"pip install of the reference failed: {(proc.stderr or proc.stdout).strip()[-300:]}"
    return "installed"


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
