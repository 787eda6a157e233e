"""Builds libb2exec.so (the C-ABI CUDA library, include/b2exec.h) in-tree with nvcc.

The library is compiled for sm_100a only; nvcc cross-compiles without a GPU.
The .so is git-ignored but travels to the GPU box with the repo snapshot.

Freshness: the fingerprint covers EVERY file under csrc/ (``*.cu``, ``*.cuh``,
``*.h`` — a glob, never a hand-kept list), the public header and the compiler
flags.  The same hash is compiled into the library (``-DB2_SRC_HASH``) and
returned by ``b2_source_hash()``, so a bench line or a test can say which
source tree the binary it ran was built from.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path
from typing import Iterable, List, Optional

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libb2exec.so"
STAMP_PATH = PKG_DIR / ".libb2exec.stamp"
PUBLIC_HEADER = PKG_DIR.parent / "include" / "b2exec.h"
SOURCE_PATTERNS = ("*.cu", "*.cuh", "*.cpp", "*.h")

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "--extended-lambda",
    "-Xcompiler",
    "-fPIC",
    "-Xcompiler",
    "-pthread",
    "-shared",
    "-cudart",
    "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def source_files(csrc: Path = CSRC) -> List[Path]:
    """Every file the library is compiled from, in a stable order."""
    found = set()
    for pat in SOURCE_PATTERNS:
        found.update(csrc.glob(pat))
    return sorted(found)


def translation_units(csrc: Path = CSRC) -> List[Path]:
    return sorted(csrc.glob("*.cu")) + sorted(csrc.glob("*.cpp"))


def fingerprint(csrc: Path = CSRC, header: Path = PUBLIC_HEADER, flags: Iterable[str] = NVCC_FLAGS) -> str:
    h = hashlib.sha256()
    for path in source_files(csrc):
        h.update(path.name.encode() + b"\0")
        h.update(path.read_bytes())
        h.update(b"\0")
    h.update(header.read_bytes())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def is_fresh(lib: Path = LIB_PATH, stamp: Path = STAMP_PATH, csrc: Path = CSRC) -> bool:
    """True when the built library matches the current sources."""
    return lib.exists() and stamp.exists() and stamp.read_text().strip() == fingerprint(csrc)


def _compile(out: Path, defines: dict, verbose: bool = False, csrc: Path = CSRC) -> str:
    sha = fingerprint(csrc)
    cmd = [_nvcc(), *NVCC_FLAGS, f'-DB2_SRC_HASH="{sha[:16]}"']
    cmd += [f"-D{k}={v}" for k, v in defines.items()]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", str(out)] + [str(s) for s in translation_units(csrc)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return sha


def build_variant(out: Path, defines: dict) -> Path:
    """Tuning builds (tools/sweep_build.py): same sources with -D overrides, written to
    ``out`` and selected at run time with B2EXEC_LIB=<out>."""
    out.parent.mkdir(parents=True, exist_ok=True)
    _compile(out, defines)
    return out


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into libb2exec.so unless it is already up to date."""
    if not force and is_fresh():
        return LIB_PATH
    tmp = LIB_PATH.with_suffix(".so.tmp")
    sha = _compile(tmp, {}, verbose=verbose)
    os.replace(tmp, LIB_PATH)
    STAMP_PATH.write_text(sha + "\n")
    return LIB_PATH


# ---- the CPython staging helper (hostext/b2host.c -> _b2host.so), host compiler only ----------
HOSTEXT_SRC = PKG_DIR / "hostext" / "b2host.c"
HOSTEXT_PATH = PKG_DIR / "_b2host.so"
HOSTEXT_STAMP = PKG_DIR / ".b2host.stamp"


def _hostext_fingerprint() -> str:
    import sysconfig

    h = hashlib.sha256()
    h.update(HOSTEXT_SRC.read_bytes())
    h.update(PUBLIC_HEADER.read_bytes())
    h.update(sysconfig.get_python_version().encode())
    return h.hexdigest()


def build_hostext(force: bool = False) -> Path:
    """Compile the staging helper against this interpreter's headers (gcc; no CUDA)."""
    import sysconfig

    if (not force and HOSTEXT_PATH.exists() and HOSTEXT_STAMP.exists()
            and HOSTEXT_STAMP.read_text().strip() == _hostext_fingerprint()):
        return HOSTEXT_PATH
    cc = os.environ.get("CC") or shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("no C compiler found for the _b2host helper (set CC)")
    tmp = HOSTEXT_PATH.with_suffix(".so.tmp")
    cmd = [cc, "-O2", "-Wall", "-shared", "-fPIC", f"-I{sysconfig.get_paths()['include']}",
           str(HOSTEXT_SRC), "-o", str(tmp)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"{cc} failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    os.replace(tmp, HOSTEXT_PATH)
    HOSTEXT_STAMP.write_text(_hostext_fingerprint() + "\n")
    return HOSTEXT_PATH


def build_all(force: bool = False, verbose: bool = False) -> Path:
    build_hostext(force)
    return build_library(force, verbose)


def built_source_hash() -> Optional[str]:
    """First 16 hex digits of the fingerprint the existing library was stamped with."""
    try:
        return STAMP_PATH.read_text().strip()[:16]
    except OSError:
        return None


if __name__ == "__main__":
    import sys

    path = build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path, built_source_hash())
