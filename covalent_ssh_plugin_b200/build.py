"""Builds libb2exec.so (the C-ABI CUDA library, include/b2exec.h) in-tree with nvcc.

The library is compiled for sm_100a only; nvcc cross-compiles without a GPU.
The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libb2exec.so"
STAMP_PATH = PKG_DIR / ".libb2exec.stamp"
SOURCES = ["b2exec.cu"]
HEADERS = ["b2_device.cuh", "b2_spans.cuh", "b2_worker.cuh"]
PUBLIC_HEADER = PKG_DIR.parent / "include" / "b2exec.h"

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "--extended-lambda",
    "-Xcompiler",
    "-fPIC",
    "-shared",
    "-cudart",
    "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _fingerprint() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        h.update((CSRC / name).read_bytes())
    h.update(PUBLIC_HEADER.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    """True when the built library matches the current sources."""
    return (
        LIB_PATH.exists()
        and STAMP_PATH.exists()
        and STAMP_PATH.read_text().strip() == _fingerprint()
    )


def build_variant(out: Path, defines: dict) -> Path:
    """Tuning builds (tools/sweep.py): same sources with -D overrides, written to
    ``out`` and selected at run time with B2EXEC_LIB=<out>."""
    out.parent.mkdir(parents=True, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{k}={v}" for k, v in defines.items()], "-o", str(out)]
    cmd += [str(CSRC / s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    return out


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into libb2exec.so unless it is already up to date."""
    if not force and is_fresh():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    tmp = LIB_PATH.with_suffix(".so.tmp")
    cmd += ["-o", str(tmp)] + [str(CSRC / s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    os.replace(tmp, LIB_PATH)
    STAMP_PATH.write_text(_fingerprint() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    import sys

    path = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
