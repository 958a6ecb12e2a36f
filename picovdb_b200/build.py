"""In-tree build of the CUDA extension (sm_100a only).

``python -m picovdb_b200.build`` compiles ``csrc/*.cu`` into ``picovdb_b200/libpicovdb_b200.so`` with
nvcc.  The library links the CUDA runtime statically and resolves driver entry points at run
time, so it loads (and its symbols can be inspected) on a machine without a GPU; every compute
entry point fails with PVDB_ERR_CUDA there.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libpicovdb_b200.so")
STAMP_PATH = os.path.join(PKG_DIR, ".libpicovdb_b200.stamp")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def _sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/picovdb_b200.h"]
    for f in files:
        path = os.path.normpath(os.path.join(CSRC, f))
        if os.path.isfile(path):
            h.update(f.encode())
            with open(path, "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build picovdb_b200's CUDA extension")
    return nvcc


def is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as f:
        return f.read().strip() == _fingerprint()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the shared library if the sources changed; return its path."""
    if not force and is_current():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-o", LIB_PATH, *_sources()]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr, file=sys.stderr)
    with open(STAMP_PATH, "w") as f:
        f.write(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
