"""In-tree build of the CUDA extension (sm_100a only).

``python -m picovdb_b200.build`` compiles every ``csrc/*.cu`` into an object file (in parallel, one
nvcc process per translation unit, cached by a hash of the source + headers + flags under
``picovdb_b200/_build/``) and links them into ``picovdb_b200/libpicovdb_b200.so``.  The library
links the CUDA runtime statically and resolves driver entry points at run time, so it loads (and
its symbols can be inspected) on a machine without a GPU; every compute entry point fails with
PVDB_ERR_CUDA there.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "_build")
LIB_PATH = os.path.join(PKG_DIR, "libpicovdb_b200.so")
STAMP_PATH = os.path.join(PKG_DIR, ".libpicovdb_b200.stamp")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    *os.environ.get("PVDB_NVCC_EXTRA", "").split(),  # e.g. -DPVDB_SCAN_THREADS=384 for tuning experiments
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]


def _sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest() -> bytes:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h", ".inl")):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(f.encode() + fh.read())
    with open(os.path.join(INCLUDE, "picovdb_b200.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(COMPILE_FLAGS).encode())
    return h.digest()


def _object_key(src: str, headers: bytes) -> str:
    h = hashlib.sha256(headers)
    with open(src, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()[:24]


def _fingerprint() -> str:
    headers = _headers_digest()
    h = hashlib.sha256(" ".join(LINK_FLAGS).encode())
    for src in _sources():
        h.update(os.path.basename(src).encode() + _object_key(src, headers).encode())
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


def build_variant(name: str, extra_flags: list[str]) -> str:
    """An instrumented / experimental copy of the library next to the product one, e.g.
    ``build_variant("stats", ["-DPVDB_BATCH_STATS"])`` -> picovdb_b200/_variants/libpicovdb_b200_stats.so
    (select it with PICOVDB_B200_LIB=...).  Never used by the product path."""
    nvcc = find_nvcc()
    out_dir = os.path.join(PKG_DIR, "_variants")
    obj_dir = os.path.join(out_dir, name)
    os.makedirs(obj_dir, exist_ok=True)
    lib = os.path.join(out_dir, f"libpicovdb_b200_{name}.so")

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.splitext(os.path.basename(src))[0] + ".o")
        cmd = [nvcc, *COMPILE_FLAGS, *extra_flags, "-I", INCLUDE, "-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n{res.stdout}\n{res.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 1)) as ex:
        objects = list(ex.map(compile_one, _sources()))
    res = subprocess.run([nvcc, *LINK_FLAGS, "-o", lib, *objects], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc link failed:\n{res.stdout}\n{res.stderr}")
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile what changed, link, return the library path."""
    if not force and is_current():
        return LIB_PATH
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = _headers_digest()
    jobs = []
    objects = []
    for src in _sources():
        stem = os.path.splitext(os.path.basename(src))[0]
        obj = os.path.join(OBJ_DIR, f"{stem}.{_object_key(src, headers)}.o")
        objects.append(obj)
        if force or not os.path.exists(obj):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *COMPILE_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-I", INCLUDE, "-c", src, "-o", obj + ".tmp"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n{res.stdout}\n{res.stderr}")
        os.replace(obj + ".tmp", obj)
        if verbose:
            print(res.stderr, file=sys.stderr)

    if jobs:
        workers = max(1, min(len(jobs), os.cpu_count() or 1))
        with ThreadPoolExecutor(max_workers=workers) as ex:
            list(ex.map(compile_one, jobs))
    res = subprocess.run([nvcc, *LINK_FLAGS, "-o", LIB_PATH, *objects], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc link failed:\n{res.stdout}\n{res.stderr}")
    keep = set(objects)
    for f in os.listdir(OBJ_DIR):  # drop objects of older source versions
        path = os.path.join(OBJ_DIR, f)
        if path not in keep:
            os.remove(path)
    with open(STAMP_PATH, "w") as f:
        f.write(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
        print(path)
