"""The C-ABI library builds, loads, and exports every symbol include/picovdb_b200.h declares.

No compute call is made here; without a GPU every entry point must fail loudly (no CPU fallback).
"""
import ctypes
import os
import re

import pytest

from picovdb_b200 import _native as N
from picovdb_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    with open(os.path.join(ROOT, "include", "picovdb_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pvdb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = B.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in N.SIGNATURES, f"{name} has no ctypes prototype"
    assert sorted(N.SIGNATURES) == declared
    assert N.load().pvdb_abi_version() == 1


def test_missing_library_is_an_import_error(tmp_path):
    with pytest.raises(ImportError, match="no CPU fallback"):
        N.load(str(tmp_path / "nope.so"))


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from picovdb_b200 import PicoVectorDB

    with pytest.raises(N.NativeError) as ei:
        PicoVectorDB(embedding_dim=4, storage_file="/tmp/never_created_pvdb")
    assert ei.value.code == N.PVDB_ERR_CUDA
    with pytest.raises(N.NativeError):
        N.device_count()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "picovdb_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert "oracle" not in src.lower(), f"{fn} mentions the oracle package"
