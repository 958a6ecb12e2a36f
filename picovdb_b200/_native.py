"""ctypes binding of ``libpicovdb_b200.so`` (see include/picovdb_b200.h).

There is no CPU fallback: :func:`load` raises when the shared library is missing and every
compute call raises :class:`NativeError` when no B200 is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libpicovdb_b200.so")

PVDB_OK = 0
PVDB_ERR_INVALID = -1
PVDB_ERR_CUDA = -2
PVDB_ERR_OOM = -3
PVDB_ERR_UNSUPPORTED = -4
PVDB_ERR_CAPACITY = -5

STORE_F32 = 0x1
STORE_BF16 = 0x2
STORE_FIXED_CAPACITY = 0x4

PREC_AUTO, PREC_F32, PREC_TF32, PREC_BF16 = 0, 1, 2, 3
SEARCH_QUERIES_NORMALIZED = 0x100
SEARCH_NO_RESCORE = 0x200
SEARCH_SCAN_ONLY = 0x400
SEARCH_NO_GUARD = 0x800
IPC_HANDLE_BYTES = 64

PRECISIONS = {"auto": PREC_AUTO, "f32": PREC_F32, "fp32": PREC_F32, "tf32": PREC_TF32, "bf16": PREC_BF16}


class NativeError(RuntimeError):
    """A C-ABI call returned a negative PVDB_ERR_* code."""

    def __init__(self, code: int, message: str) -> None:
        super().__init__(f"picovdb_b200 native error {code}: {message}")
        self.code = code
        self.message = message


class StoreInfo(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("ld_f32", C.c_int32),
        ("ld_bf16", C.c_int32),
        ("flags", C.c_int32),
        ("device", C.c_int32),
        ("reserved", C.c_int32),
        ("rows", C.c_int64),
        ("capacity", C.c_int64),
        ("active", C.c_int64),
        ("row_base", C.c_int64),
        ("device_bytes", C.c_uint64),
    ]


_P = C.c_void_p
_I64 = C.c_int64

# name -> (restype, argtypes); must list every function include/picovdb_b200.h declares
SIGNATURES = {
    "pvdb_abi_version": (C.c_int, []),
    "pvdb_last_error": (C.c_char_p, []),
    "pvdb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pvdb_store_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _I64, C.c_int]),
    "pvdb_store_destroy": (C.c_int, [_P]),
    "pvdb_store_reserve": (C.c_int, [_P, _I64]),
    "pvdb_store_info": (C.c_int, [_P, C.POINTER(StoreInfo)]),
    "pvdb_store_set_row_base": (C.c_int, [_P, _I64]),
    "pvdb_store_upsert": (C.c_int, [_P, _P, _P, _I64]),
    "pvdb_store_upsert_dev": (C.c_int, [_P, _P, _P, _I64, _I64, _P]),
    "pvdb_store_upsert_range": (C.c_int, [_P, _P, _I64, _I64]),
    "pvdb_store_upsert_range_dev": (C.c_int, [_P, _P, _I64, _I64, _P]),
    "pvdb_store_delete": (C.c_int, [_P, _P, _I64]),
    "pvdb_store_fetch": (C.c_int, [_P, _P, _I64, _P]),
    "pvdb_store_download": (C.c_int, [_P, _I64, _I64, _P]),
    "pvdb_store_upload": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "pvdb_store_download_bf16": (C.c_int, [_P, _I64, _I64, _P]),
    "pvdb_store_upload_bf16": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "pvdb_store_write_file": (C.c_int, [_P, C.c_char_p, _I64, _I64, _I64, C.c_int]),
    "pvdb_store_active_bits": (C.c_int, [_P, _P]),
    "pvdb_store_compact": (C.c_int, [_P, _P, _I64]),
    "pvdb_search": (C.c_int, [_P, _P, _I64, C.c_int, _P, C.c_int, _P, _P]),
    "pvdb_search_where": (C.c_int, [_P, _P, _I64, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, _P, _P,
                                    C.POINTER(_I64)]),
    "pvdb_store_column_write": (C.c_int, [_P, C.c_int, _P, _I64, _P, _I64]),
    "pvdb_store_column_drop": (C.c_int, [_P, C.c_int]),
    "pvdb_search_dev": (C.c_int, [_P, _P, _I64, C.c_int, _P, C.c_int, _P, _P, _P]),
    "pvdb_store_guard_stats": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64)]),
    "pvdb_exchange_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, _I64]),
    "pvdb_exchange_destroy": (C.c_int, [_P]),
    "pvdb_exchange_ipc_handle": (C.c_int, [_P, _P]),
    "pvdb_exchange_connect_ipc": (C.c_int, [_P, _P]),
    "pvdb_exchange_disconnect": (C.c_int, [_P]),
    "pvdb_exchange_connect_local": (C.c_int, [C.POINTER(_P), C.c_int]),
    "pvdb_exchange_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(_I64), C.POINTER(_I64)]),
    "pvdb_search_exchange": (C.c_int, [_P, _P, _P, _I64, C.c_int, _P, C.c_int, _P, _P]),
    "pvdb_search_exchange_dev": (C.c_int, [_P, _P, _P, _I64, C.c_int, _P, C.c_int, _P, _P, _P]),
    "pvdb_group_create": (C.c_int, [C.POINTER(_P), C.POINTER(C.c_int), C.c_int, C.c_int, _I64, C.c_int, _I64]),
    "pvdb_group_destroy": (C.c_int, [_P]),
    "pvdb_group_size": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(_I64)]),
    "pvdb_group_store": (_P, [_P, C.c_int]),
    "pvdb_group_search": (C.c_int, [_P, _P, _I64, C.c_int, _P, C.c_int, _P, _P]),
    "pvdb_merge_topk_dev": (C.c_int, [C.c_int, _P, _P, C.c_int, _I64, C.c_int, _I64, _I64, _P, _P, _P]),
    "pvdb_kernel_launches": (_I64, []),
}

_lib: Optional[C.CDLL] = None


def load(path: Optional[str] = None) -> C.CDLL:
    """dlopen the extension and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("PICOVDB_B200_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(
            f"picovdb_b200: CUDA extension {p} not found. Build it with "
            "`python -m picovdb_b200.build` (needs nvcc); there is no CPU fallback."
        )
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.pvdb_abi_version() != 1:
        raise ImportError(f"picovdb_b200: ABI version mismatch in {p}")
    if path is None:
        _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != PVDB_OK:
        msg = load().pvdb_last_error()
        raise NativeError(rc, msg.decode("utf-8", "replace") if msg else "")


def device_count() -> int:
    n = C.c_int(0)
    check(load().pvdb_device_count(C.byref(n)))
    return n.value


def kernel_launches() -> int:
    return int(load().pvdb_kernel_launches())
