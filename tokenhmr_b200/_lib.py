"""ctypes binding of libtokenhmr_b200.so (the C ABI in include/tokenhmr_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libtokenhmr_b200.so"
_lib = None


class ThmrError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ThmrError(
                f"{LIB_PATH} not found: build it with `python -m tokenhmr_b200._build` "
                "(the engine has no non-CUDA fallback)")
        _lib = ctypes.CDLL(str(LIB_PATH))
        _declare(_lib)
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = lib().thmr_last_error().decode(errors="replace")
        raise ThmrError(f"tokenhmr_b200 call failed ({status}): {msg}")


def _declare(L: ctypes.CDLL) -> None:
    L.thmr_abi_version.restype = c_int
    L.thmr_last_error.restype = c_char_p
    L.thmr_check_device_flags.restype = c_int
    L.thmr_gemm_f16.restype = c_int
    L.thmr_gemm_f16.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]
