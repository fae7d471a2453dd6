"""ctypes front-end of oracle/rc_oracle.c (torchac restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.rc_oracle_encode.restype = ctypes.c_longlong
        _lib.rc_oracle_encode.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong]
        _lib.rc_oracle_decode.restype = ctypes.c_int
        _lib.rc_oracle_decode.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong]
    return _lib


def encode_u16(cdf_u16: np.ndarray, sym: np.ndarray) -> bytes:
    """cdf_u16: [M,Lp] per-symbol rows, or [Lp] shared row.  sym: int16 [M]."""
    cdf = np.ascontiguousarray(cdf_u16, dtype=np.uint16)
    sym = np.ascontiguousarray(sym, dtype=np.int16)
    Lp = cdf.shape[-1]
    stride = Lp if cdf.ndim == 2 else 0
    n = len(sym)
    cap = n * 2 + 64  # generous; re-run if short
    while True:
        out = np.empty(cap, dtype=np.uint8)
        need = lib().rc_oracle_encode(cdf.ctypes.data, stride, Lp, sym.ctypes.data, n, out.ctypes.data, cap)
        if need <= cap:
            return out[:need].tobytes()
        cap = int(need)


def decode_u16(cdf_u16: np.ndarray, data: bytes, n: int) -> np.ndarray:
    cdf = np.ascontiguousarray(cdf_u16, dtype=np.uint16)
    Lp = cdf.shape[-1]
    stride = Lp if cdf.ndim == 2 else 0
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.empty(n, dtype=np.int16)
    lib().rc_oracle_decode(cdf.ctypes.data, stride, Lp, buf.ctypes.data if len(buf) else None, len(buf),
                           out.ctypes.data, n)
    return out


def float_cdf_to_u16(cdf_float: np.ndarray) -> np.ndarray:
    """torchac._convert_to_int_and_normalize(needs_normalization=True) [UPSTREAM]."""
    Lp = cdf_float.shape[-1]
    v = np.rint(cdf_float.astype(np.float32) * np.float32(65536 - (Lp - 1))).astype(np.int64)
    return ((v + np.arange(Lp)) & 0xFFFF).astype(np.uint16)


def encode_float_cdf(cdf_float: np.ndarray, sym: np.ndarray) -> bytes:
    """torchac.encode_float_cdf(cdf [M,Lp] f32, sym int16 [M]) (module_utils.py:28)."""
    return encode_u16(float_cdf_to_u16(np.asarray(cdf_float)), sym)


def decode_float_cdf(cdf_float: np.ndarray, data: bytes) -> np.ndarray:
    """torchac.decode_float_cdf (module_utils.py:38)."""
    c = float_cdf_to_u16(np.asarray(cdf_float))
    return decode_u16(c, data, c.shape[0])
