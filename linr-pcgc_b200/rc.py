"""Host range coder front-end (csrc/rc_host.cpp): the torchac-shaped surface of the product.

`encode_float_cdf` / `decode_float_cdf` keep torchac's call signature (models/module_utils.py:28,38;
model_compression/model_size_est.py:482,561) for code that still builds float CDFs; the fast path feeds
the 16-bit CDF midpoints the GPU already produced (`encode_binary_batch`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

import numpy as np

from . import _lib


def host_cores() -> int:
    """Host cores this process may plan with: the box's cores divided by the ranks torchrun started on it (one rank
    per GPU share the host; LOCAL_WORLD_SIZE is absent in single-process runs)."""
    n = os.cpu_count() or 8
    try:
        aff = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        aff = n
    if aff < n:
        return max(2, aff)         # the rank was bound to a slice of the cores (dist.bind_rank_cores); while the coder runs, the
                                   # launch thread only waits for it
    try:
        n //= max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except ValueError:
        pass
    return max(2, n)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def encode_binary(cdf_mid: np.ndarray, sym: np.ndarray) -> bytes:
    lib = _lib.load()
    cdf_mid = np.ascontiguousarray(cdf_mid).view(np.uint16)
    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    n = int(sym.shape[0])
    cap = n // 4 + 64
    while True:
        out = np.empty(cap, dtype=np.uint8)
        w = lib.linr_rc_encode_binary(_p(cdf_mid), _p(sym), n, _p(out), cap)
        if w >= 0:
            return out[:w].tobytes()
        cap = int(-w)


def decode_binary(cdf_mid: np.ndarray, data: bytes, n: int) -> np.ndarray:
    lib = _lib.load()
    cdf_mid = np.ascontiguousarray(cdf_mid).view(np.uint16)
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.empty(n, dtype=np.uint8)
    _lib.check(lib.linr_rc_decode_binary(_p(cdf_mid), _p(buf) if len(buf) else None, len(buf), _p(out), n), "linr_rc_decode_binary")
    return out


def decode_binary_into(cdf_mid: np.ndarray, data: bytes, out: np.ndarray) -> None:
    """decode_binary writing into a caller-owned uint8 buffer (pinned staging memory in the decoder)."""
    lib = _lib.load()
    cdf_mid = np.ascontiguousarray(cdf_mid).view(np.uint16)
    buf = np.frombuffer(data, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"]
    _lib.check(lib.linr_rc_decode_binary(_p(cdf_mid), _p(buf) if len(buf) else None, len(buf), _p(out), int(out.shape[0])),
               "linr_rc_decode_binary")


def encode_binary_batch(cdf_mids: Sequence[np.ndarray], syms: Sequence[np.ndarray], shifts: Sequence[int] | None = None,
                        threads: int | None = None) -> List[bytes]:
    """Independent streams (8 stages x S scales of a frame) on a pool of host threads.

    Stream i codes bit `shifts[i]` of every byte of `syms[i]` (packed occupancy), default bit 0."""
    lib = _lib.load()
    k = len(cdf_mids)
    if k == 0:
        return []
    threads = threads or min(host_cores(), 32)
    cdf_mids = [np.ascontiguousarray(c).view(np.uint16) for c in cdf_mids]
    syms = [np.ascontiguousarray(s, dtype=np.uint8) for s in syms]
    ns = np.array([len(s) for s in syms], dtype=np.int64)
    sh = np.zeros(k, dtype=np.int32) if shifts is None else np.asarray(shifts, dtype=np.int32)
    caps = ns // 4 + 64
    while True:
        outs = [np.empty(int(c), dtype=np.uint8) for c in caps]
        cp = (C.c_void_p * k)(*[c.ctypes.data for c in cdf_mids])
        sp = (C.c_void_p * k)(*[s.ctypes.data for s in syms])
        op = (C.c_void_p * k)(*[o.ctypes.data for o in outs])
        written = np.zeros(k, dtype=np.int64)
        rc = lib.linr_rc_encode_binary_batch(k, cp, sp, _p(sh), _p(ns), op, _p(caps), _p(written), threads)
        if rc == 0:
            return [outs[i][: int(written[i])].tobytes() for i in range(k)]
        caps = np.maximum(caps, np.abs(written))


def cdf_float_to_u16(cdf_float: np.ndarray) -> np.ndarray:
    """torchac's float -> int16 CDF conversion (needs_normalization=True)."""
    Lp = cdf_float.shape[-1]
    v = np.rint(np.asarray(cdf_float, dtype=np.float32) * np.float32(65536 - (Lp - 1))).astype(np.int64)
    return ((v + np.arange(Lp)) & 0xFFFF).astype(np.uint16)


def encode_shared(cdf_row_u16: np.ndarray, sym: np.ndarray) -> bytes:
    lib = _lib.load()
    row = np.ascontiguousarray(cdf_row_u16, dtype=np.uint16)
    sym = np.ascontiguousarray(sym, dtype=np.int16)
    n = len(sym)
    cap = 2 * n + 64
    while True:
        out = np.empty(cap, dtype=np.uint8)
        w = lib.linr_rc_encode_shared(_p(row), len(row), _p(sym), n, _p(out), cap)
        if w > 0:
            return out[:w].tobytes()
        if w == 0:
            raise _lib.LinrError("linr_rc_encode_shared: symbol outside the alphabet")
        cap = int(-w)


def decode_shared(cdf_row_u16: np.ndarray, data: bytes, n: int) -> np.ndarray:
    lib = _lib.load()
    row = np.ascontiguousarray(cdf_row_u16, dtype=np.uint16)
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.empty(n, dtype=np.int16)
    _lib.check(lib.linr_rc_decode_shared(_p(row), len(row), _p(buf) if len(buf) else None, len(buf), _p(out), n),
               "linr_rc_decode_shared")
    return out


# ---- torchac-shaped entry points (same names / argument meaning as the package the reference imports) ----
def encode_float_cdf(cdf_float, sym, needs_normalization=True, check_input_bounds=False) -> bytes:
    import torch
    cdf = cdf_float.detach().cpu().numpy() if isinstance(cdf_float, torch.Tensor) else np.asarray(cdf_float)
    s = sym.detach().cpu().numpy() if isinstance(sym, torch.Tensor) else np.asarray(sym)
    if cdf.ndim != 2:
        raise ValueError("cdf must be [M, Lp]")
    u16 = cdf_float_to_u16(cdf)
    if cdf.shape[1] == 3:
        return encode_binary(u16[:, 1].copy(), s.astype(np.uint8))
    if len(u16) and not (u16 == u16[0]).all():
        raise NotImplementedError("per-symbol CDF rows with Lp > 3 are not on the LINR-PCGC path")
    return encode_shared(u16[0], s)


def decode_float_cdf(cdf_float, byte_stream, needs_normalization=True):
    import torch
    cdf = cdf_float.detach().cpu().numpy() if isinstance(cdf_float, torch.Tensor) else np.asarray(cdf_float)
    u16 = cdf_float_to_u16(cdf)
    n = cdf.shape[0]
    if cdf.shape[1] == 3:
        return torch.from_numpy(decode_binary(u16[:, 1].copy(), byte_stream, n).astype(np.int16))
    if len(u16) and not (u16 == u16[0]).all():
        raise NotImplementedError("per-symbol CDF rows with Lp > 3 are not on the LINR-PCGC path")
    return torch.from_numpy(decode_shared(u16[0], byte_stream, n))
