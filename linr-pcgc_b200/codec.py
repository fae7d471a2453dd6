"""Frame coding on top of the C ABI: teacher-forced encode (8 streams per scale) and sequential decode.

Host-side mirror of `encode_one_frame` (encoder.py:158-203) / `CNP.encode` (models/upsample.py:219-246) and
`decode_one_frame` (decoder.py:153-176) / `CNP.decode` (models/upsample.py:249-295).  The GPU emits the 16-bit CDF
boundaries torchac would derive from the float probabilities; only those 2 B/symbol cross PCIe, and the serial range
coder (csrc/rc_host.cpp) runs on the host over the independent streams in parallel.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import rc
from .frame import Frame, RowTables, build_tables, octree_up
from .net import NetRunner


def pack_bitstream(parts: Sequence[bytes]) -> bytes:
    """models/function_utils.py:109-116: u32 count | u32 len[count] | payloads."""
    head = np.array([len(parts)] + [len(b) for b in parts], dtype="<u4").tobytes()
    return head + b"".join(parts)


def unpack_bitstream(buf: bytes) -> List[bytes]:
    """models/function_utils.py:119-132."""
    n = int(np.frombuffer(buf[:4], dtype="<u4")[0])
    lens = np.frombuffer(buf[4: 4 + 4 * n], dtype="<u4")
    out, s = [], 4 + 4 * n
    for l in lens:
        out.append(buf[s: s + int(l)])
        s += int(l)
    return out


class _Pinned:
    """Reusable pinned staging buffers (grown on demand) so encode makes no host allocation per frame."""

    def __init__(self):
        self.cdf = None
        self.occ = None

    def get(self, rows: int):
        if self.cdf is None or self.cdf.shape[1] < rows:
            cap = max(rows, 1)
            self.cdf = torch.empty((8, cap), dtype=torch.int16).pin_memory()
            self.occ = torch.empty(cap, dtype=torch.uint8).pin_memory()
        return self.cdf, self.occ


_pinned = _Pinned()


def frame_cdfs_to_host(runner: NetRunner, params: torch.Tensor, frame: Frame):
    """Forward (no grad) + D2H of the CDF boundaries and the occupancy bytes -> numpy views [8,R] u16, [R] u8."""
    t = frame.tables
    R = t.n_rows
    out = runner.forward(params, t, train=False, want_cdf=True, want_bits=False)
    h_cdf, h_occ = _pinned.get(R)
    h_cdf[:, :R].copy_(out["cdf"], non_blocking=True)
    h_occ[:R].copy_(t.occ, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h_cdf.numpy().view(np.uint16), h_occ.numpy(), R


def encode_frame(runner: NetRunner, params: torch.Tensor, frame: Frame, threads: Optional[int] = None) -> List[bytes]:
    """All scales of one frame -> list of per-scale packed bitstreams (frame%04d_scale%d.bin payloads)."""
    cdf, occ, R = frame_cdfs_to_host(runner, params, frame)
    cdfs, syms, shifts = [], [], []
    for s in range(frame.n_scales):
        a, b = frame.scale_off[s], frame.scale_off[s + 1]
        for k in range(8):
            cdfs.append(cdf[k, a:b])
            syms.append(occ[a:b])
            shifts.append(k)
    streams = rc.encode_binary_batch(cdfs, syms, shifts, threads)
    return [pack_bitstream(streams[8 * s: 8 * s + 8]) for s in range(frame.n_scales)]


def decode_scale(runner: NetRunner, params: torch.Tensor, coords: torch.Tensor, scale_idx: int, data: bytes):
    """One scale: parents `coords` (sorted unique, CUDA int32 [N,3]) + its bitstream -> uint8 occupancy [N] (CUDA)."""
    n = int(coords.shape[0])
    dev = coords.device
    scale = torch.full((n,), scale_idx, dtype=torch.uint8, device=dev)
    occ = torch.zeros(n, dtype=torch.uint8, device=dev)
    t = build_tables(coords, scale, occ)
    streams = unpack_bitstream(data)
    runner.decode_begin(params, t)
    h_cdf = torch.empty(n, dtype=torch.int16).pin_memory()
    h_sym = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_sym = torch.empty(n, dtype=torch.uint8, device=dev)
    for k in range(8):
        d_cdf, _ = runner.decode_stage(params, t, k)
        h_cdf.copy_(d_cdf, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        sym = rc.decode_binary(h_cdf.numpy().view(np.uint16), streams[k], n)
        h_sym.numpy()[:] = sym
        d_sym.copy_(h_sym, non_blocking=True)
        runner.occ_set_stage(occ, d_sym, k)
    return occ, t


def decode_frame(runner: NetRunner, params: torch.Tensor, all_bytes: Sequence[bytes], low_coords: torch.Tensor,
                 low_bits: int = 8) -> torch.Tensor:
    """decode_one_frame (decoder.py:153-176): coarse-to-fine; returns the full-resolution sorted coords (CUDA int32)."""
    cur = low_coords
    bits = max(low_bits, 1)
    for s in range(len(all_bytes) - 1, -1, -1):
        occ, _ = decode_scale(runner, params, cur, s, all_bytes[s])
        bits += 1
        cur = octree_up(cur, occ, bits)
    return cur


def pack_low_xyz(low_coords: Sequence[np.ndarray], mins: Sequence[np.ndarray]) -> bytes:
    """enc_all_frame_low_xyz (test_utils.py:199-232): uint8 xyz per frame, then int32 mins [F,3]."""
    parts = []
    for c in low_coords:
        c = np.asarray(c)
        assert c.size == 0 or int(c.max()) < 256, "downsampled xyzQ should be less than 8 bit"
        parts.append(c.astype(np.uint8).tobytes())
    parts.append(np.asarray(mins, dtype=np.int32).reshape(-1, 3).tobytes())
    return pack_bitstream(parts)


def unpack_low_xyz(buf: bytes):
    """dec_all_frame_low_xyz (test_utils.py:299-312)."""
    parts = unpack_bitstream(buf)
    mins = np.frombuffer(parts.pop(), dtype=np.int32).reshape(-1, 3)
    return [np.frombuffer(p, dtype=np.uint8).reshape(-1, 3).astype(np.int32) for p in parts], mins
