"""One-file GOP container.

The reference writes a GOP as a directory: `side_info.json`, `bins/model.bin`, `bins/low_enc_bytes.bin` and one
`bins/frame%04d_scale%d.bin` per frame and scale (encoder.py:13-18,84-146) -- 200+ small files for a 32-frame GOP.
This module packs the same payloads, untouched, into one file with a fixed little-endian header and an index, and
converts both ways (`pipeline.write_gop` / `read_gop` keep the directory layout), so a container made here can be
unpacked into exactly what the reference's decoder reads.

    magic   8 B   b"LINRGOP\\0"
    u16 container_version (1)    u16 cdf_version (1: the reference's Laplace row [c1..cL,0]; 2: [0,c1..cL])
    u16 scale_num   u16 bitdepth   u16 enc_mode   u16 reserved   u32 n_frames   u32 n_payloads
    f64 mu, b, min_param, max_param, model_bits
    u32 point_num[n_frames]      u16 n_scales[n_frames]  (a frame may have fewer scales than scale_num)
    u64 offset[n_payloads], u64 length[n_payloads]   payload order: model, low_xyz, then frame-major, scale-minor
    payload bytes
"""
from __future__ import annotations

import struct
from typing import List

import numpy as np

from .pipeline import EncodedGop

MAGIC = b"LINRGOP\0"
VERSION = 1
_HEAD = struct.Struct("<8sHHHHHHII5d")


def pack(enc: EncodedGop) -> bytes:
    side = enc.side_info
    payloads: List[bytes] = [bytes(enc.model_bytes), bytes(enc.low_enc_bytes)]
    n_scales = []
    for fb in enc.frame_bytes:
        n_scales.append(len(fb))
        payloads += [bytes(b) for b in fb]
    head = _HEAD.pack(MAGIC, VERSION, int(side.get("cdf_version", 1)), int(enc.scale_num), int(side["bitdepth"]), int(side["enc_mode"]), 0,
                      len(enc.frame_bytes), len(payloads), float(side["mu"]), float(side["b"]), float(side["min_param"]),
                      float(side["max_param"]), float(enc.model_bits))
    meta = np.asarray(enc.point_nums, dtype="<u4").tobytes() + np.asarray(n_scales, dtype="<u2").tobytes()
    lens = np.asarray([len(p) for p in payloads], dtype="<u8")
    base = len(head) + len(meta) + 16 * len(payloads)
    offs = base + np.concatenate([[0], np.cumsum(lens)[:-1]]).astype("<u8")
    return head + meta + offs.astype("<u8").tobytes() + lens.tobytes() + b"".join(payloads)


def unpack(buf: bytes) -> EncodedGop:
    if len(buf) < _HEAD.size or buf[:8] != MAGIC:
        raise ValueError("not a LINR GOP container")
    (_, ver, cdf_version, scale_num, bitdepth, enc_mode, _, n_frames, n_payloads, mu, b, mn, mx, model_bits) = _HEAD.unpack_from(buf, 0)
    if ver != VERSION:
        raise ValueError(f"container version {ver} is not supported (this reader knows {VERSION})")
    o = _HEAD.size
    point_nums = np.frombuffer(buf, dtype="<u4", count=n_frames, offset=o).astype(int).tolist()
    o += 4 * n_frames
    n_scales = np.frombuffer(buf, dtype="<u2", count=n_frames, offset=o).astype(int).tolist()
    o += 2 * n_frames
    offs = np.frombuffer(buf, dtype="<u8", count=n_payloads, offset=o)
    lens = np.frombuffer(buf, dtype="<u8", count=n_payloads, offset=o + 8 * n_payloads)
    if n_payloads != 2 + sum(n_scales) or (n_payloads and int(offs[-1] + lens[-1]) > len(buf)):
        raise ValueError("corrupt LINR GOP container (index does not match the file)")
    get = lambda i: bytes(buf[int(offs[i]): int(offs[i] + lens[i])])
    frame_bytes, i = [], 2
    for ns in n_scales:
        frame_bytes.append([get(i + s) for s in range(ns)])
        i += ns
    side = dict(mu=mu, b=b, min_param=mn, max_param=mx, enc_mode=enc_mode, bitdepth=bitdepth)
    if cdf_version != 1:
        side["cdf_version"] = cdf_version
    return EncodedGop(scale_num, side, get(0), model_bits, get(1), frame_bytes, point_nums)


def write(enc: EncodedGop, path: str) -> int:
    data = pack(enc)
    with open(path, "wb") as f:
        f.write(data)
    return len(data)


def read(path: str) -> EncodedGop:
    with open(path, "rb") as f:
        return unpack(f.read())
