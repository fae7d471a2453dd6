"""Frame coding on top of the C ABI: teacher-forced encode (8 streams per scale) and sequential decode.

Host-side mirror of `encode_one_frame` (encoder.py:158-203) / `CNP.encode` (models/upsample.py:219-246) and
`decode_one_frame` (decoder.py:153-176) / `CNP.decode` (models/upsample.py:249-295).  The GPU emits the 16-bit CDF
boundaries torchac would derive from the float probabilities; only those 2 B/symbol cross PCIe, and the serial range
coder (csrc/rc_host.cpp) runs on the host over the independent streams in parallel.
"""
from __future__ import annotations

import os
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
        """Contiguous pinned views [8, rows] int16 and [rows] uint8 (a strided D2H copy would bounce through a
        temporary and synchronise)."""
        if self.cdf is None or self.cdf.numel() < 8 * rows:
            cap = max(rows, 1)
            self.cdf = torch.empty(8 * cap, dtype=torch.int16).pin_memory()
            self.occ = torch.empty(cap, dtype=torch.uint8).pin_memory()
        return self.cdf[: 8 * rows].view(8, rows), self.occ[:rows]


_pinned = _Pinned()


def frame_cdfs_to_host(runner: NetRunner, params: torch.Tensor, frame: Frame):
    """Forward (no grad) + D2H of the CDF boundaries and the occupancy bytes -> numpy views [8,R] u16, [R] u8."""
    t = frame.tables
    R = t.n_rows
    out = runner.forward(params, t, train=False, want_cdf=True, want_bits=False)
    h_cdf, h_occ = _pinned.get(R)
    h_cdf.copy_(out["cdf"], non_blocking=True)
    h_occ.copy_(t.occ, non_blocking=True)
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


def encode_frames(runner: NetRunner, params: torch.Tensor, frames: Sequence[Frame], threads: Optional[int] = None,
                  depth: Optional[int] = None, coders: Optional[int] = None) -> List[List[bytes]]:
    """Encode several frames with the GPU and the host coder overlapped: while the range coder works on frame i
    (C code, GIL released), the network forward and the CDF download of frame i+1 are already running.  `depth` pinned
    staging sets bound the frames in flight, `coders` frames are range-coded at the same time (the 8 stage streams of
    the finest scale hold 3/4 of a frame's symbols, so one frame cannot use more than ~8 cores).  Defaults (measured,
    tools/encode_profile.py, 16 host cores: 1.45 ms/frame with 2 coders x 8 threads, 1.31 with 3 x 8): three frames in the
    coder when the host has >= 12 cores, else two; up to 8 threads each.  Same bytes as `encode_frame` frame by frame."""
    from concurrent.futures import ThreadPoolExecutor
    if not frames:
        return []
    cores = rc.host_cores()
    if coders is None:
        coders = 3 if cores >= 12 else 2
    if depth is None:
        depth = 2 * coders
    if threads is None:
        threads = max(2, min(8, cores // 2))
    cap = max(f.tables.n_rows for f in frames)
    sets = [(torch.empty(8 * max(cap, 1), dtype=torch.int16).pin_memory(), torch.empty(max(cap, 1), dtype=torch.uint8).pin_memory())
            for _ in range(min(depth, len(frames)))]
    events = [torch.cuda.Event() for _ in sets]
    pending = [None] * len(sets)
    out: List[Optional[List[bytes]]] = [None] * len(frames)

    def code(i: int, slot: int):
        f = frames[i]
        events[slot].synchronize()
        R = f.tables.n_rows
        cdf, occ = sets[slot][0][: 8 * R].view(8, R).numpy().view(np.uint16), sets[slot][1][:R].numpy()
        cdfs, syms, shifts = [], [], []
        for s in range(f.n_scales):
            a, b = f.scale_off[s], f.scale_off[s + 1]
            for k in range(8):
                cdfs.append(cdf[k, a:b])
                syms.append(occ[a:b])
                shifts.append(k)
        streams = rc.encode_binary_batch(cdfs, syms, shifts, threads)
        return [pack_bitstream(streams[8 * s: 8 * s + 8]) for s in range(f.n_scales)]

    with ThreadPoolExecutor(max_workers=max(1, coders)) as pool:
        for i, f in enumerate(frames):
            slot = i % len(sets)
            if pending[slot] is not None:           # the staging set is free once its frame has been coded
                j, fut = pending[slot]
                out[j] = fut.result()
            R = f.tables.n_rows
            res = runner.forward(params, f.tables, train=False, want_cdf=True, want_bits=False)
            sets[slot][0][: 8 * R].view(8, R).copy_(res["cdf"], non_blocking=True)   # contiguous: one async memcpy
            sets[slot][1][:R].copy_(f.tables.occ, non_blocking=True)
            events[slot].record()
            pending[slot] = (i, pool.submit(code, i, slot))
        for p in pending:
            if p is not None:
                out[p[0]] = p[1].result()
    return out  # type: ignore[return-value]


class _DecodeCtx:
    """Per-thread decode state: runner + pinned staging buffers, reused across scales and frames."""

    def __init__(self, scale_num: int, device, runner: Optional[NetRunner] = None):
        self.runner = runner if runner is not None else NetRunner(scale_num, 1, device, train=False)
        self.cap = 0
        self.h_cdf = self.h_sym = None

    def reserve(self, n: int):
        if n > self.cap:
            self.cap = max(n, 2 * self.cap, 1024)
            self.h_cdf = torch.empty(self.cap, dtype=torch.int16).pin_memory()
            self.h_sym = torch.empty(self.cap, dtype=torch.uint8).pin_memory()


def decode_scale(runner: NetRunner, params: torch.Tensor, coords: torch.Tensor, scale_idx: int, data: bytes,
                 ctx: Optional[_DecodeCtx] = None):
    """One scale: parents `coords` (sorted unique, CUDA int32 [N,3]) + its bitstream -> uint8 occupancy [N] (CUDA).
    The 8 stages are strictly sequential (CNP.decode, models/upsample.py:249-295): stage k's network input contains
    the bits decoded in stages < k."""
    n = int(coords.shape[0])
    dev = coords.device
    if ctx is None:
        ctx = _DecodeCtx(runner.S, dev, runner)
    ctx.reserve(n)
    scale = torch.full((n,), scale_idx, dtype=torch.uint8, device=dev)
    occ = torch.zeros(n, dtype=torch.uint8, device=dev)
    t = build_tables(coords, scale, occ, tile_ranges=False)   # forward only: the gradient kernels' tables are not needed
    streams = unpack_bitstream(data)
    d_sym = torch.empty(n, dtype=torch.uint8, device=dev)
    # the eight device<->host round trips of the scale run inside one C call (no interpreter lock held), so the frames
    # decoded by other host threads overlap with this one
    runner.decode_scale(params, t, streams, d_sym, ctx.h_cdf, ctx.h_sym)
    return occ, t


def decode_frame(runner: NetRunner, params: torch.Tensor, all_bytes: Sequence[bytes], low_coords: torch.Tensor,
                 low_bits: int = 8, ctx: Optional[_DecodeCtx] = None) -> torch.Tensor:
    """decode_one_frame (decoder.py:153-176): coarse-to-fine; returns the full-resolution sorted coords (CUDA int32)."""
    cur = low_coords
    bits = max(low_bits, 1)
    for s in range(len(all_bytes) - 1, -1, -1):
        occ, _ = decode_scale(runner, params, cur, s, all_bytes[s], ctx)
        bits += 1
        cur = octree_up(cur, occ, bits)
    return cur


_ctx_pool: dict = {}   # (device, scale_num) -> idle decode contexts (workspaces are large: keep them across calls)


def decode_frames(params: torch.Tensor, scale_num: int, jobs: Sequence, workers: int = 8) -> List[torch.Tensor]:
    """Decode independent frames concurrently (frames of a GOP only share the model): one host thread, CUDA stream
    and runner per in-flight frame, so the 56 device<->host round trips of a frame overlap with those of the others
    and the serial range decoders run on different cores.  jobs: (all_bytes, low_coords CUDA int32 [N,3])."""
    from concurrent.futures import ThreadPoolExecutor
    import threading
    if not jobs:
        return []
    dev = params.device
    workers = max(1, min(workers, len(jobs)))
    local = threading.local()
    key = (str(dev), scale_num)
    idle = _ctx_pool.setdefault(key, [])
    lock = threading.Lock()
    used = []
    main_stream = torch.cuda.current_stream(dev)
    ready = torch.cuda.Event()
    ready.record(main_stream)

    def run(job):
        if not hasattr(local, "ctx"):
            torch.cuda.set_device(dev)
            with lock:
                local.ctx = idle.pop() if idle else None
            if local.ctx is None:
                local.ctx = _DecodeCtx(scale_num, dev)
                local.ctx.stream = torch.cuda.Stream(dev)
            local.stream = local.ctx.stream
            with lock:
                used.append(local.ctx)
        with torch.cuda.stream(local.stream):
            local.stream.wait_event(ready)
            out = decode_frame(local.ctx.runner, params, job[0], job[1], ctx=local.ctx)
            local.stream.synchronize()
        return out

    try:
        if workers == 1:
            return [run(j) for j in jobs]
        with ThreadPoolExecutor(max_workers=workers) as pool:
            return list(pool.map(run, jobs))
    finally:
        idle.extend(used)


_POPC = {}


def decode_frames_batched(params: torch.Tensor, scale_num: int, jobs: Sequence, max_batch: int = 8, workers: Optional[int] = None,
                          threads: Optional[int] = None) -> List[torch.Tensor]:
    """Decode independent frames in LOCKSTEP (frames of a GOP only share the model): at every scale the parents of all
    frames of a batch are concatenated -- frame f shifted by f * stride in x, so the frames stay sorted, unique and out
    of each other's 3x3x3 neighbourhoods -- and every one of the 8 stages is one set of launches over all of them, one
    CDF download, one range decoder per frame on its own host thread, one symbol upload (linr_net_decode_scale_batch).
    A batch makes 56 launch sets and device<->host round trips instead of 56 per frame, and the launches are large
    enough to fill the GPU (a per-frame stage of a coarse scale is a handful of blocks).  `workers` batches run at the
    same time on their own streams and host threads, so the network passes of one overlap the range decoding of the
    others: four in flight when the host has the cores (measured, tools/decode_profile.py: 32 loot frames 3.40 ms/frame
    with 2 batches of 8 in flight, 2.29 with 4; 16 MVUB-shaped frames 2.31 with 2 x 8, 2.07 with 4 x 4), and batches no
    larger than it takes to have that many.  Bit-identical CDFs: a row sees the same neighbours in the same order as in
    its own frame.
    jobs: (all_bytes, low_coords CUDA int32 [N,3])."""
    from concurrent.futures import ThreadPoolExecutor
    import threading
    if not jobs:
        return []
    dev = params.device
    low_max = max(int(j[1].max().item()) if j[1].numel() else 0 for j in jobs)
    s_max = max(len(j[0]) for j in jobs)
    low_stride = 1 << max(1, (low_max + 2 - 1).bit_length())                  # >= low_max + 2: a gap of at least one voxel
    fit = max(1, (1 << 20) // (low_stride << s_max))                          # coordinates stay below 2^20 at the finest level
    if workers is None:
        workers = min(4, max(2, rc.host_cores() // 4))
    B = max(1, min(max_batch, fit, -(-len(jobs) // workers)))
    batches = [list(range(b0, min(b0 + B, len(jobs)))) for b0 in range(0, len(jobs), B)]
    workers = max(1, min(workers, len(batches)))
    threads = threads or max(B, rc.host_cores() // workers)
    if dev not in _POPC:
        _POPC[dev] = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=dev)
    popc = _POPC[dev]
    out: List[Optional[torch.Tensor]] = [None] * len(jobs)
    main_stream = torch.cuda.current_stream(dev)
    ready = torch.cuda.Event()
    ready.record(main_stream)

    key = (str(dev), scale_num)
    idle = _ctx_pool.setdefault(key, [])          # decode contexts (workspace, pinned staging, stream) are kept across calls
    lock = threading.Lock()

    def run(worker: int, mine):
        torch.cuda.set_device(dev)
        with lock:
            ctx = idle.pop() if idle else None
        if ctx is None:
            ctx = _DecodeCtx(scale_num, dev)
            ctx.stream = torch.cuda.Stream(dev)
        try:
            with torch.cuda.stream(ctx.stream):
                ctx.stream.wait_event(ready)
                for idx in mine:
                    _decode_batch(ctx, params, [jobs[i] for i in idx], s_max, low_stride, popc, threads, out, idx)
                ctx.stream.synchronize()
        finally:
            with lock:
                idle.append(ctx)

    if workers == 1:
        run(0, batches)
    else:
        with ThreadPoolExecutor(max_workers=workers) as pool:
            list(pool.map(lambda w: run(w, batches[w::workers]), range(workers)))
    return out  # type: ignore[return-value]


def _decode_batch(ctx: _DecodeCtx, params, batch, s_max: int, low_stride: int, popc, threads: int, out, idx):
    dev = params.device
    runner = ctx.runner
    F = len(batch)
    cur: List[Optional[torch.Tensor]] = [None] * F
    for s in range(s_max - 1, -1, -1):
        stride = low_stride << (s_max - 1 - s)                              # doubles with every level
        for f, (ab, low) in enumerate(batch):
            if len(ab) - 1 == s:
                cur[f] = low.to(torch.int32)                                # this frame's coarsest coded scale
        act = [f for f in range(F) if cur[f] is not None]
        if not act:
            continue
        seg = [0]
        for f in act:
            seg.append(seg[-1] + int(cur[f].shape[0]))
        n = seg[-1]
        fid = torch.repeat_interleave(torch.tensor(act, dtype=torch.int32, device=dev),
                                      torch.tensor([seg[i + 1] - seg[i] for i in range(len(act))], device=dev))
        coords = torch.cat([cur[f] for f in act], dim=0)
        coords[:, 0] += fid * stride
        scale = torch.full((n,), s, dtype=torch.uint8, device=dev)
        occ = torch.zeros(n, dtype=torch.uint8, device=dev)
        t = build_tables(coords, scale, occ, tile_ranges=False)
        d_sym = torch.empty(n, dtype=torch.uint8, device=dev)
        ctx.reserve(n)                                                      # pinned staging grows geometrically, reused across scales
        streams = [unpack_bitstream(batch[f][0][s]) for f in act]
        runner.decode_scale_batch(params, t, seg, streams, d_sym, ctx.h_cdf, ctx.h_sym, threads)
        bits = max(1, int(F * stride * 2 - 1).bit_length())
        child = octree_up(coords, occ, bits)                                # all frames at once: offsets double with the coordinates
        ccount = torch.cumsum(popc[occ.long()], dim=0)
        ends = ccount[torch.tensor([e - 1 for e in seg[1:]], device=dev)].cpu().tolist()
        starts = [0] + [int(e) for e in ends[:-1]]
        for i, f in enumerate(act):
            c = child[starts[i]: int(ends[i])].clone()
            c[:, 0] -= f * stride * 2
            cur[f] = c
    for f in range(F):
        out[idx[f]] = cur[f]


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
