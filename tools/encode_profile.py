"""Where the encode time of a trained GOP goes: forward only, forward + D2H, host coder alone, and the pipelined
codec.encode_frames for several (coders, depth) settings.  python tools/encode_profile.py [frames] [epochs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import linr_pcgc_b200  # noqa: F401
from linr_pcgc_b200 import codec, pipeline, rc, synth
from linr_pcgc_b200.net import NetRunner
from linr_pcgc_b200.trainer import GopTrainer

F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
E = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda")
pts = synth.make_sequence("loot", F, device=dev)
frames = pipeline.prepare_gop(pts, None, 64, dev)
S = frames[0].n_scales
mr = max(f.tables.n_rows for f in frames)
tr = GopTrainer(S, dev, seed=8807, max_rows=mr)
tr.fit(frames, E)
flat = tr.state.params
run = NetRunner(S, mr, dev, train=False)


def wall(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n / len(frames) * 1e3


print(f"forward only            : {wall(lambda: [run.forward(flat, f.tables, want_cdf=True, want_bits=False) for f in frames]):.2f} ms/frame")
print(f"forward + D2H (serial)  : {wall(lambda: [codec.frame_cdfs_to_host(run, flat, f) for f in frames]):.2f} ms/frame")
cdf, occ, R = codec.frame_cdfs_to_host(run, flat, frames[0])
cdf, occ = cdf.copy(), occ.copy()
f0 = frames[0]
cdfs, syms, shifts = [], [], []
for s in range(f0.n_scales):
    a, b = f0.scale_off[s], f0.scale_off[s + 1]
    for k in range(8):
        cdfs.append(cdf[k, a:b]), syms.append(occ[a:b]), shifts.append(k)
for th in (1, 4, 8, 16):
    t0 = time.perf_counter()
    for _ in range(5):
        st = rc.encode_binary_batch(cdfs, syms, shifts, th)
    dt = (time.perf_counter() - t0) / 5
    print(f"host coder, {th:2d} threads   : {dt * 1e3:.2f} ms/frame ({8 * R / dt / 1e6:.0f} Msym/s), {sum(len(x) for x in st)} bytes")
for coders, depth in ((1, 2), (2, 4), (3, 4), (4, 6), (4, 8)):
    print(f"encode_frames coders {coders} depth {depth}: {wall(lambda: codec.encode_frames(run, flat, frames, coders=coders, depth=depth)):.2f} ms/frame")
print(f"pipeline.encode_gop     : {wall(lambda: pipeline.encode_gop(frames, flat, S, 8, runner=run)):.2f} ms/frame")
