"""Decode throughput of an encoded GOP for several (frames per lockstep batch, batches in flight) settings of
codec.decode_frames_batched.  python tools/decode_profile.py [shape] [frames] [epochs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linr_pcgc_b200  # noqa: F401
from linr_pcgc_b200 import codec, model_compression, params as P, pipeline, synth
from linr_pcgc_b200.trainer import GopTrainer

shape = sys.argv[1] if len(sys.argv) > 1 else "loot"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 32
E = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda")
pts = synth.make_sequence(shape, F, device=dev)
frames = pipeline.prepare_gop(pts, None, 64, dev)
S = frames[0].n_scales
tr = GopTrainer(S, dev, seed=8807, max_rows=max(f.tables.n_rows for f in frames))
tr.fit(frames, E)
enc = pipeline.encode_gop(frames, tr.state.params, S, 8)
n = P.offsets(P.param_spec(enc.scale_num))[-1]
d = dict(enc.side_info)
d["final_bytes"] = enc.model_bytes
flat = model_compression.decompress_model(d, n, dev)
lows, mins = codec.unpack_low_xyz(enc.low_enc_bytes)
jobs = [(fb, torch.from_numpy(lows[i]).to(dev)) for i, fb in enumerate(enc.frame_bytes)]
sweep = ((8, 2), (8, 3), (8, 4), (4, 4), (16, 2), (6, 3), (4, 6), (11, 3))
if len(sys.argv) > 4:
    sweep = tuple(tuple(int(v) for v in a.split("x")) for a in sys.argv[4].split(","))
for mb, wk in sweep:
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dec = codec.decode_frames_batched(flat, enc.scale_num, jobs, max_batch=mb, workers=wk)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    ok = all(bool((x + torch.from_numpy(mins[i].copy()).to(dev) == pts[i]).all()) for i, x in enumerate(dec))
    print(f"{shape} {F} frames: batch {mb:2d} x {wk} in flight: {best / F * 1e3:.2f} ms/frame lossless={ok}")
