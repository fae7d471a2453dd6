import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import linr_pcgc_b200
from linr_pcgc_b200 import pipeline, synth, codec, rc, params as P
dev = torch.device("cuda")
pts = synth.make_sequence("loot", 8, device=dev)
frames = pipeline.prepare_gop(pts, None, 64, dev)
S = frames[0].n_scales
flat = P.init_flat(S, 3).to(dev)
enc = pipeline.encode_gop(frames, flat, S)
t_rc = [0.0]; n_rc = [0]
orig = rc.decode_binary_into
def timed(*a):
    t0 = time.perf_counter(); orig(*a); t_rc[0] += time.perf_counter() - t0; n_rc[0] += a[2].shape[0]
rc.decode_binary_into = timed
for w in (1, 2, 4, 8, 12):
    pipeline.decode_gop(enc, dev, workers=w)
    torch.cuda.synchronize(); t_rc[0] = 0; n_rc[0] = 0
    t0 = time.perf_counter(); d = pipeline.decode_gop(enc, dev, workers=w); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"workers {w}: {dt/len(frames)*1e3:.1f} ms/frame; rc total {t_rc[0]*1e3:.0f} ms for {n_rc[0]/1e6:.1f} M symbols ({n_rc[0]/max(t_rc[0],1e-9)/1e6:.0f} Msym/s/thread-sum)")
# encode timing
run = pipeline.NetRunner(S, max(f.tables.n_rows for f in frames), dev, train=False)
for coders in (1, 2, 4):
    codec.encode_frames(run, flat, frames, coders=coders)
    torch.cuda.synchronize(); t0 = time.perf_counter(); codec.encode_frames(run, flat, frames, coders=coders); torch.cuda.synchronize()
    print(f"encode coders {coders}: {(time.perf_counter()-t0)/len(frames)*1e3:.2f} ms/frame")
torch.cuda.synchronize(); t0 = time.perf_counter()
for f in frames: run.forward(flat, f.tables, want_cdf=True, want_bits=False)
torch.cuda.synchronize(); print(f"forward only: {(time.perf_counter()-t0)/len(frames)*1e3:.2f} ms/frame")
cdf, occ, R = codec.frame_cdfs_to_host(run, flat, frames[0])
t0 = time.perf_counter(); b = rc.encode_binary(cdf[0, :frames[0].scale_off[1]].copy(), occ[:frames[0].scale_off[1]] & 1); dt = time.perf_counter() - t0
print(f"single-stream encode {frames[0].scale_off[1]} symbols: {dt*1e3:.2f} ms")
