"""Where the end-to-end (host buffers -> bitstreams) time of one GOP goes, outside the kernels.  GPU box only."""
import sys, time, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linr_pcgc_b200
from linr_pcgc_b200 import pipeline, synth, frame
from linr_pcgc_b200.trainer import GopTrainer

dev = torch.device("cuda", 0)
F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seq = synth.make_sequence("loot", F, device=dev)
host = [p.cpu().pin_memory() for p in seq]
torch.cuda.synchronize()

def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r

ms, d = t(lambda: [p.to(dev, non_blocking=True) for p in host]); print(f"H2D {F} frames: {ms:.2f} ms ({ms/F:.3f}/frame)")
ms, fr = t(lambda: pipeline.prepare_gop(d, None, 64, dev)); print(f"prepare_gop: {ms:.2f} ms ({ms/F:.3f}/frame)")
S = fr[0].n_scales
ms, _ = t(lambda: frame.sort_unique(d[0] - d[0].min(0).values, 10)); print(f"  sort_unique: {ms:.3f} ms")
ms, _ = t(lambda: frame.octree_down(fr[0].xyz, 10)); print(f"  octree_down(level 0): {ms:.3f} ms")
ms, _ = t(lambda: frame.build_tables(fr[0].tables.coords, fr[0].tables.scale, fr[0].tables.occ)); print(f"  build_tables: {ms:.3f} ms")
mr = max(f.tables.n_rows for f in fr)
ms, tr = t(lambda: GopTrainer(S, dev, seed=1, max_rows=mr)); print(f"GopTrainer(): {ms:.2f} ms")
ms, _ = t(lambda: tr.fit(fr, 1), 2); print(f"fit 1 epoch: {ms:.2f} ms ({ms/F:.3f}/frame-iter)")
ms, enc = t(lambda: pipeline.encode_gop(fr, tr.state.params, S, 8), 2); print(f"encode_gop: {ms:.2f} ms ({ms/F:.3f}/frame)")
from linr_pcgc_b200 import model_compression
ms, _ = t(lambda: model_compression.compress_model(tr.state.params, 8)); print(f"  compress_model: {ms:.2f} ms")
ms, _ = t(lambda: [f.scale_coords(f.n_scales - 1).cpu().numpy() for f in fr]); print(f"  low coords to host: {ms:.2f} ms")

# whole call vs its parts
E = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ms_full, _ = t(lambda: pipeline.overfit_encode_gop(host, E, device=dev, seed=1)[0], 2)
ms_fit, _ = t(lambda: tr.fit(fr, E), 2)
ms_enc, _ = t(lambda: pipeline.encode_gop(fr, tr.state.params, S, 8), 2)
ms_prep, _ = t(lambda: pipeline.prepare_gop(host, None, 64, dev), 2)
print(f"overfit_encode_gop(host, {E} epochs): {ms_full:.1f} ms;  parts: prepare(host) {ms_prep:.1f} + fit {ms_fit:.1f} + encode {ms_enc:.1f} = {ms_prep + ms_fit + ms_enc:.1f}")
hp = torch.cuda.Stream(dev, priority=-1)
torch.cuda.synchronize()
torch.cuda.set_stream(hp)
ms_hp, _ = t(lambda: pipeline.overfit_encode_gop(host, E, device=dev, seed=1)[0], 2)
print(f"serial, high-priority main stream: {ms_hp:.1f} ms")
coder = pipeline.GopCoder(dev)
def piped():
    f, st, _ = pipeline.overfit_encode_gop(host, E, device=dev, seed=1, coder=coder)
    return f
ms_p, _ = t(lambda: piped(), 3)
coder.collect()
print(f"with GopCoder (steady state): {ms_p:.1f} ms")
