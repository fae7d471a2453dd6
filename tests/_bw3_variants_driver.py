"""Subprocess body of tests/test_gpu_staged.py::test_weight_gradient_v3_switches: the environment switches of the staged
weight-gradient kernel are read once per process (LINR_BW3_MASK: which classes run it; LINR_BW3_NOX=1: neighbour rows
gathered through L1 instead of staged -- the path a tile takes when its neighbour ranges do not fit the staging area).
Prints one JSON line of maximal relative errors against the lane = row kernels (tables without ranges / pair lists)."""
import dataclasses
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linr_pcgc_b200 import frame, net, synth  # noqa: E402
from linr_pcgc_b200 import params as P  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "int_mid.npz"))
frames = [frame.prepare_frame(torch.from_numpy(g["points"]).cuda(), None, 64),
          frame.prepare_frame(synth.make_sequence("mvub10", 1, device="cuda")[0], None, 64)]
gen = torch.Generator(device="cuda").manual_seed(9)
res = {"conv": 0.0, "net": 0.0, "repro": True}
for fr in frames:
    t = fr.tables
    t0 = dataclasses.replace(t, tile_rng=None, pair_cnt=None, pair_list=None, _rows=None)
    n = t.n_rows
    for cin, cout in ((8, 8), (8, 4), (4, 4)):
        x = torch.randn(n, cin, generator=gen, device="cuda")
        dy = torch.randn(n, cout, generator=gen, device="cuda")
        dW, db = net.spconv27_bwd_w(x, dy, t)
        dW2, db2 = net.spconv27_bwd_w(x, dy, t)
        res["repro"] = res["repro"] and bool(torch.equal(dW, dW2) and torch.equal(db, db2))
        dW0, db0 = net.spconv27_bwd_w(x, dy, t0)
        res["conv"] = max(res["conv"], (dW - dW0).abs().max().item() / dW0.abs().max().item(),
                          (db - db0).abs().max().item() / max(1.0, db0.abs().max().item()))
    prm = P.init_flat(fr.n_scales, 11).cuda()
    run = net.NetRunner(fr.n_scales, n, "cuda", train=True)
    grads = []
    for tab in (t, t0):
        grad = torch.empty_like(prm)
        run.forward(prm, tab, train=True, loss_scale=1.0 / fr.point_num)
        run.backward(prm, tab, grad)
        grads.append(grad.clone())
    res["net"] = max(res["net"], (grads[0] - grads[1]).abs().max().item() / grads[1].abs().max().item())
    run.close()
print("RESULT " + json.dumps(res))
