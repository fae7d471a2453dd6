"""GPU parity of the staged (bulk-copy / TMA) weight-gradient kernel and of the tables it reads.

`linr_rows.d_tile_rng` + `d_pair_cnt` / `d_pair_list` switch the 8->8 weight gradient to conv27_bwd_w3_kernel
(csrc/net_kernels.cuh): dy rows, neighbour row ranges and pair lists staged in shared memory by bulk copies, one
(dy row, neighbour row) pair per lane.  The tables are integer work: bit-exact against a numpy restatement.  The
gradient changes its summation schedule (compacted pairs), so it is compared within fp32 tolerance against the
lane = row kernel and float64, and must stay bitwise reproducible run to run.  Forward results do not depend on the
tables at all (the encoder / decoder reproducibility contract): checked bit for bit."""
import dataclasses

import numpy as np
import pytest
import torch

from test_gpu_parity import L, O, _cuda, _load, _dense_from_compact, _net_case, RTOL  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


def _no_ranges(t):
    return dataclasses.replace(t, tile_rng=None, pair_cnt=None, pair_list=None, _rows=None)


def _ranges_from_dense(nb, n):
    """numpy restatement of linr_tile_ranges on the dense [n,27] table."""
    nt = (n + 127) // 128
    out = np.zeros((nt, 6), dtype=np.int32)
    for t in range(nt):
        blk = nb[t * 128:(t + 1) * 128]
        for d in range(3):
            v = blk[:, [k for k in range(27) if k % 3 == d]]
            v = v[v >= 0]
            if len(v):
                out[t, 2 * d], out[t, 2 * d + 1] = v.min(), v.max() + 1
    return out


@pytest.mark.parametrize("name", ["tiny", "ragged", "mid"])
def test_tile_ranges_bit_exact(L, name):
    g = _load(f"int_{name}.npz")
    fr = L.frame.prepare_frame(_cuda(g["points"]), None, 64)
    t = fr.tables
    nb = _dense_from_compact(t.anchor, t.mask, t.n_rows)
    assert t.tile_rng is not None and t.tile_rng.shape == ((t.n_rows + 127) // 128, 6)
    np.testing.assert_array_equal(t.tile_rng.cpu().numpy(), _ranges_from_dense(nb, t.n_rows))


def _pair_order(rl, nbv):
    """Order of the entries inside a pair list (coords.cu pair_list_kernel): entries are classed by (row mod 4, neighbour
    mod 4); diagonal d = (neighbour - row) mod 4 gives g[d] = min class count groups of four entries with distinct rows and
    distinct neighbours modulo 4 (the shared-memory bank condition of the weight-gradient kernel), diagonals one after
    the other; what is left follows in row order.  Returns the permutation and the number of grouped entries."""
    a, b = rl & 3, nbv & 3
    d = (b - a) & 3
    rank = np.zeros(len(rl), np.int64)
    ncls = np.zeros((4, 4), np.int64)
    for i in range(len(rl)):
        rank[i] = ncls[a[i], b[i]]
        ncls[a[i], b[i]] += 1
    g = [min(ncls[x, (x + dd) & 3] for x in range(4)) for dd in range(4)]
    gb = np.concatenate([[0], np.cumsum(g)])
    head = np.full(4 * gb[4], -1, np.int64)
    tail = []
    for i in range(len(rl)):
        if rank[i] < g[d[i]]:
            head[4 * (gb[d[i]] + rank[i]) + a[i]] = i
        else:
            tail.append(i)
    return np.concatenate([head, np.array(tail, np.int64)]).astype(np.int64), int(4 * gb[4])


@pytest.mark.parametrize("name", ["tiny", "ragged", "mid"])
def test_pair_lists_bit_exact(L, name):
    """linr_pair_lists: list (t,k) = the rows of tile t with a neighbour at offset k and that neighbour, in the
    bank-conflict-avoiding order of _pair_order."""
    g = _load(f"int_{name}.npz")
    fr = L.frame.prepare_frame(_cuda(g["points"]), None, 64)
    t = fr.tables
    n = t.n_rows
    nb = _dense_from_compact(t.anchor, t.mask, n)
    cnt = t.pair_cnt.cpu().numpy()
    lst = t.pair_list.cpu().numpy().view(np.uint32)
    assert cnt.shape == ((n + 255) // 256, 32) and lst.shape == ((n + 255) // 256, 27 * 256)
    import ctypes
    order = (ctypes.c_int32 * 28)()
    L.lib.load().linr_pair_list_order(order)
    order = np.array(order).reshape(2, 14)
    assert sorted(order.reshape(-1).tolist()) == list(range(28))        # every offset once, 27 = the bias slot (no list)
    grouped = total = 0
    for ti in range(cnt.shape[0]):
        blk = nb[ti * 256:(ti + 1) * 256]
        for h in range(2):
            start = h * 14 * 256                                        # half h of the tile's storage, lists back to back
            for k in order[h]:
                if k == 27:
                    continue
                rl = np.nonzero(blk[:, k] >= 0)[0]
                assert cnt[ti, k] == len(rl)
                perm, nh = _pair_order(rl, blk[rl, k])
                want = (rl[perm].astype(np.uint32) << np.uint32(24)) | blk[rl[perm], k].astype(np.uint32)
                np.testing.assert_array_equal(lst[ti, start:start + len(rl)], want)
                got = lst[ti, start:start + nh].reshape(-1, 4)
                assert (np.sort((got >> np.uint32(24)) & 3, axis=1) == np.arange(4)).all()     # rows distinct modulo 4
                assert (np.sort(got & np.uint32(3), axis=1) == np.arange(4)).all()             # neighbours too
                grouped, total = grouped + nh, total + len(rl)
                start += (len(rl) + 3) & ~3                             # every list starts on a multiple of 4 entries
        assert (cnt[ti, 27:] == 0).all()
    if name == "mid":
        assert grouped >= 0.6 * total


def _frames(L):
    g = _load("int_mid.npz")
    small = L.frame.prepare_frame(_cuda(g["points"]), None, 64)
    big = L.frame.prepare_frame(L.synth.make_sequence("loot", 1, device="cuda")[0], None, 64)
    return [("mid", small), ("loot", big)]


def test_weight_gradient_v3_matches_v2_and_is_reproducible(L):
    gen = torch.Generator(device="cuda").manual_seed(4)
    for name, fr in _frames(L):
        t, t0 = fr.tables, _no_ranges(fr.tables)
        n = t.n_rows
        for cin, cout in ((8, 8), (8, 4), (4, 4)):
            x = torch.randn(n, cin, generator=gen, device="cuda")
            dy = torch.randn(n, cout, generator=gen, device="cuda")
            dW, db = L.net.spconv27_bwd_w(x, dy, t)
            dW2, db2 = L.net.spconv27_bwd_w(x, dy, t)
            assert torch.equal(dW, dW2) and torch.equal(db, db2), (name, cin, cout)      # run to run bitwise
            dW0, db0 = L.net.spconv27_bwd_w(x, dy, t0)
            sc = dW0.abs().max().item()
            assert (dW - dW0).abs().max().item() <= 2e-5 * sc, (name, cin, cout)         # fp32 re-association only
            assert (db - db0).abs().max().item() <= 2e-5 * max(1.0, db0.abs().max().item())
            # against float64 on the dense table
            if name == "mid":
                nb = torch.from_numpy(_dense_from_compact(t.anchor, t.mask, n)).cuda().long()
                xd = torch.cat([x.double(), torch.zeros(1, cin, device="cuda", dtype=torch.float64)])
                ref = torch.stack([xd[nb[:, k]].T @ dy.double() for k in range(27)])
                assert (dW.double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


@pytest.mark.parametrize("env", [{"LINR_BW3_MASK": "15"}, {"LINR_BW3_MASK": "15", "LINR_BW3_NOX": "1"}, {"LINR_BW3_MASK": "0"}])
def test_weight_gradient_v3_switches(L, env):
    """Every class on the staged kernel (4->4 included), the un-staged neighbour path of a tile whose ranges do not fit,
    and the kernel switched off: same gradients up to fp32 re-association, bitwise reproducible run to run."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tests", "_bw3_variants_driver.py")], capture_output=True, text=True,
                       timeout=600, cwd=root, env={**os.environ, **env})
    assert p.returncode == 0, p.stderr[-3000:]
    r = json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    assert r["repro"]
    assert r["conv"] <= 2e-5 and r["net"] <= 5e-5, r


def test_network_staged_vs_gathered(L, O):
    """Whole network: probabilities / CDFs bit-identical, gradient within fp32 re-association, training forward too."""
    g, S, sd, flat, fr = _net_case(L, O)
    cases = [(fr, flat.cuda())]
    big = L.frame.prepare_frame(L.synth.make_sequence("mvub10", 1, device="cuda")[0], None, 64)
    from linr_pcgc_b200 import params as P
    cases.append((big, P.init_flat(big.n_scales, 11).cuda()))
    for f, params in cases:
        t, t0 = f.tables, _no_ranges(f.tables)
        run = L.net.NetRunner(f.n_scales, t.n_rows, "cuda", train=True)
        outs = []
        for tab in (t, t0):
            o = run.forward(params, tab, want_probs=True, want_cdf=True, want_bits=True)
            probs, cdf = o["probs"].clone(), o["cdf"].clone()
            grad = torch.empty_like(params)
            run.forward(params, tab, train=True, loss_scale=1.0 / f.point_num)
            run.backward(params, tab, grad)
            outs.append((probs, cdf, grad.clone(), float(o["bits"].item())))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        assert abs(outs[0][3] - outs[1][3]) <= 1e-6 * abs(outs[1][3])
        ga, gb = outs[0][2], outs[1][2]
        assert torch.isfinite(ga).all()
        assert (ga - gb).abs().max().item() <= 5e-5 * gb.abs().max().item()
