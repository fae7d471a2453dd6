"""GPU: the reference-facing host API (LINR_PCGC_Model, GOP pipeline, CLI) on top of the C ABI, against the golden
fixtures recorded from the reference's own Python flow and the CPU oracle."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import linr_pcgc_b200  # noqa: F401
    from linr_pcgc_b200 import _lib
    _lib.load()
    return True


def _golden_model(env):
    from linr_pcgc_b200.model import LINR_PCGC_Model
    g = np.load(os.path.join(GOLDEN, "net_tiny.npz"), allow_pickle=False)
    S = int(g["scale_num"])
    m = LINR_PCGC_Model({"scale_num": S, "in_channel": 7, "hidden_channel_conv": 8, "block_layers": 1, "outstage": 8, "instage": 1})
    m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")})
    return g, S, m.cuda()


def _scale_inputs(g):
    """The dicts main.py:457-475 hands to the model, rebuilt from the oracle's frame preparation."""
    from oracle import linr_oracle as O
    fr = O.prepare_frame(g["points"], None, 64)
    out = []
    for sc in fr["scales"]:
        occ = torch.from_numpy(sc["occ"].astype(np.float32)).cuda()
        out.append({"coord": torch.from_numpy(sc["coord"]).cuda(), "occ_lst": [occ[:, k:k + 1].contiguous() for k in range(8)],
                    "offset_tensor": torch.from_numpy(sc["nbr7"].astype(np.float32)).cuda(), "scale_idx": sc["scale_idx"]})
    return fr, out


def test_model_forward_backward_like_main_loop(env):
    g, S, model = _golden_model(env)
    fr, scales = _scale_inputs(g)
    opt = torch.optim.Adam(model.parameters(), lr=0.01, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)   # main.py:231-237
    model.train()
    bits = 0
    per_scale = []
    for d in scales:                      # overfit_one_frame, main.py:457-475
        b = model(d)
        per_scale.append(float(b))
        bits = bits + b
    loss = bits / fr["point_num"]
    loss.backward()
    for i, b in enumerate(per_scale):
        assert abs(b - float(g[f"s{i}_scale_bits"])) <= RTOL * float(g[f"s{i}_scale_bits"])
    assert abs(float(loss) - float(g["loss"])) <= RTOL * float(g["loss"])
    grad = model.flat.grad.cpu().numpy()
    ref = g["grad_flat"]
    assert np.abs(grad - ref).max() / np.abs(ref).max() < RTOL
    opt.step()                            # torch's own Adam on the flat parameter == the reference's per-tensor Adam
    # the first Adam step moves every weight by lr * g / (|g| + 1e-8): elements whose gradient is ~1e-8 amplify the
    # 1e-7-relative gradient difference, hence the absolute tolerance of 0.2 % of the step size lr = 0.01
    np.testing.assert_allclose(model.flat.detach().cpu().numpy(), g["flat_after_adam"], rtol=1e-5, atol=2e-5)


def test_model_encode_decode_codec_signatures(env):
    g, S, model = _golden_model(env)
    fr, scales = _scale_inputs(g)
    model.eval()
    tot = 0
    for i, d in enumerate(scales):
        core = model.logic_core(d)
        assert len(core["out_cls_list"]) == 8 and core["out_cls_list"][0].shape == (len(d["coord"]), 1)
        np.testing.assert_allclose(torch.cat(core["out_cls_list"], 1).cpu().numpy(), g[f"s{i}_probs"], rtol=RTOL, atol=1e-6)
        enc = model.encode(d)             # models/model_core.py:236
        assert set(enc) == {"enc_bytes", "bits", "x_low"} and enc["bits"] == 8 * len(enc["enc_bytes"])
        tot += enc["bits"]
        dec = model.decode({"enc_bytes": enc["enc_bytes"], "coord": d["coord"], "offset_tensor": d["offset_tensor"],
                            "scale_idx": d["scale_idx"]})          # models/model_core.py:268
        assert len(dec) == 8
        for k in range(8):
            assert dec[k].shape == d["occ_lst"][k].shape and bool((dec[k] == d["occ_lst"][k]).all())
        c = model.codec(d)                # models/model_core.py:169
        assert {"bits", "enc_bytes", "enc_time", "dec_time", "bits_t"} <= set(c)
        assert abs(c["bits_t"] - float(g[f"s{i}_scale_bits"])) <= RTOL * float(g[f"s{i}_scale_bits"])
        assert c["bits"] <= enc["bits"]   # one stream + no 36-byte header
    assert abs(tot - int(g["all_bit"])) <= 0.005 * int(g["all_bit"])


def test_gop_pipeline_overfit_encode_decode(env):
    from linr_pcgc_b200 import pipeline, synth
    pts = synth.make_sequence("tiny", 3)
    host = [p.pin_memory() for p in pts]
    enc, state, losses = pipeline.overfit_encode_gop(host, epochs=6, seed=3)
    assert len(losses) == 6 and losses[-1] < losses[0] * 0.8          # it learns
    assert state.step == 18 and len(enc.frame_bytes) == 3
    dec = pipeline.decode_gop(enc)
    for d, p in zip(dec, pts):
        assert d.shape == p.shape and bool((d.cpu() == p).all())      # decoder.py:140
    # bitstream accounting (test_utils.py:145-157)
    bits = sum(8 * len(b) for fb in enc.frame_bytes for b in fb) + enc.model_bits + 8 * len(enc.low_enc_bytes)
    assert abs(enc.bpp - bits / sum(int(p.shape[0]) for p in pts)) < 1e-12
    # a second GOP warm-started from the first (main.py:102-104): its first-epoch loss starts near the converged one
    pts2 = synth.make_sequence("tiny", 3, start=3)
    _, st2, l2 = pipeline.overfit_encode_gop([p.cuda() for p in pts2], epochs=1, state=state.clone())
    assert l2[0] < losses[0] and st2.step == 21


def test_background_gop_coder_matches_serial_encode(env):
    """GopCoder (coding of GOP g on a side stream + host threads while the caller keeps training) returns the same
    bytes as the serial encode of the same parameter snapshot, even though training continues on the live vector."""
    import torch
    from linr_pcgc_b200 import pipeline, synth
    from linr_pcgc_b200.trainer import GopTrainer
    pts = synth.make_sequence("tiny", 3)
    frames = pipeline.prepare_gop([p.cuda() for p in pts])
    S = frames[0].n_scales
    tr = GopTrainer(S, "cuda", seed=5, max_rows=max(f.tables.n_rows for f in frames))
    tr.fit(frames, 2)
    snap = tr.state.params.clone()
    coder = pipeline.GopCoder("cuda")
    fut = coder.submit(frames, tr.state.params, S)
    tr.fit(frames, 2)                              # keeps changing tr.state.params while the coder runs
    enc_bg = coder.collect()
    assert fut.done() and coder.collect() is None
    enc_serial = pipeline.encode_gop(frames, snap, S)
    assert enc_bg.frame_bytes == enc_serial.frame_bytes and enc_bg.model_bytes == enc_serial.model_bytes
    assert enc_bg.low_enc_bytes == enc_serial.low_enc_bytes and enc_bg.side_info == enc_serial.side_info
    fut2, st, _ = pipeline.overfit_encode_gop([p.cuda() for p in pts], epochs=1, seed=5, coder=coder)
    dec = pipeline.decode_gop(coder.collect())
    for d, p in zip(dec, pts):
        assert bool((d.cpu() == p).all())


def test_cli_end_to_end_lossless(env, tmp_path):
    from linr_pcgc_b200 import main as cli, pointio, synth
    ori = tmp_path / "ori"
    ori.mkdir()
    for i, p in enumerate(synth.make_sequence("tiny", 4)):
        pointio.write_ply_ascii(str(ori / f"f{i:03d}.ply"), p.numpy() + np.array([7, -3, 11]))   # non-zero frame minimum
    argv = ["--overfit", "True", "--encode", "True", "--decode", "True", "--mid_test", "True", "--ori_dir", str(ori),
            "--frame_num", "4", "--gop_size", "2", "--first_epoch", "3", "--others_epoch", "2",
            "--result_dir", str(tmp_path / "out"), "--encode_dir", str(tmp_path / "enc"), "--decode_dir", str(tmp_path / "dec")]
    cli.main(argv)
    for gop in ("gop_0_1", "gop_2_3"):
        assert os.path.exists(tmp_path / "out" / gop / "model.pth")
        res = json.load(open(tmp_path / "out" / gop / "result.json"))
        assert {"epoch", "loss", "train_time", "real_bpp_all", "model_bpp", "xyzlow_bpp"} <= set(res[0])
        for f in ("side_info.json", "bins/model.bin", "bins/low_enc_bytes.bin", "bins/frame0000_scale0.bin", "bins/frame0001_scale0.bin"):
            assert os.path.exists(tmp_path / "enc" / gop / f), f
        side = json.load(open(tmp_path / "enc" / gop / "side_info.json"))
        assert set(side) == {"mu", "b", "min_param", "max_param", "enc_mode", "bitdepth"}   # encoder.py:114
    ck = torch.load(tmp_path / "out" / "gop_0_1" / "model.pth", weights_only=False)
    assert len(ck["model"]) == 161 + 4 * int(ck["scale_num"])
    dec0 = pointio.read_ply(str(tmp_path / "dec" / "gop_2_3" / "frame0001_dec.ply"))
    src = pointio.read_ply(str(ori / "f003.ply"))
    assert (dec0 == src[np.lexsort((src[:, 2], src[:, 1], src[:, 0]))]).all()


def test_background_gop_preparer_matches_in_line_preparation():
    """pipeline.GopPreparer: the next GOP uploaded and prepared on a side stream / host thread while the caller works on
    its own stream -- the same tables bit for bit as prepare_gop in line, usable on the caller's stream afterwards."""
    import dataclasses
    from linr_pcgc_b200 import pipeline, synth
    from linr_pcgc_b200.trainer import GopTrainer
    pts = [p.cpu().pin_memory() for p in synth.make_sequence("tiny", 4)]
    ref = pipeline.prepare_gop(pts, None, 64, "cuda")
    S = ref[0].n_scales
    prep = pipeline.GopPreparer("cuda")
    prep.submit(pts, S, 64)
    tr = GopTrainer(S, "cuda", seed=3, max_rows=max(f.tables.n_rows for f in ref))
    tr.fit(ref, 1)                                   # the caller's stream is busy meanwhile
    got = prep.collect()
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert a.scale_off == b.scale_off and a.point_num == b.point_num and (a.coord_min == b.coord_min).all()
        assert torch.equal(a.xyz, b.xyz)
        for k in dataclasses.fields(a.tables):
            x, y = getattr(a.tables, k.name), getattr(b.tables, k.name)
            if not isinstance(x, torch.Tensor):
                continue
            if k.name == "anchor":       # defined where the column has a neighbour (the other words are never read)
                n = a.tables.n_rows
                for c in range(9):
                    has = ((a.tables.mask.to(torch.int64) >> (3 * c)) & 7) != 0
                    assert torch.equal(x[c, :n][has], y[c, :n][has]), "anchor"
            elif k.name == "pair_list":  # entries beyond a list's length are padding
                pass
            else:
                assert torch.equal(x, y), k.name
    l0 = GopTrainer(S, "cuda", seed=5, max_rows=tr.runner.max_rows).fit(ref, 1)
    l1 = GopTrainer(S, "cuda", seed=5, max_rows=tr.runner.max_rows).fit(got, 1)
    assert l0 == l1                                  # and they train to the same bits
