"""CPU: the oracle (oracle/linr_oracle.py, oracle/rc_oracle.c) against fixtures recorded from the
reference's own Python (oracle/gen_golden.py).  Integer stages bit-exact; float stages 1e-4 rel."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import linr_oracle as O
from oracle import rc

from conftest import GOLDEN


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.mark.parametrize("name", ["tiny", "tiny_s3", "ragged", "mid"])
def test_prepare_frame_matches_reference(name):
    g = _load(f"int_{name}.npz")
    scale_num = int(g["scale_num"]) if name == "tiny_s3" else None
    fr = O.prepare_frame(g["points"], scale_num, int(g["min_point_num"]))
    assert len(fr["scales"]) == int(g["n_scales"])
    assert fr["point_num"] == int(g["point_num"])
    assert (fr["coord_min"] == g["coord_min"]).all()
    assert (fr["xyz"] == g["ori"]).all()
    for i, sc in enumerate(fr["scales"]):
        assert (sc["coord"] == g[f"s{i}_coord"]).all()
        assert (sc["occ"] == g[f"s{i}_occ"]).all()
        assert (sc["nbr7"] == g[f"s{i}_nbr7"]).all()
        assert (sc["ground_truth"] == g[f"s{i}_gt"]).all()
        assert (O.octree_up(sc["coord"], sc["occ"]) == g[f"s{i}_up"]).all()


def test_sort_quantize_lookup_match_reference():
    g = _load("int_sort.npz")
    assert (O.sort_rows_lex(g["xyz"]) == g["sorted"]).all()
    assert (O.unique_rows_lex(g["xyz"].astype(np.int64) >> 1) == g["quant2"]).all()
    assert (O.unique_rows_lex(g["xyz"]) == g["uniq"]).all()
    hit = O.lookup_rows(g["uniq"], g["query"]) >= 0
    assert (hit.astype(np.uint8) == g["hit"]).all()
    assert (O.lookup_rows(g["uniq"], g["query_idx_clamped"]) == g["idx"]).all()


def test_nbr27_consistency():
    g = _load("int_mid.npz")
    c = g["s0_coord"]
    t = O.nbr27(c)
    assert (t[:, 13] == np.arange(len(c))).all()
    # pair (i -> o, k)  <=>  (o -> i, 26-k) at stride 1
    for k in range(27):
        o = np.nonzero(t[:, k] >= 0)[0]
        assert (t[t[o, k], 26 - k] == o).all()
    # the 7 face offsets are columns of the 27 table
    cols = [13, 12, 14, 10, 16, 4, 22]
    assert ((t[:, cols] >= 0).astype(np.uint8) == g["s0_nbr7"]).all()


def _sd_from(g):
    return {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}


def test_param_spec_matches_reference_model_and_checkpoint():
    g = _load("net_tiny.npz")
    S = int(g["scale_num"])
    spec = O.param_spec(S)
    assert [n for n, _ in spec] == list(g["param_order"])
    for n, shp in spec:
        assert tuple(g["w:" + n].shape) == shp
    ck = json.load(open(os.path.join(GOLDEN, "loot_checkpoint_spec.json")))
    assert [(n, list(s)) for n, s in O.param_spec(7)] == [(n, s) for n, s in ck["params"]]
    assert sum(int(np.prod(s)) for _, s in O.param_spec(7)) == 54712
    # Appendix A.2 consistency check: lr after 7223 steps
    assert abs(O.lr_at_step(7223) - ck["lr"]) / ck["lr"] < 0.01


@pytest.mark.parametrize("fixture", ["tiny", "mid"])
def test_network_forward_backward_adam_match_reference_flow(fixture):
    g = _load(f"net_{fixture}.npz")
    S = int(g["scale_num"])
    sd = _sd_from(g)
    fr = O.prepare_frame(g["points"], None, 64)
    flat = O.flatten_params(sd, S).clone().requires_grad_(True)
    sdv = O.unflatten_params(flat, S)
    bits = O.frame_bits(sdv, fr)
    loss = bits / fr["point_num"]
    loss.backward()
    assert abs(bits.item() - float(g["bits"])) <= 1e-4 * abs(float(g["bits"]))
    for i, sc in enumerate(fr["scales"]):
        with torch.no_grad():
            _, p = O.scale_forward(sd, sc)
        np.testing.assert_allclose(p.numpy(), g[f"s{i}_probs"], rtol=1e-4, atol=1e-6)
    gref = g["grad_flat"]
    err = np.abs(flat.grad.numpy() - gref).max() / np.abs(gref).max()
    assert err < 1e-4
    # Adam (main.py:231-237): one step from zero moments
    p = [flat.detach().clone()]
    m, v = [torch.zeros_like(p[0])], [torch.zeros_like(p[0])]
    O.adam_step_reference(p, [torch.from_numpy(gref.copy())], m, v, step=1, lr=0.01)  # fed the recorded grad: tests Adam alone
    np.testing.assert_allclose(p[0].numpy(), g["flat_after_adam"], rtol=1e-5, atol=1e-6)


def test_codec_bitstreams_match_reference_flow():
    g = _load("net_tiny.npz")
    sd = _sd_from(g)
    fr = O.prepare_frame(g["points"], None, 64)
    all_bytes = O.encode_frame(sd, fr)
    tot = 0
    for i, b in enumerate(all_bytes):
        ref = g[f"s{i}_bytes"].tobytes()
        # probabilities agree to ~1e-6, so 16-bit CDFs may differ in a few symbols: sizes must agree closely
        assert abs(len(b) - len(ref)) <= max(2, 0.005 * len(ref))
        tot += len(b) * 8
    assert abs(tot - int(g["all_bit"])) <= 0.005 * int(g["all_bit"])
    dec = O.decode_frame(sd, all_bytes, fr["scales"][-1]["coord"])
    assert (dec == g["dec_coord"]).all()
    # the reference-made streams decode losslessly with the oracle decoder too
    ref_bytes = [g[f"s{i}_bytes"].tobytes() for i in range(len(all_bytes))]
    dec2 = O.decode_frame(sd, ref_bytes, fr["scales"][-1]["coord"])
    assert (dec2 == g["dec_coord"]).all()


@pytest.mark.parametrize("fixture", ["tiny", "mid"])
def test_model_compression_matches_reference(fixture):
    g = _load(f"net_{fixture}.npz")
    S = int(g["scale_num"])
    flat = O.flatten_params(_sd_from(g), S)
    c = O.compress_model(flat, 8)
    assert c["enc_mode"] == int(g["q_enc_mode"])
    assert c["mu"] == float(g["q_mu"]) and c["b"] == float(g["q_b"])
    assert float(c["min_param"]) == float(g["q_min"]) and float(c["max_param"]) == float(g["q_max"])
    np.testing.assert_array_equal(c["recon"].numpy(), g["q_recon"])
    assert c["final_bytes"] == g["q_bytes"].tobytes()
    assert c["bit_real"] == float(g["q_bit_real"])
    rec = O.decompress_model(c, flat.numel())
    np.testing.assert_array_equal(rec.numpy(), g["q_recon"])


def test_range_coder_round_trip_and_edges():
    rng = np.random.default_rng(3)
    for n in [0, 1, 2, 33, 5000]:
        p = np.clip(rng.random(n).astype(np.float32) ** 4, 0, 1)
        p[: n // 10] = 0.0        # saturated probabilities still get >= 2^-16 mass
        p[n // 10: n // 5] = 1.0
        sym = (rng.random(n) < 0.5).astype(np.int16)
        c = O.cdf_u16_binary(p)
        cdf = np.stack([np.zeros(n, np.uint16), c, np.zeros(n, np.uint16)], axis=1).reshape(-1, 3)
        b = rc.encode_u16(cdf, sym)
        assert (rc.decode_u16(cdf, b, n) == sym).all()
        if n:
            f = np.stack([np.zeros(n, np.float32), np.float32(1) - p, np.ones(n, np.float32)], 1)
            assert (rc.float_cdf_to_u16(f)[:, 1] == c).all()
            assert c.min() >= 1 and c.max() <= 65535
