"""Generate tests/golden/*.npz by running the REFERENCE's own Python (from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py

What runs is the reference's code, unmodified and imported from where it lies:
`datautils.custom_dataset.MyDataset.handle_data` (octree prep), `models.model_core.
LINR_PCGC_Model.forward/encode/decode`, `decoder.decode_one_frame`, `models.sort_functions`,
`models.quantize_functions`, `model_compression.model_size_est.Model_Estimate`.  Its absent
third-party imports are satisfied by the CPU stand-ins under oracle/me_cpu (MinkowskiEngine,
torchac, open3d) and `.cuda()` is patched to a no-op, so: integer stages are the reference's
true outputs; network/bitstream fixtures pin the reference *control flow* on top of the
restated ME/torchac arithmetic (see oracle/linr_oracle.py header).
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "me_cpu"), ROOT, REF, os.path.join(REF, "models")]

# the reference hard-wires CUDA
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
_orig_tensor = torch.tensor


def _tensor(*a, **k):
    if str(k.get("device", "")).startswith("cuda"):
        k["device"] = "cpu"
    return _orig_tensor(*a, **k)


torch.tensor = _tensor

import linr_pcgc_b200.synth as synth  # noqa: E402  (input generator only)
from datautils.custom_dataset import MyDataset  # noqa: E402
import models.model_core as model_core  # noqa: E402
import models.module_utils as module_utils  # noqa: E402
import models.sort_functions as sort_functions  # noqa: E402
import models.quantize_functions as quantize_functions  # noqa: E402
import decoder as ref_decoder  # noqa: E402
import glob_params  # noqa: E402
from model_compression.model_size_est import Model_Estimate  # noqa: E402

model_core.device = torch.device("cpu")
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
offsets_ini = glob_params.offsets_ini


def ref_prepare(points: np.ndarray, scale_num=None, min_point_num=64):
    with tempfile.TemporaryDirectory() as d:
        np.save(os.path.join(d, "f0.npy"), points)
        ds = MyDataset(d, None, scale_num, "npy", stage=8, derive_ori=True)
        ds.set_prefix_data({"offsets_ini": offsets_ini, "offset_of_neigbor": None, "min_point_num": min_point_num})
        return ds[0], ds.scale_num


def int_fixture(name: str, points: np.ndarray, scale_num=None, min_point_num=64):
    data, S = ref_prepare(points, scale_num, min_point_num)
    out = {"points": points.astype(np.int32), "scale_num": np.int32(S), "min_point_num": np.int32(min_point_num),
           "point_num": np.int32(data["point_num"]), "coord_min": np.asarray(data["coord_data_min"], np.int32),
           "ori": data["ori"].numpy().astype(np.int32)}
    for i, sc in enumerate(data["all_input_info"]):
        q = sc["xyzqsc_t"]
        out[f"s{i}_coord"] = q.get_coord().numpy().astype(np.int32)
        out[f"s{i}_nbr7"] = q.get_offset_tensor().numpy().astype(np.uint8)
        out[f"s{i}_occ"] = torch.cat(sc["occ_lst"], dim=1).numpy().astype(np.uint8)
        out[f"s{i}_gt"] = sc["ground_truth"].numpy().astype(np.int32)
        # upper_layer round trip exactly as custom_dataset.py:295
        up = module_utils.octree_level_obj.upper_layer(q.get_coord(), torch.cat(sc["occ_lst"], dim=1))
        out[f"s{i}_up"] = up.numpy().astype(np.int32)
    out["n_scales"] = np.int32(len(data["all_input_info"]))
    np.savez_compressed(os.path.join(OUT, f"int_{name}.npz"), **out)
    print("int fixture", name, "scales", out["n_scales"], "points", out["point_num"])
    return data, S


def sort_fixture():
    rng = np.random.default_rng(8807)
    xyz = rng.integers(-5, 40, size=(500, 3)).astype(np.int32)
    t = torch.from_numpy(xyz)
    srt = sort_functions.sort_by_coord_sum_c(t).numpy()
    q2 = quantize_functions.quantize(t, 2).numpy()
    uq = torch.unique(t, dim=0)
    qs = module_utils.QuickSearchCoord(uq)
    query = torch.from_numpy(rng.integers(-6, 41, size=(300, 3)).astype(np.int32))  # within the +-1 envelope
    hit = qs.search(query).numpy().astype(np.uint8)[:, 0]
    idx = qs.search_coord_idx(query.clamp(min=int(qs.minimum))).numpy()  # ref has no negative guard here
    np.savez_compressed(os.path.join(OUT, "int_sort.npz"), xyz=xyz, sorted=srt, quant2=q2, uniq=qs.coord.numpy(),
                        query=query.numpy(), hit=hit, query_idx_clamped=query.clamp(min=int(qs.minimum)).numpy(), idx=idx)
    print("sort fixture ok")


def net_fixture(name: str, points: np.ndarray, seed: int):
    data, S = ref_prepare(points, None, 64)
    torch.manual_seed(seed)
    model = model_core.LINR_PCGC_Model({"scale_num": S, "in_channel": 7, "hidden_channel_conv": 8,
                                        "block_layers": 1, "outstage": 8, "instage": 1})
    # make biases non-trivial so the fixture exercises them
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("bias") and p.abs().sum() == 0:
                p.uniform_(-0.05, 0.05)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    names = [n for n, _ in model.named_parameters()]
    out = {"points": points.astype(np.int32), "scale_num": np.int32(S)}
    for k, v in sd.items():
        out["w:" + k] = v.numpy()
    out["param_order"] = np.array(names)

    # forward + backward exactly as main.py:305-316
    import main as ref_main  # noqa  (argparse lives under __main__, safe to import)
    model.train()
    bits = ref_main.overfit_one_frame(model, data["all_input_info"])
    loss = bits / data["point_num"]
    loss.backward()
    out["bits"] = np.float64(bits.item())
    out["loss"] = np.float64(loss.item())
    out["grad_flat"] = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).numpy()
    # per-scale probabilities (teacher forced)
    with torch.no_grad():
        for i, sc in enumerate(data["all_input_info"]):
            a = dict(sc)
            a["coord"] = sc["xyzqsc_t"].get_coord()
            a["offset_tensor"] = sc["xyzqsc_t"].get_offset_tensor()
            core = model.logic_core(a)
            out[f"s{i}_probs"] = torch.cat(core["out_cls_list"], dim=1).numpy()
            out[f"s{i}_scale_bits"] = np.float64(model(a).item())
    # one Adam step as main.py:231-237,319-321
    opt = torch.optim.Adam(model.parameters(), lr=0.01, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    opt.step()
    out["flat_after_adam"] = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy()
    model.load_state_dict(sd)

    # encode / decode through the reference's codec glue (encoder.py:158-203, decoder.py:153-176)
    model.eval()
    import encoder as ref_encoder  # noqa
    enc = ref_encoder.encode_one_frame(model, data["all_input_info"], data["ori"])
    for i, b in enumerate(enc["all_bytes"]):
        out[f"s{i}_bytes"] = np.frombuffer(b, dtype=np.uint8)
    low = data["all_input_info"][-1]["xyzqsc_t"].get_coord()
    dec = ref_decoder.decode_one_frame(model, list(enc["all_bytes"]), low)
    assert (dec["dec_coord"] != data["ori"]).sum() == 0, "reference control flow: lossless round trip failed"
    out["dec_coord"] = dec["dec_coord"].numpy().astype(np.int32)
    out["all_bit"] = np.int64(enc["all_bit"])

    # model compression (model_size_est.py:390-579)
    me = Model_Estimate()
    model2 = model_core.LINR_PCGC_Model({"scale_num": S, "in_channel": 7, "hidden_channel_conv": 8,
                                         "block_layers": 1, "outstage": 8, "instage": 1})
    # the shipped artefacts use AC mode (enc_mode 2); shrink the weights' spread so the Laplace model wins
    comp = me.compress_test(model, model2, 8)
    out["q_recon"] = comp["recon_ret"].detach().numpy()
    out["q_mu"] = np.float32(comp["mu"].item())
    out["q_b"] = np.float32(comp["b"].item())
    out["q_min"] = np.float32(comp["min_param"].item())
    out["q_max"] = np.float32(comp["max_param"].item())
    out["q_enc_mode"] = np.int32(comp["enc_mode"])
    out["q_bit_real"] = np.float64(comp["bit_real"])
    out["q_bytes"] = np.frombuffer(comp["final_bytes"], dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, f"net_{name}.npz"), **out)
    print("net fixture", name, "bits", out["bits"], "all_bit", out["all_bit"], "enc_mode", out["q_enc_mode"])


def checkpoint_fixture():
    ck = torch.load(os.path.join(REF, "loot/gop_32_62/model.pth"), map_location="cpu", weights_only=False)
    spec = [(k, list(v.shape)) for k, v in ck["model"].items()]
    with open(os.path.join(OUT, "loot_checkpoint_spec.json"), "w") as f:
        json.dump({"params": spec, "adam_step": float(ck["optimizer_state_dict"]["state"][0]["step"]),
                   "lr": ck["optimizer_state_dict"]["param_groups"][0]["lr"]}, f, indent=0)
    print("checkpoint spec", len(spec))


if __name__ == "__main__":
    if sys.argv[1:] == ["net_mid"]:
        # ~40k points, 9 bit: 5 scales, ~14k parent rows -> several 256-row chunks of partial sums per kernel (the tiny
        # fixture fits in one), every scale's SCE MLP exercised
        net_fixture("mid", synth.make_sequence("tiny", 1, bits=9, target=40_000)[0].numpy(), seed=2)
        sys.exit(0)
    sort_fixture()
    tiny = synth.make_sequence("tiny", 1)[0].numpy()
    int_fixture("tiny", tiny + np.array([3, -2, 7], np.int32))     # non-zero / negative min exercised
    int_fixture("tiny_s3", tiny, scale_num=3)
    rng = np.random.default_rng(8807)
    int_fixture("ragged", np.unique(rng.integers(0, 64, size=(900, 3)).astype(np.int32), axis=0), min_point_num=8)
    mid = synth.make_sequence("tiny", 1, bits=8, target=12_000)[0].numpy()
    int_fixture("mid", mid)
    net_fixture("tiny", tiny, seed=1)
    net_fixture("mid", synth.make_sequence("tiny", 1, bits=9, target=40_000)[0].numpy(), seed=2)
    checkpoint_fixture()
