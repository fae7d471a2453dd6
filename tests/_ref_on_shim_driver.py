"""Subprocess body of tests/test_reference_on_shim.py (build container only: needs /root/reference).

Runs the REFERENCE's own, unmodified Python (models/model_core.py, upsample.py, resnet.py, function_utils.py,
module_utils.py, main.overfit_one_frame, encoder.encode_one_frame, decoder.decode_one_frame) on top of the PRODUCT's
`MinkowskiEngine` / `torchac` drop-in modules (linr_pcgc_b200.shim).
 * With a CUDA device AND a reference checkout (LINR_REFERENCE_DIR, default /root/reference) nothing is replaced: the
   reference's `.cuda()` calls are real and its modules run on the sm_100a kernels (tolerances 1e-4, fp32 on device).
 * Without a GPU (the build container) the three device entry points under the shim (kernel-map build,
   linr_spconv27_fwd/_bwd_in/_bwd_w) are replaced by CPU test doubles made from the oracle; everything above them --
   SparseTensor / CoordinateManager bookkeeping, union adds, pruning, cat, the conv modules and their autograd wiring,
   the torchac surface on the real host range coder -- is the shipped code.
Prints one JSON line that the test compares with tests/golden/net_tiny.npz (recorded from the same reference code).
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("LINR_REFERENCE_DIR", "/root/reference")
sys.path[:0] = [ROOT, REF, os.path.join(REF, "models")]
ON_GPU = torch.cuda.is_available() and os.environ.get("LINR_SHIM_DOUBLES") != "1"

if not ON_GPU:
    # the reference hard-wires CUDA
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _orig_tensor = torch.tensor

    def _tensor(*a, **k):
        if str(k.get("device", "")).startswith("cuda"):
            k["device"] = "cpu"
        return _orig_tensor(*a, **k)

    torch.tensor = _tensor
sys.modules["open3d"] = types.ModuleType("open3d")      # PLY IO only (custom_dataset.py:4); the fixture is .npy

import linr_pcgc_b200.shim as shim  # noqa: E402
from oracle import linr_oracle as O  # noqa: E402

shim.install(force=True)
import MinkowskiEngine as ME  # noqa: E402  -> linr_pcgc_b200.shim.MinkowskiEngine
import torchac  # noqa: E402           -> linr_pcgc_b200.shim.torchac
assert ME.__name__.startswith("linr_pcgc_b200.shim") and torchac.__name__.startswith("linr_pcgc_b200.shim")


# ---- CPU test doubles of the device layer ------------------------------------------------------------------
class _Tab:
    def __init__(self, nbr):
        self.nbr = nbr
        self.n_rows = int(nbr.shape[0])


def _build_tables(xyz, scale, *a, **k):
    return _Tab(torch.from_numpy(O.nbr27(xyz.numpy().astype(np.int32)).astype(np.int64)))


def _fwd(x, W, b, t, relu=False):
    y = O.conv27(x, t.nbr, W, b)
    return torch.relu(y) if relu else y


def _bwd_in(dy, W, t):
    with torch.enable_grad():      # called from inside autograd.Function.backward
        x0 = torch.zeros(t.n_rows, W.shape[1], requires_grad=True)
        return torch.autograd.grad(O.conv27(x0, t.nbr, W.detach(), None), x0, dy)[0]


def _bwd_w(x, dy, t):
    with torch.enable_grad():
        W0 = torch.zeros(27, x.shape[1], dy.shape[1], requires_grad=True)
        b0 = torch.zeros(dy.shape[1], requires_grad=True)
        gW, gb = torch.autograd.grad(O.conv27(x.detach(), t.nbr, W0, b0), (W0, b0), dy)
    return gW, gb


if not ON_GPU:
    ME._check_cuda = lambda t: None
    ME._check_coordinate_set = lambda coords, xyz: None     # validated on the device path (tests/test_gpu_me_shim.py)
    ME.build_tables = _build_tables
    ME._net.spconv27_fwd, ME._net.spconv27_bwd_in, ME._net.spconv27_bwd_w = _fwd, _bwd_in, _bwd_w

# ---- the reference's own code ---------------------------------------------------------------------------------
from datautils.custom_dataset import MyDataset  # noqa: E402
import models.model_core as model_core  # noqa: E402
import glob_params  # noqa: E402
import main as ref_main  # noqa: E402
import encoder as ref_encoder  # noqa: E402
import decoder as ref_decoder  # noqa: E402

if not ON_GPU:
    model_core.device = torch.device("cpu")
fx = np.load(os.path.join(ROOT, "tests", "golden", "net_tiny.npz"), allow_pickle=False)
points = fx["points"]
with tempfile.TemporaryDirectory() as d:
    np.save(os.path.join(d, "f0.npy"), points)
    ds = MyDataset(d, None, None, "npy", stage=8, derive_ori=True)
    ds.set_prefix_data({"offsets_ini": glob_params.offsets_ini, "offset_of_neigbor": None, "min_point_num": 64})
    data, S = ds[0], ds.scale_num
assert S == int(fx["scale_num"])
model = model_core.LINR_PCGC_Model({"scale_num": S, "in_channel": 7, "hidden_channel_conv": 8, "block_layers": 1,
                                    "outstage": 8, "instage": 1})
names = [n for n, _ in model.named_parameters()]
assert names == [str(n) for n in fx["param_order"]], "parameters() order differs from the reference run on real ME shapes"
model.load_state_dict({k[2:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("w:")})
if ON_GPU:
    model = model.cuda()

res = {"n_params": sum(p.numel() for p in model.parameters()), "on_gpu": ON_GPU}
model.train()
bits = ref_main.overfit_one_frame(model, data["all_input_info"])
loss = bits / data["point_num"]
loss.backward()
res["bits"], res["bits_fixture"] = float(bits.item()), float(fx["bits"])
g = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).cpu().numpy()
res["grad_max_abs_err"] = float(np.abs(g - fx["grad_flat"]).max())
res["grad_max_abs"] = float(np.abs(fx["grad_flat"]).max())
perr = 0.0
with torch.no_grad():
    for i, sc in enumerate(data["all_input_info"]):
        a = dict(sc)
        a["coord"] = sc["xyzqsc_t"].get_coord()
        a["offset_tensor"] = sc["xyzqsc_t"].get_offset_tensor()
        core = model.logic_core(a)
        perr = max(perr, float((torch.cat(core["out_cls_list"], dim=1).cpu() - torch.from_numpy(fx[f"s{i}_probs"])).abs().max()))
res["probs_max_abs_err"] = perr
model.eval()
enc = ref_encoder.encode_one_frame(model, data["all_input_info"], data["ori"])
res["bytes_equal"] = all(bytes(b) == fx[f"s{i}_bytes"].tobytes() for i, b in enumerate(enc["all_bytes"]))
res["all_bit"], res["all_bit_fixture"] = int(enc["all_bit"]), int(fx["all_bit"])
low = data["all_input_info"][-1]["xyzqsc_t"].get_coord()
dec = ref_decoder.decode_one_frame(model, list(enc["all_bytes"]), low)
res["lossless"] = bool((dec["dec_coord"].cpu() != data["ori"].cpu()).sum() == 0)
res["tables_built"] = len(ME._tables.d)
print("RESULT " + json.dumps(res))
