"""Parameter contract of LINR_PCGC_Model: tensor names / shapes in `model.parameters()` order.

The flat fp32 vector the kernels read is the concatenation of these tensors in this order — the same vector the
reference quantises (`torch.cat([p.view(-1) for p in model.parameters()])`, model_compression/model_size_est.py:391)
— so checkpoints (`loot/gop_32_62/model.pth` names and shapes) load by name.  Sources: models/model_core.py:31-34,
models/upsample.py:43-97, models/resnet.py:15-51, models/module_utils.py:42-81.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

Spec = List[Tuple[str, Tuple[int, ...]]]


def param_spec(scale_num: int, ch: int = 8) -> Spec:
    if ch != 8:
        raise ValueError("the sm_100a kernels are specialised for hidden_channel_conv=8 (main.py:520 default)")
    h = ch // 2
    spec: Spec = [("scale_emb.weight", (scale_num, 8))]
    for s in range(scale_num):
        spec += [(f"scale_mlp.{s}.0.weight", (16, 15)), (f"scale_mlp.{s}.0.bias", (16,)),
                 (f"scale_mlp.{s}.2.weight", (8, 16)), (f"scale_mlp.{s}.2.bias", (8,))]

    def block(prefix: str, cin: int) -> Spec:
        irn = f"{prefix}.2.layers.0"
        return [(f"{prefix}.0.kernel", (27, cin, ch)), (f"{prefix}.0.bias", (1, ch)),
                (f"{irn}.conv0_0.kernel", (27, ch, h)), (f"{irn}.conv0_0.bias", (1, h)),
                (f"{irn}.conv0_1.kernel", (27, h, h)), (f"{irn}.conv0_1.bias", (1, h)),
                (f"{irn}.conv1_0.kernel", (ch, h)), (f"{irn}.conv1_0.bias", (1, h)),
                (f"{irn}.conv1_1.kernel", (27, h, h)), (f"{irn}.conv1_1.bias", (1, h)),
                (f"{irn}.conv1_2.kernel", (h, h)), (f"{irn}.conv1_2.bias", (1, h)),
                (f"{prefix}.3.kernel", (27, ch, ch)), (f"{prefix}.3.bias", (1, ch))]

    spec += block("upsampler.block_in", 8)
    for k in range(8):
        p = f"upsampler.inner_mlps.{k}.0"
        spec += [(f"{p}.0.weight", (24, ch)), (f"{p}.0.bias", (24,)), (f"{p}.2.weight", (1, 24)), (f"{p}.2.bias", (1,))]
    for k in range(8):
        p = f"upsampler.prune_blocks.{k}.0.conv"
        spec += [(f"{p}.kernel", (27, ch, ch)), (f"{p}.bias", (1, ch))]
    for k in range(7):
        spec += block(f"upsampler.outter_blocks.{k}", k + 1)
    return spec


def offsets(spec: Spec) -> List[int]:
    out, o = [], 0
    for _, shp in spec:
        out.append(o)
        o += math.prod(shp)
    out.append(o)
    return out


def init_flat(scale_num: int, seed: int | None = None) -> torch.Tensor:
    """Random initialisation with the reference's distributions, returned as the flat CPU vector.

    nn.Embedding ~ N(0,1) (models/model_core.py:31); PointwiseMLP: xavier_uniform with ReLU gain, zero bias
    (models/module_utils.py:42-61); ME convolution kernel and bias ~ U(-1/sqrt(Cin*K), +1/sqrt(Cin*K)), K = kernel
    volume (MinkowskiEngine 0.5.4 `reset_parameters`)."""
    g = torch.Generator()
    if seed is None:
        g.seed()  # the reference never seeds (main.py:504 is parsed and unused)
    else:
        g.manual_seed(seed)
    parts, last_a = [], 0.0
    for name, shp in param_spec(scale_num):
        if name == "scale_emb.weight":
            t = torch.randn(shp, generator=g)
        elif name.endswith(".weight"):
            fan_out, fan_in = shp
            a = math.sqrt(2.0) * math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shp, generator=g) * 2 - 1) * a
        elif name.endswith(".kernel"):
            last_a = 1.0 / math.sqrt(shp[0] * shp[1] if len(shp) == 3 else shp[0])
            t = (torch.rand(shp, generator=g) * 2 - 1) * last_a
        elif len(shp) == 2:  # ME convolution bias [1, Cout], drawn with its kernel's bound
            t = (torch.rand(shp, generator=g) * 2 - 1) * last_a
        else:  # Linear bias
            t = torch.zeros(shp)
        parts.append(t.reshape(-1).float())
    return torch.cat(parts)


def named_views(flat: torch.Tensor, scale_num: int) -> Dict[str, torch.Tensor]:
    spec = param_spec(scale_num)
    offs = offsets(spec)
    assert flat.numel() == offs[-1], (flat.numel(), offs[-1])
    return {n: flat[offs[i]: offs[i + 1]].view(shp) for i, (n, shp) in enumerate(spec)}
