"""`LINR_PCGC_Model`: the reference's model class (models/model_core.py:18-286) re-hosted on the C ABI.

Same constructor dict, same method names, argument dicts and return values, same `state_dict()` keys and
`parameters()` order contract — so `main.py`'s loop, `encoder.py`, `decoder.py` and `test_utils.py` can call it in
place of the MinkowskiEngine-based class:

    forward(d) -> 0-dim tensor (bits of one scale, differentiable)          models/model_core.py:72-81
    encode(d)  -> {'enc_bytes', 'bits', 'x_low'}                            models/model_core.py:236-266
    decode(d)  -> list of 8 [N,1] float32 occupancy tensors                 models/model_core.py:268-286
    codec(d)   -> {'bits', 'enc_bytes', 'enc_time', 'dec_time', 'bits_t'}   models/model_core.py:169-227

`d` carries `coord` [N,3] int32 (sorted unique parents of one scale), `occ_lst` (8 x [N,1] float 0/1),
`offset_tensor` [N,7] float 0/1, `scale_idx`.  The kernel map of a coordinate set is built once and cached by the
tensor's storage (the reference rebuilds coordinate managers ~16x per scale per iteration).

Parameters live in ONE flat fp32 `nn.Parameter` (`flat`), the concatenation of the reference's 189 tensors in
`parameters()` order; `state_dict()` / `load_state_dict()` expose and accept the reference's names and shapes.
There is no CPU path: inputs must be CUDA tensors and the extension must be built.
"""
from __future__ import annotations

import time
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import nn

from . import _lib, codec, rc
from . import params as P
from .frame import RowTables, build_tables
from .net import NetRunner

LN2 = 0.6931471805599453


def _pack_bits(cols: torch.Tensor) -> torch.Tensor:
    """[N,k] 0/1 (any dtype) -> uint8 [N], bit i = column i."""
    k = cols.shape[1]
    w = (1 << torch.arange(k, device=cols.device, dtype=torch.int32))
    return (cols.to(torch.int32) * w).sum(dim=1).to(torch.uint8).contiguous()


class _TableCache:
    """coordinate tensor (storage pointer, rows, scale) -> RowTables; bounded FIFO."""

    def __init__(self, capacity: int = 4096):
        self.cap = capacity
        self.d: "OrderedDict[tuple, tuple]" = OrderedDict()

    def get(self, coord: torch.Tensor, scale_idx: int) -> RowTables:
        key = (coord.data_ptr(), int(coord.shape[0]), int(scale_idx), coord._version)
        hit = self.d.get(key)
        if hit is not None:
            return hit[0]
        c = coord[:, -3:].to(torch.int32).contiguous()   # tolerate a leading batch column (ME-style [N,1+3])
        n = int(c.shape[0])
        scale = torch.full((n,), int(scale_idx), dtype=torch.uint8, device=c.device)
        t = build_tables(c, scale)
        self.d[key] = (t, coord)   # keep the key tensor alive so its pointer cannot be recycled
        if len(self.d) > self.cap:
            self.d.popitem(last=False)
        return t


class _ScaleBits(torch.autograd.Function):
    """bits = sum_k BCE_sum(p_k, gt_k) / ln 2 for one scale; backward = the deterministic CUDA backward."""

    @staticmethod
    def forward(ctx, flat, model, tables):
        runner = model._train_runner(tables.n_rows)
        out = runner.forward(flat.detach(), tables, train=True, loss_scale=1.0, want_bits=True)
        ctx.model, ctx.tables, ctx.runner = model, tables, runner
        ctx.save_for_backward(flat)
        return out["bits"].to(torch.float32).reshape(()).clone()

    @staticmethod
    def backward(ctx, gout):
        (flat,) = ctx.saved_tensors
        grad = torch.empty_like(flat)
        ctx.runner.backward(flat.detach(), ctx.tables, grad)
        ctx.model._release_runner(ctx.runner)
        return grad * gout.to(grad.dtype), None, None


class LINR_PCGC_Model(nn.Module):
    def __init__(self, inargs: Dict):
        super().__init__()
        self.scale_num = int(inargs["scale_num"])
        if int(inargs.get("in_channel", 7)) != 7 or int(inargs.get("hidden_channel_conv", 8)) != 8 \
                or int(inargs.get("block_layers", 1)) != 1 or int(inargs.get("outstage", 8)) != 8 \
                or int(inargs.get("instage", 1)) != 1:
            raise ValueError("the sm_100a kernels implement the live configuration of main.py:97,218,520-521: "
                             "in_channel 7, hidden_channel_conv 8, block_layers 1, outstage 8, instage 1")
        self.spec = P.param_spec(self.scale_num)
        self.flat = nn.Parameter(P.init_flat(self.scale_num, inargs.get("seed")))
        self._tables = _TableCache()
        self._free_runners: List[NetRunner] = []
        self._infer: Optional[NetRunner] = None

    # ---- parameters under the reference's names ---------------------------------------------------------------
    def state_dict(self, *args, destination=None, prefix="", keep_vars=False, **kw):
        out = OrderedDict() if destination is None else destination
        views = P.named_views(self.flat if keep_vars else self.flat.detach(), self.scale_num)
        for n, _ in self.spec:
            out[prefix + n] = views[n]
        return out

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        views = P.named_views(self.flat.data, self.scale_num)
        missing = [n for n in views if n not in state_dict]
        extra = [n for n in state_dict if n not in views and n != "flat"]
        if strict and (missing or extra):
            raise RuntimeError(f"load_state_dict: missing {missing[:4]}..., unexpected {extra[:4]}...")
        with torch.no_grad():
            for n, v in views.items():
                if n in state_dict:
                    src = state_dict[n]
                    if tuple(src.shape) != tuple(v.shape):
                        raise RuntimeError(f"size mismatch for {n}: {tuple(src.shape)} vs {tuple(v.shape)}")
                    v.copy_(src)
        return torch.nn.modules.module._IncompatibleKeys(missing, extra)

    # ---- plumbing ---------------------------------------------------------------------------------------------
    def _dev(self):
        if not self.flat.is_cuda:
            raise _lib.LinrError("LINR_PCGC_Model has no CPU path: call .cuda() first")
        return self.flat.device

    def _train_runner(self, rows: int) -> NetRunner:
        for i, r in enumerate(self._free_runners):
            if r.max_rows >= rows:
                return self._free_runners.pop(i)
        return NetRunner(self.scale_num, rows, self._dev(), train=True)

    def _release_runner(self, r: NetRunner):
        if len(self._free_runners) < 16:
            self._free_runners.append(r)

    def _infer_runner(self, rows: int) -> NetRunner:
        if self._infer is None:
            self._infer = NetRunner(self.scale_num, rows, self._dev(), train=False)
        self._infer.reserve(rows)
        return self._infer

    def _tables_for(self, d: Dict, need_occ: bool) -> RowTables:
        coord = d["coord"]
        if not coord.is_cuda:
            raise _lib.LinrError("coord must be a CUDA tensor (no CPU path)")
        t = self._tables.get(coord, int(d["scale_idx"]))
        occ = None
        if need_occ:
            occ = _pack_bits(torch.cat([o.reshape(-1, 1) for o in d["occ_lst"]], dim=1))
        # the 7 face-neighbour bits come from the caller's offset_tensor (qscTensor.set_offset_tensor), as in the reference
        nbr7 = _pack_bits(d["offset_tensor"]) if d.get("offset_tensor") is not None else t.nbr7
        return RowTables(coords=t.coords, scale=t.scale, nbr7=nbr7, anchor=t.anchor, mask=t.mask, occ=occ)

    # ---- reference API ----------------------------------------------------------------------------------------
    def forward(self, inargs: Dict) -> torch.Tensor:
        t = self._tables_for(inargs, need_occ=True)
        if not (torch.is_grad_enabled() and self.flat.requires_grad):
            out = self._infer_runner(t.n_rows).forward(self.flat.detach(), t, want_bits=True)
            return out["bits"].to(torch.float32).reshape(()).clone()
        return _ScaleBits.apply(self.flat, self, t)

    @torch.no_grad()
    def logic_core(self, in_args: Dict) -> Dict:
        """Teacher-forced probabilities (models/model_core.py:38-68): 8 x [N,1] probs and ground truths."""
        t = self._tables_for(in_args, need_occ=True)
        run = self._infer_runner(t.n_rows)
        out = run.forward(self.flat.detach(), t, want_probs=True, want_bits=False)
        probs = out["probs"]
        return {"out_cls_list": [probs[k].reshape(-1, 1).clone() for k in range(8)],
                "ground_truth_list": [o.reshape(-1, 1) for o in in_args["occ_lst"]]}

    def _cdf_host(self, t: RowTables, want_bits: bool):
        run = self._infer_runner(t.n_rows)
        out = run.forward(self.flat.detach(), t, want_cdf=True, want_bits=want_bits)
        cdf = out["cdf"].cpu().numpy().view(np.uint16)
        occ = t.occ.cpu().numpy()
        return cdf, occ, (float(out["bits"].item()) if want_bits else None)

    @torch.no_grad()
    def encode(self, in_args: Dict) -> Dict:
        t = self._tables_for(in_args, need_occ=True)
        cdf, occ, _ = self._cdf_host(t, False)
        streams = rc.encode_binary_batch([cdf[k] for k in range(8)], [occ] * 8, list(range(8)))
        data = codec.pack_bitstream(streams)
        return {"enc_bytes": data, "bits": len(data) * 8, "x_low": in_args["coord"]}

    @torch.no_grad()
    def decode(self, inagrs: Dict) -> List[torch.Tensor]:
        coord = inagrs["coord"]
        n = int(coord.shape[0])
        base = self._tables_for({"coord": coord, "scale_idx": inagrs["scale_idx"], "offset_tensor": inagrs.get("offset_tensor")}, False)
        occ = torch.zeros(n, dtype=torch.uint8, device=coord.device)
        t = RowTables(coords=base.coords, scale=base.scale, nbr7=base.nbr7, anchor=base.anchor, mask=base.mask, occ=occ)
        run = self._infer_runner(n)
        streams = codec.unpack_bitstream(inagrs["enc_bytes"])
        flat = self.flat.detach()
        run.decode_begin(flat, t)
        out = []
        for k in range(8):
            d_cdf, _ = run.decode_stage(flat, t, k)
            sym = rc.decode_binary(d_cdf.cpu().numpy().view(np.uint16), streams[k], n)
            d_sym = torch.from_numpy(sym).to(coord.device)
            run.occ_set_stage(occ, d_sym, k)
            out.append(d_sym.to(torch.float32).reshape(-1, 1))
        return out

    @torch.no_grad()
    def codec(self, inargs: Dict) -> Dict:
        """Mid-test path: ONE range-coder stream over all 8 stages, encode + decode timed (model_core.py:169-227)."""
        torch.cuda.synchronize()
        st1 = time.time()
        t = self._tables_for(inargs, need_occ=True)
        cdf, occ, bits_t = self._cdf_host(t, True)
        flat_cdf = np.ascontiguousarray(cdf.reshape(-1))
        sym = np.concatenate([(occ >> k) & 1 for k in range(8)]).astype(np.uint8)
        st2 = time.time()
        enc_bytes = rc.encode_binary(flat_cdf, sym)
        st3 = time.time()
        recon = rc.decode_binary(flat_cdf, enc_bytes, len(sym))
        st4 = time.time()
        assert (recon != sym).sum() == 0
        return {"bits": len(enc_bytes) * 8, "enc_bytes": enc_bytes, "enc_time": st3 - st1, "dec_time": st2 - st1 + st4 - st3,
                "bits_t": bits_t}
