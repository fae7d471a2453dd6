"""Thin host wrapper over the network entry points of the C ABI (forward / backward / decode / Adam / quant).

One `NetRunner` owns the caller-side memory the ABI asks for (workspace, probability / CDF / bit-count
buffers) for up to `max_rows` rows, so a training loop makes no allocation per iteration.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .frame import RowTables


def param_count(scale_num: int) -> int:
    return int(_lib.load().linr_param_count(scale_num))


def param_offsets(scale_num: int):
    buf = (C.c_int64 * 512)()
    n = _lib.load().linr_param_offsets(scale_num, buf, 512)
    return [int(buf[i]) for i in range(n)]


FWD_GDFE, FWD_PRE, FWD_POST, FWD_ALL = 1, 2, 4, 7                      # forward phases (include/linr_b200.h)
BWD_HEADS, BWD_LDFE, BWD_GDFE, BWD_FINAL, BWD_ALL = 1, 2, 4, 8, 15       # backward phases


class NetRunner:
    def __init__(self, scale_num: int, max_rows: int, device, train: bool = True):
        self.lib = _lib.load()
        self.S = scale_num
        self.P = param_count(scale_num)
        self.device = torch.device(device)
        self.train = train
        self.max_rows = 0
        self.ws = None
        self.reserve(max_rows)
        self.bits = torch.zeros(1, dtype=torch.float64, device=self.device)
        # a training runner is one user of the device: its context takes turns at the constant weight bank
        # (linr_ctx_create / include/linr_b200.h)
        self.ctx = None
        if train:
            idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
            h = C.c_void_p()
            check(self.lib.linr_ctx_create(int(idx), C.byref(h)), "linr_ctx_create")
            self.ctx = h

    def close(self):
        """Destroy the context (waits for its last training call if the bank still holds its weights)."""
        if getattr(self, "ctx", None) is not None:
            self.lib.linr_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bank_calls(self) -> int:
        """Training calls of this runner that ran on the constant-bank kernels."""
        return int(self.lib.linr_ctx_bank_calls(self.ctx)) if self.ctx is not None else 0

    def bank_launches(self) -> int:
        return int(self.lib.linr_ctx_bank_launches(self.ctx)) if self.ctx is not None else 0

    def reserve(self, rows: int):
        if rows <= self.max_rows and self.ws is not None:
            return
        self.max_rows = max(rows, 1)
        nbytes = self.lib.linr_net_ws_bytes(self.max_rows, 1 if self.train else 0)
        self.ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        self.probs = torch.empty(8 * self.max_rows, dtype=torch.float32, device=self.device)
        self.cdf = torch.empty(8 * self.max_rows, dtype=torch.int16, device=self.device)

    # -- teacher-forced forward over all 8 stages ---------------------------------------------------------
    def forward(self, params: torch.Tensor, t: RowTables, train: bool = False, loss_scale: float = 0.0,
                want_probs: bool = False, want_cdf: bool = False, want_bits: bool = True, stages=(0, 8), phases: int = FWD_ALL,
                same_params: bool = False):
        """`stages` = (lo, hi): only the stages lo..hi-1, `phases`: which part of the pass (one rank's share of a stage
        split, linr_net_forward_stages)."""
        assert params.is_cuda and params.dtype == torch.float32 and params.numel() == self.P
        assert t.occ is not None, "teacher-forced forward needs the ground-truth occupancy"
        assert not train or self.train, "runner was created without training workspace"
        n = t.n_rows
        self.reserve(n)
        rows = t.rows()
        if train:
            self.lib.linr_ctx_set_current(self.ctx)
            if same_params:      # a later phase of the same iteration: the weights staged by the first phase are still valid
                self.lib.linr_ctx_hint_same_params(self.ctx)
        check(self.lib.linr_net_forward_stages(ptr(params), self.S, C.byref(rows), int(stages[0]), int(stages[1]), int(phases),
                                               1 if train else 0, loss_scale,
                                               ptr(self.probs) if want_probs else None, ptr(self.cdf) if want_cdf else None,
                                               ptr(self.bits) if want_bits else None, ptr(self.ws), self.ws.numel(), stream_ptr()),
              "linr_net_forward_stages")
        out = {}
        if want_bits:
            out["bits"] = self.bits
        if want_probs:
            out["probs"] = self.probs[: 8 * n].view(8, n)
        if want_cdf:
            out["cdf"] = self.cdf[: 8 * n].view(8, n)
        return out

    def backward(self, params: torch.Tensor, t: RowTables, grad: torch.Tensor, stages=(0, 8), phases: int = BWD_ALL,
                 own_gdfe: bool = True, same_params: bool = False):
        assert grad.is_cuda and grad.numel() >= self.P
        rows = t.rows()
        self.lib.linr_ctx_set_current(self.ctx)
        if same_params:
            self.lib.linr_ctx_hint_same_params(self.ctx)
        check(self.lib.linr_net_backward_stages(ptr(params), self.S, C.byref(rows), int(stages[0]), int(stages[1]), int(phases),
                                                1 if own_gdfe else 0, ptr(grad), ptr(self.ws), self.ws.numel(), stream_ptr()),
              "linr_net_backward_stages")

    def exchange_views(self, n_rows: int):
        """(g, dg): float32 [n_rows, 8] views of the workspace -- what a stage split broadcasts (g = block_in's output) and
        reduces (dg = its gradient) between the phases."""
        key = int(n_rows)
        hit = self._views.get(key) if hasattr(self, "_views") else None
        if hit is None or hit[2] is not self.ws:
            og, od = C.c_int64(), C.c_int64()
            check(self.lib.linr_net_ws_offsets(key, 1 if self.train else 0, self.S, C.byref(og), C.byref(od)), "linr_net_ws_offsets")
            nb = key * 32
            g = self.ws[og.value: og.value + nb].view(torch.float32).view(key, 8)
            dg = self.ws[od.value: od.value + nb].view(torch.float32).view(key, 8) if od.value >= 0 else None
            if not hasattr(self, "_views"):
                self._views = {}
            hit = self._views[key] = (g, dg, self.ws)
        return hit[0], hit[1]

    # -- sequential decode ----------------------------------------------------------------------------------
    def decode_begin(self, params: torch.Tensor, t: RowTables):
        self.reserve(t.n_rows)
        rows = t.rows()
        check(self.lib.linr_net_decode_begin(ptr(params), self.S, C.byref(rows), ptr(self.ws), self.ws.numel(), stream_ptr()),
              "linr_net_decode_begin")

    def decode_stage(self, params: torch.Tensor, t: RowTables, stage: int, want_probs: bool = False):
        n = t.n_rows
        rows = t.rows()
        check(self.lib.linr_net_decode_stage(ptr(params), self.S, C.byref(rows), stage,
                                             ptr(self.probs) if want_probs else None, ptr(self.cdf), ptr(self.ws),
                                             self.ws.numel(), stream_ptr()), "linr_net_decode_stage")
        return self.cdf[:n], (self.probs[:n] if want_probs else None)

    def decode_scale(self, params: torch.Tensor, t: RowTables, streams, d_sym: torch.Tensor, h_cdf: torch.Tensor,
                     h_sym: torch.Tensor):
        """The 8 sequential stages of one scale in one C call (linr_net_decode_scale): t.occ must be zero-filled and
        holds the decoded occupancy afterwards; h_cdf / h_sym are pinned host scratch of >= n_rows elements."""
        n = t.n_rows
        assert len(streams) == 8 and h_cdf.is_pinned() and h_sym.is_pinned() and h_cdf.numel() >= n and h_sym.numel() >= n
        self.reserve(n)
        bufs = [np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, np.uint8) for b in streams]
        ptrs = (C.c_void_p * 8)(*[b.ctypes.data for b in bufs])
        lens = (C.c_int64 * 8)(*[len(b) for b in streams])
        rows = t.rows()
        check(self.lib.linr_net_decode_scale(ptr(params), self.S, C.byref(rows), ptrs, lens, ptr(self.cdf), ptr(d_sym),
                                             h_cdf.data_ptr(), h_sym.data_ptr(), ptr(self.ws), self.ws.numel(), stream_ptr()),
              "linr_net_decode_scale")

    def decode_scale_batch(self, params: torch.Tensor, t: RowTables, seg_off, streams, d_sym: torch.Tensor, h_cdf: torch.Tensor,
                           h_sym: torch.Tensor, threads: int):
        """The 8 stages of one scale for SEVERAL frames in one C call (linr_net_decode_scale_batch): t holds the frames'
        parents concatenated (segment f = rows seg_off[f]..seg_off[f+1], kept apart in space by the caller);
        streams[f] = the 8 stage bitstreams of frame f."""
        n, F = t.n_rows, len(streams)
        assert len(seg_off) == F + 1 and seg_off[0] == 0 and seg_off[-1] == n
        assert h_cdf.is_pinned() and h_sym.is_pinned() and h_cdf.numel() >= n and h_sym.numel() >= n
        self.reserve(n)
        bufs = [[np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, np.uint8) for b in st] for st in streams]
        ptrs = (C.c_void_p * (8 * F))(*[b.ctypes.data for st in bufs for b in st])
        lens = (C.c_int64 * (8 * F))(*[len(b) for st in streams for b in st])
        offs = (C.c_int64 * (F + 1))(*[int(o) for o in seg_off])
        rows = t.rows()
        check(self.lib.linr_net_decode_scale_batch(ptr(params), self.S, C.byref(rows), F, offs, ptrs, lens, ptr(self.cdf), ptr(d_sym),
                                                   h_cdf.data_ptr(), h_sym.data_ptr(), int(threads), ptr(self.ws), self.ws.numel(),
                                                   stream_ptr()), "linr_net_decode_scale_batch")

    def occ_set_stage(self, occ: torch.Tensor, sym: torch.Tensor, stage: int):
        check(self.lib.linr_occ_set_stage(ptr(occ), ptr(sym), int(occ.numel()), stage, stream_ptr()), "linr_occ_set_stage")


def adam_step(params, grad, m, v, step: int, lr: float, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4):
    """torch.optim.Adam with L2 decay (main.py:231-237) on the flat buffers, one launch."""
    check(_lib.load().linr_adam_fused(ptr(params), ptr(grad), ptr(m), ptr(v), params.numel(), step, lr, b1, b2, eps, wd,
                                      stream_ptr()), "linr_adam_fused")


def param_quant(params: torch.Tensor, bitdepth: int = 8):
    """quant_uniform2 + Laplace stats (model_size_est.py:72-91,410-411) -> (q u8, recon f32, stats[min,max,mu,b])."""
    recon = torch.empty_like(params)
    stats = torch.empty(4, dtype=torch.float32, device=params.device)
    if bitdepth <= 8:
        q = torch.empty(params.numel(), dtype=torch.uint8, device=params.device)
        check(_lib.load().linr_param_quant(ptr(params), params.numel(), bitdepth, ptr(q), ptr(recon), ptr(stats), stream_ptr()),
              "linr_param_quant")
    else:      # 9..16 bit symbols (int16 storage, read as uint16)
        q = torch.empty(params.numel(), dtype=torch.int16, device=params.device)
        check(_lib.load().linr_param_quant16(ptr(params), params.numel(), bitdepth, ptr(q), ptr(recon), ptr(stats), stream_ptr()),
              "linr_param_quant16")
    return q, recon, stats


# -- single-layer entry points (ME.MinkowskiConvolution, kernel_size 3, stride 1) ------------------------------
def spconv27_fwd(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], t: RowTables, relu: bool = False):
    """y = sum_k x[row(C+delta_k)] @ W[k] + bias  (models/upsample.py:17,90,95; models/resnet.py:15-51)."""
    cin, cout = int(W.shape[1]), int(W.shape[2])
    y = torch.empty((t.n_rows, cout), dtype=torch.float32, device=x.device)
    rows = t.rows()
    check(_lib.load().linr_spconv27_fwd(ptr(x.contiguous()), cin, ptr(W.contiguous()), ptr(bias), ptr(y), cout, C.byref(rows),
                                        1 if relu else 0, stream_ptr()), "linr_spconv27_fwd")
    return y


def spconv27_bwd_in(dy: torch.Tensor, W: torch.Tensor, t: RowTables):
    cin, cout = int(W.shape[1]), int(W.shape[2])
    dx = torch.empty((t.n_rows, cin), dtype=torch.float32, device=dy.device)
    rows = t.rows()
    check(_lib.load().linr_spconv27_bwd_in(ptr(dy.contiguous()), cin, ptr(W.contiguous()), ptr(dx), cout, C.byref(rows),
                                           stream_ptr()), "linr_spconv27_bwd_in")
    return dx


def spconv27_bwd_w(x: torch.Tensor, dy: torch.Tensor, t: RowTables):
    lib = _lib.load()
    cin, cout = int(x.shape[1]), int(dy.shape[1])
    dW = torch.empty((27, cin, cout), dtype=torch.float32, device=x.device)
    db = torch.empty(cout, dtype=torch.float32, device=x.device)
    ws = torch.empty(int(lib.linr_spconv27_bwd_w_ws_bytes(t.n_rows, cin, cout)), dtype=torch.uint8, device=x.device)
    rows = t.rows()
    check(lib.linr_spconv27_bwd_w(ptr(x.contiguous()), cin, ptr(dy.contiguous()), cout, C.byref(rows), ptr(dW), ptr(db), ptr(ws),
                                  ws.numel(), stream_ptr()), "linr_spconv27_bwd_w")
    return dW, db
