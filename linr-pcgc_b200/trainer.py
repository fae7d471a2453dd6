"""Per-GOP overfitting loop: host-side mirror of `overfit_one_gop` / `overfit_one_frame`
(main.py:122-455, :457-475) with the frame's whole forward + backward + Adam running as stream-ordered kernels.

One optimiser step per frame, exactly as the reference: loss = bits / point_num (main.py:315), Adam(lr .01,
betas .9/.999, eps 1e-8, L2 1e-4) (main.py:231-237), StepLR(step_size 32, gamma .992) stepped per frame
(main.py:252,321), lr floored at min_lr after every epoch (main.py:433-437).  Later GOPs start from GOP 0's
parameters, Adam moments, step count and learning rate (main.py:102-104,241-246).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import params as P
from .frame import Frame
from .net import NetRunner, adam_step


@dataclass
class OptimState:
    """What `torch.save({'model', 'optimizer_state_dict', ...})` carries between GOPs (main.py:365-374)."""
    params: torch.Tensor
    m: torch.Tensor
    v: torch.Tensor
    step: int          # Adam step count
    sched_step: int    # StepLR counter
    lr: float

    def clone(self) -> "OptimState":
        return OptimState(self.params.clone(), self.m.clone(), self.v.clone(), self.step, self.sched_step, self.lr)


class GopTrainer:
    def __init__(self, scale_num: int, device="cuda", learning_rate: float = 0.01, gamma: float = 0.992,
                 step_size: int = 32, min_lr: float = 4e-4, decay_rate: float = 1e-4, seed: Optional[int] = None,
                 state: Optional[OptimState] = None, max_rows: int = 1, grad_hook: Optional[Callable] = None):
        self.S = scale_num
        self.device = torch.device(device)
        self.gamma, self.step_size, self.min_lr, self.wd = gamma, step_size, min_lr, decay_rate
        n = P.offsets(P.param_spec(scale_num))[-1]
        if state is None:
            flat = P.init_flat(scale_num, seed).to(self.device)
            state = OptimState(flat, torch.zeros(n, device=self.device), torch.zeros(n, device=self.device), 0, 0, learning_rate)
        self.state = state
        self.grad = torch.empty(n, dtype=torch.float32, device=self.device)
        self.runner = NetRunner(scale_num, max_rows, self.device, train=True)
        self.grad_hook = grad_hook   # e.g. an NCCL all-reduce of the flat gradient under intra-GOP data parallelism
        self.bits_log: List[torch.Tensor] = []

    # one frame-iteration (main.py:305-321)
    def step(self, frame: Frame, record_bits: bool = True):
        st = self.state
        out = self.runner.forward(st.params, frame.tables, train=True, loss_scale=1.0 / frame.point_num,
                                  want_bits=record_bits)
        self.runner.backward(st.params, frame.tables, self.grad)
        if self.grad_hook is not None:
            self.grad_hook(self.grad)
        st.step += 1
        adam_step(st.params, self.grad, st.m, st.v, st.step, st.lr, wd=self.wd)
        st.sched_step += 1
        if st.sched_step % self.step_size == 0:
            st.lr *= self.gamma
        return out.get("bits")

    def end_epoch(self):
        if self.state.lr < self.min_lr:
            self.state.lr = self.min_lr

    def fit(self, frames: Sequence[Frame], epochs: int, log: Optional[Callable[[Dict], None]] = None) -> List[float]:
        """`epochs` passes over the GOP; returns the mean loss (bits/point) per epoch (one host sync per epoch)."""
        losses = []
        bits = torch.zeros(len(frames), dtype=torch.float64, device=self.device)
        pn = torch.tensor([f.point_num for f in frames], dtype=torch.float64, device=self.device)
        for ep in range(epochs):
            for i, f in enumerate(frames):
                b = self.step(f)
                bits[i: i + 1].copy_(b)   # stream-ordered: the runner reuses its bits buffer
            self.end_epoch()
            loss = float((bits / pn).mean().item())
            losses.append(loss)
            if log is not None:
                log({"epoch": ep, "loss": loss, "lr": self.state.lr})
        return losses
