"""Per-GOP overfitting loop: host-side mirror of `overfit_one_gop` / `overfit_one_frame`
(main.py:122-455, :457-475) with the frame's whole forward + backward + Adam running as stream-ordered kernels.

One optimiser step per frame, exactly as the reference: loss = bits / point_num (main.py:315), Adam(lr .01,
betas .9/.999, eps 1e-8, L2 1e-4) (main.py:231-237), StepLR(step_size 32, gamma .992) stepped per frame
(main.py:252,321), lr floored at min_lr after every epoch (main.py:433-437).  Later GOPs start from GOP 0's
parameters, Adam moments, step count and learning rate (main.py:102-104,241-246); the StepLR counter restarts with
every GOP because the reference builds a fresh scheduler per GOP (main.py:252).

Several GPUs on ONE GOP ("stage split"): every rank steps through the same frames in the same
order and computes the stages `stages = (lo, hi)` of each frame (linr_net_forward_stages / _backward_stages); one
all-reduce(sum) of the flat gradient per frame, then the same fused Adam step on every rank, so the parameters stay
replicated bit for bit and the optimiser still steps once per frame exactly as the reference does.
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


def sched_after_step(st: OptimState, step_size: int, gamma: float) -> None:
    """torch.optim.lr_scheduler.StepLR.step() after an optimiser step (main.py:252,321): the chainable form multiplies the
    CURRENT lr by gamma every `step_size` calls -- it does not recompute it from the initial lr, so the per-epoch floor
    below and a warm start from another GOP's lr carry over."""
    st.sched_step += 1
    if st.sched_step % step_size == 0:
        st.lr *= gamma


def sched_end_epoch(st: OptimState, min_lr: float) -> None:
    """main.py:433-437: after every epoch the lr is raised back to `min_lr` if it fell below."""
    if st.lr < min_lr:
        st.lr = min_lr


def sched_new_gop(st: OptimState) -> OptimState:
    """A later GOP loads GOP 0's optimizer state dict (lr, Adam moments and step counts continue, main.py:241-246) and
    builds a FRESH StepLR (main.py:252): the decay counter restarts at 0."""
    st.sched_step = 0
    return st


class GopTrainer:
    def __init__(self, scale_num: int, device="cuda", learning_rate: float = 0.01, gamma: float = 0.992,
                 step_size: int = 32, min_lr: float = 4e-4, decay_rate: float = 1e-4, seed: Optional[int] = None,
                 state: Optional[OptimState] = None, max_rows: int = 1, grad_hook: Optional[Callable] = None,
                 stages=(0, 8), group=None):
        self.S = scale_num
        self.device = torch.device(device)
        self.gamma, self.step_size, self.min_lr, self.wd = gamma, step_size, min_lr, decay_rate
        n = P.offsets(P.param_spec(scale_num))[-1]
        if state is None:
            flat = P.init_flat(scale_num, seed).to(self.device)
            state = OptimState(flat, torch.zeros(n, device=self.device), torch.zeros(n, device=self.device), 0, 0, learning_rate)
        else:
            sched_new_gop(state)
        self.state = state
        self.grad = torch.empty(n, dtype=torch.float32, device=self.device)
        self.runner = NetRunner(scale_num, max_rows, self.device, train=True)
        self.grad_hook = grad_hook   # called with the flat gradient before the Adam step
        self.stages = (int(stages[0]), int(stages[1]))
        self.group = group           # torch.distributed group of the ranks that share this GOP (stage split)
        self.split = self.stages != (0, 8)
        self.bits_log: List[torch.Tensor] = []

    def close(self):
        """Destroy the runner's context (NetRunner.close); the trainer cannot step afterwards."""
        self.runner.close()

    def reset(self, state: Optional[OptimState] = None, seed: Optional[int] = None, learning_rate: float = 0.01):
        """Start another GOP with this trainer's workspace: from `state` (a later GOP, seeded by GOP 0: main.py:102-104)
        or from a fresh random model."""
        if state is None:
            flat = P.init_flat(self.S, seed).to(self.device)
            state = OptimState(flat, torch.zeros_like(flat), torch.zeros_like(flat), 0, 0, learning_rate)
        else:
            sched_new_gop(state)
        self.state = state

    # one frame-iteration (main.py:305-321)
    def step(self, frame: Frame, record_bits: bool = True):
        st = self.state
        out = self.runner.forward(st.params, frame.tables, train=True, loss_scale=1.0 / frame.point_num,
                                  want_bits=record_bits, stages=self.stages)
        self.runner.backward(st.params, frame.tables, self.grad, stages=self.stages)
        if self.split:
            # sum of the per-rank stage contributions = the frame's gradient; stream-ordered, no host sync
            torch.distributed.all_reduce(self.grad, op=torch.distributed.ReduceOp.SUM, group=self.group)
        if self.grad_hook is not None:
            self.grad_hook(self.grad)
        st.step += 1
        adam_step(st.params, self.grad, st.m, st.v, st.step, st.lr, wd=self.wd)
        sched_after_step(st, self.step_size, self.gamma)
        return out.get("bits")

    def end_epoch(self):
        sched_end_epoch(self.state, self.min_lr)

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
            if self.split:   # each rank holds the bits of its own stages
                torch.distributed.all_reduce(bits, op=torch.distributed.ReduceOp.SUM, group=self.group)
            loss = float((bits / pn).mean().item())
            losses.append(loss)
            if log is not None:
                log({"epoch": ep, "loss": loss, "lr": self.state.lr})
        return losses
