"""Per-GOP overfitting loop: host-side mirror of `overfit_one_gop` / `overfit_one_frame`
(main.py:122-455, :457-475) with the frame's whole forward + backward + Adam running as stream-ordered kernels.

One optimiser step per frame, exactly as the reference: loss = bits / point_num (main.py:315), Adam(lr .01,
betas .9/.999, eps 1e-8, L2 1e-4) (main.py:231-237), StepLR(step_size 32, gamma .992) stepped per frame
(main.py:252,321), lr floored at min_lr after every epoch (main.py:433-437).  Later GOPs start from GOP 0's
parameters, Adam moments, step count and learning rate (main.py:102-104,241-246); the StepLR counter restarts with
every GOP because the reference builds a fresh scheduler per GOP (main.py:252).

Several GPUs on ONE GOP ("stage split"): every rank of `ranks` steps through the same frames in the
same order and computes its share of the 8 autoregressive stages of each frame (dist.stage_range; LDFE block + head per
stage).  The first rank also owns SCE + block_in: it broadcasts g = block_in's output ([rows,8] floats) while the others
already run their LDFE blocks (which read occupancy bits, not g), and receives the sum of the ranks' dg = sum_k dh_k
(one reduce) while they run their LDFE backward.  One all-reduce(sum) of the flat 219 kB gradient per frame, then the
same fused Adam step on every rank: the parameters stay replicated bit for bit and the optimiser still steps once per
frame exactly as the reference does (main.py:305-321).
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
                 ranks: Optional[Sequence[int]] = None, group=None):
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
        # stage split: `ranks` = the global ranks that share this GOP (None / one rank: no split), `group` their process group
        self.ranks = list(ranks) if ranks is not None and len(ranks) > 1 else None
        self.group = group
        self.split = self.ranks is not None
        if self.split:
            from . import dist as D
            self.part = self.ranks.index(torch.distributed.get_rank())
            self.stages = D.stage_range(len(self.ranks), self.part)
            self.leader = self.ranks[0]          # owner of SCE + block_in
        else:
            self.part, self.stages, self.leader = 0, (0, 8), 0
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
        t, ls = frame.tables, 1.0 / frame.point_num
        if not self.split:
            out = self.runner.forward(st.params, t, train=True, loss_scale=ls, want_bits=record_bits)
            self.runner.backward(st.params, t, self.grad)
        else:
            out = self._split_iteration(st.params, t, ls, record_bits)
        if self.grad_hook is not None:
            self.grad_hook(self.grad)
        st.step += 1
        adam_step(st.params, self.grad, st.m, st.v, st.step, st.lr, wd=self.wd)
        sched_after_step(st, self.step_size, self.gamma)
        return out.get("bits")

    def _split_iteration(self, params, t, ls, record_bits):
        """Forward + backward of this rank's stages with the two exchanges of the split between the phases; every launch
        and both collectives are stream-ordered (no host sync)."""
        from . import net as N
        dist = torch.distributed
        run, own = self.runner, self.part == 0
        kw = dict(train=True, loss_scale=ls, stages=self.stages)
        run.reserve(t.n_rows)
        g, dg = run.exchange_views(t.n_rows)
        if own:
            run.forward(params, t, want_bits=False, phases=N.FWD_GDFE, **kw)
        wb = dist.broadcast(g, src=self.leader, group=self.group, async_op=True)     # overlaps the LDFE blocks below
        run.forward(params, t, want_bits=False, phases=N.FWD_PRE, same_params=own, **kw)
        wb.wait()
        out = run.forward(params, t, want_bits=record_bits, phases=N.FWD_POST, same_params=True, **kw)
        run.backward(params, t, self.grad, stages=self.stages, phases=N.BWD_HEADS, own_gdfe=own)
        wr = dist.reduce(dg, dst=self.leader, op=dist.ReduceOp.SUM, group=self.group, async_op=True)   # overlaps the LDFE backward
        run.backward(params, t, self.grad, stages=self.stages, phases=N.BWD_LDFE, own_gdfe=own, same_params=True)
        wr.wait()
        if own:
            run.backward(params, t, self.grad, stages=self.stages, phases=N.BWD_GDFE, own_gdfe=True, same_params=True)
        run.backward(params, t, self.grad, stages=self.stages, phases=N.BWD_FINAL, own_gdfe=own)
        # sum of the per-rank contributions = the frame's gradient
        dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.group)
        return out

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
