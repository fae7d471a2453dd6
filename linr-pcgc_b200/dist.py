"""Multi-GPU sharding of the codec (no counterpart in the reference, which is single-process: SURVEY.md 2.1, 8(e)).

One process per GPU (`torch.distributed`, NCCL over NVLink; gloo in the CPU tests).  Two levels, as the codec shards:

* GOPs: GOP 0 is trained first because its checkpoint, Adam moments and learning rate seed every later GOP
  (main.py:102-104,241-246); it is trained data-parallel over all ranks, or by rank 0 alone with
  `broadcast_state`.  GOPs 1.. are independent -> `plan_gops` deals them round-robin, no collective on the data path.
* frames of ONE GOP (`frame_shard` + `GradAllReduce`): each rank runs forward/backward on its own frame, the flat
  219 kB gradient is summed over ranks (one NCCL all-reduce, latency-bound) and every rank takes the same fused Adam
  step on replicated parameters, so no parameter broadcast is ever needed.  k ranks turn k per-frame steps into one
  k-frame step (mean gradient): 1/k of the optimiser steps per epoch -- a semantic change against the reference,
  reported by bench.py as a separate mode.
Encode and decode shard by frame (frames of a GOP only share the model).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def plan_gops(n_gops: int, world_size: int, first_is_seed: bool = True) -> List[List[int]]:
    """GOP indices per rank.  GOP 0 (the seed, trained before the others start) is not dealt when `first_is_seed`."""
    todo = list(range(1 if first_is_seed else 0, n_gops))
    return [todo[r::world_size] for r in range(world_size)]


def frame_shard(n_frames: int, world_size: int, r: int) -> List[int]:
    """Frames of one GOP handled by rank r in data-parallel mode; every rank gets the same number of steps
    (the tail is padded by wrapping around, so the collective count matches on all ranks)."""
    per = -(-n_frames // world_size)
    return [(r + i * world_size) % n_frames for i in range(per)]


def broadcast_state(state, src: int = 0):
    """Ship GOP 0's optimiser state (params, moments, counters, lr) to every rank."""
    if world() == 1:
        return state
    for t in (state.params, state.m, state.v):
        dist.broadcast(t, src)
    meta = torch.tensor([state.step, state.sched_step, state.lr], dtype=torch.float64, device=state.params.device)
    dist.broadcast(meta, src)
    state.step, state.sched_step, state.lr = int(meta[0].item()), int(meta[1].item()), float(meta[2].item())
    return state


class GradAllReduce:
    """grad_hook for GopTrainer: mean of the per-rank flat gradients (sum all-reduce, then 1/world)."""

    def __init__(self, average: bool = True):
        self.average = average
        self.calls = 0

    def __call__(self, grad: torch.Tensor):
        w = world()
        if w > 1:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM)
            if self.average:
                grad.mul_(1.0 / w)
        self.calls += 1


def gather_bytes(parts: Sequence[bytes], dst: int = 0):
    """Collect per-rank byte strings on `dst` (encode results of frame-sharded GOPs)."""
    if world() == 1:
        return [list(parts)]
    out = [None] * world() if rank() == dst else None
    dist.gather_object(list(parts), out, dst=dst)
    return out
