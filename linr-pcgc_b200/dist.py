"""Multi-GPU sharding of the codec (no counterpart in the reference, which is single-process: SURVEY.md 2.1, 8(e)).

One process per GPU (`torch.distributed`, NCCL over NVLink; gloo in the CPU tests).  The codec shards at two levels:

* GOPs: GOP 0 is trained first because its checkpoint, Adam moments and learning rate seed every later GOP
  (main.py:102-104,241-246).  GOPs 1.. are independent of each other -> `plan_job` deals them to groups of ranks, no
  collective between groups.
* ONE GOP on several ranks: the reference steps the optimiser once per frame (main.py:305-321), so dealing frames to
  ranks would change the result (k frames per step, 1/k of the steps: +16 % bpp measured in round 1 -- removed).
  Instead the 8 autoregressive stages of every frame are dealt to the ranks of the group (`stage_range`): a rank runs
  SCE + block_in (replicated) and the LDFE blocks + heads of its stages; the flat 219 kB gradient is summed by one
  all-reduce per frame and every rank takes the same fused Adam step, so parameters stay replicated and the training
  trajectory is the single-GPU one up to fp32 summation order (trainer.GopTrainer(stages=..., group=...)).
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


def stage_range(parts: int, part: int):
    """Stages [lo, hi) of the 8 that member `part` of a `parts`-rank group computes (contiguous, sizes differ by <= 1)."""
    if not 1 <= parts <= 8 or not 0 <= part < parts:
        raise ValueError(f"stage split needs 1..8 ranks per GOP (got part {part} of {parts})")
    return (part * 8) // parts, ((part + 1) * 8) // parts


def plan_job(n_gops: int, world_size: int):
    """Schedule of a whole sequence: a list of phases, each a list of (ranks, [gop indices]).

    Phase 0: GOP 0 on all ranks (at most 8 per GOP; extra ranks idle).  Phase 1: GOPs 1.. dealt round-robin to
    min(world, n_gops - 1) contiguous groups of ranks; a group runs its GOPs one after the other, stage-split over its
    ranks.  96 frames / gop 32 on 8 GPUs: [[(0..7, [0])], [((0..3), [1]), ((4..7), [2])]]."""
    first = list(range(min(world_size, 8)))
    phases = [[(first, [0])]]
    rest = list(range(1, n_gops))
    if rest:
        n_groups = min(world_size, len(rest))
        bounds = [(g * world_size) // n_groups for g in range(n_groups + 1)]
        groups = []
        for g in range(n_groups):
            ranks = list(range(bounds[g], bounds[g + 1]))[:8]
            groups.append((ranks, rest[g::n_groups]))
        phases.append(groups)
    return phases


def frame_share(n_frames: int, parts: int, part: int) -> List[int]:
    """Frames of one GOP that member `part` of a group codes (encode / decode shard by frame)."""
    return list(range(part, n_frames, parts))


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


_GROUPS = {}


def group_for(ranks: Sequence[int]):
    """Process group of `ranks` (None = the whole world).  Every rank must ask for the same groups in the same order
    (torch.distributed.new_group is collective); cached per rank tuple."""
    ranks = tuple(ranks)
    if world() == 1 or len(ranks) == world():
        return None
    if ranks not in _GROUPS:
        _GROUPS[ranks] = dist.new_group(list(ranks))
    return _GROUPS[ranks]


def make_groups(phases):
    """Create every group of a `plan_job` schedule on all ranks, in schedule order."""
    for phase in phases:
        for ranks, _ in phase:
            group_for(ranks)


def bind_rank_cores() -> int:
    """Bind this process to its 1/LOCAL_WORLD_SIZE slice of the host cores (round 1: 8 ranks' coder threads and launch
    threads shared 32 cores and cost 4 % at N = 8).  Returns the number of cores of the slice (0: left unbound)."""
    import os
    try:
        lw, lr = int(os.environ.get("LOCAL_WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
        cores = sorted(os.sched_getaffinity(0))
    except (AttributeError, OSError, ValueError):
        return 0
    if lw <= 1 or len(cores) < 2 * lw:
        return 0
    per = len(cores) // lw
    mine = cores[lr * per:(lr + 1) * per]
    os.sched_setaffinity(0, mine)
    return len(mine)


def gather_bytes(parts: Sequence[bytes], dst: int = 0):
    """Collect per-rank byte strings on `dst` (encode results of frame-sharded GOPs)."""
    if world() == 1:
        return [list(parts)]
    out = [None] * world() if rank() == dst else None
    dist.gather_object(list(parts), out, dst=dst)
    return out
