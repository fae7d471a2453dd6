"""Deterministic synthetic voxelised surface sequences.

The named datasets (8iVFB / Owlii / MVUB) are not available offline, so throughput and
parity are measured on closed bumpy surfaces with the same bit depth and point counts:

    r(u, t) = R * (1 + 0.15 sin(5 u_x + w t) cos(4 u_y) + 0.10 sin(7 u_z + w t))
    p       = floor(u * r + 2^(b-1) + drift(t))            then unique

Directions `u` come from a Fibonacci lattice (no RNG -> bit-identical on every host);
R is rescaled until the voxel count is within 2 % of the target.  Runs on any torch
device (float64), so the bench can build loot-sized GOPs on the GPU in milliseconds.
"""
from __future__ import annotations

import math
from typing import List

import torch

SHAPES = {
    # name: (bit depth, target points per frame)
    "plumbing": (10, 100_000),
    "loot": (10, 780_000),
    "owlii": (11, 2_500_000),
    "mvub9": (9, 300_000),
    "mvub10": (10, 300_000),
    "tiny": (7, 3_000),
}


def _directions(m: int, device) -> torch.Tensor:
    i = torch.arange(m, dtype=torch.float64, device=device)
    z = 1.0 - (2.0 * i + 1.0) / m
    phi = i * (math.pi * (3.0 - math.sqrt(5.0)))
    s = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    return torch.stack([s * torch.cos(phi), s * torch.sin(phi), z], dim=1)


def _voxelise(u: torch.Tensor, R: float, t: int, bits: int) -> torch.Tensor:
    w = 0.05 * t
    r = R * (1.0 + 0.15 * torch.sin(5.0 * u[:, 0] + w) * torch.cos(4.0 * u[:, 1]) + 0.10 * torch.sin(7.0 * u[:, 2] + w))
    c = float(1 << (bits - 1))
    drift = torch.tensor([0.3 * t, 0.1 * t, 0.0], dtype=torch.float64, device=u.device)
    p = torch.floor(u * r[:, None] + c + drift).to(torch.int64)
    p = torch.clamp(p, 0, (1 << bits) - 1)
    key = (p[:, 0] << 42) | (p[:, 1] << 21) | p[:, 2]
    key = torch.unique(key)  # sorted -> x-major lexicographic
    m = (1 << 21) - 1
    return torch.stack([(key >> 42) & m, (key >> 21) & m, key & m], dim=1).to(torch.int32)


def make_sequence(shape: str = "loot", frames: int = 1, device="cpu", oversample: int = 12,
                  bits: int | None = None, target: int | None = None, start: int = 0) -> List[torch.Tensor]:
    """Return `frames` tensors [Np,3] int32 (sorted, unique) of a temporally coherent sequence."""
    b, n = SHAPES[shape]
    bits = bits or b
    target = target or n
    u = _directions(oversample * target, device)
    R = math.sqrt(target / (4.0 * math.pi * 1.45))
    rmax = ((1 << (bits - 1)) - 8) / 1.27
    for _ in range(6):
        cnt = _voxelise(u, R, start, bits).shape[0]
        if abs(cnt - target) <= 0.02 * target:
            break
        R = min(R * math.sqrt(target / cnt), rmax)
    return [_voxelise(u, R, start + t, bits) for t in range(frames)]
