"""Deterministic synthetic weights and inputs  --  TEST INFRASTRUCTURE (see oracle/omnivggt_oracle.py).

The reference checkpoint is not available offline, so parity work uses
random weights.  Stock init would make the test vacuous (camera_adapters are zero-initialised,
omnivggt_aggregator.py:70-72; LayerScale gamma 0.01; camera/register tokens std 1e-6,
aggregator.py:136-137), so every tensor is drawn "de-zeroed" from a seeded generator with a
scale chosen per parameter role.  Values depend only on (name order, shape, seed), so the
golden generator (which loads them into the real reference modules) and the tests (which feed
them to the oracle / CUDA path) reproduce the identical state dict from the small JSON schema.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch

Schema = Dict[str, List[int]]


def _scale_for(name: str, shape: Sequence[int]) -> tuple:
    """(mean, std) for a parameter, by role."""
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "gamma":                                   # LayerScale
        return 0.25, 0.05
    if "norm" in name and leaf == "weight" and len(shape) == 1:
        return 1.0, 0.1
    if leaf == "bias":
        return 0.0, 0.05
    if leaf in ("camera_token", "register_token", "depth_placeholder", "cls_token", "register_tokens"):
        return 0.0, 0.5
    if leaf == "pos_embed":
        return 0.0, 0.2
    if leaf == "empty_pose_tokens":
        return 0.0, 0.3
    if leaf == "weight" and len(shape) >= 2:
        if "resize_layers.0" in name or "resize_layers.1" in name:   # ConvTranspose [Cin,Cout,k,k]
            fan_in = shape[0]
        else:
            fan_in = int(math.prod(shape[1:]))
        return 0.0, 1.0 / math.sqrt(fan_in)
    return 0.0, 0.1


def make_state_dict(schema: Schema, seed: int = 0, device="cpu", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for name in sorted(schema):
        shape = schema[name]
        mean, std = _scale_for(name, shape)
        t = torch.randn(shape, generator=g, dtype=torch.float32) * std + mean
        sd[name] = t.to(device=device, dtype=dtype)
    return sd


def random_rotations(n: int, g: torch.Generator) -> torch.Tensor:
    q, r = torch.linalg.qr(torch.randn(n, 3, 3, generator=g))
    q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1))[:, None, :]
    det = torch.linalg.det(q)
    q[:, :, 0] = q[:, :, 0] * det[:, None]
    return q


def make_inputs(B: int, S: int, H: int, W: int, seed: int = 1) -> Dict[str, torch.Tensor]:
    """Synthetic inputs shaped like visual_util.py:835-841 (SURVEY.md section 8d recipe)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    images = torch.rand(B, S, 3, H, W, generator=g)
    R = random_rotations(B * S, g).reshape(B, S, 3, 3)
    t = torch.randn(B, S, 3, 1, generator=g)
    extr = torch.cat([R, t], -1)
    intr = torch.zeros(B, S, 3, 3)
    intr[..., 0, 0] = 500.0 * W / 518
    intr[..., 1, 1] = 500.0 * W / 518
    intr[..., 0, 2] = W / 2
    intr[..., 1, 2] = H / 2
    intr[..., 2, 2] = 1.0
    depth = 0.5 + 4.0 * torch.rand(B, S, H, W, 1, generator=g)
    mask = (torch.rand(B, S, H, W, generator=g) > 0.2).float()
    depth = depth * mask[..., None]
    return dict(images=images, extrinsics=extr, intrinsics=intr, depth=depth, mask=mask)
