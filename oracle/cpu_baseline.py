"""CPU baseline for bench.py  --  TEST INFRASTRUCTURE (see oracle/omnivggt_oracle.py).

Bounded CPU sample for the ``cpu_baseline`` field of the product bench line (``kind: "port"``): the oracle port on the
workload's shapes (1 scene x S views @ 518 x 518, fp32, all host threads).  A full forward costs ~95 s on 8 threads
 -- the whole-forward measurement is ``bench.py --impl reference`` (the reference itself from
oracle/_ref); here each sample times ONE of each repeated unit at the full workload shape and scales by the unit counts
of the real model:
    t = t_patch_embed + 24 * t_dino_block + 24 * (t_frame_block + t_global_block) + 8 * t_dpt_frame(2 heads)
        + 4/trunk * t_camera_head
Every unit is executed at full width (C = 1024, 16 heads, L = 8 * 1374 tokens, DPT features 256) on real-shaped data."""
from __future__ import annotations

import time
from typing import Dict, Tuple

import torch

from . import omnivggt_oracle as O
from .synth import make_state_dict

_CACHE: Dict[str, object] = {}


def _schema():
    """Names + shapes of ONE of each repeated unit, read from the reference's own key schema (tests/golden/full.schema.json,
    written by oracle/make_golden_full.py): block / adapter indices other than 0 are dropped."""
    import json
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "full.schema.json")
    full = json.load(open(path))["schema"]
    rep = re.compile(r"\.(blocks|frame_blocks|global_blocks|trunk|pose_embeddings|camera_adapters)\.(\d+)\.")
    return {k: v for k, v in full.items() if all(int(m.group(2)) == 0 for m in rep.finditer(k))}


def _state():
    if "sd" not in _CACHE:
        _CACHE["sd"] = make_state_dict(_schema(), 0)
    return _CACHE["sd"]


@torch.no_grad()
def sample(S: int = 8, H: int = 518, W: int = 518) -> Tuple[float, Dict[str, float]]:
    """Returns (estimated seconds per view-set, per-unit timings)."""
    sd = _state()
    cfg = O.OracleConfig(dpt_layers=(0, 0, 0, 0))
    C, hp, wp = 1024, H // 14, W // 14
    P, T = hp * wp, hp * wp + 5
    g = torch.Generator().manual_seed(0)
    t = {}

    img = torch.rand(S, 3, H, W, generator=g)
    t0 = time.perf_counter()
    x = O.conv_patch_embed(sd, "aggregator.patch_embed.patch_embed", img, 14)
    x = torch.cat([sd["aggregator.patch_embed.cls_token"].expand(S, -1, -1), x], 1) + sd["aggregator.patch_embed.pos_embed"]
    x = torch.cat([x[:, :1], sd["aggregator.patch_embed.register_tokens"].expand(S, -1, -1), x[:, 1:]], 1)
    t["patch_embed"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    x = O.block(sd, "aggregator.patch_embed.blocks.0", x, 16, None, False, 0.0, 1e-6)
    t["dino_block"] = time.perf_counter() - t0

    tokens = torch.randn(S, T, C, generator=g)
    yy, xx = torch.meshgrid(torch.arange(hp), torch.arange(wp), indexing="ij")
    pos = torch.cat([torch.zeros(5, 2, dtype=torch.long), torch.stack([yy.reshape(-1), xx.reshape(-1)], -1) + 1])
    pos = pos[None].expand(S, -1, -1)
    t0 = time.perf_counter()
    tokens = O.block(sd, "aggregator.frame_blocks.0", tokens, 16, pos, True, 100.0, 1e-5)
    t["frame_block"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    tokens = O.block(sd, "aggregator.global_blocks.0", tokens.reshape(1, S * T, C), 16, pos.reshape(1, S * T, 2), True,
                     100.0, 1e-5)
    t["global_block"] = time.perf_counter() - t0

    inter = {0: torch.randn(1, 1, T, 2 * C, generator=g)}
    t0 = time.perf_counter()
    O.dpt_head(sd, "depth_head", inter, H, W, 5, cfg, "exp")
    O.dpt_head(sd, "point_head", inter, H, W, 5, cfg, "inv_log")
    t["dpt_frame_2heads"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.camera_head(sd, "camera_head", torch.randn(1, S, T, 2 * C, generator=g)[:, :, :1].expand(1, S, 1, 2 * C), cfg)
    t["camera_head_trunk1"] = time.perf_counter() - t0

    total = (t["patch_embed"] + 24 * t["dino_block"] + 24 * (t["frame_block"] + t["global_block"]) +
             S * t["dpt_frame_2heads"] + 4 * t["camera_head_trunk1"])
    return total, t


SAMPLE_DESC = ("oracle port, fp32, cfg2 shapes (1 scene x 8 views @518x518): one DINOv2 block, one frame block, one global "
               "block (L=10992), the DPT depth+point heads on 1 frame and a depth-1 camera trunk are timed at full width and "
               "scaled by the model's unit counts (24/24/24/8/4)")
