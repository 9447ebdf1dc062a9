"""tests/golden/postprocess.safetensors: outputs of the UNMODIFIED reference post-processing functions
(utils/pose_enc.py:65-130, utils/geometry.py:151-180) and of numpy.percentile as inference.py:132-133 uses it, on seeded
inputs (build container only; TEST INFRASTRUCTURE).        python oracle/make_golden_post.py"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
from safetensors.torch import save_file

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_shims import import_reference  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def make_post_inputs(S=3, H=42, W=70, seed=21):
    g = torch.Generator().manual_seed(seed)
    pose = torch.randn(1, S, 9, generator=g)
    pose[..., 3:7] = torch.nn.functional.normalize(pose[..., 3:7], dim=-1) * (0.8 + 0.4 * torch.rand(1, S, 1, generator=g))
    pose[..., 7:] = 0.6 + 0.5 * torch.rand(1, S, 2, generator=g)          # fov in radians
    depth = 0.3 + 5.0 * torch.rand(1, S, H, W, 1, generator=g)
    conf = 1.0 + torch.rand(1, S, H, W, generator=g).pow(3) * 8.0
    return pose, depth, conf


def main():
    import_reference()
    from omnivggt.utils.pose_enc import pose_encoding_to_extri_intri
    from omnivggt.utils.geometry import unproject_depth_map_to_point_map
    pose, depth, conf = make_post_inputs()
    H, W = depth.shape[2:4]
    ext, intr = pose_encoding_to_extri_intri(pose, (H, W))
    world = unproject_depth_map_to_point_map(depth[0].numpy(), ext[0].numpy(), intr[0].numpy())
    out = {"extrinsic": ext.contiguous(), "intrinsic": intr.contiguous(), "world_points_from_depth": torch.from_numpy(world)}
    for pct in (0.0, 37.5, 50.0, 99.9, 100.0):
        flat = conf.numpy().reshape(-1)
        thr = np.percentile(flat, pct)                                  # inference.py:132
        out[f"thr_{pct}"] = torch.tensor(float(thr), dtype=torch.float64)
        out[f"mask_{pct}"] = torch.from_numpy((flat >= thr) & (flat > 0.1)).to(torch.uint8)
    save_file(out, os.path.join(GOLDEN, "postprocess.safetensors"))
    print({k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
