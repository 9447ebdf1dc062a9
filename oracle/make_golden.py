"""Generate tests/golden/* by running the UNMODIFIED reference (build container only).

    python oracle/make_golden.py

Builds reduced-size instances of the reference's own classes (ZeroAggregator, DPTHead, CameraHead,
DinoVisionTransformer; every hot-path class is constructor-parametrisable),
loads deterministic de-zeroed weights (oracle/synth.py), runs the reference forward composition of
models/omnivggt.py:20-68 on seeded inputs and stores: the state-dict schema (names + shapes, JSON)
and the reference outputs (safetensors).  tests/test_oracle_golden.py replays the same cases through
oracle/omnivggt_oracle.py; GPU tests replay them through the CUDA path.
"""
from __future__ import annotations

import json
import os
import sys
from functools import partial

import torch
from safetensors.torch import save_file

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_shims import import_reference  # noqa: E402
from oracle.synth import make_inputs, make_state_dict  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

# model variants -------------------------------------------------------------------------------
VARIANTS = {
    # conv patchifier, C=128 (2 heads x 64), 4+4 blocks
    # (DPT channel counts are multiples of 64 -- the tap-GEMM conv needs Cin % 64 == 0, features // 2 included)
    "mini_conv": dict(embed_dim=128, depth=4, patch_embed="conv", features=128,
                      out_channels=[64, 128, 256, 256], cam_heads=2, cam_trunk=2, img_size=56),
    # reduced DINOv2 patchifier (2 ViT blocks) in front of the same aggregator
    "mini_dino": dict(embed_dim=128, depth=4, patch_embed="dino", features=128,
                      out_channels=[64, 128, 256, 256], cam_heads=2, cam_trunk=1, img_size=56),
}

# cases: name -> (variant, B, S, H, W, depth_idx, cam_idx, input_seed)
CASES = {
    "conv_images_only": ("mini_conv", 1, 3, 56, 56, [], [], 1),
    "conv_partial_aux_b2": ("mini_conv", 2, 4, 56, 56, [0, 2], [0, 1, 3], 2),
    "conv_full_aux_rect": ("mini_conv", 1, 2, 42, 70, [0, 1], [0, 1], 3),
    "conv_single_cam": ("mini_conv", 1, 3, 56, 56, [1], [0], 4),
    "conv_chunked_s9": ("mini_conv", 1, 9, 42, 42, [3], [0, 5], 5),
    "dino_square": ("mini_dino", 1, 2, 56, 56, [0], [0, 1], 6),
    "dino_rect_interp": ("mini_dino", 1, 2, 42, 70, [], [], 7),
}


class RefModel(torch.nn.Module):
    """Reference modules composed exactly as models/omnivggt.py:11-17, at reduced size."""

    def __init__(self, agg, dpt, cam, vit, layers, v):
        super().__init__()
        C = v["embed_dim"]
        self.aggregator = agg.ZeroAggregator(img_size=v["img_size"], patch_size=14, embed_dim=C, depth=v["depth"],
                                             num_heads=C // 64, patch_embed="conv", pose_hidden_dim=9)
        if v["patch_embed"] == "dino":
            self.aggregator.patch_embed = vit.DinoVisionTransformer(
                img_size=v["img_size"], patch_size=14, embed_dim=C, depth=2, num_heads=2, mlp_ratio=4,
                block_fn=partial(layers.NestedTensorBlock, attn_class=layers.MemEffAttention),
                num_register_tokens=4, interpolate_antialias=True, interpolate_offset=0.0, block_chunks=0,
                init_values=1.0)
        idx = list(range(v["depth"]))[-4:]
        self.camera_head = cam.CameraHead(dim_in=2 * C, num_heads=v["cam_heads"], trunk_depth=v["cam_trunk"])
        self.point_head = dpt.DPTHead(dim_in=2 * C, output_dim=4, activation="inv_log", conf_activation="expp1",
                                      features=v["features"], out_channels=v["out_channels"],
                                      intermediate_layer_idx=idx)
        self.depth_head = dpt.DPTHead(dim_in=2 * C, output_dim=2, activation="exp", conf_activation="expp1",
                                      features=v["features"], out_channels=v["out_channels"],
                                      intermediate_layer_idx=idx)

    @torch.no_grad()
    def forward(self, images, extrinsics, intrinsics, depth, mask, depth_gt_index, camera_gt_index):
        toks, psi = self.aggregator(images=images, extrinsics=extrinsics, intrinsics=intrinsics, depth=depth,
                                    mask=mask, depth_gt_index=depth_gt_index, camera_gt_index=camera_gt_index)
        out = {}
        pl = self.camera_head(toks)
        out["pose_enc"] = pl[-1]
        for i, p in enumerate(pl):
            out[f"pose_enc_list.{i}"] = p
        out["depth"], out["depth_conf"] = self.depth_head(toks, images=images, patch_start_idx=psi)
        out["world_points"], out["world_points_conf"] = self.point_head(toks, images=images, patch_start_idx=psi)
        out["agg_last"] = toks[-1]
        out["agg_first"] = toks[0]
        return out


def main():
    agg, dpt, cam, vit, layers = import_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    models = {}
    for vname, v in VARIANTS.items():
        torch.manual_seed(0)
        m = RefModel(agg, dpt, cam, vit, layers, v).eval()
        schema = {k: list(t.shape) for k, t in m.state_dict().items()}
        sd = make_state_dict(schema, seed=0)
        m.load_state_dict(sd, strict=True)
        with open(os.path.join(GOLDEN, f"{vname}.schema.json"), "w") as f:
            json.dump({"variant": v, "schema": schema}, f, indent=0, sort_keys=True)
        models[vname] = m
        print(vname, len(schema), "tensors", sum(t.numel() for t in sd.values()) / 1e6, "M params")
    index = {}
    for cname, (vname, B, S, H, W, didx, cidx, seed) in CASES.items():
        inp = make_inputs(B, S, H, W, seed=seed)
        out = models[vname](inp["images"], inp["extrinsics"], inp["intrinsics"], inp["depth"], inp["mask"],
                            list(didx), list(cidx))
        save_file({k: v.contiguous().clone() for k, v in out.items()}, os.path.join(GOLDEN, f"{cname}.safetensors"))
        index[cname] = dict(variant=vname, B=B, S=S, H=H, W=W, depth_gt_index=didx, camera_gt_index=cidx,
                            input_seed=seed, weight_seed=0)
        print(cname, {k: tuple(v.shape) for k, v in out.items() if k in ("depth", "pose_enc", "world_points")})
    with open(os.path.join(GOLDEN, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
