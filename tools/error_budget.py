"""Where does the deviation of the dense outputs come from?  (GPU box; diagnostic, not a test.)

Runs the packed unmodified reference (oracle/_ref/omnivggt_ref.zip) on the GPU in true fp32 (TF32 off) as the truth and measures,
against it, on one synthetic scene:
  yard        the reference under torch.autocast(bf16) (SURVEY section 8d's yardstick; its heads stay fp32)
  ours        this package end to end
  agg_only    OUR aggregator snapshots (bf16 slots) fed to the reference's fp32 heads      -> aggregator + snapshot rounding
  heads_only  the reference's fp32 tokens fed to OUR DPT heads                               -> DPT kernels alone
  snap_bf16   the reference's tokens rounded to bf16, reference fp32 heads                   -> snapshot rounding alone
  sim:<group> the reference's fp32 heads with ONE group of convolutions given bf16-rounded inputs and weights
  tf32        the reference's fp32 heads with cuDNN TF32 convolutions allowed (PyTorch's default on a GPU)
"""
from __future__ import annotations

import contextlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KEYS = ("depth", "depth_conf", "world_points", "world_points_conf")
GROUPS = {"projects": ("projects.",), "resize": ("resize_layers.",), "layer_rn": ("scratch.layer",),
          "refinenet": ("scratch.refinenet",), "output_conv1": ("scratch.output_conv1",),
          "tail3x3": ("scratch.output_conv2.0",), "tail1x1": ("scratch.output_conv2.2",)}


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp(min=1e-12))


def bf(t):
    return t.bfloat16().float()


@contextlib.contextmanager
def bf16_group(head, prefixes):
    saved, hooks = [], []
    for name, m in head.named_modules():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)) and any(name.startswith(p) for p in prefixes):
            saved.append((m, m.weight.data))
            m.weight.data = bf(m.weight.data)
            hooks.append(m.register_forward_pre_hook(lambda mod, args: (bf(args[0]),) + tuple(args[1:])))
    try:
        yield len(saved)
    finally:
        for m, w in saved:
            m.weight.data = w
        for h in hooks:
            h.remove()


def main():
    S = int(os.environ.get("OVG_EB_VIEWS", 4))
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    from bench import synth_inputs
    from omnivggt_official_b200 import OmniVGGT
    from oracle.vendor_ref import import_reference_zip
    with torch.device(dev):
        model = OmniVGGT(init_seed=None)
    model.randomize_(seed=0)
    model.eval()
    model.use_cuda_graph = False
    Ref = import_reference_zip()
    with torch.device(dev):
        ref = Ref()
    ref.load_state_dict(model.state_dict(), strict=True)
    ref = ref.to(dev).eval()
    inp = {k: v.to(dev) for k, v in synth_inputs(1, S, seed=1).items()}
    images = inp["images"]
    out = {}
    with torch.no_grad():
        toks, psi = ref.aggregator(images=images, extrinsics=inp["extrinsics"], intrinsics=inp["intrinsics"],
                                   depth=inp["depth"], mask=inp["mask"], depth_gt_index=[], camera_gt_index=[])
        toks = [t.float() for t in toks]

        def heads(tl):
            d, dc = ref.depth_head(tl, images=images, patch_start_idx=psi)
            p, pc = ref.point_head(tl, images=images, patch_start_idx=psi)
            return {"depth": d, "depth_conf": dc, "world_points": p, "world_points_conf": pc}

        truth = heads(toks)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yard = ref(**inp, depth_gt_index=[], camera_gt_index=[])
        out["yard"] = {k: rel(yard[k], truth[k]) for k in KEYS}
        ours = model(images=images)
        out["ours"] = {k: rel(ours[k], truth[k]) for k in KEYS}

        # our aggregator -> reference fp32 heads
        eng = model.engine()
        H, W = images.shape[-2:]
        K = S
        layers = list(model.dpt_layers)
        from omnivggt_official_b200 import torch_parts as TP
        from omnivggt_official_b200.model import _RESNET_MEAN, _RESNET_STD
        P = (H // 14) * (W // 14)
        pos = TP.dino_pos_embed(model.aggregator.patch_embed, P, H, W, 14).float()
        patch = eng.dino_patchify(images.float().view(K, 3, H, W), pos, _RESNET_MEAN, _RESNET_STD)
        inj = TP.injection_vectors(eng.inj_pack, None, [], 1, S, None)
        slots, _ = eng.aggregate(patch, inj, None, None, [], 1, S, H, W, set(layers))
        mix = list(toks)
        for li in layers:
            mix[li] = slots[li].float().view(1, S, -1, slots[li].shape[-1])
        out["agg_only"] = {k: rel(v, truth[k]) for k, v in heads(mix).items()}
        out["agg_tokens_rel_l2"] = {str(li): rel(mix[li], toks[li]) for li in layers}

        # reference tokens -> our heads
        rs = {li: toks[li].view(K, -1, toks[li].shape[-1]).bfloat16().contiguous() for li in layers}
        eng.warm_tables(H, W)
        d = eng.dpt("depth_head", rs, layers, K, H, W, head_act=0)
        p = eng.dpt("point_head", rs, layers, K, H, W, head_act=1)
        mine = {"depth": d[0].view(1, K, H, W, 1), "depth_conf": d[1].view(1, K, H, W),
                "world_points": p[0].view(1, K, H, W, 3), "world_points_conf": p[1].view(1, K, H, W)}
        out["heads_only"] = {k: rel(mine[k], truth[k]) for k in KEYS}

        snap = list(toks)
        for li in layers:
            snap[li] = bf(toks[li])
        out["snap_bf16"] = {k: rel(v, truth[k]) for k, v in heads(snap).items()}

        for g, pre in GROUPS.items():
            with bf16_group(ref.depth_head, pre) as n1, bf16_group(ref.point_head, pre) as n2:
                out[f"sim:{g}"] = {k: rel(v, truth[k]) for k, v in heads(toks).items()}
                out[f"sim:{g}"]["convs"] = n1 + n2
        allp = tuple(p for v in GROUPS.values() for p in v)
        with bf16_group(ref.depth_head, allp), bf16_group(ref.point_head, allp):
            out["sim:all+snap"] = {k: rel(v, truth[k]) for k, v in heads(snap).items()}
        torch.backends.cudnn.allow_tf32 = True
        out["tf32"] = {k: rel(v, truth[k]) for k, v in heads(toks).items()}
        torch.backends.cudnn.allow_tf32 = False
    for k, v in out.items():
        print(k, json.dumps(v))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "error_budget.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
