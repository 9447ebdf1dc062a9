"""Full-size parity pins: run the UNMODIFIED reference ``OmniVGGT`` (24 + 24 aggregator blocks, DINOv2 ViT-L patchifier,
camera head, both DPT heads; reference omnivggt/models/omnivggt.py:10-68) once per case at 518 x 518 on CPU fp32 and store
its outputs (build container only; TEST INFRASTRUCTURE).

    python oracle/make_golden_full.py            # ~6 min, ~12 GB of host memory

Cases (BASELINE.json configs):
  full_cfg1      configs[0]: 1 scene x 4 views @ 518^2, images only
  full_aux_s3    configs[2]-shaped: 1 scene x 3 views @ 518^2, depth + camera aux on every view
  full_mixed_s5  configs[4]-shaped: 1 scene x 5 views @ 518^2, partial depth_gt_index / camera_gt_index

Weights are oracle/synth.make_state_dict(schema, seed 0) over the reference's own 1 505-key schema (stored as
tests/golden/full.schema.json, names + shapes only), so the GPU tests rebuild the identical state dict without the
reference.  Dense outputs are stored on a pixel lattice (every STRIDE-th row / column, fp32): the fixtures stay small
and the comparison still covers every frame, every output channel and the whole image extent.
"""
from __future__ import annotations

import json
import os
import sys
import time

import torch
from safetensors.torch import save_file

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_shims import import_reference  # noqa: E402
from oracle.synth import make_inputs, make_state_dict  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
STRIDE = 7   # 518 = 74 * 7

# name -> (S, depth_gt_index, camera_gt_index, input seed)
CASES = {
    "full_cfg1": (4, [], [], 11),
    "full_aux_s3": (3, [0, 1, 2], [0, 1, 2], 12),
    "full_mixed_s5": (5, [0, 3], [0, 2, 4], 13),
}


def lattice(t: torch.Tensor) -> torch.Tensor:
    """[B,S,H,W,...] -> every STRIDE-th pixel (offset STRIDE // 2)."""
    o = STRIDE // 2
    return t[:, :, o::STRIDE, o::STRIDE].contiguous().clone()


def main():
    import_reference()
    from omnivggt.models.omnivggt import OmniVGGT as RefOmniVGGT
    torch.set_num_threads(os.cpu_count() or 8)
    t0 = time.time()
    m = RefOmniVGGT().eval()
    schema = {k: list(t.shape) for k, t in m.state_dict().items()}
    sd = make_state_dict(schema, seed=0)
    m.load_state_dict(sd, strict=True)
    del sd
    with open(os.path.join(GOLDEN, "full.schema.json"), "w") as f:
        json.dump({"schema": schema}, f, indent=0, sort_keys=True)
    print(f"reference OmniVGGT built: {len(schema)} tensors, {time.time() - t0:.0f} s", flush=True)
    index = {}
    for name, (S, didx, cidx, seed) in CASES.items():
        inp = make_inputs(1, S, 518, 518, seed=seed)
        t0 = time.time()
        with torch.no_grad():
            out = m(images=inp["images"], extrinsics=inp["extrinsics"], intrinsics=inp["intrinsics"], depth=inp["depth"],
                    mask=inp["mask"], depth_gt_index=list(didx), camera_gt_index=list(cidx))
        dt = time.time() - t0
        store = {"pose_enc": out["pose_enc"].contiguous().clone()}
        for i, p in enumerate(out["pose_enc_list"]):
            store[f"pose_enc_list.{i}"] = p.contiguous().clone()
        for k in ("depth", "depth_conf", "world_points", "world_points_conf"):
            store[k] = lattice(out[k].float())
        save_file(store, os.path.join(GOLDEN, f"{name}.safetensors"))
        stats = {k: [float(v.abs().mean()), float(v.abs().max())] for k, v in store.items() if "list" not in k}
        index[name] = dict(S=S, H=518, W=518, depth_gt_index=didx, camera_gt_index=cidx, input_seed=seed, weight_seed=0,
                           stride=STRIDE, cpu_forward_s=round(dt, 1), stats=stats)
        print(name, f"{dt:.0f} s", stats, flush=True)
    with open(os.path.join(GOLDEN, "full_index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
