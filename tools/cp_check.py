"""Context-parallel check / benchmark, one process per GPU (torchrun):
    torchrun --nproc-per-node N tools/cp_check.py [--full] [--views S]
Every rank builds the same model, runs the scene once WITHOUT context parallelism (all views on its own GPU) and once WITH the
views sharded over the ranks, and compares the dense outputs of its views bit for bit (the per-row arithmetic is identical: same
GEMM / attention kernels, same KV order).  --full: the full architecture at 518 x 518, with CUDA-event timing of both modes."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnivggt_official_b200 import OmniVGGT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--full", action="store_true")
ap.add_argument("--views", type=int, default=0)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)

if args.full:
    H = W = 518
    S = args.views or 2 * world
    with torch.device(dev):
        m = OmniVGGT(init_seed=None)
else:
    H, W = 56, 70
    S = args.views or 2 * world
    with torch.device(dev):
        m = OmniVGGT(img_size=56, embed_dim=128, depth=4, patch_embed="dino", dino_depth=2, dino_heads=2, dpt_features=128,
                     dpt_out_channels=(64, 128, 256, 256), dpt_layers=(0, 1, 2, 3), camera_heads=2, camera_trunk_depth=2,
                     init_seed=None)
m.randomize_(0).eval()          # same seed on every rank: identical replicas without a broadcast
m.use_cuda_graph = False
g = torch.Generator().manual_seed(7)
images = torch.rand(1, S, 3, H, W, generator=g).to(dev)
q, r = torch.linalg.qr(torch.randn(S, 3, 3, generator=g))
q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1))[:, None, :]
q[:, :, 0] = q[:, :, 0] * torch.linalg.det(q)[:, None]
extr = torch.cat([q, torch.randn(S, 3, 1, generator=g)], -1)[None].to(dev)
intr = torch.zeros(1, S, 3, 3)
intr[..., 0, 0] = intr[..., 1, 1] = 500.0 * W / 518
intr[..., 0, 2], intr[..., 1, 2], intr[..., 2, 2] = W / 2, H / 2, 1.0
intr = intr.to(dev)
mask = (torch.rand(1, S, H, W, generator=g) > 0.2).float()
depth = ((0.5 + 4 * torch.rand(1, S, H, W, 1, generator=g)) * mask[..., None]).to(dev)
mask = mask.to(dev)
didx = [i for i in range(S) if i % 3 != 1]
cidx = [0] + [i for i in range(1, S) if i % 2 == 0]
kw = dict(images=images, extrinsics=extr, intrinsics=intr, depth=depth, mask=mask, depth_gt_index=didx, camera_gt_index=cidx)


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, float(t.item())


from omnivggt_official_b200 import _lib
lib = _lib.lib()
# as shipped (long attention sequences split the tiles of their last CTA wave over the keys): timing, agreement within 2e-3
ref_s, ms_single = timed(lambda: m(**kw), args.steps if args.full else 1)
# bit-identity: with one CTA per tile the summation order of a tile does not depend on the number of tiles in the launch
lib.ovg_runtime_attention_split(0)
ref = m(**kw)
m.enable_context_parallel()
lib.ovg_runtime_attention_split(1)
out_s, ms_cp = timed(lambda: m(**kw), args.steps if args.full else 1)
lib.ovg_runtime_attention_split(0)
out = m(**kw)
v0, v1 = out["view_range"]
res = {"rank": rank, "world": world, "views": S, "view_range": [v0, v1], "ms_single_gpu": ms_single, "ms_context_parallel": ms_cp}
ok = True
res["split_attention_rel_l2"] = {k: float((out_s[k].float() - ref_s[k][:, v0:v1].float()).norm() / ref_s[k][:, v0:v1].float().norm())
                                 for k in ("depth", "world_points")}
ok &= all(v < 5e-3 for v in res["split_attention_rel_l2"].values())     # two roundings of the same sums, amplified by 48 blocks
for k in ("depth", "depth_conf", "world_points", "world_points_conf"):
    same = torch.equal(out[k], ref[k][:, v0:v1])
    rel = float((out[k].float() - ref[k][:, v0:v1].float()).norm() / ref[k][:, v0:v1].float().norm())
    res[k] = {"bit_identical": bool(same), "rel_l2": rel}
    ok &= rel < 1e-6
same = torch.equal(out["pose_enc"], ref["pose_enc"])
res["pose_enc"] = {"bit_identical": bool(same), "rel_l2": float((out["pose_enc"] - ref["pose_enc"]).norm() / ref["pose_enc"].norm())}
ok &= res["pose_enc"]["rel_l2"] < 1e-6
# the same under CUDA-graph replay (third call of a signature captures; a forward makes no collective call)
m.use_cuda_graph = True
for _ in range(3):
    m(**kw)
outg, ms_cpg = timed(lambda: m(**kw), args.steps if args.full else 2)
res["ms_context_parallel_graph"] = ms_cpg
res["graph_replay_bit_identical"] = bool(all(torch.equal(outg[k], ref[k][:, v0:v1]) for k in ("depth", "world_points")) and
                                         torch.equal(outg["pose_enc"], ref["pose_enc"]))
ok &= res["graph_replay_bit_identical"]
lib.ovg_runtime_attention_split(1)
m._graphs = {}                  # the captured graphs carry the switch: time the shipped configuration
for _ in range(3):
    m(**kw)
_, ms_cpg = timed(lambda: m(**kw), args.steps if args.full else 2)
res["ms_context_parallel_graph"] = ms_cpg
res["ok"] = bool(ok)
gathered = [None] * world
dist.all_gather_object(gathered, res)
if rank == 0:
    for r_ in gathered:
        print(json.dumps(r_))
    print("CP_CHECK", "PASS" if all(r_["ok"] for r_ in gathered) else "FAIL",
          f"eager: single {ms_single:.2f} ms, cp {ms_cp:.2f} ms ({ms_single / ms_cp:.2f}x); cp graph replay {ms_cpg:.2f} ms on {world} GPUs")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
