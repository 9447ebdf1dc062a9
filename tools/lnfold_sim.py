"""Numerical feasibility of folding LayerNorm into the following GEMM (GPU box; diagnostic).  In the packed reference (true fp32
elsewhere) every block's norm1 -> qkv and norm2 -> fc1 is replaced by an emulation of
  baseline:  bf16(LN(x)) @ bf16(W)^T + b                      (what the kernels do today)
  fold:      rstd * (bf16(x) @ bf16(W diag(gamma))^T - mean * colsum) + (W beta + b)
and the deviation of the model outputs from the fp32 forward is compared."""
import json, os, sys
import torch
import torch.nn as nn
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KEYS = ("pose_enc", "depth", "depth_conf", "world_points", "world_points_conf")


def bf(t):
    return t.bfloat16().float()


class Stash(nn.Module):
    def forward(self, x):
        self.x = x
        return x


class Folded(nn.Module):
    def __init__(self, stash, ln, lin, mode):
        super().__init__()
        self.stash, self.ln, self.lin, self.mode = stash, ln, lin, mode

    def forward(self, _):
        x = self.stash.x.float()
        W, b = self.lin.weight.float(), self.lin.bias.float()
        g, be, eps = self.ln.weight.float(), self.ln.bias.float(), self.ln.eps
        if self.mode == "baseline":
            return bf(F.layer_norm(x, x.shape[-1:], g, be, eps)) @ bf(W).t() + b
        Wp = bf(W * g[None])
        mu = x.mean(-1, keepdim=True)
        var = (x * x).mean(-1, keepdim=True) - mu * mu
        rstd = torch.rsqrt(var + eps)
        return rstd * (bf(x) @ Wp.t() - mu * Wp.sum(1)[None]) + (W @ be + b)


def patch(ref, mode):
    saved = []
    blocks = list(ref.aggregator.frame_blocks) + list(ref.aggregator.global_blocks) + list(ref.aggregator.patch_embed.blocks)
    for blk in blocks:
        s1, s2 = Stash(), Stash()
        saved.append((blk, blk.norm1, blk.attn.qkv, blk.norm2, blk.mlp.fc1))
        f1, f2 = Folded(s1, blk.norm1, blk.attn.qkv, mode), Folded(s2, blk.norm2, blk.mlp.fc1, mode)
        blk.norm1, blk.attn.qkv, blk.norm2, blk.mlp.fc1 = s1, f1, s2, f2
    return saved


def unpatch(saved):
    for blk, n1, q, n2, f1 in saved:
        blk.norm1, blk.attn.qkv, blk.norm2, blk.mlp.fc1 = n1, q, n2, f1


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    from bench import synth_inputs
    from omnivggt_official_b200 import OmniVGGT
    from oracle.vendor_ref import import_reference_zip
    with torch.device(dev):
        model = OmniVGGT(init_seed=None)
    model.randomize_(seed=0)
    Ref = import_reference_zip()
    with torch.device(dev):
        ref = Ref()
    ref.load_state_dict(model.state_dict(), strict=True)
    ref = ref.to(dev).eval()
    del model
    inp = {k: v.to(dev) for k, v in synth_inputs(1, int(os.environ.get("OVG_EB_VIEWS", 4)), seed=1).items()}
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp(min=1e-12))
    with torch.no_grad():
        truth = ref(**inp, depth_gt_index=[], camera_gt_index=[])
        out = {}
        for mode in ("baseline", "fold"):
            saved = patch(ref, mode)
            o = ref(**inp, depth_gt_index=[], camera_gt_index=[])
            unpatch(saved)
            out[mode] = {k: rel(o[k], truth[k]) for k in KEYS}
    for k, v in out.items():
        print(k, json.dumps(v))


if __name__ == "__main__":
    main()
