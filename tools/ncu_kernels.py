"""The block-level kernels at cfg2 shapes, one launch each inside a cudaProfiler range (for `ncu --set full`):
   0 qkv fused (LN + RoPE epilogue)   1 proj (residual)   2 fc1 (GELU)   3 fc2 (residual)   4 global attention   5 layernorm
   6 DPT tail (3x3 conv 128->32 + 1x1 + activations, 8 frames @ 518^2)   7 bilinear upsampling 296^2 -> 518^2 (128 ch)
   (6 + 7 = the two-kernel tail; the product runs the fused tail:)   7b tail_tables + fusedtail_kernel (fp16)   4b frame attention
and the modality gather / scatter kernels at the cfg5 shapes (24 views, 6 depth views):
   8 assemble_tokens   9 inject_snapshot (with bf16 slot)   10 depth_stats   11 depth_im2col   12 image_im2col
NCU_ONLY=attn restricts the profiled range to the global attention.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnivggt_official_b200 import ops  # noqa: E402

BF16 = torch.bfloat16
dev = "cuda"
S, T, C = 8, 1374, 1024
M = S * T
g = torch.Generator(device=dev).manual_seed(0)


def rn(*s):
    return torch.randn(*s, device=dev, generator=g)


a = rn(M, C).to(BF16)
h = rn(M, 4 * C).to(BF16)
wqkv, bqkv = (rn(3 * C, C) * C ** -0.5).to(BF16), rn(3 * C)
wproj, bproj = (rn(C, C) * C ** -0.5).to(BF16), rn(C)
w1, b1 = (rn(4 * C, C) * C ** -0.5).to(BF16), rn(4 * C)
w2, b2 = (rn(C, 4 * C) * (4 * C) ** -0.5).to(BF16), rn(C)
gamma = rn(C)
x = rn(M, C)
ones, zeros = torch.ones(64, device=dev), torch.zeros(64, device=dev)
cos, sin = ops.rope_tables(38, dev)
q = torch.empty(1, 16, M, 64, device=dev, dtype=BF16)
k, v = torch.empty_like(q), torch.empty_like(q)
o = torch.empty(1, M, C, device=dev, dtype=BF16)
hid = torch.empty(M, 4 * C, device=dev, dtype=BF16)
ln_w, ln_b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
ln_out = torch.empty(M, C, device=dev, dtype=BF16)


Fr, hh, ww = 8, 518, 518
xt = torch.zeros(Fr, hh + 2, ww + 2, 128, device=dev, dtype=BF16)
xt[:, 1:-1, 1:-1] = torch.randn(Fr, hh, ww, 128, device=dev, generator=g).to(BF16)
wt = (rn(32, 9 * 128) * (9 * 128) ** -0.5).to(BF16)
bt, w2t, b2t = rn(32) * 0.1, (rn(4, 32) * 32 ** -0.5).contiguous(), rn(4) * 0.1
preds = torch.zeros(Fr, hh, ww, 3, device=dev)
conf = torch.zeros(Fr, hh, ww, device=dev)
taps = [(ky - 1) * (ww + 2) + (kx - 1) for ky in range(3) for kx in range(3)]
xs = torch.zeros(Fr, 298, 298, 128, device=dev, dtype=BF16)
xs[:, 1:-1, 1:-1] = torch.randn(Fr, 296, 296, 128, device=dev, generator=g).to(BF16)
xu = torch.empty(Fr, hh + 2, ww + 2, 128, device=dev, dtype=BF16)
tx, ty = rn(ww, 64), rn(hh, 64)


# ---- modality gather / scatter at cfg5 shapes: K = 24 frames, depth aux on 6 of them
K5, P5 = 24, 1369
patch5 = rn(K5, P5, C)
x5 = torch.empty(K5, T, C, device=dev)
cam_tok, reg_tok, inj0, placeholder = rn(2, C), rn(2, 4, C), rn(K5, C), rn(C)
has_depth = torch.zeros(K5, dtype=torch.int32, device=dev)
didx = [0, 3, 4, 9, 15, 22]
has_depth[didx] = 1
slot5 = torch.empty(K5 * T, 2 * C, device=dev, dtype=BF16)
depth5 = 0.5 + 4 * torch.rand(1, K5, 518, 518, device=dev, generator=g)
mask5 = (torch.rand(1, K5, 518, 518, device=dev, generator=g) > 0.2).float()
idx5 = torch.tensor(didx, dtype=torch.int32, device=dev)
scratch5 = torch.zeros(ops.L.DEPTH_SCRATCH_DOUBLES(1), dtype=torch.float64, device=dev)
cols5 = torch.empty(len(didx) * P5, 392, device=dev, dtype=BF16)
img5 = torch.rand(K5, 3, 518, 518, device=dev, generator=g)
icol5 = torch.empty(K5 * P5, 592, device=dev, dtype=BF16)
xs16, wt16 = xs.to(torch.float16), wt.to(torch.float16)
qf = (torch.randn(8, 16, T, 64, device=dev, generator=g) * 0.18).to(BF16)
kf, vf = torch.randn_like(qf), torch.randn_like(qf)
of = torch.empty(8, T, C, device=dev, dtype=BF16)
att_scratch = ops.attention_scratch(dev)
q.copy_((torch.randn(1, 16, M, 64, device=dev, generator=g) * 0.18).to(BF16))
k.copy_(torch.randn(1, 16, M, 64, device=dev, generator=g).to(BF16))
v.copy_(torch.randn(1, 16, M, 64, device=dev, generator=g).to(BF16))
ONLY = os.environ.get("NCU_ONLY", "")


def run():
    if ONLY == "attn":
        ops.attention(q, k, v, o, 1, 16, M)
        return
    if ONLY != "scatter":
        run_blocks()
    ops.assemble_tokens(x5, patch5, cam_tok, reg_tok, inj0, placeholder, has_depth, K5, K5, T, 4, C)
    ops.inject_snapshot(x5.view(K5 * T, C), inj0, slot5, None, K5, T, C, 0)
    ops.depth_im2col(depth5, mask5, idx5, scratch5, cols5, 1, K5, len(didx), 518, 518, 14)
    ops.image_im2col(img5, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), icol5, K5, 518, 518, 14)


def run_blocks():
    ops.qkv_proj(a, wqkv, bqkv, ones, zeros, ones, zeros, q, k, v, ntok=M, T=T, nspecial=5, wp=37, rope_cos=cos, rope_sin=sin)
    ops.linear_resid(a, wproj, bproj, gamma, x)
    ops.linear_bf16(a, w1, b1, act=ops.L.ACT_GELU, out=hid)
    ops.linear_resid(h, w2, b2, gamma, x)
    ops.attention(q, k, v, o, 1, 16, M)
    ops.layernorm(x, ln_out, ln_w, ln_b)
    ops.gemm(xt.reshape(-1, 128), wt, taps=taps, epi=ops.L.EPI_HEADTAIL, bias=bt, w2=w2t, b2=b2t, outc=4, head_act=1,
             preds=preds, conf=conf, rowmap=ops.L.ROWS_PAD, gh=hh, gw=ww)
    ops.upsample_bilinear(xs, xu, tx, ty, Fr, 296, 296, hh, ww, 128)
    ops.dpt_tail(xs16, tx, ty, wt16, bt, w2t, b2t, 1, Fr, 296, 296, hh, ww)
    ops.attention(qf, kf, vf, of, 8, 16, T)
    ops.attention(q, k, v, o, 1, 16, M, scratch=att_scratch)       # global attention with the KV-split tail tiles + merge


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
