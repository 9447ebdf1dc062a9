"""The block-level kernels at cfg2 shapes, one launch each inside a cudaProfiler range (for `ncu --set full`):
   0 qkv fused (LN + RoPE epilogue)   1 proj (residual)   2 fc1 (GELU)   3 fc2 (residual)   4 global attention   5 layernorm
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


def run():
    ops.qkv_proj(a, wqkv, bqkv, ones, zeros, ones, zeros, q, k, v, ntok=M, T=T, nspecial=5, wp=37, rope_cos=cos, rope_sin=sin)
    ops.linear_resid(a, wproj, bproj, gamma, x)
    ops.linear_bf16(a, w1, b1, act=ops.L.ACT_GELU, out=hid)
    ops.linear_resid(h, w2, b2, gamma, x)
    ops.attention(q, k, v, o, 1, 16, M)
    ops.layernorm(x, ln_out, ln_w, ln_b)


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
