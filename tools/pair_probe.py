"""Isolate the tensor-pipe rate of the single-CTA (block_n 256) vs CTA-pair (block_n 512) GEMM kernels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnivggt_official_b200 import ops
BF16 = torch.bfloat16
def t(fn, it=20):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it
for (M, N, K) in ((256, 256, 8192), (256, 256, 65536 // 4), (10992, 3072, 1024)):
    a = torch.randn(M, K, device="cuda").to(BF16); w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(BF16)
    out = torch.empty(M, N, device="cuda", dtype=BF16)
    for bn in (256, 512):
        ms = t(lambda: ops.linear_bf16(a, w, None, out=out, block_n=bn))
        kb = K // 64
        print(f"M={M} N={N} K={K} bn={bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TF/s   per-kblock {ms*1e-3*1.9e9/kb/max(1,(M//256)*(N//256)//74 if M>256 else 1):.0f} cyc@1.9GHz")
