"""LayerNorm A/B (GPU box): the aggregator shape (10 992 x 1024 fp32 -> bf16, affine) and the DPT shape (8 x 1369 of 1374 rows x 2048
bf16 -> fp16, row gather), L2 flushed between launches and back to back; OVG_LIB_PATH selects the library build."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnivggt_official_b200 import ops

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, iters=20, cold=True):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if cold:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


M, C = 10992, 1024
x = torch.randn(M, C, device=dev)
out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
w, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
res = {"tag": sys.argv[1] if len(sys.argv) > 1 else "", "agg_cold_us": timeit(lambda: ops.layernorm(x, out, w, b)),
       "agg_warm_us": timeit(lambda: ops.layernorm(x, out, w, b), cold=False)}
# back to back inside a CUDA graph, rotating over 8 input buffers (360 MB > L2): what a launch costs inside the captured forward
xs8 = [torch.randn(M, C, device=dev) for _ in range(8)]
os8 = [torch.empty(M, C, device=dev, dtype=torch.bfloat16) for _ in range(8)]
for i in range(8):
    ops.layernorm(xs8[i], os8[i], w, b)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for rep in range(4):
        for i in range(8):
            ops.layernorm(xs8[i], os8[i], w, b)
res["agg_graph_us"] = timeit(g.replay, iters=10, cold=False) / 32
res["agg_graph_gbs"] = M * C * 6 / res["agg_graph_us"] / 1e3
del xs8, os8
ref = torch.nn.functional.layer_norm(x, (C,), w, b, 1e-5)
res["agg_rel"] = float((out.float() - ref).norm() / ref.norm())
T, P, K, C2 = 1374, 1369, 8, 2048
xs = torch.randn(K * T, C2, device=dev).bfloat16()
o2 = torch.empty(K * P, C2, device=dev, dtype=torch.float16)
res["dpt_cold_us"] = timeit(lambda: ops.layernorm(xs, o2, None, None, 1e-5, grp_out=P, grp_in=T, grp_off=5))
ref2 = torch.nn.functional.layer_norm(xs.float().view(K, T, C2)[:, 5:], (C2,)).reshape(K * P, C2)
res["dpt_rel"] = float((o2.float() - ref2).norm() / ref2.norm())
res["agg_cold_gbs"] = M * C * 6 / res["agg_cold_us"] / 1e3
print(json.dumps(res))
