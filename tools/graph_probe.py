"""Does whole-forward CUDA-graph replay beat eager launch (1000+ launches per step)?  cfg2 shapes."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnivggt_official_b200 import OmniVGGT
with torch.device("cuda"):
    m = OmniVGGT(init_seed=None)
m.randomize_(0).eval()
img = torch.rand(1, 8, 3, 518, 518, device="cuda")
for _ in range(3):
    out = m(images=img)
torch.cuda.synchronize()
def timed(fn, n=10):
    torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s.record()
    for _ in range(n): fn()
    e.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n, (t1 - t0) / n * 1e3
gpu_ms, cpu_ms = timed(lambda: m(images=img))
print(f"eager: {gpu_ms:.2f} ms/step (GPU events), CPU issue time {cpu_ms:.2f} ms/step")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    m(images=img)
torch.cuda.current_stream().wait_stream(s)
try:
    with torch.cuda.graph(g):
        gout = m(images=img)
    gpu_ms2, cpu_ms2 = timed(g.replay)
    print(f"graph: {gpu_ms2:.2f} ms/step, CPU {cpu_ms2:.2f} ms/step")
    ref = m(images=img)
    g.replay(); torch.cuda.synchronize()
    print("graph == eager:", all(torch.equal(gout[k], ref[k]) for k in ("depth", "world_points", "pose_enc")))
except Exception as ex:
    print("graph capture failed:", repr(ex)[:500])
