"""One full-size forward (cfg2: 8 views @ 518^2) inside a cudaProfiler range, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/profile_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnivggt_official_b200 import OmniVGGT  # noqa: E402

S = int(os.environ.get("S", 8))
with torch.device("cuda"):
    m = OmniVGGT(init_seed=None)
m.randomize_(0).eval()
img = torch.rand(1, S, 3, 518, 518, device="cuda")
for _ in range(int(os.environ.get("WARM", 2))):
    m(images=img)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = m(images=img)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(out["depth"].mean()))
