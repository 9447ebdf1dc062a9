"""Alternate the device-resident loop and the streaming pipeline (pipeline.py) on one model: how much of the host<->device copies
is hidden?  (GPU box; diagnostic.)"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synth_inputs, OUT_KEYS
from omnivggt_official_b200 import OmniVGGT
from omnivggt_official_b200.pipeline import StreamingPipeline

dev = torch.device("cuda", 0)
with torch.device(dev):
    m = OmniVGGT(init_seed=None)
m.randomize_(seed=0)
m.eval()
host = {"images": synth_inputs(1, 8, seed=1)["images"].pin_memory()}
devin = {k: v.to(dev) for k, v in host.items()}
pipe = StreamingPipeline(m, slots=int(os.environ.get("SLOTS", 2)), out_keys=OUT_KEYS)
hout = None


def resident(n):
    for _ in range(n):
        m(**devin)


def inline(n):
    global hout
    for _ in range(n):
        o = m(**{k: v.to(dev, non_blocking=True) for k, v in host.items()})
        if hout is None:
            hout = {k: torch.empty(o[k].shape, dtype=o[k].dtype).pin_memory() for k in OUT_KEYS}
        for k in OUT_KEYS:
            hout[k].copy_(o[k], non_blocking=True)
        torch.cuda.current_stream().synchronize()


def piped(n):
    pend = []
    for _ in range(n):
        pend.append(pipe.submit(host))
        if len(pend) > 1:
            pipe.result(pend.pop(0))
    pipe.result(pend.pop(0))
    pipe.drain()


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    fn(n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


for f in (resident, inline, piped):
    f(4)
for rnd in range(3):
    for f in (resident, inline, piped):
        ev, wall = timed(f, 20)
        print(f"round {rnd} {f.__name__:9s} {ev:7.3f} ms/step (events)  {wall:7.3f} ms/step (host clock)", flush=True)
