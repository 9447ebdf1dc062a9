#!/usr/bin/env python
"""Benchmark of the OmniVGGT hot path (BASELINE.json metric: view-sets/sec, N-view 518^2 batches).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one full OmniVGGT.forward over one view-set per GPU.  N = 1 workload = BASELINE configs[1]
("cfg2": 1 scene x 8 views @ 518 x 518, images-only, bf16 kernels, random-init weights of the full architecture).
N > 1 (torchrun): weak scaling, one independent view-set per rank per step, weights broadcast once from rank 0 over NCCL.
Prints ONE JSON line on rank 0 (see DESIGN.md section "Measurement" for the field definitions).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S_VIEWS, IMG = 8, 518
T_TOK = (IMG // 14) ** 2 + 5
WORKLOAD = f"cfg2: 1 scene x {S_VIEWS} views @ {IMG}x{IMG}, images-only, per GPU per step (BASELINE.json configs[1])"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1393.7), d.get("hbm_gbs", 6489.9), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.p = index, None

    def __enter__(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None
        time.sleep(0.25)
        return self

    def __exit__(self, *a):
        self.out = ""
        if self.p is not None:
            time.sleep(0.15)
            self.p.terminate()
            try:
                self.out = self.p.communicate(timeout=5)[0]
            except Exception:
                self.p.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v == "Active":
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference path on the host cores (the Python reference itself cannot travel to the
    GPU box).  Rank 0 only."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; this arm must use every host core it can get, so the variables
    # are dropped BEFORE torch / MKL / OpenMP initialise (setting the thread count afterwards left MKL single-threaded:
    # 246 s per view-set on a 128-thread box instead of ~70 s) and torch picks its default (one thread per physical core)
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.pop(var, None)
    import torch
    from oracle import cpu_baseline as cb
    for _ in range(max(args.warmup, 0) and 1):          # one warm-up sample is enough to page in MKL / weights
        cb.sample()
    secs = []
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        est, _ = cb.sample()
        secs.append(est)
        if time.perf_counter() - t_begin > 240:          # keep the whole run within a few minutes
            break
    est = statistics.median(secs)
    val = 1.0 / est
    line = {"impl": "reference", "metric": "view_sets_per_sec", "value": val, "unit": "view-sets/s", "n_gpus": args.gpus,
            "steps": len(secs), "warmup": 1 if args.warmup else 0, "ms_per_step": est * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "device": "host CPU"},
            "cpu_baseline": {"value": val, "unit": "view-sets/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": cb.SAMPLE_DESC},
            "e2e": {"value": val, "unit": "view-sets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--views", type=int, default=S_VIEWS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from omnivggt_official_b200 import OmniVGGT, _lib
    from omnivggt_official_b200.dist import broadcast_weights, max_over_ranks

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    S = args.views
    with torch.device(dev):
        model = OmniVGGT(init_seed=None)
    model.randomize_(seed=0 if rank == 0 else 1000 + rank)     # non-zero ranks are overwritten by the broadcast
    bcast_bytes = 0
    if world > 1:
        bcast_bytes = broadcast_weights(model, src=0)
    model.eval()
    eng = model.engine()
    lib = _lib.lib()

    g = torch.Generator().manual_seed(1 + rank)
    host_images = torch.rand(1, S, 3, IMG, IMG, generator=g).pin_memory()
    dev_images = host_images.to(dev)
    out_keys = ("pose_enc", "depth", "depth_conf", "world_points", "world_points_conf")
    host_out = None

    def step_resident():
        return model(images=dev_images)

    def step_e2e():
        nonlocal host_out
        img = host_images.to(dev, non_blocking=True)
        out = model(images=img)
        if host_out is None:
            host_out = {k: torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory() for k in out_keys}
        for k in out_keys:
            host_out[k].copy_(out[k], non_blocking=True)
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        return max_over_ranks(ms, dev) if world > 1 else ms

    for _ in range(args.warmup):        # (the third call of a shape captures the CUDA graph that later calls replay)
        step_resident()
    with ClockSampler(local) as cs:
        total_ms = timed(step_resident, args.steps)
    clocks = cs.summary()
    ms_step = total_ms / args.steps
    value = world * 1e3 / ms_step

    for _ in range(args.warmup):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps) / args.steps

    # Kernel-level pass: the product path replays a CUDA graph, inside which single launches cannot be bracketed by
    # events or counted by the library, so the same K steps are run once more with eager launches to time the 24
    # global-attention launches per step and to count libovg launches per step.
    graph_mode = model.use_cuda_graph
    model.use_cuda_graph = False
    step_resident()
    eng.attn_events = []
    l0 = lib.ovg_launch_count()
    timed(step_resident, args.steps)
    launches = lib.ovg_launch_count() - l0
    events, eng.attn_events = eng.attn_events, None
    model.use_cuda_graph = graph_mode
    h2d = host_images.numel() * host_images.element_size()
    d2h = sum(t.numel() * t.element_size() for t in host_out.values())

    # ---- roofline of the dominant kernel: global attention (24 launches / step), timed live with CUDA events
    peak_tf, peak_hbm, peak_src = measured_peaks()
    att_ms = [a.elapsed_time(b) for a, b, _, _ in events]
    L = S * T_TOK
    att_flops = 4.0 * L * L * 1024                    # 4 L^2 C per launch (QK^T + PV, 16 heads x 64)
    att_avg = sum(att_ms) / max(len(att_ms), 1)
    achieved = att_flops / (att_avg * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "attn_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"S{S}")
    roofline = {"kernel": "ovg::attn1_kernel (global attention)", "bound": "tensor", "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic, "peak_source": peak_src,
                "launches_timed": len(att_ms), "avg_launch_ms": att_avg,
                "share_of_step": sum(att_ms) / args.steps / ms_step,
                "timed_in": "separate eager pass of the same K steps (launches inside the replayed CUDA graph cannot be bracketed)"}

    line = {"metric": "view_sets_per_sec", "value": value, "unit": "view-sets/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD if S == S_VIEWS else f"1 scene x {S} views @ {IMG}x{IMG} per GPU per step",
                       "views": S, "parallelism": f"dp{world} (scene-sharded, NCCL weight broadcast {bcast_bytes} B at start-up)",
                       "weights": "random-init, full architecture (1217.5 M params)",
                       "l2": "no flush needed: each step streams >2 GB of weights+activations, far beyond the 126 MB L2",
                       "dino": "frozen DINOv2 patchifier on the libovg kernels",
                       "launch": "CUDA graph replay" if model.use_cuda_graph else "eager"},
            "clocks": clocks,
            "e2e": {"value": world * 1e3 / e2e_ms, "unit": "view-sets/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms},
            "gpu_launches": int(launches), "roofline": roofline}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline as cb
        import torch as _t
        est, parts = cb.sample()
        line["cpu_baseline"] = {"value": 1.0 / est, "unit": "view-sets/s", "cores": _t.get_num_threads(), "kind": "port",
                                "sample": cb.SAMPLE_DESC, "seconds_per_view_set": est}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
