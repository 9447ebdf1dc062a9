"""Scene-level data parallelism: one process per GPU, a full weight replica per rank obtained
by ONE broadcast of a flat parameter arena from rank 0 (NCCL over NVLink on the GPU box, gloo in the CPU tests), then no
data-path communication: scenes (view-sets) are independent (B is a pure batch dim in the reference,
models/aggregator.py:317-318), so rank r simply takes scenes r, r+world, ...  The reference has no distributed runtime
(SURVEY.md section 2.3); this replaces nothing, it is new."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_scenes(num_scenes: int, rank: int, world: int) -> List[int]:
    return list(range(rank, num_scenes, world))


@torch.no_grad()
def broadcast_weights(module: torch.nn.Module, src: int = 0, bucket_bytes: int = 64 << 20) -> int:
    """Broadcast all parameters and buffers from `src`.  Tensors of 1 MiB and more (the weight matrices: > 99% of the bytes)
    are broadcast IN PLACE, one message each -- no staging copy, no transient memory; the many small ones (biases, norms,
    LayerScale, tokens) travel in flat same-dtype buckets so that NVSwitch launch latency is paid once per bucket.
    Packed kernel-layout replicas (engine, CUDA graphs) are invalidated afterwards.  Returns the bytes broadcast."""
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    total = 0
    small = {}
    for t in tensors:
        nbytes = t.numel() * t.element_size()
        if nbytes >= (1 << 20) and t.is_contiguous():
            dist.broadcast(t, src=src)
            total += nbytes
        else:
            small.setdefault(t.dtype, []).append(t)
    for dtype, ts in small.items():
        bucket, size = [], 0

        def flush():
            nonlocal bucket, size, total
            if not bucket:
                return
            flat = torch.cat([t.reshape(-1) for t in bucket])
            dist.broadcast(flat, src=src)
            off = 0
            for t in bucket:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
            total += flat.numel() * flat.element_size()
            bucket, size = [], 0
        for t in ts:
            bucket.append(t)
            size += t.numel() * t.element_size()
            if size >= bucket_bytes:
                flush()
        flush()
    if hasattr(module, "_invalidate"):
        module._invalidate()          # in-place parameter edits make the packed bf16 replicas / captured graphs stale
    return total


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
