"""Context parallelism for ONE scene with many views: the views are sharded over the GPUs of a node,
every rank runs the per-token work of its own views, and the single SDPA over all views of the reference's global blocks
(models/aggregator.py:312-341) becomes: K / V rows stored straight into every rank's full-length buffer by the QKV GEMM epilogue
(peer-mapped memory, NVLink) -> flag barrier -> the rank's queries attend to all keys; the S camera-token rows the camera head
needs travel the same way.  torch.distributed is only plumbing here: it provides the peer-mapped allocation (symmetric memory) at
set-up; a forward makes no collective call, so it can be captured into a CUDA graph like the single-GPU forward.  The reference has no distributed runtime: this replaces nothing, it is new."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib as L


class ContextParallel:
    def __init__(self, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("context parallelism spans the GPUs of one NVSwitch node (at most 8 ranks)")
        self.device = device
        self._key = None
        self._desc: Optional[L.ContextParallelDesc] = None
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)

    def desc(self, heads: int, ntok_total: int, views_total: int) -> L.ContextParallelDesc:
        """Peer-mapped K / V double buffers, camera-token gather buffer and barrier flags for a scene of `ntok_total` tokens;
        cached per size."""
        key = (heads, ntok_total, views_total)
        if self._key != key:
            import torch.distributed._symmetric_memory as symm_mem
            one = heads * ntok_total * 64 * 2                       # one bf16 [heads, ntok_total, 64] buffer
            one = (one + 255) // 256 * 256
            cam = (views_total * 2 * heads * 64 * 4 + 255) // 256 * 256      # fp32 [views_total, 2C]
            total = 4 * one + cam + 256                             # K0 K1 V0 V1 | camera tokens | flags
            buf = symm_mem.empty(total, dtype=torch.uint8, device=self.device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, self.group)
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)                                # every rank has zeroed its flags before anyone signals
            d = L.ContextParallelDesc()
            d.rank, d.world = self.rank, self.world
            for r in range(self.world):
                base = int(hdl.buffer_ptrs[r])
                d.k_peers[0][r], d.k_peers[1][r] = base, base + one
                d.v_peers[0][r], d.v_peers[1][r] = base + 2 * one, base + 3 * one
                d.cam_peers[r] = base + 4 * one
                d.flag_peers[r] = base + 4 * one + cam
            d.epoch_counter = self.epoch.data_ptr()
            d.views_total = views_total
            self.cam_all = buf[4 * one:4 * one + views_total * 2 * heads * 64 * 4].view(torch.float32).view(views_total, 2 * heads * 64)
            self._buf, self._hdl, self._desc, self._key = buf, hdl, d, key
        return self._desc

    def local_views(self, S: int):
        if S % self.world:
            raise ValueError(f"{S} views cannot be split evenly over {self.world} ranks")
        n = S // self.world
        return self.rank * n, n

    def local_indices(self, idx: List[int], S: int) -> List[int]:
        v0, n = self.local_views(S)
        return [i - v0 for i in idx if v0 <= i < v0 + n]
