"""Host <-> device streaming around ``OmniVGGT.forward``: the H2D copy of request i+1 and the D2H read of request i-1 run on their
own CUDA streams under the forward of request i (pinned host buffers, events, no host synchronisation until a result is asked for).
This is what a serving loop around reference inference.py:343-390 (inputs .to(device), model(**inputs), predictions .cpu()) needs on a
PCIe-attached GPU: ~77 MB cross the bus per 8-view scene (25.8 MB in, 51.5 MB out), 1.5 ms of a 48 ms forward if left in line."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

OUT_KEYS = ("pose_enc", "depth", "depth_conf", "world_points", "world_points_conf")


class StreamingPipeline:
    def __init__(self, model, slots: int = 2, out_keys=OUT_KEYS):
        self.model, self.slots, self.out_keys = model, slots, tuple(out_keys)
        self.device = next(model.parameters()).device
        self.s_in, self.s_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self._host_out: List[Optional[Dict[str, torch.Tensor]]] = [None] * slots
        self._dev_in: List[Optional[Dict[str, torch.Tensor]]] = [None] * slots
        self._done: List[Optional[torch.cuda.Event]] = [None] * slots
        self._free: List[Optional[torch.cuda.Event]] = [None] * slots     # compute of the slot's previous request has consumed its inputs
        self._n = 0

    def submit(self, host_inputs: Dict[str, torch.Tensor], depth_gt_index=None, camera_gt_index=None) -> int:
        """host_inputs: pinned CPU tensors keyed like forward()'s tensor arguments.  Returns a ticket for result()."""
        k = self._n % self.slots
        if self._done[k] is not None:
            self._done[k].synchronize()          # the slot's pinned output buffers are being handed out again
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.s_in):
            if self._free[k] is not None:
                self.s_in.wait_event(self._free[k])
            if self._dev_in[k] is None or any(self._dev_in[k][n].shape != t.shape for n, t in host_inputs.items()):
                self._dev_in[k] = {n: torch.empty(t.shape, dtype=t.dtype, device=self.device) for n, t in host_inputs.items()}
            for n, t in host_inputs.items():
                self._dev_in[k][n].copy_(t, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.s_in)
        main.wait_event(ready)
        out = self.model(**self._dev_in[k], depth_gt_index=depth_gt_index, camera_gt_index=camera_gt_index)
        self._free[k] = torch.cuda.Event()
        self._free[k].record(main)
        computed = torch.cuda.Event()
        computed.record(main)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(computed)
            if self._host_out[k] is None or any(self._host_out[k][n].shape != out[n].shape for n in self.out_keys):
                self._host_out[k] = {n: torch.empty(out[n].shape, dtype=out[n].dtype).pin_memory() for n in self.out_keys}
            for n in self.out_keys:
                out[n].record_stream(self.s_out)
                self._host_out[k][n].copy_(out[n], non_blocking=True)
            self._done[k] = torch.cuda.Event()
            self._done[k].record(self.s_out)
        self._n += 1
        return self._n - 1

    def result(self, ticket: int) -> Dict[str, torch.Tensor]:
        """Pinned host tensors of request `ticket` (valid until `slots` further requests have been submitted)."""
        k = ticket % self.slots
        self._done[k].synchronize()
        return self._host_out[k]

    def drain(self):
        for ev in self._done:
            if ev is not None:
                ev.synchronize()
        torch.cuda.current_stream(self.device).wait_stream(self.s_out)
