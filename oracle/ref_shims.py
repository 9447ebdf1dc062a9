"""Import the UNMODIFIED reference from /root/reference  --  build-container only, TEST INFRASTRUCTURE.

The reference cannot be imported as-is offline: ``omnivggt.utils.misc``
pulls in ``evo``/``matplotlib`` at import time and ``Aggregator.__build_patch_embed__`` calls
``torch.hub.load`` (network).  This module installs the two shims *without touching the
reference tree* and is used only by ``oracle/make_golden.py``.  /root/reference does not exist
on the GPU box; nothing under tests/, bench.py or smoke() imports this file at run time there.
"""
from __future__ import annotations

import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = "/root/reference"


def import_reference():
    for mod in ("evo", "evo.main_ape", "evo.main_rpe", "evo.core", "evo.core.sync", "evo.core.metrics",
                "evo.core.trajectory", "evo.tools", "evo.tools.file_interface", "evo.tools.plot",
                "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(mod, MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torch

    class _NoHub:
        def state_dict(self):
            return {}

    torch.hub.load = lambda *a, **k: _NoHub()     # aggregator.py:191-193 loads with strict=False
    import omnivggt.models.omnivggt_aggregator as agg
    import omnivggt.heads.dpt_head as dpt
    import omnivggt.heads.camera_head as cam
    import omnivggt.layers.vision_transformer as vit
    import omnivggt.layers as layers
    return agg, dpt, cam, vit, layers
