"""Recipe that packs the UNMODIFIED reference model package into oracle/_ref/omnivggt_ref.zip  --  TEST / BASELINE
INFRASTRUCTURE (build container only; /root/reference does not exist on the GPU box).

    python oracle/vendor_ref.py

The reference is pure Python (no build system, SURVEY.md quick facts), so "building" it for the GPU box means making its
model package importable there: the .py files of /root/reference/omnivggt/{models,heads,layers,utils} are zipped
byte-for-byte (Python imports packages from zip archives).  oracle/_ref/ is git-ignored -- no reference source enters the
repository history -- but it travels to the GPU box with the snapshot, where ``bench.py --impl reference`` times the
reference's own ``OmniVGGT.forward`` on the host cores and the ``gpu_torch_baseline`` leg times it on the GPU with the
library kernels it dispatches to.  Nothing in the product package imports it.
"""
from __future__ import annotations

import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref", "omnivggt_ref.zip")
SUBPACKAGES = ("models", "heads", "layers", "utils")


def build(force: bool = False) -> str:
    if not os.path.isdir(os.path.join(REF, "omnivggt")):
        raise FileNotFoundError(f"{REF}/omnivggt not found (the recipe runs in the build container only)")
    if os.path.exists(OUT) and not force:
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    n = 0
    with zipfile.ZipFile(OUT, "w", zipfile.ZIP_DEFLATED) as z:
        dirs_seen, have_init = {"omnivggt"}, set()
        for sub in SUBPACKAGES:
            for root, dirs, files in os.walk(os.path.join(REF, "omnivggt", sub)):
                dirs[:] = [d for d in dirs if d != "__pycache__"]
                rel_dir = os.path.relpath(root, REF)
                dirs_seen.add(rel_dir)
                for f in sorted(files):
                    if f.endswith(".py"):
                        full = os.path.join(root, f)
                        z.write(full, os.path.join(rel_dir, f))
                        n += 1
                        if f == "__init__.py":
                            have_init.add(rel_dir)
        # the reference relies on implicit namespace packages (omnivggt/, models/, heads/ have no __init__.py); a zip
        # archive needs regular packages, so EMPTY __init__.py members are added where the reference has none
        for d in sorted(dirs_seen - have_init):
            z.writestr(os.path.join(d, "__init__.py"), "")
    print(f"packed {n} reference modules -> {OUT}")
    return OUT


def import_reference_zip():
    """Import the packed reference (with the two import shims of oracle/ref_shims.py: evo / matplotlib stubs and a
    torch.hub.load stub).  Returns the reference's OmniVGGT class, or raises FileNotFoundError."""
    if not os.path.exists(OUT):
        raise FileNotFoundError(OUT)
    from unittest.mock import MagicMock
    for mod in ("evo", "evo.main_ape", "evo.main_rpe", "evo.core", "evo.core.sync", "evo.core.metrics",
                "evo.core.trajectory", "evo.tools", "evo.tools.file_interface", "evo.tools.plot",
                "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(mod, MagicMock())
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import torch

    class _NoHub:
        def state_dict(self):
            return {}

    torch.hub.load = lambda *a, **k: _NoHub()     # aggregator.py:191-193 loads with strict=False
    from omnivggt.models.omnivggt import OmniVGGT
    return OmniVGGT


if __name__ == "__main__":
    build(force=True)
