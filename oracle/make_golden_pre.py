"""tests/golden/preprocess.json: SHA-256 (and a few samples) of the outputs of the UNMODIFIED reference loader
(visual_util.py:679-841 load_images_and_cameras) on the seeded synthetic folders of oracle/synth_folder.py
(build container only; TEST INFRASTRUCTURE).        python oracle/make_golden_pre.py"""
from __future__ import annotations

import hashlib
import importlib.machinery
import json
import os
import sys
import tempfile
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.synth_folder import FOLDERS, make_folder  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def summarize(out) -> dict:
    images, extr, intr, dep, mask, didx, cidx = out
    img_u8 = np.rint(np.asarray(images, np.float64) * 255).astype(np.uint8)       # ToTensor output is uint8 / 255 exactly
    return {"images_shape": list(img_u8.shape), "images_u8_sha256": digest(img_u8), "images_f32_sha256": digest(np.asarray(images, np.float32)),
            "depth_shape": list(np.asarray(dep).shape), "depth_sha256": digest(np.asarray(dep, np.float32)),
            "mask_sha256": digest(np.asarray(mask, np.float32)),
            "extrinsics": np.asarray(extr, np.float64).round(7).tolist(), "intrinsics": np.asarray(intr, np.float64).round(5).tolist(),
            "depth_indices": list(didx), "camera_indices": list(cidx),
            "images_samples": img_u8[:, :, ::97, ::89].tolist()}


def main():
    for mod in ("evo", "evo.main_ape", "evo.main_rpe", "evo.core", "evo.core.sync", "evo.core.metrics", "evo.core.trajectory",
                "evo.tools", "evo.tools.file_interface", "evo.tools.plot", "matplotlib", "matplotlib.pyplot", "onnxruntime", "trimesh",
                "viser", "viser.transforms", "pillow_heif", "imageio", "imageio.v2", "imageio.v3"):
        m = MagicMock()
        m.__spec__ = importlib.machinery.ModuleSpec(mod, None)
        m.__name__, m.__path__ = mod, []
        sys.modules.setdefault(mod, m)
    sys.path.insert(0, "/root/reference")
    import visual_util as vu
    res = {}
    with tempfile.TemporaryDirectory() as root:
        for name in FOLDERS:
            d = make_folder(root, name, seed=0)
            out = vu.load_images_and_cameras(d["images"], d["cameras"], d["depths"])
            res[name] = summarize([t.numpy() if hasattr(t, "numpy") else t for t in out])
            print(name, res[name]["images_shape"], res[name]["depth_indices"], res[name]["camera_indices"])
    import PIL
    import cv2
    res["_versions"] = {"pillow": PIL.__version__, "opencv": cv2.__version__}
    with open(os.path.join(GOLDEN, "preprocess.json"), "w") as f:
        json.dump(res, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
