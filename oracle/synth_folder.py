"""Deterministic synthetic scene folders in the layout reference visual_util.py:679-841 reads (images/*.png, cameras/*.txt,
depths/*.npy|png)  --  TEST INFRASTRUCTURE.  PNG only (lossless: the decoded pixels depend on nothing but the seed)."""
from __future__ import annotations

import os
from typing import Dict

import numpy as np

# name -> (image width, height, number of views); "wide": no crop (392 x 518), "tall": crop to 518 x 518
FOLDERS = {"wide": (640, 480, 3), "tall": (300, 500, 2)}


def _image(rng, h, w, rgba=False):
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([127 + 90 * np.sin(xx / 17.0 + c) * np.cos(yy / 23.0 - c) for c in range(3)], -1)
    img = np.clip(base + rng.normal(0, 25, (h, w, 3)), 0, 255).astype(np.uint8)
    if rgba:
        alpha = np.clip(255 * (0.5 + 0.5 * np.sin(xx / 31.0)), 0, 255).astype(np.uint8)
        img = np.concatenate([img, alpha[..., None]], -1)
    return img


def make_folder(root: str, name: str, seed: int = 0) -> Dict[str, str]:
    import cv2
    from PIL import Image
    w, h, n = FOLDERS[name]
    rng = np.random.default_rng(seed + len(name))
    d = {k: os.path.join(root, name, k) for k in ("images", "cameras", "depths")}
    for p in d.values():
        os.makedirs(p, exist_ok=True)
    for i in range(n):
        stem = f"frame-{i:04d}"
        Image.fromarray(_image(rng, h, w, rgba=(i == 1))).save(os.path.join(d["images"], stem + ".png"))
        if i != 1:                                        # view 1 has no camera
            ang = 0.3 * i
            R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
            c2w = np.concatenate([R, rng.normal(0, 1, (3, 1))], 1)
            K = np.array([[0.8 * w, 0, w / 2 + 3.5], [0, 0.8 * w, h / 2 - 2.25], [0, 0, 1]])
            with open(os.path.join(d["cameras"], stem + ".txt"), "w") as f:
                f.write("# camera-to-world 3x4, then K 3x3\n")
                for row in c2w:
                    f.write(" ".join(repr(float(v)) for v in row) + "\n")
                for row in K:
                    f.write(" ".join(repr(float(v)) for v in row) + "\n")
        if i == 0:                                        # float depth at another resolution, with invalid values
            dep = (0.5 + 6 * rng.random((h // 2 + 3, w // 2 + 1))).astype(np.float32)
            dep[::11, ::7] = np.inf
            dep[5::13, 3::9] = 1e10
            dep[2::17, 1::5] = 0.0
            np.save(os.path.join(d["depths"], stem + ".npy"), dep)
        elif i == n - 1:                                  # 16-bit PNG depth (the reference transposes it: visual_util.py:771)
            dep16 = rng.integers(0, 90, (w // 3, h // 3)).astype(np.uint16)
            cv2.imwrite(os.path.join(d["depths"], stem + ".png"), dep16)
    return d
