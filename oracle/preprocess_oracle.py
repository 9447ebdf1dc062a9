"""CPU restatement of the reference's input pipeline  --  TEST INFRASTRUCTURE (see oracle/omnivggt_oracle.py).

``load_views`` follows reference visual_util.py:679-841 (load_images_and_cameras) per view, on in-memory arrays, with the
same third-party calls the reference makes (Pillow ``Image.resize(..., BICUBIC)``, OpenCV ``cv2.resize(..., INTER_NEAREST)``):
these libraries (pinned by the image: Pillow 12.2, OpenCV 4.13) are the arithmetic oracle of the two resampling steps.
``pil_bicubic_coeffs`` / ``cv2_nearest_index`` restate the two libraries' published index / coefficient arithmetic (Pillow
src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc; OpenCV imgproc/resize.cpp: resizeNN) -- the tables the
CUDA kernels consume; tests/test_preprocess.py checks them against the libraries themselves.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def target_geometry(width: int, height: int, target_size: int = 518) -> Tuple[int, int, int, int]:
    """(new_width, new_height, crop_start_y, final_height): reference visual_util.py:731-747."""
    new_width = target_size
    new_height = round(height * (new_width / width) / 14) * 14
    crop = (new_height - target_size) // 2 if new_height > target_size else 0
    return new_width, new_height, crop, min(new_height, target_size)


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_coeffs(in_size: int, out_size: int):
    """Pillow's fixed-point bicubic taps for one axis: (xmin int32 [out], xcount int32 [out], kk int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    xcnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        n = hi - lo
        w = [_bicubic((x + lo - center + 0.5) * ss) for x in range(n)]
        ww = sum(w)              # left-to-right double accumulation, as the C loop
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmin[xx], xcnt[xx] = lo, n
    return xmin, xcnt, kk


def pil_resize_u8(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """uint8 [h, w, 3] -> uint8 [new_h, new_w, 3] with Pillow's two-pass fixed-point bicubic (horizontal first)."""
    h, w, _ = img.shape
    cur = img.astype(np.int64)
    for axis, (n_in, n_out) in ((1, (w, new_w)), (0, (h, new_h))):
        if n_in == n_out:
            continue
        xmin, xcnt, kk = pil_bicubic_coeffs(n_in, n_out)
        out_shape = list(cur.shape)
        out_shape[axis] = n_out
        out = np.empty(out_shape, np.int64)
        for o in range(n_out):
            sl = [slice(None)] * 3
            sl[axis] = slice(xmin[o], xmin[o] + xcnt[o])
            k = kk[o, :xcnt[o]].astype(np.int64)
            shape = [1, 1, 1]
            shape[axis] = -1
            acc = (cur[tuple(sl)] * k.reshape(shape)).sum(axis=axis) + (1 << (PRECISION_BITS - 1))
            dst = [slice(None)] * 3
            dst[axis] = o
            out[tuple(dst)] = np.clip(acc >> PRECISION_BITS, 0, 255)
        cur = out
    return cur.astype(np.uint8)


def cv2_nearest_index(src: int, dst: int) -> np.ndarray:
    """Source index of every destination index for cv2.resize(..., INTER_NEAREST) (imgproc/resize.cpp: resizeNN)."""
    inv = 1.0 / (dst / src)
    return np.minimum(np.floor(np.arange(dst) * inv).astype(np.int64), src - 1).astype(np.int32)


def se3_inverse_3x4(c2w: np.ndarray) -> np.ndarray:
    """camera-to-world [3,4] / [4,4] -> world-to-camera [3,4] (reference utils/geometry.py:269-318)."""
    R, t = c2w[:3, :3], c2w[:3, 3:4]
    return np.concatenate([R.T, -R.T @ t], 1).astype(np.float32)


def load_views(images: Sequence[np.ndarray], cameras: Sequence[Optional[Tuple[np.ndarray, np.ndarray]]],
               depths: Sequence[Optional[np.ndarray]], target_size: int = 518, max_depth: float = 100.0):
    """In-memory version of reference visual_util.py:719-841.  images: uint8 RGB [h, w, 3]; cameras: (c2w 3x4, K 3x3) or None;
    depths: float32 [h', w'] as loaded (before the validity filter) or None.  Returns the reference's 7-tuple as numpy."""
    import cv2
    from PIL import Image
    imgs, extr, intr, dmaps, masks, didx, cidx = [], [], [], [], [], [], []
    for i, (im, cam, dep) in enumerate(zip(images, cameras, depths)):
        h, w, _ = im.shape
        nw, nh, crop, fh = target_geometry(w, h, target_size)
        sx, sy = nw / w, nh / h
        r = np.asarray(Image.fromarray(im).resize((nw, nh), Image.Resampling.BICUBIC))
        r = r[crop:crop + fh]
        imgs.append(np.ascontiguousarray(r.transpose(2, 0, 1)).astype(np.float32) / np.float32(255))     # ToTensor
        if dep is not None:
            d = dep.astype(np.float32).copy()
            d[~np.isfinite(d)] = 0
            d[d > max_depth] = 0
            d[d < 1e-5] = 0
            d = cv2.resize(d, (nw, nh), interpolation=cv2.INTER_NEAREST)[crop:crop + fh]
            didx.append(i)
            dmaps.append(d)
            masks.append(d > 1e-5)
        else:
            dmaps.append(np.zeros((fh, nw), np.float32))
            masks.append(np.zeros((fh, nw), bool))
        if cam is not None:
            c2w, K = cam
            K = K.astype(np.float32).copy()
            K[0, 0] *= sx; K[1, 1] *= sy; K[0, 2] *= sx; K[1, 2] *= sy
            if nh > target_size:
                K[1, 2] -= crop
            cidx.append(i)
            extr.append(se3_inverse_3x4(c2w.astype(np.float32)))
            intr.append(K)
        else:
            extr.append(np.zeros((3, 4), np.float32))
            intr.append(np.zeros((3, 3), np.float32))
    return (np.stack(imgs), np.stack(extr)[None], np.stack(intr)[None], np.stack(dmaps)[None, ..., None].astype(np.float32),
            np.stack(masks)[None].astype(np.float32), didx, cidx)
