"""GPU input pipeline: drop-in for reference ``visual_util.load_images_and_cameras`` (visual_util.py:679-841).

Decoding (PNG / JPEG / .npy / camera .txt parsing, RGBA -> RGB on white) stays on the host, as file I/O; everything the reference
then does per view with Pillow / OpenCV / numpy -- bicubic resize to width 518, height to a multiple of 14, centre crop, ToTensor,
depth validity filter + nearest resize + crop + mask, intrinsics rescale, camera-to-world -> world-to-camera -- runs in libovg
kernels on the device and returns the model's input tuple as CUDA tensors.  The small per-size tap / index tables are computed on
the host with the arithmetic the two libraries publish (Pillow Resample.c, OpenCV resizeNN) and cached on the device, so the image
tensor equals the reference's bit for bit (tests/test_preprocess.py compares against Pillow / OpenCV and the reference loader)."""
from __future__ import annotations

import glob
import math
import os
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

PRECISION_BITS = 32 - 8 - 2
_tables: Dict[tuple, tuple] = {}


def target_geometry(width: int, height: int, target_size: int = 518) -> Tuple[int, int, int, int]:
    """(new_width, new_height, crop_start_y, final_height)   -- visual_util.py:731-747."""
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


def bicubic_taps(in_size: int, out_size: int, device) -> tuple:
    """Pillow's fixed-point bicubic taps of one axis as device tensors (kmin, kcnt, kk [out, ksize], ksize)."""
    key = ("bicubic", in_size, out_size, str(device))
    if key not in _tables:
        scale = in_size / out_size
        filterscale = max(scale, 1.0)
        support = 2.0 * filterscale
        ksize = int(math.ceil(support)) * 2 + 1
        kmin, kcnt = np.zeros(out_size, np.int32), np.zeros(out_size, np.int32)
        kk = np.zeros((out_size, ksize), np.int32)
        ss = 1.0 / filterscale
        for xx in range(out_size):
            center = (xx + 0.5) * scale
            lo = max(int(center - support + 0.5), 0)
            hi = min(int(center + support + 0.5), in_size)
            w = [_bicubic((x + lo - center + 0.5) * ss) for x in range(hi - lo)]
            ww = sum(w)
            if ww != 0.0:
                w = [v / ww for v in w]
            for x, v in enumerate(w):
                kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
            kmin[xx], kcnt[xx] = lo, hi - lo
        _tables[key] = (torch.from_numpy(kmin).to(device), torch.from_numpy(kcnt).to(device), torch.from_numpy(kk).to(device), ksize)
    return _tables[key]


def nearest_index(src: int, dst: int, device) -> torch.Tensor:
    """Source index per destination index of cv2.resize(..., INTER_NEAREST)."""
    key = ("nearest", src, dst, str(device))
    if key not in _tables:
        inv = 1.0 / (dst / src)
        idx = np.minimum(np.floor(np.arange(dst) * inv).astype(np.int64), src - 1).astype(np.int32)
        _tables[key] = (torch.from_numpy(idx).to(device),)
    return _tables[key][0]


@torch.no_grad()
def preprocess_views(images: Sequence, cameras: Optional[Sequence] = None, depths: Optional[Sequence] = None,
                     target_size: int = 518, max_depth: float = 100.0, device="cuda", depth_transposed: Optional[Sequence[bool]] = None):
    """images: uint8 RGB arrays / tensors [h, w, 3]; cameras: per view (camera-to-world 3x4 or 4x4, K 3x3) or None; depths: per view
    float32 [h', w'] as loaded or None.  Returns (images [S,3,H,W], extrinsics [1,S,3,4], intrinsics [1,S,3,3], depth [1,S,H,W,1],
    mask [1,S,H,W], depth_indices, camera_indices) -- CUDA tensors, the tuple reference visual_util.py:835-841 returns."""
    lib = L.lib()
    dev = torch.device(device)
    S = len(images)
    cameras = list(cameras) if cameras is not None else [None] * S
    depths = list(depths) if depths is not None else [None] * S
    depth_transposed = list(depth_transposed) if depth_transposed is not None else [False] * S
    geoms = []
    for im in images:
        h, w = int(im.shape[0]), int(im.shape[1])
        geoms.append((h, w) + target_geometry(w, h, target_size))
    fh, nw = geoms[0][5], geoms[0][2]
    if any(g[5] != fh for g in geoms):
        raise ValueError("all views of a scene must resize to the same height (the reference stacks them: visual_util.py:835)")
    out = torch.empty(S, 3, fh, nw, device=dev, dtype=torch.float32)
    dmap = torch.zeros(S, fh, nw, device=dev, dtype=torch.float32)
    mask = torch.zeros(S, fh, nw, device=dev, dtype=torch.float32)
    c2w = torch.zeros(S, 3, 4, dtype=torch.float32)
    kin = torch.zeros(S, 3, 3, dtype=torch.float32)
    geom = torch.zeros(S, 3, dtype=torch.float32)
    has = torch.zeros(S, dtype=torch.int32)
    didx, cidx = [], []
    st = L.stream()
    keep = []
    for i, (im, (h, w, _, nh, crop, _)) in enumerate(zip(images, geoms)):
        src = torch.as_tensor(np.array(im, dtype=np.uint8, copy=True) if isinstance(im, np.ndarray) else im, dtype=torch.uint8).to(dev).contiguous()
        assert src.shape == (h, w, 3), "images are uint8 RGB [h, w, 3]"
        hk = bicubic_taps(w, nw, dev) if w != nw else (None, None, None, 0)
        vk = bicubic_taps(h, nh, dev) if h != nh else (None, None, None, 0)
        tmp = torch.empty(h, nw, 3, device=dev, dtype=torch.uint8) if w != nw else None
        L.check(lib.ovg_preprocess_image(src.data_ptr(), h, w, nw, nh, crop, fh, L.ptr(hk[0]), L.ptr(hk[1]), L.ptr(hk[2]), hk[3],
                                         L.ptr(vk[0]), L.ptr(vk[1]), L.ptr(vk[2]), vk[3], L.ptr(tmp), out[i].data_ptr(), st))
        keep += [src, tmp]
        dep = depths[i]
        if dep is not None:
            d = torch.as_tensor(np.ascontiguousarray(dep, dtype=np.float32) if isinstance(dep, np.ndarray) else dep).float().to(dev).contiguous()
            rows, cols = d.shape
            if depth_transposed[i]:         # the reference transposes PNG depth maps after reading them (visual_util.py:771)
                sh, sw, rs, cs = cols, rows, 1, cols
            else:
                sh, sw, rs, cs = rows, cols, cols, 1
            L.check(lib.ovg_preprocess_depth(d.data_ptr(), rs, cs, nearest_index(sh, nh, dev).data_ptr(),
                                             nearest_index(sw, nw, dev).data_ptr(), crop, fh, nw, float(max_depth),
                                             dmap[i].data_ptr(), mask[i].data_ptr(), st))
            keep.append(d)
            didx.append(i)
        cam = cameras[i]
        if cam is not None:
            e, k = np.asarray(cam[0], np.float32), np.asarray(cam[1], np.float32)
            c2w[i] = torch.from_numpy(e[:3, :4].copy())
            kin[i] = torch.from_numpy(k.copy())
            geom[i] = torch.tensor([np.float32(nw / w), np.float32(nh / h), float(crop) if nh > target_size else -1.0])
            has[i] = 1
            cidx.append(i)
    w2c = torch.empty(S, 3, 4, device=dev, dtype=torch.float32)
    kout = torch.empty(S, 3, 3, device=dev, dtype=torch.float32)
    c2w_d, kin_d, geom_d, has_d = c2w.to(dev), kin.to(dev), geom.to(dev), has.to(dev)
    L.check(lib.ovg_prepare_cameras(c2w_d.data_ptr(), kin_d.data_ptr(), geom_d.data_ptr(), has_d.data_ptr(), w2c.data_ptr(),
                                    kout.data_ptr(), S, st))
    torch.cuda.current_stream().synchronize()     # the staging tensors in `keep` are released after the kernels ran
    return out, w2c[None], kout[None], dmap[None, ..., None], mask[None], didx, cidx


def read_camera_txt(path: str):
    """3 lines 3x4 camera-to-world + 3 lines 3x3 K, '#' comments allowed (visual_util.py:843-891)."""
    try:
        lines = [ln.strip() for ln in open(path) if ln.strip() and not ln.strip().startswith("#")]
        if len(lines) < 6:
            return None
        e = [[float(x) for x in lines[i].split()] for i in range(3)]
        k = [[float(x) for x in lines[i].split()] for i in range(3, 6)]
        if any(len(r) != 4 for r in e) or any(len(r) != 3 for r in k):
            return None
        return np.array(e, np.float32), np.array(k, np.float32)
    except Exception:
        return None


def load_images_and_cameras(image_folder: str, camera_folder: Optional[str] = None, depth_folder: Optional[str] = None,
                            target_size: int = 518, max_depth: float = 100, device="cuda"):
    """Same signature, file layout and return tuple as reference visual_util.load_images_and_cameras (visual_util.py:679-841);
    files are decoded on the host (Pillow / numpy), the rest runs on the GPU."""
    from PIL import Image
    paths = sorted(glob.glob(os.path.join(image_folder, "*")))
    paths = [p for p in paths if p.lower().endswith((".png", ".jpg", ".jpeg"))]
    images, cams, deps, transposed = [], [], [], []
    for p in paths:
        stem = Path(p).stem
        img = Image.open(p)
        if img.mode == "RGBA":                                   # white background (visual_util.py:722-726)
            img = Image.alpha_composite(Image.new("RGBA", img.size, (255, 255, 255, 255)), img)
        images.append(np.asarray(img.convert("RGB")))
        dep, tr = None, False
        if depth_folder is not None:
            for cand in (os.path.join(depth_folder, stem + ".npy"), os.path.join(depth_folder, stem + ".png")):
                if os.path.exists(cand):                         # both present: the later candidate wins, as in the reference loop
                    if cand.endswith(".npy"):
                        dep, tr = np.load(cand).astype(np.float32), False
                    else:
                        dep, tr = np.asarray(Image.open(cand)).astype(np.float32), True
        deps.append(dep)
        transposed.append(tr)
        cam = None
        if camera_folder is not None and os.path.exists(os.path.join(camera_folder, stem + ".txt")):
            cam = read_camera_txt(os.path.join(camera_folder, stem + ".txt"))
        cams.append(cam)
    return preprocess_views(images, cams, deps, target_size, max_depth, device, transposed)
