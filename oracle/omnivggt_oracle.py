"""CPU oracle for the OmniVGGT hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional fp32 PyTorch restatement of the reference algorithm (aggregator with
depth/camera token injection, DPT depth/point heads, camera head, DINOv2 patchifier),
driven directly by a reference-schema ``state_dict``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module; the product package (``omnivggt-official_b200``) never does.

Parity pin: the reference ships no tests / golden vectors, so this
restatement is pinned against outputs of the reference itself, imported in the build
container by ``oracle/make_golden.py`` (fixtures under ``tests/golden``) and checked by
``tests/test_oracle_golden.py``.

Every function cites the reference file:line (relative to /root/reference) it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

RESNET_MEAN = (0.485, 0.456, 0.406)  # omnivggt/models/aggregator.py:22
RESNET_STD = (0.229, 0.224, 0.225)   # omnivggt/models/aggregator.py:23


@dataclass
class OracleConfig:
    """Hyper-parameters that are not recoverable from tensor shapes."""
    patch_size: int = 14
    head_dim: int = 64                      # aggregator heads = embed_dim / 64 (omnivggt_aggregator.py:23)
    rope_freq: float = 100.0                # omnivggt_aggregator.py:35
    num_register_tokens: int = 4            # omnivggt_aggregator.py:25
    dpt_layers: Sequence[int] = (4, 11, 17, 23)   # heads/dpt_head.py:52
    camera_head_heads: int = 16             # heads/camera_head.py:27
    camera_iters: int = 4                   # heads/camera_head.py:83
    dino_heads: int = 16                    # layers/vision_transformer.py:369-380 (vit_large)
    frames_chunk: int = 8                   # heads/dpt_head.py:133


# ----------------------------------------------------------------------------- primitives

def layer_norm(x: Tensor, w: Optional[Tensor], b: Optional[Tensor], eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def linear(x: Tensor, sd: SD, name: str) -> Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def rope_tables(max_pos: int, half_dim: int, base: float, dtype=torch.float32):
    """cos/sin tables [max_pos, half_dim/2]; layers/rope.py:86-117 (angle cast before cos/sin)."""
    exponents = torch.arange(0, half_dim, 2).float() / half_dim
    inv_freq = 1.0 / (base ** exponents)
    ang = torch.arange(max_pos, dtype=inv_freq.dtype)[:, None] * inv_freq[None, :]
    ang = ang.to(dtype)
    return ang.cos(), ang.sin()


def rope_1d(z: Tensor, p: Tensor, cos_t: Tensor, sin_t: Tensor) -> Tensor:
    """z [..., N, D] (D = one spatial half), p [Bx, N] int positions; layers/rope.py:119-152.
    Pairs are (z[i], z[i + D/2]) sharing the angle p * base^(-2i/D)."""
    d2 = z.shape[-1] // 2
    c = cos_t[p][:, None]          # [Bx,1,N,D/2]
    s = sin_t[p][:, None]
    a, b = z[..., :d2], z[..., d2:]
    return torch.cat([a * c - b * s, b * c + a * s], dim=-1)


def rope_2d(t: Tensor, pos: Tensor, base: float) -> Tensor:
    """t [Bx, H, N, Dh], pos [Bx, N, 2] (y, x); layers/rope.py:154-188."""
    half = t.shape[-1] // 2
    cos_t, sin_t = rope_tables(int(pos.max()) + 1, half, base, t.dtype)
    v, h = t[..., :half], t[..., half:]
    return torch.cat([rope_1d(v, pos[..., 0], cos_t, sin_t),
                      rope_1d(h, pos[..., 1], cos_t, sin_t)], dim=-1)


def attention(sd: SD, pre: str, x: Tensor, heads: int, pos: Optional[Tensor], qk_norm: bool,
              rope_base: float) -> Tensor:
    """layers/attention.py:50-77."""
    Bx, N, C = x.shape
    dh = C // heads
    qkv = linear(x, sd, pre + ".qkv").reshape(Bx, N, 3, heads, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if qk_norm:  # nn.LayerNorm(head_dim), eps 1e-5 (attention.py:43-44)
        q = layer_norm(q, sd[pre + ".q_norm.weight"], sd[pre + ".q_norm.bias"], 1e-5)
        k = layer_norm(k, sd[pre + ".k_norm.weight"], sd[pre + ".k_norm.bias"], 1e-5)
    if pos is not None:
        q = rope_2d(q, pos, rope_base)
        k = rope_2d(k, pos, rope_base)
    o = F.scaled_dot_product_attention(q, k, v)          # attention.py:61-66 (default scale dh^-0.5, no mask)
    o = o.transpose(1, 2).reshape(Bx, N, C)
    return linear(o, sd, pre + ".proj")


def mlp(sd: SD, pre: str, x: Tensor) -> Tensor:
    """layers/mlp.py:34-40; exact (erf) GELU."""
    return linear(F.gelu(linear(x, sd, pre + ".fc1")), sd, pre + ".fc2")


def block(sd: SD, pre: str, x: Tensor, heads: int, pos: Optional[Tensor], qk_norm: bool,
          rope_base: float, eps: float) -> Tensor:
    """layers/block.py:81-107 (eval path) with LayerScale layers/layer_scale.py:26-27."""
    h = attention(sd, pre + ".attn", layer_norm(x, sd[pre + ".norm1.weight"], sd[pre + ".norm1.bias"], eps),
                  heads, pos, qk_norm, rope_base)
    if pre + ".ls1.gamma" in sd:
        h = h * sd[pre + ".ls1.gamma"]
    x = x + h
    h = mlp(sd, pre + ".mlp", layer_norm(x, sd[pre + ".norm2.weight"], sd[pre + ".norm2.bias"], eps))
    if pre + ".ls2.gamma" in sd:
        h = h * sd[pre + ".ls2.gamma"]
    return x + h


# ----------------------------------------------------------------------------- patchifier

def conv_patch_embed(sd: SD, pre: str, img: Tensor, patch: int) -> Tensor:
    """layers/patch_embed.py:68-81: conv k=s=patch, flatten row-major -> [K, P, C]."""
    assert img.shape[-2] % patch == 0 and img.shape[-1] % patch == 0
    y = F.conv2d(img, sd[pre + ".proj.weight"], sd[pre + ".proj.bias"], stride=patch)
    return y.flatten(2).transpose(1, 2)


def dino_pos_embed(sd: SD, pre: str, npatch: int, h_img: int, w_img: int, patch: int) -> Tensor:
    """layers/vision_transformer.py:180-212 (interpolate_offset=0.0, antialias=True as built at
    models/aggregator.py:152-186). NB the reference passes (w=H_img, h=W_img) swapped names:
    prepare_tokens_with_masks unpacks ``B, nc, w, h = x.shape`` (vision_transformer.py:215)."""
    pe = sd[pre + ".pos_embed"]
    n = pe.shape[1] - 1
    w, h = h_img, w_img
    if npatch == n and w == h:
        return pe
    pe = pe.float()
    dim = pe.shape[-1]
    w0, h0 = w // patch, h // patch
    m = int(math.sqrt(n))
    grid = F.interpolate(pe[:, 1:].reshape(1, m, m, dim).permute(0, 3, 1, 2), size=(w0, h0),
                         mode="bicubic", antialias=True)
    grid = grid.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat([pe[:, :1], grid], dim=1)


def dino_patchify(sd: SD, pre: str, img: Tensor, cfg: OracleConfig) -> Tensor:
    """DINOv2 ViT forward_features -> x_norm_patchtokens; layers/vision_transformer.py:214-271."""
    x = conv_patch_embed(sd, pre + ".patch_embed", img, cfg.patch_size)
    K, P, C = x.shape
    x = torch.cat([sd[pre + ".cls_token"].expand(K, -1, -1), x], dim=1)
    x = x + dino_pos_embed(sd, pre, P, img.shape[-2], img.shape[-1], cfg.patch_size)
    nreg = 0
    if pre + ".register_tokens" in sd:
        reg = sd[pre + ".register_tokens"]
        nreg = reg.shape[1]
        x = torch.cat([x[:, :1], reg.expand(K, -1, -1), x[:, 1:]], dim=1)
    i = 0
    while f"{pre}.blocks.{i}.norm1.weight" in sd:   # eps 1e-6, LayerScale, no rope / qk-norm
        x = block(sd, f"{pre}.blocks.{i}", x, cfg.dino_heads, None, False, 0.0, 1e-6)
        i += 1
    x = layer_norm(x, sd[pre + ".norm.weight"], sd[pre + ".norm.bias"], 1e-6)
    return x[:, 1 + nreg:]


# ----------------------------------------------------------------------------- pose encoding

def se3_inverse(m: Tensor) -> Tensor:
    """utils/geometry.py:269-318 for [N,4,4]."""
    R, t = m[:, :3, :3], m[:, :3, 3:]
    out = torch.eye(4, dtype=m.dtype).repeat(len(m), 1, 1)
    out[:, :3, :3] = R.transpose(1, 2)
    out[:, :3, 3:] = -(R.transpose(1, 2) @ t)
    return out


def rotmat_to_quat_xyzw(R: Tensor) -> Tensor:
    """utils/rotation.py:47-109 + standardize :126-138 (scalar-last, w >= 0)."""
    m = R.reshape(R.shape[:-2] + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.unbind(-1)
    q_abs = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22,
                         1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1).clamp(min=0).sqrt()
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1)], -2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    best = q_abs.argmax(-1)
    rijk = torch.gather(cand, -2, best[..., None, None].expand(best.shape + (1, 4))).squeeze(-2)
    q = rijk[..., [1, 2, 3, 0]]
    return torch.where(q[..., 3:4] < 0, -q, q)


def normalize_extrinsics(E: Tensor) -> Tensor:
    """models/omnivggt_aggregator.py:85-105: re-base on first selected cam; scale by the mean
    camera-'centre' distance (translation column of the re-based world->cam matrices)."""
    B, S = E.shape[:2]
    Eh = torch.zeros(B, S, 4, 4, dtype=E.dtype)
    Eh[:, :, :3] = E
    Eh[:, :, 3, 3] = 1.0
    new = Eh @ se3_inverse(Eh[:, 0])[:, None]
    if S > 1:
        c = new[:, :, :3, 3]
        scale = (c - c[:, :1]).norm(dim=-1)[:, 1:].mean(dim=1, keepdim=True).clamp(min=1e-6)
        new[:, :, :3, 3] = new[:, :, :3, 3] / scale[..., None]
    return new[:, :, :3]


def pose_encoding(E: Tensor, Kmat: Tensor, H: int, W: int) -> Tensor:
    """utils/pose_enc.py:49-58: [t(3), quat xyzw(4), fov_h, fov_w]."""
    fov_h = 2 * torch.atan((H / 2) / Kmat[..., 1, 1])
    fov_w = 2 * torch.atan((W / 2) / Kmat[..., 0, 0])
    return torch.cat([E[..., :3, 3], rotmat_to_quat_xyzw(E[..., :3, :3]),
                      fov_h[..., None], fov_w[..., None]], -1).float()


def normalize_depth(d: Tensor, m: Tensor, eps: float = 1e-8) -> Tensor:
    """models/omnivggt_aggregator.py:107-128. d [B,V,H,W], m [B,V,H,W] -> [B,V,H,W]."""
    out = torch.zeros_like(d)
    for b in range(d.shape[0]):
        valid = d[b][m[b] > 0]
        if valid.numel() == 0:
            continue
        out[b] = d[b] / (valid.mean() + eps) * m[b]
    return out


# ----------------------------------------------------------------------------- aggregator

def special_tokens(tok: Tensor, B: int, S: int) -> Tensor:
    """models/aggregator.py:343-366: slot 0 -> view 0, slot 1 -> views 1..S-1."""
    first = tok[:, 0:1].expand(B, 1, *tok.shape[2:])
    rest = tok[:, 1:2].expand(B, S - 1, *tok.shape[2:])
    return torch.cat([first, rest], 1).reshape(B * S, *tok.shape[2:])


def aggregator(sd: SD, images: Tensor, extrinsics, intrinsics, depth, mask,
               depth_idx: List[int], cam_idx: List[int], cfg: OracleConfig,
               keep: Optional[Sequence[int]] = None, pre: str = "aggregator"):
    """models/omnivggt_aggregator.py:130-305. Returns {layer: [B,S,T,2C]} for layers in ``keep``
    (all layers when None) and patch_start_idx."""
    B, S, Cin, H, W = images.shape
    if Cin != 3:
        raise ValueError(f"Expected 3 input channels, got {Cin}")
    mean = torch.tensor(RESNET_MEAN).view(1, 1, 3, 1, 1)
    std = torch.tensor(RESNET_STD).view(1, 1, 3, 1, 1)
    img = ((images - mean) / std).reshape(B * S, Cin, H, W)
    if pre + ".patch_embed.cls_token" in sd:
        patches = dino_patchify(sd, pre + ".patch_embed", img, cfg)
    else:
        patches = conv_patch_embed(sd, pre + ".patch_embed", img, cfg.patch_size)
    K, P, C = patches.shape
    heads = C // cfg.head_dim
    depth_layers = 0
    while f"{pre}.frame_blocks.{depth_layers}.norm1.weight" in sd:
        depth_layers += 1

    cam_tok = special_tokens(sd[pre + ".camera_token"], B, S)       # [K,1,C]
    reg_tok = special_tokens(sd[pre + ".register_token"], B, S)     # [K,R,C]

    pose = None
    g0 = torch.zeros(K, 1, C)
    cam_rows = None
    if len(cam_idx) != 0:
        idx = torch.tensor(cam_idx)
        pose = pose_encoding(normalize_extrinsics(extrinsics[:, idx]), intrinsics[:, idx], H, W)
        cam_rows = (torch.arange(B)[:, None] * S + idx[None]).reshape(-1)
        g0[cam_rows] = linear(pose, sd, pre + ".pose_embeddings.0").reshape(-1, 1, C)
    if len(depth_idx) != 0:
        idx = torch.tensor(depth_idx)
        dsel, msel = depth[:, idx].squeeze(-1), mask[:, idx]
        dn = normalize_depth(dsel, msel)
        dm = torch.stack([dn.reshape(-1, H, W), msel.reshape(-1, H, W)], 1)
        dtok = conv_patch_embed(sd, pre + ".depth_patch_embed", dm, cfg.patch_size)
        dfull = sd[pre + ".depth_placeholder"].expand(K, P, C).clone()
        rows = (torch.arange(B)[:, None] * S + idx[None]).reshape(-1)
        dfull[rows] = dtok
    else:
        dfull = sd[pre + ".depth_placeholder"].expand(K, P, C)

    cam_tok = cam_tok + linear(g0, sd, pre + ".camera_adapters.0")   # bias hits ALL frames (:211)
    tokens = torch.cat([cam_tok, reg_tok, patches + dfull], 1)       # [K,T,C]
    T = tokens.shape[1]
    nspecial = T - P
    hp, wp = H // cfg.patch_size, W // cfg.patch_size
    yy, xx = torch.meshgrid(torch.arange(hp), torch.arange(wp), indexing="ij")
    pos = torch.stack([yy.reshape(-1), xx.reshape(-1)], -1) + 1                 # rope.py:53-56, :219
    pos = torch.cat([torch.zeros(nspecial, 2, dtype=pos.dtype), pos], 0)[None].expand(K, -1, -1)

    out = {}
    for i in range(depth_layers):
        tokens = block(sd, f"{pre}.frame_blocks.{i}", tokens.reshape(K, T, C), heads, pos, True,
                       cfg.rope_freq, 1e-5)
        g = torch.zeros(K, 1, C)
        if pose is not None:
            g[cam_rows] = linear(pose, sd, f"{pre}.pose_embeddings.{i + 1}").reshape(-1, 1, C)
        inj = linear(g, sd, f"{pre}.camera_adapters.{i + 1}")         # [K,1,C] (:273-287)
        tokens = torch.cat([tokens[:, :1] + inj, tokens[:, 1:]], 1)   # (:301) zeros elsewhere
        frame_i = tokens.reshape(B, S, T, C)
        tokens = block(sd, f"{pre}.global_blocks.{i}", tokens.reshape(B, S * T, C), heads,
                       pos.reshape(B, S * T, 2), True, cfg.rope_freq, 1e-5)
        if keep is None or i in keep:
            out[i] = torch.cat([frame_i, tokens.reshape(B, S, T, C)], -1)   # (:248-251)
    return out, nspecial


# ----------------------------------------------------------------------------- DPT head

def uv_posembed(C: int, h: int, w: int, aspect: float) -> Tensor:
    """heads/utils.py:11-108 + heads/dpt_head.py:262-272 -> [C,h,w] (already x0.1)."""
    diag = (aspect ** 2 + 1.0) ** 0.5
    sx, sy = aspect / diag, 1.0 / diag
    xs = torch.linspace(-sx * (w - 1) / w, sx * (w - 1) / w, w, dtype=torch.float32)
    ys = torch.linspace(-sy * (h - 1) / h, sy * (h - 1) / h, h, dtype=torch.float32)
    uu, vv = torch.meshgrid(xs, ys, indexing="xy")          # [h,w]
    q = C // 4
    omega = 1.0 / (100.0 ** (torch.arange(q, dtype=torch.double) / q))

    def sc(p):
        o = p.reshape(-1).double()[:, None] * omega[None]
        return torch.cat([o.sin(), o.cos()], 1).float()

    emb = torch.cat([sc(uu), sc(vv)], -1).reshape(h, w, C)
    return (emb * 0.1).permute(2, 0, 1)


def conv(sd: SD, name: str, x: Tensor, stride=1, padding=0) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def rcu(sd: SD, pre: str, x: Tensor) -> Tensor:
    """heads/dpt_head.py:379-399. The in-place ReLU (:315,:389) makes the skip add relu(x)."""
    r = F.relu(x)
    return conv(sd, pre + ".conv2", F.relu(conv(sd, pre + ".conv1", r, padding=1)), padding=1) + r


def fusion(sd: SD, pre: str, x: Tensor, skip: Optional[Tensor], size=None) -> Tensor:
    """heads/dpt_head.py:445-469."""
    if skip is not None:
        x = x + rcu(sd, pre + ".resConfUnit1", skip)
    x = rcu(sd, pre + ".resConfUnit2", x)
    if size is None:
        size = (x.shape[-2] * 2, x.shape[-1] * 2)
    x = F.interpolate(x, size=size, mode="bilinear", align_corners=True)
    return conv(sd, pre + ".out_conv", x)


def inverse_log(y: Tensor) -> Tensor:
    """heads/head_act.py:115-125."""
    return torch.sign(y) * torch.expm1(torch.abs(y))


def dpt_head(sd: SD, pre: str, inter: Dict[int, Tensor], H: int, W: int, nspecial: int,
             cfg: OracleConfig, activation: str):
    """heads/dpt_head.py:128-304 + heads/head_act.py:61-112. Frame-chunked like the reference."""
    first = inter[cfg.dpt_layers[0]]
    B, S = first.shape[:2]
    preds, confs = [], []
    for s0 in range(0, S, cfg.frames_chunk if cfg.frames_chunk < S else S):
        s1 = min(S, s0 + (cfg.frames_chunk if cfg.frames_chunk < S else S))
        p, c = _dpt_chunk(sd, pre, {k: v[:, s0:s1] for k, v in inter.items()}, H, W, nspecial, cfg,
                          activation)
        preds.append(p)
        confs.append(c)
    return torch.cat(preds, 1), torch.cat(confs, 1)


def _dpt_chunk(sd, pre, inter, H, W, nspecial, cfg, activation):
    ph, pw = H // cfg.patch_size, W // cfg.patch_size
    feats = []
    for lvl, li in enumerate(cfg.dpt_layers):
        x = inter[li][:, :, nspecial:]
        B, S = x.shape[:2]
        x = x.reshape(B * S, ph * pw, -1)
        x = layer_norm(x, sd[pre + ".norm.weight"], sd[pre + ".norm.bias"], 1e-5)
        x = x.permute(0, 2, 1).reshape(B * S, -1, ph, pw)
        x = conv(sd, f"{pre}.projects.{lvl}", x)
        x = x + uv_posembed(x.shape[1], ph, pw, W / H)
        if lvl == 0:
            x = F.conv_transpose2d(x, sd[pre + ".resize_layers.0.weight"], sd[pre + ".resize_layers.0.bias"], stride=4)
        elif lvl == 1:
            x = F.conv_transpose2d(x, sd[pre + ".resize_layers.1.weight"], sd[pre + ".resize_layers.1.bias"], stride=2)
        elif lvl == 3:
            x = conv(sd, pre + ".resize_layers.3", x, stride=2, padding=1)
        feats.append(x)
    l1, l2, l3, l4 = [conv(sd, f"{pre}.scratch.layer{i + 1}_rn", f, padding=1) for i, f in enumerate(feats)]
    o = fusion(sd, pre + ".scratch.refinenet4", l4, None, size=l3.shape[2:])
    o = fusion(sd, pre + ".scratch.refinenet3", o, l3, size=l2.shape[2:])
    o = fusion(sd, pre + ".scratch.refinenet2", o, l2, size=l1.shape[2:])
    o = fusion(sd, pre + ".scratch.refinenet1", o, l1)
    o = conv(sd, pre + ".scratch.output_conv1", o, padding=1)
    o = F.interpolate(o, size=(ph * cfg.patch_size, pw * cfg.patch_size), mode="bilinear", align_corners=True)
    o = o + uv_posembed(o.shape[1], o.shape[2], o.shape[3], W / H)
    o = conv(sd, pre + ".scratch.output_conv2.2",
             F.relu(conv(sd, pre + ".scratch.output_conv2.0", o, padding=1)))
    f = o.permute(0, 2, 3, 1)
    xyz, conf = f[..., :-1], f[..., -1]
    if activation == "exp":
        pts = torch.exp(xyz)
    elif activation == "inv_log":
        pts = inverse_log(xyz)
    else:
        raise ValueError(activation)
    conf = 1 + conf.exp()                                   # "expp1"
    return pts.reshape(B, S, *pts.shape[1:]), conf.reshape(B, S, *conf.shape[1:])


# ----------------------------------------------------------------------------- camera head

def camera_head(sd: SD, pre: str, last: Tensor, cfg: OracleConfig) -> List[Tensor]:
    """heads/camera_head.py:83-154. ``last`` = aggregated tokens of the final layer [B,S,T,2C]."""
    tok = layer_norm(last[:, :, 0], sd[pre + ".token_norm.weight"], sd[pre + ".token_norm.bias"], 1e-5)
    B, S, C = tok.shape
    ntrunk = 0
    while f"{pre}.trunk.{ntrunk}.norm1.weight" in sd:
        ntrunk += 1
    pred, outs = None, []
    for _ in range(cfg.camera_iters):
        inp = sd[pre + ".empty_pose_tokens"].expand(B, S, -1) if pred is None else pred
        mod = linear(F.silu(linear(inp, sd, pre + ".embed_pose")), sd, pre + ".poseLN_modulation.1")
        shift, scale, gate = mod.chunk(3, dim=-1)
        h = gate * (layer_norm(tok, None, None, 1e-6) * (1 + scale) + shift) + tok
        for i in range(ntrunk):
            h = block(sd, f"{pre}.trunk.{i}", h, cfg.camera_head_heads, None, False, 0.0, 1e-5)
        delta = mlp(sd, pre + ".pose_branch",
                    layer_norm(h, sd[pre + ".trunk_norm.weight"], sd[pre + ".trunk_norm.bias"], 1e-5))
        pred = delta if pred is None else pred + delta
        outs.append(torch.cat([pred[..., :7], F.relu(pred[..., 7:])], -1))   # head_act.py:12-35
    return outs


# ----------------------------------------------------------------------------- model API

@torch.no_grad()
def omnivggt_forward(sd: SD, images: Tensor, extrinsics=None, intrinsics=None, depth=None, mask=None,
                     depth_gt_index=None, camera_gt_index=None, cfg: Optional[OracleConfig] = None
                     ) -> Dict[str, object]:
    """models/omnivggt.py:20-68."""
    cfg = cfg or OracleConfig()
    if images.dim() == 4:
        images = images.unsqueeze(0)
    depth_gt_index = list(depth_gt_index or [])
    camera_gt_index = list(camera_gt_index or [])
    B, S, _, H, W = images.shape
    nlayers = 0
    while f"aggregator.frame_blocks.{nlayers}.norm1.weight" in sd:
        nlayers += 1
    keep = set(cfg.dpt_layers) | {nlayers - 1}
    inter, nspecial = aggregator(sd, images, extrinsics, intrinsics, depth, mask, depth_gt_index,
                                 camera_gt_index, cfg, keep=keep)
    out: Dict[str, object] = {}
    if "camera_head.token_norm.weight" in sd:
        pl = camera_head(sd, "camera_head", inter[nlayers - 1], cfg)
        out["pose_enc"], out["pose_enc_list"] = pl[-1], pl
    if "depth_head.norm.weight" in sd:
        out["depth"], out["depth_conf"] = dpt_head(sd, "depth_head", inter, H, W, nspecial, cfg, "exp")
    if "point_head.norm.weight" in sd:
        out["world_points"], out["world_points_conf"] = dpt_head(sd, "point_head", inter, H, W, nspecial,
                                                                   cfg, "inv_log")
    out["images"] = images
    return out
