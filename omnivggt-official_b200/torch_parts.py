"""PyTorch host code that the north star keeps in PyTorch: the frozen DINOv2 patchifier, the camera head (perf
negligible), the pose encoding of the auxiliary cameras, the 25 per-layer camera
injection vectors and the sin/cos UV position-embedding tables.  None of this is on the measured CUDA hot path."""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------ blocks
def _ln(x, wb, eps):
    return F.layer_norm(x, (x.shape[-1],), wb.weight.to(x.dtype), wb.bias.to(x.dtype), eps)


def _lin(x, wb):
    return F.linear(x, wb.weight.to(x.dtype), wb.bias.to(x.dtype))


def _lin_lp(x, wb, dtype):
    """Linear with weights held in a cached low-precision replica (fp32 in / fp32 out).  The camera-head GEMMs have
    M = B*S rows (8 at cfg2) and are pure weight streaming: bf16 weights halve the bytes and use the tensor-core GEMV path
    instead of the fp32 SIMT sgemm."""
    if dtype == torch.float32:
        return _lin(x, wb)
    lp = wb.__dict__.get("_lp")
    if lp is None or lp[0].dtype != dtype or lp[0].device != wb.weight.device:
        lp = (wb.weight.detach().to(dtype), wb.bias.detach().to(dtype))
        wb.__dict__["_lp"] = lp
    return F.linear(x.to(dtype), lp[0], lp[1]).float()


def clear_lp_cache(module):
    for m in module.modules():
        m.__dict__.pop("_lp", None)


def torch_block(bp, x: Tensor, heads: int, eps: float, lin=_lin) -> Tensor:
    """Pre-LN block without RoPE / QK-norm (reference layers/block.py:81-107 as used by DINOv2 and the camera
    trunk), on library kernels (cuBLAS + SDPA)."""
    Bx, N, C = x.shape
    h = _ln(x, bp.norm1, eps)
    qkv = lin(h, bp.attn.qkv).reshape(Bx, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(Bx, N, C)
    x = x + lin(o, bp.attn.proj) * bp.ls1.gamma.to(x.dtype)
    h = lin(F.gelu(lin(_ln(x, bp.norm2, eps), bp.mlp.fc1)), bp.mlp.fc2)
    return x + h * bp.ls2.gamma.to(x.dtype)


# ------------------------------------------------------------------------------------------------ DINOv2
def dino_pos_embed(dp, npatch: int, h_img: int, w_img: int, patch: int) -> Tensor:
    """reference layers/vision_transformer.py:180-212 with interpolate_offset=0.0, antialias=True
    (models/aggregator.py:152-186)."""
    pe = dp.pos_embed
    n = pe.shape[1] - 1
    if npatch == n and h_img == w_img:
        return pe
    pe = pe.float()
    dim = pe.shape[-1]
    m = int(math.sqrt(n))
    grid = F.interpolate(pe[:, 1:].reshape(1, m, m, dim).permute(0, 3, 1, 2), size=(h_img // patch, w_img // patch),
                         mode="bicubic", antialias=True)
    return torch.cat([pe[:, :1], grid.permute(0, 2, 3, 1).reshape(1, -1, dim)], dim=1)


def dino_patchify(dp, img: Tensor, patch: int, dtype: torch.dtype) -> Tensor:
    """Frozen DINOv2 ViT -> x_norm_patchtokens [K, P, C] (reference layers/vision_transformer.py:214-271)."""
    x = F.conv2d(img.to(dtype), dp.patch_embed.proj.weight.to(dtype), dp.patch_embed.proj.bias.to(dtype), stride=patch)
    x = x.flatten(2).transpose(1, 2)
    K, P, C = x.shape
    x = torch.cat([dp.cls_token.to(dtype).expand(K, -1, -1), x], 1)
    x = x + dino_pos_embed(dp, P, img.shape[-2], img.shape[-1], patch).to(dtype)
    reg = dp.register_tokens.to(dtype)
    x = torch.cat([x[:, :1], reg.expand(K, -1, -1), x[:, 1:]], 1)
    for blk in dp.blocks:
        x = torch_block(blk, x, dp.heads, 1e-6)
    x = _ln(x, dp.norm, 1e-6)
    return x[:, 1 + reg.shape[1]:]


# ------------------------------------------------------------------------------------------------ camera head
def camera_head(cp, cam_tokens: Tensor, iters: int = 4, dtype: torch.dtype = torch.float32) -> List[Tensor]:
    """reference heads/camera_head.py:83-154.  cam_tokens: fp32 [B, S, 2C] (token 0 of the last layer).
    ``dtype`` = precision of the large weight matrices (activations, norms, residuals and the pose update stay fp32)."""
    lin = (lambda x, wb: _lin_lp(x, wb, dtype))
    tok = _ln(cam_tokens, cp.token_norm, 1e-5)
    B, S, C = tok.shape
    pred, outs = None, []
    for _ in range(iters):
        inp = cp.empty_pose_tokens.expand(B, S, -1) if pred is None else pred
        mod = lin(F.silu(_lin(inp, cp.embed_pose)), cp.poseLN_modulation["1"])
        shift, scale, gate = mod.chunk(3, dim=-1)
        h = gate * (F.layer_norm(tok, (C,), None, None, 1e-6) * (1 + scale) + shift) + tok
        for blk in cp.trunk:
            h = torch_block(blk, h, cp.heads, 1e-5, lin)
        delta = _lin(F.gelu(lin(_ln(h, cp.trunk_norm, 1e-5), cp.pose_branch.fc1)), cp.pose_branch.fc2)
        pred = delta if pred is None else pred + delta
        outs.append(torch.cat([pred[..., :7], F.relu(pred[..., 7:])], -1))     # heads/head_act.py:12-35
    return outs


# ------------------------------------------------------------------------------------------------ pose encoding
def rotmat_to_quat_xyzw(R: Tensor) -> Tensor:
    """reference utils/rotation.py:47-109,:126-138 (scalar-last, real part >= 0)."""
    m = R.reshape(R.shape[:-2] + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.unbind(-1)
    q_abs = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22],
                        -1).clamp(min=0).sqrt()
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1)], -2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    best = q_abs.argmax(-1)
    rijk = torch.gather(cand, -2, best[..., None, None].expand(best.shape + (1, 4))).squeeze(-2)
    q = torch.cat([rijk[..., 1:], rijk[..., :1]], -1)     # rijk -> ijkr (no index tensors: CUDA-graph capturable)
    return torch.where(q[..., 3:4] < 0, -q, q)


def aux_pose_encoding(extr: Tensor, intr: Tensor, H: int, W: int) -> Tensor:
    """Selected world->cam [B,Sc,3,4] + intrinsics -> [B,Sc,9] = [t, quat xyzw, fov_h, fov_w]
    (reference omnivggt_aggregator.py:85-105 + utils/pose_enc.py:49-58 + utils/geometry.py:269-318)."""
    B, Sc = extr.shape[:2]
    E = torch.zeros(B, Sc, 4, 4, dtype=torch.float32, device=extr.device)
    E[:, :, :3] = extr.float()
    E[:, :, 3, 3] = 1.0
    R0, t0 = E[:, 0, :3, :3], E[:, 0, :3, 3:]
    inv0 = torch.eye(4, device=extr.device).repeat(B, 1, 1)
    inv0[:, :3, :3] = R0.transpose(1, 2)
    inv0[:, :3, 3:] = -(R0.transpose(1, 2) @ t0)
    new = E @ inv0[:, None]
    if Sc > 1:
        c = new[:, :, :3, 3]
        scale = (c - c[:, :1]).norm(dim=-1)[:, 1:].mean(dim=1, keepdim=True).clamp(min=1e-6)
        new[:, :, :3, 3] = new[:, :, :3, 3] / scale[..., None]
    fov_h = 2 * torch.atan((H / 2) / intr[..., 1, 1].float())
    fov_w = 2 * torch.atan((W / 2) / intr[..., 0, 0].float())
    return torch.cat([new[:, :, :3, 3], rotmat_to_quat_xyzw(new[:, :, :3, :3]), fov_h[..., None], fov_w[..., None]], -1)


def pack_injection(ap):
    """Stack the depth+1 pose-embedding / camera-adapter layers once (weights are frozen at inference)."""
    return (torch.stack([m.weight for m in ap.pose_embeddings]).float().contiguous(),      # [L+1, C, 9]
            torch.stack([m.bias for m in ap.pose_embeddings]).float().contiguous(),        # [L+1, C]
            torch.stack([m.weight for m in ap.camera_adapters]).float().contiguous(),      # [L+1, C, C]
            torch.stack([m.bias for m in ap.camera_adapters]).float().contiguous())        # [L+1, C]


def injection_vectors(pack, pose: Optional[Tensor], cam_idx: List[int], B: int, S: int,
                      rows: Optional[Tensor] = None) -> Tensor:
    """All depth+1 camera injection vectors [L+1, K, C] fp32 in one shot: they depend only on the inputs, not on
    the token stream (reference omnivggt_aggregator.py:172-179,:211,:273-287).  Frames without a camera receive the
    adapter *bias* (the adapter is applied to a zero row).  ``pack`` = pack_injection(aggregator params)."""
    Wp, bp, Wa, ba = pack
    L1, C = ba.shape
    out = ba[:, None, :].expand(L1, B * S, C).clone()
    if pose is not None and len(cam_idx):
        g = torch.einsum("brn,lcn->lbrc", pose.float(), Wp) + bp[:, None, None, :]      # [L+1,B,Sc,C]
        inj = torch.einsum("lbrc,ldc->lbrd", g, Wa) + ba[:, None, None, :]
        if rows is None:     # frame rows b*S + idx (pass a cached device tensor to avoid a host->device copy per call)
            rows = (torch.arange(B)[:, None] * S + torch.tensor(cam_idx)[None]).reshape(-1).to(out.device)
        out[:, rows] = inj.reshape(L1, -1, C)
    return out.contiguous()


# ------------------------------------------------------------------------------------------------ UV pos-embed
def uv_posembed_table(C: int, h: int, w: int, aspect: float, device) -> Tensor:
    """[h*w, C] fp32, already scaled by 0.1 (reference heads/utils.py:11-108, heads/dpt_head.py:262-272)."""
    diag = (aspect ** 2 + 1.0) ** 0.5
    sx, sy = aspect / diag, 1.0 / diag
    xs = torch.linspace(-sx * (w - 1) / w, sx * (w - 1) / w, w, dtype=torch.float32)
    ys = torch.linspace(-sy * (h - 1) / h, sy * (h - 1) / h, h, dtype=torch.float32)
    uu, vv = torch.meshgrid(xs, ys, indexing="xy")
    q = C // 4
    omega = 1.0 / (100.0 ** (torch.arange(q, dtype=torch.double) / q))

    def sc(p):
        o = p.reshape(-1).double()[:, None] * omega[None]
        return torch.cat([o.sin(), o.cos()], 1).float()

    return (torch.cat([sc(uu), sc(vv)], -1) * 0.1).contiguous().to(device)


def uv_posembed_separable(C: int, h: int, w: int, aspect: float, device) -> Tuple[Tensor, Tensor]:
    """The same embedding in separable form: channels [0, C/2) depend only on x, [C/2, C) only on y (heads/utils.py:26-30
    concatenates emb_x and emb_y).  Returns (tx [w, C/2], ty [h, C/2]) fp32, x0.1."""
    diag = (aspect ** 2 + 1.0) ** 0.5
    sx, sy = aspect / diag, 1.0 / diag
    xs = torch.linspace(-sx * (w - 1) / w, sx * (w - 1) / w, w, dtype=torch.float32)
    ys = torch.linspace(-sy * (h - 1) / h, sy * (h - 1) / h, h, dtype=torch.float32)
    q = C // 4
    omega = 1.0 / (100.0 ** (torch.arange(q, dtype=torch.double) / q))

    def sc(p):
        o = p.double()[:, None] * omega[None]
        return (torch.cat([o.sin(), o.cos()], 1).float() * 0.1).contiguous().to(device)

    return sc(xs), sc(ys)
