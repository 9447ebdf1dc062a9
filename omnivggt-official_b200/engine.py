"""Hot-path engine: weight repacking into kernel layouts, workspace management and the launch sequence of the
aggregator (reference models/omnivggt_aggregator.py:130-305) and the DPT heads (reference
heads/dpt_head.py:128-304) on libovg.  All arithmetic happens in the CUDA library; this file only sequences
kernels on the current stream and owns device buffers."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import ops
from .torch_parts import pack_injection, uv_posembed_separable, uv_posembed_table

BF16, F32 = torch.bfloat16, torch.float32


def _bf(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(BF16).contiguous()


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(F32).contiguous()


def _conv3x3_w(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin], K order (tap = ky*3+kx, cin): matches the row-shifted tap GEMM."""
    return _bf(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))


@dataclass
class BlockPack:
    ln1_w: torch.Tensor; ln1_b: torch.Tensor; w_qkv: torch.Tensor; b_qkv: torch.Tensor
    qn_w: torch.Tensor; qn_b: torch.Tensor; kn_w: torch.Tensor; kn_b: torch.Tensor
    w_proj: torch.Tensor; b_proj: torch.Tensor; g1: torch.Tensor
    ln2_w: torch.Tensor; ln2_b: torch.Tensor; w_fc1: torch.Tensor; b_fc1: torch.Tensor
    w_fc2: torch.Tensor; b_fc2: torch.Tensor; g2: torch.Tensor


def pack_block(bp) -> BlockPack:
    a = bp.attn
    qk = hasattr(a, "q_norm")       # aggregator blocks have q/k LayerNorm; DINOv2 blocks do not
    return BlockPack(_f32(bp.norm1.weight), _f32(bp.norm1.bias), _bf(a.qkv.weight), _f32(a.qkv.bias),
                     _f32(a.q_norm.weight) if qk else None, _f32(a.q_norm.bias) if qk else None,
                     _f32(a.k_norm.weight) if qk else None, _f32(a.k_norm.bias) if qk else None,
                     _bf(a.proj.weight), _f32(a.proj.bias), _f32(bp.ls1.gamma),
                     _f32(bp.norm2.weight), _f32(bp.norm2.bias), _bf(bp.mlp.fc1.weight), _f32(bp.mlp.fc1.bias),
                     _bf(bp.mlp.fc2.weight), _f32(bp.mlp.fc2.bias), _f32(bp.ls2.gamma))


class DPTPack:
    """Kernel-layout weights of one DPT head.  The shared LayerNorm affine (heads/dpt_head.py:66,:227) is folded
    into the 1x1 projections: W (g*xhat + b) + c = (W*g) xhat + (W b + c)."""

    def __init__(self, hp):
        g, b = hp.norm.weight.detach().float(), hp.norm.bias.detach().float()
        self.proj_w, self.proj_b = [], []
        for pr in hp.projects:
            w = pr.weight.detach().float().flatten(1)            # [oc, 2C]
            self.proj_w.append(_bf(w * g[None]))
            self.proj_b.append(_f32(w @ b + pr.bias.detach().float()))
        self.oc = [w.shape[0] for w in self.proj_w]
        r0, r1, r3 = hp.resize_layers["0"], hp.resize_layers["1"], hp.resize_layers["3"]
        # ConvTranspose2d weight [Cin, Cout, k, k] -> rows (ky, kx, cout), cols cin
        self.up_w = [_bf(r.weight.detach().permute(2, 3, 1, 0).reshape(-1, r.weight.shape[0])) for r in (r0, r1)]
        self.up_b = [_f32(r0.bias), _f32(r1.bias)]
        self.down_w, self.down_b = _conv3x3_w(r3.weight.detach()), _f32(r3.bias)
        s = hp.scratch
        self.rn_w = [_conv3x3_w(getattr(s, f"layer{i + 1}_rn").weight.detach()) for i in range(4)]
        self.feat = self.rn_w[0].shape[0]
        self.fus = []
        for name in ("refinenet1", "refinenet2", "refinenet3", "refinenet4"):
            f = getattr(s, name)
            d = {"oc_w": _bf(f.out_conv.weight.detach().flatten(1)), "oc_b": _f32(f.out_conv.bias)}
            for u in ("resConfUnit1", "resConfUnit2"):
                if hasattr(f, u):
                    ru = getattr(f, u)
                    d[u] = (_conv3x3_w(ru.conv1.weight.detach()), _f32(ru.conv1.bias),
                            _conv3x3_w(ru.conv2.weight.detach()), _f32(ru.conv2.bias))
            self.fus.append(d)
        self.oc1_w, self.oc1_b = _conv3x3_w(s.output_conv1.weight.detach()), _f32(s.output_conv1.bias)
        self.oc2_w, self.oc2_b = _conv3x3_w(s.output_conv2["0"].weight.detach()), _f32(s.output_conv2["0"].bias)
        self.w2, self.b2 = _f32(s.output_conv2["2"].weight.detach().flatten(1)), _f32(s.output_conv2["2"].bias)
        self.outc = self.w2.shape[0]


class Workspace:
    """Named device buffers, reused across calls (stable addresses keep the TMA descriptor cache hot)."""

    def __init__(self, device):
        self.device = device
        self.bufs: Dict[str, torch.Tensor] = {}
        self.version = 0          # bumped on every (re)allocation: captured CUDA graphs hold raw pointers

    def get(self, name: str, shape: Sequence[int], dtype=BF16, zero: bool = False) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        key = name
        buf = self.bufs.get(key)
        if buf is None or buf.dtype != dtype or buf.numel() < n:
            buf = torch.empty(max(n, 1), device=self.device, dtype=dtype)
            self.bufs[key] = buf
            self.version += 1
        v = buf[:n].view(*shape)
        if zero:
            v.zero_()
        return v


def _taps(w: int) -> List[int]:
    return [(ky - 1) * (w + 2) + (kx - 1) for ky in range(3) for kx in range(3)]


class Engine:
    def __init__(self, model):
        self.m = model
        self.device = next(model.parameters()).device
        ag = model.aggregator
        self.C = ag.camera_token.shape[-1]
        self.R = ag.register_token.shape[2]
        self.depth = len(ag.frame_blocks)
        self.heads = self.C // 64
        if self.C % 64:
            raise ValueError(f"embed_dim {self.C}: the attention / QKV kernels are built for head_dim 64 (embed_dim % 64 == 0)")
        self.patch = model.patch_size
        self.frame = [pack_block(b) for b in ag.frame_blocks]
        self.glob = [pack_block(b) for b in ag.global_blocks]
        self.cam_tok = _f32(ag.camera_token.reshape(2, self.C))
        self.reg_tok = _f32(ag.register_token.reshape(2, self.R, self.C))
        self.placeholder = _f32(ag.depth_placeholder.reshape(self.C))
        dw = ag.depth_patch_embed.proj.weight.detach()
        self.depth_w = _bf(dw.flatten(1))                         # [C, 2*patch*patch], K order (ch, ky, kx)
        self.depth_b = _f32(ag.depth_patch_embed.proj.bias)
        self.ones_c = torch.ones(self.C, device=self.device, dtype=F32)
        self.inj_pack = pack_injection(ag)
        # frozen DINOv2 patchifier on the same kernels: reference
        # layers/vision_transformer.py:214-271 -- blocks without RoPE / q-k norm, LayerNorm eps 1e-6, LayerScale gammas
        self.dino = None
        pe = ag.patch_embed
        if hasattr(pe, "blocks") and getattr(model, "dino_backend", "ovg") == "ovg":
            if self.C // pe.heads != 64:
                raise ValueError(f"DINOv2 patchifier with head_dim {self.C // pe.heads}: libovg attention needs head_dim 64 "
                                 "(use dino_backend='torch' for other widths)")
            w = pe.patch_embed.proj.weight.detach().flatten(1)                       # [C, 3*p*p]
            kpad = (w.shape[1] + 7) // 8 * 8
            wpad = torch.zeros(w.shape[0], kpad, device=self.device, dtype=BF16)
            wpad[:, :w.shape[1]] = w.to(BF16)
            self.dino = dict(blocks=[pack_block(b) for b in pe.blocks], w=wpad, b=_f32(pe.patch_embed.proj.bias),
                             norm_w=_f32(pe.norm.weight), norm_b=_f32(pe.norm.bias), heads=pe.heads,
                             nreg=pe.register_tokens.shape[1], kpad=kpad)
        self.dpt_packs = {name: DPTPack(getattr(model, name)) for name in ("depth_head", "point_head")
                    if getattr(model, name, None) is not None}
        self.ws = Workspace(self.device)
        self.attn_events = None      # bench.py sets this to a list to time the global-attention launches
        self._idx_cache: Dict[tuple, torch.Tensor] = {}   # small device index tensors (no per-call H2D copies)
        self._rope: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._tables: Dict[tuple, torch.Tensor] = {}

    # ------------------------------------------------------------------------------------------ helpers
    def rope(self, maxpos: int):
        if maxpos not in self._rope:
            self._rope[maxpos] = ops.rope_tables(maxpos, self.device)
        return self._rope[maxpos]

    def cached(self, key: tuple, make) -> torch.Tensor:
        # never evicted: captured CUDA graphs hold raw pointers to these tensors (a few KB per input signature)
        t = self._idx_cache.get(key)
        if t is None:
            t = make().to(self.device)
            self._idx_cache[key] = t
        return t

    def table(self, C: int, h: int, w: int, aspect: float) -> torch.Tensor:
        key = (C, h, w, round(aspect, 9))
        if key not in self._tables:
            self._tables[key] = uv_posembed_table(C, h, w, aspect, self.device)
        return self._tables[key]

    def table_xy(self, C: int, h: int, w: int, aspect: float):
        key = ("xy", C, h, w, round(aspect, 9))
        if key not in self._tables:
            self._tables[key] = uv_posembed_separable(C, h, w, aspect, self.device)
        return self._tables[key]

    def warm_tables(self, H: int, W: int):
        """Build the UV position-embedding tables of this image size on the CURRENT stream (the two DPT heads run on
        different streams and share the cached tables; the first use must not race with the copy that fills them)."""
        hp, wp = H // self.patch, W // self.patch
        for pk in self.dpt_packs.values():
            for oc in pk.oc:
                self.table(oc, hp, wp, W / H)
            self.table_xy(pk.feat // 2, hp * self.patch, wp * self.patch, W / H)

    # ------------------------------------------------------------------------------------------ DINOv2 patchifier
    def dino_patchify(self, images: torch.Tensor, pos_embed: torch.Tensor, mean, std) -> torch.Tensor:
        """images fp32 [K,3,H,W] in [0,1] -> x_norm_patchtokens fp32 [K,P,C] (reference
        layers/vision_transformer.py:214-271).  pos_embed: fp32 [1, 1+P, C], already interpolated to this grid."""
        d, ws, C = self.dino, self.ws, self.C
        K, _, H, W = images.shape
        hp, wp = H // self.patch, W // self.patch
        P, nreg = hp * wp, d["nreg"]
        Td = 1 + nreg + P
        pe = self.m.aggregator.patch_embed
        # token assembly: [cls + pos0, registers, pos_patches] broadcast over frames; the patch-embedding GEMM adds on top
        base = torch.cat([pe.cls_token.float() + pos_embed[:, :1], pe.register_tokens.float(), pos_embed[:, 1:]], 1)
        x = ws.get("dino_x", (K, Td, C), F32)
        x.copy_(base.expand(K, -1, -1))
        cols = ws.get("dino_cols", (K * P, d["kpad"]))
        ops.image_im2col(images.contiguous(), mean, std, cols, K, H, W, self.patch)
        rows = self.cached(("dino_rows", K, Td, P), lambda: (
            (torch.arange(K) * Td)[:, None] + (1 + nreg) + torch.arange(P)[None]).reshape(-1).to(torch.int32))
        x2 = x.view(K * Td, C)
        ops.linear_resid(cols, d["w"], d["b"], self.ones_c, x2, row_index=rows)
        for bp in d["blocks"]:
            self.block(bp, x2, K, Td, Td, 1, None, eps=1e-6)
        out = ws.get("dino_out", (K, P, C), F32)
        ops.layernorm(x2, out.view(K * P, C), d["norm_w"], d["norm_b"], 1e-6, rows=K * P, grp_out=P, grp_in=Td, grp_off=1 + nreg)
        return out

    # ------------------------------------------------------------------------------------------ aggregator
    def block(self, bp: BlockPack, x2: torch.Tensor, batch: int, ntok: int, T: int, wp: int, rope, eps: float = 1e-5):
        """x += g1 * proj(attn(LN1 x)); x += g2 * fc2(gelu(fc1(LN2 x)))   (reference layers/block.py:81-107)."""
        M, C = x2.shape
        ws = self.ws
        xn = ws.get("xn", (M, C))
        q = ws.get("q", (batch, self.heads, ntok, 64))
        k = ws.get("k", (batch, self.heads, ntok, 64))
        v = ws.get("v", (batch, self.heads, ntok, 64))
        o = ws.get("o", (M, C))
        h = ws.get("h", (M, 4 * C))
        ops.layernorm(x2, xn, bp.ln1_w, bp.ln1_b, eps)
        ops.qkv_proj(xn, bp.w_qkv, bp.b_qkv, bp.qn_w, bp.qn_b, bp.kn_w, bp.kn_b, q, k, v, ntok=ntok, T=T,
                     nspecial=self.R + 1, wp=wp, rope_cos=rope[0] if rope else None, rope_sin=rope[1] if rope else None)
        if self.attn_events is not None and ntok > T:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.attention(q, k, v, o, batch, self.heads, ntok)
            e1.record()
            self.attn_events.append((e0, e1, batch, ntok))
        else:
            ops.attention(q, k, v, o, batch, self.heads, ntok)
        ops.linear_resid(o, bp.w_proj, bp.b_proj, bp.g1, x2)
        ops.layernorm(x2, xn, bp.ln2_w, bp.ln2_b, eps)
        ops.linear_bf16(xn, bp.w_fc1, bp.b_fc1, act=L.ACT_GELU, out=h)
        ops.linear_resid(h, bp.w_fc2, bp.b_fc2, bp.g2, x2)

    def aggregate(self, patch_tokens: torch.Tensor, inj: torch.Tensor, depth: Optional[torch.Tensor],
                  mask: Optional[torch.Tensor], depth_idx: List[int], B: int, S: int, H: int, W: int,
                  keep: Sequence[int]):
        """patch_tokens fp32 [K,P,C]; inj fp32 [depth+1,K,C].  Returns ({layer: bf16 slot [K,T,2C]}, cam fp32 [K,2C])."""
        C, R = self.C, self.R
        K = B * S
        hp, wp = H // self.patch, W // self.patch
        P = hp * wp
        T = P + R + 1
        ws = self.ws
        x = ws.get("x", (K, T, C), F32)
        def _has_depth():
            h = torch.zeros(B, S, dtype=torch.int32)
            if len(depth_idx):
                h[:, depth_idx] = 1
            return h.reshape(K)
        has_depth = self.cached(("has_depth", B, S, tuple(depth_idx)), _has_depth)
        ops.assemble_tokens(x, patch_tokens, self.cam_tok, self.reg_tok, inj[0], self.placeholder, has_depth, K, S, T, R, C)
        x2 = x.view(K * T, C)
        if len(depth_idx):
            Sd = len(depth_idx)
            idx = self.cached(("depth_idx", tuple(depth_idx)), lambda: torch.tensor(depth_idx, dtype=torch.int32))
            kk = 2 * self.patch * self.patch
            cols = ws.get("depth_cols", (B * Sd * P, kk))
            scratch = ws.get("depth_scratch", (B * 128 * 2,), torch.float64)
            d32 = depth.reshape(B, S, H, W).to(F32).contiguous()
            m32 = mask.reshape(B, S, H, W).to(F32).contiguous()
            ops.depth_im2col(d32, m32, idx, scratch, cols, B, S, Sd, H, W, self.patch)
            rows = self.cached(("depth_rows", B, S, T, P, tuple(depth_idx)), lambda: (
                ((torch.arange(B)[:, None] * S + torch.tensor(depth_idx)[None]) * T)[:, :, None] + (R + 1) +
                torch.arange(P)[None, None]).reshape(-1).to(torch.int32))
            ops.linear_resid(cols, self.depth_w, self.depth_b, self.ones_c, x2, row_index=rows)
        if max(hp, wp) + 1 > 64:
            raise ValueError(f"{H}x{W} input: the fused RoPE epilogue holds 64 positions per axis (at most 882 px per side)")
        rope = self.rope(max(hp, wp) + 1)
        slots: Dict[int, torch.Tensor] = {}
        cam_out = ws.get("cam_out", (K, 2 * C), F32)
        last = self.depth - 1
        for i in range(self.depth):
            self.block(self.frame[i], x2, K, T, T, wp, rope)
            kept = i in keep
            slot = ws.get(f"slot{i}", (K * T, 2 * C)) if kept else None
            ops.inject_snapshot(x2, inj[i + 1], slot, cam_out if i == last else None, K, T, C, 0)
            self.block(self.glob[i], x2, B, S * T, T, wp, rope)
            if kept or i == last:
                ops.inject_snapshot(x2, None, slot, cam_out if i == last else None, K, T, C, C)
            if kept:
                slots[i] = slot.view(K, T, 2 * C)
        return slots, cam_out

    # ------------------------------------------------------------------------------------------ DPT head
    def conv3x3(self, src, w, bias, dst, F_, h, wd, *, act=L.ACT_NONE, skip1=None, skip2=None):
        cin = src.shape[-1]
        ops.gemm(src.reshape(-1, cin), w, taps=_taps(wd), epi=L.EPI_BF16, bias=bias, act=act, out=dst, ldo=dst.shape[-1],
                 skip1=skip1, skip2=skip2, rowmap=L.ROWS_PAD, gh=h, gw=wd)

    def dpt_chunk(self, pk: DPTPack, slots: List[torch.Tensor], f0: int, Fc: int, H: int, W: int, head_act: int,
                  preds: torch.Tensor, conf: torch.Tensor, ns: str = ""):
        """One frame chunk of one head (reference heads/dpt_head.py:185-304).  slots: 4 x bf16 [K,T,2C].
        ``ns`` namespaces the workspace buffers so that the two heads can run concurrently on different streams."""
        R = self.R
        _ws = self.ws

        class _NS:        # thin view of the workspace with prefixed buffer names
            @staticmethod
            def get(name, shape, dtype=BF16, zero=False):
                return _ws.get(ns + name, shape, dtype, zero)
        ws = _NS
        hp, wp = H // self.patch, W // self.patch
        P, T, C2 = hp * wp, hp * wp + R + 1, 2 * self.C
        aspect = W / H
        f = pk.feat
        sizes = [(4 * hp, 4 * wp), (2 * hp, 2 * wp), (hp, wp), ((hp - 1) // 2 + 1, (wp - 1) // 2 + 1)]
        lr = []
        for lvl in range(4):
            oc = pk.oc[lvl]
            xhat = ws.get("dpt_xhat", (Fc * P, C2))
            ops.layernorm(slots[lvl][f0:f0 + Fc].reshape(Fc * T, C2), xhat, None, None, 1e-5, rows=Fc * P,
                          grp_out=P, grp_in=T, grp_off=R + 1)
            tab = self.table(oc, hp, wp, aspect)
            lh, lw = sizes[lvl]
            feat = ws.get(f"dpt_feat{lvl}", (Fc, lh + 2, lw + 2, oc))
            if lvl == 2:
                feat.zero_()
                ops.gemm(xhat, pk.proj_w[lvl], epi=L.EPI_BF16, bias=pk.proj_b[lvl], table=tab, table_rows=P, out=feat,
                         ldo=oc, rowmap=L.ROWS_DENSE2PAD, gh=hp, gw=wp)
            else:
                dense = ws.get("dpt_dense", (Fc * P, oc))
                ops.gemm(xhat, pk.proj_w[lvl], epi=L.EPI_BF16, bias=pk.proj_b[lvl], table=tab, table_rows=P, out=dense,
                         ldo=oc)
                feat.zero_()
                if lvl < 2:
                    ps = 4 if lvl == 0 else 2
                    ops.gemm(dense, pk.up_w[lvl], epi=L.EPI_BF16, bias=pk.up_b[lvl], out=feat, ldo=oc,
                             rowmap=L.ROWS_PIXSHUF, gh=hp, gw=wp, ps=ps, cout=oc)
                else:
                    cols = ws.get("dpt_cols", (Fc * lh * lw, 9 * oc))
                    ops.im2col3x3s2(dense, cols, Fc, hp, wp, oc)
                    ops.gemm(cols, pk.down_w, epi=L.EPI_BF16, bias=pk.down_b, out=feat, ldo=oc,
                             rowmap=L.ROWS_DENSE2PAD, gh=lh, gw=lw)
            # layerN_rn (no bias); only relu(l_rn) is ever consumed (in-place ReLU quirk, dpt_head.py:315,:389)
            l = ws.get(f"dpt_lr{lvl}", (Fc, lh + 2, lw + 2, f))
            self.conv3x3(feat, pk.rn_w[lvl], None, l, Fc, lh, lw, act=L.ACT_RELU)
            lr.append(l)
        # fusion: refinenet4 -> 3 -> 2 -> 1
        X = None
        for lvl in (3, 2, 1, 0):
            fu = pk.fus[lvl]
            lh, lw = sizes[lvl]
            shp = (Fc, lh + 2, lw + 2, f)
            if X is None:
                U = lr[lvl]
            else:
                c1w, c1b, c2w, c2b = fu["resConfUnit1"]
                t1 = ws.get("dpt_t", shp)
                self.conv3x3(lr[lvl], c1w, c1b, t1, Fc, lh, lw, act=L.ACT_RELU)
                U = ws.get("dpt_u", shp)
                self.conv3x3(t1, c2w, c2b, U, Fc, lh, lw, act=L.ACT_RELU, skip1=lr[lvl], skip2=X)
            c1w, c1b, c2w, c2b = fu["resConfUnit2"]
            t2 = ws.get("dpt_t", shp)
            self.conv3x3(U, c1w, c1b, t2, Fc, lh, lw, act=L.ACT_RELU)
            V = ws.get("dpt_v", shp)
            self.conv3x3(t2, c2w, c2b, V, Fc, lh, lw, skip1=U)
            # out_conv (1x1) commutes with the bilinear resize; apply it at the low resolution
            Wv = ws.get("dpt_w", shp)
            ops.gemm(V.reshape(-1, f), fu["oc_w"], epi=L.EPI_BF16, bias=fu["oc_b"], out=Wv, ldo=f, rowmap=L.ROWS_PAD,
                     gh=lh, gw=lw)
            th, tw = sizes[lvl - 1] if lvl > 0 else (2 * lh, 2 * lw)
            X = ws.get(f"dpt_x{lvl}", (Fc, th + 2, tw + 2, f))
            ops.upsample_bilinear(Wv, X, None, None, Fc, lh, lw, th, tw, f)
        th, tw = 2 * sizes[0][0], 2 * sizes[0][1]
        o1 = ws.get("dpt_o1", (Fc, th + 2, tw + 2, f // 2))
        self.conv3x3(X, pk.oc1_w, pk.oc1_b, o1, Fc, th, tw)
        Hh, Ww = hp * self.patch, wp * self.patch
        up = ws.get("dpt_up", (Fc, Hh + 2, Ww + 2, f // 2))
        tx, ty = self.table_xy(f // 2, Hh, Ww, aspect)
        ops.upsample_bilinear(o1, up, tx, ty, Fc, th, tw, Hh, Ww, f // 2)
        ops.gemm(up.reshape(-1, f // 2), pk.oc2_w, taps=_taps(Ww), epi=L.EPI_HEADTAIL, bias=pk.oc2_b, w2=pk.w2, b2=pk.b2,
                 outc=pk.outc, head_act=head_act, preds=preds[f0:f0 + Fc], conf=conf[f0:f0 + Fc], rowmap=L.ROWS_PAD,
                 gh=Hh, gw=Ww)

    def dpt_alloc(self, name: str, K: int, H: int, W: int):
        pk = self.dpt_packs[name]
        return (torch.empty(K, H, W, pk.outc - 1, device=self.device, dtype=F32),
                torch.empty(K, H, W, device=self.device, dtype=F32))

    def dpt(self, name: str, slots: Dict[int, torch.Tensor], layers: Sequence[int], K: int, H: int, W: int,
            head_act: int, chunk: int = 8, out=None):
        pk = self.dpt_packs[name]
        preds, conf = out if out is not None else self.dpt_alloc(name, K, H, W)
        sl = [slots[i] for i in layers]
        for f0 in range(0, K, chunk):
            self.dpt_chunk(pk, sl, f0, min(chunk, K - f0), H, W, head_act, preds, conf, ns=name + ".")
        return preds, conf
