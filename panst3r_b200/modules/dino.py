"""CUDA DINOv2 encoder behind the reference's `DinoV2Encoder(image, true_shape) -> (b, N, 1024)` seam
(reference src/panst3r/model/dino.py:49-71, driven from engine/dino.py:8-20).

Parameter names are HuggingFace `Dinov2Model`'s (`dinov2.embeddings.*`, `dinov2.encoder.layer.N.*`,
`dinov2.layernorm.*`) so the reference's state dict loads unchanged.  The forward pass does not call
transformers: preprocessing + im2col, patch GEMM with the (bicubically pre-interpolated) position embedding
fused as a broadcast residual, 24 x (LN, QKV GEMM, tcgen05 attention over 1+N tokens, proj GEMM with
LayerScale+residual, LN, MLP GEMMs with LayerScale+residual), final LN that drops the CLS row.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .common import bias_of, cat_f32, cat_w16, f32, fold_stats, ln_linear, prepared, w16
from .must3r import oriented


class _Holder(nn.Module):
    pass


class _DinoLayer(nn.Module):
    def __init__(self, dim, mlp_ratio, eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attention = _Holder()
        self.attention.attention = _Holder()
        for n in ("query", "key", "value"):
            setattr(self.attention.attention, n, nn.Linear(dim, dim))
        self.attention.output = _Holder()
        self.attention.output.dense = nn.Linear(dim, dim)
        self.layer_scale1 = _Holder()
        self.layer_scale1.lambda1 = nn.Parameter(torch.ones(dim))
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = _Holder()
        self.mlp.fc1 = nn.Linear(dim, int(dim * mlp_ratio))
        self.mlp.fc2 = nn.Linear(int(dim * mlp_ratio), dim)
        self.layer_scale2 = _Holder()
        self.layer_scale2.lambda1 = nn.Parameter(torch.ones(dim))

    def qkv_weight(self):
        a = self.attention.attention
        return cat_w16([a.query.weight, a.key.weight, a.value.weight])

    def qkv_bias(self):
        a = self.attention.attention
        return cat_f32([a.query.bias, a.key.bias, a.value.bias])


class DinoV2Encoder(nn.Module):
    def __init__(self, dino_model="facebook/dinov2-large", output_stride=16, landscape_only=True, hidden_size=1024,
                 depth=24, num_heads=16, patch_size=14, image_size=518, mlp_ratio=4.0, eps=1e-6):
        super().__init__()
        self.output_stride, self.patch_size, self.embed_dim, self.num_heads = output_stride, patch_size, hidden_size, num_heads
        self.eps = eps
        self.fold_ln = False  # per-view stage: stand-alone LayerNorm kernels (see common.FOLD_LN_MAX_ROWS)
        self.kpad = ((3 * patch_size * patch_size + 7) // 8) * 8  # 588 -> 592: TMA row pitch must be 16 B aligned
        npos = (image_size // patch_size) ** 2
        d = self.dinov2 = _Holder()
        d.embeddings = _Holder()
        d.embeddings.cls_token = nn.Parameter(torch.zeros(1, 1, hidden_size))
        d.embeddings.mask_token = nn.Parameter(torch.zeros(1, hidden_size))
        d.embeddings.position_embeddings = nn.Parameter(torch.zeros(1, npos + 1, hidden_size))
        d.embeddings.patch_embeddings = _Holder()
        d.embeddings.patch_embeddings.projection = nn.Conv2d(3, hidden_size, kernel_size=patch_size, stride=patch_size)
        d.encoder = _Holder()
        d.encoder.layer = nn.ModuleList([_DinoLayer(hidden_size, mlp_ratio, eps) for _ in range(depth)])
        d.layernorm = nn.LayerNorm(hidden_size, eps=eps)

    # ---- prepared constants ----------------------------------------------------------------------
    def _patch_weight(self):
        wt = self.dinov2.embeddings.patch_embeddings.projection.weight
        return prepared("dino_pw", [wt], lambda: F.pad(wt.detach().reshape(wt.shape[0], -1), (0, self.kpad - 3 * self.patch_size ** 2))
                        .to(torch.bfloat16).contiguous())

    def _pos_embed(self, gh: int, gw: int):
        """HF `interpolate_pos_encoding` (bicubic, align_corners=False) — a constant of (gh, gw); weight
        preprocessing done once per token grid, not per forward.  Returns (cls+pos0 bf16 [D], patch pos bf16 [N, D])."""
        pe = self.dinov2.embeddings.position_embeddings
        cls = self.dinov2.embeddings.cls_token

        def build():
            D = pe.shape[-1]
            npos = pe.shape[1] - 1
            s = int(round(npos ** 0.5))
            patch = pe.detach()[:, 1:].float()
            if not (gh * gw == npos and gh == gw):
                patch = F.interpolate(patch.reshape(1, s, s, D).permute(0, 3, 1, 2), size=(gh, gw), mode="bicubic",
                                      align_corners=False).permute(0, 2, 3, 1).reshape(1, gh * gw, D)
            c = (cls.detach().float() + pe.detach()[:, :1].float()).reshape(1, D)
            return torch.cat([c, patch.reshape(gh * gw, D)], 0).to(torch.bfloat16).contiguous()

        t = prepared(f"dino_pos_{gh}x{gw}", [pe, cls], build)
        return t[0], t[1:]

    @torch.no_grad()
    def forward(self, image: torch.Tensor, true_shape, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """image fp32 (b,3,H,W) in [-1,1] -> bf16 (b, N, D) patch tokens (CLS dropped).  `out`: optional bf16
        destination rows (b*N, D) with any row stride."""
        image, H, W = oriented(image, true_shape)  # portrait views are stored transposed (model/dino.py:25-33)
        b = image.shape[0]
        P, D, Hh = self.patch_size, self.embed_dim, self.num_heads
        gh, gw = H // self.output_stride, W // self.output_stride
        N = gh * gw
        T = N + 1
        dev = image.device
        a = ops.dino_preprocess_patchify(image.float(), gh * P, gw * P, P, self.kpad)
        cls_pos, patch_pos = self._pos_embed(gh, gw)
        x = torch.empty((b, T, D), device=dev, dtype=torch.bfloat16)
        x[:, 0] = cls_pos  # constant row (cls_token + pos[0]); pure data movement
        proj = self.dinov2.embeddings.patch_embeddings.projection
        # patch GEMM writes row (view, n) to x[view, 1 + n] and adds the position embedding of token n
        ops.gemm(a, self._patch_weight(), bias=f32(proj.bias), residual=patch_pos, res_mod_rows=N, out=x[:, 1:],
                 rows_per_batch=N, batch_stride=T * D, out_ld=D)
        xr = x.view(b * T, D)
        # LayerNorms are folded into the GEMMs that consume them (common.ln_linear); the residual GEMMs leave the row
        # statistics behind.  Only the first norm1 runs as a kernel: its input (CLS row + remapped patch rows) has no
        # single producing GEMM.
        s1, s2 = fold_stats(b * T, D, dev, 2, enable=self.fold_ln)
        st = None
        o = torch.empty((b, T, D), device=dev, dtype=torch.bfloat16)
        for lyr in self.dinov2.encoder.layer:
            a_ = lyr.attention.attention
            qkv = ln_linear(xr, st, lyr.norm1, [a_.query.weight, a_.key.weight, a_.value.weight],
                            [a_.query.bias, a_.key.bias, a_.value.bias], self.eps, plain_w=lyr.qkv_weight,
                            plain_b=lyr.qkv_bias).view(b, T, 3, Hh, D // Hh)
            # 1 + N tokens: the N patch queries fill whole 256-query CTAs of the tensor-core kernel, the CLS query (one row
            # per view and head) goes to the single-query kernel instead of occupying a fourth CTA per (view, head)
            ops.attention(qkv[:, 1:, 0], qkv[:, :, 1], qkv[:, :, 2], out=o[:, 1:])
            ops.attention(qkv[:, :1, 0], qkv[:, :, 1], qkv[:, :, 2], out=o[:, :1])
            dense = lyr.attention.output.dense
            ops.gemm(o.view(b * T, D), w16(dense.weight), bias=bias_of(dense), col_scale=f32(lyr.layer_scale1.lambda1),
                     residual=xr, out=xr, stats_out=s1)
            h = ln_linear(xr, s1, lyr.norm2, [lyr.mlp.fc1.weight], [lyr.mlp.fc1.bias], self.eps, act=ops.ACT_GELU)
            ops.gemm(h, w16(lyr.mlp.fc2.weight), bias=bias_of(lyr.mlp.fc2), col_scale=f32(lyr.layer_scale2.lambda1),
                     residual=xr, out=xr, stats_out=s2)
            st = s2
        if out is None:
            out = torch.empty((b * N, D), device=dev, dtype=torch.bfloat16)
        ln = self.dinov2.layernorm
        ops.layernorm(x[:, 1:], f32(ln.weight), f32(ln.bias), self.eps, out=out, x_rows=(b * N, D, D, N, T * D))
        return out.view(b, N, D) if out.is_contiguous() else out.unflatten(0, (b, N))
