"""Oracle restatement of the upstream croco building blocks (TEST INFRASTRUCTURE ONLY).

The reference imports these from the un-vendored `croco` package
(`from croco.models.blocks import Block, Mlp, CrossAttention, DropPath` — reference
src/panst3r/model/blocks.py:7, model/input_mixer.py:5, model/upscalers/pixel_shuffle.py:7) and
`get_pos_embed('RoPE100')` from must3r (model/input_mixer.py:6,16).  croco @ branch croco_module
(README.md:70, unpinned commit).  Restated from the published architecture:
pre-norm ViT blocks, packed qkv with rows [q|k|v], per-head softmax(q k^T hd^-0.5) v, 2-D RoPE.
Parameter names match upstream so that state dicts interchange.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


# The reference toggles xformers' memory-efficient attention globally (gradio_panst3r.py:25,
# `toggle_memory_efficient_attention(enabled=has_xformers)`); xformers is not installable here, torch's fused SDPA
# stands in for it (SURVEY §8d "GPU reference baseline").  Off by default: the CPU oracle uses the plain softmax form.
_FUSED_ATTENTION = False


def toggle_memory_efficient_attention(enabled: bool = True):
    global _FUSED_ATTENTION
    _FUSED_ATTENTION = bool(enabled)


def _attend(q, k, v, scale, mask=None):
    """softmax(q k^T scale [masked]) v on (B, H, N, hd) tensors; mask: bool, True = blocked."""
    if _FUSED_ATTENTION:
        am = None if mask is None else ~mask
        return F.scaled_dot_product_attention(q, k, v, attn_mask=am, scale=scale)
    attn = (q @ k.transpose(-2, -1)) * scale
    if mask is not None:
        attn = attn.masked_fill(mask, float("-inf"))
    return attn.softmax(dim=-1) @ v


class RoPE2D(nn.Module):
    """2-D rotary embedding, base frequency `freq` (Appendix A.2; curope kernels.cu semantics, fp32 angles).

    tokens (B, H, N, D), positions (B, N, 2) = (y, x).  First D/2 channels rotate with y, second with x; inside a
    half of width Dh=D/2 the pairs are (j, j + Dh/2) with angle p * freq^(-j / (Dh/2)).
    """

    def __init__(self, freq: float = 100.0):
        super().__init__()
        self.base = freq

    def forward(self, tokens: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        B, H, N, D = tokens.shape
        assert D % 4 == 0
        Q = D // 4
        inv = self.base ** (-torch.arange(Q, dtype=torch.float32, device=tokens.device) / Q)
        out = torch.empty_like(tokens)
        tf = tokens.float()
        for half in range(2):
            ang = positions[..., half].to(torch.float32)[:, None, :, None] * inv  # B 1 N Q
            c, s = ang.cos(), ang.sin()
            lo = half * (D // 2)
            u, v = tf[..., lo:lo + Q], tf[..., lo + Q:lo + 2 * Q]
            out[..., lo:lo + Q] = (u * c - v * s).to(tokens.dtype)
            out[..., lo + Q:lo + 2 * Q] = (v * c + u * s).to(tokens.dtype)
        return out


def get_pos_embed(name: str):
    assert name.startswith("RoPE"), name
    return RoPE2D(float(name[len("RoPE"):]))


class DropPath(nn.Module):  # identity at inference
    def __init__(self, p: float = 0.0):
        super().__init__()

    def forward(self, x):
        return x


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, bias=True, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Attention(nn.Module):
    def __init__(self, dim, rope=None, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.rope = rope

    def forward(self, x, xpos):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).transpose(1, 3)  # B H 3 N hd
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        if self.rope is not None:
            q, k = self.rope(q, xpos), self.rope(k, xpos)
        x = _attend(q, k, v, self.scale).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, rope=None):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, rope=rope, num_heads=num_heads, qkv_bias=qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), act_layer=act_layer)

    def forward(self, x, xpos):
        x = x + self.attn(self.norm1(x), xpos)
        x = x + self.mlp(self.norm2(x))
        return x


class CrossAttention(nn.Module):
    def __init__(self, dim, rope=None, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.projq = nn.Linear(dim, dim, bias=qkv_bias)
        self.projk = nn.Linear(dim, dim, bias=qkv_bias)
        self.projv = nn.Linear(dim, dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.rope = rope

    def forward(self, query, key, value, qpos, kpos, mask=None):
        B, Nq, C = query.shape
        Nk = key.shape[1]
        H = self.num_heads
        q = self.projq(query).reshape(B, Nq, H, C // H).permute(0, 2, 1, 3)
        k = self.projk(key).reshape(B, Nk, H, C // H).permute(0, 2, 1, 3)
        v = self.projv(value).reshape(B, Nk, H, C // H).permute(0, 2, 1, 3)
        if self.rope is not None:
            q, k = self.rope(q, qpos), self.rope(k, kpos)
        x = _attend(q, k, v, self.scale, mask).transpose(1, 2).reshape(B, Nq, C)  # mask: True = blocked
        return self.proj(x)


class PatchEmbedDust3R(nn.Module):
    """Conv2d(3, D, P, P) -> tokens (B, N, D), integer (y, x) positions row-major (Appendix A.3)."""

    def __init__(self, img_size=(512, 512), patch_size=16, in_chans=3, embed_dim=1024):
        super().__init__()
        self.patch_size = (patch_size, patch_size)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x, true_shape=None):
        B, C, H, W = x.shape
        # Batches are stored landscape (W >= H); a portrait view (true_shape H > W) is the transposed picture: it is
        # embedded in its true orientation, as dust3r's ManyAR_PatchEmbed does (the convention the reference's own
        # `dinov2_transpose` / `transpose_to_landscape` wrappers assume, model/dino.py:25-33, utils.py:36-49).
        if true_shape is not None and int(true_shape.reshape(-1, 2)[0, 0]) > int(true_shape.reshape(-1, 2)[0, 1]):
            x = x.transpose(2, 3)
        x = self.proj(x)
        h, w = x.shape[-2:]
        ys, xs = torch.meshgrid(torch.arange(h, device=x.device), torch.arange(w, device=x.device), indexing="ij")
        pos = torch.stack([ys.flatten(), xs.flatten()], dim=-1)[None].expand(B, -1, -1).contiguous()
        return x.flatten(2).transpose(1, 2), pos
