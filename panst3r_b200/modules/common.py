"""Shared host-side helpers for the CUDA modules: parameter preparation caches and the ViT block runner.

Modules in this package are *parameter containers with the reference's state-dict names* whose forward
passes call only panst3r_b200.ops (the C ABI).  Weights are converted once to the kernel formats
(bf16 [N, K] matrices, fp32 bias / norm vectors, fused or permuted variants) and cached per parameter version.
"""
from __future__ import annotations

from typing import Callable, Dict, Sequence, Tuple

import torch
import torch.nn as nn

from .. import ops

_cache: Dict[Tuple, Tuple[Tuple, torch.Tensor]] = {}


def _ver(ps: Sequence[torch.Tensor]) -> Tuple:
    return tuple((p.data_ptr(), p._version, str(p.device), p.dtype) for p in ps)


def prepared(key: str, params: Sequence[torch.Tensor], build: Callable[[], torch.Tensor]) -> torch.Tensor:
    """Cache `build()` until any of `params` changes (load_state_dict / .to() bump version or data_ptr)."""
    k = (key, tuple(id(p) for p in params))
    v = _ver(params)
    hit = _cache.get(k)
    if hit is not None and hit[0] == v:
        return hit[1]
    with torch.no_grad():
        t = build()
    _cache[k] = (v, t)
    return t


def _check_cuda(p: torch.Tensor):
    if not p.is_cuda:
        raise ops._l.Pst3rError("panst3r_b200 modules run on CUDA only (no CPU fallback): move the module with .cuda()")


def w16(p: torch.Tensor) -> torch.Tensor:
    """bf16 contiguous 2-D weight [N, K] (conv kernels flattened)."""
    _check_cuda(p)
    return prepared("w16", [p], lambda: p.detach().reshape(p.shape[0], -1).to(torch.bfloat16).contiguous())


def f32(p: torch.Tensor) -> torch.Tensor:
    _check_cuda(p)
    return prepared("f32", [p], lambda: p.detach().to(torch.float32).contiguous())


def b16(p: torch.Tensor) -> torch.Tensor:
    _check_cuda(p)
    return prepared("b16", [p], lambda: p.detach().to(torch.bfloat16).contiguous())


def cat_w16(ps: Sequence[torch.Tensor]) -> torch.Tensor:
    return prepared("catw16", list(ps), lambda: torch.cat([p.detach().reshape(p.shape[0], -1) for p in ps], 0)
                    .to(torch.bfloat16).contiguous())


def cat_f32(ps: Sequence[torch.Tensor]) -> torch.Tensor:
    return prepared("catf32", list(ps), lambda: torch.cat([p.detach().reshape(-1) for p in ps], 0).to(torch.float32).contiguous())


_rope_tables: Dict[Tuple, torch.Tensor] = {}


def rope_table(maxpos: int, head_dim: int, base: float, device) -> torch.Tensor:
    """fp32 [maxpos, head_dim/4, 2] (cos, sin) of p * base^(-j / (head_dim/4)) — constant of (shape, base)."""
    key = (maxpos, head_dim, base, str(device))
    t = _rope_tables.get(key)
    if t is None:
        Q = head_dim // 4
        inv = base ** (-torch.arange(Q, dtype=torch.float64) / Q)
        ang = torch.arange(maxpos, dtype=torch.float64)[:, None] * inv
        t = torch.stack([ang.cos(), ang.sin()], -1).to(torch.float32).to(device).contiguous()
        _rope_tables[key] = t
    return t


_pos_grids: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}


def pos_grid(h: int, w: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """(int64 [h*w, 2], int32 [h*w, 2]) row-major (y, x) token positions (PatchEmbedDust3R convention)."""
    key = (h, w, str(device))
    t = _pos_grids.get(key)
    if t is None:
        ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        p = torch.stack([ys.flatten(), xs.flatten()], -1).to(device)
        t = (p.contiguous(), p.to(torch.int32).contiguous())
        _pos_grids[key] = t
    return t


class ViTBlockParams(nn.Module):
    """croco `Block` parameter container: norm1, attn.{qkv,proj}, norm2, mlp.{fc1,fc2}."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, eps: float = 1e-6, qkv_bias: bool = True):
        super().__init__()
        self.dim, self.num_heads, self.eps = dim, num_heads, eps
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = nn.Module()
        self.attn.qkv = nn.Linear(dim, 3 * dim, bias=qkv_bias)
        self.attn.proj = nn.Linear(dim, dim)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = nn.Module()
        self.mlp.fc1 = nn.Linear(dim, int(dim * mlp_ratio))
        self.mlp.fc2 = nn.Linear(int(dim * mlp_ratio), dim)


def bias_of(lin: nn.Linear):
    return None if lin.bias is None else f32(lin.bias)


def self_attention(x: torch.Tensor, blk, B: int, N: int, rope, in_place: bool = True) -> torch.Tensor:
    """x (+)= proj(attn(rope(qkv(norm1(x)))));  x bf16 [B*N, D].  rope = (table, pos_i32 [B*N, 2]) or None."""
    D, H = blk.dim, blk.num_heads
    h = ops.layernorm(x, f32(blk.norm1.weight), f32(blk.norm1.bias), blk.eps)
    qkv = ops.gemm(h, w16(blk.attn.qkv.weight), bias=bias_of(blk.attn.qkv),
                   rope=None if rope is None else (rope[0], rope[1], 2 * D))
    q5 = qkv.view(B, N, 3, H, D // H)
    o = ops.attention(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2])
    return ops.gemm(o.view(B * N, D), w16(blk.attn.proj.weight), bias=bias_of(blk.attn.proj), residual=x,
                    out=x if in_place else None)


def mlp_residual(x: torch.Tensor, norm: nn.LayerNorm, mlp, eps: float, out=None, act: int = ops.ACT_GELU) -> torch.Tensor:
    """out = x + fc2(act(fc1(norm(x))))  (in place when out is None)"""
    h = ops.layernorm(x, f32(norm.weight), f32(norm.bias), eps)
    h = ops.gemm(h, w16(mlp.fc1.weight), bias=bias_of(mlp.fc1), act=act)
    return ops.gemm(h, w16(mlp.fc2.weight), bias=bias_of(mlp.fc2), residual=x, out=x if out is None else out)


def vit_block(x: torch.Tensor, blk: ViTBlockParams, B: int, N: int, rope) -> torch.Tensor:
    x = self_attention(x, blk, B, N, rope)
    return mlp_residual(x, blk.norm2, blk.mlp, blk.eps)
