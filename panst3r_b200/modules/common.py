"""Shared host-side helpers for the CUDA modules: parameter preparation caches and the ViT block runner.

Modules in this package are *parameter containers with the reference's state-dict names* whose forward
passes call only panst3r_b200.ops (the C ABI).  Weights are converted once to the kernel formats
(bf16 [N, K] matrices, fp32 bias / norm vectors, fused or permuted variants) and cached per parameter version.
"""
from __future__ import annotations

import weakref
from typing import Callable, Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import ops

_cache: Dict[Tuple, Tuple[Tuple, Tuple, torch.Tensor]] = {}
_cache_sweep_at = 4096


def _ver(ps: Sequence[torch.Tensor]) -> Tuple:
    return tuple((p.data_ptr(), p._version, str(p.device), p.dtype) for p in ps)


def prepared(key: str, params: Sequence[torch.Tensor], build: Callable[[], torch.Tensor]) -> torch.Tensor:
    """Cache `build()` until any of `params` changes (load_state_dict / .to() bump version or data_ptr).
    For PARAMETERS (long-lived tensors) only — activations are converted per call.  Entries hold weak references to
    their parameters: a hit requires the very same live objects (CPython recycles id()s, the caching allocator
    recycles addresses), and entries whose parameters died are dropped."""
    global _cache_sweep_at
    k = (key, tuple(id(p) for p in params))
    v = _ver(params)
    hit = _cache.get(k)
    if hit is not None and hit[0] == v and all(r() is p for r, p in zip(hit[1], params)):
        return hit[2]
    with torch.no_grad():
        t = build()
    _cache[k] = (v, tuple(weakref.ref(p) for p in params), t)
    if len(_cache) > _cache_sweep_at:  # drop entries of parameters that no longer exist
        for kk in [kk for kk, e in _cache.items() if any(r() is None for r in e[1])]:
            del _cache[kk]
        _cache_sweep_at = max(4096, 2 * len(_cache))
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


def _split_cat(w32: torch.Tensor) -> torch.Tensor:
    """fp32 [N, K] -> bf16 [N, 2K] rows [hi | lo] with hi = bf16(w), lo = bf16(w - hi)."""
    hi = w32.to(torch.bfloat16)
    lo = (w32 - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], 1).contiguous()


def _as_split(buf: torch.Tensor) -> "ops.Split":
    k = buf.shape[-1] // 2
    return ops.Split(buf[..., :k], k)


def wsplit(p: torch.Tensor) -> "ops.Split":
    """Reference-precision weight operand [N, K] (ops.Split: bf16 hi | lo parts of the fp32 parameter)."""
    _check_cuda(p)
    return _as_split(prepared("wsplit", [p], lambda: _split_cat(p.detach().reshape(p.shape[0], -1).float())))


def cat_wsplit(ps: Sequence[torch.Tensor]) -> "ops.Split":
    return _as_split(prepared("catwsplit", list(ps),
                              lambda: _split_cat(torch.cat([p.detach().reshape(p.shape[0], -1).float() for p in ps], 0))))


def psplit(p: torch.Tensor) -> "ops.Split":
    """A parameter used as an ACTIVATION (learned queries, embeddings) in split form."""
    _check_cuda(p)
    return _as_split(prepared("psplit", [p], lambda: _split_cat(p.detach().reshape(-1, p.shape[-1]).float())))


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


# Folding pays where the chain is latency bound (M = 768 rows per step of the sequential memory build: one ~4 us
# LayerNorm launch less per Linear).  Measured on B200 (profiles/r01_stage_times.md): at M = 12288 the heavier
# epilogues cost more (+0.6 ms per 24-block encoder) than the LayerNorm kernels they replace, so the per-view stages
# (encoder, DINOv2, render) keep the stand-alone kernel — for every batch size, so that a rank holding a shard of the
# views computes bit-identical per-view results to a single GPU holding all of them (tests/test_gpu_dist.py).
FOLD_LN_MAX_ROWS = 3072


def fold_stats(rows: int, dim: int, device, count: int = 1, enable: bool = True):
    """`count` statistics buffers for a residual stream of `rows` rows, or Nones when folding is not used."""
    if not enable or rows > FOLD_LN_MAX_ROWS or dim % 64 != 0:
        return (None,) * count if count > 1 else None
    bufs = tuple(ops.new_stats(rows, dim, device) for _ in range(count))
    return bufs if count > 1 else bufs[0]


def folded_ln(norm: nn.LayerNorm, weights: Sequence[torch.Tensor], biases: Sequence[Optional[torch.Tensor]]):
    """LayerNorm folded into the Linear(s) that consume it:  LN(x) W^T + b = rstd (x (gamma o W)^T - mu colsum) + (W beta + b).
    Returns (bf16 gamma-scaled weights [N, K], fp32 colsum [N] of exactly those bf16 weights, fp32 bias' [N]);
    several Linears (q | k | v) are concatenated along N.  Cached per parameter version."""
    ps = [norm.weight, norm.bias] + list(weights) + [b for b in biases if b is not None]
    for p in ps:
        _check_cuda(p)

    def build():
        W = torch.cat([w.detach().reshape(w.shape[0], -1).float() for w in weights], 0)
        b = torch.cat([(bb.detach().float() if bb is not None else torch.zeros(w.shape[0], device=W.device))
                       for w, bb in zip(weights, biases)], 0)
        wf = (W * norm.weight.detach().float()[None, :]).to(torch.bfloat16).contiguous()
        return wf, wf.float().sum(1).contiguous(), (W @ norm.bias.detach().float() + b).contiguous()

    return prepared("folded_ln", ps, build)


def ln_linear(x: torch.Tensor, stats: Optional[torch.Tensor], norm: nn.LayerNorm, weights, biases, eps: float,
              plain_w=None, plain_b=None, **gemm_kw) -> torch.Tensor:
    """Linear(LayerNorm(x)).  With `stats` (partial row sums left by the GEMM that produced x) the normalisation is
    folded into the GEMM epilogue and never materialised; without, LayerNorm runs as its own kernel."""
    if stats is None:
        h = ops.layernorm(x, f32(norm.weight), f32(norm.bias), eps)
        w = plain_w() if plain_w is not None else w16(weights[0])
        b = plain_b() if plain_b is not None else (None if biases[0] is None else f32(biases[0]))
        return ops.gemm(h, w, bias=b, **gemm_kw)
    wf, colsum, bf = folded_ln(norm, weights, biases)
    return ops.gemm(x, wf, bias=bf, ln=(stats, colsum, eps), **gemm_kw)


def self_attention(x: torch.Tensor, blk, B: int, N: int, rope, in_place: bool = True,
                   stats: Optional[torch.Tensor] = None, stats_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x (+)= proj(attn(rope(qkv(norm1(x)))));  x bf16 [B*N, D].  rope = (table, pos_i32 [B*N, 2]) or None.
    stats: LayerNorm statistics of x (norm1 folds into the QKV GEMM); stats_out: receives those of the result."""
    D, H = blk.dim, blk.num_heads
    qkv = ln_linear(x, stats, blk.norm1, [blk.attn.qkv.weight], [blk.attn.qkv.bias], blk.eps,
                    rope=None if rope is None else (rope[0], rope[1], 2 * D))
    q5 = qkv.view(B, N, 3, H, D // H)
    o = ops.attention(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2])
    return ops.gemm(o.view(B * N, D), w16(blk.attn.proj.weight), bias=bias_of(blk.attn.proj), residual=x,
                    out=x if in_place else None, stats_out=stats_out)


def mlp_residual(x: torch.Tensor, norm: nn.LayerNorm, mlp, eps: float, out=None, act: int = ops.ACT_GELU,
                 stats: Optional[torch.Tensor] = None, stats_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = x + fc2(act(fc1(norm(x))))  (in place when out is None)"""
    h = ln_linear(x, stats, norm, [mlp.fc1.weight], [mlp.fc1.bias], eps, act=act)
    return ops.gemm(h, w16(mlp.fc2.weight), bias=bias_of(mlp.fc2), residual=x, out=x if out is None else out,
                    stats_out=stats_out)


def vit_block(x: torch.Tensor, blk: ViTBlockParams, B: int, N: int, rope, stats=None, stats_pair=None):
    """One croco Block.  With `stats` (+ two scratch statistics buffers) both LayerNorms are folded into the GEMMs
    that consume them; returns (x, stats of x)."""
    if stats is None:
        x = self_attention(x, blk, B, N, rope)
        return mlp_residual(x, blk.norm2, blk.mlp, blk.eps), None
    s1, s2 = stats_pair  # rotate: stats is one of them, always the one written last
    x = self_attention(x, blk, B, N, rope, stats=stats, stats_out=s1)
    return mlp_residual(x, blk.norm2, blk.mlp, blk.eps, stats=s1, stats_out=s2), s2
