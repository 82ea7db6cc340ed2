"""Deterministic, construction-order-independent synthetic weights (TEST INFRASTRUCTURE ONLY).

Every tensor is drawn from a generator seeded by a hash of its state-dict key, so the reference module, the
oracle module and the CUDA module all receive bit-identical parameters through `load_state_dict`, whatever
order their constructors create parameters in.  Scales keep activations O(1) through deep stacks
(random-init `exp`-type heads otherwise overflow, SURVEY.md §7 "hard parts").
"""
from __future__ import annotations

import hashlib
import math

import torch


def _seed(key: str, seed: int) -> int:
    return int.from_bytes(hashlib.sha256(f"{seed}:{key}".encode()).digest()[:8], "little") % (2 ** 63 - 1)


def synth_tensor(key: str, shape, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(_seed(key, seed))
    shape = tuple(shape)
    leaf = key.split(".")[-1]
    if len(shape) == 0:
        return torch.zeros(())
    if leaf in ("weight", "in_proj_weight") and len(shape) >= 2:
        fan_in = math.prod(shape[1:])
        return torch.randn(shape, generator=g) * (fan_in ** -0.5)
    if ("norm" in key or leaf == "weight") and len(shape) == 1 and leaf == "weight":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)  # norm gains
    if leaf in ("bias", "in_proj_bias"):
        return 0.02 * torch.randn(shape, generator=g)
    if leaf == "lambda1":  # DINOv2 LayerScale
        return 0.5 + 0.1 * torch.randn(shape, generator=g)
    if leaf in ("cls_token", "mask_token", "position_embeddings", "image2_embed", "lang_embed"):
        return 0.02 * torch.randn(shape, generator=g)
    if leaf == "biases":  # ImplicitFeaturizer phase offsets
        return torch.randn(shape, generator=g)
    return 0.5 * torch.randn(shape, generator=g)  # embeddings (query_feat, query_embed, level_embed)


def synth_state_dict(module: torch.nn.Module, seed: int = 0, prefix: str = "") -> dict:
    out = {}
    for k, v in module.state_dict().items():
        t = synth_tensor(prefix + k, v.shape, seed)
        out[k] = t.to(v.dtype) if v.dtype.is_floating_point else v.clone()
    return out


def load_synth(module: torch.nn.Module, seed: int = 0, prefix: str = "") -> torch.nn.Module:
    module.load_state_dict(synth_state_dict(module, seed, prefix), strict=True)
    return module


def synth_class_embeddings(names, dim=768, seed: int = 0) -> dict:
    return {n: synth_tensor(f"class_embeddings.{n}", (dim,), seed) for n in names}
