"""Oracle restatement of the upstream MUSt3R encoder / decoder (TEST INFRASTRUCTURE ONLY) — PARITY UNPINNED.

The reference constructs `Dust3rEncoder(img_size=[512,512], patch_embed='PatchEmbedDust3R')` and
`MUSt3R(img_size=[512,512], feedback_type='single_mlp', memory_mode='norm_y')` from the un-vendored `must3r`
package (reference configs/base.yaml:7-15, src/panst3r/panst3r.py:9; pyproject.toml:14 pins no commit) and only
ever calls them through two seams:
    x, pos = encoder(img, true_shape)                                     (engine/must3r.py:17-24)
    mem, pointmaps, feats = decoder(x, pos, true_shape, mem, render=, return_feats=True)   (:45, :93, :116)
with `mem = (mem_vals, mem_labels, mem_nimgs, mem_protected_imgs, mem_protected_tokens)` (:76) and
`mem_vals[l]` a (B, Nmem, D) tensor (:77-80, :104-106).  This file restates the published MUSt3R architecture
 behind exactly those seams.  Design decisions that cannot be verified offline are
listed in DESIGN.md ("MUSt3R restatement choices"); the CUDA path mirrors THIS file and parity is CUDA-vs-oracle.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from .blocks import Attention, Block, CrossAttention, Mlp, PatchEmbedDust3R, get_pos_embed


class Dust3rEncoder(nn.Module):
    """CroCo/DUSt3R ViT-L/16 encoder: patch-embed conv -> depth x RoPE Block -> LayerNorm(eps 1e-6)."""

    def __init__(self, img_size=(512, 512), patch_embed="PatchEmbedDust3R", patch_size=16, embed_dim=1024, depth=24,
                 num_heads=16, mlp_ratio=4.0, pos_embed="RoPE100"):
        super().__init__()
        assert patch_embed == "PatchEmbedDust3R"
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.patch_embed = PatchEmbedDust3R(img_size, patch_size, 3, embed_dim)
        self.rope = get_pos_embed(pos_embed)
        norm = partial(nn.LayerNorm, eps=1e-6)
        self.blocks_enc = nn.ModuleList(
            [Block(embed_dim, num_heads, mlp_ratio, qkv_bias=True, norm_layer=norm, rope=self.rope) for _ in range(depth)])
        self.norm_enc = norm(embed_dim)

    def forward(self, img, true_shape):
        x, pos = self.patch_embed(img, true_shape)
        for blk in self.blocks_enc:
            x = blk(x, pos)
        return self.norm_enc(x), pos


class MemoryDecoderBlock(nn.Module):
    """RoPE self-attention within the view -> cross-attention to memory tokens (no RoPE on memory keys: the memory
    5-tuple carries no positions) -> MLP.  `norm_y` is applied when tokens are WRITTEN to memory
    (memory_mode='norm_y'), so the read side uses the stored values as they are."""

    def __init__(self, dim, num_heads, mlp_ratio, rope, norm):
        super().__init__()
        self.norm1 = norm(dim)
        self.attn = Attention(dim, rope=rope, num_heads=num_heads, qkv_bias=True)
        self.norm2 = norm(dim)
        self.cross_attn = CrossAttention(dim, rope=None, num_heads=num_heads, qkv_bias=True)
        self.norm3 = norm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.norm_y = norm(dim)

    def forward(self, x, xpos, mem, mask, n=1):
        """x (B*n, N, D): the n views of each scene; mem (B, Nmem, D): that scene's memory; mask (B*n, 1, N, Nmem) or None."""
        x = x + self.attn(self.norm1(x), xpos)
        Bn, N, D = x.shape
        if mask is None:
            # render: the views only READ the memory, so they fold into the query axis — one K/V projection per scene
            # instead of one per view (row-wise identical to expanding the memory over the views)
            q = self.norm2(x).reshape(Bn // n, n * N, D)
            x = x + self.cross_attn(q, mem, mem, None, None).reshape(Bn, N, D)
        else:
            Nm = mem.shape[1]
            mem_b = mem[:, None].expand(Bn // n, n, Nm, D).reshape(Bn, Nm, D)
            x = x + self.cross_attn(self.norm2(x), mem_b, mem_b, None, None, mask)
        x = x + self.mlp(self.norm3(x))
        return x


class LinearHead(nn.Module):
    """LayerNorm -> Linear(D, C * P * P) -> pixel_shuffle(P) -> (B, H, W, C) raw pointmap channels
    (pts3d 3 + pts3d_local 3 + conf 1; activations are applied by the caller, tools/demo_panst3r.py:220-221)."""

    def __init__(self, dim, patch_size, channels):
        super().__init__()
        self.patch_size, self.channels = patch_size, channels
        self.proj = nn.Linear(dim, channels * patch_size * patch_size)

    def forward(self, tokens, hw):
        B, N, D = tokens.shape
        h, w = hw[0] // self.patch_size, hw[1] // self.patch_size
        f = self.proj(tokens).transpose(-1, -2).reshape(B, -1, h, w)
        return F.pixel_shuffle(f, self.patch_size).permute(0, 2, 3, 1)


class MUSt3R(nn.Module):
    def __init__(self, img_size=(512, 512), feedback_type="single_mlp", memory_mode="norm_y", enc_embed_dim=1024,
                 embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, patch_size=16, pos_embed="RoPE100",
                 head_channels=7):
        super().__init__()
        assert feedback_type == "single_mlp" and memory_mode == "norm_y"
        self.embed_dim, self.depth, self.patch_size = embed_dim, depth, patch_size
        norm = partial(nn.LayerNorm, eps=1e-6)
        self.rope = get_pos_embed(pos_embed)
        self.feat_embed_enc_to_dec = nn.Linear(enc_embed_dim, embed_dim)
        self.image2_embed = nn.Parameter(torch.zeros(1, 1, embed_dim))
        nn.init.normal_(self.image2_embed, std=0.02)
        self.blocks_dec = nn.ModuleList(
            [MemoryDecoderBlock(embed_dim, num_heads, mlp_ratio, self.rope, norm) for _ in range(depth)])
        self.feedback_layer = Mlp(embed_dim, int(mlp_ratio * embed_dim), embed_dim)
        self.norm_dec = norm(embed_dim)
        self.head_dec = LinearHead(embed_dim, patch_size, head_channels)

    def forward(self, x, pos, true_shape, mem=None, render=False, return_feats=False):
        """x (B, n, N, Denc), pos (B, n, N, 2), true_shape (B, n, 2).  Returns (mem, pointmaps (B,n,H,W,C), feats)."""
        B, n, N, _ = x.shape
        H, W = int(true_shape[0, 0, 0]), int(true_shape[0, 0, 1])
        h = self.feat_embed_enc_to_dec(x)
        # every view except the scene's very first one is tagged "not the reference image"
        tag = torch.ones(n, device=x.device, dtype=h.dtype)
        if mem is None and not render:
            tag[0] = 0
        h = h + tag.view(1, n, 1, 1) * self.image2_embed

        if mem is None:
            assert not render, "render needs a memory"
            mem_vals = [h.new_zeros(B, 0, self.embed_dim) for _ in range(self.depth)]
            mem_labels = torch.zeros(B, 0, dtype=torch.long, device=x.device)
            mem_nimgs = 0
        else:
            mem_vals, mem_labels, mem_nimgs = mem[0], mem[1], mem[2]

        new_labels = (mem_nimgs + torch.arange(n, device=x.device)).view(1, n, 1).expand(B, n, N)  # (B, n, N)
        layer_in = []
        hv = h.reshape(B * n, N, self.embed_dim)
        posv = pos.reshape(B * n, N, 2)
        feats = [h]
        for l, blk in enumerate(self.blocks_dec):
            layer_in.append(hv.view(B, n, N, -1))
            if render:
                mem_l = mem_vals[l]
                mask = None
            else:
                # candidates = stored memory + this batch's own (normalised) layer inputs; a view never attends
                # to tokens carrying its own label (the 2-view initialisation attends "to the other image")
                fresh = blk.norm_y(hv).view(B, n * N, -1)
                mem_l = torch.cat([mem_vals[l], fresh], dim=1)
                labels = torch.cat([mem_labels, new_labels.reshape(B, n * N)], dim=1)  # (B, Nmem + nN)
                mask = labels[:, None, None, :] == new_labels.reshape(B, n, N)[..., None]  # (B, n, N, Ncand)
                mask = mask.reshape(B * n, 1, N, -1)
            hv = blk(hv, posv, mem_l, mask, n)
            feats.append(hv.view(B, n, N, -1))

        pointmaps = self.head_dec(self.norm_dec(hv), (H, W)).view(B, n, H, W, -1)
        if H > W:  # portrait: predicted in the true orientation, returned in the landscape storage convention
            pointmaps = pointmaps.transpose(2, 3)

        if not render:
            # feedback ('single_mlp'): one MLP of the last-layer tokens is added to what every layer stores
            fb = self.feedback_layer(hv).view(B, n, N, -1)
            new_vals = []
            for l, blk in enumerate(self.blocks_dec):
                stored = blk.norm_y(layer_in[l] + fb).reshape(B, n * N, -1)
                new_vals.append(torch.cat([mem_vals[l], stored], dim=1))
            mem = (new_vals, torch.cat([mem_labels, new_labels.reshape(B, n * N)], dim=1), mem_nimgs + n, None, None)
        return mem, pointmaps, (feats if return_feats else None)
