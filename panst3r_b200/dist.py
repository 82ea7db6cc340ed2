"""Multi-GPU execution of one scene: views sharded across ranks (one process per GPU, torch.distributed / NCCL).

The path shards per view: patch-embed + encoder, DINOv2, the render pass, the feature upscaler, the
full-resolution mask einsum and the pointmap head are independent per view and run on the owning rank only.
Two things are not per-view and are REPLICATED on every rank (identical results, no further exchange):
the sequential memory build (engine/must3r.py:40-54 is a chain over views by construction) and the 200-query mask
transformer (one query set attends to all views, mask_transformer.py:134-146).  Exchanges (all over NVLink):
  1. all-gather of encoder tokens (V, N, 1024) bf16 ahead of the cross-view decoder  — the one the north star names;
  2. all-gather of the stride-16 head features (V, N, 768) and the centre-pooled mask features (V, N, Cm) ahead of
     the replicated query decoder (1.2 MB + 0.4 MB per view).
Outputs stay sharded: each rank returns pointmaps / mask logits of its own views; class logits and queries are
replicated.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def partition_views(num_views: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous blocks [start, end) per rank; the first (num_views % world) ranks get one extra view."""
    base, extra = divmod(num_views, world)
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((s, s + n))
        s += n
    return out


def gather_rows(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """All-gather row blocks of possibly different heights: local (counts[rank], ...) -> (sum(counts), ...), in
    rank order.  Equal counts use one all_gather_into_tensor; ragged counts pad to the maximum."""
    world = dist.get_world_size(group)
    assert len(counts) == world and local.shape[0] == counts[dist.get_rank(group)]
    local = local.contiguous()
    if len(set(counts)) == 1:
        out = torch.empty((world * counts[0], *local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    m = max(counts)
    pad = torch.zeros((m, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


class ShardedPanSt3R:
    """forward(imgs, true_shape, classes) -> (panout, pointmaps) for THIS rank's views.

    `imgs` is the full (1, V, 3, H, W) batch on every rank (each rank touches only its slice); returned
    `pred_masks` / `pointmaps` cover views [start, end) = partition_views(V, world)[rank]."""

    def __init__(self, model, rank: int, world: int, group=None):
        self.model, self.rank, self.world, self.group = model, rank, world, group

    @torch.no_grad()
    def __call__(self, imgs: torch.Tensor, true_shape: torch.Tensor, classes, outdevice=None):
        from . import ops
        from .panst3r import DEC_DIM, DINO_DIM, ENC_DIM
        m = self.model
        B, V, _, H, W = imgs.shape
        if B != 1:
            raise ops._l.Pst3rError("one scene per call (B == 1)")
        parts = partition_views(V, self.world)
        counts = [e - s for s, e in parts]
        s, e = parts[self.rank]
        nv = e - s
        ts = true_shape.cpu() if true_shape.is_cuda else true_shape
        if int(ts.reshape(-1, 2)[0, 0]) != H or int(ts.reshape(-1, 2)[0, 1]) != W:
            raise ops._l.Pst3rError("the sharded runner handles landscape scenes (true_shape == tensor shape) only")
        P = m.must3r_encoder.patch_size
        hs, ws = H // P, W // P
        N = hs * ws
        dev = imgs.device
        # ---- per-view producers on the owning rank, written into one concatenated feature buffer
        cat = torch.empty((1, max(nv, 1), N, ENC_DIM + DEC_DIM + DINO_DIM), device=dev, dtype=torch.bfloat16)
        rows = cat.view(-1, cat.shape[-1])
        my_imgs, my_ts = imgs[:, s:e], ts[:, s:e]
        cur = torch.cuda.current_stream()
        side = m._side_stream or torch.cuda.Stream()
        m._side_stream = side
        dino_sms = m.dino_sms if (nv > 0 and m.overlap_dino) else 0
        if nv > 0:
            x_loc, _ = m.forward_must3r_encoder(my_imgs, my_ts, out=rows[:, :ENC_DIM])
            x_loc = x_loc[0]
            # DINOv2 of the local views shares the GPU with the (replicated, latency-bound) memory build below
            side.wait_stream(cur)
            with torch.cuda.stream(side), ops.sm_budget(dino_sms):
                m.forward_dino(my_imgs, my_ts, out=rows[:, ENC_DIM + DEC_DIM:])
        else:
            x_loc = torch.empty((0, N, ENC_DIM), device=dev, dtype=torch.bfloat16)
        # ---- exchange 1: encoder tokens of every view, then the replicated sequential memory build
        x_all = gather_rows(x_loc, counts, self.group).view(1, V, N, ENC_DIM)
        from .modules.common import pos_grid
        pos_all = pos_grid(hs, ws, dev)[0][None, None].expand(1, V, N, 2)
        with ops.sm_budget(max(ops.num_sms() - dino_sms, 8) if dino_sms else 0):
            mem = m.build_memory(x_all, pos_all, ts)
        if nv > 0:
            cur.wait_stream(side)
        pointmaps = None
        mt = m.panoptic_decoder.mask_transformer
        if nv > 0:
            _, pointmaps, _ = m.must3r_decoder(x_all[:, s:e], pos_all[:, s:e], my_ts, mem, render=True, return_feats="last",
                                               feats_out=rows[:, ENC_DIM:ENC_DIM + DEC_DIM])
            src_loc, mask_f = m.panoptic_decoder.upscaler.forward_nhwc(rows, nv, hs, ws, f16_extra_bias=mt.level_embed.weight)
            Cm = mask_f.shape[-1]
            pooled_loc = ops.center_pool8(mask_f).view(nv * N, Cm)
        else:
            Cm = mt.mask_dim
            src_loc = torch.empty((0, mt.hidden_dim), device=dev, dtype=torch.bfloat16)
            pooled_loc = torch.empty((0, Cm), device=dev, dtype=torch.bfloat16)
            mask_f = torch.empty((0, 8 * hs, 8 * ws, Cm), device=dev, dtype=torch.bfloat16)
        # ---- exchange 2: stride-16 features + centre-pooled mask features for the replicated query decoder
        tok_counts = [c * N for c in counts]
        src_all = gather_rows(src_loc, tok_counts, self.group)
        pooled_all = gather_rows(pooled_loc, tok_counts, self.group)
        cls_emb = m.panoptic_decoder.text_encoder(classes, device=dev)
        panout = mt.forward_nhwc(src_all, mask_f, (hs, ws), cls_emb, deep_supervision=m.panoptic_decoder.deep_supervision,
                                 pooled=pooled_all)
        return panout, pointmaps
