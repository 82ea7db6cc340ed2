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


def _gather_any(local, counts, group=None):
    """gather_rows for bf16 tensors or packed ops.Split matrices (their [hi | lo] rows travel as they are)."""
    from . import ops
    from .modules.common import _as_split
    if isinstance(local, ops.Split):
        return _as_split(gather_rows(local.full(), counts, group))
    return gather_rows(local, counts, group)


def allreduce_minmax(minmax: Optional[torch.Tensor], device, group=None) -> torch.Tensor:
    """Batch-global per-channel (min, max) of LoftUp's MinMaxScaler (model/upscalers/loftup.py:14-19) across ranks:
    the reference reduces over the WHOLE batch of views, so view-sharded ranks exchange their 3 x 2 extrema
    (one MAX all-reduce of [-min, max]).  A rank without views contributes the neutral element."""
    t = torch.full((3, 2), float("-inf"), device=device, dtype=torch.float32)
    if minmax is not None:
        t = torch.stack([-minmax[:, 0], minmax[:, 1]], dim=1).contiguous()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return torch.stack([-t[:, 0], t[:, 1]], dim=1).contiguous()


class ShardedPanSt3R:
    """forward(imgs, true_shape, classes) -> (panout, pointmaps) for THIS rank's views.

    `imgs` is the (1, V, 3, Hs, Ws) batch on every rank — only the slice of the local views is read, so a caller may
    leave the rest unfilled (bench.py uploads just that slice).  Returned `pred_masks` / `pointmaps` cover views
    [start, end) = partition_views(V, world)[rank].  Handles the v1 and v2 heads, landscape and portrait scenes
    (views stored transposed, true_shape (H > W)), ragged partitions (V % world != 0, ranks without views); all views
    of a scene share one shape (mixed-shape scenes go through the single-GPU forward_inference_multi_ar)."""

    def __init__(self, model, rank: int, world: int, group=None):
        self.model, self.rank, self.world, self.group = model, rank, world, group

    @torch.no_grad()
    def __call__(self, imgs: torch.Tensor, true_shape: torch.Tensor, classes, outdevice=None):
        from . import ops
        from .modules.common import pos_grid
        from .modules.must3r import _hw
        from .modules.panoptic import PixelShuffleUpscaler
        from .panst3r import DEC_DIM, DINO_DIM, ENC_DIM
        m = self.model
        B, V = imgs.shape[:2]
        if B != 1:
            raise ops._l.Pst3rError("one scene per call (B == 1)")
        parts = partition_views(V, self.world)
        counts = [e - s for s, e in parts]
        s, e = parts[self.rank]
        nv = e - s
        ts = true_shape.cpu() if true_shape.is_cuda else true_shape
        H, W = _hw(ts)  # true size; portrait views (H > W) are stored transposed and processed in their true orientation
        P = m.must3r_encoder.patch_size
        hs, ws = H // P, W // P
        N = hs * ws
        dev = imgs.device
        pd = m.panoptic_decoder
        mt = pd.mask_transformer
        pr = pd.precise
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
            if m.overlap_dino:
                # DINOv2 of the local views shares the GPU with the (replicated, latency-bound) memory build below
                side.wait_stream(cur)
                with torch.cuda.stream(side), ops.sm_budget(dino_sms):
                    m.forward_dino(my_imgs, my_ts, out=rows[:, ENC_DIM + DEC_DIM:])
            else:
                m.forward_dino(my_imgs, my_ts, out=rows[:, ENC_DIM + DEC_DIM:])
        else:
            x_loc = torch.empty((0, N, ENC_DIM), device=dev, dtype=torch.bfloat16)
        # ---- exchange 1: encoder tokens of every view, then the replicated sequential memory build
        x_all = gather_rows(x_loc, counts, self.group).view(1, V, N, ENC_DIM)
        pos_all = pos_grid(hs, ws, dev)[0][None, None].expand(1, V, N, 2)
        with ops.sm_budget(max(ops.num_sms() - dino_sms, 8) if dino_sms else 0):
            mem = m.build_memory(x_all, pos_all, ts)
        if nv > 0 and m.overlap_dino:
            cur.wait_stream(side)
        pointmaps = None
        portrait = H > W
        grid = (ws, hs) if portrait else (hs, ws)  # the head works in the landscape storage convention
        loftup = not isinstance(pd.upscaler, PixelShuffleUpscaler)
        if loftup:  # v2: LoftUp's MinMaxScaler is batch-global -> 3-channel min/max all-reduce across the view shards
            pd.upscaler.minmax_reduce = lambda mm: allreduce_minmax(mm, dev, self.group)
        try:
            if nv > 0:
                _, pointmaps, _ = m.must3r_decoder(x_all[:, s:e], pos_all[:, s:e], my_ts, mem, render=True, return_feats="last",
                                                   feats_out=rows[:, ENC_DIM:ENC_DIM + DEC_DIM])
                src_loc, mask_f, grid, portrait, _ = pd._stack_features(None, my_imgs, my_ts, cat)
                Cm = mask_f.shape[-1]
                pooled_loc = ops.center_pool8(mask_f).view(nv * N, Cm)
            else:
                if loftup:
                    allreduce_minmax(None, dev, self.group)
                Cm = mt.mask_dim
                mk = (lambda *sh: ops.Split.empty(sh, dev)) if pr else (lambda *sh: torch.empty(sh, device=dev, dtype=torch.bfloat16))
                src_loc, pooled_loc, mask_f = mk(0, mt.hidden_dim), mk(0, Cm), mk(0, 8 * grid[0], 8 * grid[1], Cm)
        finally:
            if loftup:
                pd.upscaler.minmax_reduce = None
        # ---- exchange 2: stride-16 features + centre-pooled mask features for the replicated query decoder
        tok_counts = [c * N for c in counts]
        src_all = _gather_any(src_loc, tok_counts, self.group)
        pooled_all = _gather_any(pooled_loc, tok_counts, self.group)
        cls_emb = pd.text_encoder(classes, device=dev, precise=pr)
        panout = mt.forward_nhwc(src_all, mask_f, grid, cls_emb, deep_supervision=pd.deep_supervision,
                                 pooled=pooled_all, portrait=portrait, precise=pr)
        if outdevice is not None:
            panout = pd._to_device(panout, outdevice, dev)
            pointmaps = None if pointmaps is None else pointmaps.to(outdevice)
        return panout, pointmaps
