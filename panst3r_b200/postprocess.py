"""GPU panoptic post-processing with the reference's interface
(reference src/panst3r/engine/postprocess.py:8-131: `panoptic_inference_v1`, `panoptic_inference_v2`).

    results = panoptic_inference_v2(mask_cls, mask_pred, true_shape, label_mode='sigmoid', cls_threshold=0.1,
                                    temperature=None, mask_threshold=0.25, overlap_threshold=0.5, niters=2,
                                    void_confidence=0.1, device=None, multi_ar=False)
    results[b] = {'pan': int32 segment ids, 'segments_info': [{'id', 'query_id', 'category_id'}], 'conf': fp32}

The per-pixel work (sigmoid, bilinear resize to the image size, score-weighted argmax over the surviving queries,
the two per-query pixel counts of the filtering rule, the final id / confidence maps) runs in
libpanst3r_b200.so (csrc/postprocess.cu); the host only walks the <= Q surviving queries of each round over two small
counter arrays, exactly the reference's loop (:77-113) without its per-query `.item()` reductions over full-size maps.
Only label_mode='sigmoid' without temperature (configs/base.yaml, tools/demo_panst3r.py) is implemented.

`LazyMasks` (what `PanopticDecoder.lazy_masks = True` puts into `pred_masks`): the mask einsum
"bqc,bnchw->bnqhw" (mask_transformer.py:279-280) is NOT evaluated by the head; `mask_pred` then holds the final
mask embeddings and the pixel features, and every post-processing round produces the logits chunk by chunk (one band of
one view per tcgen05 GEMM launch) into ONE scratch buffer small enough to stay in the 126 MB L2, consumed at once by
the band form of the argmax kernel.  The (V, Q, h, w) fp32 tensor the reference writes, copies and re-reads
(629 MB at 16 views of 512x384) never exists; ids, segments and confidences are those of the materialised path.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import ops


class LazyMasks:
    """Mask logits of one stack of equally shaped views in factored form: logits[v, q, y, x] = <feats[v, y, x, :], embed[q, :]>.
    feats: (V, hm, wm, C) pixel-major, embed: (Q, C); both bf16 tensors, or both ops.Split pairs (reference-precision head).
    Stands where the reference's `pred_masks` tensor stands: `shape` / `dim()` / indexing by batch and view follow a
    (1, V, Q, hm, wm) tensor (ndim 5: leading batch of one; 4: a stack; 3: one view), `.to(cuda device)` is the identity and
    `.to('cpu')` / `materialize()` evaluate the full tensor with the same GEMM the eager head uses."""

    # bytes of fp32 logits produced per GEMM launch: 200 queries x 96 rows x 256 pixels = 19.7 MB, half a 512x384 view
    # (measured: profiles/r02_lazy_masks.md)
    scratch_bytes = 20 << 20
    # chunks in flight, each with its own scratch.  1: the logits stay in the L2 (DRAM traffic = features + maps + 10 %);
    # 2 with 40 MB chunks halves the time of a round but writes most of the logits to DRAM once (profiles/r02_lazy_masks.md)
    streams = 1
    _lanes = {}          # (device, launching stream) -> [[side stream, scratch buffer], ...]

    def __init__(self, feats, embed, ndim: int = 4):
        if feats.shape[-1] != embed.shape[-1] or len(feats.shape) != 4 or len(embed.shape) != 2:
            raise ops._l.Pst3rError(f"LazyMasks: feats (V, hm, wm, C) / embed (Q, C) expected, got {tuple(feats.shape)} / {tuple(embed.shape)}")
        if isinstance(feats, ops.Split) != isinstance(embed, ops.Split):
            raise ops._l.Pst3rError("LazyMasks: features and embeddings must be both bf16 or both split pairs")
        if ndim == 3 and feats.shape[0] != 1:
            raise ops._l.Pst3rError("LazyMasks: a single-view handle needs exactly one view")
        self.feats, self.embed, self._ndim = feats, embed, ndim

    # ---- tensor-like surface -------------------------------------------------------------------------
    @property
    def shape(self):
        V, hm, wm, _ = self.feats.shape
        core = (self.embed.shape[0], hm, wm)
        return torch.Size({3: core, 4: (V, *core), 5: (1, V, *core)}[self._ndim])

    def dim(self) -> int:
        return self._ndim

    def __len__(self) -> int:
        return self.shape[0]

    @property
    def device(self):
        return self.feats.device

    dtype = torch.float32
    is_cuda = True

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            out = self
            for i in idx:
                out = out[i]
            return out
        if idx is None:
            if self._ndim == 5:
                raise ops._l.Pst3rError("LazyMasks: already has a batch dimension")
            return LazyMasks(self.feats, self.embed, self._ndim + 1)
        if self._ndim == 5:
            if isinstance(idx, slice):
                if idx.indices(1) != (0, 1, 1):
                    raise IndexError("LazyMasks: the batch dimension has one entry")
                return self
            if int(idx) not in (0, -1):
                raise IndexError("LazyMasks: the batch dimension has one entry")
            return LazyMasks(self.feats, self.embed, 4)
        if self._ndim == 4:
            V = self.feats.shape[0]
            if isinstance(idx, slice):
                a, b, st = idx.indices(V)
                if st != 1 or b <= a:
                    raise IndexError("LazyMasks: contiguous, non-empty view ranges only")
                return LazyMasks(self.feats[a:b], self.embed, 4)
            i = int(idx) + (V if int(idx) < 0 else 0)
            if not 0 <= i < V:
                raise IndexError(f"LazyMasks: view {idx} of {V}")
            return LazyMasks(self.feats[i:i + 1], self.embed, 3)
        raise IndexError("LazyMasks: a single view is indexed by materialising it")

    def to(self, device=None, *a, **k):
        """Same CUDA device / fp32: the handle itself.  Anything else (another device, another dtype) needs the values:
        the tensor is evaluated and converted."""
        if device is None or device == torch.float32:
            return self
        if not isinstance(device, torch.dtype) and torch.device(device).type == "cuda":
            return self
        return self.materialize().to(device, *a, **k)

    def float(self):
        return self

    def contiguous(self):
        return self

    # ---- evaluation ----------------------------------------------------------------------------------
    def logits(self, v: int, r0: int, r1: int, out: torch.Tensor = None) -> torch.Tensor:
        """fp32 [Q, r1 - r0, wm]: source rows [r0, r1) of view v — one GEMM launch with a plane-major store."""
        V, hm, wm, C = self.feats.shape
        Q = self.embed.shape[0]
        n = (r1 - r0) * wm
        if out is None:
            out = torch.empty((Q, r1 - r0, wm), device=self.device, dtype=torch.float32)
        rows = self.feats.view(V * hm * wm, C)[(v * hm + r0) * wm:(v * hm + r1) * wm]
        ops.gemm(rows, self.embed, out=out, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=n, batch_stride=0, ldt=n)
        return out

    def materialize(self) -> torch.Tensor:
        """The tensor this object stands for, (…, Q, hm, wm) fp32 with this handle's leading dims (one GEMM launch)."""
        V, hm, wm, C = self.feats.shape
        Q = self.embed.shape[0]
        mk = torch.empty((V, Q, hm, wm), device=self.device, dtype=torch.float32)
        ops.gemm(self.feats.view(V * hm * wm, C), self.embed, out=mk, store_mode=ops.STORE_TRANSPOSED,
                 rows_per_batch=hm * wm, batch_stride=Q * hm * wm, ldt=hm * wm)
        return {3: mk[0], 4: mk, 5: mk[None]}[self._ndim]

    def band_plan(self, H: int, scratch_bytes: int = None):
        """[(y0, rows, src_row0, src_rows)]: output-row bands (multiples of the kernel's 32-row tiles) whose source rows fit
        the scratch budget, each with the source rows its bilinear resize reads (one spare row on either side: the
        kernel's fp32 index arithmetic is re-checked by the library)."""
        _, hm, wm, _ = self.feats.shape
        Q = self.embed.shape[0]
        budget = self.scratch_bytes if scratch_bytes is None else int(scratch_bytes)
        max_src = max(4, budget // (Q * wm * 4))
        scale = hm / H

        def src_range(y0, rows):
            lo = int(np.floor(scale * (y0 + 0.5) - 0.5)) - 1
            hi = int(np.floor(scale * (y0 + rows - 1 + 0.5) - 0.5)) + 2
            return max(lo, 0), min(hi, hm - 1)

        step = max(32, int((max_src - 4) / scale) // 32 * 32)
        plan, y0 = [], 0
        while y0 < H:
            rows = min(step, H - y0)
            lo, hi = src_range(y0, rows)
            plan.append((y0, rows, lo, hi - lo + 1))
            y0 += rows
        return plan

    def panoptic_argmax(self, keep_idx, keep_scores, size, mask_threshold, area_half, area_won, scratch_bytes: int = None,
                        streams: int = None):
        """`ops.panoptic_argmax` without the logits tensor: per view and band, GEMM into a scratch buffer -> band argmax.
        A band's argmax is a short grid of latency-bound CTAs (one 32 x 32 tile each, walking the kept queries), so the
        (view, band) chunks are dealt round-robin to `streams` side streams with one scratch buffer each: the GEMM of
        one chunk runs next to the argmax of another.  Forked from / joined to the current stream; the counters are
        integer atomics, so the result does not depend on the interleaving."""
        V, hm, wm, _ = self.feats.shape
        Q = self.embed.shape[0]
        H, W = int(size[0]), int(size[1])
        dev = self.device
        ids = torch.empty((V, H, W), device=dev, dtype=torch.int32)
        win = torch.empty((V, H, W), device=dev, dtype=torch.float32)
        plan = self.band_plan(H, scratch_bytes)
        need = max(p_[3] for p_ in plan) * Q * wm
        n_str = max(1, min(int(self.streams if streams is None else streams), V * len(plan)))
        cur = torch.cuda.current_stream(dev)
        if n_str == 1:  # everything on the launching stream
            lanes = LazyMasks._lanes.setdefault((dev, cur.cuda_stream, "solo"), [[cur, None]])
            lanes[0][0] = cur
        else:           # side streams of this (device, launching stream), created once
            lanes = LazyMasks._lanes.setdefault((dev, cur.cuda_stream), [])
            while len(lanes) < n_str:
                lanes.append([torch.cuda.Stream(dev), None])
        for lane in lanes[:n_str]:
            if lane[0] is not cur:
                lane[0].wait_stream(cur)  # fork: ids / win / counters / keep lists are ready
            with torch.cuda.stream(lane[0]):
                if lane[1] is None or lane[1].numel() < need:
                    lane[1] = torch.empty(need, device=dev, dtype=torch.float32)
        c = 0
        for v in range(V):
            for y0, rows, s0, sr in plan:
                st, buf = lanes[c % n_str]
                c += 1
                with torch.cuda.stream(st):
                    chunk = self.logits(v, s0, s0 + sr, out=buf[:Q * sr * wm].view(Q, sr, wm))
                    ops.panoptic_argmax(chunk[None], keep_idx, keep_scores, (H, W), mask_threshold, area_half, area_won,
                                        out=(ids[v:v + 1], win[v:v + 1]), band=(y0, rows, s0, hm))
        for lane in lanes[:n_str]:
            if lane[0] is not cur:
                cur.wait_stream(lane[0])  # join
        return ids, win


def _argmax(g, *args):
    return g.panoptic_argmax(*args) if isinstance(g, LazyMasks) else ops.panoptic_argmax(g, *args)


def _views(mask_pred, true_shape, multi_ar: bool):
    """-> list over batch of lists of (masks fp32 CUDA [n, Q, h, w], (H, W)) view groups of equal shape."""
    if multi_ar:
        ts = np.asarray(true_shape.cpu() if torch.is_tensor(true_shape) else true_shape).reshape(-1, 2)
        groups = []
        for m, (h, w) in zip(mask_pred, ts):
            m4 = m if m.dim() == 4 else m[None]
            groups.append((m4, (int(h), int(w))))
        return [groups]
    size = tuple(int(s) for s in (true_shape.tolist() if torch.is_tensor(true_shape) else true_shape))
    return [[(mask_pred[b], size)] for b in range(len(mask_pred))]


@torch.no_grad()
def panoptic_inference_v2(mask_cls, mask_pred, true_shape, label_mode="sigmoid", cls_threshold=0.1, temperature=None,
                          mask_threshold=0.25, overlap_threshold=0.5, niters=2, void_confidence=0.1, device=None,
                          multi_ar=False):
    if label_mode != "sigmoid" or temperature is not None:
        raise NotImplementedError("the CUDA post-processing implements label_mode='sigmoid' without temperature")
    batches = _views(mask_pred, true_shape, multi_ar)
    results = []
    for b, groups in enumerate(batches):
        dev = groups[0][0].device
        if dev.type != "cuda":
            raise ops._l.Pst3rError("panoptic_inference (panst3r_b200) runs on CUDA tensors only (no CPU fallback)")
        groups = [(g if isinstance(g, LazyMasks) or (g.dtype == torch.float32 and g.is_contiguous()) else g.float().contiguous(), s)
                  for g, s in groups]
        scores_d, labels_d = ops.class_scores(mask_cls[b].to(dev).float().contiguous())
        scores, labels = scores_d.cpu().numpy(), labels_d.cpu().numpy()
        keep = np.nonzero(scores > np.float32(cls_threshold))[0].astype(np.int32)
        maps = None  # per group (ids, win) of the last round that ran
        lut_d = None
        segments: List[dict] = []
        nkeep_last = 0
        for _ in range(niters):
            segments, lut_d, maps, nkeep_last = [], None, None, 0
            if keep.size == 0:
                break
            keep_d = torch.from_numpy(keep).to(dev)
            sc_d = scores_d[keep_d.long()].contiguous()
            areas = torch.zeros((2, keep.size), device=dev, dtype=torch.int32)
            maps = [_argmax(g, keep_d, sc_d, size, mask_threshold, areas[0], areas[1]) for g, size in groups]
            area_half, area_won = areas.cpu().numpy()
            lut = np.zeros(keep.size, dtype=np.int32)
            selected = []
            for k in range(keep.size):
                if area_won[k] > 0 and area_half[k] > 0 and area_won[k] / area_half[k] >= overlap_threshold:
                    selected.append(k)
                    lut[k] = len(segments) + 1
                    segments.append({"id": int(lut[k]), "query_id": int(keep[k]), "category_id": int(labels[keep[k]])})
            lut_d, nkeep_last = torch.from_numpy(lut).to(dev), keep.size
            keep = keep[np.asarray(selected, dtype=np.int64)] if selected else keep[:0]
        pans, confs = [], []
        for gi, (g, size) in enumerate(groups):
            n = g.shape[0]
            if maps is None:  # nothing survived the class threshold / the previous round
                pans.append(torch.zeros((n, *size), device=dev, dtype=torch.int32))
                confs.append(torch.full((n, *size), float(void_confidence), device=dev, dtype=torch.float32))
            else:
                p_, c_ = ops.panoptic_finalize(maps[gi][0], maps[gi][1], lut_d if nkeep_last else None, mask_threshold, void_confidence)
                pans.append(p_)
                confs.append(c_)
        if device is not None:
            pans, confs = [p_.to(device) for p_ in pans], [c_.to(device) for c_ in confs]
        if multi_ar:
            pan, conf = [p_[0] for p_ in pans], [c_[0] for c_ in confs]
        else:
            pan, conf = pans[0], confs[0]
        results.append({"pan": pan, "segments_info": segments, "conf": conf})
    return results


def panoptic_inference_v1(*args, mask_threshold=0.5, overlap_threshold=0.8, **kwargs):
    """postprocess.py:8-10"""
    return panoptic_inference_v2(*args, mask_threshold=mask_threshold, overlap_threshold=overlap_threshold, niters=1, **kwargs)
