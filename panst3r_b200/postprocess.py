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
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import ops


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
        groups = [(g.float().contiguous() if g.dtype != torch.float32 or not g.is_contiguous() else g, s) for g, s in groups]
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
            maps = [ops.panoptic_argmax(g, keep_d, sc_d, size, mask_threshold, areas[0], areas[1]) for g, size in groups]
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
