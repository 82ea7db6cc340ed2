"""CPU restatement of the reference's panoptic post-processing (TEST INFRASTRUCTURE ONLY — never imported by the
product path).  Follows /root/reference/src/panst3r/engine/postprocess.py:

  * :18-27   mask logits -> sigmoid -> bilinear resize (align_corners=False) to the image size
  * :38-45   label_mode 'sigmoid': score / label = max over classes of sigmoid(class logits); keep score > cls_threshold
  * :63      score-weighted probabilities
  * :66-120  `niters` rounds of: per-pixel argmax over the surviving queries; a query keeps its segment when the
             part of its >= mask_threshold area that wins the argmax is at least overlap_threshold of its >= 0.5
             area; segment ids are consecutive in query order; conf = the winning query's mask probability

Pinned against the reference function itself (tests/test_oracle_vs_reference.py, this container only) and against
tests/golden/postprocess_v2_*.pt generated from it by oracle/make_golden.py.  Only label_mode='sigmoid' without
temperature is restated (the configuration of configs/base.yaml and tools/demo_panst3r.py).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn.functional as F


def class_scores(mask_cls_i: torch.Tensor):
    """(Q, K) class logits -> (scores (Q,), labels (Q,))   [postprocess.py:39]"""
    return mask_cls_i.sigmoid().max(-1)


def upsampled_probabilities(mask_logits: torch.Tensor, size) -> torch.Tensor:
    """(V, Q, h, w) logits -> (V, Q, H, W) probabilities   [postprocess.py:24-25]"""
    return F.interpolate(mask_logits.sigmoid(), size=tuple(int(s) for s in size), mode="bilinear", align_corners=False)


@torch.no_grad()
def panoptic_inference_v2(mask_cls, mask_pred, true_shape, cls_threshold=0.1, mask_threshold=0.25, overlap_threshold=0.5,
                          niters=2, void_confidence=0.1, multi_ar=False):
    """mask_cls (B, Q, K); mask_pred: multi_ar -> list of V tensors (Q, h, w) or (1, Q, h, w), true_shape (V, 2);
    otherwise a (B, V, Q, h, w) tensor and true_shape = (H, W).  Returns the reference's list of
    {'pan', 'segments_info', 'conf'} (one entry per batch element)."""
    if multi_ar:
        ups = []
        for m, ts in zip(mask_pred, true_shape):
            m4 = m if m.dim() == 4 else m[None]
            ups.append(upsampled_probabilities(m4.float(), ts)[0])  # (Q, H_i, W_i)
        Hm, Wm = max(u.shape[-2] for u in ups), max(u.shape[-1] for u in ups)
        probs = torch.zeros((1, len(ups), ups[0].shape[0], Hm, Wm))
        for i, u in enumerate(ups):
            probs[0, i, :, :u.shape[-2], :u.shape[-1]] = u
    else:
        probs = torch.stack([upsampled_probabilities(mask_pred[b].float(), true_shape) for b in range(len(mask_pred))])
    results = []
    for b in range(mask_cls.shape[0]):
        scores, labels = class_scores(mask_cls[b].float())
        keep = torch.nonzero(scores > cls_threshold).flatten()
        masks = probs[b].transpose(0, 1)[keep]  # (Qk, V, H, W)
        sc, cl, qid = scores[keep], labels[keep], keep.clone()
        pan = torch.zeros(probs.shape[1:2] + probs.shape[-2:], dtype=torch.int32)
        conf = torch.full(pan.shape, float(void_confidence))
        segments: List[dict] = []
        for _ in range(niters):
            pan = torch.zeros_like(pan)
            conf = torch.full(pan.shape, float(void_confidence))
            segments = []
            if masks.shape[0] == 0:
                break
            winner = (sc.view(-1, 1, 1, 1) * masks).argmax(0)
            selected = []
            for k in range(masks.shape[0]):
                area_half = int((masks[k] >= 0.5).sum())
                won = (winner == k) & (masks[k] >= mask_threshold)
                area_won = int(won.sum())
                if area_won == 0 or area_half == 0 or area_won / area_half < overlap_threshold:
                    continue
                selected.append(k)
                seg_id = len(segments) + 1
                pan[won] = seg_id
                conf[won] = masks[k][won]
                segments.append({"id": seg_id, "query_id": int(qid[k]), "category_id": int(cl[k])})
            sel = torch.tensor(selected, dtype=torch.int64)
            masks, sc, cl, qid = masks[sel], sc[sel], cl[sel], qid[sel]
        if multi_ar:
            pan = [pan[i, :int(h), :int(w)].contiguous() for i, (h, w) in enumerate(true_shape)]
            conf = [conf[i, :int(h), :int(w)].contiguous() for i, (h, w) in enumerate(true_shape)]
        results.append({"pan": pan, "segments_info": segments, "conf": conf})
    return results


def synthetic_scene(V: int, Q: int, K: int, h: int, w: int, seed: int, blobs: int = 6):
    """Well-conditioned synthetic head outputs: `blobs` queries own smooth regions with confident logits, the others
    are weak / low-score distractors.  All values are fp16-representable, so every implementation reads identical
    inputs.  Returns (mask_cls (1, Q, K), mask_logits (1, V, Q, h, w))."""
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    logits = torch.empty(V, Q, h, w)
    for q in range(Q):
        cy, cx = torch.rand(V, generator=g), torch.rand(V, generator=g)
        rad = 0.15 + 0.25 * torch.rand(1, generator=g)
        amp = 7.0 if q < blobs else 2.0 * torch.rand(1, generator=g).item()
        d2 = (ys[None] - cy[:, None, None]) ** 2 + (xs[None] - cx[:, None, None]) ** 2
        logits[:, q] = amp * (1.0 - d2 / rad ** 2).clamp(min=-1.0) + 0.3 * torch.randn(V, h, w, generator=g)
    cls = torch.randn(1, Q, K, generator=g) - 3.0
    for q in range(Q):
        if q < blobs or torch.rand(1, generator=g).item() < 0.3:
            cls[0, q, int(torch.randint(0, K, (1,), generator=g))] = 1.0 + 3.0 * torch.rand(1, generator=g).item()
    return cls.half().float(), logits[None].half().float()
