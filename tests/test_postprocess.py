"""Panoptic post-processing (reference engine/postprocess.py:8-131).

CPU: the oracle restatement (oracle/postprocess.py) equals the golden outputs of the REFERENCE function
(tests/golden/postprocess_synthetic.pt, written by `python -m oracle.make_golden postprocess`) bit for bit, and the
reference function itself where its checkout is present.
GPU: the CUDA path (panst3r_b200/postprocess.py -> csrc/postprocess.cu through the C ABI) reproduces the same
segments; pixel ids are exact wherever the decision is not a floating-point tie (score-weighted top-2 margin, or
distance to a threshold, above 1e-5 — the fused kernel evaluates sigmoid / bilinear weights in a different but
equivalent order than ATen, so values agree to ~1e-7, not bitwise).
"""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN
from oracle import postprocess as op
from oracle import ref_import

CASES = torch.load(os.path.join(GOLDEN, "postprocess_synthetic.pt"))
PARAMS = {"v2": dict(), "v1": dict(mask_threshold=0.5, overlap_threshold=0.8, niters=1)}


def scene_inputs(case):
    V, Q, K, h, w, seed = case["scene"]
    cls, logits = op.synthetic_scene(V, Q, K, h, w, seed)
    sizes = case["sizes"]
    masks = [logits[0, i][None, ..., :sizes[i][0] // 2, :sizes[i][1] // 2].contiguous() for i in range(V)]
    return cls, masks, np.array(sizes)


@pytest.mark.parametrize("ci", range(len(CASES)))
@pytest.mark.parametrize("which", ["v2", "v1"])
def test_oracle_equals_reference_golden(ci, which):
    cls, masks, ts = scene_inputs(CASES[ci])
    g = CASES[ci]["out"][which]
    o = op.panoptic_inference_v2(cls, masks, ts, multi_ar=True, **PARAMS[which])[0]
    assert o["segments_info"] == g["segments_info"]
    for a, b in zip(o["pan"], g["pan"]):
        assert torch.equal(a.to(torch.int16), b)
    for a, b in zip(o["conf"], g["conf"]):
        assert torch.equal(a, b)


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present (GPU box)")
def test_oracle_equals_reference_function_random_scenes():
    pp = ref_import.load_reference().postprocess
    for seed in range(10, 14):
        cls, logits = op.synthetic_scene(2, 30, 9, 12, 20, seed, blobs=5)
        masks = [logits[0, i][None] for i in range(2)]
        ts = np.array([[24, 40]] * 2)
        r = pp.panoptic_inference_v2(cls.clone(), [m.clone() for m in masks], ts, label_mode="sigmoid", device="cpu", multi_ar=True)[0]
        o = op.panoptic_inference_v2(cls, masks, ts, multi_ar=True)[0]
        assert r["segments_info"] == o["segments_info"]
        assert all(torch.equal(a, b) for a, b in zip(r["pan"], o["pan"])) and all(torch.equal(a, b) for a, b in zip(r["conf"], o["conf"]))
    # nothing above the class threshold: empty result, void confidence everywhere (postprocess.py:70-72)
    cls = torch.full((1, 5, 3), -8.0)
    o = op.panoptic_inference_v2(cls, [torch.randn(1, 5, 4, 6)], np.array([[8, 12]]), multi_ar=True)[0]
    r = pp.panoptic_inference_v2(cls.clone(), [torch.randn(1, 5, 4, 6)], np.array([[8, 12]]), label_mode="sigmoid", device="cpu", multi_ar=True)[0]
    assert o["segments_info"] == r["segments_info"] == [] and int(o["pan"][0].abs().max()) == 0
    assert torch.equal(o["conf"][0], r["conf"][0])


def _decidable(cls, masks, ts, mask_threshold, eps=1e-5):
    """Per view: pixels whose argmax / threshold decisions are separated by more than eps in the oracle."""
    scores, _ = op.class_scores(cls[0])
    keep = scores > 0.1
    out = []
    for m, size in zip(masks, ts):
        up = op.upsampled_probabilities(m, size)[0][keep]
        w = scores[keep].view(-1, 1, 1) * up
        top2 = w.topk(min(2, w.shape[0]), dim=0).values
        margin = top2[0] - top2[1] if w.shape[0] > 1 else torch.ones_like(top2[0])
        thr_gap = (up - mask_threshold).abs().amin(0)
        out.append((margin > eps) & (thr_gap > eps))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(CASES)))
@pytest.mark.parametrize("which", ["v2", "v1"])
def test_cuda_postprocess_matches_reference_golden(ci, which):
    from panst3r_b200 import postprocess as pp
    cls, masks, ts = scene_inputs(CASES[ci])
    g = CASES[ci]["out"][which]
    fn = pp.panoptic_inference_v2 if which == "v2" else pp.panoptic_inference_v1
    r = fn(cls.cuda(), [m.cuda() for m in masks], ts, label_mode="sigmoid", device="cpu", multi_ar=True)[0]
    assert r["segments_info"] == g["segments_info"]
    safe = _decidable(cls, masks, ts, PARAMS[which].get("mask_threshold", 0.25))
    for a, b, c, d, s in zip(r["pan"], g["pan"], r["conf"], g["conf"], safe):
        assert a.shape == b.shape and a.dtype == torch.int32
        assert s.float().mean() > 0.95
        assert torch.equal(a[s].to(torch.int16), b[s])
        assert (a.to(torch.int16) == b).float().mean() > 0.999
        assert (c[s] - d[s]).abs().max() < 1e-5


@pytest.mark.gpu
def test_cuda_postprocess_stacked_tensor_and_edge_cases():
    from panst3r_b200 import ops, postprocess as pp
    # (B, V, Q, h, w) tensor form (multi_ar=False) equals the per-view list form; non-integer scale factors
    cls, logits = op.synthetic_scene(3, 20, 6, 24, 32, 21)
    o = op.panoptic_inference_v2(cls, logits, (45, 70))[0]
    r = pp.panoptic_inference_v2(cls.cuda(), logits.cuda(), (45, 70))[0]
    assert r["segments_info"] == o["segments_info"] and r["pan"].shape == (3, 45, 70)
    assert (r["pan"].cpu() == o["pan"]).float().mean() > 0.999
    rl = pp.panoptic_inference_v2(cls.cuda(), [logits[0, i].cuda() for i in range(3)], np.array([[45, 70]] * 3), multi_ar=True)[0]
    assert all(torch.equal(rl["pan"][i], r["pan"][i]) for i in range(3))
    # nothing survives the class threshold
    e = pp.panoptic_inference_v2(torch.full((1, 20, 6), -8.0).cuda(), logits.cuda(), (48, 64))[0]
    assert e["segments_info"] == [] and int(e["pan"].abs().max()) == 0 and torch.all(e["conf"] == 0.1)
    # class scores: max / first argmax of the sigmoid, saturated ties included
    lg = torch.randn(50, 33, device="cuda") * 4
    lg[3, 5] = lg[3, 9] = 40.0
    s, l = ops.class_scores(lg)
    rs, rl_ = lg.sigmoid().max(-1)
    assert torch.allclose(s, rs, atol=1e-6) and torch.equal(l.long(), rl_) and int(l[3]) == 5
    with pytest.raises(ops._l.Pst3rError):
        pp.panoptic_inference_v2(cls, logits, (48, 64))  # CPU tensors: no fallback


@pytest.mark.gpu
def test_cuda_postprocess_full_size_properties():
    """16 views x 200 queries at 512x384: every pixel gets the best surviving query (checked against torch on one view),
    counters add up, a second identical call is bit-identical."""
    from panst3r_b200 import ops
    torch.manual_seed(0)
    V, Q, hm, wm = 16, 200, 192, 256
    masks = torch.randn(V, Q, hm, wm, device="cuda") * 3
    scores = torch.rand(Q, device="cuda")
    keep = torch.arange(0, Q, 2, device="cuda", dtype=torch.int32)
    sc = scores[keep.long()].contiguous()
    areas = torch.zeros(2, keep.numel(), device="cuda", dtype=torch.int32)
    ids, win = ops.panoptic_argmax(masks, keep, sc, (384, 512), 0.25, areas[0], areas[1])
    a2 = torch.zeros_like(areas)
    ids2, win2 = ops.panoptic_argmax(masks, keep, sc, (384, 512), 0.25, a2[0], a2[1])
    assert torch.equal(ids, ids2) and torch.equal(win, win2) and torch.equal(areas, a2)
    up = torch.nn.functional.interpolate(masks[5:6, ::2].sigmoid(), size=(384, 512), mode="bilinear", align_corners=False)[0]
    w = sc.view(-1, 1, 1) * up
    top2 = w.topk(2, dim=0)
    safe = (top2.values[0] - top2.values[1]) > 1e-5
    assert safe.float().mean() > 0.99 and torch.equal(ids[5][safe].long(), top2.indices[0][safe])
    assert (win[5] - up.gather(0, ids[5][None].long())[0]).abs().max() < 1e-5
    assert int(areas[1].sum()) == int((win >= 0.25).sum())
    assert abs(int(areas[0].sum()) - int((torch.nn.functional.interpolate(masks[:, ::2].sigmoid(), size=(384, 512), mode="bilinear",
                                                                         align_corners=False) >= 0.5).sum())) < 200
