"""Oracle pinning: the restated panoptic head reproduces the golden vectors generated from the reference's own
modules (oracle/make_golden.py), and — when the reference checkout is present — the reference modules directly."""
import pytest
import torch

from helpers import CLASSES, build_oracle_head, golden_files, head_inputs, relmax
from oracle import ref_import
from oracle import weights as W


@pytest.mark.parametrize("path", golden_files())
def test_oracle_head_matches_golden(path):
    g = torch.load(path)
    m = build_oracle_head(g["variant"])
    feats, imgs, pos, ts = head_inputs(g["V"], g["H"], g["W"], g["input_seed"], g["portrait"])
    with torch.no_grad():
        out = m(feats, imgs, pos, ts, CLASSES)
    # fixtures are fp32 (the larger one keeps its first aux head in fp16); the oracle is bit-exact against the reference
    # modules here, 1e-6 leaves room for a different CPU / BLAS build on the GPU box
    assert relmax(out["pred_logits"], g["pred_logits"]) < 1e-5
    assert relmax(out["out_queries"], g["out_queries"]) < 1e-5
    assert relmax(out["pred_masks"], g["pred_masks"]) < 1e-5
    assert relmax(out["aux_outputs"][0]["pred_masks"], g["aux0_masks"]) < (1e-3 if g["aux0_masks"].dtype == torch.float16 else 1e-5)
    assert g["memq_masks_equal_full"]
    with torch.no_grad():
        mq = m(feats, imgs, pos, ts, CLASSES, memory_queries=out["out_queries"])
    assert torch.equal(mq["pred_masks"], out["pred_masks"])  # config-3 reuse path == final prediction head only


def test_oracle_conditioned_golden():
    """The well-conditioned fixture (class logits O(1)): oracle == reference outputs, argmax ids of the post-processing
    front half, > 90 % of the pixels decided by a top-2 margin above 1e-4."""
    import os
    from helpers import GOLDEN
    g = torch.load(os.path.join(GOLDEN, "head_v1_conditioned.pt"))
    m = build_oracle_head(g["variant"], cls_logit_scale=g["cls_logit_scale"])
    feats, imgs, pos, ts = head_inputs(g["V"], g["H"], g["W"], g["input_seed"])
    with torch.no_grad():
        out = m(feats, imgs, pos, ts, CLASSES)
    assert relmax(out["pred_masks"], g["pred_masks"]) < 1e-5 and relmax(out["pred_logits"], g["pred_logits"]) < 1e-5
    assert g["pred_logits"].abs().max() > 1.0 and g["pred_masks"].abs().max() > 1.0
    assert (g["margin"] > 1e-4).float().mean() > 0.9
    scores = out["pred_logits"].sigmoid().max(-1).values[0]
    up = torch.nn.functional.interpolate(out["pred_masks"][0].sigmoid(), size=(g["H"], g["W"]), mode="bilinear", align_corners=False)
    ids = (scores[None, :, None, None] * up).argmax(1)
    safe = g["margin"] > 1e-6
    assert torch.equal(ids[safe].to(torch.int16), g["ids"][safe])


def test_oracle_multi_ar_head_matches_golden():
    """multi_ar=True (three stacks: two landscape shapes + one portrait) vs the reference's output lists."""
    import os
    from helpers import GOLDEN
    from oracle.make_golden import multi_ar_head_inputs
    g = torch.load(os.path.join(GOLDEN, "head_v1_multi_ar.pt"))
    m = build_oracle_head("v1")
    args = multi_ar_head_inputs()
    with torch.no_grad():
        out = m(*args, CLASSES, multi_ar=True)
        mq = m(*args, CLASSES, multi_ar=True, memory_queries=out["out_queries"])
    assert relmax(out["pred_logits"], g["pred_logits"]) < 1e-5 and relmax(out["out_queries"], g["out_queries"]) < 1e-5
    assert len(out["pred_masks"]) == len(g["pred_masks"]) == 3
    for a, b, c, d in zip(out["pred_masks"], g["pred_masks"], out["aux_outputs"][0]["pred_masks"], g["aux0_masks"]):
        assert a.shape == b.shape and relmax(a, b) < 1e-5 and relmax(c, d) < 1e-5
    assert g["memq_masks_equal_full"] and all(torch.equal(a, b) for a, b in zip(mq["pred_masks"], out["pred_masks"]))


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present (GPU box)")
def test_oracle_multi_ar_head_equals_reference_modules():
    from oracle.make_golden import build_ref_head, multi_ar_head_inputs
    ref = ref_import.load_reference()
    args = multi_ar_head_inputs()
    for variant in ("v1", "v2"):
        r, o = build_ref_head(ref, variant), build_oracle_head(variant)
        with torch.no_grad():
            ro = r(*args, CLASSES, multi_ar=True, outdevice="cpu")
            oo = o(*args, CLASSES, multi_ar=True)
        assert torch.equal(ro["pred_logits"], oo["pred_logits"]) and torch.equal(ro["out_queries"], oo["out_queries"])
        for a, b in zip(ro["pred_masks"], oo["pred_masks"]):
            assert torch.equal(a, b)  # bit-exact, portrait stack included


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("variant,portrait", [("v1", False), ("v2", False), ("v1", True), ("v2", True)])
def test_oracle_head_equals_reference_modules(variant, portrait):
    from oracle.make_golden import build_ref_head
    ref = ref_import.load_reference()
    r = build_ref_head(ref, variant)
    o = build_oracle_head(variant)
    assert set(r.state_dict().keys()) == set(o.state_dict().keys())
    feats, imgs, pos, ts = head_inputs(2, 48, 64, seed=11, portrait=portrait)
    with torch.no_grad():
        ro = r(feats, imgs, pos, ts, CLASSES)
        oo = o(feats, imgs, pos, ts, CLASSES)
    tol = 0.0 if not portrait else 1e-5  # landscape is bit-exact; portrait differs by memory-order of the same sums
    for k in ("pred_logits", "pred_masks", "out_queries"):
        assert relmax(oo[k], ro[k]) <= tol, k
    for a, b in zip(oo["aux_outputs"], ro["aux_outputs"]):
        assert relmax(a["pred_masks"], b["pred_masks"]) <= tol


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present (GPU box)")
def test_argmax_ids_match_reference_postprocess_front_half():
    """engine/postprocess.py:18-27,63,77: sigmoid -> bilinear x2 -> score-weighted argmax over queries."""
    import os
    from helpers import GOLDEN
    g = torch.load(os.path.join(GOLDEN, "argmax_v1_V2_32x48.pt"))
    m = build_oracle_head("v1")
    feats, imgs, pos, ts = head_inputs(2, 32, 48, seed=5)
    with torch.no_grad():
        out = m(feats, imgs, pos, ts, CLASSES)
    scores = out["pred_logits"].sigmoid().max(-1).values[0]
    up = torch.nn.functional.interpolate(out["pred_masks"][0].sigmoid(), size=(32, 48), mode="bilinear", align_corners=False)
    ids = (scores[None, :, None, None] * up).argmax(1)
    assert torch.equal(ids.to(torch.int16), g["ids"])
