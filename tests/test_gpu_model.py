"""GPU parity of the CUDA modules / façade against the CPU oracle and the reference-generated golden vectors.

Tolerances (max |a-b| / max |b| per tensor):
  * single modules fed identical bf16-representable weights: 2e-2 — bf16 activations (2^-8 per rounding) accumulated
    over the stack; the fp32 oracle keeps fp32 activations;
  * first prediction head (no attention-mask feedback yet): 2e-2;
  * later prediction heads with the oracle's block masks forced: 3e-2; free-running they additionally depend on
    sign(mask logit) decisions of near-zero logits (ill-conditioned by construction) -> looser bound + a
    bound on the fraction of flipped mask bits;
  * argmax ids: exact wherever the golden top-2 margin exceeds the logit tolerance.
"""
import os

import pytest
import torch

from helpers import CLASSES, GOLDEN, bf16_weights, build_oracle_head, golden_files, head_inputs, relmax
from oracle import weights as W
from oracle.panst3r import build_panst3r as build_oracle

pytestmark = pytest.mark.gpu

TOL = 2e-2


@pytest.fixture(scope="module")
def pair():
    from panst3r_b200.panst3r import build_panst3r
    torch.manual_seed(0)
    depth = (2, 2, 2)
    o = build_oracle("v1", *depth)
    sd = bf16_weights(W.synth_state_dict(o, seed=3))
    o.load_state_dict(sd)
    m = build_panst3r("v1", *depth)
    m.load_state_dict(sd, strict=True)  # identical state-dict surface
    m = m.cuda()
    classes = [f"c{i}" for i in range(9)]
    ce = W.synth_class_embeddings(classes)
    o.panoptic_decoder.text_encoder.class_embeddings = ce
    m.panoptic_decoder.text_encoder.class_embeddings = ce
    V, H, Wd = 3, 64, 96
    g = torch.Generator().manual_seed(1)
    imgs = torch.rand(1, V, 3, H, Wd, generator=g) * 2 - 1
    ts = torch.tensor([[[H, Wd]] * V])
    return o, m, imgs, ts, classes


def test_encoder_dino_parity(pair):
    o, m, imgs, ts, _ = pair
    xo, poso = o.forward_must3r_encoder(imgs, ts)
    xm, posm = m.forward_must3r_encoder(imgs.cuda(), ts)
    assert xm.dtype == torch.bfloat16 and torch.equal(posm.cpu(), poso)
    assert relmax(xm, xo) < TOL
    assert relmax(m.forward_dino(imgs.cuda(), ts), o.forward_dino(imgs, ts)) < TOL
    # the same stages with every LayerNorm folded into the consuming GEMM (off by default for per-view stages)
    m.must3r_encoder.fold_ln = m.dino_encoder.fold_ln = True
    try:
        xf, _ = m.forward_must3r_encoder(imgs.cuda(), ts)
        assert relmax(xf, xo) < TOL and relmax(xf, xm) < TOL
        assert relmax(m.forward_dino(imgs.cuda(), ts), o.forward_dino(imgs, ts)) < TOL
    finally:
        m.must3r_encoder.fold_ln = m.dino_encoder.fold_ln = False


def test_decoder_memory_and_render_parity(pair):
    o, m, imgs, ts, _ = pair
    xo, poso = o.forward_must3r_encoder(imgs, ts)
    xc, pc = xo.cuda().bfloat16(), poso.cuda()
    memo = o.build_memory(xo, poso, ts)
    memm = m.build_memory(xc, pc, ts)
    assert memm[2] == memo[2] == imgs.shape[1] and torch.equal(memm[1].cpu(), memo[1])
    for a, b in zip(memm[0], memo[0]):
        assert a.shape == b.shape and relmax(a, b) < TOL
    pmo, yo = o.render(xo, poso, ts, memo)
    mem2, pmm, fm = m.must3r_decoder(xc, pc, ts, memm, render=True, return_feats=True)
    assert mem2 is memm and len(fm) == 3
    assert relmax(fm[-1], yo) < TOL and relmax(pmm, pmo) < TOL
    # render is per-view independent: a sub-batch gives bit-identical rows (chunk invariance, utils.batched_map)
    _, pm1, f1 = m.must3r_decoder(xc[:, 1:2], pc[:, 1:2], ts[:, 1:2], memm, render=True, return_feats="last")
    assert torch.equal(pm1, pmm[:, 1:2]) and torch.equal(f1[-1], fm[-1][:, 1:2])
    # the reference's sliced render path hands the decoder a plain (expanded) memory tuple without our bank
    plain = ([t.clone() for t in memm[0]], memm[1], memm[2], None, None)
    _, pm2, _ = m.must3r_decoder(xc[:, 1:2], pc[:, 1:2], ts[:, 1:2], plain, render=True, return_feats="last")
    assert relmax(pm2, pm1) < 1e-6
    # render with folded LayerNorms (the memory build above already runs folded)
    m.must3r_decoder.fold_ln_render = True
    try:
        _, pmf, ff = m.must3r_decoder(xc, pc, ts, memm, render=True, return_feats="last")
        assert relmax(ff[-1], yo) < TOL and relmax(pmf, pmo) < TOL
    finally:
        m.must3r_decoder.fold_ln_render = False


@pytest.mark.parametrize("path", golden_files("head_v1*.pt"))
def test_head_against_reference_golden(path):
    """CUDA PanopticDecoder in its bf16 mode vs outputs of the REFERENCE's own modules (tests/golden,
    oracle/make_golden.py).  The reference-precision mode is held to 1e-3 in tests/test_gpu_precise.py."""
    from panst3r_b200.modules.panoptic import PanopticDecoder, PixelShuffleUpscaler
    g = torch.load(path)
    m = PanopticDecoder(upscaler=PixelShuffleUpscaler(input_dim=2816), precision="bf16").eval()  # fp32 mode: test_gpu_precise.py
    o = build_oracle_head("v1")
    m.load_state_dict(o.state_dict(), strict=True)
    m = m.cuda()
    m.text_encoder.class_embeddings = o.text_encoder.class_embeddings
    feats, imgs, pos, ts = head_inputs(g["V"], g["H"], g["W"], g["input_seed"], portrait=g["portrait"])
    out = m(tuple(f.cuda() for f in feats), imgs.cuda(), pos.cuda(), ts, CLASSES)
    V, H, Wd = g["V"], g["H"], g["W"]
    if g["portrait"]:  # predicted in the true (48 x 32) orientation, stored transposed (utils.transpose_to_landscape)
        f16, f2 = m.upscaler((torch.cat(feats, -1)[0].cuda(), None), (Wd, H))
        f16, f2 = [f16[0].swapaxes(2, 3)], f2.swapaxes(2, 3)
    else:
        f16, f2 = m.upscaler((torch.cat(feats, -1)[0].cuda(), None), (H, Wd))
    assert relmax(f16[0], g["fpn0"]) < TOL and relmax(f2, g["mask_feats"]) < TOL
    assert out["pred_masks"].shape == g["pred_masks"].shape and out["pred_masks"].dtype == torch.float32
    assert relmax(out["aux_outputs"][0]["pred_masks"], g["aux0_masks"]) < TOL
    assert relmax(out["aux_outputs"][0]["pred_logits"], g["aux0_logits"]) < TOL
    assert relmax(out["pred_masks"], g["pred_masks"]) < 0.25
    assert relmax(out["pred_logits"], g["pred_logits"]) < 0.25
    # memory-query path (BASELINE config 3): exactly the final prediction head
    mq = m(tuple(f.cuda() for f in feats), imgs.cuda(), pos.cuda(), ts, CLASSES, memory_queries=out["out_queries"])
    assert torch.equal(mq["pred_masks"], out["pred_masks"]) and set(mq) == {"pred_logits", "pred_masks"}


def _oracle_masks(o, fpn, mask_f, ts, cls):
    rec = []
    orig = o.mask_transformer.forward_prediction_heads

    def hook(*a, **k):
        r = orig(*a, **k)
        if r[2] is not None:
            rec.append(r[2].clone())
        return r
    o.mask_transformer.forward_prediction_heads = hook
    try:
        out = o.mask_transformer(fpn, mask_f, ts, cls)
    finally:
        o.mask_transformer.forward_prediction_heads = orig
    return out, rec


def test_query_decoder_layers_with_forced_masks():
    """Every decoder layer in isolation from threshold flips: the oracle's block masks are forced into the CUDA path."""
    from panst3r_b200.modules.panoptic import PanopticDecoder, PixelShuffleUpscaler
    from test_gpu_kernels import pack_bits
    o = build_oracle_head("v1")
    sd = bf16_weights(o.state_dict())
    o.load_state_dict(sd)
    m = PanopticDecoder(upscaler=PixelShuffleUpscaler(input_dim=2816)).eval()
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    V, H, Wd = 3, 64, 96
    feats, imgs, pos, ts = head_inputs(V, H, Wd, seed=9)
    cat = torch.cat(feats, -1).bfloat16().float()
    with torch.no_grad():
        fpn, mask_f = o.upscaler((cat[0], imgs[0]), (H, Wd))
        cls = torch.stack([o.text_encoder.class_embeddings[c] for c in CLASSES])
        cls = cls / cls.norm(dim=-1, keepdim=True)
        ref, masks = _oracle_masks(o, [fpn[0][None]], mask_f[None], ts, cls)
    assert len(masks) == 7
    forced = []
    for mk in masks[:6]:
        mk = mk.clone()
        mk[torch.where(mk.sum(-1) == mk.shape[-1])] = False
        forced.append(pack_bits(mk[:1].cuda()))  # heads share one mask
    mt = m.mask_transformer
    hs, ws = H // 16, Wd // 16
    src, mf = m.upscaler.forward_nhwc(cat.cuda().bfloat16().view(V * hs * ws, -1), V, hs, ws, f16_extra_bias=mt.level_embed.weight)
    cls_emb = m.text_encoder(CLASSES, device="cuda") if (setattr(m.text_encoder, "class_embeddings", o.text_encoder.class_embeddings) or True) else None
    out = mt.forward_nhwc(src, mf, (hs, ws), cls_emb, mask_override=forced)
    for i, (a, b) in enumerate(zip(out["aux_outputs"], ref["aux_outputs"])):
        assert relmax(a["pred_masks"], b["pred_masks"]) < 3e-2, f"head {i}"
        assert relmax(a["pred_logits"], b["pred_logits"]) < 3e-2, f"head {i}"
    assert relmax(out["pred_masks"], ref["pred_masks"]) < 3e-2
    assert relmax(out["pred_logits"], ref["pred_logits"]) < 3e-2
    assert relmax(out["out_queries"], ref["out_queries"]) < 3e-2
    # free-running: our own masks differ from the oracle's only on near-zero logits
    free = mt.forward_nhwc(src, mf, (hs, ws), cls_emb)
    _, _, bits0 = mt.prediction_heads(mt.query_feat.weight.detach().cuda().bfloat16(), mf,
                                      __import__("panst3r_b200.ops", fromlist=["x"]).center_pool8(mf).view(V * hs * ws, -1), cls_emb, False)
    mine = ((bits0[0].to(torch.int64)[..., None] >> torch.arange(32, device="cuda")) & 1).bool().view(200, -1)[:, :V * hs * ws]
    theirs = masks[0][0].clone()
    theirs[theirs.all(-1)] = False
    flipped = (mine.cpu() != theirs).float().mean().item()
    assert flipped < 0.02, f"{flipped:.4f} of the first block mask differs"
    assert relmax(free["pred_masks"], ref["pred_masks"]) < 0.3


def test_argmax_ids_margin_aware():
    """Instance ids (postprocess.py:18-27,63,77 front half) from CUDA masks equal the reference's wherever the golden
    top-2 margin exceeds the score tolerance."""
    from panst3r_b200.modules.panoptic import PanopticDecoder, PixelShuffleUpscaler
    g = torch.load(os.path.join(GOLDEN, "argmax_v1_V2_32x48.pt"))
    o = build_oracle_head("v1")
    m = PanopticDecoder(upscaler=PixelShuffleUpscaler(input_dim=2816), deep_supervision=False, precision="bf16").eval()
    m.load_state_dict(o.state_dict(), strict=True)
    m = m.cuda()
    m.text_encoder.class_embeddings = o.text_encoder.class_embeddings
    feats, imgs, pos, ts = head_inputs(2, 32, 48, seed=5)
    gold = torch.load(os.path.join(GOLDEN, "head_v1_V2_32x48.pt"))
    full = m(tuple(f.cuda() for f in feats), imgs.cuda(), pos.cuda(), ts, CLASSES)
    assert full["aux_outputs"] == []  # deep_supervision=False keeps only the final prediction
    # final prediction head on the REFERENCE's final queries (the config-3 memory-query path): well conditioned,
    # unlike the free-running 6-layer decoder whose block masks flip on near-zero logits of random weights
    out = m(tuple(f.cuda() for f in feats), imgs.cuda(), pos.cuda(), ts, CLASSES, memory_queries=gold["out_queries"].cuda())
    assert relmax(out["pred_masks"], gold["pred_masks"]) < TOL and relmax(out["pred_logits"], gold["pred_logits"]) < TOL
    scores = out["pred_logits"].sigmoid().max(-1).values[0]
    up = torch.nn.functional.interpolate(out["pred_masks"][0].sigmoid(), size=(32, 48), mode="bilinear", align_corners=False)
    weighted = scores[None, :, None, None] * up
    ids = weighted.argmax(1).cpu()
    gs = gold["pred_logits"].sigmoid().max(-1).values[0]
    gup = torch.nn.functional.interpolate(gold["pred_masks"].float()[0].sigmoid(), size=(32, 48), mode="bilinear", align_corners=False)
    # 200 random-weight queries are near-tied at most pixels; ids must be EXACT wherever the golden top-2 margin
    # exceeds twice the per-pixel score error (anything else is undecidable at any finite precision)
    tol = 2.0 * (weighted.cpu() - gs[None, :, None, None] * gup).abs().amax(dim=1)
    safe = g["margin"].float() > tol
    assert safe.float().mean().item() > 0.02, f"only {safe.float().mean().item():.4f} of the pixels are decidable"
    assert torch.equal(ids[safe], g["ids"].long()[safe])
    assert (ids == g["ids"].long()).float().mean().item() > 0.5


def test_forward_and_multi_ar_vs_oracle(pair):
    o, m, imgs, ts, classes = pair
    pan_o, pm_o = o(imgs, ts, classes)
    pan, pm = m(imgs.cuda(), ts, classes)
    assert set(pan) == {"pred_logits", "pred_masks", "aux_outputs", "out_queries"} and len(pan["aux_outputs"]) == 6
    assert pm.shape == pm_o.shape and pm.dtype == torch.float32 and relmax(pm, pm_o) < TOL
    assert relmax(pan["aux_outputs"][0]["pred_masks"], pan_o["aux_outputs"][0]["pred_masks"]) < TOL
    assert relmax(pan["pred_masks"], pan_o["pred_masks"]) < 0.25
    # determinism: the same inputs give bit-identical outputs
    pan_b, pm_b = m(imgs.cuda(), ts, classes)
    assert torch.equal(pm, pm_b) and torch.equal(pan["pred_masks"], pan_b["pred_masks"])
    # all-keyframe multi_ar == forward (return order differs, panst3r.py:284 vs :296)
    pms, pan2 = m.forward_inference_multi_ar(list(imgs[0].cuda()), ts[0], classes)
    assert torch.equal(torch.stack(pms), pm[0]) and torch.equal(torch.stack(pan2["pred_masks"]), pan["pred_masks"][0])
    # keyframes + render-only frames (BASELINE config 3 path) vs the oracle's restatement of the same path
    pms3, pan3 = m.forward_inference_multi_ar(list(imgs[0].cuda()), ts[0], classes, num_keyframes=2)
    pmso, pano = o.forward_inference_multi_ar(list(imgs[0]), ts[0], classes, num_keyframes=2)
    for a, b in zip(pms3, pmso):
        assert relmax(a, b) < TOL
    assert len(pan3["pred_masks"]) == 3 and pan3["pred_masks"][1].shape == pano["pred_masks"][1].shape
    # outdevice: results land on the host like the reference's demo (tools/demo_panst3r.py:232-233)
    pan_c, pm_c = m(imgs.cuda(), ts, classes, outdevice="cpu")
    assert pm_c.device.type == "cpu" and pan_c["pred_masks"].device.type == "cpu" and torch.equal(pm_c, pm.cpu())


def test_forward_portrait_vs_oracle(pair):
    """All-portrait scene: tensors stored landscape (64 x 96), true_shape (96, 64) — the reference's convention
    (model/dino.py:25-33, utils.py:36-49).  Every stage runs in the true orientation; dense outputs come back in the
    landscape storage layout."""
    o, m, imgs, ts, classes = pair
    tsp = ts.flip(-1)
    pan_o, pm_o = o(imgs, tsp, classes)
    pan, pm = m(imgs.cuda(), tsp, classes)
    assert pm.shape == pm_o.shape == (1, imgs.shape[1], 64, 96, 7) and relmax(pm, pm_o) < TOL
    assert pan["pred_masks"].shape == pan_o["pred_masks"].shape
    assert relmax(pan["aux_outputs"][0]["pred_masks"], pan_o["aux_outputs"][0]["pred_masks"]) < TOL
    assert relmax(pan["aux_outputs"][0]["pred_logits"], pan_o["aux_outputs"][0]["pred_logits"]) < TOL
    # not the landscape result in disguise
    _, pm_l = m(imgs.cuda(), ts, classes)
    assert relmax(pm, pm_l) > 0.05
    assert relmax(m.forward_dino(imgs.cuda(), tsp), o.forward_dino(imgs, tsp)) < TOL


def test_head_multi_ar_against_reference_golden():
    """PanopticDecoder(multi_ar=True) — three stacks (two landscape shapes + one portrait) — vs the REFERENCE's output
    lists (tests/golden/head_v1_multi_ar.pt, oracle/make_golden.py::make_multi_ar_golden)."""
    from oracle.make_golden import multi_ar_head_inputs
    from panst3r_b200.modules.panoptic import PanopticDecoder, PixelShuffleUpscaler
    g = torch.load(os.path.join(GOLDEN, "head_v1_multi_ar.pt"))
    o = build_oracle_head("v1")
    m = PanopticDecoder(upscaler=PixelShuffleUpscaler(input_dim=2816), precision="bf16").eval()
    m.load_state_dict(o.state_dict(), strict=True)
    m = m.cuda()
    m.text_encoder.class_embeddings = o.text_encoder.class_embeddings
    in_feats, imgs, pos, ts = multi_ar_head_inputs()
    cu = lambda lst: [t.cuda() for t in lst]  # noqa: E731
    out = m(tuple(cu(f) for f in in_feats), cu(imgs), cu(pos), ts, CLASSES, multi_ar=True)
    assert len(out["pred_masks"]) == 3 and len(out["aux_outputs"]) == 6
    for a, b, c, d in zip(out["pred_masks"], g["pred_masks"], out["aux_outputs"][0]["pred_masks"], g["aux0_masks"]):
        assert a.shape == b.shape and a.dtype == torch.float32
        assert relmax(c, d) < TOL          # first head: no mask feedback yet
        assert relmax(a, b) < 0.25         # free-running final head (ill-conditioned with random weights, see module doc)
    assert relmax(out["aux_outputs"][0]["pred_logits"], g["aux0_logits"]) < TOL
    # well-conditioned check of the last head on every stack: the reference's final queries through the memory-query path
    mq = m(tuple(cu(f) for f in in_feats), cu(imgs), cu(pos), ts, CLASSES, multi_ar=True, memory_queries=g["out_queries"].cuda())
    for a, b in zip(mq["pred_masks"], g["pred_masks"]):
        assert relmax(a, b) < TOL
    assert relmax(mq["pred_logits"], g["pred_logits"]) < TOL


def test_forward_inference_multi_ar_mixed_shapes_vs_oracle(pair):
    """Views of different shape and orientation in one scene (must3r stack_views, panst3r.py:203-206, 257-263): per-view
    stages run per stack, the memory build walks the keyframes in order, the head sees one stack per shape."""
    o, m, _, _, classes = pair
    g = torch.Generator().manual_seed(11)
    shapes = [(64, 96), (64, 96), (64, 64), (64, 96), (64, 96), (64, 64)]
    ts = torch.tensor([[64, 96], [64, 96], [64, 64], [96, 64], [64, 96], [64, 64]])  # view 3: portrait, stored transposed
    imgs = [torch.rand(3, *sh, generator=g) * 2 - 1 for sh in shapes]
    for kf in (None, 4):  # all keyframes; 4 keyframes (linspace -> views 0, 1, 3, 5) + 2 render-only frames
        pms_o, pan_o = o.forward_inference_multi_ar(imgs, ts, classes, num_keyframes=kf)
        pms, pan = m.forward_inference_multi_ar([im.cuda() for im in imgs], ts, classes, num_keyframes=kf)
        assert len(pms) == len(pan["pred_masks"]) == 6
        for a, b in zip(pms, pms_o):
            assert a.shape == b.shape and relmax(a, b) < TOL
        for a, b in zip(pan["pred_masks"], pan_o["pred_masks"]):
            assert a.shape == b.shape and torch.isfinite(a).all()
        if kf is not None:  # render-only frames decoded with the ORACLE's final queries: well-conditioned comparison
            assert relmax(pan["pred_logits"], pan_o["pred_logits"]) < 0.25
    # same-shape lists keep going through the single-stack path: identical to forward()
    same = [imgs[i].cuda() for i in (0, 1, 4)]
    pms1, pan1 = m.forward_inference_multi_ar(same, ts[[0, 1, 4]], classes)
    pan2, pm2 = m(torch.stack(same)[None], ts[[0, 1, 4]][None], classes)
    assert torch.equal(torch.stack(pms1), pm2[0]) and torch.equal(torch.stack(pan1["pred_masks"]), pan2["pred_masks"][0])
    # the initialisation pair must share a shape
    with pytest.raises(Exception):
        m.forward_inference_multi_ar([imgs[0].cuda(), imgs[2].cuda(), imgs[1].cuda()], ts[[0, 2, 1]], classes)


def test_full_depth_full_resolution_properties():
    """Full-depth model at 512x384 (4 keyframes): finite outputs, reference shapes, render chunk invariance."""
    from panst3r_b200.panst3r import build_panst3r
    import bench
    with torch.device("cuda"):
        m = build_panst3r("v1")
    bench.init_weights_(m)
    g = torch.Generator().manual_seed(7)
    classes = bench.CLASSES[:20]
    m.panoptic_decoder.text_encoder.class_embeddings = {c: torch.randn(768, generator=g) for c in classes}
    V = 4
    imgs, ts = bench.make_inputs(V, "cuda")
    pan, pm = m(imgs.cuda(), ts, classes)
    assert pm.shape == (1, V, 384, 512, 7) and pan["pred_masks"].shape == (1, V, 200, 192, 256)
    assert pan["pred_logits"].shape == (1, 200, 20) and pan["out_queries"].shape == (200, 1, 768)
    for t in (pm, pan["pred_masks"], pan["pred_logits"]):
        assert torch.isfinite(t).all()
    pan2, pm2 = m(imgs.cuda(), ts, classes)
    assert torch.equal(pm, pm2) and torch.equal(pan["pred_masks"], pan2["pred_masks"])


def test_from_checkpoint_roundtrip(tmp_path, pair):
    from panst3r_b200.panst3r import PanSt3R
    o, m, imgs, ts, classes = pair
    args = dict(must3r_encoder="Dust3rEncoder(img_size=[512, 512], patch_embed='PatchEmbedDust3R', depth=2)",
                must3r_decoder="MUSt3R(img_size=[512, 512], feedback_type='single_mlp', memory_mode='norm_y', depth=2)",
                dino_encoder="DinoV2Encoder(depth=2)",
                panoptic_decoder="PanopticDecoder(input_mixer=None, upscaler=PixelShuffleUpscaler(input_dim=2816), label_mode='sigmoid', text_encoder='siglip')")
    p = tmp_path / "ckpt.pth"
    torch.save({"args": args, "weights": {k: v.cpu() for k, v in m.state_dict().items()}}, p)
    m2 = PanSt3R.from_checkpoint(str(p)).cuda()
    m2.panoptic_decoder.text_encoder.class_embeddings = m.panoptic_decoder.text_encoder.class_embeddings
    _, pm = m(imgs.cuda(), ts, classes)
    _, pm2 = m2(imgs.cuda(), ts, classes)
    assert torch.equal(pm, pm2)


def test_trunk_error_within_reference_policy_noise(pair):
    """DINOv2 / encoder / decoder run under bf16 autocast in the reference (panst3r.py:174, 204) — its outputs then differ
    from an fp32 run by ~1e-2.  The CUDA trunk (bf16 operands, fp32 accumulation / statistics) must sit inside that
    policy noise: error vs the fp32 oracle <= 2.5 x the error of the oracle run under the reference's own policy."""
    o, m, imgs, ts, classes = pair
    pan_o, pm_o = o(imgs, ts, classes)
    pan_a, pm_a = o(imgs, ts, classes, amp=True)
    pan, pm = m(imgs.cuda(), ts, classes)
    for name, ours, pol, ref in (("pointmaps", pm, pm_a, pm_o),
                                 ("first-head masks", pan["aux_outputs"][0]["pred_masks"], pan_a["aux_outputs"][0]["pred_masks"],
                                  pan_o["aux_outputs"][0]["pred_masks"])):
        e_ours, e_pol = relmax(ours, ref), relmax(pol, ref)
        print(f"{name}: CUDA {e_ours:.2e}, reference policy {e_pol:.2e}")
        assert e_ours < max(2.5 * e_pol, 5e-3), (name, e_ours, e_pol)


def test_side_stream_split_kv_workspaces_do_not_collide():
    """DINOv2 on the side stream and the encoder / memory build on the main stream may BOTH run split-KV
    attention (a 1-view stack at 512x384: 64 DINO CTAs -> 2 splits, decoder 36-72 CTAs -> 2-4 splits).  Partials live in
    per-stream workspaces: overlapped and serial execution give bit-identical results."""
    from panst3r_b200.panst3r import build_panst3r
    import bench
    with torch.device("cuda"):
        m = build_panst3r("v1", 2, 2, 2)
    bench.init_weights_(m)
    g = torch.Generator().manual_seed(7)
    classes = bench.CLASSES[:5]
    m.panoptic_decoder.text_encoder.class_embeddings = {c: torch.randn(768, generator=g) for c in classes}
    imgs, ts = bench.make_inputs(3, "cuda")
    imgs = imgs.cuda()
    # views 0, 1 share a stack; view 2 has another shape -> a 1-view stack at full resolution
    lst = [imgs[0, 0], imgs[0, 1], imgs[0, 2][:, :, :384].contiguous()]
    tsl = torch.tensor([[384, 512], [384, 512], [384, 384]])
    res = {}
    for overlap in (True, False):
        m.overlap_dino = overlap
        outs = []
        for _ in range(3):
            pms, pan = m.forward_inference_multi_ar(lst, tsl, classes)
            outs.append((torch.cat([p.flatten() for p in pms]), torch.cat([p.flatten() for p in pan["pred_masks"]])))
        torch.cuda.synchronize()
        assert all(torch.equal(outs[0][0], o_[0]) and torch.equal(outs[0][1], o_[1]) for o_ in outs[1:]), "non-deterministic"
        res[overlap] = outs[0]
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])


def test_memory_is_a_value_copy_on_write(pair):
    """`mem` behaves like the reference's value (torch.cat builds new tensors): a second update branching
    from an OLD memory must not overwrite the tokens of the first."""
    o, m, imgs, ts, _ = pair
    xo, poso = o.forward_must3r_encoder(imgs, ts)
    xc, pc = xo.cuda().bfloat16(), poso.cuda()
    mem2, _, _ = m.must3r_decoder(xc[:, :2], pc[:, :2], ts[:, :2], None, render=False, compute_pointmaps=False)
    mem3a, _, _ = m.must3r_decoder(xc[:, 2:3], pc[:, 2:3], ts[:, 2:3], mem2, render=False, compute_pointmaps=False)
    snap = [t.clone() for t in mem3a[0]]
    # branch again from mem2 with ANOTHER view: must leave mem3a intact
    mem3b, _, _ = m.must3r_decoder(xc[:, 1:2], pc[:, 1:2], ts[:, 1:2], mem2, render=False, compute_pointmaps=False)
    for a, b in zip(mem3a[0], snap):
        assert torch.equal(a, b)
    assert not torch.equal(mem3b[0][0][:, -1], mem3a[0][0][:, -1])


def test_from_checkpoint_reports_and_remaps(tmp_path, pair):
    from panst3r_b200.lib import Pst3rError
    from panst3r_b200.panst3r import PanSt3R
    o, m, imgs, ts, classes = pair
    args = dict(must3r_encoder="Dust3rEncoder(img_size=[512, 512], patch_embed='PatchEmbedDust3R', depth=2)",
                must3r_decoder="MUSt3R(img_size=[512, 512], feedback_type='single_mlp', memory_mode='norm_y', depth=2)",
                dino_encoder="DinoV2Encoder(depth=2)",
                panoptic_decoder="PanopticDecoder(input_mixer=None, upscaler=PixelShuffleUpscaler(input_dim=2816), label_mode='sigmoid', text_encoder='siglip')",
                postprocess_default="standard_v1", qubo_enabled=False)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    # croco-style names for the encoder blocks are remapped
    alt = {k.replace("must3r_encoder.blocks_enc.", "must3r_encoder.enc_blocks.").replace("must3r_encoder.norm_enc.", "must3r_encoder.enc_norm."): v
           for k, v in sd.items()}
    p = tmp_path / "alt.pth"
    torch.save({"args": args, "weights": alt}, p)
    m2 = PanSt3R.from_checkpoint(str(p))
    assert m2.postprocess_default == "standard_v1" and m2.qubo_enabled is False
    assert all(not r["missing"] and not r["unexpected"] for r in m2.load_report.values() if isinstance(r, dict))
    # a checkpoint whose decoder names match nothing must not load silently
    bad = {k.replace("must3r_decoder.", "must3r_decoder.zz_"): v for k, v in sd.items()}
    p2 = tmp_path / "bad.pth"
    torch.save({"args": args, "weights": bad}, p2)
    with pytest.raises(Pst3rError):
        PanSt3R.from_checkpoint(str(p2))
    with pytest.warns(UserWarning):
        m3 = PanSt3R.from_checkpoint(str(p2), allow_partial=True)
    assert len(m3.load_report["must3r_decoder"]["missing"]) == m3.load_report["must3r_decoder"]["parameters"]


def test_forward_batch_of_scenes(pair):
    """B > 1 (panst3r.py:286-296 is batch-generic): a batch is a loop over independent scenes."""
    o, m, imgs, ts, classes = pair
    g = torch.Generator().manual_seed(5)
    imgs2 = torch.rand(imgs.shape, generator=g) * 2 - 1
    both_i, both_t = torch.cat([imgs, imgs2], 0).cuda(), torch.cat([ts, ts], 0)
    pan, pm = m(both_i, both_t, classes)
    pan1, pm1 = m(imgs2.cuda(), ts, classes)
    assert pm.shape[0] == 2 and pan["pred_masks"].shape[0] == 2 and pan["out_queries"].shape[1] == 2
    assert torch.equal(pm[1:], pm1) and torch.equal(pan["pred_masks"][1:], pan1["pred_masks"])
