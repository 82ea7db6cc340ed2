"""End-to-end bring-up check: CUDA PanSt3R vs the CPU oracle on a small scene (development tool)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.panst3r import build_panst3r as build_oracle
from oracle import weights as W
from panst3r_b200.panst3r import build_panst3r

def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()

def main(variant="v1", depth=(2, 2, 2), V=3, H=64, Wd=96):
    torch.manual_seed(0)
    o = build_oracle(variant, *depth)
    sd = W.synth_state_dict(o, seed=3)
    sd = {k: (v.to(torch.bfloat16).float() if v.is_floating_point() else v) for k, v in sd.items()}  # bf16-representable weights
    o.load_state_dict(sd)
    m = build_panst3r(variant, *depth)
    missing = m.load_state_dict(sd, strict=False)
    print("missing", missing.missing_keys[:5], "unexpected", missing.unexpected_keys[:5])
    m = m.cuda()
    classes = [f"c{i}" for i in range(9)]
    ce = W.synth_class_embeddings(classes)
    o.panoptic_decoder.text_encoder.class_embeddings = ce
    m.panoptic_decoder.text_encoder.class_embeddings = ce
    g = torch.Generator().manual_seed(1)
    imgs = torch.rand(1, V, 3, H, Wd, generator=g) * 2 - 1
    ts = torch.tensor([[[H, Wd]] * V])
    t = time.time(); po, pmo = o(imgs, ts, classes); print("oracle fwd %.2fs" % (time.time() - t))
    # stage-wise
    xo, poso = o.forward_must3r_encoder(imgs, ts)
    do = o.forward_dino(imgs, ts)
    ic = imgs.cuda()
    xm, posm = m.forward_must3r_encoder(ic, ts)
    dm = m.forward_dino(ic, ts)
    print("encoder rel", rel(xm, xo), "dino rel", rel(dm, do))
    memo = o.build_memory(xo, poso, ts)
    memm = m.build_memory(xm, posm, ts)
    for l in range(len(memo[0])):
        print(" mem layer", l, rel(memm[0][l], memo[0][l]))
    pmr, yo = o.render(xo, poso, ts, memo)
    _, pmm, fm = m.must3r_decoder(xm, posm, ts, memm, render=True, return_feats="last")
    print("render feats rel", rel(fm[-1], yo), "pointmaps rel", rel(pmm, pmr))
    torch.cuda.synchronize()
    t = time.time(); pm_, pmm2 = m(ic, ts, classes); torch.cuda.synchronize(); print("cuda fwd %.3fs" % (time.time() - t))
    print("FULL pointmaps rel", rel(pmm2, pmo))
    for k in ("pred_logits", "pred_masks", "out_queries"):
        print("FULL", k, tuple(pm_[k].shape), "rel", rel(pm_[k], po[k]))
    for i, (a, b) in enumerate(zip(pm_["aux_outputs"], po["aux_outputs"])):
        print("  aux", i, rel(a["pred_masks"], b["pred_masks"]), rel(a["pred_logits"], b["pred_logits"]))
    # head alone on oracle features (isolates the panoptic head)
    cat = torch.cat([xo, yo, do], -1).cuda().to(torch.bfloat16)
    ph = m.panoptic_decoder(None, ic, posm, ts, classes, cat_feats=cat)
    for k in ("pred_logits", "pred_masks", "out_queries"):
        print("HEAD-ONLY", k, "rel", rel(ph[k], po[k]))
    # multi_ar path
    pms, pan2 = m.forward_inference_multi_ar(list(ic[0]), ts[0], classes, num_keyframes=2)
    pmso, pan2o = o.forward_inference_multi_ar(list(imgs[0]), ts[0], classes, num_keyframes=2)
    print("multi_ar pointmaps", [rel(a, b) for a, b in zip(pms, pmso)], "masks", [rel(a, b) for a, b in zip(pan2["pred_masks"], pan2o["pred_masks"])])

if __name__ == "__main__":
    main()
