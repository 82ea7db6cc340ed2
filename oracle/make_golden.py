"""Generate golden vectors FROM THE REFERENCE'S OWN MODULES (run in the build container, where /root/reference exists).

    python -m oracle.make_golden

The reference's panoptic-head classes are imported from /root/reference/src (oracle/ref_import.py), loaded with the
deterministic key-hashed weights of oracle/weights.py, run on seeded inputs, and the outputs are stored under
tests/golden/.  Tests then check (a) the oracle restatement and (b) the CUDA path against these files without needing
the reference checkout (the GPU box has none).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os

import torch

from . import ref_import, weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CLASSES = [f"c{i}" for i in range(7)]


def head_inputs(V, H, Wd, seed, portrait=False):
    hs, ws = H // 16, Wd // 16
    N = hs * ws
    g = torch.Generator().manual_seed(seed)
    feats = tuple(torch.randn(1, V, N, d, generator=g) for d in (1024, 768, 1024))
    imgs = torch.rand(1, V, 3, H, Wd, generator=g) * 2 - 1
    ys, xs = torch.meshgrid(torch.arange(hs), torch.arange(ws), indexing="ij")
    pos = torch.stack([ys.flatten(), xs.flatten()], -1)[None, None].expand(1, V, -1, -1).contiguous()
    ts = torch.tensor([[[Wd, H] if portrait else [H, Wd]] * V])
    return feats, imgs, pos, ts


def build_ref_head(ref, variant):
    if variant == "v1":
        m = ref.PanopticDecoder(upscaler=ref.PixelShuffleUpscaler(input_dim=2816), text_encoder="siglip", fixed_vocab=True)
    else:
        m = ref.PanopticDecoder(input_mixer=ref.InputMixer([512, 512], 16, 2816, 768),
                                upscaler=ref.LoftUpUpscaler(input_dim=768, dim=384), mask_dim=384,
                                text_encoder="siglip", fixed_vocab=True)
    m.eval()
    m.load_state_dict(W.synth_state_dict(m, seed=1))
    m.text_encoder.class_embeddings = W.synth_class_embeddings(CLASSES)
    return m


def main():
    ref = ref_import.load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    cases = [("v1", 2, 32, 48, False), ("v2", 2, 32, 48, False), ("v1", 3, 64, 96, False), ("v1", 2, 32, 48, True)]
    for variant, V, H, Wd, portrait in cases:
        m = build_ref_head(ref, variant)
        feats, imgs, pos, ts = head_inputs(V, H, Wd, seed=5, portrait=portrait)
        with torch.no_grad():
            out = m(feats, imgs, pos, ts, CLASSES)
            mq = m(feats, imgs, pos, ts, CLASSES, memory_queries=out["out_queries"])
            x = torch.cat(feats, -1).flatten(0, 1)
            if m.input_mixer is not None and not callable(getattr(m.input_mixer, "__name__", None)):
                try:
                    x = m.input_mixer(x, pos.flatten(0, 1))
                except TypeError:
                    pass
            fpn, mask_f = m._upscaler_wrapper((x, imgs.flatten(0, 1)), ts.flatten(0, 1))
        blob = {
            "variant": variant, "V": V, "H": H, "W": Wd, "portrait": portrait, "classes": CLASSES, "weight_seed": 1,
            "input_seed": 5,
            "pred_logits": out["pred_logits"], "pred_masks": out["pred_masks"].half(), "out_queries": out["out_queries"],
            "aux0_masks": out["aux_outputs"][0]["pred_masks"].half(), "aux0_logits": out["aux_outputs"][0]["pred_logits"],
            "memq_masks_equal_full": bool(torch.equal(mq["pred_masks"], out["pred_masks"])),
            "fpn0": fpn[0].half(), "mask_feats": mask_f.half(),
        }
        name = f"head_{variant}_V{V}_{H}x{Wd}{'_portrait' if portrait else ''}.pt"
        torch.save(blob, os.path.join(GOLDEN, name))
        print("wrote", name, {k: tuple(v.shape) for k, v in blob.items() if torch.is_tensor(v)})
    # post-processing front half (argmax ids) on the v1 case: defines "bit-exact argmax instance ids"
    m = build_ref_head(ref, "v1")
    feats, imgs, pos, ts = head_inputs(2, 32, 48, seed=5)
    with torch.no_grad():
        out = m(feats, imgs, pos, ts, CLASSES)
        scores = out["pred_logits"].sigmoid().max(-1).values[0]  # (Q,)
        masks = out["pred_masks"][0].sigmoid()  # (V, Q, h, w)
        up = torch.nn.functional.interpolate(masks, size=(32, 48), mode="bilinear", align_corners=False)
        ids = (scores[None, :, None, None] * up).argmax(1)
        top2 = (scores[None, :, None, None] * up).topk(2, dim=1).values
    torch.save({"ids": ids.to(torch.int16), "margin": (top2[:, 0] - top2[:, 1]).half()}, os.path.join(GOLDEN, "argmax_v1_V2_32x48.pt"))
    print("wrote argmax_v1_V2_32x48.pt")


if __name__ == "__main__":
    main()
