"""Shared test helpers (CPU oracle side)."""
import glob
import os

import torch

from oracle import panoptic as op
from oracle import weights as W
from oracle.make_golden import CLASSES, GOLDEN, head_inputs  # noqa: F401


def golden_files(pattern="head_*.pt"):
    """Single-stack head fixtures (the multi aspect-ratio fixture has its own tests)."""
    return sorted(f for f in glob.glob(os.path.join(GOLDEN, pattern)) if "multi_ar" not in f and "conditioned" not in f)


def build_oracle_head(variant, cls_logit_scale=None):
    if variant == "v1":
        m = op.PanopticDecoder(upscaler=op.PixelShuffleUpscaler(input_dim=2816))
    else:
        m = op.PanopticDecoder(input_mixer=op.InputMixer([512, 512], 16, 2816, 768),
                               upscaler=op.LoftUpUpscaler(input_dim=768, dim=384), mask_dim=384)
    m.eval()
    m.load_state_dict(W.synth_state_dict(m, seed=1))
    m.text_encoder.class_embeddings = W.synth_class_embeddings(CLASSES)
    if cls_logit_scale is not None:
        with torch.no_grad():
            m.mask_transformer.cls_logit_scale.fill_(cls_logit_scale)
    return m


def relmax(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def bf16_weights(sd):
    """Round every floating tensor to a bf16-representable value (what the tensor cores consume)."""
    return {k: (v.to(torch.bfloat16).float() if v.is_floating_point() else v) for k, v in sd.items()}
