"""Import the REFERENCE's own panoptic-head modules from /root/reference/src (this container only).

TEST INFRASTRUCTURE ONLY.  The reference's package __init__ imports the un-vendored `must3r`, so the package
is registered as a path-only stub and the two upstream packages it needs (`croco.models.blocks`,
`must3r.model.blocks.pos_embed`) are served by the oracle restatement in oracle/blocks.py.  Nothing is copied:
the reference sources are executed where they lie.  Used by oracle/make_golden.py and
tests/test_oracle_vs_reference.py (skipped when /root/reference is absent, e.g. on the GPU box).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "panst3r"))


def _stub(name: str, path: str | None = None) -> types.ModuleType:
    m = types.ModuleType(name)
    if path is not None:
        m.__path__ = [path]
    sys.modules[name] = m
    return m


def load_reference():
    """Returns a namespace with the reference's PanopticDecoder, MaskTransformer, PixelShuffleUpscaler,
    LoftUpUpscaler, InputMixer, CrossonlyDecoderBlock, TextEncoder, batched_map, postprocess module."""
    if not available():
        raise RuntimeError("/root/reference is not present")
    from . import blocks as ob

    if "panst3r" not in sys.modules or not hasattr(sys.modules["panst3r"], "_oracle_stub"):
        pk = _stub("panst3r", os.path.join(REF_SRC, "panst3r"))
        pk._oracle_stub = True
        _stub("panst3r.model", os.path.join(REF_SRC, "panst3r", "model"))
        _stub("panst3r.model.upscalers", os.path.join(REF_SRC, "panst3r", "model", "upscalers"))
        _stub("panst3r.engine", os.path.join(REF_SRC, "panst3r", "engine"))
        # upstream packages -> oracle restatement
        _stub("croco")
        _stub("croco.models")
        cb = _stub("croco.models.blocks")
        for n in ("Mlp", "Attention", "Block", "CrossAttention", "DropPath"):
            setattr(cb, n, getattr(ob, n))
        _stub("must3r")
        _stub("must3r.model")
        _stub("must3r.model.blocks")
        pe = _stub("must3r.model.blocks.pos_embed")
        pe.get_pos_embed = ob.get_pos_embed
        if "torchvision" not in sys.modules:
            try:
                import torchvision  # noqa: F401
            except Exception:  # model/dino.py imports torchvision.transforms at module level
                tv = _stub("torchvision")
                tvt = _stub("torchvision.transforms")
                tv.transforms = tvt

    ns = types.SimpleNamespace()
    ns.utils = importlib.import_module("panst3r.utils")
    ns.mask_transformer = importlib.import_module("panst3r.model.mask_transformer")
    ns.text_encoder = importlib.import_module("panst3r.model.text_encoder")
    ns.blocks = importlib.import_module("panst3r.model.blocks")
    ns.input_mixer = importlib.import_module("panst3r.model.input_mixer")
    ns.pixel_shuffle = importlib.import_module("panst3r.model.upscalers.pixel_shuffle")
    ns.loftup = importlib.import_module("panst3r.model.upscalers.loftup")
    ns.panoptic_decoder = importlib.import_module("panst3r.model.panoptic_decoder")
    ns.postprocess = importlib.import_module("panst3r.engine.postprocess")
    ns.PanopticDecoder = ns.panoptic_decoder.PanopticDecoder
    ns.MaskTransformer = ns.mask_transformer.MaskTransformer
    ns.PixelShuffleUpscaler = ns.pixel_shuffle.PixelShuffleUpscaler
    ns.LoftUpUpscaler = ns.loftup.LoftUpUpscaler
    ns.InputMixer = ns.input_mixer.InputMixer
    ns.CrossonlyDecoderBlock = ns.blocks.CrossonlyDecoderBlock
    ns.TextEncoder = ns.text_encoder.TextEncoder
    return ns
