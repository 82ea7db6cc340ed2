"""CPU oracle for the PanSt3R forward path — TEST INFRASTRUCTURE ONLY.

Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it, and only as the checker or the CPU baseline.

Pinning status:
  * panoptic head (PanopticDecoder, MaskTransformer, PixelShuffle / LoftUp upscalers, InputMixer,
    CrossonlyDecoderBlock, TextEncoder fixed-vocab): PINNED — oracle/panoptic.py is checked against the
    reference's own modules imported from /root/reference/src (oracle/ref_import.py) and against the
    committed fixtures tests/golden/*.pt generated from them by oracle/make_golden.py.
  * MUSt3R encoder / decoder / RoPE2D / pointmap head (oracle/must3r.py): PARITY UNPINNED — the upstream
    packages must3r/dust3r/croco are not vendored in the reference checkout nor installable here; the
    restatement follows their published architecture and is anchored only on the reference's call sites
    (engine/must3r.py:17-24,45,93) plus closed-form known-answer tests.
  * DINOv2: HuggingFace transformers.Dinov2Model (third-party, same version on both boxes).
"""
