"""In-graph time of every stage of one PanSt3R.forward (development tool; each stage captured in its own CUDA graph,
timed with CUDA events over several replays).  Usage:  python tools/stage_times.py [V] [variant] [dino SM list | head]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from panst3r_b200 import ops  # noqa: E402
from panst3r_b200.panst3r import DEC_DIM, ENC_DIM, build_panst3r  # noqa: E402


def timed_graph(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    n0 = ops.launches
    fn()
    n = ops.launches - n0
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, n


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    variant = sys.argv[2] if len(sys.argv) > 2 else "v1"
    with torch.device("cuda"):
        m = build_panst3r(variant)
    bench.init_weights_(m)
    g = torch.Generator().manual_seed(7)
    m.panoptic_decoder.text_encoder.class_embeddings = {c: torch.randn(768, generator=g) for c in bench.CLASSES}
    m.overlap_dino = False
    imgs, ts = bench.make_inputs(V, "cuda")
    imgs = imgs.cuda()
    cat, rows, x, pos, _ = m._features(imgs, ts)
    out = {}
    if len(sys.argv) > 3 and sys.argv[3] == "head":
        # panoptic head only: auxiliary mask GEMMs on the side stream (MaskTransformer.overlap_aux_masks) vs in line
        m(imgs, ts, bench.CLASSES)  # fills `cat` through the whole trunk
        mt = m.panoptic_decoder.mask_transformer
        for prec in ("fp32", "bf16"):
            m.panoptic_decoder.precision = prec
            mt.overlap_aux_masks = False
            out[f"head_{prec}_in_line"] = timed_graph(lambda: m.panoptic_decoder(None, imgs, pos, ts, bench.CLASSES, cat_feats=cat), 10)
            mt.overlap_aux_masks = True
            for sms in (32, 64, 96):
                mt.aux_mask_sms = sms
                out[f"head_{prec}_aux_side_{sms}sm"] = timed_graph(lambda: m.panoptic_decoder(None, imgs, pos, ts, bench.CLASSES, cat_feats=cat), 10)
            m.panoptic_decoder.lazy_masks = True
            out[f"head_{prec}_lazy_masks"] = timed_graph(lambda: m.panoptic_decoder(None, imgs, pos, ts, bench.CLASSES, cat_feats=cat), 10)
            m.panoptic_decoder.lazy_masks = False
        print(json.dumps({"views": V, "variant": variant, "stages": {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in out.items()}}))
        return
    out["dino"] = timed_graph(lambda: m.forward_dino(imgs, ts, out=rows[:, ENC_DIM + DEC_DIM:]))
    out["encoder"] = timed_graph(lambda: m.forward_must3r_encoder(imgs, ts, out=rows[:, :ENC_DIM]))
    out["memory_build"] = timed_graph(lambda: m.build_memory(x, pos, ts))
    with ops.sm_budget(84):
        out["memory_build_84sm"] = timed_graph(lambda: m.build_memory(x, pos, ts))
    with ops.sm_budget(64):
        out["dino_64sm"] = timed_graph(lambda: m.forward_dino(imgs, ts, out=rows[:, ENC_DIM + DEC_DIM:]))
    mem = m.build_memory(x, pos, ts)
    out["render"] = timed_graph(lambda: m.must3r_decoder(x, pos, ts, mem, render=True, return_feats="last",
                                                         feats_out=rows[:, ENC_DIM:ENC_DIM + DEC_DIM]))
    out["panoptic_head"] = timed_graph(lambda: m.panoptic_decoder(None, imgs, pos, ts, bench.CLASSES, cat_feats=cat))
    out["forward_serial"] = timed_graph(lambda: m(imgs, ts, bench.CLASSES))
    m.overlap_dino = True
    for sms in [int(a) for a in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "48", "64", "80"])]:
        m.dino_sms = sms
        out[f"forward_dino_{sms}sm"] = timed_graph(lambda: m(imgs, ts, bench.CLASSES))
    res = {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in out.items()}
    print(json.dumps({"views": V, "variant": variant, "stages": res}))


if __name__ == "__main__":
    main()
