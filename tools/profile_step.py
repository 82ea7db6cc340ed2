"""Profiling driver for ncu (never a bench number).  Usage under gpurun:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py step
  ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/prof \
      python tools/profile_step.py kernels
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from panst3r_b200 import ops  # noqa: E402
from panst3r_b200.panst3r import build_panst3r  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "step"
    V = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    cudart = torch.cuda.cudart()
    if mode == "step":
        with torch.device("cuda"):
            m = build_panst3r("v1")
        bench.init_weights_(m)
        g = torch.Generator().manual_seed(7)
        m.panoptic_decoder.text_encoder.class_embeddings = {c: torch.randn(768, generator=g) for c in bench.CLASSES}
        m.overlap_dino = False
        m.panoptic_decoder.mask_transformer.overlap_aux_masks = False
        imgs, ts = bench.make_inputs(V, "cuda")
        imgs = imgs.cuda()
        for _ in range(2):
            m(imgs, ts, bench.CLASSES)
        torch.cuda.synchronize()
        cudart.cudaProfilerStart()
        m(imgs, ts, bench.CLASSES)
        torch.cuda.synchronize()
        cudart.cudaProfilerStop()
    else:
        r = lambda *s: torch.randn(*s, device="cuda").bfloat16()  # noqa: E731
        q, k, v = r(V, 768, 12, 64), r(1, V * 768, 12, 64), r(1, V * 768, 12, 64)      # render cross-attention
        qs, ks, vs = r(V, 768, 16, 64), r(V, 768, 16, 64), r(V, 768, 16, 64)           # encoder self-attention
        a, w = r(V * 768, 1024), r(4096, 1024)                                         # ViT-L fc1
        o = torch.empty(V * 768, 4096, device="cuda", dtype=torch.bfloat16)
        a2, w2 = r(768, 768), r(2304, 768)                                             # memory-build QKV (latency bound)
        feats, emb = r(V * 192 * 256, 256), r(200, 256)                                # mask einsum
        mo = torch.empty(V, 200, 192, 256, device="cuda", dtype=torch.float32)
        x = r(V * 768, 1024)
        gam, bet = torch.ones(1024, device="cuda"), torch.zeros(1024, device="cuda")
        # reference-precision head (split-bf16 operands): upscaler fc2 with accumulator promotion, mask einsum with TMA stores
        ah, wh = ops.Split.from_float(torch.randn(V * 768, 11264, device="cuda")), ops.Split.from_float(torch.randn(2048, 11264, device="cuda") * 0.01)
        oh = ops.Split.empty((V * 4 * 768, 512), "cuda")
        fs, es = ops.Split.from_float(feats.float()), ops.Split.from_float(emb.float())
        a3, w3, res3 = r(768, 3072), r(768, 3072), r(768, 768)                         # memory-build fc2: split-K cluster kernel / unsplit

        def run():
            ops.attention(q, k, v)
            ops.attention(qs, ks, vs)
            ops.gemm(a, w, out=o, act=ops.ACT_GELU, bias=torch.zeros(4096, device="cuda"))
            ops.gemm(a2, w2)
            ops.gemm(feats, emb, out=mo, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=192 * 256, batch_stride=200 * 192 * 256, ldt=192 * 256)
            ops.layernorm(x, gam, bet, 1e-6)
            ops.gemm(ah, wh, out=oh, store_mode=ops.STORE_PIXSHUF2, grid=(24 * V, 32))
            ops.gemm(fs, es, out=mo, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=192 * 256, batch_stride=200 * 192 * 256, ldt=192 * 256)
            prev = ops.set_split_k(True)
            ops.gemm(a3, w3, residual=res3)
            ops.set_split_k(prev)
            ops.gemm(a3, w3, residual=res3)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        cudart.cudaProfilerStart()
        run()
        torch.cuda.synchronize()
        cudart.cudaProfilerStop()


if __name__ == "__main__":
    main()
