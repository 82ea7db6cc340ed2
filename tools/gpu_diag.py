"""Kernel bring-up diagnostics: runs every C-ABI kernel against a torch fp32 reference of the same op and
prints one line per case (never stops at the first failure).  Used under gpurun during development;
the judged parity tests live in tests/.
"""
from __future__ import annotations

import json
import math
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panst3r_b200 import ops  # noqa: E402

RESULTS = []


def report(name, got, ref, tol=2e-2):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-12
    nan = bool(torch.isnan(got).any().item())
    ok = (not nan) and err <= tol * scale
    RESULTS.append(dict(name=name, ok=ok, max_err=err, ref_max=scale, rel=err / scale, nan=nan))
    print(f"{'PASS' if ok else 'FAIL'} {name}: max_err={err:.4e} ref_max={scale:.4e} rel={err/scale:.3e} nan={nan}", flush=True)
    return ok


def guarded(name):
    def deco(fn):
        def run(*a, **k):
            try:
                fn(*a, **k)
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                RESULTS.append(dict(name=name, ok=False, error=repr(e)))
                print(f"ERROR {name}: {e!r}", flush=True)
                traceback.print_exc()
        return run
    return deco


def rnd(*shape, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, device="cuda") * scale).to(dtype)


@guarded("gemm")
def diag_gemm():
    torch.manual_seed(0)
    for (M, N, K) in [(128, 256, 64), (128, 256, 128), (256, 512, 256), (300, 200, 96), (1000, 3072, 1024),
                      (768, 768, 3072), (200, 768, 768), (12288, 1024, 1024), (130, 72, 2816)]:
        a = rnd(M, K)
        w = rnd(N, K, scale=K ** -0.5)
        bias = torch.randn(N, device="cuda")
        ref = a.float() @ w.float().t() + bias
        out = ops.gemm(a, w, bias=bias, out_dtype=torch.float32)
        report(f"gemm f32 M{M} N{N} K{K}", out, ref, tol=1e-3)
        out = ops.gemm(a, w, bias=bias)
        report(f"gemm bf16 M{M} N{N} K{K}", out, ref, tol=1e-2)
    # epilogues
    M, N, K = 512, 1024, 512
    a, w = rnd(M, K), rnd(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    res = rnd(M, N)
    ls = torch.randn(N, device="cuda")
    ref = torch.nn.functional.gelu(a.float() @ w.float().t() + bias)
    report("gemm gelu", ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, out_dtype=torch.float32), ref, tol=1e-3)
    ref = (a.float() @ w.float().t() + bias) * ls + res.float()
    report("gemm layerscale+residual", ops.gemm(a, w, bias=bias, col_scale=ls, residual=res, out_dtype=torch.float32), ref, tol=1e-3)
    ref = torch.relu(a.float() @ w.float().t() + bias)
    report("gemm relu", ops.gemm(a, w, bias=bias, act=ops.ACT_RELU, out_dtype=torch.float32), ref, tol=1e-3)
    # strided A (slice of a wider buffer) and strided out
    big = rnd(M, 2 * K)
    a_s = big[:, K:]
    outbuf = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a_s, w, out=outbuf[:, N:])
    report("gemm strided A/out", outbuf[:, N:], a_s.float() @ w.float().t(), tol=1e-2)
    report("gemm strided out untouched", outbuf[:, :N], torch.zeros(M, N, device="cuda"), tol=0)
    # transposed store (mask einsum layout): rows = pixels of V views, cols = queries
    V, HW, Q, Cm = 2, 1536, 200, 256
    feats = rnd(V * HW, Cm)
    emb = rnd(Q, Cm, scale=Cm ** -0.5)
    out = torch.empty(V, Q, HW, device="cuda", dtype=torch.float32)
    ops.gemm(feats, emb, out=out, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=HW, batch_stride=Q * HW, ldt=HW)
    ref = torch.einsum("qc,vpc->vqp", emb.float(), feats.float().view(V, HW, Cm))
    report("gemm transposed store (mask einsum)", out, ref, tol=1e-3)
    # pixel shuffle store
    B, gh, gw, Cout = 2, 6, 8, 64
    a = rnd(B * gh * gw, 128)
    w = rnd(Cout * 4, 128, scale=128 ** -0.5)
    out = torch.empty(B * 2 * gh * 2 * gw, Cout, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, out=out, store_mode=ops.STORE_PIXSHUF2, grid=(gh, gw))
    y = (a.float() @ w.float().t()).view(B, gh, gw, Cout * 4).permute(0, 3, 1, 2)
    ref = torch.nn.functional.pixel_shuffle(y, 2).permute(0, 2, 3, 1).reshape(B * 2 * gh * 2 * gw, Cout)
    report("gemm pixel_shuffle store", out, ref, tol=1e-2)
    # depth-to-space store (pointmap head), weights permuted (i, j, c)
    P, Cc = 16, 7
    a = rnd(B * gh * gw, 768)
    w_ref = rnd(Cc * P * P, 768, scale=768 ** -0.5)  # rows ordered (c, i, j) as in the reference LinearHead
    bias_ref = torch.randn(Cc * P * P, device="cuda")
    w_perm = w_ref.view(Cc, P, P, 768).permute(1, 2, 0, 3).reshape(P * P * Cc, 768).contiguous()
    b_perm = bias_ref.view(Cc, P, P).permute(1, 2, 0).reshape(-1).contiguous()
    out = torch.empty(B, gh * P, gw * P, Cc, device="cuda", dtype=torch.float32)
    ops.gemm(a, w_perm, bias=b_perm, out=out, store_mode=ops.STORE_D2S, grid=(gh, gw), d2s=(P, Cc))
    y = (a.float() @ w_ref.float().t() + bias_ref).view(B, gh, gw, Cc * P * P).permute(0, 3, 1, 2)
    ref = torch.nn.functional.pixel_shuffle(y, P).permute(0, 2, 3, 1)
    report("gemm depth-to-space store", out, ref, tol=1e-3)


def rope_ref(t, pos, base=100.0):
    # t [B, N, H, D] float, pos [B, N, 2]
    B, N, H, D = t.shape
    Q = D // 4
    inv = base ** (-torch.arange(Q, device=t.device, dtype=torch.float32) / Q)
    out = t.clone()
    for half in range(2):
        ang = pos[..., half].float()[..., None] * inv  # B N Q
        c, s = ang.cos()[:, :, None, :], ang.sin()[:, :, None, :]
        u = t[..., half * D // 2: half * D // 2 + Q]
        v = t[..., half * D // 2 + Q: half * D // 2 + 2 * Q]
        out[..., half * D // 2: half * D // 2 + Q] = u * c - v * s
        out[..., half * D // 2 + Q: half * D // 2 + 2 * Q] = v * c + u * s
    return out


@guarded("rope")
def diag_rope():
    torch.manual_seed(1)
    B, N, H, D = 2, 48, 12, 64
    t = rnd(B, N, H, D)
    ys, xs = torch.meshgrid(torch.arange(6), torch.arange(8), indexing="ij")
    pos = torch.stack([ys.flatten(), xs.flatten()], -1)[None].expand(B, -1, -1).contiguous().to("cuda", torch.int32)
    ref = rope_ref(t.float(), pos)
    got = ops.rope2d_(t.clone(), pos)
    report("rope2d standalone", got, ref, tol=1e-2)
    # fused in the QKV GEMM epilogue
    dim = H * D
    x = rnd(B * N, dim)
    w = rnd(3 * dim, dim, scale=dim ** -0.5)
    bias = torch.randn(3 * dim, device="cuda")
    maxpos = 8
    Q = D // 4
    inv = 100.0 ** (-torch.arange(Q, device="cuda", dtype=torch.float32) / Q)
    ang = torch.arange(maxpos, device="cuda", dtype=torch.float32)[:, None] * inv
    cs = torch.stack([ang.cos(), ang.sin()], -1).contiguous()
    got = ops.gemm(x, w, bias=bias, out_dtype=torch.float32, rope=(cs, pos.view(-1, 2), 2 * dim))
    qkv = (x.float() @ w.float().t() + bias).view(B, N, 3, H, D)
    ref = torch.stack([rope_ref(qkv[:, :, 0], pos), rope_ref(qkv[:, :, 1], pos), qkv[:, :, 2]], 2).view(B * N, 3 * dim)
    report("rope fused in gemm epilogue", got, ref, tol=1e-3)


def attn_ref(q, k, v, scale, mask=None):
    # q [B,Nq,H,hd], k/v [B,Nk,H,hd]; mask bool [B,Nq,Nk] True = blocked
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale
    if mask is not None:
        s = s.masked_fill(mask[:, None], float("-inf"))
    p = s.softmax(-1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v.float()).reshape(q.shape[0], q.shape[1], -1)


@guarded("attention")
def diag_attention():
    torch.manual_seed(2)
    for (B, H, Nq, Nk, hd, splits) in [(1, 1, 128, 128, 64, 1), (1, 2, 128, 256, 64, 1), (2, 4, 300, 500, 64, 1),
                                       (2, 12, 768, 768, 64, 1), (1, 12, 768, 1536, 64, 2), (1, 1, 128, 128, 96, 1),
                                       (2, 4, 200, 1000, 96, 1), (1, 8, 200, 3000, 96, 4), (1, 16, 769, 769, 64, 1)]:
        q, k, v = rnd(B, Nq, H, hd), rnd(B, Nk, H, hd), rnd(B, Nk, H, hd)
        ref = attn_ref(q, k, v, hd ** -0.5)
        got = ops.attention(q, k, v, kv_splits=splits)
        report(f"attention B{B} H{H} Nq{Nq} Nk{Nk} hd{hd} splits{splits}", got, ref, tol=2e-2)
    # fused QKV layout [B, N, 3, H, hd]
    B, N, H, hd = 2, 384, 12, 64
    qkv = rnd(B, N, 3, H, hd)
    got = ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2])
    report("attention fused-qkv strides", got, attn_ref(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], hd ** -0.5), tol=2e-2)
    # shared K/V across the batch (render pass)
    q, k, v = rnd(3, 256, 12, 64), rnd(1, 640, 12, 64), rnd(1, 640, 12, 64)
    got = ops.attention(q, k, v)
    report("attention shared kv", got, attn_ref(q, k.expand(3, -1, -1, -1), v.expand(3, -1, -1, -1), 0.125), tol=2e-2)
    # block mask
    B, H, Nq, Nk, hd = 1, 8, 200, 1536, 96
    q, k, v = rnd(B, Nq, H, hd), rnd(B, Nk, H, hd), rnd(B, Nk, H, hd)
    mask = torch.rand(B, Nq, Nk, device="cuda") < 0.6
    mask[:, 5] = False
    words = ((Nk + 127) // 128) * 4
    mb = torch.zeros(B, Nq, words * 32, device="cuda", dtype=torch.bool)
    mb[:, :, :Nk] = mask
    weights = (1 << torch.arange(32, device="cuda", dtype=torch.int64))
    bits = (mb.view(B, Nq, words, 32).to(torch.int64) * weights).sum(-1)
    bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32).contiguous()
    for splits in (1, 3):
        got = ops.attention(q, k, v, mask_bits=bits, kv_splits=splits)
        report(f"attention masked hd96 splits{splits}", got, attn_ref(q, k, v, hd ** -0.5, mask), tol=2e-2)


@guarded("elementwise")
def diag_elementwise():
    torch.manual_seed(3)
    x = rnd(1000, 1024)
    g, b = torch.randn(1024, device="cuda"), torch.randn(1024, device="cuda")
    ref = torch.nn.functional.layer_norm(x.float(), (1024,), g, b, 1e-6)
    report("layernorm bf16", ops.layernorm(x, g, b, 1e-6), ref, tol=1e-2)
    report("layernorm f32 out", ops.layernorm(x, g, b, 1e-6, out_dtype=torch.float32), ref, tol=1e-4)
    add = rnd(1000, 1024)
    sum_out = torch.empty_like(x)
    got = ops.layernorm(x, g, b, 1e-5, add=add, sum_out=sum_out, out_dtype=torch.float32)
    report("layernorm add", got, torch.nn.functional.layer_norm(x.float() + add.float(), (1024,), g, b, 1e-5), tol=1e-4)
    report("layernorm sum_out", sum_out, x.float() + add.float(), tol=1e-2)
    img = torch.rand(2, 3, 64, 96, device="cuda") * 2 - 1
    got = ops.patchify(img, 16)
    ref = torch.nn.functional.unfold(img, 16, stride=16).transpose(1, 2).reshape(-1, 768)
    report("patchify", got, ref, tol=1e-2)
    Ho, Wo = 64 // 16 * 14, 96 // 16 * 14
    got = ops.dino_preprocess_patchify(img, Ho, Wo, 14, 592)
    y = img * 0.5 + 0.5
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda").view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda").view(1, 3, 1, 1)
    y = torch.nn.functional.interpolate((y - mean) / std, size=(Ho, Wo), mode="bilinear", align_corners=False)
    ref = torch.nn.functional.unfold(y, 14, stride=14).transpose(1, 2).reshape(-1, 588)
    report("dino preprocess+patchify", got[:, :588], ref, tol=1e-2)
    report("dino patchify pad zero", got[:, 588:], torch.zeros_like(got[:, 588:]), tol=0)
    f = rnd(2, 16, 24, 64)
    ref = torch.nn.functional.interpolate(f.float().permute(0, 3, 1, 2), size=(2, 3), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    report("center_pool8 == bilinear/8", ops.center_pool8(f), ref, tol=1e-2)
    lg = torch.randn(200, 1536, device="cuda")
    lg[7] = -lg[7].abs()
    bits = ops.attn_mask_bits(lg, 1500)
    blocked = lg[:, :1500] < 0
    blocked[blocked.all(-1)] = False
    unpacked = ((bits[0].to(torch.int64)[..., None] >> torch.arange(32, device="cuda")) & 1).bool().view(200, -1)
    report("attn_mask_bits", unpacked[:, :1500].float(), blocked.float(), tol=0)
    report("attn_mask_bits tail clear", unpacked[:, 1500:].float(), torch.zeros_like(unpacked[:, 1500:]).float(), tol=0)
    xx = torch.randn(200, 768, device="cuda")
    report("l2norm_rows", ops.l2norm_rows(xx, 1e-7, torch.float32), xx / (xx.norm(dim=-1, keepdim=True) + 1e-7), tol=1e-5)
    t = rnd(2, 100, 48)
    report("nhwc_to_nchw", ops.nhwc_to_nchw_f32(t), t.float().transpose(1, 2), tol=0)
    a_, b_ = rnd(6, 10, 64), rnd(10, 64)
    report("add_bcast", ops.add_bcast(a_.view(60, 64), b_), (a_.float() + b_.float()).view(60, 64), tol=1e-2)


def bench_gemm():
    print("--- gemm timing ---", flush=True)
    for (M, N, K) in [(12288, 3072, 1024), (12288, 4096, 1024), (12288, 1024, 4096), (12288, 1024, 1024), (768, 768, 768)]:
        a, w = rnd(M, K), rnd(N, K)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, w, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(a, w, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"gemm M{M} N{N} K{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        e0.record()
        for _ in range(10):
            torch.matmul(a, w.t())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"   cuBLAS: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)


def bench_attention():
    print("--- attention timing ---", flush=True)
    for (B, H, Nq, Nk, hd) in [(16, 16, 768, 768, 64), (16, 12, 768, 12288, 64), (1, 12, 768, 6144, 64), (16, 4, 49152 // 8, 768, 96)]:
        q, k, v = rnd(B, Nq, H, hd), rnd(B, Nk, H, hd), rnd(B, Nk, H, hd)
        for _ in range(2):
            ops.attention(q, k, v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.attention(q, k, v)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 4.0 * B * H * Nq * Nk * hd
        print(f"attention B{B} H{H} Nq{Nq} Nk{Nk} hd{hd}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "rope", "elementwise", "attention", "bench"]
    print(torch.cuda.get_device_name(0), flush=True)
    t0 = time.time()
    if "gemm" in which:
        diag_gemm()
    if "rope" in which:
        diag_rope()
    if "elementwise" in which:
        diag_elementwise()
    if "attention" in which:
        diag_attention()
    if "bench" in which:
        try:
            bench_gemm()
            bench_attention()
        except Exception as e:  # noqa: BLE001
            print("bench error", repr(e))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/diag.json", "w") as f:
        json.dump(RESULTS, f, indent=1)
    nfail = sum(1 for r in RESULTS if not r.get("ok"))
    print(f"diag done in {time.time()-t0:.1f}s: {len(RESULTS)-nfail} pass, {nfail} fail", flush=True)
