"""GPU parity of the reference-precision head mode (PanopticDecoder.precision == "fp32", the default).

The reference runs the panoptic head in fp32 (src/panst3r/panst3r.py:236-245); the CUDA head keeps activations and
weights as split bf16 pairs (ops.Split: hi + lo, 16 mantissa bits) and accumulates the three bf16 tensor-core products
in fp32.  Bars (max |a - b| / max |b| per tensor, the north star's definition):
  * split kernels vs an fp64 statement of the op: 3e-5 (observed ~5e-6; a plain bf16 operand pair gives 4e-3);
  * the whole head on identical inputs vs the REFERENCE-generated goldens (tests/golden/head_*.pt, fp32): 1e-3 on every
    output of all seven prediction heads under the reference's own sign decisions; free-running, decisions may differ
    only where the reference's logit is inside the rounding band (|logit| < 1e-4 of the maximum), and whenever none
    differs the free-running outputs meet 1e-3 too;
  * argmax instance ids (engine/postprocess.py:18-27, 63, 77): exact on every pixel whose golden top-2 margin exceeds
    twice the measured score error; > 90 % of the pixels are decidable on the conditioned fixture.
"""
import os

import pytest
import torch

from helpers import CLASSES, GOLDEN, build_oracle_head, golden_files, head_inputs, relmax

pytestmark = pytest.mark.gpu

TOL_SPLIT = 3e-5
TOL_HEAD = 1e-3  # BASELINE.json: logits <= 1e-3 vs ref


@pytest.fixture(scope="module")
def ops():
    from panst3r_b200 import lib, ops as o
    assert lib.load().pst3r_check_device() == 0, lib.load().pst3r_last_error()
    torch.manual_seed(0)
    return o


def r64(*shape, scale=1.0):
    return torch.randn(*shape, device="cuda", dtype=torch.float64) * scale


def split_of(ops, x64):
    return ops.Split.from_float(x64.float().contiguous())


def val(s):
    """fp64 value of a Split / tensor as stored"""
    return (s.hi.double() + s.lo.double()) if hasattr(s, "hi") else s.double()


def test_split_roundtrip_precision(ops):
    x = r64(300, 768, scale=3.0)
    s = split_of(ops, x)
    assert s.packed() and s.hi.shape == (300, 768) and s.lo_off == 768
    # hi + lo reproduces the fp32 value to 2^-17 relative (bf16 x 2); the conversion back to fp32 is what was stored
    assert ((val(s) - x.float().double()).abs() / x.abs().clamp_min(1e-3)).max().item() < 2 ** -15
    back = ops.convert(s, torch.empty(300, 768, device="cuda"))
    assert torch.equal(back.double(), val(s).float().double())
    b16 = ops.convert(s, torch.empty(300, 768, device="cuda", dtype=torch.bfloat16))
    assert relmax(b16, x) < 4e-3


@pytest.mark.parametrize("M,N,K", [(200, 768, 768), (256, 512, 256), (300, 200, 96), (130, 72, 2816), (1000, 3072, 1024),
                                   (768, 768, 3072), (4864, 1024, 1024), (129, 264, 200)])
def test_split_gemm(ops, M, N, K):
    """A_hi B_hi + A_lo B_hi + A_hi B_lo with fp32 accumulation vs fp64; (4864, 1024, 1024) takes the 2-CTA kernel."""
    a, w, bias = r64(M, K), r64(N, K, scale=K ** -0.5), r64(N)
    ref = a.float().double() @ w.float().double().t() + bias.float().double()
    sa, sw = split_of(ops, a), split_of(ops, w)
    got32 = ops.gemm(sa, sw, bias=bias.float(), out_dtype=torch.float32)
    assert relmax(got32, ref) < TOL_SPLIT
    got = ops.gemm(sa, sw, bias=bias.float(), out_dtype="split")
    assert relmax(val(got), ref) < TOL_SPLIT
    # the same product on plain bf16 operands is two orders of magnitude coarser: the split terms do the work
    plain = ops.gemm(a.float().bfloat16(), w.float().bfloat16(), bias=bias.float(), out_dtype=torch.float32)
    assert relmax(plain, ref) > 20 * relmax(got32, ref)
    # two-term form: exactly representable (bf16) activations x split weights
    ab = a.float().bfloat16()
    got2 = ops.gemm(ab, sw, bias=bias.float(), out_dtype=torch.float32)
    assert relmax(got2, ab.double() @ w.float().double().t() + bias.float().double()) < TOL_SPLIT


def test_split_gemm_epilogues(ops):
    M, N, K = 384, 512, 256
    a, w, bias, res = r64(M, K), r64(N, K, scale=K ** -0.5), r64(N), r64(M, N)
    sa, sw, sres = split_of(ops, a), split_of(ops, w), split_of(ops, res)
    y = a.float().double() @ w.float().double().t() + bias.float().double()
    gelu = torch.nn.functional.gelu(y)
    assert relmax(val(ops.gemm(sa, sw, bias=bias.float(), act=ops.ACT_GELU, out_dtype="split")), gelu) < TOL_SPLIT
    assert relmax(val(ops.gemm(sa, sw, bias=bias.float(), act=ops.ACT_RELU, out_dtype="split")), y.relu()) < TOL_SPLIT
    assert relmax(val(ops.gemm(sa, sw, bias=bias.float(), residual=sres, out_dtype="split")), y + val(sres)) < TOL_SPLIT
    assert relmax(ops.gemm(sa, sw, bias=bias.float(), residual=res.float(), out_dtype=torch.float32), y + res.float().double()) < TOL_SPLIT
    # in place: out aliases the split residual
    x = split_of(ops, res)
    ops.gemm(sa, sw, bias=bias.float(), residual=x, out=x)
    assert relmax(val(x), y + val(sres)) < TOL_SPLIT
    # pixel_shuffle(2) store of split pairs
    B, gh, gw, Cout = 2, 6, 8, 64
    a2, w2 = r64(B * gh * gw, 128), r64(Cout * 4, 128, scale=128 ** -0.5)
    out = ops.Split.empty((B * 2 * gh * 2 * gw, Cout), "cuda")
    ops.gemm(split_of(ops, a2), split_of(ops, w2), out=out, store_mode=ops.STORE_PIXSHUF2, grid=(gh, gw))
    y2 = (a2.float().double() @ w2.float().double().t()).view(B, gh, gw, Cout * 4).permute(0, 3, 1, 2)
    assert relmax(val(out), torch.nn.functional.pixel_shuffle(y2, 2).permute(0, 2, 3, 1).reshape(-1, Cout)) < TOL_SPLIT
    # transposed stores: fp32 planes (mask einsum) and split rows (V^T of the precise attention)
    V, HW, Q, Cm = 2, 24 * 16, 200, 256
    f, e = r64(V * HW, Cm), r64(Q, Cm, scale=Cm ** -0.5)
    mk = torch.empty(V, Q, HW, device="cuda")
    ops.gemm(split_of(ops, f), split_of(ops, e), out=mk, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=HW, batch_stride=Q * HW, ldt=HW)
    ref = (f.float().double() @ e.float().double().t()).view(V, HW, Q).transpose(1, 2)
    assert relmax(mk, ref) < TOL_SPLIT
    vT = ops.Split.empty((Q, V * HW), "cuda")
    ops.gemm(split_of(ops, f), split_of(ops, e), out=vT, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=V * HW, batch_stride=0, ldt=2 * V * HW)
    assert relmax(val(vT), (f.float().double() @ e.float().double().t()).t()) < TOL_SPLIT


def test_row_kernels_on_split_operands(ops):
    x, add = r64(333, 768, scale=2.0), r64(333, 768)
    g, b = torch.randn(768, device="cuda"), torch.randn(768, device="cuda")
    sx, sadd = split_of(ops, x), split_of(ops, add)
    ref = torch.nn.functional.layer_norm(val(sx) + val(sadd), (768,), g.double(), b.double(), 1e-5)
    so = ops.Split.empty((333, 768), "cuda")
    got = ops.layernorm(sx, g, b, 1e-5, add=sadd, sum_out=so)
    assert isinstance(got, ops.Split) and relmax(val(got), ref) < TOL_SPLIT and relmax(val(so), val(sx) + val(sadd)) < TOL_SPLIT
    assert relmax(ops.layernorm(sx, g, b, 1e-5, out_dtype=torch.float32), torch.nn.functional.layer_norm(val(sx), (768,), g.double(), b.double(), 1e-5)) < 1e-5
    # broadcast add with mixed kinds (split activations + fp32 sine position embedding)
    pe = torch.randn(37, 768, device="cuda")
    s9 = split_of(ops, r64(9 * 37, 768))
    assert relmax(val(ops.add_bcast(s9, pe)), val(s9) + pe.double().repeat(9, 1)) < TOL_SPLIT
    # centre pooling of a split feature map
    f = r64(2, 16, 24, 64)
    sf = split_of(ops, f.view(-1, 64)).view(2, 16, 24, 64)
    refp = torch.nn.functional.interpolate(val(sf).permute(0, 3, 1, 2), size=(2, 3), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    assert relmax(val(ops.center_pool8(sf)), refp) < TOL_SPLIT
    xx = torch.randn(200, 768, device="cuda")
    assert relmax(val(ops.l2norm_rows(xx, 1e-7, "split")), xx.double() / (xx.double().norm(dim=-1, keepdim=True) + 1e-7)) < TOL_SPLIT
    t = split_of(ops, r64(200, 48)).view(2, 100, 48)
    assert relmax(ops.nhwc_to_nchw_f32(t), val(t).transpose(1, 2)) < 1e-6


@pytest.mark.parametrize("Nk", [1400, 12, 201])
def test_precise_attention_pieces(ops, Nk):
    """The unfused reference-precision attention of the query decoder: per-head QK^T (batched split GEMM over head
    slices, K = 96 = 1.5 k-blocks), masked row softmax, per-head PV against V^T (K = Nk not a multiple of 64, nor of 8
    for tiny scenes: the lo parts then sit at the next multiple of 8)."""
    from test_gpu_kernels import pack_bits
    H, Q, hd = 8, 200, 96
    d = H * hd
    q, k, v = r64(Q, d), r64(Nk, 6 * d), r64(Nk, d)
    sq, sk = split_of(ops, q), split_of(ops, k)
    mask = torch.rand(1, Q, Nk, device="cuda") < 0.6
    mask[:, 5] = False
    mask[:, :, 0] = False  # no fully blocked row (the mask builder clears those, mask_transformer.py:172)
    bits = pack_bits(mask)
    layer = 3
    S = torch.empty(H, Q, Nk, device="cuda")
    ops.gemm_batched(sq.view(Q, H, hd).permute(1, 0, 2), sk[:, layer * d:(layer + 1) * d].view(Nk, H, hd).permute(1, 0, 2),
                     alpha=hd ** -0.5, out=S)
    kq = val(sk)[:, layer * d:(layer + 1) * d].view(Nk, H, hd)
    ref_s = torch.einsum("qhd,khd->hqk", val(sq).view(Q, H, hd), kq) * hd ** -0.5
    assert relmax(S, ref_s) < TOL_SPLIT
    P = ops.softmax_rows(S, Q, bits)
    ref_p = S.double().masked_fill(mask, float("-inf")).softmax(-1)
    assert P.lo_off % 8 == 0 and relmax(val(P), ref_p) < TOL_SPLIT and val(P)[:, mask[0]].abs().max().item() == 0
    # V^T through the transposed split store, then O_h = P_h V_h
    w = r64(d, 64, scale=0.125)
    src = r64(Nk, 64)
    vT = ops.Split.empty((d, Nk), "cuda", align=8)
    ops.gemm(split_of(ops, src), split_of(ops, w), out=vT, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=Nk, batch_stride=0,
             ldt=vT.hi.stride(0))
    o = ops.Split.empty((Q, d), "cuda")
    ops.gemm_batched(P, vT.view(H, hd, Nk), out=o.view(Q, H, hd).permute(1, 0, 2))
    ref_o = torch.einsum("hqk,hdk->qhd", val(P), val(vT).view(H, hd, Nk)).reshape(Q, d)
    assert relmax(val(o), ref_o) < TOL_SPLIT
    # no mask, tiny key set (the decoder's 200 x 200 self-attention)
    S2 = torch.randn(H, Q, Q, device="cuda")
    assert relmax(val(ops.softmax_rows(S2, Q)), S2.double().softmax(-1)) < TOL_SPLIT


def _cuda_head(variant="v1", precision="fp32", deep_supervision=True, cls_logit_scale=None):
    from panst3r_b200.modules.panoptic import InputMixer, LoftUpUpscaler, PanopticDecoder, PixelShuffleUpscaler
    o = build_oracle_head(variant, cls_logit_scale=cls_logit_scale)
    if variant == "v1":
        m = PanopticDecoder(upscaler=PixelShuffleUpscaler(input_dim=2816), precision=precision, deep_supervision=deep_supervision)
    else:
        m = PanopticDecoder(input_mixer=InputMixer([512, 512], 16, 2816, 768), upscaler=LoftUpUpscaler(input_dim=768, dim=384),
                            mask_dim=384, precision=precision, deep_supervision=deep_supervision)
    m = m.eval()
    m.load_state_dict(o.state_dict(), strict=True)  # the exact fp32 weights the golden run used
    m = m.cuda()
    m.text_encoder.class_embeddings = o.text_encoder.class_embeddings
    return o, m


BAND = 1e-4  # |reference pooled logit| / max below which a sign decision is inside the split-bf16 rounding band


def _ref_bits(pooled):
    """Reference block masks (mask_transformer.py:264-272, 172) from the golden's downsampled logits: blocked iff the
    logit is negative; a query whose keys are all blocked attends everywhere."""
    from test_gpu_kernels import pack_bits
    out = []
    for p in pooled:
        blocked = p < 0
        blocked[blocked.all(-1)] = False
        out.append(pack_bits(blocked[None].cuda()))
    return out


def _unpack(bits, nk):
    return ((bits[0].to(torch.int64)[..., None] >> torch.arange(32, device=bits.device)) & 1).bool().flatten(1)[:, :nk].cpu()


def _run_head(m, feats, imgs, pos, ts, forced=None, **kw):
    """-> (output dict, block masks the six layers used as bool (Q, Nk))"""
    mt = m.mask_transformer
    mt.mask_override, mt.bits_record = forced, []
    try:
        out = m(tuple(f.cuda() for f in feats), imgs.cuda(), pos.cuda(), ts, CLASSES, **kw)
        used = list(mt.bits_record)
    finally:
        mt.mask_override, mt.bits_record = None, None
    return out, used


def _errors(out, g):
    e = {"pred_masks": relmax(out["pred_masks"], g["pred_masks"]), "pred_logits": relmax(out["pred_logits"], g["pred_logits"]),
         "out_queries": relmax(out["out_queries"], g["out_queries"])}
    for i, gl in enumerate(g["aux_logits"]):
        e[f"aux{i}_logits"] = relmax(out["aux_outputs"][i]["pred_logits"], gl)
    if "aux0_masks" in g:
        e["aux0_masks"] = relmax(out["aux_outputs"][0]["pred_masks"], g["aux0_masks"])
    return e


def _check_decisions(used, pooled):
    """The FIRST layer at which the free-running CUDA decoder takes a different sign decision than the reference may only
    do so for logits inside the rounding band of zero (after that point the two runs follow different — equally valid —
    branches of a discontinuous function and later logits differ by more than rounding).  Returns the number of
    differing decisions at that first layer (0: identical decisions throughout)."""
    for i, (bits, p) in enumerate(zip(used, pooled)):
        mine = _unpack(bits, p.shape[1])
        theirs = p < 0
        theirs[theirs.all(-1)] = False
        diff = mine != theirs
        if diff.any():
            worst = (p.abs()[diff] / p.abs().max()).max().item()
            assert worst < BAND, f"layer {i}: a decision differs at |logit|/max = {worst:.2e} (outside the rounding band)"
            return int(diff.sum())
    return 0


@pytest.mark.parametrize("path", golden_files("head_v*.pt"))
def test_precise_head_against_reference_golden(path):
    """Every output of the v1 and v2 heads (PixelShuffle upscaler; InputMixer + LoftUp; query decoder;
    prediction heads; mask einsum) on identical fp32 inputs and weights vs the outputs of the REFERENCE's own modules at
    the north-star tolerance.

    The six-layer query decoder feeds sign(mask logit) decisions back as attention masks (mask_transformer.py:264-272):
    it is a DISCONTINUOUS function, and logits closer to zero than the arithmetic's resolution (2e-5 of the maximum for
    16-mantissa-bit split operands; two fp32 implementations meet the same issue at 1e-7) can land on either side.
    So: (1) with the reference's decisions forced, all seven heads must agree to 1e-3 — arithmetic parity;
    (2) free-running, every decision that differs from the reference's must sit inside the rounding band, and when none
    differs the free-running outputs must agree to 1e-3 as well."""
    g = torch.load(path)
    o, m = _cuda_head(g["variant"])
    assert m.precision == "fp32"
    feats, imgs, pos, ts = head_inputs(g["V"], g["H"], g["W"], g["input_seed"], portrait=g["portrait"])
    H, Wd = g["H"], g["W"]
    x_up = torch.cat(feats, -1)[0].cuda()
    if g["variant"] == "v2":  # the upscaler consumes the InputMixer's tokens (rows a9, a11)
        from panst3r_b200 import ops
        V, hs, ws = g["V"], H // 16, Wd // 16
        xm = m.input_mixer.forward_rows(ops.Split.from_float(x_up.reshape(V * hs * ws, -1).contiguous()), V, hs, ws, precise=True)
        x_up = ops.convert(xm, torch.empty(xm.shape, device="cuda")).view(V, hs * ws, -1)
    if g["portrait"]:
        f16, f2 = m.upscaler((x_up, None), (Wd, H), precise=True)
        f16, f2 = [f16[0].swapaxes(2, 3)], f2.swapaxes(2, 3)
    else:
        f16, f2 = m.upscaler((x_up, imgs[0].cuda()), (H, Wd), precise=True)
    tol_up = TOL_HEAD if g["fpn0"].dtype == torch.float32 else 2e-3  # the larger fixture stores these in fp16
    assert relmax(f16[0], g["fpn0"]) < tol_up and relmax(f2, g["mask_feats"]) < tol_up
    # (1) the reference's sign decisions forced
    forced, _ = _run_head(m, feats, imgs, pos, ts, forced=_ref_bits(g["pooled_logits"]))
    e_forced = _errors(forced, g)
    print(os.path.basename(path), "forced decisions:", {k: f"{v:.1e}" for k, v in e_forced.items()})
    for k, v in e_forced.items():
        assert v < (2e-3 if (k == "aux0_masks" and g["aux0_masks"].dtype == torch.float16) else TOL_HEAD), (k, v)
    # (2) free-running
    out, used = _run_head(m, feats, imgs, pos, ts)
    flips = _check_decisions(used, g["pooled_logits"])
    e_free = _errors(out, g)
    print(os.path.basename(path), f"free-running, {flips} decisions differ:", {k: f"{v:.1e}" for k, v in e_free.items()})
    assert out["pred_masks"].dtype == torch.float32 and out["out_queries"].dtype == torch.float32
    assert e_free["aux0_logits"] < TOL_HEAD and e_free.get("aux0_masks", 0.0) < 2e-3  # no decision precedes head 0
    if flips == 0:
        for k, v in e_free.items():
            assert v < (2e-3 if (k == "aux0_masks" and g["aux0_masks"].dtype == torch.float16) else TOL_HEAD), (k, v)
    mq = m(tuple(f.cuda() for f in feats), imgs.cuda(), pos.cuda(), ts, CLASSES, memory_queries=out["out_queries"])
    assert torch.equal(mq["pred_masks"], out["pred_masks"]) and set(mq) == {"pred_logits", "pred_masks"}


def _ids(out, H, W):
    scores = out["pred_logits"].sigmoid().max(-1).values[0]
    up = torch.nn.functional.interpolate(out["pred_masks"][0].sigmoid(), size=(H, W), mode="bilinear", align_corners=False)
    return (scores[None, :, None, None] * up).cpu()


def test_argmax_ids_exact_on_conditioned_fixture():
    """a well-conditioned fixture — class logits O(1), mask logits O(1), and every one of the
    reference's 57 600 sign(mask logit) decisions at least 6.8e-5 of the largest logit away from zero (input seed chosen
    by tools/seed_search.py) — and the post-processing front half's score-weighted argmax (engine/postprocess.py:18-27,
    63, 77).  The FREE-RUNNING six-layer decoder takes exactly the reference's decisions, its final mask / class logits
    agree to 1e-3, more than 90 % of the pixels are decidable and the instance ids are bit-exact on every one of them.
    The same holds with the reference's decisions forced."""
    g = torch.load(os.path.join(GOLDEN, "head_v1_conditioned.pt"))
    o, m = _cuda_head("v1", cls_logit_scale=g["cls_logit_scale"])
    feats, imgs, pos, ts = head_inputs(g["V"], g["H"], g["W"], g["input_seed"])
    gs = g["pred_logits"].sigmoid().max(-1).values[0]
    gup = torch.nn.functional.interpolate(g["pred_masks"][0].sigmoid(), size=(g["H"], g["W"]), mode="bilinear", align_corners=False)
    gw = gs[None, :, None, None] * gup
    results = {}
    for name, forced in (("forced", _ref_bits(g["pooled_logits"])), ("free", None)):
        out, used = _run_head(m, feats, imgs, pos, ts, forced=forced)
        flips = 0 if forced is not None else _check_decisions(used, g["pooled_logits"])
        e = _errors(out, g)
        weighted = _ids(out, g["H"], g["W"])
        tol = 2.0 * (weighted - gw).abs().amax(dim=1)  # twice the measured per-pixel score error
        safe = g["margin"] > tol
        ids = weighted.argmax(1)
        frac = safe.float().mean().item()
        print(f"conditioned fixture, {name}: {flips} decisions differ; final masks {e['pred_masks']:.1e}, class logits "
              f"{e['pred_logits']:.1e}; decidable pixels {frac:.4f}; ids equal on {(ids == g['ids'].long()).float().mean().item():.4f}")
        assert torch.equal(ids[safe], g["ids"].long()[safe])
        results[name] = (e, frac, flips, out, safe)
    assert g["min_decision_margin"] > 5e-5
    for name in ("forced", "free"):
        e, frac, flips, out, safe = results[name]
        assert flips == 0, f"{name}: {flips} sign decisions differ on the well-conditioned fixture"
        assert e["pred_masks"] < TOL_HEAD and e["pred_logits"] < TOL_HEAD and frac > 0.9, (name, e, frac)
    # through the CUDA post-processing front half as well (fused sigmoid -> bilinear -> score-weighted argmax)
    from panst3r_b200 import ops
    sc, _ = ops.class_scores(out["pred_logits"][0].contiguous())
    keep = torch.arange(200, device="cuda", dtype=torch.int32)
    z = torch.zeros(200, device="cuda", dtype=torch.int32)
    ids_k, _ = ops.panoptic_argmax(out["pred_masks"][0].contiguous(), keep, sc, (g["H"], g["W"]), 0.25, z, z.clone())
    assert torch.equal(ids_k.cpu().long()[safe], g["ids"].long()[safe])


def test_precise_vs_bf16_head_modes():
    """Both precision modes on the same inputs: the bf16 mode stays within its documented 2e-2 on the first head, the
    fp32 mode within 1e-3 everywhere; outputs have identical shapes / dtypes / keys."""
    g = torch.load(os.path.join(GOLDEN, "head_v1_V2_32x48.pt"))
    feats, imgs, pos, ts = head_inputs(g["V"], g["H"], g["W"], g["input_seed"])
    res = {}
    for prec in ("fp32", "bf16"):
        _, m = _cuda_head("v1", precision=prec)
        res[prec] = m(tuple(f.cuda() for f in feats), imgs.cuda(), pos.cuda(), ts, CLASSES)
    assert set(res["fp32"]) == set(res["bf16"])
    assert res["fp32"]["pred_masks"].shape == res["bf16"]["pred_masks"].shape
    e32 = relmax(res["fp32"]["aux_outputs"][0]["pred_masks"], g["aux0_masks"])
    e16 = relmax(res["bf16"]["aux_outputs"][0]["pred_masks"], g["aux0_masks"])
    assert e32 < TOL_HEAD and e16 < 2e-2 and e16 > 10 * e32


def test_precise_head_multi_ar_against_reference_golden():
    from oracle.make_golden import multi_ar_head_inputs
    g = torch.load(os.path.join(GOLDEN, "head_v1_multi_ar.pt"))
    o, m = _cuda_head("v1")
    in_feats, imgs, pos, ts = multi_ar_head_inputs()
    cu = lambda lst: [t.cuda() for t in lst]  # noqa: E731
    out = m(tuple(cu(f) for f in in_feats), cu(imgs), cu(pos), ts, CLASSES, multi_ar=True)
    for a, b, c, d in zip(out["pred_masks"], g["pred_masks"], out["aux_outputs"][0]["pred_masks"], g["aux0_masks"]):
        assert a.shape == b.shape and relmax(c, d) < TOL_HEAD  # first head: no sign decision precedes it
        print(f"multi_ar stack, free-running final masks: {relmax(a, b):.1e}")
        assert relmax(a, b) < 5e-2  # free-running: sign decisions inside the rounding band may differ (see above)
    assert relmax(out["aux_outputs"][0]["pred_logits"], g["aux0_logits"]) < TOL_HEAD
    mq = m(tuple(cu(f) for f in in_feats), cu(imgs), cu(pos), ts, CLASSES, multi_ar=True, memory_queries=g["out_queries"].cuda())
    for a, b in zip(mq["pred_masks"], g["pred_masks"]):
        assert relmax(a, b) < TOL_HEAD


def test_precise_head_batched_scenes():
    """B > 1 (the reference's forward is batch-generic, panst3r.py:286-296): two scenes in one call equal two calls."""
    o, m = _cuda_head("v1", deep_supervision=False)
    f1, i1, p1, ts = head_inputs(2, 32, 48, seed=5)
    f2, i2, p2, _ = head_inputs(2, 32, 48, seed=6)
    cat = lambda a, b: torch.cat([a, b], 0).cuda()  # noqa: E731
    both = m(tuple(cat(a, b) for a, b in zip(f1, f2)), cat(i1, i2), cat(p1, p2), torch.cat([ts, ts], 0), CLASSES)
    one = m(tuple(f.cuda() for f in f2), i2.cuda(), p2.cuda(), ts, CLASSES)
    assert both["pred_masks"].shape[0] == 2 and both["out_queries"].shape == (200, 2, 768)
    assert torch.equal(both["pred_masks"][1:], one["pred_masks"]) and torch.equal(both["pred_logits"][1:], one["pred_logits"])
