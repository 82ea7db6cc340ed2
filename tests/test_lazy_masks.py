"""Lazy mask logits: post-processing that never materialises the (V, Q, h, w) mask-logit tensor the
reference builds (engine/postprocess.py:18-27, model/mask_transformer.py:279-280).

CPU: the tensor-like bookkeeping of `postprocess.LazyMasks` (shapes, batch / view indexing, band plans) and the argument
checks of the band entry point.
GPU: the band form of the argmax kernel equals the whole-map launch bit for bit; `LazyMasks` logits equal the eager
head's mask GEMM; `panoptic_inference_v1/_v2` on the lazy handle return the segments, ids and confidences of the
materialised path; `PanSt3R.forward_panoptic` equals forward() + panoptic_inference on the free-running model.
"""
import ctypes

import numpy as np
import pytest
import torch



def _lazy(V=3, hm=12, wm=16, C=8, Q=5, device="cpu"):
    from panst3r_b200.postprocess import LazyMasks
    g = torch.Generator().manual_seed(0)
    f = torch.randn(V, hm, wm, C, generator=g).to(torch.bfloat16).to(device)
    e = torch.randn(Q, C, generator=g).to(torch.bfloat16).to(device)
    return LazyMasks(f, e), f, e


def test_lazy_masks_is_shaped_and_indexed_like_the_tensor():
    from panst3r_b200.lib import Pst3rError
    from panst3r_b200.postprocess import LazyMasks
    lz, f, e = _lazy()
    assert lz.shape == (3, 5, 12, 16) and lz.dim() == 4 and len(lz) == 3 and lz.dtype == torch.float32
    b = lz[None]
    assert b.shape == (1, 3, 5, 12, 16) and b.dim() == 5 and len(b) == 1
    assert b[0].shape == lz.shape and b[0:1] is b and b[0, 2].shape == (5, 12, 16) and b[0, 2].dim() == 3
    assert b[0, 1].feats.data_ptr() == f[1:2].data_ptr() and b[0, 1][None].shape == (1, 5, 12, 16)
    assert lz[1:3].shape == (2, 5, 12, 16) and lz[-1].feats.data_ptr() == f[2:3].data_ptr()
    with pytest.raises(IndexError):
        b[1]
    with pytest.raises(IndexError):
        lz[3]
    with pytest.raises(IndexError):
        lz[::2]
    with pytest.raises(Pst3rError):
        b[None]
    with pytest.raises(Pst3rError):
        LazyMasks(f, e[:, :4])
    assert lz.to("cuda:0") is lz and lz.float() is lz and lz.contiguous() is lz  # no work, no copy


@pytest.mark.parametrize("H,hm,budget", [(384, 192, 20 << 20), (384, 192, 1 << 30), (384, 192, 1 << 20), (45, 24, 1 << 16),
                                         (512, 128, 3 << 20), (33, 33, 0)])
def test_band_plans_cover_the_map_and_hold_every_source_row(H, hm, budget):
    """Bands tile [0, H) in multiples of 32 rows; each band's source range contains every row the kernel's fp32 index
    arithmetic (csrc/postprocess.cu src_index, PyTorch's align_corners=False rule) reads for its output rows."""
    from panst3r_b200.postprocess import LazyMasks
    wm, Q = 256, 200
    lz = LazyMasks(torch.zeros(1, hm, wm, 8, dtype=torch.bfloat16), torch.zeros(Q, 8, dtype=torch.bfloat16))
    plan = lz.band_plan(H, budget)
    assert plan[0][0] == 0 and sum(p[1] for p in plan) == H
    scale = np.float32(hm) / np.float32(H)
    y = 0
    for y0, rows, s0, sr in plan:
        assert y0 == y and rows > 0 and (y0 % 32) == 0 and 0 <= s0 and s0 + sr <= hm
        y += rows
        for d in (y0, y0 + rows - 1):
            s = max(np.float32(scale * (np.float32(d) + np.float32(0.5)) - np.float32(0.5)), np.float32(0))
            i0 = min(int(s), hm - 1)
            i1 = min(i0 + 1, hm - 1)
            assert s0 <= i0 and i1 < s0 + sr
        if len(plan) > 1 and budget >= (36 * Q * wm * 4):
            assert sr * Q * wm * 4 <= budget
    if budget >= hm * Q * wm * 4 + 4 * Q * wm * 4:
        assert plan == [(0, H, 0, hm)]


def test_band_entry_point_rejects_bad_bands_without_a_gpu():
    from panst3r_b200 import lib as L
    lib = L.load()
    f = ctypes.c_float()
    i = ctypes.c_int32()
    pf, pi = ctypes.addressof(f), ctypes.addressof(i)

    def call(src_row0, src_rows, y0, rows, qs=96 * 128):
        return lib.pst3r_panoptic_argmax_band(pf, 0, qs, 1, 96, 128, src_row0, src_rows, pi, pf, 1, 192, 256, y0, rows, 0.25,
                                              pi, pf, 192 * 256, 256, pi, pi, None)
    assert call(0, 96, 0, 0) == -1 and b"bad band" in lib.pst3r_last_error()
    assert call(0, 97, 0, 192) == -1
    assert call(0, 16, 0, 64) == -1 and b"read source rows" in lib.pst3r_last_error()  # rows 0..63 read source rows 0..32
    assert call(20, 40, 32, 64) == -1                                                   # rows 32..95 start at source row 15
    assert call(0, 48, 0, 64, qs=47 * 128) == -1 and b"pitch" in lib.pst3r_last_error()


def test_forward_panoptic_host_logic_without_a_gpu():
    """Unknown post-processing names are refused before any work; on a CPU module the forward refuses (no CPU fallback) and
    the `lazy_masks` switch is restored; a batch of lazy handles is refused by the head."""
    from panst3r_b200.lib import Pst3rError
    from panst3r_b200.panst3r import build_panst3r
    m = build_panst3r("v1", 1, 1, 1)
    imgs, ts = torch.zeros(1, 2, 3, 32, 48), torch.tensor([[[32, 48]] * 2])
    with pytest.raises(NotImplementedError):
        m.forward_panoptic(imgs, ts, ["a"], postprocess="qubo")
    assert m.postprocess_default == "standard_v2" and m.panoptic_decoder.lazy_masks is False
    with pytest.raises(Pst3rError):
        m.forward_panoptic(imgs, ts, ["a"])
    assert m.panoptic_decoder.lazy_masks is False
    m.panoptic_decoder.lazy_masks = True
    with pytest.raises(Pst3rError, match="one scene per call"):
        m.panoptic_decoder(None, torch.zeros(2, 2, 3, 32, 48), None, torch.tensor([[[32, 48]] * 2] * 2), ["a"],
                           cat_feats=torch.zeros(2, 2, 6, 2816, dtype=torch.bfloat16))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("H,W,hm,wm", [(192, 256, 96, 128), (90, 140, 48, 64), (64, 96, 64, 96)])
def test_band_launches_equal_the_whole_map_launch(H, W, hm, wm):
    from panst3r_b200 import ops
    torch.manual_seed(1)
    V, Q = 2, 24
    masks = torch.randn(V, Q, hm, wm, device="cuda") * 3
    keep = torch.arange(0, Q, 2, device="cuda", dtype=torch.int32)
    sc = torch.rand(keep.numel(), device="cuda")
    a = torch.zeros(2, keep.numel(), device="cuda", dtype=torch.int32)
    ids, win = ops.panoptic_argmax(masks, keep, sc, (H, W), 0.25, a[0], a[1])
    b = torch.zeros_like(a)
    ids2 = torch.full_like(ids, -7)
    win2 = torch.full_like(win, -7.0)
    scale = hm / H
    for y0 in range(0, H, 32):
        rows = min(32, H - y0)
        lo = max(int(np.floor(scale * (y0 + 0.5) - 0.5)) - 1, 0)
        hi = min(int(np.floor(scale * (y0 + rows - 0.5) - 0.5)) + 2, hm - 1)
        for v in range(V):  # a band of ONE view in its own dense buffer, as LazyMasks produces it
            chunk = masks[v:v + 1, :, lo:hi + 1].contiguous()
            ops.panoptic_argmax(chunk, keep, sc, (H, W), 0.25, b[0], b[1], out=(ids2[v:v + 1], win2[v:v + 1]),
                                band=(y0, rows, lo, hm))
    assert torch.equal(ids, ids2) and torch.equal(win, win2) and torch.equal(a, b)
    with pytest.raises(ops._l.Pst3rError):  # a band that lacks a source row it reads
        ops.panoptic_argmax(masks[:1, :, 8:16].contiguous(), keep, sc, (H, W), 0.25, b[0], b[1],
                            out=(ids2[:1], win2[:1]), band=(0, 32, 8, hm))


def _features(V, hm, wm, C, Q, split, seed=0, regions=False):
    """Pixel features / mask embeddings (bf16 tensors or split pairs) + the fp32 values they represent.
    regions: every 16 x 16 block of pixels is owned by one query (logit ~ +4 for the owner, ~ -4 +- 0.5 for the others),
    so that segments survive the filtering rule of both post-processing rounds."""
    from panst3r_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    f = torch.randn(V, hm, wm, C, device="cuda", generator=g) * 0.5
    e = torch.randn(Q, C, device="cuda", generator=g) * 0.5
    if regions:
        e[:, 0] = 0
        ehat = e / (e * e).sum(1, keepdim=True)
        yy, xx = torch.meshgrid(torch.arange(hm, device="cuda"), torch.arange(wm, device="cuda"), indexing="ij")
        owner = ((yy // 16) * (wm // 16) + xx // 16)[None] + 5 * torch.arange(V, device="cuda")[:, None, None]
        f = 8.0 * ehat[owner % Q] + 0.04 * f
        f[..., 0] = 1.0
        e[:, 0] = -4.0
    if split:
        return ops.Split.from_float(f), ops.Split.from_float(e), f, e
    fb, eb = ops.to_bf16(f), ops.to_bf16(e)
    return fb, eb, fb.float(), eb.float()


def _class_logits(Q, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    cls = torch.randn(1, Q, K, generator=g) - 3.0
    for q in range(0, Q, 2):  # every other query carries a confident class
        cls[0, q, q % K] = 1.0 + 3.0 * torch.rand(1, generator=g).item()
    return cls.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("split", [False, True])
def test_lazy_logits_equal_the_mask_gemm(split):
    """materialize() is the eager head's launch; band launches reproduce its planes (same kernel, same k order)."""
    from panst3r_b200 import ops
    from panst3r_b200.postprocess import LazyMasks
    V, hm, wm, C, Q = 3, 96, 128, 256, 200
    f, e, f32, e32 = _features(V, hm, wm, C, Q, split)
    lz = LazyMasks(f, e)
    full = lz.materialize()
    ref = torch.einsum("qc,vhwc->vqhw", e32.double(), f32.double()).float()
    assert (full - ref).abs().max() / ref.abs().max() < (2e-5 if split else 1e-5)  # operands are exact in both modes
    mk = torch.empty((V, Q, hm, wm), device="cuda", dtype=torch.float32)
    ops.gemm(f.view(V * hm * wm, C), e, out=mk, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=hm * wm,
             batch_stride=Q * hm * wm, ldt=hm * wm)
    assert torch.equal(mk, full) and lz[None].materialize().shape == (1, V, Q, hm, wm) and torch.equal(lz[1].materialize(), full[1])
    for v, r0, r1 in [(0, 0, 96), (1, 17, 50), (2, 64, 96)]:
        part = lz.logits(v, r0, r1)
        d = (part - full[v, :, r0:r1]).abs().max().item()
        print(f"[lazy] split={split} view {v} rows [{r0}, {r1}): max |band - whole| = {d:.3e}")
        assert d <= 2e-6 * ref.abs().max().item()
    assert torch.equal(lz.to("cpu"), full.cpu())


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["v2", "v1"])
@pytest.mark.parametrize("split", [False, True])
def test_lazy_postprocess_equals_materialised(which, split):
    from panst3r_b200 import postprocess as pp
    V, hm, wm, C, Q, K = 3, 96, 128, 256, 60, 7
    f, e, f32, e32 = _features(V, hm, wm, C, Q, split, seed=3, regions=True)
    cls = _class_logits(Q, K, 5)
    lz = pp.LazyMasks(f, e)
    full = lz.materialize()
    fn = pp.panoptic_inference_v2 if which == "v2" else pp.panoptic_inference_v1
    kw = dict(cls_threshold=0.05)
    for scratch in (None, 1 << 20):  # default plan, and a plan of many small bands
        old = pp.LazyMasks.scratch_bytes
        if scratch:
            pp.LazyMasks.scratch_bytes = scratch
        try:
            a = fn(cls, lz[None], (2 * hm, 2 * wm), **kw)[0]
        finally:
            pp.LazyMasks.scratch_bytes = old
        b = fn(cls, full[None], (2 * hm, 2 * wm), **kw)[0]
        assert a["segments_info"] == b["segments_info"]
        same = (a["pan"] == b["pan"]).float().mean().item()
        print(f"[lazy] {which} split={split} scratch={scratch}: {len(a['segments_info'])} segments, identical ids {same:.6f}")
        assert len(a["segments_info"]) >= 10 and int(a["pan"].max()) == len(a["segments_info"])
        assert same > 0.9999 and (a["conf"] - b["conf"]).abs().max() < 1e-5
    # multi_ar form: one lazy handle per view, as forward_inference_multi_ar returns them
    ts = np.array([[2 * hm, 2 * wm]] * V)
    a = fn(cls, [lz[None][0, i] for i in range(V)], ts, multi_ar=True, **kw)[0]
    b = fn(cls, [full[i] for i in range(V)], ts, multi_ar=True, **kw)[0]
    assert a["segments_info"] == b["segments_info"]
    assert all((x == y).float().mean() > 0.9999 for x, y in zip(a["pan"], b["pan"]))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_panoptic_equals_forward_plus_postprocess(precision):
    """Free-running model, 3 views of 64x96: PanSt3R.forward_panoptic (lazy masks, no auxiliary heads) returns the class
    logits, queries, pointmaps and panoptic results of forward() + panoptic_inference_v2 on the materialised masks."""
    from panst3r_b200 import postprocess as pp
    from panst3r_b200.panst3r import build_panst3r
    from panst3r_b200.postprocess import LazyMasks
    import bench
    with torch.device("cuda"):
        m = build_panst3r("v1", 2, 2, 2, head_precision=precision)
    bench.init_weights_(m)
    classes = [f"c{i}" for i in range(6)]
    g = torch.Generator().manual_seed(1)
    m.panoptic_decoder.text_encoder.class_embeddings = {c: torch.randn(768, generator=g) for c in classes}
    V, H, W = 3, 64, 96
    imgs = (torch.rand(1, V, 3, H, W, generator=g) * 2 - 1).cuda()
    ts = torch.tensor([[[H, W]] * V])
    panout, pm = m(imgs, ts, classes)
    kw = dict(cls_threshold=0.0)  # random weights: keep every query so that the argmax has work to do
    ref = pp.panoptic_inference_v2(panout["pred_logits"], panout["pred_masks"], (H, W), **kw)[0]
    res, lazy_out, pm2 = m.forward_panoptic(imgs, ts, classes, postprocess="standard_v2", **kw)
    assert m.panoptic_decoder.lazy_masks is False  # restored
    lz = lazy_out["pred_masks"]
    assert isinstance(lz, LazyMasks) and lz.shape == panout["pred_masks"].shape and lazy_out["aux_outputs"] == []
    assert torch.equal(pm, pm2) and torch.equal(lazy_out["pred_logits"], panout["pred_logits"])
    assert torch.equal(lazy_out["out_queries"], panout["out_queries"])
    d = (lz.materialize() - panout["pred_masks"]).abs().max().item()
    print(f"[lazy] {precision}: max |lazy.materialize() - pred_masks| = {d:.3e}")
    assert d <= 1e-6 * panout["pred_masks"].abs().max().item()
    assert res[0]["segments_info"] == ref["segments_info"]
    assert (res[0]["pan"] == ref["pan"]).float().mean() > 0.999 and (res[0]["conf"] - ref["conf"]).abs().max() < 1e-5
    with pytest.raises(NotImplementedError):
        m.forward_panoptic(imgs, ts, classes, postprocess="qubo")


@pytest.mark.gpu
def test_lazy_masks_through_forward_inference_multi_ar():
    """Mixed shapes / a portrait view, 4 keyframes + 2 render-only frames (the memory-query path): with
    `PanopticDecoder.lazy_masks` every entry of `pred_masks` is a one-view handle whose logits are the eager path's, and the
    reference-style call `panoptic_inference_v2(..., size, multi_ar=True)` (tools/demo_panst3r.py:236-242) returns the same
    segments and ids as on the materialised list."""
    import bench
    from panst3r_b200 import postprocess as pp
    from panst3r_b200.panst3r import build_panst3r
    with torch.device("cuda"):
        m = build_panst3r("v1", 2, 2, 2)
    bench.init_weights_(m)
    classes = [f"c{i}" for i in range(6)]
    g = torch.Generator().manual_seed(11)
    m.panoptic_decoder.text_encoder.class_embeddings = {c: torch.randn(768, generator=g) for c in classes}
    shapes = [(64, 96), (64, 96), (64, 64), (64, 96), (64, 96), (64, 64)]
    ts = torch.tensor([[64, 96], [64, 96], [64, 64], [96, 64], [64, 96], [64, 64]])  # view 3: portrait, stored transposed
    imgs = [(torch.rand(3, *sh, generator=g) * 2 - 1).cuda() for sh in shapes]
    pms, pan = m.forward_inference_multi_ar(imgs, ts, classes, num_keyframes=4)
    m.panoptic_decoder.lazy_masks = True
    pms_l, pan_l = m.forward_inference_multi_ar(imgs, ts, classes, num_keyframes=4)
    m.panoptic_decoder.lazy_masks = False
    assert all(torch.equal(a, b) for a, b in zip(pms, pms_l)) and torch.equal(pan["pred_logits"], pan_l["pred_logits"])
    for a, b in zip(pan["pred_masks"], pan_l["pred_masks"]):
        assert isinstance(b, pp.LazyMasks) and b.shape == a.shape
        assert (b.materialize() - a).abs().max() <= 1e-6 * a.abs().max()
    kw = dict(cls_threshold=0.0)
    size = ts.numpy()
    r0 = pp.panoptic_inference_v2(pan["pred_logits"], pan["pred_masks"], size, multi_ar=True, **kw)[0]
    r1 = pp.panoptic_inference_v2(pan_l["pred_logits"], pan_l["pred_masks"], size, multi_ar=True, **kw)[0]
    assert r0["segments_info"] == r1["segments_info"]
    for a, b, t in zip(r0["pan"], r1["pan"], ts.tolist()):
        assert tuple(a.shape) == tuple(t) and (a == b).float().mean() > 0.999
    # leaving the GPU materialises the handle (outdevice='cpu', the reference demo's pattern)
    m.panoptic_decoder.lazy_masks = True
    _, pan_c = m.forward_inference_multi_ar(imgs[:3], ts[:3], classes, outdevice="cpu")
    m.panoptic_decoder.lazy_masks = False
    assert all(torch.is_tensor(x) and x.device.type == "cpu" for x in pan_c["pred_masks"])


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_aux_masks_on_the_side_stream_are_bit_identical(precision):
    """MaskTransformer.overlap_aux_masks: the auxiliary heads' full-resolution mask GEMMs run on a side stream (64 SMs)
    next to the following decoder layer.  Same kernels, same operands: every output equals the in-line run bit for bit,
    eagerly and from a replayed CUDA graph, at 512x384 (the GEMMs are long enough to overlap the whole next layer)."""
    import bench
    from panst3r_b200.panst3r import build_panst3r
    with torch.device("cuda"):
        m = build_panst3r("v1", 2, 2, 2, head_precision=precision)
    bench.init_weights_(m)
    classes = bench.CLASSES[:12]
    g = torch.Generator().manual_seed(3)
    m.panoptic_decoder.text_encoder.class_embeddings = {c: torch.randn(768, generator=g) for c in classes}
    imgs, ts = bench.make_inputs(4, "cuda")
    imgs = imgs.cuda()
    mt = m.panoptic_decoder.mask_transformer

    def flat(o):
        pan, pm = o
        return [pm, pan["pred_logits"], pan["pred_masks"], pan["out_queries"]] + \
            [a[k] for a in pan["aux_outputs"] for k in ("pred_logits", "pred_masks")]
    mt.overlap_aux_masks = False
    ref = [t.clone() for t in flat(m(imgs, ts, classes))]
    assert len(ref) == 4 + 2 * 6
    mt.overlap_aux_masks = True
    for _ in range(3):
        out = flat(m(imgs, ts, classes))
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(out, ref))
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        gout = m(imgs, ts, classes)
    for _ in range(3):
        gr.replay()
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(flat(gout), ref))


@pytest.mark.gpu
def test_lazy_postprocess_full_size():
    """16 views x 200 queries at 512x384 (BASELINE config 2): lazy == materialised on every pixel; prints both timings."""
    from panst3r_b200 import postprocess as pp
    V, hm, wm, C, Q, K = 16, 192, 256, 256, 200, 20
    f, e, _, _ = _features(V, hm, wm, C, Q, True, seed=7, regions=True)
    cls = _class_logits(Q, K, 9)
    lz = pp.LazyMasks(f, e)
    full = lz.materialize()
    b = pp.panoptic_inference_v2(cls, full[None], (384, 512))[0]
    a = pp.panoptic_inference_v2(cls, lz[None], (384, 512))[0]
    assert len(a["segments_info"]) >= 50
    same = (a["pan"] == b["pan"]).float().mean().item()
    print(f"[lazy] full size: {len(a['segments_info'])} segments, identical ids {same:.7f}, "
          f"bitwise {torch.equal(a['pan'], b['pan']) and torch.equal(a['conf'], b['conf'])}")
    assert a["segments_info"] == b["segments_info"] and same > 0.99999 and (a["conf"] - b["conf"]).abs().max() < 1e-5

    def timed(fn, n=3):
        fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / n
    t_mat = timed(lambda: pp.panoptic_inference_v2(cls, lz.materialize()[None], (384, 512)))
    t_lazy = timed(lambda: pp.panoptic_inference_v2(cls, lz[None], (384, 512)))
    print(f"[lazy] 16 x 200 x 192 x 256: mask GEMM + post-processing {t_mat:.2f} ms materialised, {t_lazy:.2f} ms lazy")
