"""Closed-form known-answer tests for the restated upstream pieces (parity-unpinned part of the oracle) and the
identities the CUDA design relies on."""
import math

import torch
import torch.nn.functional as F

from oracle import weights as W
from oracle.blocks import RoPE2D
from oracle.must3r import MUSt3R, Dust3rEncoder
from oracle.panoptic import sine_position_embedding
from oracle.panst3r import build_panst3r


def grid_pos(h, w, B=1):
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    return torch.stack([ys.flatten(), xs.flatten()], -1)[None].expand(B, -1, -1).contiguous()


def test_rope2d_closed_form():
    torch.manual_seed(0)
    rope = RoPE2D(100.0)
    t = torch.randn(2, 3, 12, 64)
    pos = grid_pos(3, 4, 2)
    out = rope(t, pos)
    # norm preserving rotation, identity at position (0, 0)
    assert torch.allclose(out.norm(dim=-1), t.norm(dim=-1), atol=1e-5)
    assert torch.allclose(out[:, :, 0], t[:, :, 0], atol=1e-6)
    # explicit pair formula: D=64 -> halves of 32, pairs (j, j+16), angle p * 100^(-j/16)
    b, h, n, j = 1, 2, 7, 5
    y, x = pos[b, n].tolist()
    for half, p in ((0, y), (1, x)):
        th = p * 100.0 ** (-j / 16)
        u, v = t[b, h, n, half * 32 + j], t[b, h, n, half * 32 + j + 16]
        assert abs(out[b, h, n, half * 32 + j] - (u * math.cos(th) - v * math.sin(th))) < 1e-5
        assert abs(out[b, h, n, half * 32 + j + 16] - (v * math.cos(th) + u * math.sin(th))) < 1e-5
    # relative-position property: <rope(q,p1), rope(k,p2)> depends only on p1 - p2
    q, k = torch.randn(1, 1, 1, 64), torch.randn(1, 1, 1, 64)
    def dot(p1, p2):
        a = rope(q, torch.tensor([[p1]]))
        c = rope(k, torch.tensor([[p2]]))
        return (a * c).sum().item()
    assert abs(dot([3, 5], [1, 2]) - dot([7, 9], [5, 6])) < 1e-4


def test_bilinear_down8_is_centre_2x2_mean():
    torch.manual_seed(1)
    x = torch.randn(2, 5, 16, 24)
    ref = F.interpolate(x, size=(2, 3), mode="bilinear", align_corners=False)
    mine = x.view(2, 5, 2, 8, 3, 8)[:, :, :, 3:5, :, 3:5].mean(dim=(3, 5))
    assert torch.equal(ref, mine) or torch.allclose(ref, mine, atol=1e-7)


def test_pixel_shuffle_index_identity():
    x = torch.arange(2 * 8 * 3 * 5, dtype=torch.float32).view(2, 8, 3, 5)
    out = F.pixel_shuffle(x, 2)
    for b, c, y, xx, i, j in [(0, 0, 0, 0, 0, 0), (1, 1, 2, 4, 1, 0), (0, 1, 1, 3, 1, 1)]:
        assert out[b, c, 2 * y + i, 2 * xx + j] == x[b, 4 * c + 2 * i + j, y, xx]


def test_sine_pe_layout():
    pe = sine_position_embedding(3, 4, 384, "cpu")
    assert pe.shape == (768, 3, 4) and pe.abs().max() <= 1.0
    y1 = 1 / (3 + 1e-6) * 2 * math.pi
    assert abs(pe[0, 0, 0] - math.sin(y1)) < 1e-6 and abs(pe[1, 0, 0] - math.cos(y1)) < 1e-6
    x2 = 2 / (4 + 1e-6) * 2 * math.pi
    assert abs(pe[384, 0, 1] - math.sin(x2)) < 1e-6
    assert abs(pe[384 + 3, 0, 1] - math.cos(x2 / 10000 ** (2 / 384))) < 1e-6


def small_decoder():
    torch.manual_seed(0)
    enc = W.load_synth(Dust3rEncoder(depth=1), seed=2).eval()
    dec = W.load_synth(MUSt3R(depth=2), seed=2).eval()
    return enc, dec


def test_restated_decoder_memory_semantics():
    enc, dec = small_decoder()
    V, H, Wd = 3, 32, 48
    imgs = torch.rand(V, 3, H, Wd) * 2 - 1
    ts = torch.tensor([[H, Wd]] * V)
    with torch.no_grad():
        x, pos = enc(imgs, ts)
        x, pos, ts = x[None], pos[None], ts[None]
        mem, pm, feats = dec(x[:, :2], pos[:, :2], ts[:, :2], None, render=False, return_feats=True)
        N = x.shape[2]
        assert len(mem[0]) == 2 and mem[0][0].shape == (1, 2 * N, 768) and mem[2] == 2
        assert mem[1].tolist() == [[0] * N + [1] * N]
        assert pm.shape == (1, 2, H, Wd, 7) and feats[-1].shape == (1, 2, N, 768)
        mem2, _, _ = dec(x[:, 2:3], pos[:, 2:3], ts[:, 2:3], mem, render=False, return_feats=True)
        assert mem2[0][1].shape == (1, 3 * N, 768) and torch.equal(mem2[0][0][:, :2 * N], mem[0][0])
        # render leaves the memory untouched and is per-view independent (chunk invariant)
        mem3, pm_all, f_all = dec(x, pos, ts, mem2, render=True, return_feats=True)
        assert mem3 is mem2
        _, pm_1, f_1 = dec(x[:, 1:2], pos[:, 1:2], ts[:, 1:2], mem2, render=True, return_feats=True)
        assert torch.allclose(pm_all[:, 1:2], pm_1, atol=1e-5) and torch.allclose(f_all[-1][:, 1:2], f_1[-1], atol=1e-5)
        # during the 2-view initialisation a view attends only to the OTHER view: perturbing view 1's input
        # changes view 0's output, while view 0's own memory write is excluded from its candidates
        x2 = x.clone()
        x2[:, 1] += 1.0
        _, _, fb = dec(x2[:, :2], pos[:, :2], ts[:, :2], None, render=False, return_feats=True)
        assert not torch.allclose(fb[-1][:, 0], feats[-1][:, 0])


def test_facade_paths_agree():
    m = build_panst3r("v1", 1, 1, 1)
    W.load_synth(m, seed=4)
    classes = ["a", "b", "c"]
    m.panoptic_decoder.text_encoder.class_embeddings = W.synth_class_embeddings(classes)
    assert m.get_must3r_mem_batches(5) == [2, 1, 1, 1] and m.get_must3r_mem_batches(2) == [2]
    V, H, Wd = 3, 32, 48
    imgs = torch.rand(1, V, 3, H, Wd) * 2 - 1
    ts = torch.tensor([[[H, Wd]] * V])
    pan, pm = m(imgs, ts, classes)
    pms, pan2 = m.forward_inference_multi_ar(list(imgs[0]), ts[0], classes)  # all keyframes
    assert torch.allclose(torch.stack(pms), pm[0], atol=1e-5)
    assert torch.allclose(torch.stack(pan2["pred_masks"]), pan["pred_masks"][0], atol=1e-4)
    # keyframe subset: order is restored, render-only frames reuse the keyframes' queries
    pms3, pan3 = m.forward_inference_multi_ar(list(imgs[0]), ts[0], classes, num_keyframes=2)
    assert len(pms3) == V and pan3["pred_masks"][1].shape == pan["pred_masks"][0, 1].shape


def test_portrait_convention_identities():
    """Portrait views are stored transposed (the reference's batch convention, model/dino.py:25-33, utils.py:36-49):
    every stage must see the picture in its true orientation, dense outputs come back in the storage layout."""
    torch.manual_seed(0)
    m = build_panst3r("v1", 1, 1, 1)
    m.load_state_dict(W.synth_state_dict(m, seed=3))
    classes = ["a", "b", "c"]
    m.panoptic_decoder.text_encoder.class_embeddings = W.synth_class_embeddings(classes)
    g = torch.Generator().manual_seed(2)
    stored = torch.rand(1, 2, 3, 32, 48, generator=g) * 2 - 1          # landscape storage
    ts_p = torch.tensor([[[48, 32]] * 2])                               # true shape: portrait
    true = stored.transpose(-1, -2).contiguous()                        # the pictures as they really are (48 x 32)
    ts_t = torch.tensor([[[32, 48]] * 2])                               # "no transposition needed" flag for `true`
    with torch.no_grad():
        # encoder / DINOv2: tokens of the stored-transposed batch == tokens of the pictures embedded as they are
        xa, pa = m.forward_must3r_encoder(stored, ts_p)
        xb, pb = m.forward_must3r_encoder(true, ts_t)
        assert torch.equal(pa, pb) and torch.allclose(xa, xb, atol=1e-5)
        assert torch.allclose(m.forward_dino(stored, ts_p), m.forward_dino(true, ts_t), atol=1e-4)
        pan, pm = m(stored, ts_p, classes)
    assert pm.shape == (1, 2, 32, 48, 7) and pan["pred_masks"].shape == (1, 2, 200, 16, 24)  # storage layout
    with torch.no_grad():
        _, pm_l = m(stored, torch.tensor([[[32, 48]] * 2]), classes)
    assert (pm - pm_l).abs().max() > 1e-3  # and it is not the landscape interpretation of the same tensor


def test_multi_ar_facade_reduces_to_single_shape_path():
    """The per-view mixed-shape restatement equals the stacked single-shape one when all views share a shape."""
    torch.manual_seed(0)
    m = build_panst3r("v1", 1, 1, 1)
    m.load_state_dict(W.synth_state_dict(m, seed=3))
    classes = ["a", "b", "c"]
    m.panoptic_decoder.text_encoder.class_embeddings = W.synth_class_embeddings(classes)
    g = torch.Generator().manual_seed(4)
    imgs = [torch.rand(3, 32, 48, generator=g) * 2 - 1 for _ in range(4)]
    ts = torch.tensor([[32, 48]] * 4)
    with torch.no_grad():
        a = m.forward_inference_multi_ar(imgs, ts, classes, num_keyframes=2)
        b = m._forward_inference_mixed(imgs, ts, classes, 2)
    for x, y in zip(a[0] + a[1]["pred_masks"], b[0] + b[1]["pred_masks"]):
        assert x.shape == y.shape and torch.allclose(x, y, atol=1e-4, rtol=1e-4)
