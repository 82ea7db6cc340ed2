"""Oracle PanSt3R façade (TEST INFRASTRUCTURE ONLY): the reference's orchestration restated over oracle modules.

Follows /root/reference/src/panst3r/panst3r.py:47-86 (forward_dino / forward_must3r_encoder /
forward_must3r_decoder), :169-284 (forward_inference_multi_ar, single aspect ratio, linspace keyframes) and
:286-296 (forward); engine/must3r.py:28-69 (memory build loop) and :71-94 (render, unsliced branch).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .must3r import Dust3rEncoder, MUSt3R
from .panoptic import DinoV2Encoder, InputMixer, LoftUpUpscaler, PanopticDecoder, PixelShuffleUpscaler


class PanSt3R(nn.Module):
    def __init__(self, must3r_encoder, must3r_decoder, dino_encoder, panoptic_decoder):
        super().__init__()
        self.must3r_encoder = must3r_encoder
        self.must3r_decoder = must3r_decoder
        self.dino_encoder = dino_encoder
        self.panoptic_decoder = panoptic_decoder
        self.must3r_params = dict(init_num_views=2, batch_num_views=1, render_iterations=1)

    def get_must3r_mem_batches(self, n_imgs):  # panst3r.py:65-70
        mem_batches = [self.must3r_params["init_num_views"]]
        while (s := sum(mem_batches)) != n_imgs:
            mem_batches.append(min(self.must3r_params["batch_num_views"], n_imgs - s))
        return mem_batches

    def forward_dino(self, imgs, true_shape):
        B, V = imgs.shape[:2]
        return self.dino_encoder(imgs.flatten(0, 1), true_shape.flatten(0, 1)).unflatten(0, (B, V))

    def forward_must3r_encoder(self, imgs, true_shape):
        B, V = imgs.shape[:2]
        x, pos = self.must3r_encoder(imgs.flatten(0, 1), true_shape.flatten(0, 1))
        return x.unflatten(0, (B, V)), pos.unflatten(0, (B, V))

    def build_memory(self, x, pos, true_shape):  # engine/must3r.py:28-69
        edges = [0] + np.cumsum(self.get_must3r_mem_batches(x.shape[1])).tolist()
        mem = None
        for a, b in zip(edges[:-1], edges[1:]):
            mem, _, _ = self.must3r_decoder(x[:, a:b].contiguous(), pos[:, a:b].contiguous(),
                                            true_shape[:, a:b].contiguous(), mem, render=False, return_feats=True)
        return mem

    def render(self, x, pos, true_shape, mem):  # engine/must3r.py:71-94
        _, pointmaps, feats = self.must3r_decoder(x, pos, true_shape, mem, render=True, return_feats=True)
        return pointmaps, feats[-1]

    @torch.no_grad()
    def forward(self, imgs, true_shape, classes, max_bs=None, outdevice=None, amp=False):
        """amp=True: the reference's inference precision policy (panst3r.py:174, 204, 236-245): DINOv2, encoder and
        decoder under torch.autocast(bf16), the panoptic head in fp32.  amp=False: fp32 everywhere."""
        with torch.autocast(imgs.device.type, dtype=torch.bfloat16, enabled=bool(amp)):
            x_dino = self.forward_dino(imgs, true_shape)
            x, pos = self.forward_must3r_encoder(imgs, true_shape)
            mem = self.build_memory(x, pos, true_shape)
            pointmaps, y = self.render(x, pos, true_shape, mem)
        panout = self.panoptic_decoder((x.float(), y.float(), x_dino.float()), imgs, pos, true_shape, classes)
        return panout, pointmaps.float()

    @torch.no_grad()
    def forward_inference_multi_ar(self, imgs, true_shape, classes, num_keyframes=None, max_bs=None, outdevice=None):
        """imgs: list of (3,H,W) tensors; true_shape (N,2).  Returns (pointmaps list, panout).  Views of different
        shape / orientation go through the per-view restatement below (panst3r.py:169-284 with must3r's stack_views)."""
        if len({(tuple(t), tuple(im.shape[-2:])) for im, t in zip(imgs, true_shape.tolist())}) > 1:
            return self._forward_inference_mixed(imgs, true_shape, classes, num_keyframes)
        N = len(imgs)
        if num_keyframes is None or num_keyframes > N:
            num_keyframes = N
            keyframes = list(range(N))
        else:
            keyframes = np.linspace(0, N - 1, num_keyframes, dtype=int).tolist()
        not_keyframes = sorted(set(range(N)).difference(keyframes))
        order = keyframes + not_keyframes
        im = torch.stack([imgs[i] for i in order])[None]
        ts = true_shape[order][None]
        x, pos = self.forward_must3r_encoder(im, ts)
        k = num_keyframes
        mem = self.build_memory(x[:, :k], pos[:, :k], ts[:, :k])
        pm_kf, y_kf = self.render(x[:, :k], pos[:, :k], ts[:, :k], mem)
        d_kf = self.forward_dino(im[:, :k], ts[:, :k])
        pan = self.panoptic_decoder((x[:, :k], y_kf, d_kf), im[:, :k], pos[:, :k], ts[:, :k], classes)
        pointmaps = list(pm_kf[0])
        masks = list(pan["pred_masks"][0])
        for i in range(k, N):  # render-only frames reuse the keyframes' final queries (panst3r.py:127-167)
            sl = slice(i, i + 1)
            pm_i, y_i = self.render(x[:, sl], pos[:, sl], ts[:, sl], mem)
            d_i = self.forward_dino(im[:, sl], ts[:, sl])
            out_i = self.panoptic_decoder((x[:, sl], y_i, d_i), im[:, sl], pos[:, sl], ts[:, sl], classes,
                                          memory_queries=pan["out_queries"])
            pointmaps.append(pm_i[0, 0])
            masks.append(out_i["pred_masks"][0, 0])
        inv = np.argsort(order)
        return [pointmaps[i] for i in inv], {"pred_logits": pan["pred_logits"], "pred_masks": [masks[i] for i in inv],
                                              "out_queries": pan["out_queries"]}


def _mixed(self, imgs, true_shape, classes, num_keyframes=None):
    """Mixed aspect ratios, view by view: encoder / DINOv2 / render are per-view independent; the memory build walks the
    keyframes in order (mem_batches [2, 1, 1, ...]); the head sees one stack per distinct shape (keyframes first)."""
    N = len(imgs)
    if num_keyframes is None or num_keyframes > N:
        num_keyframes, keyframes = N, list(range(N))
    else:
        keyframes = np.linspace(0, N - 1, num_keyframes, dtype=int).tolist()
    order = keyframes + sorted(set(range(N)).difference(keyframes))
    k = num_keyframes
    im = [imgs[i][None, None] for i in order]                      # (1, 1, 3, H, W) each
    ts = [true_shape[i][None, None] for i in order]                # (1, 1, 2)
    enc = [self.forward_must3r_encoder(a, t) for a, t in zip(im, ts)]
    x, pos = [e[0] for e in enc], [e[1] for e in enc]
    edges = [0] + np.cumsum(self.get_must3r_mem_batches(k)).tolist()
    mem = None
    for a, b in zip(edges[:-1], edges[1:]):
        mem, _, _ = self.must3r_decoder(torch.cat(x[a:b], 1), torch.cat(pos[a:b], 1), torch.cat(ts[a:b], 1), mem,
                                        render=False, return_feats=True)
    ren = [self.render(x[p], pos[p], ts[p], mem) for p in range(N)]
    dino = [self.forward_dino(im[p], ts[p]) for p in range(N)]
    stacks = {}
    for p in range(N):
        stacks.setdefault((tuple(ts[p].flatten().tolist()), tuple(im[p].shape[-2:])), []).append(p)
    stacks = list(stacks.values())

    def head(sel, queries=None):
        groups = [[p for p in idx if sel(p)] for idx in stacks]
        groups = [g_ for g_ in groups if g_]
        cat1 = lambda lst, g_: torch.cat([lst[p] for p in g_], 1)  # noqa: E731
        feats = tuple([cat1(src, g_) for g_ in groups] for src in (x, [r[1] for r in ren], dino))
        out = self.panoptic_decoder(feats, [cat1(im, g_) for g_ in groups], [cat1(pos, g_) for g_ in groups],
                                    [cat1(ts, g_) for g_ in groups], classes, multi_ar=True, memory_queries=queries)
        return groups, out

    gk, pan = head(lambda p: p < k)
    masks = [None] * N
    for g_, mk in zip(gk, pan["pred_masks"]):
        for j, p in enumerate(g_):
            masks[p] = mk[0, j]
    if k < N:
        gn, out_n = head(lambda p: p >= k, pan["out_queries"])
        for g_, mk in zip(gn, out_n["pred_masks"]):
            for j, p in enumerate(g_):
                masks[p] = mk[0, j]
    inv = np.argsort(order)
    return [ren[i][0][0, 0] for i in inv], {"pred_logits": pan["pred_logits"], "pred_masks": [masks[i] for i in inv],
                                            "out_queries": pan["out_queries"]}


PanSt3R._forward_inference_mixed = torch.no_grad()(_mixed)


def build_panst3r(variant: str = "v1", enc_depth=24, dec_depth=12, dino_depth=24, mixer_layers=3) -> PanSt3R:
    """Full-size widths (ViT-L encoder 1024/16 heads, decoder 768/12 heads, DINOv2-L); depths reducible for tests."""
    enc = Dust3rEncoder(depth=enc_depth)
    dec = MUSt3R(depth=dec_depth)
    dino = DinoV2Encoder(depth=dino_depth)
    if variant == "v1":
        pd = PanopticDecoder(upscaler=PixelShuffleUpscaler(input_dim=2816))
    elif variant == "v2":
        pd = PanopticDecoder(input_mixer=InputMixer([512, 512], 16, 2816, 768, num_layers=mixer_layers),
                             upscaler=LoftUpUpscaler(input_dim=768, dim=384), mask_dim=384)
    else:
        raise ValueError(variant)
    return PanSt3R(enc, dec, dino, pd).eval()
