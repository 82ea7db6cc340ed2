"""Oracle restatement of the reference's panoptic head in plain fp32 PyTorch (TEST INFRASTRUCTURE ONLY).

Follows, function by function (paths relative to /root/reference/src/panst3r):
  PixelShuffleUpscaler   model/upscalers/pixel_shuffle.py:9-59
  MinMaxScaler / ImplicitFeaturizer / LoftUpUpscaler   model/upscalers/loftup.py:9-190
  CrossonlyDecoderBlock  model/blocks.py:9-35
  InputMixer             model/input_mixer.py:8-29
  TextEncoder (fixed vocabulary only)   model/text_encoder.py:94-103
  PositionEmbeddingSine / MaskTransformer   model/mask_transformer.py:12-288, 487-527
  PanopticDecoder        model/panoptic_decoder.py:16-77
  DinoV2Encoder          model/dino.py:49-71
  batched_map chunking is the identity for batch_size=None (utils.py:156), so views are processed in one chunk.
Parameter names equal the reference's, so the same state dict loads into both (tests/test_oracle_vs_reference.py
checks outputs against the reference modules imported from /root/reference; tests/golden/ holds their outputs).
Only single-aspect-ratio, all-landscape or all-portrait batches are restated (what the benchmark uses).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .blocks import Block, CrossAttention, Mlp, get_pos_embed


# ----------------------------------------------------------------------------------------------- upscalers
class PixelShuffleUpscaler(nn.Module):
    def __init__(self, input_dim, patch_size=16, hidden_dim_factor=4, fp_dim=(768, 512, 384, 256)):
        super().__init__()
        self.patch_size = patch_size
        f = hidden_dim_factor
        self.proj_8 = Mlp(input_dim, int(f * input_dim), fp_dim[1] * 4)
        self.proj_4 = Mlp(fp_dim[1], int(f * fp_dim[1]), fp_dim[2] * 4)
        self.proj_2 = Mlp(fp_dim[2], int(f * fp_dim[2]), fp_dim[3] * 4)
        self.proj_16 = Mlp(input_dim, int(f * input_dim), fp_dim[0])

    @staticmethod
    def _to_map(tokens, h, w):  # (B, h*w, C) -> (B, C, h, w)
        return tokens.transpose(1, 2).reshape(tokens.shape[0], -1, h, w)

    def forward(self, inputs, img_shape):
        feats = inputs[0]
        H, W = img_shape
        hs, ws = H // self.patch_size, W // self.patch_size
        f8 = F.pixel_shuffle(self._to_map(self.proj_8(feats), hs, ws), 2)
        f4 = F.pixel_shuffle(self._to_map(self.proj_4(f8.flatten(2).transpose(1, 2)), 2 * hs, 2 * ws), 2)
        f2 = F.pixel_shuffle(self._to_map(self.proj_2(f4.flatten(2).transpose(1, 2)), 4 * hs, 4 * ws), 2)
        f16 = self._to_map(self.proj_16(feats), hs, ws)
        return [f16], f2


class MinMaxScaler(nn.Module):
    def forward(self, x):  # per channel over the WHOLE batch (loftup.py:14-19)
        lo = x.amin(dim=(0, 2, 3), keepdim=True)
        hi = x.amax(dim=(0, 2, 3), keepdim=True)
        return (x - lo) / (hi - lo).clamp_min(1e-4) - 0.5


class ImplicitFeaturizer(nn.Module):
    def __init__(self, color_feats=True, n_freqs=10, learn_bias=False):
        super().__init__()
        self.color_feats, self.n_freqs = color_feats, n_freqs
        self.dim_multiplier = 5 if color_feats else 2
        self.learn_bias = learn_bias
        if learn_bias:
            self.biases = nn.Parameter(torch.randn(2, self.dim_multiplier, n_freqs))

    def forward(self, img):
        b, _, h, w = img.shape
        gy = torch.linspace(-1, 1, h, device=img.device).view(1, 1, h, 1).expand(b, 1, h, w)
        gx = torch.linspace(-1, 1, w, device=img.device).view(1, 1, 1, w).expand(b, 1, h, w)
        base = torch.cat([gy, gx] + ([img] if self.color_feats else []), dim=1)  # (b, m, h, w)
        freqs = torch.exp(torch.linspace(-2, 10, self.n_freqs, device=img.device)).view(1, -1, 1, 1, 1)
        arg = base.unsqueeze(1) * freqs  # (b, n_freqs, m, h, w)
        if self.learn_bias:
            # NB the parameter is (2, m, n_freqs) but is *reshaped* (not transposed) to (n_freqs, m) (loftup.py:62-63)
            s_arg = arg + self.biases[0].reshape(1, self.n_freqs, self.dim_multiplier, 1, 1)
            c_arg = arg + self.biases[1].reshape(1, self.n_freqs, self.dim_multiplier, 1, 1)
        else:
            s_arg = c_arg = arg
        out = [torch.sin(s_arg).flatten(1, 2), torch.cos(c_arg).flatten(1, 2)]
        if self.color_feats:
            out.append(img)
        return torch.cat(out, dim=1)


class CrossonlyDecoderBlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, norm_mem=True):
        super().__init__()
        self.cross_attn = CrossAttention(dim, rope=None, num_heads=num_heads, qkv_bias=False)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.norm_y = nn.LayerNorm(dim) if norm_mem else nn.Identity()

    def forward(self, x, y, xpos=None, ypos=None):
        y_ = self.norm_y(y)
        x = x + self.cross_attn(self.norm2(x), y_, y_, xpos, ypos)
        x = x + self.mlp(self.norm3(x))
        return x, y


class LoftUpUpscaler(nn.Module):
    def __init__(self, input_dim, dim, output_stride=2, patch_size=16, color_feats=True, n_freqs=20, num_heads=4,
                 num_layers=2):
        super().__init__()
        self.output_stride, self.patch_size = output_stride, patch_size
        self.patch_embed = nn.Conv2d(input_dim, input_dim, kernel_size=1)
        start_dim = 5 * n_freqs * 2 + 3 if color_feats else 2 * n_freqs * 2
        self.lr_pe = ImplicitFeaturizer(color_feats=False, n_freqs=5, learn_bias=True)
        self.lr_input_proj = nn.Sequential(nn.Linear(input_dim + 20, dim), nn.LayerNorm(dim))
        self.fourier_feat = nn.Sequential(MinMaxScaler(), ImplicitFeaturizer(color_feats, n_freqs=n_freqs, learn_bias=True))
        self.first_conv = nn.Sequential(
            nn.GroupNorm(1, start_dim), nn.Conv2d(start_dim, dim, 3, padding=1), nn.GroupNorm(8, dim), nn.ReLU(),
            nn.Conv2d(dim, dim, 3, padding=1), nn.GroupNorm(8, dim), nn.ReLU())
        self.ca_transformer_blocks = nn.ModuleList(
            [CrossonlyDecoderBlock(dim, num_heads, mlp_ratio=1) for _ in range(num_layers)])
        self.ca_transformer_norm = nn.LayerNorm(dim)

    def forward(self, inputs, img_shape):
        lr, img = inputs
        H, W = img_shape
        B = lr.shape[0]
        lr_map = lr.transpose(1, 2).reshape(B, -1, H // self.patch_size, W // self.patch_size)
        patch_feats = self.patch_embed(lr_map)
        if H > W:
            img = img.transpose(2, 3)
        if self.output_stride != 1:
            img = F.interpolate(img, scale_factor=1.0 / self.output_stride, mode="bilinear", align_corners=False)
        x = self.first_conv(self.fourier_feat(img))
        _, Ch, Ho, Wo = x.shape
        x = x.flatten(2).transpose(1, 2)
        lr_tok = torch.cat([lr_map, self.lr_pe(lr_map)], dim=1).flatten(2).transpose(1, 2)
        lr_tok = self.lr_input_proj(lr_tok)
        for blk in self.ca_transformer_blocks:
            x, _ = blk(x, lr_tok)
        x = self.ca_transformer_norm(x)
        return [patch_feats], x.transpose(1, 2).reshape(B, Ch, Ho, Wo)


class InputMixer(nn.Module):
    def __init__(self, img_size, patch_size, in_dim, hidden_dim, num_heads=12, num_layers=3, ff_dim_mult=4):
        super().__init__()
        self.in_proj = nn.Linear(in_dim, hidden_dim)
        self.rope = get_pos_embed("RoPE100")
        self.mixer_blk = nn.ModuleList(
            [Block(hidden_dim, num_heads, mlp_ratio=ff_dim_mult, rope=self.rope, qkv_bias=True) for _ in range(num_layers)])
        self.mixer_norm = nn.LayerNorm(hidden_dim)

    def forward(self, x, pos):
        x = self.in_proj(x)
        for blk in self.mixer_blk:
            x = blk(x, pos)
        return self.mixer_norm(x)


# ----------------------------------------------------------------------------------------------- text
class TextEncoder(nn.Module):
    """Fixed-vocabulary mode only: dictionary lookup + L2 normalisation (text_encoder.py:94-103)."""

    def __init__(self, model_name="siglip", out_dim=768, fixed_vocab=True):
        super().__init__()
        assert fixed_vocab, "the HF text tower is out of scope (needs network weights)"
        self.embed_dim = {"siglip": 768, "siglip2": 768, "clip": 512}[model_name]
        self.class_embeddings = {}

    def forward(self, classes: List[str]):
        e = torch.stack([self.class_embeddings[c] for c in classes])
        return e / e.norm(dim=-1, keepdim=True)


# ----------------------------------------------------------------------------------------------- mask transformer
def sine_position_embedding(h, w, num_pos_feats, device, temperature=10000.0):
    """PositionEmbeddingSine(normalize=True) for an unmasked (h, w) grid -> (2*num_pos_feats, h, w)."""
    eps, scale = 1e-6, 2 * math.pi
    y = torch.arange(1, h + 1, dtype=torch.float32, device=device) / (h + eps) * scale
    x = torch.arange(1, w + 1, dtype=torch.float32, device=device) / (w + eps) * scale
    i = torch.arange(num_pos_feats, dtype=torch.float32, device=device)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)

    def enc(v):  # (n,) -> (n, num_pos_feats): sin on even, cos on odd channels
        a = v[:, None] / dim_t
        return torch.stack([a[:, 0::2].sin(), a[:, 1::2].cos()], dim=2).flatten(1)

    py = enc(y)[:, None, :].expand(h, w, num_pos_feats)
    px = enc(x)[None, :, :].expand(h, w, num_pos_feats)
    return torch.cat([py, px], dim=2).permute(2, 0, 1)


class _SelfAttentionLayer(nn.Module):
    def __init__(self, d, nhead):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, nhead)
        self.norm = nn.LayerNorm(d)

    def forward(self, tgt, query_pos):
        qk = tgt + query_pos
        return self.norm(tgt + self.self_attn(qk, qk, value=tgt)[0])


class _CrossAttentionLayer(nn.Module):
    def __init__(self, d, nhead):
        super().__init__()
        self.multihead_attn = nn.MultiheadAttention(d, nhead)
        self.norm = nn.LayerNorm(d)

    def forward(self, tgt, memory, memory_mask, pos, query_pos):
        out = self.multihead_attn(query=tgt + query_pos, key=memory + pos, value=memory, attn_mask=memory_mask)[0]
        return self.norm(tgt + out)


class _FFNLayer(nn.Module):
    def __init__(self, d, ff):
        super().__init__()
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        self.norm = nn.LayerNorm(d)

    def forward(self, tgt):
        return self.norm(tgt + self.linear2(F.relu(self.linear1(tgt))))


class _MLP(nn.Module):
    def __init__(self, i, h, o, n):
        super().__init__()
        dims = [i] + [h] * (n - 1) + [o]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for k, l in enumerate(self.layers):
            x = l(x) if k == len(self.layers) - 1 else F.relu(l(x))
        return x


class MaskTransformer(nn.Module):
    def __init__(self, in_dim, hidden_dim, ff_dim, mask_dim, num_queries, num_heads, dec_layers, lang_dim=768,
                 num_feature_levels=1, landscape_only=True):
        super().__init__()
        assert num_feature_levels == 1 and list(in_dim) == [hidden_dim], "only the identity input_proj branch works upstream"
        self.num_heads, self.num_layers, self.num_queries = num_heads, dec_layers, num_queries
        self.hidden_dim = hidden_dim
        self.self_attn_layers = nn.ModuleList(_SelfAttentionLayer(hidden_dim, num_heads) for _ in range(dec_layers))
        self.cross_attn_layers = nn.ModuleList(_CrossAttentionLayer(hidden_dim, num_heads) for _ in range(dec_layers))
        self.ffn_layers = nn.ModuleList(_FFNLayer(hidden_dim, ff_dim) for _ in range(dec_layers))
        self.decoder_norm = nn.LayerNorm(hidden_dim)
        self.query_feat = nn.Embedding(num_queries, hidden_dim)
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.level_embed = nn.Embedding(1, hidden_dim)
        self.input_proj = nn.ModuleList([nn.Sequential()])
        self.lang_embed = nn.Linear(hidden_dim, lang_dim)
        self.cls_logit_scale = nn.Parameter(torch.ones([]))
        self.mask_embed = _MLP(hidden_dim, hidden_dim, mask_dim, 3)

    def forward_prediction_heads(self, output, mask_feats, cls_embeddings, attn_mask_target_size=None):
        """output (Q, B, C); mask_feats (B, V, Cm, Hm, Wm).  Returns class logits (B,Q,K), mask logits (B,V,Q,Hm,Wm),
        boolean attention mask (B*heads, Q, V*h*w) with True = blocked (or None).
        Multi aspect ratio (mask_transformer.py:215-275 with multi_ar=True): mask_feats and attn_mask_target_size are
        lists with one entry per stack; mask logits come back as a list, the attention mask covers the concatenated
        tokens of all stacks in stack order."""
        multi = isinstance(mask_feats, (list, tuple))
        mfs = list(mask_feats) if multi else [mask_feats]
        sizes = (list(attn_mask_target_size) if multi else [attn_mask_target_size]) if attn_mask_target_size is not None else None
        dec = self.decoder_norm(output).transpose(0, 1)
        lang = self.lang_embed(dec)
        lang = lang / (lang.norm(dim=-1, keepdim=True) + 1e-7)
        logits = self.cls_logit_scale.exp() * lang @ cls_embeddings.unsqueeze(0).transpose(1, 2)
        emb = self.mask_embed(dec)
        masks = [torch.einsum("bqc,bvchw->bvqhw", emb, mf) for mf in mfs]
        attn_mask = None
        if sizes is not None:
            smalls = []
            for mk, size in zip(masks, sizes):
                B, V, Q, _, _ = mk.shape
                small = F.interpolate(mk.flatten(0, 1), size=tuple(size), mode="bilinear", align_corners=False)
                smalls.append(small.view(B, V, Q, -1).permute(0, 2, 1, 3).flatten(2))  # (B, Q, V*h*w)
            small = torch.cat(smalls, dim=2)
            attn_mask = (small.sigmoid().unsqueeze(1).repeat(1, self.num_heads, 1, 1).flatten(0, 1) < 0.5).bool().detach()
        return logits, (masks if multi else masks[0]), attn_mask

    def _stack_tokens(self, f, true_shape):
        """One stack f (B, V, C, h, w) -> (memory tokens (V*h*w, B, C) + level embedding, their positional encoding)."""
        B, V, Cc, h, w = f.shape
        Ht, Wt = int(true_shape[0, 0, 0]), int(true_shape[0, 0, 1])
        if Wt >= Ht:
            pe = sine_position_embedding(h, w, self.hidden_dim // 2, f.device).flatten(1)  # (C, h*w)
        else:  # portrait views: PE of the (w, h) grid, flattened in ITS row-major order and applied as-is to the
            # landscape-stored tokens (mask_transformer.py:112-115 — restated literally, no transpose back)
            pe = sine_position_embedding(w, h, self.hidden_dim // 2, f.device).flatten(1)
        pos = pe.t()[:, None, :].expand(h * w, B, Cc).repeat(V, 1, 1)  # (V*h*w, B, C)
        src = f.permute(0, 2, 1, 3, 4).flatten(2).permute(2, 0, 1) + self.level_embed.weight[0][None, None]
        return src, pos

    def forward(self, fpn_f, mask_feats, true_shape, cls_embeddings, deep_supervision=True, multi_ar=False):
        """multi_ar: fpn_f[0], mask_feats, true_shape are lists with one entry per stack (mask_transformer.py:126-146)."""
        if multi_ar:
            toks = [self._stack_tokens(f, ts) for f, ts in zip(fpn_f[0], true_shape)]
            src, pos = torch.cat([t[0] for t in toks], 0), torch.cat([t[1] for t in toks], 0)
            B = fpn_f[0][0].shape[0]
            size = [tuple(f.shape[-2:]) for f in fpn_f[0]]
        else:
            src, pos = self._stack_tokens(fpn_f[0], true_shape)
            B = fpn_f[0].shape[0]
            size = tuple(fpn_f[0].shape[-2:])
        h, w = size if not multi_ar else (None, None)
        query_embed = self.query_embed.weight.unsqueeze(1).repeat(1, B, 1)
        output = self.query_feat.weight.unsqueeze(1).repeat(1, B, 1)
        cls, msk, attn_mask = self.forward_prediction_heads(output, mask_feats, cls_embeddings, size)
        pred_cls, pred_msk = ([cls], [msk]) if deep_supervision else ([], [])
        for i in range(self.num_layers):
            attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False
            output = self.cross_attn_layers[i](output, src, attn_mask, pos, query_embed)
            output = self.self_attn_layers[i](output, query_embed)
            output = self.ffn_layers[i](output)
            cls, msk, attn_mask = self.forward_prediction_heads(output, mask_feats, cls_embeddings, size)
            if deep_supervision or i == self.num_layers - 1:
                pred_cls.append(cls)
                pred_msk.append(msk)
        return {
            "pred_logits": pred_cls[-1],
            "pred_masks": pred_msk[-1],
            "aux_outputs": [{"pred_logits": a, "pred_masks": b} for a, b in zip(pred_cls[:-1], pred_msk[:-1])],
            "out_queries": output.detach(),
        }


class PanopticDecoder(nn.Module):
    def __init__(self, input_mixer=None, upscaler=None, fpn_dim=(768,), hidden_dim=768, mask_dim=256, ff_dim=2048,
                 num_queries=200, num_heads=8, dec_layers=6, text_encoder="siglip", fixed_vocab=True,
                 label_mode="sigmoid", landscape_only=True, deep_supervision=True):
        super().__init__()
        assert upscaler is not None and label_mode == "sigmoid"
        self.input_mixer = input_mixer
        self.upscaler = upscaler
        self.text_encoder = TextEncoder(text_encoder, out_dim=hidden_dim, fixed_vocab=fixed_vocab)
        self.mask_transformer = MaskTransformer(list(fpn_dim), hidden_dim, ff_dim, mask_dim, num_queries, num_heads,
                                                dec_layers, lang_dim=self.text_encoder.embed_dim,
                                                num_feature_levels=len(fpn_dim), landscape_only=landscape_only)
        self.deep_supervision = deep_supervision

    def _stack_features(self, cat, in_imgs, pos, true_shape):
        """One stack of equally shaped views: cat (B, V, N, 2816) -> ([fpn (B, V, C, h, w)], mask_f (B, V, Cm, Hm, Wm))."""
        B, V = cat.shape[:2]
        x = cat.flatten(0, 1)
        if self.input_mixer is not None:
            x = self.input_mixer(x, pos.flatten(0, 1))
        ts = true_shape.flatten(0, 1)
        H, W = int(ts[0, 0]), int(ts[0, 1])
        if W >= H:
            fpn, mask_f = self.upscaler((x, in_imgs.flatten(0, 1)), (H, W))
        else:  # portrait: predict in the true (portrait) shape, then swap spatial dims back to the landscape
            # storage convention (utils.transpose_to_landscape, dims=(2,3); utils.py:46-49)
            fpn, mask_f = self.upscaler((x, in_imgs.flatten(0, 1)), (H, W))
            fpn, mask_f = [t.swapaxes(2, 3) for t in fpn], mask_f.swapaxes(2, 3)
        return [t.unflatten(0, (B, V)) for t in fpn], mask_f.unflatten(0, (B, V))

    def forward(self, in_feats, in_imgs, pos, true_shape, classes, max_bs=None, outdevice=None, memory_queries=None,
                multi_ar=False):
        """multi_ar (panoptic_decoder.py:44-45, 53-76): every tensor argument is a list with one entry per stack of
        equally shaped views; `pred_masks` comes back as a list with one tensor per stack."""
        if multi_ar:
            cats = [torch.cat(parts, dim=-1) for parts in zip(*in_feats)]
            per = [self._stack_features(c, im, p, ts) for c, im, p, ts in zip(cats, in_imgs, pos, true_shape)]
            fpn = [[st[0][lvl] for st in per] for lvl in range(len(per[0][0]))]
            mask_f = [st[1] for st in per]
        else:
            fpn, mask_f = self._stack_features(torch.cat(in_feats, dim=-1), in_imgs, pos, true_shape)
        cls_emb = self.text_encoder(classes).to(mask_f[0].device if multi_ar else mask_f.device)
        if memory_queries is None:
            return self.mask_transformer(fpn, mask_f, true_shape, cls_emb, deep_supervision=self.deep_supervision,
                                         multi_ar=multi_ar)
        logits, masks, _ = self.mask_transformer.forward_prediction_heads(memory_queries, mask_f, cls_emb)
        return {"pred_logits": logits, "pred_masks": masks}


# ----------------------------------------------------------------------------------------------- DINOv2
class DinoV2Encoder(nn.Module):
    """[-1,1] -> ImageNet normalisation -> bilinear resize to (H/16*14, W/16*14) -> HF Dinov2Model -> drop CLS.
    The HF model is built from a config (no network): random-init weights."""

    def __init__(self, dinov2: Optional[nn.Module] = None, output_stride=16, hidden_size=1024, depth=24, heads=16):
        super().__init__()
        if dinov2 is None:
            from transformers import Dinov2Config, Dinov2Model
            dinov2 = Dinov2Model(Dinov2Config(hidden_size=hidden_size, num_hidden_layers=depth, num_attention_heads=heads,
                                              patch_size=14, image_size=518)).eval()
        self.dinov2 = dinov2
        self.output_stride = output_stride
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1), persistent=False)
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1), persistent=False)

    def forward(self, image, true_shape):
        x = (image * 0.5 + 0.5 - self.mean) / self.std
        P = self.dinov2.config.patch_size
        h, w = [s // self.output_stride * P for s in image.shape[-2:]]
        x = F.interpolate(x, size=(h, w), mode="bilinear", align_corners=False)
        Ht, Wt = int(true_shape[0, 0]), int(true_shape[0, 1])
        if Wt < Ht:
            x = x.transpose(2, 3)
        return self.dinov2(pixel_values=x).last_hidden_state[:, 1:]
