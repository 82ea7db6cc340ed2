"""CUDA panoptic head: PanopticDecoder / MaskTransformer / PixelShuffleUpscaler / TextEncoder with the
reference's constructor arguments, forward signatures and state-dict names
(reference src/panst3r/model/panoptic_decoder.py:16-77, mask_transformer.py:12-288,
upscalers/pixel_shuffle.py:9-59, text_encoder.py:94-103).

B200-first restructuring (results are those of the reference ops on the same inputs):
  * feature maps stay pixel-major (NHWC bf16): F.pixel_shuffle is the store pattern of the producing GEMM and
    the mask einsum "bqc,bnchw->bnqhw" is a tcgen05 GEMM (pixels x queries) with a plane-major fp32 store;
  * the 8x bilinear downsample feeding the attention mask equals the mean of the centre 2x2 logits and is linear
    in the features, so it is computed from centre-pooled features (1/64 of the pixels) as a second small GEMM;
  * K/V in-projections of the (layer-invariant) memory tokens of all 6 decoder layers are two batched GEMMs;
  * level_embed is folded into the proj_16.fc2 bias; the sine PE is a per-grid constant;
  * the boolean attention mask is a bitmask consumed directly by the attention kernel.

Numerics.  The reference runs this head in fp32 even when the trunk runs under bf16 autocast (panst3r.py:236-245).
`precise=True` (PanopticDecoder.precision == "fp32", the default) reproduces that policy on the bf16 tensor cores:
activations and weights are `ops.Split` pairs (hi + lo bf16, 16 mantissa bits), every GEMM accumulates
A_hi B_hi + A_lo B_hi + A_hi B_lo in fp32, LayerNorm / softmax / GELU are evaluated in fp32, and the 200-query attention
of the query decoder is evaluated unfused (QK^T GEMM -> masked row softmax -> PV GEMM) so that its probabilities keep
16 bits too.  `precision == "bf16"` keeps plain bf16 operands and the fused flash-attention kernels (faster, ~1e-2).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import ops
from .common import _as_split, _split_cat, b16, bias_of, cat_f32, cat_w16, f32, prepared, psplit, w16, wsplit
from .must3r import _hw, oriented


class _Mlp(nn.Module):
    def __init__(self, i, h, o):
        super().__init__()
        self.fc1 = nn.Linear(i, h)
        self.fc2 = nn.Linear(h, o)


class PixelShuffleUpscaler(nn.Module):
    def __init__(self, input_dim, patch_size=16, hidden_dim_factor=4, fp_dim=(768, 512, 384, 256), **kwargs):
        super().__init__()
        self.patch_size, self.fp_dim = patch_size, list(fp_dim)
        f = hidden_dim_factor
        self.proj_8 = _Mlp(input_dim, int(f * input_dim), fp_dim[1] * 4)
        self.proj_4 = _Mlp(fp_dim[1], int(f * fp_dim[1]), fp_dim[2] * 4)
        self.proj_2 = _Mlp(fp_dim[2], int(f * fp_dim[2]), fp_dim[3] * 4)
        self.proj_16 = _Mlp(input_dim, int(f * input_dim), fp_dim[0])
        self.mask_dim = fp_dim[3]

    @torch.no_grad()
    def forward_nhwc(self, feats, b: int, hs: int, ws: int, f16_extra_bias: Optional[torch.Tensor] = None,
                     precise: bool = False):
        """feats bf16 (or ops.Split) rows (b*hs*ws, input_dim) -> (f16 (b*hs*ws, 768) token-major, mask feats
        (b, 8hs, 8ws, 256) pixel-major); bf16 tensors, or ops.Split pairs when `precise`.  f16_extra_bias is added to
        the f16 output (level_embed)."""
        dev = feats.device
        W = wsplit if precise else w16
        act = "split" if precise else torch.bfloat16

        def mlp_shuffle(x, m: _Mlp, gh, gw):
            h = ops.gemm(x, W(m.fc1.weight), bias=bias_of(m.fc1), act=ops.ACT_GELU, out_dtype=act)
            cout = m.fc2.weight.shape[0] // 4
            out = ops._new((b * 2 * gh * 2 * gw, cout), ops._dkind(act), dev)
            ops.gemm(h, W(m.fc2.weight), bias=bias_of(m.fc2), out=out, store_mode=ops.STORE_PIXSHUF2, grid=(gh, gw))
            return out

        f8 = mlp_shuffle(feats, self.proj_8, hs, ws)
        f4 = mlp_shuffle(f8, self.proj_4, 2 * hs, 2 * ws)
        f2 = mlp_shuffle(f4, self.proj_2, 4 * hs, 4 * ws)
        h = ops.gemm(feats, W(self.proj_16.fc1.weight), bias=bias_of(self.proj_16.fc1), act=ops.ACT_GELU, out_dtype=act)
        bias16 = f32(self.proj_16.fc2.bias)
        if f16_extra_bias is not None:
            bias16 = prepared("f16bias", [self.proj_16.fc2.bias, f16_extra_bias],
                              lambda: (self.proj_16.fc2.bias.detach().float() + f16_extra_bias.detach().float().view(-1)).contiguous())
        f16 = ops.gemm(h, W(self.proj_16.fc2.weight), bias=bias16, out_dtype=act)
        return f16, f2.view(b, 8 * hs, 8 * ws, self.mask_dim)

    @torch.no_grad()
    def forward(self, inputs, img_shape, precise: bool = False):
        """Reference signature: (feats (b, N, C), imgs), (H, W) -> ([f16 (b,768,hs,ws)], mask_feats (b,256,H/2,W/2)) fp32 NCHW."""
        feats = inputs[0]
        H, W = img_shape
        hs, ws = H // self.patch_size, W // self.patch_size
        b = feats.shape[0]
        x = _head_input(feats.reshape(b * hs * ws, -1), precise)
        f16, f2 = self.forward_nhwc(x, b, hs, ws, precise=precise)
        f16_nchw = ops.nhwc_to_nchw_f32(f16.view(b, hs * ws, -1)).view(b, -1, hs, ws)
        f2_nchw = ops.nhwc_to_nchw_f32(f2.view(b, 64 * hs * ws, -1)).view(b, -1, 8 * hs, 8 * ws)
        return [f16_nchw], f2_nchw


def _head_input(x: torch.Tensor, precise: bool):
    """Head input rows: bf16 stays bf16 (exact; two-term products in precise mode); fp32 becomes an ops.Split when
    precise, bf16 otherwise."""
    if x.dtype == torch.bfloat16:
        return x
    x = x.float().contiguous()
    return ops.Split.from_float(x) if precise else ops.to_bf16(x)


class TextEncoder(nn.Module):
    """Fixed-vocabulary mode: dictionary lookup + L2 normalisation (text_encoder.py:94-103).  The HF text tower
    (`set_vocab`) runs once per vocabulary, needs downloaded weights and is out of scope: provide embeddings
    through `class_embeddings` (same attribute as the reference)."""

    def __init__(self, model_name="siglip", out_dim=768, fixed_vocab=True):
        super().__init__()
        self.embed_dim = {"siglip": 768, "siglip2": 768, "clip": 512}[model_name]
        self.fixed_vocab = fixed_vocab
        self.class_embeddings: Dict[str, torch.Tensor] = {}
        self._norm_cache = {}

    def set_vocab(self, classes, device=None):
        raise NotImplementedError("the HF text tower is outside the CUDA hot path; assign `class_embeddings` directly")

    @torch.no_grad()
    def forward(self, classes: List[str], device=None, precise: bool = False):
        assert all(c in self.class_embeddings for c in classes), "Missing classes in vocabulary"
        embs = [self.class_embeddings[c] for c in classes]
        # identity + version of every embedding tensor: re-assigned / updated vectors for the same names must not hit
        key = (tuple(classes), tuple((id(e), e._version, e.data_ptr()) for e in embs), str(torch.device(device)), precise)
        hit = self._norm_cache.get(key)
        if hit is None:
            e = torch.stack(embs).to(device=device, dtype=torch.float32).contiguous()
            hit = (ops.l2norm_rows(e, 0.0, "split" if precise else torch.bfloat16), embs)  # embs kept alive: ids stay unique
            self._norm_cache[key] = hit
            while len(self._norm_cache) > 4:
                self._norm_cache.pop(next(iter(self._norm_cache)))
        return hit[0]


class _MHA(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = nn.Linear(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)


class _SALayer(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.self_attn = _MHA(d)
        self.norm = nn.LayerNorm(d)


class _CALayer(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.multihead_attn = _MHA(d)
        self.norm = nn.LayerNorm(d)


class _FFN(nn.Module):
    def __init__(self, d, ff):
        super().__init__()
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        self.norm = nn.LayerNorm(d)


class _MaskMLP(nn.Module):
    def __init__(self, i, h, o, n):
        super().__init__()
        dims = [i] + [h] * (n - 1) + [o]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))


def sine_pe(h: int, w: int, num_pos_feats: int, temperature: float = 10000.0) -> torch.Tensor:
    """PositionEmbeddingSine(normalize=True) on an unmasked (h, w) grid -> fp32 (h*w, 2*num_pos_feats), row-major
    tokens, channels [pos_y | pos_x] (mask_transformer.py:504-527).  A constant of the grid shape."""
    eps, scale = 1e-6, 2 * math.pi
    y = torch.arange(1, h + 1, dtype=torch.float32) / (h + eps) * scale
    x = torch.arange(1, w + 1, dtype=torch.float32) / (w + eps) * scale
    i = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)

    def enc(v):
        a = v[:, None] / dim_t
        return torch.stack([a[:, 0::2].sin(), a[:, 1::2].cos()], dim=2).flatten(1)

    py = enc(y)[:, None, :].expand(h, w, num_pos_feats)
    px = enc(x)[None, :, :].expand(h, w, num_pos_feats)
    return torch.cat([py, px], dim=2).reshape(h * w, 2 * num_pos_feats)


class MaskTransformer(nn.Module):
    def __init__(self, in_dim, hidden_dim, ff_dim, mask_dim, num_queries, num_heads, dec_layers, lang_dim=768,
                 normalize_before=False, num_feature_levels=1, enforce_input_project=False, two_stage=False,
                 landscape_only=False):
        super().__init__()
        in_dim = [in_dim] * num_feature_levels if isinstance(in_dim, int) else list(in_dim)
        assert num_feature_levels == 1 and in_dim == [hidden_dim] and not two_stage and not normalize_before
        self.hidden_dim, self.num_heads, self.num_layers, self.num_queries = hidden_dim, num_heads, dec_layers, num_queries
        self.mask_dim = mask_dim
        self.self_attn_layers = nn.ModuleList(_SALayer(hidden_dim) for _ in range(dec_layers))
        self.cross_attn_layers = nn.ModuleList(_CALayer(hidden_dim) for _ in range(dec_layers))
        self.ffn_layers = nn.ModuleList(_FFN(hidden_dim, ff_dim) for _ in range(dec_layers))
        self.decoder_norm = nn.LayerNorm(hidden_dim)
        self.query_feat = nn.Embedding(num_queries, hidden_dim)
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.level_embed = nn.Embedding(num_feature_levels, hidden_dim)
        self.input_proj = nn.ModuleList([nn.Sequential()])
        self.lang_embed = nn.Linear(hidden_dim, lang_dim)
        self.cls_logit_scale = nn.Parameter(torch.ones([]))
        self.mask_embed = _MaskMLP(hidden_dim, hidden_dim, mask_dim, 3)
        # test instrumentation (the query decoder is a discontinuous function of sign(mask logit) decisions):
        # `mask_override`: list of 6 block-mask bit tensors forced into the layers; `bits_record`: list that receives the
        # block mask each layer actually used.  Both None in normal operation.
        self.mask_override: Optional[List[torch.Tensor]] = None
        self.bits_record: Optional[list] = None
        # The masks of the auxiliary heads (deep supervision: the initial prediction and layers 0 .. L-2) feed nothing but the
        # output dict, while the decoder layer that follows them is a chain of small, latency-bound launches on 200 query rows:
        # their HBM-bound full-resolution GEMMs run on a side stream, on `aux_mask_sms` SMs, next to that layer, and are
        # joined before forward_nhwc returns.  Bit-identical to the in-line run (tests/test_lazy_masks.py).  Measured at 16 views
        # of 512x384 (profiles/r02_lazy_masks.md): the bf16 head gains 0.4 of 6.0 ms; the fp32-grade head, whose decoder
        # layers are real tensor work on split operands, gains nothing (15.84 -> 15.79 ms).  None = on for the bf16 head
        # only; True / False (or PST3R_AUX_OVERLAP=1 / 0) force it.
        env = os.environ.get("PST3R_AUX_OVERLAP")
        self.overlap_aux_masks: Optional[bool] = None if env is None else env != "0"
        self.aux_mask_sms = 64
        self._aux_stream = None

    # ---- prepared -------------------------------------------------------------------------------
    def _kv_weights(self, precise: bool = False):
        d = self.hidden_dim
        ws = [l.multihead_attn.in_proj_weight for l in self.cross_attn_layers]
        bs = [l.multihead_attn.in_proj_bias for l in self.cross_attn_layers]
        bk = prepared("mt_bk", bs, lambda: torch.cat([b.detach()[d:2 * d] for b in bs], 0).float().contiguous())
        bv = prepared("mt_bv", bs, lambda: torch.cat([b.detach()[2 * d:] for b in bs], 0).float().contiguous())
        if precise:
            wk = _as_split(prepared("mt_wk_s", ws, lambda: _split_cat(torch.cat([w.detach()[d:2 * d].float() for w in ws], 0))))
            wv = _as_split(prepared("mt_wv_s", ws, lambda: _split_cat(torch.cat([w.detach()[2 * d:].float() for w in ws], 0))))
            return wk, bk, wv, bv
        wk = prepared("mt_wk", ws, lambda: torch.cat([w.detach()[d:2 * d] for w in ws], 0).to(torch.bfloat16).contiguous())
        wv = prepared("mt_wv", ws, lambda: torch.cat([w.detach()[2 * d:] for w in ws], 0).to(torch.bfloat16).contiguous())
        return wk, bk, wv, bv

    def _pos(self, h, w, device, portrait: bool = False, precise: bool = False):
        """Sine PE rows for the (h, w) token grid of one view (bf16, or fp32 when precise).  Portrait views (stored
        transposed): the reference embeds the transposed map and applies its rows, in THAT map's raster order, to the
        stored tokens as they are (get_pe_with_transpose, mask_transformer.py:106-119) — restated literally."""
        dt = torch.float32 if precise else torch.bfloat16
        gh, gw = (w, h) if portrait else (h, w)
        return prepared(f"mt_pe_{h}x{w}_{'p' if portrait else 'l'}_{'f' if precise else 'b'}", [self.level_embed.weight],
                        lambda: sine_pe(gh, gw, self.hidden_dim // 2).to(device=device, dtype=dt).contiguous())

    @staticmethod
    def _slice_w(p: torch.Tensor, a: int, b: int, tag: str, precise: bool = False):
        if p.dim() == 2 and precise:
            return _as_split(prepared(f"mt_slice_s_{tag}_{a}_{b}", [p], lambda: _split_cat(p.detach()[a:b].float())))
        return prepared(f"mt_slice_{tag}_{a}_{b}", [p], lambda: (p.detach()[a:b].to(torch.bfloat16) if p.dim() == 2 else
                                                                 p.detach()[a:b].float()).contiguous())

    # ---- prediction heads (mask_transformer.py:215-288) -----------------------------------------------
    @torch.no_grad()
    def prediction_heads(self, output, mask_feats, pooled, cls_emb, want_masks: bool, precise: bool = False,
                         lazy: bool = False, side=None):
        """output (Q, C) [batch 1]; mask_feats (V, Hm, Wm, Cm) pixel-major; pooled (V*h*w, Cm) or None — bf16 tensors,
        or ops.Split pairs when `precise`.
        Returns (class logits fp32 (Q, K), mask logits fp32 (V, Q, Hm, Wm) | None, mask bits int32 (1, Q, W) | None).
        lazy: the mask logits come back as a `postprocess.LazyMasks` (embeddings + features; the einsum is left to the
        post-processing, which evaluates it band by band through the L2).
        side = (stream, keep_alive list): the full-resolution mask GEMM is launched on that stream, on `aux_mask_sms` SMs,
        so that it runs next to the following decoder layer (whose launches are small and latency bound); the caller joins
        the stream before it returns and keeps `keep_alive` until then."""
        Q = output.shape[0]
        W = wsplit if precise else w16
        act = "split" if precise else torch.bfloat16
        dev = output.device
        dec = ops.layernorm(output, f32(self.decoder_norm.weight), f32(self.decoder_norm.bias), 1e-5)
        lang = ops.gemm(dec, W(self.lang_embed.weight), bias=bias_of(self.lang_embed), out_dtype=torch.float32)
        lang = ops.l2norm_rows(lang, 1e-7, act)
        scale = prepared("mt_scale", [self.cls_logit_scale], lambda: self.cls_logit_scale.detach().float().exp().cpu())
        logits = ops.gemm(lang, cls_emb, alpha=float(scale), out_dtype=torch.float32)
        e = dec
        nl = len(self.mask_embed.layers)
        for i, l in enumerate(self.mask_embed.layers):
            e = ops.gemm(e, W(l.weight), bias=bias_of(l), act=ops.ACT_RELU if i < nl - 1 else ops.ACT_NONE, out_dtype=act)
        masks = None
        if want_masks:
            def plane_major(mf):
                if lazy:
                    from ..postprocess import LazyMasks
                    return LazyMasks(mf, e)
                V, Hm, Wm, Cm = mf.shape
                mk = torch.empty((V, Q, Hm, Wm), device=dev, dtype=torch.float32)
                ops.gemm(mf.view(V * Hm * Wm, Cm), e, out=mk, store_mode=ops.STORE_TRANSPOSED,
                         rows_per_batch=Hm * Wm, batch_stride=Q * Hm * Wm, ldt=Hm * Wm)
                return mk
            # several aspect-ratio stacks (multi_ar): one mask tensor per stack, as the reference returns them
            def all_planes():
                return [plane_major(mf) for mf in mask_feats] if isinstance(mask_feats, (list, tuple)) else plane_major(mask_feats)
            if side is not None and not lazy:
                stream, keep_alive = side
                stream.wait_stream(torch.cuda.current_stream())  # fork: the embeddings (and the features) are complete
                keep_alive.append(e)                             # read by the side stream: not to be recycled before the join
                with torch.cuda.stream(stream), ops.sm_budget(self.aux_mask_sms):
                    masks = all_planes()
            else:
                masks = all_planes()
        bits = None
        if pooled is not None:
            nk = pooled.shape[0]
            small = torch.empty((Q, nk), device=dev, dtype=torch.float32)
            ops.gemm(pooled, e, out=small, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=nk, batch_stride=0, ldt=nk)
            bits = ops.attn_mask_bits(small, nk)
        return logits, masks, bits

    def _mha(self, q_in, k, v, m: _MHA, mask_bits, residual, precise: bool = False):
        """residual + out_proj(attention(q_in Wq^T + bq, k, v)).  k, v already projected:
        bf16 mode: (1, Nk, H, hd) tensors for the fused flash-attention kernel;
        precise:   k = ops.Split (H, Nk, hd), v = ops.Split (H, hd, Nk) [V^T] for the unfused reference-precision attention
                   S = (q k^T) / sqrt(hd)  ->  masked row softmax (fp32)  ->  P v, all on split operands."""
        d, H = self.hidden_dim, self.num_heads
        hd = d // H
        wq, bq = self._slice_w(m.in_proj_weight, 0, d, "wq", precise), self._slice_w(m.in_proj_bias, 0, d, "bq")
        if not precise:
            q = ops.gemm(q_in, wq, bias=bq).view(1, -1, H, hd)
            o = ops.attention(q, k, v, mask_bits=mask_bits)
            return ops.gemm(o.view(-1, d), w16(m.out_proj.weight), bias=bias_of(m.out_proj), residual=residual)
        q = ops.gemm(q_in, wq, bias=bq, out_dtype="split")
        o = attention_precise(q, k, v, hd ** -0.5, mask_bits)
        return ops.gemm(o, wsplit(m.out_proj.weight), bias=bias_of(m.out_proj), residual=residual, out_dtype="split")

    @torch.no_grad()
    def forward_nhwc(self, src, mask_feats, hw, cls_emb,
                     deep_supervision: bool = True, pooled=None,
                     mask_override: Optional[List[torch.Tensor]] = None, portrait=False, precise: bool = False,
                     lazy_masks: bool = False):
        """src (V*h*w, C) = stride-16 features + level_embed, views flattened view-major (batch 1);
        mask_feats (V', Hm, Wm, Cm) — the views whose full-resolution masks this call produces (all V, or this
        rank's shard); pooled (V*h*w, Cm): centre-pooled mask features of ALL views (computed from mask_feats
        when omitted).  bf16 tensors, or ops.Split pairs when `precise` (see the module docstring).
        Returns the reference's output dict (batch dim 1).
        Multi aspect ratio (mask_transformer.py:126-146 with multi_ar=True): src / mask_feats / hw / portrait are
        LISTS with one entry per stack of equally shaped views; the memory tokens of all stacks are concatenated in
        stack order (each with the PE of its own grid) and `pred_masks` comes back as a list with one tensor per stack.
        lazy_masks: `pred_masks` is a `postprocess.LazyMasks` (the final einsum is left to the post-processing) and the
        auxiliary heads are not evaluated (`aux_outputs` is empty): the layers only need the pooled attention masks."""
        if lazy_masks:
            deep_supervision = False
        if mask_override is None:
            mask_override = self.mask_override
        multi = isinstance(src, (list, tuple))
        srcs, mfs, hws = (list(src), list(mask_feats), list(hw)) if multi else ([src], [mask_feats], [hw])
        ports = list(portrait) if isinstance(portrait, (list, tuple)) else [portrait] * len(srcs)
        d, H, Q = self.hidden_dim, self.num_heads, self.num_queries
        hd = d // H
        dev = srcs[0].device
        act = "split" if precise else torch.bfloat16
        cat0 = _cat_rows
        # key = memory + pos (pos of view 0 of each stack tiled over its views, :139-141)
        src_pos = [ops.add_bcast(s_, self._pos(h_, w_, dev, p_, precise)) for s_, (h_, w_), p_ in zip(srcs, hws, ports)]
        src = srcs[0] if len(srcs) == 1 else cat0(srcs)
        src_pos = src_pos[0] if len(src_pos) == 1 else cat0(src_pos)
        Nk = src.shape[0]
        wk, bk, wv, bv = self._kv_weights(precise)
        L = self.num_layers
        if precise:
            k_all = ops.gemm(src_pos, wk, bias=bk, out_dtype="split")  # (Nk, L*d)
            vT = ops.Split.empty((L * d, Nk), dev, align=8)             # V^T: the PV GEMM wants its B operand K-major
            ops.gemm(src, wv, bias=bv, out=vT, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=Nk, batch_stride=0,
                     ldt=vT.hi.stride(0))
            k_of = lambda i: k_all[:, i * d:(i + 1) * d].view(Nk, H, hd).permute(1, 0, 2)  # noqa: E731
            v_of = lambda i: vT[i * d:(i + 1) * d].view(H, hd, Nk)                          # noqa: E731
        else:
            k_all = ops.gemm(src_pos, wk, bias=bk).view(1, Nk, L, H, hd)
            v_all = ops.gemm(src, wv, bias=bv).view(1, Nk, L, H, hd)
            k_of = lambda i: k_all[:, :, i]  # noqa: E731
            v_of = lambda i: v_all[:, :, i]  # noqa: E731
        if pooled is None:
            pl = [ops.center_pool8(mf).view(-1, mf.shape[-1]) for mf in mfs]
            pooled = pl[0] if len(pl) == 1 else cat0(pl)
        mask_feats = mfs if multi else mfs[0]
        qe = psplit(self.query_embed.weight) if precise else b16(self.query_embed.weight)
        output = psplit(self.query_feat.weight) if precise else b16(self.query_feat.weight)
        pred_cls, pred_msk = [], []
        side = None
        overlap = (not precise) if self.overlap_aux_masks is None else self.overlap_aux_masks
        if deep_supervision and overlap and not lazy_masks:
            if self._aux_stream is None:
                self._aux_stream = torch.cuda.Stream()
            side = (self._aux_stream, [])
        cls, msk, bits = self.prediction_heads(output, mask_feats, pooled, cls_emb, want_masks=deep_supervision, precise=precise,
                                               side=side)
        if deep_supervision:
            pred_cls.append(cls)
            pred_msk.append(msk)
        for i in range(L):
            ca, sa, ff = self.cross_attn_layers[i], self.self_attn_layers[i], self.ffn_layers[i]
            W = wsplit if precise else w16
            # masked cross-attention (post-norm): tgt = LN(tgt + MHA(tgt + query_pos, memory + pos, memory))
            if self.bits_record is not None:
                self.bits_record.append(bits.clone())
            if mask_override is not None:  # test hook: force the block mask of layer i (isolates threshold flips)
                bits = mask_override[i]
            t = self._mha(ops.add_bcast(output, qe), k_of(i), v_of(i), ca.multihead_attn, bits, output, precise)
            output = ops.layernorm(t, f32(ca.norm.weight), f32(ca.norm.bias), 1e-5)
            # self-attention: q = k = tgt + query_pos, v = tgt
            m = sa.self_attn
            qk_in = ops.add_bcast(output, qe)
            wk_s, bk_s = self._slice_w(m.in_proj_weight, d, 2 * d, "wk", precise), self._slice_w(m.in_proj_bias, d, 2 * d, "bk")
            wv_s, bv_s = self._slice_w(m.in_proj_weight, 2 * d, 3 * d, "wv", precise), self._slice_w(m.in_proj_bias, 2 * d, 3 * d, "bv")
            if precise:
                kk = ops.gemm(qk_in, wk_s, bias=bk_s, out_dtype="split").view(Q, H, hd).permute(1, 0, 2)
                vvT = ops.Split.empty((d, Q), dev, align=8)
                ops.gemm(output, wv_s, bias=bv_s, out=vvT, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=Q, batch_stride=0,
                         ldt=vvT.hi.stride(0))
                t = self._mha(qk_in, kk, vvT.view(H, hd, Q), m, None, output, True)
            else:
                kk = ops.gemm(qk_in, wk_s, bias=bk_s)
                vv = ops.gemm(output, wv_s, bias=bv_s)
                t = self._mha(qk_in, kk.view(1, Q, H, hd), vv.view(1, Q, H, hd), m, None, output)
            output = ops.layernorm(t, f32(sa.norm.weight), f32(sa.norm.bias), 1e-5)
            # FFN
            hmid = ops.gemm(output, W(ff.linear1.weight), bias=bias_of(ff.linear1), act=ops.ACT_RELU, out_dtype=act)
            t = ops.gemm(hmid, W(ff.linear2.weight), bias=bias_of(ff.linear2), residual=output, out_dtype=act)
            output = ops.layernorm(t, f32(ff.norm.weight), f32(ff.norm.bias), 1e-5)
            last = i == L - 1
            cls, msk, bits = self.prediction_heads(output, mask_feats, None if last else pooled, cls_emb,
                                                   want_masks=deep_supervision or last, precise=precise, lazy=lazy_masks,
                                                   side=None if last else side)
            if deep_supervision or last:
                pred_cls.append(cls)
                pred_msk.append(msk)
        if side is not None:  # join: the auxiliary mask planes are complete before anything downstream reads them
            torch.cuda.current_stream().wait_stream(side[0])
            side[1].clear()
        b1 = (lambda t: [x[None] for x in t]) if multi else (lambda t: t[None])  # batch dim 1 (per stack when multi_ar)
        return {
            "pred_logits": pred_cls[-1][None],
            "pred_masks": b1(pred_msk[-1]),
            "aux_outputs": [{"pred_logits": a[None], "pred_masks": b1(b_)} for a, b_ in zip(pred_cls[:-1], pred_msk[:-1])],
            "out_queries": (ops.convert(output, torch.empty((Q, d), device=dev, dtype=torch.float32)) if precise else output).view(Q, 1, d),
        }


def attention_precise(q, k_heads, vT_heads, scale: float, mask_bits=None, out=None):
    """Reference-precision attention on split operands, unfused: S = scale * q k^T (batched GEMM over heads) -> row softmax
    in fp32 (optionally block-masked) -> O = P v (batched GEMM against V^T).
    q: ops.Split (Nq, H*hd) rows; k_heads: ops.Split (H, Nk, hd); vT_heads: ops.Split (H, hd, Nk).  -> ops.Split (Nq, H*hd).
    This is what croco's `CrossAttention` / `Attention` and nn.MultiheadAttention compute in the reference's fp32 head
    (model/blocks.py:29-35, mask_transformer.py:314,372); the probabilities never drop to bf16."""
    H, Nk, hd = k_heads.shape
    Nq = q.shape[0]
    S = torch.empty((H, Nq, Nk), device=q.device, dtype=torch.float32)
    ops.gemm_batched(q.view(Nq, H, hd).permute(1, 0, 2), k_heads, alpha=scale, out=S)
    P = ops.softmax_rows(S, Nq, mask_bits)
    o = out if out is not None else ops.Split.empty((Nq, H * hd), q.device)
    ops.gemm_batched(P, vT_heads, out=o.view(Nq, H, hd).permute(1, 0, 2))
    return o


def _cat_rows(parts):
    """Row-wise concatenation of bf16 tensors or packed ops.Split matrices (pure data movement)."""
    if isinstance(parts[0], ops.Split):
        return _as_split(torch.cat([p_.full() for p_ in parts], 0))
    return torch.cat(parts, 0)


class PanopticDecoder(nn.Module):
    def __init__(self, input_mixer=None, upscaler=None, fpn_dim=[768], hidden_dim=768, mask_dim=256, ff_dim=2048,
                 num_queries=200, num_heads=8, dec_layers=6, text_encoder="siglip", fixed_vocab=True,
                 label_mode="sigmoid", two_stage=False, landscape_only=True, deep_supervision=True, precision="fp32",
                 lazy_masks=False):
        super().__init__()
        assert upscaler is not None, "Upscaler module must be provided"
        assert label_mode == "sigmoid" and not two_stage
        self.input_mixer = input_mixer
        self.upscaler = upscaler
        self.text_encoder = TextEncoder(text_encoder, out_dim=hidden_dim, fixed_vocab=fixed_vocab)
        self.label_mode = label_mode
        self.mask_transformer = MaskTransformer(list(fpn_dim), hidden_dim, ff_dim, mask_dim, num_queries, num_heads,
                                                dec_layers, lang_dim=self.text_encoder.embed_dim,
                                                num_feature_levels=len(fpn_dim), landscape_only=landscape_only)
        self.deep_supervision = deep_supervision
        # "fp32": the reference's policy for this head (panst3r.py:236-245) on split-bf16 tensor-core operands;
        # "bf16": plain bf16 operands + fused flash attention (see the module docstring)
        assert precision in ("fp32", "bf16")
        self.precision = precision
        # True: `pred_masks` comes back as a `postprocess.LazyMasks` — the full-resolution mask einsum is evaluated by
        # `panoptic_inference_v1/_v2` band by band and never materialised (no auxiliary masks; one scene per call)
        self.lazy_masks = bool(lazy_masks)

    @property
    def precise(self) -> bool:
        return self.precision == "fp32"

    @torch.no_grad()
    def forward(self, in_feats, in_imgs, pos, true_shape, classes, max_bs=None, outdevice=None, memory_queries=None,
                multi_ar=False, cat_feats: Optional[torch.Tensor] = None):
        """Reference signature (panoptic_decoder.py:41).  in_feats = (x_enc, y_dec, x_dino), each (B, V, N, C_i);
        `cat_feats`: optional pre-concatenated bf16 (B, V, N, 2816) buffer the producers already wrote into
        (then in_feats is ignored)."""
        if multi_ar:
            return self._forward_multi_ar(in_feats, in_imgs, pos, true_shape, classes, outdevice, memory_queries, cat_feats)
        B = cat_feats.shape[0] if cat_feats is not None else in_feats[0].shape[0]
        if B > 1:  # scenes are independent: one pass per scene, outputs concatenated along the batch dimension
            if self.lazy_masks:
                raise ops._l.Pst3rError("lazy_masks: one scene per call (a batch of lazy mask handles cannot be concatenated)")
            from ..panst3r import merge_panouts
            ts = true_shape.cpu() if torch.is_tensor(true_shape) and true_shape.is_cuda else true_shape
            outs = []
            for b in range(B):
                sl = slice(b, b + 1)
                mq = None if memory_queries is None else memory_queries[:, sl]
                outs.append(self.forward(None if in_feats is None else tuple(f[sl] for f in in_feats), in_imgs[sl],
                                         None if pos is None else pos[sl], ts[sl], classes, max_bs=max_bs, outdevice=outdevice,
                                         memory_queries=mq, cat_feats=None if cat_feats is None else cat_feats[sl]))
            return merge_panouts(outs)
        pr = self.precise
        src, mask_f, grid, portrait, dev = self._stack_features(in_feats, in_imgs, true_shape, cat_feats)
        mt = self.mask_transformer
        cls_emb = self.text_encoder(classes, device=dev, precise=pr)
        if memory_queries is None:
            out = mt.forward_nhwc(src, mask_f, grid, cls_emb, deep_supervision=self.deep_supervision, portrait=portrait,
                                  precise=pr, lazy_masks=self.lazy_masks)
        else:
            logits, masks, _ = mt.prediction_heads(self._queries(memory_queries), mask_f, None, cls_emb, want_masks=True,
                                                   precise=pr, lazy=self.lazy_masks)
            out = {"pred_logits": logits[None], "pred_masks": masks[None]}
        return self._to_device(out, outdevice, dev)

    def _queries(self, memory_queries):
        mt = self.mask_transformer
        q = memory_queries.reshape(mt.num_queries, mt.hidden_dim)
        if self.precise:
            return ops.Split.from_float(q.contiguous() if q.dtype in (torch.float32, torch.bfloat16) else q.float().contiguous())
        return q if q.dtype == torch.bfloat16 else ops.to_bf16(q.float().contiguous())

    @staticmethod
    def _to_device(out, outdevice, dev):
        if outdevice is None or torch.device(outdevice) == dev:
            return out

        def mv(v):
            if torch.is_tensor(v) or hasattr(v, "materialize"):  # a LazyMasks handle materialises when it leaves the GPU
                return v.to(outdevice)
            if isinstance(v, dict):
                return {k: mv(x) for k, x in v.items()}
            return [mv(x) for x in v]
        return {k: mv(v) for k, v in out.items()}

    def _stack_features(self, in_feats, in_imgs, true_shape, cat_feats):
        """One stack of equally shaped views -> (stride-16 tokens + level_embed (V*N, 768), mask features
        (V, Hm, Wm, Cm) pixel-major, token grid, portrait flag, device), both in the landscape storage convention."""
        pr = self.precise
        if cat_feats is None:
            # producers that did not write into a shared buffer: concatenate (pure data movement).  fp32 features keep
            # their precision through the reference-precision head (split pairs); bf16 features are exact as they are.
            if pr and any(t.dtype != torch.bfloat16 for t in in_feats):
                cat_feats = torch.cat([t.float() for t in in_feats], dim=-1)
            else:
                parts = [t if t.dtype == torch.bfloat16 else ops.to_bf16(t.float().contiguous()) for t in in_feats]
                cat_feats = torch.cat(parts, dim=-1)
        B, V, N, Cc = cat_feats.shape
        if B != 1:
            raise ops._l.Pst3rError("_stack_features handles one scene (PanopticDecoder.forward loops over the batch)")
        H, W = _hw(true_shape)  # true size; portrait (H > W) views are predicted in their true orientation and
        portrait = H > W        # returned in the landscape storage convention (utils.transpose_to_landscape, dims=(2, 3))
        P = self.upscaler.patch_size
        hs, ws = H // P, W // P
        dev = cat_feats.device
        x = _head_input(cat_feats.reshape(V * N, Cc), pr)
        mt = self.mask_transformer
        if self.input_mixer is not None:
            x = self.input_mixer.forward_rows(x, V, hs, ws, precise=pr)
        if isinstance(self.upscaler, PixelShuffleUpscaler):
            src, mask_f = self.upscaler.forward_nhwc(x, V, hs, ws, f16_extra_bias=mt.level_embed.weight, precise=pr)
        else:
            imgs_t, _, _ = oriented(in_imgs.reshape(V, 3, *in_imgs.shape[-2:]), true_shape)
            src, mask_f = self.upscaler.forward_nhwc(x, imgs_t, V, hs, ws, f16_extra_bias=mt.level_embed.weight, precise=pr)
        if portrait:  # swap the spatial dims of both outputs (pure data movement); the head then sees a (ws, hs) grid
            src = _swap_grid(src, V, hs, ws)
            mask_f = _swap_grid(mask_f, V, mask_f.shape[1], mask_f.shape[2], keep4d=True)
            hs, ws = ws, hs
        return src, mask_f, (hs, ws), portrait, dev

    def _forward_multi_ar(self, in_feats, in_imgs, pos, true_shape, classes, outdevice, memory_queries, cat_feats):
        """panoptic_decoder.py:44-45, 53-76 with multi_ar=True: every argument is a list with one entry per stack of
        equally shaped views; `pred_masks` (and the aux masks) come back as lists with one (1, n_i, Q, h_i, w_i) tensor
        per stack.  LoftUp's batch-global MinMaxScaler stays per stack, as in the reference (each stack is one
        `batched_map` call, utils.py:90-160)."""
        n_st = len(true_shape)
        pr = self.precise
        cats = cat_feats if cat_feats is not None else [None] * n_st
        feats = [tuple(f[i] for f in in_feats) if in_feats is not None else None for i in range(n_st)]
        st = [self._stack_features(feats[i], in_imgs[i], true_shape[i], cats[i]) for i in range(n_st)]
        dev = st[0][4]
        mt = self.mask_transformer
        cls_emb = self.text_encoder(classes, device=dev, precise=pr)
        if memory_queries is None:
            out = mt.forward_nhwc([s_[0] for s_ in st], [s_[1] for s_ in st], [s_[2] for s_ in st], cls_emb,
                                  deep_supervision=self.deep_supervision, portrait=[s_[3] for s_ in st], precise=pr,
                                  lazy_masks=self.lazy_masks)
        else:
            logits, masks, _ = mt.prediction_heads(self._queries(memory_queries), [s_[1] for s_ in st], None, cls_emb,
                                                   want_masks=True, precise=pr, lazy=self.lazy_masks)
            out = {"pred_logits": logits[None], "pred_masks": [m_[None] for m_ in masks]}
        return self._to_device(out, outdevice, dev)


def _swap_grid(t, V: int, gh: int, gw: int, keep4d: bool = False):
    """(V*gh*gw, C) rows or (V, gh, gw, C) map -> the same tokens on the transposed (gw, gh) grid (pure data movement);
    bf16 tensors or packed ops.Split matrices."""
    if isinstance(t, ops.Split):
        buf = t.full().reshape(V, gh, gw, -1).transpose(1, 2).contiguous()
        out = _as_split(buf if keep4d else buf.view(V * gh * gw, -1))
        return out
    buf = t.reshape(V, gh, gw, -1).transpose(1, 2).contiguous()
    return buf if keep4d else buf.view(V * gh * gw, -1)


# =====================================================================================================
# v2 head pieces: InputMixer (reference model/input_mixer.py:8-29) and LoftUpUpscaler
# (reference model/upscalers/loftup.py:84-190, model/blocks.py:9-35)
# =====================================================================================================
class InputMixer(nn.Module):
    """Linear 2816 -> 768, three croco RoPE `Block`s (12 heads, LayerNorm eps 1e-5, qkv bias), LayerNorm."""

    def __init__(self, img_size, patch_size, in_dim, hidden_dim, num_heads=12, num_layers=3, ff_dim_mult=4):
        super().__init__()
        from .common import ViTBlockParams
        self.hidden_dim, self.num_heads = hidden_dim, num_heads
        self.in_proj = nn.Linear(in_dim, hidden_dim)
        self.mixer_blk = nn.ModuleList([ViTBlockParams(hidden_dim, num_heads, ff_dim_mult, eps=1e-5) for _ in range(num_layers)])
        self.mixer_norm = nn.LayerNorm(hidden_dim)

    @torch.no_grad()
    def forward_rows(self, x, V: int, hs: int, ws: int, precise: bool = False):
        """x rows (V*hs*ws, in_dim), bf16 or ops.Split -> rows (V*hs*ws, hidden_dim): bf16, or ops.Split when `precise`
        (split operands, LayerNorm in fp32, unfused reference-precision attention per view)."""
        from .common import pos_grid, rope_table, vit_block
        N = hs * ws
        D, H = self.hidden_dim, self.num_heads
        hd = D // H
        _, pos32 = pos_grid(hs, ws, x.device)
        rope = (rope_table(max(hs, ws), hd, 100.0, x.device), pos32.repeat(V, 1))
        if not precise:
            if isinstance(x, ops.Split):
                x = ops.convert(x, torch.empty(x.shape, device=x.device, dtype=torch.bfloat16))
            h = ops.gemm(x, w16(self.in_proj.weight), bias=bias_of(self.in_proj))
            for blk in self.mixer_blk:
                h, _ = vit_block(h, blk, V, N, rope)
            return ops.layernorm(h, f32(self.mixer_norm.weight), f32(self.mixer_norm.bias), 1e-5)
        dev = x.device
        h = ops.gemm(x, wsplit(self.in_proj.weight), bias=bias_of(self.in_proj), out_dtype="split")
        for blk in self.mixer_blk:
            qkv_w, qkv_b = blk.attn.qkv.weight, blk.attn.qkv.bias
            w_qk = MaskTransformer._slice_w(qkv_w, 0, 2 * D, "mix_qk", True)
            w_v = MaskTransformer._slice_w(qkv_w, 2 * D, 3 * D, "mix_v", True)
            b_qk = None if qkv_b is None else MaskTransformer._slice_w(qkv_b, 0, 2 * D, "mix_bqk")
            b_v = None if qkv_b is None else MaskTransformer._slice_w(qkv_b, 2 * D, 3 * D, "mix_bv")
            hn = ops.layernorm(h, f32(blk.norm1.weight), f32(blk.norm1.bias), blk.eps)
            qk = ops.gemm(hn, w_qk, bias=b_qk, rope=(rope[0], rope[1], 2 * D), out_dtype="split")  # (V*N, 2D), RoPE on q and k
            vT = ops.Split.empty((V, D, N), dev, align=8)  # V^T per view: the PV GEMM wants its B operand K-major
            ops.gemm(hn, w_v, bias=b_v, out=vT, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=N,
                     batch_stride=vT.hi.stride(0), ldt=vT.hi.stride(1))
            o = ops.Split.empty((V * N, D), dev)
            for v in range(V):
                r = slice(v * N, (v + 1) * N)
                attention_precise(qk[r, :D], qk[r, D:2 * D].view(N, H, hd).permute(1, 0, 2), vT[v].view(H, hd, N), hd ** -0.5,
                                  out=o[r])
            h = ops.gemm(o, wsplit(blk.attn.proj.weight), bias=bias_of(blk.attn.proj), residual=h, out=h)
            hn = ops.layernorm(h, f32(blk.norm2.weight), f32(blk.norm2.bias), blk.eps)
            mid = ops.gemm(hn, wsplit(blk.mlp.fc1.weight), bias=bias_of(blk.mlp.fc1), act=ops.ACT_GELU, out_dtype="split")
            h = ops.gemm(mid, wsplit(blk.mlp.fc2.weight), bias=bias_of(blk.mlp.fc2), residual=h, out=h)
        return ops.layernorm(h, f32(self.mixer_norm.weight), f32(self.mixer_norm.bias), 1e-5)

    def forward(self, x, pos):
        b, N, _ = x.shape
        raise ops._l.Pst3rError("call forward_rows (the CUDA PanopticDecoder does); token grids are needed for RoPE tables")


class _ImplicitFeaturizerParams(nn.Module):
    def __init__(self, dim_multiplier, n_freqs):
        super().__init__()
        self.biases = nn.Parameter(torch.randn(2, dim_multiplier, n_freqs))


class _CrossAttn(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.projq = nn.Linear(dim, dim, bias=False)
        self.projk = nn.Linear(dim, dim, bias=False)
        self.projv = nn.Linear(dim, dim, bias=False)
        self.proj = nn.Linear(dim, dim)


class CrossonlyDecoderBlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.dim, self.num_heads = dim, num_heads
        self.cross_attn = _CrossAttn(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio), dim)
        self.norm_y = nn.LayerNorm(dim)


class LoftUpUpscaler(nn.Module):
    def __init__(self, input_dim, dim, output_stride=2, patch_size=16, color_feats=True, n_freqs=20, num_heads=4,
                 num_layers=2, lr_pe_type="sine"):
        super().__init__()
        assert color_feats and lr_pe_type == "sine" and output_stride == 2
        self.input_dim, self.dim, self.patch_size, self.n_freqs, self.num_heads = input_dim, dim, patch_size, n_freqs, num_heads
        self.mask_dim = dim
        self.patch_embed = nn.Conv2d(input_dim, input_dim, kernel_size=1)
        start_dim = 5 * n_freqs * 2 + 3
        self.start_dim = start_dim
        self.lr_pe = _ImplicitFeaturizerParams(2, 5)
        self.lr_input_proj = nn.Sequential(nn.Linear(input_dim + 20, dim), nn.LayerNorm(dim))
        self.fourier_feat = nn.Sequential(nn.Identity(), _ImplicitFeaturizerParams(5, n_freqs))  # .1.biases
        self.first_conv = nn.Sequential(
            nn.GroupNorm(1, start_dim), nn.Conv2d(start_dim, dim, 3, padding=1), nn.GroupNorm(8, dim), nn.ReLU(),
            nn.Conv2d(dim, dim, 3, padding=1), nn.GroupNorm(8, dim), nn.ReLU())
        self.ca_transformer_blocks = nn.ModuleList([CrossonlyDecoderBlock(dim, num_heads, mlp_ratio=1) for _ in range(num_layers)])
        self.ca_transformer_norm = nn.LayerNorm(dim)

    # ---- prepared constants ----------------------------------------------------------------------
    @staticmethod
    def _conv_w(conv: nn.Conv2d, cpad: int, precise: bool = False):
        wt = conv.weight

        def build():
            O, Cc = wt.shape[:2]
            t = torch.zeros((O, 3, 3, cpad), device=wt.device, dtype=torch.float32)
            t[..., :Cc] = wt.detach().float().permute(0, 2, 3, 1)  # [O, C, ky, kx] -> [O, ky, kx, C]
            t = t.reshape(O, 9 * cpad)
            return _split_cat(t) if precise else t.to(torch.bfloat16).contiguous()
        buf = prepared(f"conv_w_{cpad}_{'s' if precise else 'b'}", [wt], build)
        return _as_split(buf) if precise else buf

    def _tables(self, Hh, Wh, device):
        def build():
            gy = torch.linspace(-1, 1, Hh, device=device)
            gx = torch.linspace(-1, 1, Wh, device=device)
            fr = torch.exp(torch.linspace(-2, 10, self.n_freqs, device=device))
            return torch.cat([gy, gx, fr]).float().contiguous()
        t = prepared(f"loftup_tab_{Hh}x{Wh}", [self.ca_transformer_norm.weight], build)
        return t[:Hh], t[Hh:Hh + Wh], t[Hh + Wh:]

    def _lr_pe_term(self, hs, ws, device, precise: bool = False):
        """Constant of (grid, weights): lr_input_proj.0 applied to the 20 sine-PE channels of the low-res grid, plus its
        bias -> bf16 (hs*ws, dim).  (ImplicitFeaturizer(color_feats=False, n_freqs=5, learn_bias=True), loftup.py:99.)"""
        lin = self.lr_input_proj[0]
        b = self.lr_pe.biases

        def build():
            gy = torch.linspace(-1, 1, hs, device=device).view(1, hs, 1).expand(1, hs, ws)
            gx = torch.linspace(-1, 1, ws, device=device).view(1, 1, ws).expand(1, hs, ws)
            base = torch.cat([gy, gx], 0)  # (2, hs, ws)
            fr = torch.exp(torch.linspace(-2, 10, 5, device=device)).view(5, 1, 1, 1)
            arg = base.unsqueeze(0) * fr  # (5, 2, hs, ws)
            bb = b.detach().float()
            s = torch.sin(arg + bb[0].reshape(5, 2, 1, 1)).flatten(0, 1)
            c = torch.cos(arg + bb[1].reshape(5, 2, 1, 1)).flatten(0, 1)
            pe = torch.cat([s, c], 0).flatten(1).t()  # (hs*ws, 20)
            wpe = lin.weight.detach().float()[:, self.input_dim:]
            return (pe @ wpe.t() + lin.bias.detach().float()).to(torch.float32 if precise else torch.bfloat16).contiguous()
        return prepared(f"loftup_lrpe_{hs}x{ws}_{'f' if precise else 'b'}", [lin.weight, lin.bias, b], build)

    def _lr_feat_w(self, precise: bool = False):
        lin = self.lr_input_proj[0]
        if precise:
            return _as_split(prepared("loftup_lrw_s", [lin.weight], lambda: _split_cat(lin.weight.detach()[:, :self.input_dim].float())))
        return prepared("loftup_lrw", [lin.weight], lambda: lin.weight.detach()[:, :self.input_dim].to(torch.bfloat16).contiguous())

    @torch.no_grad()
    def forward_nhwc(self, feats: torch.Tensor, imgs: torch.Tensor, b: int, hs: int, ws: int,
                     f16_extra_bias: Optional[torch.Tensor] = None, precise: bool = False):
        """feats bf16 rows (b*hs*ws, input_dim) (InputMixer output), imgs fp32 (b,3,H,W) ->
        (f16 bf16 (b*hs*ws, input_dim) = patch_embed(feats) [+ level_embed], mask feats bf16 (b, H/2, W/2, dim))."""
        dev = feats.device
        N = hs * ws
        D, Hh, Wh = self.dim, hs * self.patch_size // 2, ws * self.patch_size // 2
        W = wsplit if precise else w16
        act = "split" if precise else torch.bfloat16
        if isinstance(feats, ops.Split) and not precise:
            feats = ops.convert(feats, torch.empty(feats.shape, device=dev, dtype=torch.bfloat16))
        # stride-16 features for the query decoder: 1x1 conv == GEMM
        pb = f32(self.patch_embed.bias)
        if f16_extra_bias is not None:
            pb = prepared("loftup_f16bias", [self.patch_embed.bias, f16_extra_bias],
                          lambda: (self.patch_embed.bias.detach().float() + f16_extra_bias.detach().float().view(-1)).contiguous())
        f16 = ops.gemm(feats, W(self.patch_embed.weight), bias=pb, out_dtype=act)
        # guidance branch: x0.5 image -> MinMaxScaler (batch global) -> Fourier features -> GN(1) -> 2 x (conv3x3, GN(8), ReLU)
        half, minmax = ops.loftup_guidance(imgs.float())
        if getattr(self, "minmax_reduce", None) is not None:  # view-sharded runs: batch-global extrema across ranks
            minmax = self.minmax_reduce(minmax)
        gy, gx, fr = self._tables(Hh, Wh, dev)
        gn0, c1, gn1, c2, gn2 = (self.first_conv[i] for i in (0, 1, 2, 4, 5))
        ld0 = ((self.start_dim + 7) // 8) * 8
        x0 = ops.loftup_fourier_gn(half, minmax, gy, gx, fr, f32(self.fourier_feat[1].biases).view(-1), f32(gn0.weight),
                                   f32(gn0.bias), gn0.eps, ld0, split=precise)
        x0 = x0[..., :self.start_dim] if ld0 != self.start_dim else x0
        cpad1 = ((self.start_dim + 63) // 64) * 64
        x1 = ops.conv3x3_nhwc(x0, self._conv_w(c1, cpad1, precise), cpad1, bias=f32(c1.bias))
        ops.groupnorm_nhwc_(x1, gn1.num_groups, f32(gn1.weight), f32(gn1.bias), gn1.eps, True)
        cpad2 = ((D + 63) // 64) * 64
        x2 = ops.conv3x3_nhwc(x1, self._conv_w(c2, cpad2, precise), cpad2, bias=f32(c2.bias))
        ops.groupnorm_nhwc_(x2, gn2.num_groups, f32(gn2.weight), f32(gn2.bias), gn2.eps, True)
        x = x2.view(b * Hh * Wh, D)
        # low-res tokens: Linear([feats | sine PE]) + LN, the PE part folded into a per-token constant
        ln = self.lr_input_proj[1]
        lr = ops.gemm(feats, self._lr_feat_w(precise), residual=self._lr_pe_term(hs, ws, dev, precise), res_mod_rows=N, out_dtype=act)
        lr = ops.layernorm(lr, f32(ln.weight), f32(ln.bias), ln.eps)
        H4, hd = self.num_heads, D // self.num_heads
        P = Hh * Wh
        for blk in self.ca_transformer_blocks:
            ca = blk.cross_attn
            y_ = ops.layernorm(lr, f32(blk.norm_y.weight), f32(blk.norm_y.bias), 1e-5)
            qn = ops.layernorm(x, f32(blk.norm2.weight), f32(blk.norm2.bias), 1e-5)
            if precise:
                # unfused reference-precision cross-attention, one view at a time (49 152 high-res queries x 768 low-res
                # keys x 4 heads: the score matrix of a view is what croco's CrossAttention materialises, blocks.py:29-35)
                k = ops.gemm(y_, wsplit(ca.projk.weight), out_dtype="split")  # (b*N, D)
                vT = ops.Split.empty((b, D, N), dev, align=8)
                ops.gemm(y_, wsplit(ca.projv.weight), out=vT, store_mode=ops.STORE_TRANSPOSED, rows_per_batch=N,
                         batch_stride=vT.hi.stride(0), ldt=vT.hi.stride(1))
                q = ops.gemm(qn, wsplit(ca.projq.weight), out_dtype="split")  # (b*P, D)
                o = ops.Split.empty((b * P, D), dev)
                for v in range(b):
                    attention_precise(q[v * P:(v + 1) * P], k[v * N:(v + 1) * N].view(N, H4, hd).permute(1, 0, 2),
                                      vT[v].view(H4, hd, N), hd ** -0.5, out=o[v * P:(v + 1) * P])
            else:
                kv = ops.gemm(y_, cat_w16([ca.projk.weight, ca.projv.weight])).view(b, N, 2 * D)
                q = ops.gemm(qn, w16(ca.projq.weight))
                o = ops.attention(q.view(b, P, H4, hd), kv[:, :, :D].unflatten(-1, (H4, hd)), kv[:, :, D:].unflatten(-1, (H4, hd)))
                o = o.view(b * P, D)
            x = ops.gemm(o, W(ca.proj.weight), bias=bias_of(ca.proj), residual=x, out=x)
            h = ops.layernorm(x, f32(blk.norm3.weight), f32(blk.norm3.bias), 1e-5)
            h = ops.gemm(h, W(blk.mlp.fc1.weight), bias=bias_of(blk.mlp.fc1), act=ops.ACT_GELU, out_dtype=act)
            x = ops.gemm(h, W(blk.mlp.fc2.weight), bias=bias_of(blk.mlp.fc2), residual=x, out=x)
        x = ops.layernorm(x, f32(self.ca_transformer_norm.weight), f32(self.ca_transformer_norm.bias), 1e-5)
        return f16, x.view(b, Hh, Wh, D)

    @torch.no_grad()
    def forward(self, inputs, img_shape, precise: bool = False):
        """Reference signature: ((lr_feats (b,N,C), img (b,3,H,W)), (H, W)) -> ([patch_feats (b,C,hs,ws)], (b,dim,H/2,W/2)) fp32."""
        lr, img = inputs
        H, W = img_shape
        hs, ws = H // self.patch_size, W // self.patch_size
        b = lr.shape[0]
        x = _head_input(lr.reshape(b * hs * ws, -1), precise)
        f16, mf = self.forward_nhwc(x, img, b, hs, ws, precise=precise)
        f16_nchw = ops.nhwc_to_nchw_f32(f16.view(b, hs * ws, -1)).view(b, -1, hs, ws)
        mf_nchw = ops.nhwc_to_nchw_f32(mf.view(b, (H // 2) * (W // 2), self.dim)).view(b, self.dim, H // 2, W // 2)
        return [f16_nchw], mf_nchw
