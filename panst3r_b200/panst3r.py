"""PanSt3R façade on the CUDA modules — same surface as the reference's `PanSt3R(nn.Module)`
(reference src/panst3r/panst3r.py:19-325): constructor arguments, the four sub-module attribute names
(= state-dict prefixes), `forward`, `forward_inference_multi_ar`, `set_vocab`, `from_checkpoint`.

    forward(imgs, true_shape, classes, max_bs=None, outdevice=None) -> (panout, pointmaps)          (:286-296)
    forward_inference_multi_ar(imgs, true_shape, classes, num_keyframes=None, use_retrieval=False,
                               max_bs=None, outdevice=None, amp=False) -> (pointmaps, panout)       (:169-284)

Orchestration follows engine/must3r.py:28-69 (sequential memory build, mem_batches [2,1,1,...]) and :71-94
(render of every view against the final memory).  B200-first choices: the three feature producers write straight
into one (V, N, 2816) bf16 buffer (the torch.cat at panoptic_decoder.py:44-47 disappears); DINOv2 (throughput-bound,
needed only by the head) runs on a side stream NEXT TO the sequential memory build (latency-bound chain of small
kernels that cannot fill the GPU), each on its own share of the SMs (`dino_sms`, pst3r_set_sm_budget); `max_bs`
chunking is unnecessary (everything is one batch, results are chunk-invariant for v1) and accepted only for signature parity.
Multi-GPU: see panst3r_b200/dist.py (views sharded across ranks, one all-gather of encoder tokens).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .modules.dino import DinoV2Encoder
from .modules.must3r import MUSt3R, Dust3rEncoder
from .modules.panoptic import PanopticDecoder, PixelShuffleUpscaler

ENC_DIM, DEC_DIM, DINO_DIM = 1024, 768, 1024


def select_keyframes(n_views: int, num_keyframes):
    """(keyframes, processing order = keyframes + remaining views, k) — panst3r.py:181-196 (linspace selection)."""
    if num_keyframes is None or num_keyframes > n_views:
        num_keyframes, keyframes = n_views, list(range(n_views))
    else:
        keyframes = np.linspace(0, n_views - 1, num_keyframes, dtype=int).tolist()
    rest = sorted(set(range(n_views)).difference(set(keyframes)))
    assert len(keyframes) + len(rest) == n_views
    return keyframes, keyframes + rest, num_keyframes


def stack_views(true_shapes, stored_shapes):
    """Positions of equally shaped views grouped into stacks, in order of first appearance; positions ascend inside a
    stack (must3r `stack_views`, panst3r.py:203-206).  A view's shape key is (true (H, W), stored tensor (Hs, Ws))."""
    stacks = {}
    for p, (t, s_) in enumerate(zip(true_shapes, stored_shapes)):
        stacks.setdefault((tuple(int(v) for v in t), tuple(int(v) for v in s_)), []).append(p)
    return list(stacks.values())


def merge_panouts(outs):
    """Concatenate single-scene output dicts along the batch dimension (out_queries is (Q, B, C))."""
    if len(outs) == 1:
        return outs[0]
    m = {"pred_logits": torch.cat([o["pred_logits"] for o in outs], 0),
         "pred_masks": torch.cat([o["pred_masks"] for o in outs], 0)}
    if "aux_outputs" in outs[0]:
        m["aux_outputs"] = [{k: torch.cat([o["aux_outputs"][i][k] for o in outs], 0) for k in outs[0]["aux_outputs"][i]}
                            for i in range(len(outs[0]["aux_outputs"]))]
    if "out_queries" in outs[0]:
        m["out_queries"] = torch.cat([o["out_queries"] for o in outs], 1)
    return m


class PanSt3R(nn.Module):
    def __init__(self, must3r_encoder: nn.Module, must3r_decoder: nn.Module, dino_encoder: nn.Module,
                 panoptic_decoder: nn.Module, retrieval=None, preserve_gpu_mem: bool = False,
                 postprocess_default: str = "standard_v2", qubo_enabled: bool = True,
                 must3r_encoder_requires_grad=False, must3r_decoder_requires_grad=False, verbose: bool = False):
        super().__init__()
        self.must3r_encoder = must3r_encoder
        self.must3r_decoder = must3r_decoder
        self.dino_encoder = dino_encoder
        self.panoptic_decoder = panoptic_decoder
        self.retrieval = retrieval
        self.preserve_gpu_mem = preserve_gpu_mem
        self.verbose = verbose
        self.must3r_params = dict(init_num_views=2, batch_num_views=1, render_iterations=1)
        self.postprocess_default = postprocess_default
        self.qubo_enabled = qubo_enabled
        self.overlap_dino = True
        # SMs given to DINOv2 while it runs next to the memory build (the build gets the rest); 0 = no partition.
        # Measured (profiles/r01_stage_times.md): partitioning costs more than it hides on one GPU, plain fork wins.
        self.dino_sms = 0
        self._side_stream: Optional[torch.cuda.Stream] = None

    # ---- reference helper methods ----------------------------------------------------------------
    def get_must3r_mem_batches(self, n_imgs):
        mem_batches = [self.must3r_params["init_num_views"]]
        while (s := sum(mem_batches)) != n_imgs:
            mem_batches.append(min(self.must3r_params["batch_num_views"], n_imgs - s))
        return mem_batches

    def forward_dino(self, imgs, true_shape, max_bs=None, verbose=None, out=None):
        B, V = imgs.shape[:2]
        x = self.dino_encoder(imgs.flatten(0, 1), true_shape.flatten(0, 1), out=out)
        return x.unflatten(0, (B, V))

    def forward_must3r_encoder(self, imgs, true_shape, max_bs=None, out=None):
        B, V = imgs.shape[:2]
        x, pos = self.must3r_encoder(imgs.flatten(0, 1), true_shape.flatten(0, 1), out=out)
        return x.unflatten(0, (B, V)), pos.unflatten(0, (B, V))

    def build_memory(self, x, pos, true_shape):
        """engine/must3r.py:28-69 — V-1 dependent decoder passes; first-pass pointmaps are discarded by the caller
        (panst3r.py:77) so they are not computed."""
        self.must3r_decoder.reserve_views = x.shape[1]
        edges = [0] + np.cumsum(self.get_must3r_mem_batches(x.shape[1])).tolist()
        mem = None
        for a, b in zip(edges[:-1], edges[1:]):
            mem, _, _ = self.must3r_decoder(x[:, a:b], pos[:, a:b], true_shape[:, a:b], mem, render=False,
                                            return_feats=False, compute_pointmaps=False)
        return mem

    def forward_must3r_decoder(self, x_must3r, pos_must3r, true_shape, max_bs=None, feats_out=None):
        mem = self.build_memory(x_must3r, pos_must3r, true_shape)
        _, pointmaps, feats = self.must3r_decoder(x_must3r, pos_must3r, true_shape, mem, render=True,
                                                  return_feats="last", feats_out=feats_out)
        return feats[-1], pointmaps, mem

    # ---- shared core -----------------------------------------------------------------------------
    def _features(self, imgs, true_shape):
        """Run DINOv2 (side stream) + MUSt3R encoder into one concatenated (B, V, N, 2816) bf16 buffer."""
        B, V, _, H, W = imgs.shape
        if not imgs.is_cuda or not next(self.parameters()).is_cuda:
            raise ops._l.Pst3rError("PanSt3R (panst3r_b200) runs on CUDA sm_100 only: move the module and inputs to the "
                                    "GPU; there is no CPU fallback")
        if B != 1:
            raise ops._l.Pst3rError("the CUDA path processes one scene per call (B == 1)")
        P = self.must3r_encoder.patch_size
        N = (H // P) * (W // P)
        cat = torch.empty((B, V, N, ENC_DIM + DEC_DIM + DINO_DIM), device=imgs.device, dtype=torch.bfloat16)
        rows = cat.view(B * V * N, -1)
        cur = torch.cuda.current_stream()
        x, pos = self.forward_must3r_encoder(imgs, true_shape, out=rows[:, :ENC_DIM])
        if self.overlap_dino:  # also legal under CUDA-graph capture: the side stream forks from / joins the capturing one
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream()
            side = self._side_stream
            side.wait_stream(cur)  # fork after the encoder: DINOv2 shares the GPU with the memory build that follows
            with torch.cuda.stream(side), ops.sm_budget(self.dino_sms):
                self.forward_dino(imgs, true_shape, out=rows[:, ENC_DIM + DEC_DIM:])
            join = side
        else:
            self.forward_dino(imgs, true_shape, out=rows[:, ENC_DIM + DEC_DIM:])
            join = None
        return cat, rows, x, pos, join

    def _build_memory_shared(self, x, pos, true_shape, join):
        """Memory build on the SMs DINOv2 leaves free; joins the side stream before returning so that the
        full-width render pass that follows has the GPU to itself."""
        if join is None or not self.dino_sms:
            return self.build_memory(x, pos, true_shape)
        with ops.sm_budget(max(ops.num_sms() - self.dino_sms, 8)):
            mem = self.build_memory(x, pos, true_shape)
        torch.cuda.current_stream().wait_stream(join)
        return mem

    @torch.no_grad()
    def forward(self, imgs, true_shape, classes, max_bs=None, outdevice=None):
        """imgs fp32 (1, V, 3, H, W) in [-1, 1] on CUDA; true_shape (1, V, 2) (H, W); returns (panout, pointmaps)."""
        ts = true_shape.cpu() if true_shape.is_cuda else true_shape
        if imgs.shape[0] > 1:
            # batch-generic like the reference (panst3r.py:286-296; engine/must3r.py:95-108 keeps one memory per
            # scene): scenes are independent, so a batch is a loop over single-scene passes
            outs = [self.forward(imgs[b:b + 1], ts[b:b + 1], classes, max_bs=max_bs, outdevice=outdevice)
                    for b in range(imgs.shape[0])]
            return merge_panouts([o[0] for o in outs]), torch.cat([o[1] for o in outs], 0)
        cat, rows, x, pos, join = self._features(imgs, ts)
        mem = self._build_memory_shared(x, pos, ts, join)
        _, pointmaps, _ = self.must3r_decoder(x, pos, ts, mem, render=True, return_feats="last",
                                              feats_out=rows[:, ENC_DIM:ENC_DIM + DEC_DIM])
        if join is not None:
            torch.cuda.current_stream().wait_stream(join)
        panout = self.panoptic_decoder(None, imgs, pos, ts, classes, outdevice=outdevice, cat_feats=cat)
        if outdevice is not None:
            pointmaps = pointmaps.to(outdevice)
        return panout, pointmaps

    @torch.no_grad()
    def forward_panoptic(self, imgs, true_shape, classes, postprocess: Optional[str] = None, max_bs=None, **pp_kwargs):
        """`forward()` followed by the reference's standard post-processing (tools/demo_panst3r.py:232-242:
        `panoptic_inference_v1/_v2(pan_out['pred_logits'], pan_out['pred_masks'], size)`) for callers that want segment
        ids, not mask logits: the head leaves its final mask einsum to the post-processing (`PanopticDecoder.lazy_masks`),
        which evaluates it band by band through the L2 (panst3r_b200/postprocess.py, LazyMasks) — the (V, Q, H/2, W/2)
        fp32 tensor (629 MB at 16 views of 512x384) is never written, re-read or copied to the host.
        Returns (pan_preds, panout, pointmaps); pan_preds[b] = {'pan', 'segments_info', 'conf'} as the reference's
        functions return it, panout['pred_masks'] is the LazyMasks handle (`.materialize()` gives the tensor)."""
        from . import postprocess as pp
        which = postprocess or self.postprocess_default
        fns = {"standard_v1": pp.panoptic_inference_v1, "standard_v2": pp.panoptic_inference_v2}
        if which not in fns:
            raise NotImplementedError(f"postprocess={which!r}: only 'standard_v1' / 'standard_v2' run on the GPU "
                                      "(panoptic_inference_qubo is out of scope)")
        pd = self.panoptic_decoder
        prev, pd.lazy_masks = pd.lazy_masks, True
        try:
            panout, pointmaps = self.forward(imgs, true_shape, classes, max_bs=max_bs)
        finally:
            pd.lazy_masks = prev
        ts = true_shape.cpu() if torch.is_tensor(true_shape) and true_shape.is_cuda else true_shape
        size = tuple(int(v) for v in ts[0][0])
        pan_preds = fns[which](panout["pred_logits"], panout["pred_masks"], size, label_mode=pd.label_mode, **pp_kwargs)
        return pan_preds, panout, pointmaps

    @torch.no_grad()
    def forward_inference_multi_ar(self, imgs: List[torch.Tensor], true_shape, classes, num_keyframes=None,
                                   use_retrieval=False, max_bs=None, outdevice=None, amp=False):
        """Keyframes build the memory and run the full panoptic head; the remaining frames are rendered against the
        frozen memory and decoded with the keyframes' final queries (panst3r.py:169-284, panoptic_decoder.py:70-76).
        Views may differ in aspect ratio / orientation: they are grouped into stacks of equal `true_shape`
        (must3r `stack_views`, panst3r.py:203-206, 257-263), every per-view stage runs once per stack, the memory
        build walks the keyframes in order (the two views of the initialisation pair must share a shape)."""
        if use_retrieval:
            raise NotImplementedError("retrieval keyframe selection needs asmk + a retrieval checkpoint (out of scope)")
        N = len(imgs)
        keyframes, order, k = select_keyframes(N, num_keyframes)
        ts_o = (true_shape.cpu() if true_shape.is_cuda else true_shape)[order]
        ims = [imgs[i] for i in order]
        # ---- stacks of equally shaped views (positions in the reordered list, ascending: keyframes first)
        stacks = stack_views(ts_o.tolist(), [im.shape[-2:] for im in ims])
        where = {p: (si, j) for si, idx in enumerate(stacks) for j, p in enumerate(idx)}
        st = []
        join = None
        for idx in stacks:  # per-view producers (encoder, DINOv2 on the side stream) per stack
            im = torch.stack([ims[p] for p in idx])[None]
            ts_s = ts_o[idx][None]
            cat, rows, x, pos, join = self._features(im, ts_s)
            st.append(dict(idx=idx, im=im, ts=ts_s, cat=cat, rows=rows, x=x, pos=pos, kc=sum(1 for p in idx if p < k)))
        # ---- sequential memory build over the keyframes (engine/must3r.py:28-69)
        self.must3r_decoder.reserve_views = k
        edges = [0] + np.cumsum(self.get_must3r_mem_batches(k)).tolist()
        mem = None
        for a, b in zip(edges[:-1], edges[1:]):
            sel = [where[p] for p in range(a, b)]
            xs = [st[si]["x"][:, j] for si, j in sel]
            if any(t.shape != xs[0].shape for t in xs) or any(not torch.equal(st[si]["ts"][0, j], st[sel[0][0]]["ts"][0, sel[0][1]])
                                                              for si, j in sel):
                raise ops._l.Pst3rError("the views of one memory-update batch (the initialisation pair) must share a shape")
            si0 = sel[0][0]
            if len(sel) == 1 or all(si == si0 for si, _ in sel) and [j for _, j in sel] == list(range(sel[0][1], sel[0][1] + len(sel))):
                j0 = sel[0][1]
                xb, pb, tb = st[si0]["x"][:, j0:j0 + len(sel)], st[si0]["pos"][:, j0:j0 + len(sel)], st[si0]["ts"][:, j0:j0 + len(sel)]
            else:
                xb = torch.stack(xs, 1)
                pb = torch.stack([st[si]["pos"][:, j] for si, j in sel], 1)
                tb = torch.stack([st[si]["ts"][:, j] for si, j in sel], 1)
            mem, _, _ = self.must3r_decoder(xb, pb, tb, mem, render=False, return_feats=False, compute_pointmaps=False)
        # ---- render every view against the final memory, stack by stack (identical per-view results)
        for s_ in st:
            _, s_["pm"], _ = self.must3r_decoder(s_["x"], s_["pos"], s_["ts"], mem, render=True, return_feats="last",
                                                 feats_out=s_["rows"][:, ENC_DIM:ENC_DIM + DEC_DIM])
        if join is not None:
            torch.cuda.current_stream().wait_stream(join)
        # ---- panoptic head: keyframes (the leading kc views of each stack) with the full query decoder ...
        kf = [s_ for s_ in st if s_["kc"] > 0]
        if len(st) == 1:
            s_ = st[0]
            kc = s_["kc"]
            pan_kf = self.panoptic_decoder(None, s_["im"][:, :kc], s_["pos"][:, :kc], s_["ts"][:, :kc], classes,
                                           cat_feats=s_["cat"][:, :kc])
            kf_masks = [pan_kf["pred_masks"]]
        else:
            pan_kf = self.panoptic_decoder(None, [s_["im"][:, :s_["kc"]] for s_ in kf], [s_["pos"][:, :s_["kc"]] for s_ in kf],
                                           [s_["ts"][:, :s_["kc"]] for s_ in kf], classes, multi_ar=True,
                                           cat_feats=[s_["cat"][:, :s_["kc"]] for s_ in kf])
            kf_masks = pan_kf["pred_masks"]
        masks = [None] * N
        for s_, mk in zip(kf, kf_masks):
            for j in range(s_["kc"]):
                masks[s_["idx"][j]] = mk[0, j]
        # ... and the remaining frames with the keyframes' final queries (memory-query path)
        nk = [s_ for s_ in st if s_["kc"] < len(s_["idx"])]
        if nk:
            if len(st) == 1:
                s_ = nk[0]
                kc = s_["kc"]
                pan_nk = self.panoptic_decoder(None, s_["im"][:, kc:], s_["pos"][:, kc:], s_["ts"][:, kc:], classes,
                                               cat_feats=s_["cat"][:, kc:], memory_queries=pan_kf["out_queries"])
                nk_masks = [pan_nk["pred_masks"]]
            else:
                pan_nk = self.panoptic_decoder(None, [s_["im"][:, s_["kc"]:] for s_ in nk], [s_["pos"][:, s_["kc"]:] for s_ in nk],
                                               [s_["ts"][:, s_["kc"]:] for s_ in nk], classes, multi_ar=True,
                                               cat_feats=[s_["cat"][:, s_["kc"]:] for s_ in nk],
                                               memory_queries=pan_kf["out_queries"])
                nk_masks = pan_nk["pred_masks"]
            for s_, mk in zip(nk, nk_masks):
                for j in range(s_["kc"], len(s_["idx"])):
                    masks[s_["idx"][j]] = mk[0, j - s_["kc"]]
        pms = [None] * N
        for s_ in st:
            for j, p in enumerate(s_["idx"]):
                pms[p] = s_["pm"][0, j]
        inv = np.argsort(order)
        panout = {"pred_logits": pan_kf["pred_logits"], "pred_masks": [masks[i] for i in inv],
                  "out_queries": pan_kf["out_queries"]}
        pms = [pms[i] for i in inv]
        if outdevice is not None:
            pms = [p.to(outdevice) for p in pms]
            panout["pred_masks"] = [m.to(outdevice) for m in panout["pred_masks"]]
        return pms, panout

    def set_vocab(self, class_names, device=None):
        self.panoptic_decoder.text_encoder.set_vocab(class_names, device=device)

    # upstream parameter-name variants -> the names this package (and oracle/must3r.py) uses.  The MUSt3R classes live in
    # the un-vendored `must3r` package (pyproject.toml:14): croco-style names are accepted next to ours.
    KEY_REMAP = (
        ("must3r_encoder.enc_blocks.", "must3r_encoder.blocks_enc."),
        ("must3r_encoder.enc_norm.", "must3r_encoder.norm_enc."),
        ("must3r_decoder.dec_blocks.", "must3r_decoder.blocks_dec."),
        ("must3r_decoder.dec_norm.", "must3r_decoder.norm_dec."),
        ("must3r_decoder.decoder_embed.", "must3r_decoder.feat_embed_enc_to_dec."),
        ("must3r_decoder.feedback.", "must3r_decoder.feedback_layer."),
        ("must3r_decoder.head.proj.", "must3r_decoder.head_dec.proj."),
        ("must3r_decoder.downstream_head.proj.", "must3r_decoder.head_dec.proj."),
    )

    @classmethod
    def remap_state_dict(cls, weights: dict) -> dict:
        out = {}
        for k, v in weights.items():
            for old, new in cls.KEY_REMAP:
                if k.startswith(old):
                    k = new + k[len(old):]
                    break
            out[k] = v
        return out

    def load_checkpoint_weights(self, weights: dict, allow_partial: bool = False) -> dict:
        """load_state_dict(strict=False) as the reference does (panst3r.py:323) — but never silently: returns (and keeps
        in `self.load_report`) the missing / unexpected keys per sub-module, warns about them, and raises when a whole
        sub-module received nothing (a name mismatch that would leave it at random init) unless allow_partial."""
        import warnings
        res = self.load_state_dict(self.remap_state_dict(weights), strict=False)
        own = list(self.state_dict().keys())
        report = {}
        for sub in ("must3r_encoder", "must3r_decoder", "dino_encoder", "panoptic_decoder"):
            total = sum(1 for k in own if k.startswith(sub + "."))
            miss = [k for k in res.missing_keys if k.startswith(sub + ".")]
            unexp = [k for k in res.unexpected_keys if k.startswith(sub + ".")]
            report[sub] = {"parameters": total, "missing": miss, "unexpected": unexp}
        report["other_unexpected"] = [k for k in res.unexpected_keys
                                      if not k.startswith(("must3r_encoder.", "must3r_decoder.", "dino_encoder.", "panoptic_decoder."))]
        self.load_report = report
        bad = [s_ for s_ in report if isinstance(report[s_], dict) and report[s_]["parameters"] and
               len(report[s_]["missing"]) == report[s_]["parameters"]]
        for s_, r in report.items():
            if isinstance(r, dict) and (r["missing"] or r["unexpected"]):
                warnings.warn(f"PanSt3R checkpoint: {s_}: {len(r['missing'])}/{r['parameters']} parameters missing "
                              f"(e.g. {r['missing'][:3]}), {len(r['unexpected'])} unexpected (e.g. {r['unexpected'][:3]})")
        if bad and not allow_partial:
            raise ops._l.Pst3rError(f"checkpoint loaded NO parameter of {bad}: its key names do not match this package "
                                    f"(see PanSt3R.KEY_REMAP / model.load_report); pass allow_partial=True to proceed anyway")
        return report

    @classmethod
    def from_checkpoint(cls, checkpoint_path, retrieval_path=None, allow_partial: bool = False):
        """Loads the reference's checkpoint format: {'args': Namespace of constructor strings, 'weights': state dict}
        (panst3r.py:301-325).  The constructor strings are evaluated against this package's CUDA classes."""
        ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        assert "args" in ckpt, "Checkpoint must contain 'args' with model parameters."
        from .modules import panoptic as _p
        ns = {"Dust3rEncoder": Dust3rEncoder, "MUSt3R": MUSt3R, "DinoV2Encoder": DinoV2Encoder,
              "PanopticDecoder": PanopticDecoder, "PixelShuffleUpscaler": PixelShuffleUpscaler,
              "InputMixer": _p.InputMixer, "LoftUpUpscaler": _p.LoftUpUpscaler}
        a = ckpt["args"]
        get = (lambda k, d=None: a.get(k, d)) if isinstance(a, dict) else (lambda k, d=None: getattr(a, k, d))
        model = cls(must3r_encoder=eval(get("must3r_encoder"), ns), must3r_decoder=eval(get("must3r_decoder"), ns),
                    dino_encoder=eval(get("dino_encoder"), ns), panoptic_decoder=eval(get("panoptic_decoder"), ns),
                    retrieval=ckpt.get("retrieval"),
                    postprocess_default=get("postprocess_default", "standard_v2"),
                    qubo_enabled=get("qubo_enabled", True))
        model.load_checkpoint_weights(ckpt["weights"], allow_partial=allow_partial)
        return model


def build_panst3r(variant: str = "v1", enc_depth=24, dec_depth=12, dino_depth=24, mixer_layers=3,
                  head_precision: str = "fp32") -> PanSt3R:
    """Reference configuration (configs/base.yaml, base_v2.yaml) with reducible depths for tests.  head_precision:
    "fp32" = the reference's policy for the panoptic head (panst3r.py:236-245), "bf16" = plain bf16 operands."""
    enc = Dust3rEncoder(depth=enc_depth)
    dec = MUSt3R(depth=dec_depth)
    dino = DinoV2Encoder(depth=dino_depth)
    if variant == "v1":
        pd = PanopticDecoder(upscaler=PixelShuffleUpscaler(input_dim=ENC_DIM + DEC_DIM + DINO_DIM), precision=head_precision)
    elif variant == "v2":
        from .modules.panoptic import InputMixer, LoftUpUpscaler
        pd = PanopticDecoder(input_mixer=InputMixer([512, 512], 16, 2816, 768, num_layers=mixer_layers),
                             upscaler=LoftUpUpscaler(input_dim=768, dim=384), mask_dim=384, precision=head_precision)
    else:
        raise ValueError(variant)
    return PanSt3R(enc, dec, dino, pd).eval()
