"""CUDA MUSt3R encoder / decoder behind the reference's two operator seams.

    x, pos = encoder(img, true_shape)                                    (reference engine/must3r.py:17-24)
    mem, pointmaps, feats = decoder(x, pos, true_shape, mem, render=, return_feats=)   (:45, :93, :116)

Same constructor arguments, parameter names and memory 5-tuple as oracle/must3r.py (the restatement of the
un-vendored upstream classes named at reference configs/base.yaml:7-15); all arithmetic runs in
libpanst3r_b200.so (tcgen05 GEMMs with fused bias/RoPE/GELU/residual epilogues, tcgen05 flash attention).
B200-first differences from a literal translation:
  * the memory bank stores, next to the raw memory tokens, their K/V projections per layer — projected once
    when tokens are appended instead of on every decoder call;
  * all views of a render call attend one shared K/V copy (kv batch stride 0) instead of an expanded memory;
  * RoPE is applied in the QKV GEMM epilogue; the pointmap pixel-shuffle is the GEMM's store pattern.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .. import ops
from .common import (ViTBlockParams, bias_of, cat_f32, cat_w16, f32, fold_stats, ln_linear, mlp_residual, pos_grid,
                     prepared, rope_table, self_attention, vit_block, w16)


def _hw(true_shape) -> tuple:
    """True (H, W) of the views of a call (all equal).  H > W: portrait views — by the reference's storage convention
    (model/dino.py:25-33, utils.py:36-49) their image / dense-output tensors are kept transposed (landscape)."""
    ts = true_shape.reshape(-1, 2)
    if ts.is_cuda:
        ts = ts.cpu()
    H, W = int(ts[0, 0]), int(ts[0, 1])
    if not bool((ts == ts[0:1]).all()):
        raise ops._l.Pst3rError("all views of a call must share one true_shape (stack views by aspect ratio first)")
    return H, W


def oriented(img: torch.Tensor, true_shape) -> tuple:
    """Stored image batch (b, 3, Hs, Ws) -> (image batch in the true orientation, H, W).  Portrait views are stored
    transposed; transposing back is pure data movement."""
    H, W = _hw(true_shape)
    Hs, Ws = img.shape[-2:]
    if (Hs, Ws) == (H, W):
        return img, H, W
    if (Hs, Ws) == (W, H):
        return img.transpose(-1, -2).contiguous(), H, W
    raise ops._l.Pst3rError(f"image tensor {Hs}x{Ws} matches neither true_shape {H}x{W} nor its transpose")


class Dust3rEncoder(nn.Module):
    def __init__(self, img_size=(512, 512), patch_embed="PatchEmbedDust3R", patch_size=16, embed_dim=1024, depth=24,
                 num_heads=16, mlp_ratio=4.0, pos_embed="RoPE100"):
        super().__init__()
        assert patch_embed == "PatchEmbedDust3R" and pos_embed.startswith("RoPE")
        self.patch_size, self.embed_dim, self.num_heads = patch_size, embed_dim, num_heads
        self.rope_base = float(pos_embed[4:])
        self.patch_embed = nn.Module()
        self.patch_embed.proj = nn.Conv2d(3, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.blocks_enc = nn.ModuleList([ViTBlockParams(embed_dim, num_heads, mlp_ratio, eps=1e-6) for _ in range(depth)])
        self.norm_enc = nn.LayerNorm(embed_dim, eps=1e-6)
        self.fold_ln = False  # per-view stage: stand-alone LayerNorm kernels (see common.FOLD_LN_MAX_ROWS)

    @torch.no_grad()
    def forward(self, img: torch.Tensor, true_shape, out: Optional[torch.Tensor] = None):
        """img fp32 (b,3,H,W) in [-1,1] -> (x bf16 (b,N,D), pos int64 (b,N,2)).  `out`: optional bf16 destination
        (b*N rows, D columns, any row stride) for the normalised tokens."""
        img, H, W = oriented(img, true_shape)
        b = img.shape[0]
        P, D = self.patch_size, self.embed_dim
        h, w = H // P, W // P
        N = h * w
        pos64, pos32 = pos_grid(h, w, img.device)
        pos_rows = pos32.repeat(b, 1) if b > 1 else pos32
        rope = (rope_table(max(h, w), D // self.num_heads, self.rope_base, img.device), pos_rows)
        a = ops.patchify(img.float(), P)
        # every LayerNorm that feeds a Linear is folded into that GEMM: the GEMM producing the residual stream leaves
        # per-row partial sums behind (stats), the consuming GEMM normalises in its epilogue
        s1, s2 = fold_stats(b * N, D, img.device, 2, enable=self.fold_ln)
        x = ops.gemm(a, w16(self.patch_embed.proj.weight), bias=f32(self.patch_embed.proj.bias), stats_out=s2)
        st = s2
        for blk in self.blocks_enc:
            x, st = vit_block(x, blk, b, N, rope, stats=st, stats_pair=(s1, s2))
        if out is None:
            out = torch.empty((b * N, D), device=img.device, dtype=torch.bfloat16)
        ops.layernorm(x, f32(self.norm_enc.weight), f32(self.norm_enc.bias), 1e-6, out=out)
        xo = out.view(b, N, D) if out.is_contiguous() else out.unflatten(0, (b, N))
        return xo, pos64[None].expand(b, N, 2)


class _DecBlock(ViTBlockParams):
    """norm1/attn/norm2 from ViTBlockParams + cross_attn.{projq,projk,projv,proj}, norm3, mlp, norm_y."""

    def __init__(self, dim, num_heads, mlp_ratio):
        super().__init__(dim, num_heads, mlp_ratio, eps=1e-6)
        self.cross_attn = nn.Module()
        for n in ("projq", "projk", "projv", "proj"):
            setattr(self.cross_attn, n, nn.Linear(dim, dim))
        self.norm3 = nn.LayerNorm(dim, eps=1e-6)
        self.norm_y = nn.LayerNorm(dim, eps=1e-6)

    def kv_weight(self):
        return cat_w16([self.cross_attn.projk.weight, self.cross_attn.projv.weight])

    def kv_bias(self):
        return cat_f32([self.cross_attn.projk.bias, self.cross_attn.projv.bias])


class MemoryBank:
    """Memory tokens (L, B, cap, D) and their K|V projections (L, B, cap, 2D) of all decoder layers, bf16,
    capacity-doubling.  One allocation per kind so that the per-layer appends of a memory update are ONE
    strided-batched LayerNorm + ONE strided-batched GEMM; tok[l] / kv[l] are the per-layer (B, cap, .) views."""

    def __init__(self, B: int, depth: int, dim: int, device, capacity: int):
        self.B, self.depth, self.dim, self.n = B, depth, dim, 0
        self.cap = max(capacity, 1)
        self._alloc(device)

    def _alloc(self, device):
        self.tok_all = torch.empty((self.depth, self.B, self.cap, self.dim), device=device, dtype=torch.bfloat16)
        self.kv_all = torch.empty((self.depth, self.B, self.cap, 2 * self.dim), device=device, dtype=torch.bfloat16)
        self.tok = [self.tok_all[l] for l in range(self.depth)]
        self.kv = [self.kv_all[l] for l in range(self.depth)]

    def fork(self, n: int) -> "MemoryBank":
        """Independent bank holding the first n tokens (copy-on-write for non-tail appends)."""
        nb = MemoryBank(self.B, self.depth, self.dim, self.tok_all.device, self.cap)
        nb.tok_all[:, :, :n].copy_(self.tok_all[:, :, :n])
        nb.kv_all[:, :, :n].copy_(self.kv_all[:, :, :n])
        nb.n = n
        return nb

    def reserve(self, n_total: int):
        if n_total <= self.cap:
            return
        old_tok, old_kv = self.tok_all, self.kv_all
        self.cap = max(n_total, 2 * self.cap)
        self._alloc(old_tok.device)
        self.tok_all[:, :, :self.n].copy_(old_tok[:, :, :self.n])
        self.kv_all[:, :, :self.n].copy_(old_kv[:, :, :self.n])


class MemVals(list):
    """`mem_vals` of the reference's memory tuple: a list of (B, Nmem, D) tensors, plus the bank that owns them."""
    bank: Optional[MemoryBank] = None


class MUSt3R(nn.Module):
    def __init__(self, img_size=(512, 512), feedback_type="single_mlp", memory_mode="norm_y", enc_embed_dim=1024,
                 embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, patch_size=16, pos_embed="RoPE100",
                 head_channels=7):
        super().__init__()
        assert feedback_type == "single_mlp" and memory_mode == "norm_y"
        self.embed_dim, self.depth, self.num_heads, self.patch_size = embed_dim, depth, num_heads, patch_size
        self.head_channels = head_channels
        self.rope_base = float(pos_embed[4:])
        self.feat_embed_enc_to_dec = nn.Linear(enc_embed_dim, embed_dim)
        self.image2_embed = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.blocks_dec = nn.ModuleList([_DecBlock(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.feedback_layer = nn.Module()
        self.feedback_layer.fc1 = nn.Linear(embed_dim, int(mlp_ratio * embed_dim))
        self.feedback_layer.fc2 = nn.Linear(int(mlp_ratio * embed_dim), embed_dim)
        self.norm_dec = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head_dec = nn.Module()
        self.head_dec.proj = nn.Linear(embed_dim, head_channels * patch_size * patch_size)
        self.reserve_views = 0  # capacity hint (views) for the memory bank, set by the caller
        self.fold_ln_render = False  # LayerNorm folding is for the sequential memory build; render is a per-view stage

    # ---- prepared (fused / permuted) parameters -------------------------------------------------
    def _embed_bias(self, tagged: bool):
        b, e = self.feat_embed_enc_to_dec.bias, self.image2_embed
        if not tagged:
            return f32(b)
        return prepared("embed_bias_tagged", [b, e], lambda: (b.detach().float() + e.detach().float().view(-1)).contiguous())

    def _head_weight(self):
        P, C, D = self.patch_size, self.head_channels, self.embed_dim
        wt = self.head_dec.proj.weight
        # reference row order (c, i, j) -> (i, j, c): the D2S epilogue then writes C-contiguous pixels
        return prepared("head_w", [wt], lambda: wt.detach().view(C, P, P, D).permute(1, 2, 0, 3).reshape(P * P * C, D)
                        .to(torch.bfloat16).contiguous())

    def _stacked(self, what: str):
        """Per-layer parameters of the memory update stacked along a leading layer axis (batched launches)."""
        blks = list(self.blocks_dec)
        if what == "kv_w":
            ps = [p for b in blks for p in (b.cross_attn.projk.weight, b.cross_attn.projv.weight)]
            return prepared("stk_kv_w", ps, lambda: torch.stack([b.kv_weight() for b in blks]).contiguous())
        if what == "kv_b":
            ps = [p for b in blks for p in (b.cross_attn.projk.bias, b.cross_attn.projv.bias)]
            return prepared("stk_kv_b", ps, lambda: torch.stack([b.kv_bias() for b in blks]).contiguous())
        attr = {"ny_g": "weight", "ny_b": "bias"}[what]
        ps = [getattr(b.norm_y, attr) for b in blks]
        return prepared("stk_" + what, ps, lambda: torch.stack([p.detach().float() for p in ps]).contiguous())

    def _head_bias(self):
        P, C = self.patch_size, self.head_channels
        bs = self.head_dec.proj.bias
        return prepared("head_b", [bs], lambda: bs.detach().view(C, P, P).permute(1, 2, 0).reshape(-1).float().contiguous())

    # ---- pieces -----------------------------------------------------------------------------------
    def _cross_attention(self, x, blk: _DecBlock, B, n, N, kv_list, mask_bits, stats=None, stats_out=None):
        """x += proj(attn(q = projq(norm2 x), K|V));  kv_list[b]: bf16 (1, Nk, 2D) shared by the n views of batch b.
        stats / stats_out: LayerNorm statistics of x (norm2 folds into the q projection) / of the result."""
        D, H = self.embed_dim, self.num_heads
        hd = D // H
        q = ln_linear(x, stats, blk.norm2, [blk.cross_attn.projq.weight], [blk.cross_attn.projq.bias], 1e-6).view(B, n, N, H, hd)
        o = torch.empty((B * n, N, D), device=x.device, dtype=torch.bfloat16)
        for b in range(B):
            kv = kv_list[b]
            Nk = kv.shape[1]
            k = kv[:, :, :D].unflatten(-1, (H, hd))
            v = kv[:, :, D:].unflatten(-1, (H, hd))
            ops.attention(q[b], k, v, mask_bits=mask_bits, out=o[b * n:(b + 1) * n])
        return ops.gemm(o.view(B * n * N, D), w16(blk.cross_attn.proj.weight), bias=bias_of(blk.cross_attn.proj),
                        residual=x, out=x, stats_out=stats_out)

    @staticmethod
    def _own_view_mask(n: int, N: int, n_mem: int, device) -> torch.Tensor:
        """int32 (n, 1, W): view i may not attend to candidate keys [n_mem + i*N, n_mem + (i+1)*N) (its own tokens)."""
        nk = n_mem + n * N
        words = ((nk + 127) // 128) * 4
        kidx = torch.arange(words * 32, device=device)[None]
        lo = (torch.arange(n, device=device) * N + n_mem)[:, None]
        blocked = (kidx >= lo) & (kidx < lo + N)
        packed = (blocked.view(n, words, 32).to(torch.int64) << torch.arange(32, device=device)).sum(-1)
        packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed).to(torch.int32)
        return packed.view(n, 1, words).contiguous()

    @torch.no_grad()
    def forward(self, x, pos, true_shape, mem=None, render=False, return_feats=False, compute_pointmaps=True,
                feats_out: Optional[torch.Tensor] = None):
        """x bf16/fp32 (B, n, N, Denc); pos (B, n, N, 2); true_shape (B, n, 2).
        Returns (mem, pointmaps fp32 (B, n, H, W, C) | None, feats list | None).
        return_feats: True -> every layer's tokens (feats[-1] = last block output); 'last' -> [last] only.
        feats_out: optional bf16 destination rows (B*n*N, D) for the last block's output."""
        B, n, N, Denc = x.shape
        D, L = self.embed_dim, self.depth
        H, W = _hw(true_shape)
        dev = x.device
        h_, w_ = H // self.patch_size, W // self.patch_size
        rows = B * n * N
        xin = x if x.dtype == torch.bfloat16 else ops.to_bf16(x.float().contiguous().view(rows, Denc))
        xin = xin.reshape(B, n, N, Denc)
        _, pos32 = pos_grid(h_, w_, dev)
        rope = (rope_table(max(h_, w_), D // self.num_heads, self.rope_base, dev), pos32.repeat(B * n, 1))

        # enc -> dec embedding; every view except the scene's first is tagged with image2_embed (fused as a bias)
        # memory update: the inputs of all L layers live in one (L, rows, D) stack (batched norm_y / K|V append)
        stack = None if render else torch.empty((L, rows, D), device=dev, dtype=torch.bfloat16)
        hx = torch.empty((B, n, N, D), device=dev, dtype=torch.bfloat16) if render else stack[0].view(B, n, N, D)
        we = w16(self.feat_embed_enc_to_dec.weight)
        # LayerNorm statistics of the residual stream (three scratch buffers rotate through the block)
        sa, sb, sc = fold_stats(rows, D, dev, 3, enable=(not render) or self.fold_ln_render)
        sa4 = None if sa is None else sa.view(B, n, N, -1, 2)
        first_untagged = mem is None and not render
        if first_untagged:
            for b in range(B):
                ops.gemm(xin[b, 0], we, bias=self._embed_bias(False), out=hx[b, 0], stats_out=None if sa is None else sa4[b, 0])
                if n > 1:
                    ops.gemm(xin[b, 1:], we, bias=self._embed_bias(True), out=hx[b, 1:],
                             stats_out=None if sa is None else sa4[b, 1:])
        else:
            ops.gemm(xin, we, bias=self._embed_bias(True), out=hx, stats_out=sa)
        cur = hx.view(rows, D)

        bank: Optional[MemoryBank] = None
        if mem is not None:
            mem_vals, mem_labels, mem_nimgs = mem[0], mem[1], mem[2]
            bank = getattr(mem_vals, "bank", None)
            n_mem = mem_vals[0].shape[1]
            if bank is not None and not render and bank.n != n_mem:
                # `mem` is a VALUE in the reference (torch.cat builds new tensors).  Our bank appends in place, so an
                # update that does not start from the bank's tail (branching from / re-running an earlier memory)
                # first moves to a private copy: the other holder's tokens and K|V stay untouched.
                bank = bank.fork(n_mem)
        else:
            if render:
                raise ops._l.Pst3rError("render=True needs a memory")
            mem_vals, mem_labels, mem_nimgs, n_mem = None, None, 0, 0

        keep_all = return_feats is True
        feats: List[torch.Tensor] = [hx] if keep_all else []

        def stored_kv(l):
            if bank is not None:
                return [bank.kv[l][b:b + 1, :n_mem] for b in range(B)]
            # foreign memory tuple (e.g. the reference's sliced render path): project on the fly
            blk = self.blocks_dec[l]
            mv = mem_vals[l]  # an activation: converted per call, never through the parameter cache
            mv = mv.contiguous() if mv.dtype == torch.bfloat16 else ops.to_bf16(mv.float().contiguous())
            kv = ops.gemm(mv, blk.kv_weight(), bias=blk.kv_bias())
            return [kv[b:b + 1] for b in range(kv.shape[0])]

        if render:
            for l, blk in enumerate(self.blocks_dec):
                in_place = not keep_all
                cur = self_attention(cur, blk, B * n, N, rope, in_place=in_place, stats=sa, stats_out=sb)
                cur = self._cross_attention(cur, blk, B, n, N, stored_kv(l), None, stats=sb, stats_out=sc)
                last = l == L - 1
                cur = mlp_residual(cur, blk.norm3, blk.mlp, 1e-6, out=feats_out if (last and feats_out is not None) else None,
                                   stats=sc, stats_out=sa)
                if keep_all:
                    feats.append(cur.view(B, n, N, D) if cur.is_contiguous() else cur.unflatten(0, (B, n, N)))
            new_mem = mem
        else:
            layer_in = []
            mask_bits = self._own_view_mask(n, N, n_mem, dev) if n > 1 else None
            # The memory update is a sequential chain of small (n * N rows) projections: split their reductions over two-CTA
            # clusters (csrc/gemm_splitk.cuh).  Only here: the chain's shapes are the same on every rank of a sharded run,
            # while the per-view stages stay on the unsplit kernels, whose results do not depend on how many views a rank holds.
            prev_split_k = ops.set_split_k(True)
            try:
                for l, blk in enumerate(self.blocks_dec):
                    layer_in.append(cur)
                    nxt = self_attention(cur, blk, B * n, N, rope, in_place=False, stats=sa, stats_out=sb)
                    if n == 1:
                        if n_mem == 0:
                            raise ops._l.Pst3rError("a single first view has nothing to attend to (init needs >= 2 views)")
                        kvs = stored_kv(l)
                    else:
                        fresh = ops.layernorm(cur, f32(blk.norm_y.weight), f32(blk.norm_y.bias), 1e-6)
                        kvf = ops.gemm(fresh, blk.kv_weight(), bias=blk.kv_bias()).view(B, n * N, 2 * D)
                        kvs = []
                        old = stored_kv(l) if n_mem > 0 else None
                        for b in range(B):
                            kvs.append(kvf[b:b + 1] if old is None else torch.cat([old[b], kvf[b:b + 1]], dim=1))
                    nxt = self._cross_attention(nxt, blk, B, n, N, kvs, mask_bits, stats=sb, stats_out=sc)
                    cur = mlp_residual(nxt, blk.norm3, blk.mlp, 1e-6,
                                       out=stack[l + 1] if l < L - 1 else (feats_out if feats_out is not None else None),
                                       stats=sc, stats_out=sa)
                    if keep_all:
                        feats.append(cur.view(B, n, N, D) if cur.is_contiguous() else cur.unflatten(0, (B, n, N)))
                # feedback + memory write: stored_l = norm_y_l(layer_in_l + feedback(x_L)); K|V projected once, here
                fb = ops.gemm(cur, w16(self.feedback_layer.fc1.weight), bias=bias_of(self.feedback_layer.fc1), act=ops.ACT_GELU)
                fb = ops.gemm(fb, w16(self.feedback_layer.fc2.weight), bias=bias_of(self.feedback_layer.fc2))
            finally:
                ops.set_split_k(prev_split_k)
            if bank is None:
                cap = max(self.reserve_views, mem_nimgs + n) * N
                bank = MemoryBank(B, L, D, dev, cap)
                if n_mem > 0:  # adopt a foreign memory
                    bank.reserve(n_mem + n * N)
                    for l, blk in enumerate(self.blocks_dec):
                        bank.tok[l][:, :n_mem].copy_(mem_vals[l])
                        for b in range(B):
                            ops.gemm(bank.tok[l][b, :n_mem], blk.kv_weight(), bias=blk.kv_bias(), out=bank.kv[l][b, :n_mem])
                    bank.n = n_mem
            bank.reserve(n_mem + n * N)
            if B == 1:  # all L layers in two launches
                dst = bank.tok_all[:, 0, n_mem:n_mem + n * N]
                ops.layernorm_batched(stack, self._stacked("ny_g"), self._stacked("ny_b"), 1e-6, add=fb, out=dst)
                ops.gemm_batched(dst, self._stacked("kv_w"), bias=self._stacked("kv_b"),
                                 out=bank.kv_all[:, 0, n_mem:n_mem + n * N])
            else:
                for l, blk in enumerate(self.blocks_dec):
                    for b in range(B):
                        dst = bank.tok[l][b, n_mem:n_mem + n * N]
                        sl = slice(b * n * N, (b + 1) * n * N)
                        ops.layernorm(layer_in[l][sl], f32(blk.norm_y.weight), f32(blk.norm_y.bias), 1e-6, add=fb[sl], out=dst)
                        ops.gemm(dst, blk.kv_weight(), bias=blk.kv_bias(), out=bank.kv[l][b, n_mem:n_mem + n * N])
            bank.n = n_mem + n * N
            vals = MemVals([bank.tok[l][:, :bank.n] for l in range(L)])
            vals.bank = bank
            new_labels = (mem_nimgs + torch.arange(n, device=dev)).view(1, n, 1).expand(B, n, N).reshape(B, n * N)
            labels = new_labels if mem_labels is None else torch.cat([mem_labels, new_labels], dim=1)
            new_mem = (vals, labels, mem_nimgs + n, None, None)

        pointmaps = None
        if compute_pointmaps:
            hn = ops.layernorm(cur, f32(self.norm_dec.weight), f32(self.norm_dec.bias), 1e-6)
            pointmaps = torch.empty((B, n, H, W, self.head_channels), device=dev, dtype=torch.float32)
            ops.gemm(hn, self._head_weight(), bias=self._head_bias(), out=pointmaps, store_mode=ops.STORE_D2S,
                     grid=(h_, w_), d2s=(self.patch_size, self.head_channels))
            if H > W:  # portrait: predicted in the true orientation, returned in the landscape storage convention
                pointmaps = pointmaps.transpose(2, 3)
        if return_feats == "last":
            feats = [cur.view(B, n, N, D) if cur.is_contiguous() else cur.unflatten(0, (B, n, N))]
        return new_mem, pointmaps, (feats if return_feats else None)
