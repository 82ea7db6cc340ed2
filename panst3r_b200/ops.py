"""Thin PyTorch-facing wrappers over the C ABI (include/panst3r_b200.h).

PyTorch is plumbing here: it owns device memory and the stream; every wrapper passes raw pointers,
strides and the current CUDA stream to libpanst3r_b200.so.  No wrapper computes anything in torch.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import lib as _l

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
STORE_PLAIN, STORE_TRANSPOSED, STORE_PIXSHUF2, STORE_D2S = 0, 1, 2, 3

# launch counter (bench.py reports gpu_launches from this) and algorithmic FLOPs issued per kernel class
launches = 0
flop_count = {}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _l.Pst3rError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise _l.Pst3rError(f"{name}: expected dtype {dtype}, got {t.dtype}")


def _rows2d(t: torch.Tensor, name: str) -> Tuple[int, int, int]:
    """(rows, cols, ld) of a tensor viewed as a row-strided 2-D matrix with contiguous columns."""
    if t.dim() < 2:
        raise _l.Pst3rError(f"{name}: need >=2 dims")
    if t.stride(-1) != 1:
        raise _l.Pst3rError(f"{name}: last dim must be contiguous")
    if t.dim() > 2:
        # leading dims must collapse onto a single row stride
        ld = t.stride(-2)
        exp = ld * t.shape[-2]
        for d in range(t.dim() - 3, -1, -1):
            if t.shape[d] != 1 and t.stride(d) != exp:
                raise _l.Pst3rError(f"{name}: leading dims are not collapsible (shape {tuple(t.shape)}, strides {t.stride()})")
            exp *= t.shape[d]
        rows = 1
        for d in t.shape[:-1]:
            rows *= d
        return rows, t.shape[-1], ld
    return t.shape[0], t.shape[1], t.stride(0)


KIND_BF16, KIND_F32, KIND_SPLIT = 0, 1, 2


class Split:
    """Reference-precision matrix for the panoptic head (the reference runs it in fp32, panst3r.py:236-245): every value
    x is held as two bf16 numbers, hi = bf16(x) and lo = bf16(x - hi) (16 mantissa bits), so that fp32-grade products
    run on the bf16 tensor cores as A_hi B_hi + A_lo B_hi + A_hi B_lo.  `hi` is a bf16 view [..., cols] (contiguous
    last dim); the lo parts live `lo_off` elements after their hi parts in the same storage (a fresh matrix is one
    [rows, 2*cols] buffer, rows = [hi | lo], lo_off = cols).  Row / column slices keep lo_off."""
    __slots__ = ("hi", "lo_off")

    def __init__(self, hi: torch.Tensor, lo_off: int):
        self.hi, self.lo_off = hi, int(lo_off)

    @staticmethod
    def empty(shape, device, align: int = 1) -> "Split":
        """align > 1: the lo parts start at the next multiple of `align` columns (a GEMM operand needs lo_off % 8 == 0;
        such a matrix is not `packed()`: only the GEMM and the kernels taking an explicit lo offset accept it)."""
        *lead, cols = shape
        pitch = -(-cols // align) * align
        buf = torch.empty((*lead, 2 * pitch), device=device, dtype=torch.bfloat16)
        return Split(buf[..., :cols], pitch)

    @staticmethod
    def from_float(x: torch.Tensor) -> "Split":
        """fp32 / bf16 CUDA tensor -> split copy (one conversion kernel)."""
        out = Split.empty(x.shape, x.device)
        convert(x, out)
        return out

    @property
    def lo(self) -> torch.Tensor:
        return self.hi.as_strided(self.hi.shape, self.hi.stride(), self.hi.storage_offset() + self.lo_off)

    @property
    def shape(self):
        return self.hi.shape

    @property
    def device(self):
        return self.hi.device

    @property
    def cols(self) -> int:
        return self.hi.shape[-1]

    def float(self) -> torch.Tensor:
        return self.hi.float() + self.lo.float()

    def view(self, *shape) -> "Split":
        """Reshape (the lo parts follow their hi parts at the same offset, so any view that keeps elements in place
        within their rows is valid; the last dim may be split into (heads, hd))."""
        return Split(self.hi.view(*shape), self.lo_off)

    def permute(self, *dims) -> "Split":
        return Split(self.hi.permute(*dims), self.lo_off)

    def __getitem__(self, idx) -> "Split":
        return Split(self.hi[idx], self.lo_off)

    def full(self) -> torch.Tensor:
        """The underlying bf16 [..., 2*cols] rows [hi | lo] of a packed matrix (for pure data movement)."""
        if not self.packed():
            raise _l.Pst3rError("Split.full: not a packed [hi | lo] matrix")
        return self.hi.as_strided((*self.hi.shape[:-1], 2 * self.cols), self.hi.stride())

    def packed(self) -> bool:
        """rows are exactly [hi(cols) | lo(cols)] (what the row kernels expect)"""
        return self.lo_off == self.cols


def _kind(t) -> int:
    if isinstance(t, Split):
        return KIND_SPLIT
    if t.dtype == torch.float32:
        return KIND_F32
    if t.dtype == torch.bfloat16:
        return KIND_BF16
    raise _l.Pst3rError(f"unsupported dtype {t.dtype}")


def _base(t) -> torch.Tensor:
    return t.hi if isinstance(t, Split) else t


def _need_cuda(t, name: str):
    if not _base(t).is_cuda:
        raise _l.Pst3rError(f"{name}: expected a CUDA tensor (no CPU fallback)")


def _packed(t, name: str):
    if isinstance(t, Split) and not t.packed():
        raise _l.Pst3rError(f"{name}: split rows must be [hi | lo] with lo_off == cols")


def _dkind(out_dtype) -> int:
    """KIND_* of an out_dtype argument (torch.bfloat16 / torch.float32 / "split")."""
    if out_dtype == "split":
        return KIND_SPLIT
    if out_dtype is torch.float32:
        return KIND_F32
    if out_dtype is torch.bfloat16:
        return KIND_BF16
    raise _l.Pst3rError(f"unsupported out_dtype {out_dtype}")


def _new(shape, kind: int, device):
    if kind == KIND_SPLIT:
        return Split.empty(shape, device)
    return torch.empty(shape, device=device, dtype=torch.float32 if kind == KIND_F32 else torch.bfloat16)


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         col_scale: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, alpha: float = 1.0,
         out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
         store_mode: int = STORE_PLAIN, rows_per_batch: int = 0, batch_stride: int = 0, ldt: int = 0,
         grid: Tuple[int, int] = (0, 0), d2s: Tuple[int, int] = (0, 0),
         rope: Optional[Tuple[torch.Tensor, torch.Tensor, int]] = None, res_mod_rows: int = 0,
         out_ld: int = 0, ln: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None,
         stats_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = epilogue(a @ w.T).  a: bf16 [..., K] (row-strided), w: bf16 [N, K].
    Reference-precision mode: w is a `Split` weight; a is a `Split` (three product terms) or a plain bf16 tensor (exactly
    representable inputs, two terms); `out` / `residual` may be `Split`s, out_dtype="split" allocates one.
    PLAIN store with rows_per_batch > 0: `out` is only a base pointer, rows are remapped (pass out_ld).
    ln = (stats fp32 [M, K/32, 2], colsum fp32 [N], eps): LayerNorm(a) folded into the GEMM (w = gamma-scaled weights,
    bias includes W beta); stats_out fp32 [M, N/32, 2]: per-chunk (sum, sum^2) of the stored bf16 rows for the next fold."""
    global launches
    lib = _l.load()
    e = _l.GemmEpilogue()
    if isinstance(w, Split):
        e.split_terms = 3 if isinstance(a, Split) else 2
        e.a_lo_off = a.lo_off if isinstance(a, Split) else 0
        e.b_lo_off = w.lo_off
    elif isinstance(a, Split):
        raise _l.Pst3rError("gemm: a split activation needs split weights")
    ab, wb = _base(a), _base(w)
    _need(ab, torch.bfloat16, "gemm.a")
    _need(wb, torch.bfloat16, "gemm.w")
    M, K, lda = _rows2d(ab, "gemm.a")
    N, K2, ldb = _rows2d(wb, "gemm.w")
    if K != K2:
        raise _l.Pst3rError(f"gemm: K mismatch {K} vs {K2}")
    if out is None and store_mode == STORE_PLAIN and rows_per_batch == 0:
        if out_dtype == "split":
            out = Split.empty((*ab.shape[:-1], N), ab.device)
        else:
            out = torch.empty((*ab.shape[:-1], N), device=ab.device, dtype=out_dtype)
    if out is None:
        raise _l.Pst3rError("gemm: out must be given for remapped / non-plain store modes")
    ob = _base(out)
    if store_mode == STORE_PLAIN and rows_per_batch > 0:
        if out_ld <= 0:
            raise _l.Pst3rError("gemm: remapped PLAIN store needs out and out_ld")
        e.ldo = out_ld
    elif store_mode == STORE_PLAIN:
        _, oc, ldo = _rows2d(ob, "gemm.out")
        if oc != N:
            raise _l.Pst3rError("gemm: out has wrong number of columns")
        e.ldo = ldo
    else:
        e.ldo = ob.stride(-2) if ob.dim() >= 2 else 0
    e.out = ob.data_ptr()
    e.out_kind = _kind(out)
    if isinstance(out, Split):
        e.out_lo_off = out.lo_off
    e.act = act
    if bias is not None:
        _need(bias, torch.float32, "gemm.bias")
    if col_scale is not None:
        _need(col_scale, torch.float32, "gemm.col_scale")
    e.bias = _ptr(bias)
    e.col_scale = _ptr(col_scale)
    if residual is not None:
        rb = _base(residual)
        _need_cuda(residual, "gemm.residual")
        _, _, ldr = _rows2d(rb, "gemm.residual")
        e.residual = rb.data_ptr()
        e.ldr = ldr
        e.res_mod_rows = res_mod_rows
        e.res_kind = _kind(residual)
        if isinstance(residual, Split):
            e.res_lo_off = residual.lo_off
    e.alpha = alpha
    e.store_mode = store_mode
    e.rows_per_batch, e.batch_stride, e.ldt = rows_per_batch, batch_stride, ldt
    e.grid_h, e.grid_w = grid
    e.d2s_patch, e.d2s_ch = d2s
    if ln is not None:
        st, colsum, eps = ln
        _need(st, torch.float32, "gemm.ln_stats")
        _need(colsum, torch.float32, "gemm.ln_colsum")
        if not st.is_contiguous() or st.numel() != M * ((K + 31) // 32) * 2 or colsum.numel() != N:
            raise _l.Pst3rError(f"gemm: ln stats must be contiguous [M, K/32, 2] (got {tuple(st.shape)} for M={M} K={K})")
        e.ln_stats, e.ln_slots, e.ln_colsum, e.ln_eps = st.data_ptr(), (K + 31) // 32, colsum.data_ptr(), eps
    if stats_out is not None:
        _need(stats_out, torch.float32, "gemm.stats_out")
        if not stats_out.is_contiguous() or stats_out.numel() != M * ((N + 31) // 32) * 2:
            raise _l.Pst3rError("gemm: stats_out must be contiguous [M, N/32, 2]")
        e.stats_out = stats_out.data_ptr()
    if rope is not None:
        cs, pos, rope_cols = rope
        _need(cs, torch.float32, "gemm.rope_cs")
        _need(pos, torch.int32, "gemm.rope_pos")
        e.rope_cs, e.rope_pos, e.rope_cols, e.rope_maxpos = cs.data_ptr(), pos.data_ptr(), rope_cols, cs.shape[0]
    _l.check(lib.pst3r_gemm_bf16(ab.data_ptr(), lda, wb.data_ptr(), ldb, M, N, K, C.byref(e), _stream()), "pst3r_gemm_bf16")
    launches += 1
    flop_count["gemm"] = flop_count.get("gemm", 0.0) + 2.0 * M * N * K * max(1, e.split_terms)
    return out


def gemm_batched(a, w, *, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, alpha: float = 1.0, out):
    """out[b] = act(alpha * a[b] @ w[b].T + bias[b]) for b < L in ONE launch.  a: bf16 [L, M, K], w: bf16 [L, N, K],
    bias: fp32 [L, N], out: bf16/fp32 [L, M, N]; any batch / row strides, contiguous last dim.  `Split` operands /
    output select the reference-precision mode (per-head QK^T and PV products of the query decoder's attention)."""
    global launches
    lib = _l.load()
    e = _l.GemmEpilogue()
    if isinstance(w, Split):
        e.split_terms = 3 if isinstance(a, Split) else 2
        e.a_lo_off = a.lo_off if isinstance(a, Split) else 0
        e.b_lo_off = w.lo_off
    elif isinstance(a, Split):
        raise _l.Pst3rError("gemm_batched: a split activation needs split weights")
    ab, wb, ob = _base(a), _base(w), _base(out)
    _need(ab, torch.bfloat16, "gemm_batched.a")
    _need(wb, torch.bfloat16, "gemm_batched.w")
    if ab.dim() != 3 or wb.dim() != 3 or ob.dim() != 3 or ab.stride(-1) != 1 or wb.stride(-1) != 1 or ob.stride(-1) != 1:
        raise _l.Pst3rError("gemm_batched: expected 3-D operands with contiguous last dim")
    L, M, K = ab.shape
    L2, N, K2 = wb.shape
    if L2 != L or K2 != K or tuple(ob.shape) != (L, M, N):
        raise _l.Pst3rError(f"gemm_batched: shape mismatch a{tuple(ab.shape)} w{tuple(wb.shape)} out{tuple(ob.shape)}")
    if ob.dtype not in (torch.float32, torch.bfloat16) or not ob.is_cuda:
        raise _l.Pst3rError("gemm_batched: out must be CUDA bf16 or fp32")
    e.out, e.ldo, e.out_kind, e.act, e.alpha, e.store_mode = ob.data_ptr(), ob.stride(1), _kind(out), act, alpha, STORE_PLAIN
    if isinstance(out, Split):
        e.out_lo_off = out.lo_off
    bias_bs = 0
    if bias is not None:
        _need(bias, torch.float32, "gemm_batched.bias")
        if tuple(bias.shape) != (L, N) or bias.stride(1) != 1:
            raise _l.Pst3rError("gemm_batched: bias must be [L, N]")
        e.bias, bias_bs = bias.data_ptr(), bias.stride(0)
    _l.check(lib.pst3r_gemm_bf16_batched(ab.data_ptr(), ab.stride(1), ab.stride(0), wb.data_ptr(), wb.stride(1), wb.stride(0), M, N, K,
                                         L, C.byref(e), ob.stride(0), bias_bs, _stream()), "pst3r_gemm_bf16_batched")
    launches += 1
    flop_count["gemm"] = flop_count.get("gemm", 0.0) + 2.0 * L * M * N * K * max(1, e.split_terms)
    return out


def layernorm_batched(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *,
                      add: Optional[torch.Tensor] = None, out: torch.Tensor) -> torch.Tensor:
    """out[b] = LN(x[b] + add) * gamma[b] + beta[b] in ONE launch.  x, out: bf16 [L, rows, D] (any batch / row strides);
    gamma, beta: fp32 [L, D]; add: bf16 [rows, D] shared by all b."""
    global launches
    lib = _l.load()
    _need(x, torch.bfloat16, "layernorm_batched.x")
    _need(out, torch.bfloat16, "layernorm_batched.out")
    _need(gamma, torch.float32, "layernorm_batched.gamma")
    _need(beta, torch.float32, "layernorm_batched.beta")
    if x.dim() != 3 or tuple(out.shape) != tuple(x.shape) or x.stride(-1) != 1 or out.stride(-1) != 1:
        raise _l.Pst3rError("layernorm_batched: expected [L, rows, D] in/out with contiguous last dim")
    L, rows, D = x.shape
    if tuple(gamma.shape) != (L, D) or tuple(beta.shape) != (L, D) or gamma.stride() != beta.stride() or gamma.stride(1) != 1:
        raise _l.Pst3rError("layernorm_batched: gamma/beta must be [L, D] with equal strides")
    ld_add = 0
    if add is not None:
        _need(add, torch.bfloat16, "layernorm_batched.add")
        r2, c2, ld_add = _rows2d(add, "layernorm_batched.add")
        if (r2, c2) != (rows, D):
            raise _l.Pst3rError("layernorm_batched: add must be [rows, D]")
    _l.check(lib.pst3r_layernorm_batched(x.data_ptr(), x.stride(1), x.stride(0), _ptr(add), ld_add, gamma.data_ptr(),
                                         beta.data_ptr(), gamma.stride(0), eps, out.data_ptr(), out.stride(1), out.stride(0),
                                         rows, L, D, _stream()), "pst3r_layernorm_batched")
    launches += 1
    return out


def set_sm_budget(n: int) -> int:
    """Limit the SMs subsequent launches may fill (0 = all); returns the previous budget."""
    return _l.load().pst3r_set_sm_budget(int(n))


def set_pdl(on: bool) -> bool:
    """Programmatic dependent launch on / off for the following launches; returns the previous setting."""
    return bool(_l.load().pst3r_set_pdl(int(bool(on))))


def set_split_k(on: bool) -> bool:
    """Split-K cluster kernel for small all-bf16 GEMMs on / off for the following launches; returns the previous setting."""
    return bool(_l.load().pst3r_set_split_k(int(bool(on))))


def num_sms() -> int:
    return _l.load().pst3r_num_sms()


class sm_budget:
    """with ops.sm_budget(n): launches inside size their persistent grids / wave heuristics for n SMs."""

    def __init__(self, n: int):
        self.n = n

    def __enter__(self):
        self.prev = set_sm_budget(self.n)
        return self

    def __exit__(self, *exc):
        set_sm_budget(0 if self.prev >= num_sms() else self.prev)
        return False


def new_stats(rows: int, cols: int, device) -> torch.Tensor:
    """Per-row partial LayerNorm statistics buffer for `gemm(..., stats_out=)`: fp32 [rows, cols/32, 2]."""
    return torch.empty((rows, (cols + 31) // 32, 2), device=device, dtype=torch.float32)


_ws_cache = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Scratch buffer (split-KV partials, GroupNorm statistics) of the CURRENT stream.  One buffer per (device,
    stream): DINOv2 runs on a side stream next to the encoder / memory build, and two kernels that both split their
    KV range must never share partial-result storage.  A buffer is allocated while its stream is current and only
    ever used on it, so the caching allocator's stream-ordered reuse keeps a replaced (grown) buffer alive until the
    kernels queued on it have run."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, scale: Optional[float] = None,
              mask_bits: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
              kv_splits: int = 0) -> torch.Tensor:
    """q: bf16 [B, Nq, H, hd]; k, v: bf16 [B or 1, Nk, H, hd] (any strides, hd contiguous).
    mask_bits: int32/uint32 [B or 1, Nq, W] words, bit set = key blocked.  Returns bf16 [B, Nq, H*hd]."""
    global launches
    lib = _l.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _need(t, torch.bfloat16, f"attention.{n}")
        if t.dim() != 4 or t.stride(-1) != 1:
            raise _l.Pst3rError(f"attention.{n}: expected [B, N, H, hd] with contiguous hd")
    B, Nq, H, hd = q.shape
    Bk, Nk, Hk, hdk = k.shape
    if (Hk, hdk) != (H, hd) or v.shape != k.shape or Bk not in (1, B):
        raise _l.Pst3rError("attention: shape mismatch")
    if out is None:
        out = torch.empty((B, Nq, H * hd), device=q.device, dtype=torch.bfloat16)
    a = _l.AttnArgs()
    a.q, a.q_sb, a.q_sn, a.q_sh = q.data_ptr(), q.stride(0), q.stride(1), q.stride(2)
    shared = (Bk == 1 and B > 1) or k.stride(0) == 0
    a.k, a.k_sb, a.k_sn, a.k_sh = k.data_ptr(), (0 if shared else k.stride(0)), k.stride(1), k.stride(2)
    a.v, a.v_sb, a.v_sn, a.v_sh = v.data_ptr(), (0 if shared else v.stride(0)), v.stride(1), v.stride(2)
    if B == 1:  # batch stride is irrelevant; keep it non-zero so that K/V are not flagged shared
        a.k_sb = a.k_sb or Nk * k.stride(1)
        a.v_sb = a.v_sb or Nk * v.stride(1)
    a.o, a.o_sb, a.o_sn = out.data_ptr(), out.stride(0), out.stride(1)
    a.B, a.H, a.Nq, a.Nk, a.head_dim = B, H, Nq, Nk, hd
    a.scale = float(scale if scale is not None else hd ** -0.5)
    if mask_bits is not None:
        if mask_bits.dtype not in (torch.int32, torch.uint32) or mask_bits.dim() != 3 or mask_bits.stride(-1) != 1:
            raise _l.Pst3rError("attention.mask_bits: expected int32 [B or 1, Nq, W]")
        a.mask_bits = mask_bits.data_ptr()
        a.mask_sb = 0 if mask_bits.shape[0] == 1 else mask_bits.stride(0)
        a.mask_sq = 0 if (mask_bits.shape[1] == 1 and Nq > 1) else mask_bits.stride(1)  # one row for all queries
    splits = kv_splits if kv_splits > 0 else lib.pst3r_attention_auto_splits(B, H, Nq, Nk)
    a.kv_splits = splits
    need = lib.pst3r_attention_workspace_bytes(B, H, Nq, hd, splits)
    if need > 0:
        ws = _workspace(need, q.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    _l.check(lib.pst3r_attention(C.byref(a), _stream()), "pst3r_attention")
    launches += 1 if splits <= 1 else 2
    flop_count["attention"] = flop_count.get("attention", 0.0) + 4.0 * B * H * Nq * Nk * hd
    return out


def layernorm(x, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *,
              add=None, out=None, out_dtype=torch.bfloat16, sum_out=None,
              x_rows: Optional[Tuple[int, int, int, int, int]] = None):
    """y = LN(x [+ add]) * gamma + beta.  x / add / out / sum_out: bf16, fp32 or `Split` (out_dtype="split" allocates one;
    a Split input defaults to a Split output).
    x_rows = (rows, dim, ldx, rows_per_batch, batch_stride): x is only a base pointer with remapped rows."""
    global launches
    lib = _l.load()
    xb = _base(x)
    _need_cuda(x, "layernorm.x")
    _packed(x, "layernorm.x")
    if xb.dtype not in (torch.bfloat16, torch.float32):
        raise _l.Pst3rError("layernorm.x: expected CUDA bf16/fp32")
    x_rpb, x_bs = 0, 0
    if x_rows is not None:
        rows, dim, ldx, x_rpb, x_bs = x_rows
        if out is None:
            raise _l.Pst3rError("layernorm: remapped input needs out")
    else:
        rows, dim, ldx = _rows2d(xb, "layernorm.x")
    _need(gamma, torch.float32, "layernorm.gamma")
    _need(beta, torch.float32, "layernorm.beta")
    if out is None:
        if isinstance(x, Split) and out_dtype is torch.bfloat16:
            out_dtype = "split"
        out = _new(xb.shape, _dkind(out_dtype), xb.device)
    _packed(out, "layernorm.out")
    _, _, ldy = _rows2d(_base(out), "layernorm.out")
    ld_add, add_kind = 0, 0
    if add is not None:
        _need_cuda(add, "layernorm.add")
        _packed(add, "layernorm.add")
        _, _, ld_add = _rows2d(_base(add), "layernorm.add")
        add_kind = _kind(add)
    ld_sum, sum_kind = 0, 0
    if sum_out is not None:
        _need_cuda(sum_out, "layernorm.sum_out")
        _packed(sum_out, "layernorm.sum_out")
        _, _, ld_sum = _rows2d(_base(sum_out), "layernorm.sum_out")
        sum_kind = _kind(sum_out)
    _l.check(lib.pst3r_layernorm(xb.data_ptr(), _kind(x), ldx, _ptr(None if add is None else _base(add)), add_kind, ld_add,
                                 gamma.data_ptr(), beta.data_ptr(), eps, _base(out).data_ptr(), _kind(out), ldy,
                                 _ptr(None if sum_out is None else _base(sum_out)), sum_kind, ld_sum, rows, dim, x_rpb, x_bs,
                                 _stream()), "pst3r_layernorm")
    launches += 1
    return out


def rope2d_(tokens: torch.Tensor, pos: torch.Tensor, base: float = 100.0, fwd: float = 1.0) -> torch.Tensor:
    """In-place 2-D RoPE on bf16 tokens [B, N, H, D] (D contiguous) with int32 positions [B, N, 2]."""
    global launches
    lib = _l.load()
    _need(tokens, torch.bfloat16, "rope2d.tokens")
    _need(pos, torch.int32, "rope2d.pos")
    B, N, H, D = tokens.shape
    if tokens.stride(-1) != 1 or not pos.is_contiguous():
        raise _l.Pst3rError("rope2d: bad strides")
    _l.check(lib.pst3r_rope2d(tokens.data_ptr(), tokens.stride(0), tokens.stride(1), tokens.stride(2), pos.data_ptr(),
                              B, N, H, D, base, fwd, _stream()), "pst3r_rope2d")
    launches += 1
    return tokens


def add_bcast(a, b, out=None):
    """out[r] = a[r] + b[r % b_rows] on row matrices of any kind (bf16 / fp32 / Split); out defaults to a's kind."""
    global launches
    lib = _l.load()
    for t, n in ((a, "a"), (b, "b")):
        _need_cuda(t, f"add_bcast.{n}")
        _packed(t, f"add_bcast.{n}")
    rows, cols, lda = _rows2d(_base(a), "add_bcast.a")
    brows, bcols, ldb = _rows2d(_base(b), "add_bcast.b")
    if bcols != cols:
        raise _l.Pst3rError("add_bcast: column mismatch")
    if out is None:
        out = _new(_base(a).shape, _kind(a), _base(a).device)
    _packed(out, "add_bcast.out")
    _, _, ldo = _rows2d(_base(out), "add_bcast.out")
    _l.check(lib.pst3r_add_bcast(_base(a).data_ptr(), _kind(a), lda, _base(b).data_ptr(), _kind(b), ldb, brows,
                                 _base(out).data_ptr(), _kind(out), ldo, rows, cols, _stream()), "pst3r_add_bcast")
    launches += 1
    return out


def convert(x, out):
    """out = x between bf16 / fp32 / Split row matrices of equal logical shape."""
    global launches
    lib = _l.load()
    _need_cuda(x, "convert.x")
    _need_cuda(out, "convert.out")
    _packed(x, "convert.x")
    _packed(out, "convert.out")
    xb, ob = _base(x), _base(out)
    if xb.dim() < 2:
        xb, ob = xb.reshape(1, -1), ob.reshape(1, -1)
    if not isinstance(x, Split) and xb.stride(-1) != 1:
        xb = xb.contiguous()
    rows, cols, ldx = _rows2d(xb, "convert.x")
    r2, c2, ldy = _rows2d(ob, "convert.out")
    if (rows, cols) != (r2, c2):
        raise _l.Pst3rError("convert: shape mismatch")
    _l.check(lib.pst3r_convert(xb.data_ptr(), _kind(x), ldx, ob.data_ptr(), _kind(out), ldy, rows, cols, _stream()), "pst3r_convert")
    launches += 1
    return out


def softmax_rows(S: torch.Tensor, Q: int, mask_bits: Optional[torch.Tensor] = None, out=None):
    """Row softmax over the last dim of S fp32 [..., Nk] (rows = (head, query), query = row % Q) with an optional block
    mask int32 [1, Q, W] (bit set = key blocked) -> probabilities as a packed `Split` [..., Nk] (default) or `out`."""
    global launches
    lib = _l.load()
    _need(S, torch.float32, "softmax_rows.S")
    rows, Nk, lds = _rows2d(S, "softmax_rows.S")
    if out is None:
        out = Split.empty(S.shape, S.device, align=8)  # feeds the P V GEMM: lo offset a multiple of 8
    _, _, ldo = _rows2d(_base(out), "softmax_rows.out")
    mptr, msq = None, 0
    if mask_bits is not None:
        if mask_bits.dtype not in (torch.int32, torch.uint32) or mask_bits.stride(-1) != 1:
            raise _l.Pst3rError("softmax_rows.mask_bits: expected int32 [1, Q, W]")
        mptr, msq = mask_bits.data_ptr(), mask_bits.stride(-2)
    _l.check(lib.pst3r_softmax_rows(S.data_ptr(), lds, rows, Nk, mptr, msq, Q, _base(out).data_ptr(), _kind(out), ldo,
                                    out.lo_off if isinstance(out, Split) else 0, _stream()), "pst3r_softmax_rows")
    launches += 1
    return out


def to_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    global launches
    lib = _l.load()
    _need(x, torch.float32, "to_bf16.x")
    rows, cols, ldx = _rows2d(x, "to_bf16.x")
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _, _, ldy = _rows2d(out, "to_bf16.out")
    _l.check(lib.pst3r_cast_f32_to_bf16(x.data_ptr(), ldx, out.data_ptr(), ldy, rows, cols, _stream()), "cast")
    launches += 1
    return out


def to_f32(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    global launches
    lib = _l.load()
    _need(x, torch.bfloat16, "to_f32.x")
    rows, cols, ldx = _rows2d(x, "to_f32.x")
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    _, _, ldy = _rows2d(out, "to_f32.out")
    _l.check(lib.pst3r_cast_bf16_to_f32(x.data_ptr(), ldx, out.data_ptr(), ldy, rows, cols, _stream()), "cast")
    launches += 1
    return out


def patchify(img: torch.Tensor, P: int, ld: Optional[int] = None) -> torch.Tensor:
    """img fp32 [B,3,H,W] -> bf16 [B*(H/P)*(W/P), ld] (Conv2d weight flattening order)."""
    global launches
    lib = _l.load()
    _need(img, torch.float32, "patchify.img")
    img = img.contiguous()
    B, _, H, W = img.shape
    ld = ld or 3 * P * P
    out = torch.empty((B * (H // P) * (W // P), ld), device=img.device, dtype=torch.bfloat16)
    _l.check(lib.pst3r_patchify(img.data_ptr(), B, H, W, P, out.data_ptr(), ld, _stream()), "pst3r_patchify")
    launches += 1
    return out


def dino_preprocess_patchify(img: torch.Tensor, Ho: int, Wo: int, P: int, ld: int) -> torch.Tensor:
    global launches
    lib = _l.load()
    _need(img, torch.float32, "dino_preprocess.img")
    img = img.contiguous()
    B, _, H, W = img.shape
    out = torch.empty((B * (Ho // P) * (Wo // P), ld), device=img.device, dtype=torch.bfloat16)
    _l.check(lib.pst3r_dino_preprocess_patchify(img.data_ptr(), B, H, W, Ho, Wo, P, out.data_ptr(), ld, _stream()),
             "pst3r_dino_preprocess_patchify")
    launches += 1
    return out


def center_pool8(feats):
    """feats bf16 [B, Hm, Wm, C] (or a packed `Split` of that shape) -> [B, Hm/8, Wm/8, C] of the same kind
    (mean of the centre 2x2 of each 8x8 cell)."""
    global launches
    lib = _l.load()
    fb = _base(feats)
    _need(fb, torch.bfloat16, "center_pool8.feats")
    _packed(feats, "center_pool8.feats")
    B, Hm, Wm, Cc = fb.shape
    if isinstance(feats, Split):
        if fb.stride() != (Hm * Wm * 2 * Cc, Wm * 2 * Cc, 2 * Cc, 1):
            raise _l.Pst3rError("center_pool8: split map must be dense [B, Hm, Wm, 2C]")
        out = Split.empty((B, Hm // 8, Wm // 8, Cc), fb.device)
    else:
        fb = fb.contiguous()
        out = torch.empty((B, Hm // 8, Wm // 8, Cc), device=fb.device, dtype=torch.bfloat16)
    _l.check(lib.pst3r_center_pool8(fb.data_ptr(), _kind(feats), B, Hm, Wm, Cc, _base(out).data_ptr(), _stream()), "pst3r_center_pool8")
    launches += 1
    return out


def attn_mask_bits(logits_t: torch.Tensor, Nk: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """logits_t fp32 [Q, >=Nk] -> int32 [1, Q, ceil(Nk/128)*4] block mask (bit set: logit < 0)."""
    global launches
    lib = _l.load()
    _need(logits_t, torch.float32, "attn_mask_bits.logits")
    Q, _, ld = _rows2d(logits_t, "attn_mask_bits.logits")
    words = ((Nk + 127) // 128) * 4
    if out is None:
        out = torch.empty((1, Q, words), device=logits_t.device, dtype=torch.int32)
    _l.check(lib.pst3r_attn_mask_bits(logits_t.data_ptr(), ld, Q, Nk, out.data_ptr(), _stream()), "pst3r_attn_mask_bits")
    launches += 1
    return out


def l2norm_rows(x: torch.Tensor, eps: float, out_dtype=torch.bfloat16):
    """y = x / (||x|| + eps) per row; fp32 in, bf16 / fp32 / "split" out."""
    global launches
    lib = _l.load()
    _need(x, torch.float32, "l2norm_rows.x")
    rows, cols, ldx = _rows2d(x, "l2norm_rows.x")
    out = _new(x.shape, _dkind(out_dtype), x.device)
    _, _, ldy = _rows2d(_base(out), "l2norm_rows.out")
    _l.check(lib.pst3r_l2norm_rows(x.data_ptr(), ldx, _base(out).data_ptr(), _kind(out), ldy, rows, cols,
                                   eps, _stream()), "pst3r_l2norm_rows")
    launches += 1
    return out


def nhwc_to_nchw_f32(x) -> torch.Tensor:
    """bf16 (or packed Split) [B, HW, C] -> fp32 [B, C, HW]"""
    global launches
    lib = _l.load()
    xb = _base(x)
    _need(xb, torch.bfloat16, "nhwc_to_nchw_f32.x")
    _packed(x, "nhwc_to_nchw_f32.x")
    B, HW, Cc = xb.shape
    if isinstance(x, Split):
        if xb.stride() != (HW * 2 * Cc, 2 * Cc, 1):
            raise _l.Pst3rError("nhwc_to_nchw_f32: split map must be dense [B, HW, 2C]")
    else:
        xb = xb.contiguous()
    out = torch.empty((B, Cc, HW), device=xb.device, dtype=torch.float32)
    _l.check(lib.pst3r_nhwc_to_nchw_f32(xb.data_ptr(), _kind(x), B, HW, Cc, out.data_ptr(), _stream()), "pst3r_nhwc_to_nchw_f32")
    launches += 1
    return out


def class_scores(logits: torch.Tensor):
    """fp32 [Q, K] class logits -> (scores fp32 [Q], labels int32 [Q]) = max / argmax of sigmoid(logits)."""
    global launches
    lib = _l.load()
    _need(logits, torch.float32, "class_scores.logits")
    Q, K, ld = _rows2d(logits, "class_scores.logits")
    scores = torch.empty(Q, device=logits.device, dtype=torch.float32)
    labels = torch.empty(Q, device=logits.device, dtype=torch.int32)
    _l.check(lib.pst3r_class_scores(logits.data_ptr(), ld, Q, K, scores.data_ptr(), labels.data_ptr(), _stream()), "pst3r_class_scores")
    launches += 1
    return scores, labels


def panoptic_argmax(masks: torch.Tensor, keep_idx: torch.Tensor, keep_scores: torch.Tensor, size: Tuple[int, int],
                    mask_threshold: float, area_half: torch.Tensor, area_won: torch.Tensor, *, out=None, band=None):
    """masks fp32 [V, Q, hm, wm] mask logits -> (ids int32 [V, H, W], win fp32 [V, H, W]); accumulates the per-query
    pixel counts into area_half / area_won (int32 [nkeep], zeroed by the caller).
    out = (ids, win): write into these maps instead of fresh ones.
    band = (y0, rows, src_row0, hm): only output rows [y0, y0 + rows); `masks` is then [V, Q, src_rows, wm], the source rows
    [src_row0, src_row0 + src_rows) of a mask grid of height hm (pst3r_panoptic_argmax_band)."""
    global launches
    lib = _l.load()
    _need(masks, torch.float32, "panoptic_argmax.masks")
    if masks.dim() != 4 or masks.stride(-1) != 1 or masks.stride(-2) != masks.shape[-1]:
        raise _l.Pst3rError("panoptic_argmax.masks: expected [V, Q, hm, wm] with dense planes")
    V, Q, src_rows, wm = masks.shape
    H, W = int(size[0]), int(size[1])
    y0, rows, src_row0, hm = (0, H, 0, src_rows) if band is None else (int(b) for b in band)
    nkeep = keep_idx.numel()
    if nkeep:
        _need(keep_idx, torch.int32, "panoptic_argmax.keep_idx")
        _need(keep_scores, torch.float32, "panoptic_argmax.keep_scores")
        _need(area_half, torch.int32, "panoptic_argmax.area_half")
        _need(area_won, torch.int32, "panoptic_argmax.area_won")
    if out is None:
        ids = torch.empty((V, H, W), device=masks.device, dtype=torch.int32)
        win = torch.empty((V, H, W), device=masks.device, dtype=torch.float32)
    else:
        ids, win = out
        _need(ids, torch.int32, "panoptic_argmax.ids")
        _need(win, torch.float32, "panoptic_argmax.win")
        if tuple(ids.shape) != (V, H, W) or tuple(win.shape) != (V, H, W) or not (ids.is_contiguous() and win.is_contiguous()):
            raise _l.Pst3rError(f"panoptic_argmax: out maps must be contiguous [{V}, {H}, {W}]")
    _l.check(lib.pst3r_panoptic_argmax_band(masks.data_ptr(), masks.stride(0), masks.stride(1), V, hm, wm, src_row0, src_rows,
                                            _ptr(keep_idx) if nkeep else None, _ptr(keep_scores) if nkeep else None, nkeep, H, W,
                                            y0, rows, float(mask_threshold), ids.data_ptr(), win.data_ptr(), H * W, W,
                                            _ptr(area_half) if nkeep else None, _ptr(area_won) if nkeep else None, _stream()),
             "pst3r_panoptic_argmax_band")
    launches += 1
    return ids, win


def panoptic_finalize(ids: torch.Tensor, win: torch.Tensor, lut: Optional[torch.Tensor], mask_threshold: float,
                      void_confidence: float):
    """(ids, win) winner maps + lut int32 [nkeep] (segment id or 0) -> (pan int32, conf fp32) of the same shape."""
    global launches
    lib = _l.load()
    _need(ids, torch.int32, "panoptic_finalize.ids")
    _need(win, torch.float32, "panoptic_finalize.win")
    if not (ids.is_contiguous() and win.is_contiguous()):
        raise _l.Pst3rError("panoptic_finalize: expected contiguous maps")
    nkeep = 0 if lut is None else lut.numel()
    if nkeep:
        _need(lut, torch.int32, "panoptic_finalize.lut")
    pan, conf = torch.empty_like(ids), torch.empty_like(win)
    _l.check(lib.pst3r_panoptic_finalize(ids.data_ptr(), win.data_ptr(), _ptr(lut) if nkeep else None, nkeep, float(mask_threshold),
                                         float(void_confidence), pan.data_ptr(), conf.data_ptr(), ids.numel(), _stream()),
             "pst3r_panoptic_finalize")
    launches += 1
    return pan, conf


def conv3x3_nhwc(x, w, cpad: int, *, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, out=None):
    """3x3 / stride 1 / zero-pad 1 convolution of a pixel-major bf16 map x [V, H, W, ld>=C] with tap-major weights
    w bf16 [O, 9*cpad] as an implicit tcgen05 GEMM.  Returns bf16 [V, H, W, O].  Reference-precision mode: x and w are
    `Split`s (pixel rows [hi | lo]; weight rows [hi(9 cpad) | lo(9 cpad)]), the result a packed `Split`."""
    global launches
    lib = _l.load()
    split = isinstance(w, Split)
    if split != isinstance(x, Split):
        raise _l.Pst3rError("conv3x3: x and w must both be Split or both plain bf16")
    xb, wb = _base(x), _base(w)
    _need(xb, torch.bfloat16, "conv3x3.x")
    _need(wb, torch.bfloat16, "conv3x3.w")
    V, H, W, Cc = xb.shape
    if xb.stride(-1) != 1 or xb.stride(2) != xb.stride(-2) or xb.stride(1) != W * xb.stride(2) or xb.stride(0) != H * xb.stride(1):
        raise _l.Pst3rError("conv3x3.x: expected a dense pixel-major map")
    O = wb.shape[0]
    if wb.shape[1] != 9 * cpad or wb.stride(0) != (2 if split else 1) * 9 * cpad or (split and w.lo_off != 9 * cpad):
        raise _l.Pst3rError("conv3x3.w: expected contiguous [O, 9*cpad] (split: rows [hi | lo])")
    if out is None:
        out = Split.empty((V, H, W, O), xb.device) if split else torch.empty((V, H, W, O), device=xb.device, dtype=torch.bfloat16)
    ob = _base(out)
    e = _l.GemmEpilogue()
    e.out, e.ldo, e.out_kind, e.act, e.alpha = ob.data_ptr(), ob.stride(2), _kind(out), act, 1.0
    if split:
        e.split_terms, e.a_lo_off, e.b_lo_off, e.out_lo_off = 3, x.lo_off, w.lo_off, out.lo_off
    if bias is not None:
        _need(bias, torch.float32, "conv3x3.bias")
        e.bias = bias.data_ptr()
    _l.check(lib.pst3r_conv3x3_nhwc(xb.data_ptr(), xb.stride(2), V, H, W, Cc, wb.data_ptr(), cpad, O, C.byref(e), _stream()),
             "pst3r_conv3x3_nhwc")
    launches += 1
    flop_count["gemm"] = flop_count.get("gemm", 0.0) + 2.0 * V * H * W * O * 9 * Cc * (3 if split else 1)
    return out


def _loftup_ws(V: int, Cc: int, groups: int, device) -> torch.Tensor:
    return _workspace(_l.load().pst3r_loftup_workspace_bytes(V, Cc, groups) + (1 << 16), device)


def loftup_guidance(img: torch.Tensor):
    """img fp32 [V,3,H,W] -> (half fp32 [V,3,H/2,W/2], minmax fp32 [3,2] over the whole batch)."""
    global launches
    lib = _l.load()
    _need(img, torch.float32, "loftup_guidance.img")
    img = img.contiguous()
    V, _, H, W = img.shape
    half = torch.empty((V, 3, H // 2, W // 2), device=img.device, dtype=torch.float32)
    minmax = torch.empty((3, 2), device=img.device, dtype=torch.float32)
    ws = _loftup_ws(V, 3, 1, img.device)
    _l.check(lib.pst3r_loftup_guidance(img.data_ptr(), V, H, W, half.data_ptr(), minmax.data_ptr(), ws.data_ptr(), _stream()),
             "pst3r_loftup_guidance")
    launches += 2
    return half, minmax


def loftup_fourier_gn(half, minmax, gy, gx, freqs, biases, gamma, beta, eps: float, ld: int, split: bool = False):
    """-> pixel-major [V, Hh, Wh, ld]: GroupNorm(1)(ImplicitFeaturizer(MinMaxScaler(half))); bf16, or a packed `Split`."""
    global launches
    lib = _l.load()
    for t, n in ((half, "half"), (minmax, "minmax"), (gy, "gy"), (gx, "gx"), (freqs, "freqs"), (biases, "biases"),
                 (gamma, "gamma"), (beta, "beta")):
        _need(t, torch.float32, f"loftup_fourier_gn.{n}")
    V, _, Hh, Wh = half.shape
    nf = freqs.numel()
    out = Split.empty((V, Hh, Wh, ld), half.device) if split else torch.empty((V, Hh, Wh, ld), device=half.device, dtype=torch.bfloat16)
    ws = _loftup_ws(V, 10 * nf + 3, 1, half.device)
    _l.check(lib.pst3r_loftup_fourier_gn(half.data_ptr(), minmax.data_ptr(), gy.data_ptr(), gx.data_ptr(), freqs.data_ptr(),
                                         biases.data_ptr(), V, Hh, Wh, nf, gamma.data_ptr(), beta.data_ptr(), eps,
                                         _base(out).data_ptr(), _kind(out), ld, ws.data_ptr(), _stream()), "pst3r_loftup_fourier_gn")
    launches += 3
    return out


def groupnorm_nhwc_(x, groups: int, gamma, beta, eps: float, relu: bool):
    """In-place GroupNorm(groups) (+ReLU) on a dense pixel-major map [V, ..., C]: bf16, or a packed `Split`."""
    global launches
    lib = _l.load()
    xb = _base(x)
    _need(xb, torch.bfloat16, "groupnorm.x")
    _packed(x, "groupnorm.x")
    V, Cc = xb.shape[0], xb.shape[-1]
    npix = xb.numel() // (V * Cc)
    if isinstance(x, Split):
        if not x.full().is_contiguous():
            raise _l.Pst3rError("groupnorm.x: expected a dense split map")
    elif not xb.is_contiguous():
        raise _l.Pst3rError("groupnorm.x: expected contiguous")
    ws = _loftup_ws(V, Cc, groups, xb.device)
    _l.check(lib.pst3r_groupnorm_nhwc(xb.data_ptr(), _kind(x), V, npix, Cc, groups, gamma.data_ptr(), beta.data_ptr(), eps, int(relu),
                                      ws.data_ptr(), _stream()), "pst3r_groupnorm_nhwc")
    launches += 3
    return x


# ---------------------------------------------------------------------------------------------------
# Per-kernel-kind profiler (CUDA events on the launching stream around every C-ABI call).  Used by bench.py
# to find the dominant kernel of a step; never active inside a timed region.
# ---------------------------------------------------------------------------------------------------
class _Profiler:
    def __init__(self):
        self.records = []

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for kind, e0, e1 in self.records:
            d = agg.setdefault(kind, {"ms": 0.0, "calls": 0})
            d["ms"] += e0.elapsed_time(e1)
            d["calls"] += 1
        total = sum(d["ms"] for d in agg.values()) or 1.0
        for d in agg.values():
            d["share"] = d["ms"] / total
            d["ms"] = round(d["ms"], 4)
        return dict(sorted(agg.items(), key=lambda kv: -kv[1]["ms"]))


_prof: Optional[_Profiler] = None


class profiler:
    def __enter__(self):
        global _prof
        _prof = _Profiler()
        return _prof

    def __exit__(self, *exc):
        global _prof
        _prof = None
        return False


def _wrap(fn, kind_of):
    import functools

    @functools.wraps(fn)
    def inner(*a, **k):
        if _prof is None:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        _prof.records.append((kind_of(a, k), e0, e1))
        return r
    return inner


def _gemm_kind(a, k):
    sm = k.get("store_mode", STORE_PLAIN)
    return "gemm" + {STORE_PLAIN: "", STORE_TRANSPOSED: "_transposed_store", STORE_PIXSHUF2: "_pixshuf_store", STORE_D2S: "_d2s_store"}[sm]


def _attn_kind(a, k):
    q, kk = a[0], a[1]
    return f"attention_hd{q.shape[-1]}" + ("_masked" if k.get("mask_bits") is not None else "") + \
        ("_mem" if kk.shape[1] > 2 * q.shape[1] else "")


gemm = _wrap(gemm, _gemm_kind)
gemm_batched = _wrap(gemm_batched, lambda a, k: "gemm_batched")
layernorm_batched = _wrap(layernorm_batched, lambda a, k: "layernorm_batched")
attention = _wrap(attention, _attn_kind)
for _n in ("layernorm", "rope2d_", "add_bcast", "convert", "softmax_rows", "to_bf16", "to_f32", "patchify", "dino_preprocess_patchify", "center_pool8",
           "attn_mask_bits", "l2norm_rows", "nhwc_to_nchw_f32", "conv3x3_nhwc", "loftup_guidance", "loftup_fourier_gn",
           "groupnorm_nhwc_"):
    globals()[_n] = _wrap(globals()[_n], (lambda name: (lambda a, k: name))(_n))
