"""ctypes loader for libpanst3r_b200.so — the C-ABI boundary (include/panst3r_b200.h).

There is no CPU fallback: if the library is missing or the device is not sm_100 the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libpanst3r_b200.so")


ABI_VERSION = 3  # PST3R_ABI_VERSION of include/panst3r_b200.h these ctypes declarations were written against


class Pst3rError(RuntimeError):
    pass


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("out", C.c_void_p),
        ("ldo", C.c_int64),
        ("out_kind", C.c_int32),
        ("act", C.c_int32),
        ("bias", C.c_void_p),
        ("col_scale", C.c_void_p),
        ("residual", C.c_void_p),
        ("ldr", C.c_int64),
        ("res_mod_rows", C.c_int32),
        ("alpha", C.c_float),
        ("store_mode", C.c_int32),
        ("rows_per_batch", C.c_int64),
        ("batch_stride", C.c_int64),
        ("ldt", C.c_int64),
        ("grid_h", C.c_int32),
        ("grid_w", C.c_int32),
        ("d2s_patch", C.c_int32),
        ("d2s_ch", C.c_int32),
        ("rope_cs", C.c_void_p),
        ("rope_pos", C.c_void_p),
        ("rope_cols", C.c_int32),
        ("rope_maxpos", C.c_int32),
        ("ln_stats", C.c_void_p),
        ("ln_slots", C.c_int32),
        ("ln_colsum", C.c_void_p),
        ("ln_eps", C.c_float),
        ("stats_out", C.c_void_p),
        ("split_terms", C.c_int32),
        ("a_lo_off", C.c_int64),
        ("b_lo_off", C.c_int64),
        ("out_lo_off", C.c_int64),
        ("res_kind", C.c_int32),
        ("res_lo_off", C.c_int64),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_sb", C.c_int64), ("q_sn", C.c_int64), ("q_sh", C.c_int64),
        ("k", C.c_void_p), ("k_sb", C.c_int64), ("k_sn", C.c_int64), ("k_sh", C.c_int64),
        ("v", C.c_void_p), ("v_sb", C.c_int64), ("v_sn", C.c_int64), ("v_sh", C.c_int64),
        ("o", C.c_void_p), ("o_sb", C.c_int64), ("o_sn", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Nq", C.c_int32), ("Nk", C.c_int32), ("head_dim", C.c_int32),
        ("scale", C.c_float),
        ("mask_bits", C.c_void_p), ("mask_sb", C.c_int64), ("mask_sq", C.c_int64),
        ("kv_splits", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


# name -> (restype, argtypes); every symbol declared in include/panst3r_b200.h
_i32, _i64, _f, _p = C.c_int32, C.c_int64, C.c_float, C.c_void_p
SIGNATURES = {
    "pst3r_last_error": (C.c_char_p, []),
    "pst3r_version": (C.c_int, []),
    "pst3r_check_device": (C.c_int, []),
    "pst3r_num_sms": (C.c_int, []),
    "pst3r_set_sm_budget": (C.c_int, [_i32]),
    "pst3r_set_pdl": (C.c_int, [_i32]),
    "pst3r_set_split_k": (C.c_int, [_i32]),
    "pst3r_gemm_bf16_batched": (C.c_int, [_p, _i64, _i64, _p, _i64, _i64, _i32, _i32, _i32, _i32, C.POINTER(GemmEpilogue),
                                          _i64, _i64, _p]),
    "pst3r_layernorm_batched": (C.c_int, [_p, _i64, _i64, _p, _i64, _p, _p, _i64, _f, _p, _i64, _i64, _i32, _i32, _i32, _p]),
    "pst3r_gemm_bf16": (C.c_int, [_p, _i64, _p, _i64, _i32, _i32, _i32, C.POINTER(GemmEpilogue), _p]),
    "pst3r_conv3x3_nhwc": (C.c_int, [_p, _i64, _i32, _i32, _i32, _i32, _p, _i32, _i32, C.POINTER(GemmEpilogue), _p]),
    "pst3r_loftup_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "pst3r_loftup_guidance": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _p, _p]),
    "pst3r_loftup_fourier_gn": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _f, _p, _i32, _i64, _p, _p]),
    "pst3r_groupnorm_nhwc": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _p, _p, _f, _i32, _p, _p]),
    "pst3r_attention_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32, _i32]),
    "pst3r_attention_auto_splits": (_i32, [_i32, _i32, _i32, _i32]),
    "pst3r_attention": (C.c_int, [C.POINTER(AttnArgs), _p]),
    "pst3r_layernorm": (C.c_int, [_p, _i32, _i64, _p, _i32, _i64, _p, _p, _f, _p, _i32, _i64, _p, _i32, _i64, _i32, _i32, _i32,
                                  _i64, _p]),
    "pst3r_rope2d": (C.c_int, [_p, _i64, _i64, _i64, _p, _i32, _i32, _i32, _i32, _f, _f, _p]),
    "pst3r_add_bcast": (C.c_int, [_p, _i32, _i64, _p, _i32, _i64, _i32, _p, _i32, _i64, _i32, _i32, _p]),
    "pst3r_convert": (C.c_int, [_p, _i32, _i64, _p, _i32, _i64, _i32, _i32, _p]),
    "pst3r_softmax_rows": (C.c_int, [_p, _i64, _i32, _i32, _p, _i64, _i32, _p, _i32, _i64, _i64, _p]),
    "pst3r_cast_f32_to_bf16": (C.c_int, [_p, _i64, _p, _i64, _i32, _i32, _p]),
    "pst3r_cast_bf16_to_f32": (C.c_int, [_p, _i64, _p, _i64, _i32, _i32, _p]),
    "pst3r_patchify": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _i64, _p]),
    "pst3r_dino_preprocess_patchify": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i64, _p]),
    "pst3r_center_pool8": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _p, _p]),
    "pst3r_attn_mask_bits": (C.c_int, [_p, _i64, _i32, _i32, _p, _p]),
    "pst3r_l2norm_rows": (C.c_int, [_p, _i64, _p, _i32, _i64, _i32, _i32, _f, _p]),
    "pst3r_nhwc_to_nchw_f32": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "pst3r_class_scores": (C.c_int, [_p, _i64, _i32, _i32, _p, _p, _p]),
    "pst3r_panoptic_argmax": (C.c_int, [_p, _i64, _i64, _i32, _i32, _i32, _p, _p, _i32, _i32, _i32, _f, _p, _p, _i64, _i32,
                                        _p, _p, _p]),
    "pst3r_panoptic_argmax_band": (C.c_int, [_p, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _p, _p, _i32, _i32, _i32, _i32, _i32,
                                             _f, _p, _p, _i64, _i32, _p, _p, _p]),
    "pst3r_panoptic_finalize": (C.c_int, [_p, _p, _p, _i32, _f, _f, _p, _p, _i64, _p]),
}

_lib = None


def load(path: str | None = None) -> C.CDLL:
    """Load the shared library and bind every declared symbol; raises Pst3rError if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise Pst3rError(
            f"{p} not found: the CUDA extension is not built (run `python -m panst3r_b200.build` or "
            f"__graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.pst3r_version() != ABI_VERSION:  # a stale .so with other struct layouts would misbehave silently
        raise Pst3rError(f"{p} has ABI version {lib.pst3r_version()}, this package expects {ABI_VERSION}: rebuild it "
                         f"(python -m panst3r_b200.build)")
    if path is None:
        _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().pst3r_last_error()
        raise Pst3rError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
