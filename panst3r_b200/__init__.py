"""panst3r_b200 — Blackwell-native (sm_100a) implementation of PanSt3R's multi-view inference hot path.

Host code is PyTorch plumbing (memory, streams, torch.distributed); all arithmetic runs in the hand-written
CUDA library behind the C ABI in include/panst3r_b200.h (panst3r_b200/lib/libpanst3r_b200.so).
"""
from .lib import Pst3rError, load  # noqa: F401

__all__ = ["Pst3rError", "load", "build_panst3r", "PanSt3R"]


def __getattr__(name):  # lazy: importing the package must not require CUDA
    if name in ("PanSt3R", "build_panst3r"):
        from . import panst3r as _p
        return getattr(_p, name)
    raise AttributeError(name)
