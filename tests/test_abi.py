"""The C-ABI library loads and exports exactly what include/panst3r_b200.h declares (no compute calls: CPU box)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    __graft_entry__.build()
    from panst3r_b200 import lib
    return lib


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "panst3r_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pst3r_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound(built):
    lib = built.load()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in built.SIGNATURES, f"{s} has no ctypes signature"
    assert set(built.SIGNATURES) == set(syms)


def test_struct_layouts_match_header(built):
    # field order of the ctypes mirrors equals the C declaration order
    src = open(os.path.join(ROOT, "include", "panst3r_b200.h")).read()
    body = re.search(r"typedef struct pst3r_gemm_epilogue \{(.*?)\} pst3r_gemm_epilogue;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"([A-Za-z_0-9]+)\s*$", part.strip())[0])
    assert names == [f[0] for f in built.GemmEpilogue._fields_]
    body = re.search(r"typedef struct pst3r_attn_args \{(.*?)\} pst3r_attn_args;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            for part in decl.split(","):
                names.append(re.findall(r"([A-Za-z_0-9]+)\s*$", part.strip())[0])
    assert names == [f[0] for f in built.AttnArgs._fields_]
    assert ctypes.sizeof(built.GemmEpilogue) % 8 == 0 and ctypes.sizeof(built.AttnArgs) % 8 == 0


def test_version_and_error_string(built):
    lib = built.load()
    assert lib.pst3r_version() == built.ABI_VERSION == 3
    hdr = open(os.path.join(ROOT, "include", "panst3r_b200.h")).read()
    assert re.search(r"#define\s+PST3R_ABI_VERSION\s+3\b", hdr)
    assert isinstance(lib.pst3r_last_error(), bytes)


def test_host_side_switches_without_a_gpu(built):
    """pst3r_set_pdl / pst3r_set_split_k / pst3r_set_sm_budget are host-side launch settings: they work, and return the previous
    value, without a device (split-K is off by default: only the memory build switches it on around its own launches)."""
    lib = built.load()
    assert lib.pst3r_set_split_k(1) == 0 and lib.pst3r_set_split_k(0) == 1 and lib.pst3r_set_split_k(0) == 0
    prev = lib.pst3r_set_pdl(0)
    assert prev in (0, 1) and lib.pst3r_set_pdl(prev) == 0


def test_bad_arguments_are_rejected_without_a_gpu(built):
    lib = built.load()
    e = built.GemmEpilogue()
    rc = lib.pst3r_gemm_bf16(None, 8, None, 8, 128, 128, 64, ctypes.byref(e), None)
    assert rc == -1 and b"null" in lib.pst3r_last_error()
    a = built.AttnArgs()
    assert lib.pst3r_attention(ctypes.byref(a), None) == -1


def test_no_cpu_fallback(built):
    from panst3r_b200 import ops
    from panst3r_b200.lib import Pst3rError
    x = torch.zeros(128, 64, dtype=torch.bfloat16)
    with pytest.raises(Pst3rError):
        ops.gemm(x, x)
    with pytest.raises(Pst3rError):
        ops.layernorm(x, torch.ones(64), torch.zeros(64), 1e-6)
    from panst3r_b200.panst3r import build_panst3r
    m = build_panst3r("v1", 1, 1, 1)  # CPU module: forward must refuse
    imgs = torch.zeros(1, 2, 3, 32, 48)
    with pytest.raises(Pst3rError):
        m(imgs, torch.tensor([[[32, 48]] * 2]), ["a"])


def test_missing_library_fails_loudly(built, tmp_path):
    with pytest.raises(built.Pst3rError):
        built.load(str(tmp_path / "nope.so"))


def test_product_never_imports_oracle():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import panst3r_b200, panst3r_b200.ops, panst3r_b200.panst3r, "
            "panst3r_b200.dist; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "panst3r_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f"{f} imports the oracle"
                assert "/root/reference" not in txt, f"{f} reads the reference checkout"


def test_keyframe_selection_and_view_stacking():
    """Host logic of forward_inference_multi_ar (reference panst3r.py:181-206): linspace keyframes first, then the
    remaining views; views grouped into stacks of equal (true shape, stored shape) with ascending positions."""
    from panst3r_b200.panst3r import select_keyframes, stack_views
    kf, order, k = select_keyframes(6, 4)
    assert kf == [0, 1, 3, 5] and order == [0, 1, 3, 5, 2, 4] and k == 4
    assert select_keyframes(3, None) == ([0, 1, 2], [0, 1, 2], 3) and select_keyframes(3, 9)[2] == 3
    ts = [(64, 96), (64, 96), (96, 64), (64, 64), (64, 64), (64, 96)]
    stored = [(64, 96), (64, 96), (64, 96), (64, 64), (64, 64), (64, 96)]
    assert stack_views(ts, stored) == [[0, 1, 5], [2], [3, 4]]
    assert stack_views(ts[:2], stored[:2]) == [[0, 1]]
