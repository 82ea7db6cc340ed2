"""In-graph cost of the latency-bound kernels of the sequential memory build (development tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panst3r_b200 import ops

def timed_graph(fn, reps=50, iters=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / reps * 1e3  # us per call

r = lambda *s: torch.randn(*s, device="cuda").bfloat16()
x = r(768, 768); g_, b_ = torch.ones(768, device="cuda"), torch.zeros(768, device="cuda")
bias = torch.zeros(3072, device="cuda")
for (N, K, name) in [(2304, 768, "qkv"), (768, 768, "proj"), (3072, 768, "fc1"), (768, 3072, "fc2"), (1536, 768, "kv")]:
    a, w = r(768, K), r(N, K)
    out = torch.empty(768, N, device="cuda", dtype=torch.bfloat16)
    print(f"gemm {name} M768 N{N} K{K}: {timed_graph(lambda: ops.gemm(a, w, bias=bias[:N], out=out)):.2f} us/call", flush=True)
y = torch.empty_like(x)
print(f"layernorm 768x768: {timed_graph(lambda: ops.layernorm(x, g_, b_, 1e-6, out=y)):.2f} us/call")
qkv = r(1, 768, 3, 12, 64)
o = torch.empty(1, 768, 768, device="cuda", dtype=torch.bfloat16)
for sp in (0, 1):
    print(f"self-attn 768x768 h12 splits={sp}: {timed_graph(lambda: ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=o, kv_splits=sp)):.2f} us/call")
q = r(1, 768, 12, 64)
for nk in (1536, 6144, 11520):
    kv = r(1, nk, 1536)
    k, v = kv[:, :, :768].unflatten(-1, (12, 64)), kv[:, :, 768:].unflatten(-1, (12, 64))
    for sp in (0, 1, 2, 8):
        print(f"cross-attn Nk={nk} splits={sp}: {timed_graph(lambda: ops.attention(q, k, v, out=o, kv_splits=sp)):.2f} us/call")
# an empty-ish kernel for the floor
z = torch.zeros(8, 8, device="cuda", dtype=torch.bfloat16)
print(f"tiny add_bcast: {timed_graph(lambda: ops.add_bcast(z, z, out=z)):.2f} us/call")
