"""Times the hd-64 attention kernel at the in-step shapes and checks it against torch SDPA, including a peaky-score case
that forces the lazily kept reference maximum to be raised inside later tiles (development tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panst3r_b200 import ops

def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

r = lambda *s: torch.randn(*s, device="cuda").bfloat16()
for name, B, H, Nq, Nk, shared in [("render cross", 16, 12, 768, 12288, True), ("encoder self", 16, 16, 768, 768, False),
                                   ("dino self", 16, 16, 769, 769, False), ("membuild cross", 1, 12, 768, 11520, True),
                                   ("membuild self", 1, 12, 768, 768, False)]:
    q = r(B, Nq, H, 64)
    k, v = r(1 if shared else B, Nk, H, 64), r(1 if shared else B, Nk, H, 64)
    us = timed(lambda: ops.attention(q, k, v))
    fl = 4.0 * B * H * Nq * Nk * 64
    ref = torch.nn.functional.scaled_dot_product_attention(q[:1].transpose(1, 2).float(), k[:1].transpose(1, 2).float(), v[:1].transpose(1, 2).float())
    got = ops.attention(q, k, v)[:1].view(1, Nq, H, 64).transpose(1, 2).float()
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    print(f"{name:16s} B{B} H{H} Nq{Nq} Nk{Nk}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  rel err {err:.2e}", flush=True)
# peaky scores: the reference maximum must be raised inside later tiles
q = r(2, 512, 4, 64); k = r(2, 2048, 4, 64); v = r(2, 2048, 4, 64)
k[:, 1500:1510] *= 12.0
ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float())
got = ops.attention(q, k, v).view(2, 512, 4, 64).transpose(1, 2).float()
print("peaky rel err", ((got - ref).abs().max() / ref.abs().max()).item())
