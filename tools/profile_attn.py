import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panst3r_b200 import ops
r = lambda *s: torch.randn(*s, device="cuda").bfloat16()
V = 16
q, k, v = r(V, 768, 12, 64), r(1, V * 768, 12, 64), r(1, V * 768, 12, 64)
for _ in range(2):
    ops.attention(q, k, v)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ops.attention(q, k, v)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
