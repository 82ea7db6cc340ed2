"""CUPTI per-kernel device-time sums of one eager PanSt3R.forward (16 keyframes, 512x384, v1, single stream).
Tree-agnostic development tool: `python tools/kernel_sums.py [repo_root] [head_precision]` — used to compare two builds of
the library on the same box (e.g. this tree vs the round-1 tree)."""
import os
import re
import sys

root = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prec = sys.argv[2] if len(sys.argv) > 2 else None
sys.path.insert(0, root)
os.chdir(root)
os.environ.setdefault("PST3R_PDL", "0")
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
import bench  # noqa: E402
from panst3r_b200.panst3r import build_panst3r  # noqa: E402

with torch.device("cuda"):
    m = build_panst3r("v1", head_precision=prec) if prec else build_panst3r("v1")
bench.init_weights_(m)
g = torch.Generator().manual_seed(7)
m.panoptic_decoder.text_encoder.class_embeddings = {c: torch.randn(768, generator=g) for c in bench.CLASSES}
m.overlap_dino = False
m.panoptic_decoder.mask_transformer.overlap_aux_masks = False
imgs, ts = bench.make_inputs(16, "cuda")
imgs = imgs.cuda()
for _ in range(3):
    m(imgs, ts, bench.CLASSES)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m(imgs, ts, bench.CLASSES)
    torch.cuda.synchronize()
agg = {}
for ev in prof.key_averages():
    t = float(getattr(ev, "device_time_total", 0.0) or 0.0)
    if t <= 0:
        continue
    mm = re.search(r"pst3r::(\w+)(<[^(]*>)?", ev.key)
    name = (mm.group(1) + (mm.group(2) or "")) if mm else "other"
    a = agg.setdefault(name, [0.0, 0])
    a[0] += t / 1e3
    a[1] += ev.count
tot = sum(a[0] for a in agg.values())
print(f"root {root}  total {tot:.2f} ms")
for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"  {k:45s} {ms:8.3f} ms  {n:5d} calls  {1e3 * ms / n:8.2f} us/call")
