"""Development tool: scans input seeds for the conditioned head fixture (oracle/make_golden.py::COND) and reports the seed whose
reference run keeps every sign(mask logit) decision farthest from zero.  python tools/seed_search.py V H W n_seeds (needs /root/reference)."""
import sys, time, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import ref_import
from oracle.make_golden import build_ref_head, head_inputs, record_pooled_logits, _pooled, CLASSES
ref = ref_import.load_reference()
torch.set_num_threads(8)
V,H,W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
m = build_ref_head(ref, "v1", cls_logit_scale=3.0)
best=(0,None); t0=time.time()
for seed in range(100, 100+int(sys.argv[4])):
    feats, imgs, pos, ts = head_inputs(V,H,W, seed=seed)
    with torch.no_grad():
        with record_pooled_logits(m) as rec:
            out = m(feats, imgs, pos, ts, CLASSES)
    pl = _pooled(rec)
    mn = min((p.abs().min()/p.abs().max()).item() for p in pl)
    if mn > best[0]: best=(mn,seed); print(seed, f"{mn:.2e}", f"{time.time()-t0:.0f}s", flush=True)
print("best", best)
