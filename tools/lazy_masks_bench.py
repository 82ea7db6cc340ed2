"""Lazy mask logits vs the materialised tensor at BASELINE config 2's size (16 views x 200 queries, 192 x 256 mask grid,
384 x 512 output): timing with CUDA events, and DRAM traffic under ncu (never a bench number).

  python tools/lazy_masks_bench.py time                      -> JSON on stdout
  ncu --profile-from-start off --cache-control none --clock-control none \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/lazy_ncu.csv \
      python tools/lazy_masks_bench.py ncu                   -> launch segments in gpurun_out/lazy_ncu_segments.json
  python tools/lazy_masks_bench.py parse gpurun_out/lazy_ncu.csv gpurun_out/lazy_ncu_segments.json   -> markdown table

`--cache-control none` matters: ncu's default flushes the caches before every kernel, which would hide exactly the
L2 residency this path is built on.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

V, HM, WM, C, Q, K = 16, 192, 256, 256, 200, 20
SIZE = (384, 512)
# (scratch MB per chunk, chunks in flight): 8 / 20 / 40 MB = 6 / 2 / 1 band(s) per view
CONFIGS = [(8, 1), (20, 1), (40, 1), (20, 2), (20, 3), (10, 3), (40, 2)]


def setup(split: bool):
    import torch
    from panst3r_b200 import ops, postprocess as pp
    g = torch.Generator(device="cuda").manual_seed(0)
    f = torch.randn(V, HM, WM, C, device="cuda", generator=g) * 0.5
    e = torch.randn(Q, C, device="cuda", generator=g) * 0.5
    lz = pp.LazyMasks(ops.Split.from_float(f), ops.Split.from_float(e)) if split else pp.LazyMasks(ops.to_bf16(f), ops.to_bf16(e))
    keep = torch.arange(0, Q, 2, device="cuda", dtype=torch.int32)  # 100 surviving queries
    sc = torch.rand(keep.numel(), device="cuda")
    areas = torch.zeros(2, keep.numel(), device="cuda", dtype=torch.int32)
    return lz, keep, sc, areas


def variants(lz, keep, sc, areas):
    """name -> (callable running ONE argmax round incl. the mask GEMM(s), number of kernel launches)"""
    from panst3r_b200 import ops
    out = {"materialised": (lambda: ops.panoptic_argmax(lz.materialize(), keep, sc, SIZE, 0.25, areas[0], areas[1]), 2)}
    for mb, ns in CONFIGS:
        n = 2 * V * len(lz.band_plan(SIZE[0], mb << 20))
        out[f"lazy_{mb}MB_x{ns}"] = (lambda mb=mb, ns=ns: lz.panoptic_argmax(keep, sc, SIZE, 0.25, areas[0], areas[1],
                                                                           scratch_bytes=mb << 20, streams=ns), n)
    return out


def flush():
    import torch
    torch.empty(512 << 20, device="cuda", dtype=torch.uint8).fill_(1)  # > 126 MB L2
    torch.cuda.synchronize()


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "time"
    if mode == "parse":
        rows = list(csv.reader(l for l in open(sys.argv[2]) if l.startswith('"')))
        hdr = rows[0]
        name_i, metric_i, val_i, id_i = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
        per = {}
        for r in rows[1:]:
            per.setdefault(int(r[id_i]), {})[r[metric_i]] = float(r[val_i].replace(",", ""))
            per[int(r[id_i])]["name"] = r[name_i]
        ids = sorted(per)
        segs = json.load(open(sys.argv[3]))
        print("| head precision | variant | launches | DRAM read MB | DRAM write MB | sum of kernel times us |")
        print("|---|---|---|---|---|---|")
        pos = 0
        for s in segs:
            chunk = [per[i] for i in ids[pos:pos + s["launches"]]]
            pos += s["launches"]
            rd = sum(c.get("dram__bytes_read.sum", 0) for c in chunk)
            wr = sum(c.get("dram__bytes_write.sum", 0) for c in chunk)
            t = sum(c.get("gpu__time_duration.sum", 0) for c in chunk)
            print(f"| {s['precision']} | {s['variant']} | {s['launches']} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {t / 1e3:.1f} |")
        assert pos == len(ids), f"{len(ids)} launches in the CSV, {pos} in the segment list"
        return

    import torch
    cudart = torch.cuda.cudart()
    if mode in ("ncu", "ncufull"):
        # ncufull: `ncu --set full` of four launches — the whole-map mask GEMM + argmax, then one band GEMM + band argmax
        segs = []
        for split in ((True,) if mode == "ncufull" else (True, False)):
            lz, keep, sc, areas = setup(split)
            for name, (fn, n) in variants(lz, keep, sc, areas).items():
                if mode == "ncufull" and name not in ("materialised", "lazy_20MB_x1"):
                    continue
                fn()
                fn()
                flush()
                cudart.cudaProfilerStart()
                fn()
                torch.cuda.synchronize()
                cudart.cudaProfilerStop()
                segs.append({"precision": "fp32-grade (split bf16)" if split else "bf16", "variant": name, "launches": n})
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(segs, open(os.path.join(ROOT, "gpurun_out", "lazy_ncu_segments.json"), "w"))
        return

    from panst3r_b200 import postprocess as pp
    res = {"shape": {"views": V, "queries": Q, "mask_grid": [HM, WM], "output": list(SIZE), "kept_queries": 100}, "rows": []}
    for split in (True, False):
        lz, keep, sc, areas = setup(split)
        for name, (fn, n) in variants(lz, keep, sc, areas).items():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            t0.record()
            for _ in range(reps):
                fn()
            t1.record()
            torch.cuda.synchronize()
            res["rows"].append({"precision": "split" if split else "bf16", "variant": name, "launches": n,
                                "ms_per_round": t0.elapsed_time(t1) / reps})
        # the reference-facing call (class scores, two rounds with their host-side filtering, finalize)
        cls = torch.randn(1, Q, K, device="cuda") * 2
        for name, arg in (("materialised", None), ("lazy", lz)):
            def call():
                m = lz.materialize()[None] if arg is None else lz[None]
                return pp.panoptic_inference_v2(cls, m, SIZE)
            call()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(5):
                call()
            t1.record()
            torch.cuda.synchronize()
            res["rows"].append({"precision": "split" if split else "bf16", "variant": f"panoptic_inference_v2, {name}",
                                "ms_per_call": t0.elapsed_time(t1) / 5})
        res.setdefault("peak_alloc_mb", {})["split" if split else "bf16"] = torch.cuda.max_memory_allocated() / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
