#!/usr/bin/env python
"""bench.py — keyframe-views/sec of the PanSt3R forward hot path at 512x384 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference path on the host CPU cores

A step = one `PanSt3R.forward()` over one synthetic scene: 16 keyframes at 512x384 (W x H), v1 PixelShuffle head,
bf16 tensor-core math with fp32 accumulation, random-init weights of the reference architecture (BASELINE config 2).
N > 1: the SAME scene, views sharded across ranks (strong scaling; panst3r_b200/dist.py): per-view stages run on
the owning rank, one NCCL all-gather of encoder tokens precedes the (replicated) sequential memory build.
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "keyframe-views/sec @512x384 fwd"
UNIT = "views/s"
H_IMG, W_IMG = 384, 512
CLASSES = [f"class_{i}" for i in range(100)]


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(V, device, pinned=False):
    import torch
    g = torch.Generator().manual_seed(0)
    imgs = torch.rand(1, V, 3, H_IMG, W_IMG, generator=g) * 2 - 1
    if pinned:
        imgs = imgs.pin_memory()
    ts = torch.tensor([[[H_IMG, W_IMG]] * V])
    return imgs, ts


def init_weights_(model, seed=1):
    """Random-init weights of the reference architecture, scaled so activations stay O(1) through 24+12 layers."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            leaf = name.split(".")[-1]
            if p.dim() >= 2 and leaf in ("weight", "in_proj_weight"):
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g, device="cuda") * fan_in ** -0.5)
            elif leaf == "weight" and p.dim() == 1:
                p.fill_(1.0)
            elif leaf in ("bias", "in_proj_bias"):
                p.copy_(torch.randn(p.shape, generator=g, device="cuda") * 0.02)
            elif leaf == "lambda1":
                p.fill_(0.5)
            elif p.dim() == 0:
                p.fill_(0.0)
            else:
                p.copy_(torch.randn(p.shape, generator=g, device="cuda") * 0.02 if "embed" not in name or "position" in name
                        else torch.randn(p.shape, generator=g, device="cuda") * 0.5)


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU path (oracle port: reference panoptic-head code restated + restated MUSt3R; the upstream
    packages are not installable) on all host cores, on a bounded sample of the same workload."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.panst3r import build_panst3r as build_oracle
    from oracle import weights as OW
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Vs = args.ref_views
    torch.manual_seed(1)
    model = build_oracle(args.variant)
    model.panoptic_decoder.text_encoder.class_embeddings = OW.synth_class_embeddings(CLASSES)
    imgs, ts = make_inputs(Vs, "cpu")
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            model(imgs, ts, CLASSES)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = Vs * len(times) / total
    sample = (f"{Vs} keyframes of the 512x384 scene per step (full-depth fp32 forward incl. memory build, render, "
              f"v1 head); per-view cost grows with scene size, so this flatters the CPU at V=16")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{Vs}-keyframe sample of the 16-keyframe 512x384 batch, {args.variant} head, CPU fp32",
                   "views_per_step": Vs, "variant": args.variant},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    wd = int(os.environ.get("PST3R_BENCH_WATCHDOG", "0"))
    if wd > 0:  # debugging aid: dump every thread's stack and exit if the run has not finished after `wd` seconds
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from panst3r_b200 import ops
    from panst3r_b200.lib import load
    from panst3r_b200.panst3r import build_panst3r
    load()  # fails loudly if the CUDA extension is missing
    V = args.views
    with torch.device("cuda"):
        model = build_panst3r(args.variant)
    init_weights_(model)
    g = torch.Generator().manual_seed(7)
    model.panoptic_decoder.text_encoder.class_embeddings = {c: torch.randn(768, generator=g) for c in CLASSES}
    if world > 1:
        from panst3r_b200.dist import ShardedPanSt3R
        runner = ShardedPanSt3R(model, rank, world)
    else:
        runner = model
    host_imgs, ts = make_inputs(V, dev, pinned=True)
    dev_imgs = host_imgs.to(dev)
    keyframes = args.keyframes if 0 < args.keyframes < V else 0
    if keyframes and world > 1:
        raise SystemExit("--keyframes (BASELINE config 3) is a single-GPU workload")

    def step_device():
        if keyframes:  # BASELINE config 3: K keyframes build the memory, the other frames are render-only
            pms, pan = model.forward_inference_multi_ar(list(dev_imgs[0]), ts[0], CLASSES, num_keyframes=keyframes)
            return pan, pms
        return runner(dev_imgs, ts, CLASSES)

    def flat(o):  # the tensors a caller reads back: class logits, per-view mask logits, pointmaps
        pan, pm = o
        masks = pan["pred_masks"]
        return [pan["pred_logits"]] + (list(masks) if isinstance(masks, (list, tuple)) else [masks]) + \
            (list(pm) if isinstance(pm, (list, tuple)) else [pm])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also builds every prepared-weight / constant cache) ----
    for _ in range(max(args.warmup, 3)):
        out = step_device()
    barrier()
    launches0 = ops.launches
    out = step_device()
    launches_per_step = ops.launches - launches0
    barrier()

    # ---- optional CUDA graph over the whole forward ----
    graph = None
    if args.graph:  # NCCL all-gathers are graph-capturable; falls back to eager launches if capture fails
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                gout = step_device()
            graph.replay()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] CUDA graph capture failed ({e!r}); timing eager launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def run_step():
        if graph is not None:
            graph.replay()
            return gout
        return step_device()

    for _ in range(2):
        run_step()
    barrier()

    # ---- timed region: `value` (inputs resident in HBM) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = run_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    units = keyframes or V  # the metric counts keyframe views
    value = units * args.steps / (ms_total / 1e3)

    # ---- e2e: host buffers in, host buffers out, through the public forward() ----
    res_dev = flat(out)
    res_host = [torch.empty(r.shape, dtype=r.dtype).pin_memory() for r in res_dev]
    h2d = host_imgs.numel() * host_imgs.element_size()
    d2h = sum(r.numel() * r.element_size() for r in res_dev)

    # Serving-loop pipelining: step i's H2D (copy-in stream) and step i-1's D2H (copy-out stream) overlap the
    # forward of step i.  Every step still moves its own inputs from pinned host memory and lands its own
    # pred_logits / pred_masks / pointmaps in pinned host memory inside the timed region.
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    in_stage = [torch.empty_like(dev_imgs) for _ in range(2)]
    out_stage = [[torch.empty_like(r) for r in res_dev] for _ in range(2)]
    res_host2 = [res_host, [torch.empty(r.shape, dtype=r.dtype).pin_memory() for r in res_dev]]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_staged = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        for i in range(n):
            b_ = i & 1
            with torch.cuda.stream(s_in):           # host -> device of this step's images
                in_stage[b_].copy_(host_imgs, non_blocking=True)
                ev_in[b_].record(s_in)
            main.wait_event(ev_in[b_])
            dev_imgs.copy_(in_stage[b_], non_blocking=True)
            o = run_step()
            main.wait_event(ev_out[b_])             # staging buffer b_ must have been drained (step i-2)
            for st, r in zip(out_stage[b_], flat(o)):
                st.copy_(r, non_blocking=True)
            ev_staged[b_].record(main)
            with torch.cuda.stream(s_out):          # device -> host of this step's results
                s_out.wait_event(ev_staged[b_])
                for hbuf, st in zip(res_host2[b_], out_stage[b_]):
                    hbuf.copy_(st, non_blocking=True)
                ev_out[b_].record(s_out)
        main.wait_stream(s_out)

    e2e_loop(2)
    barrier()
    w0 = time.perf_counter()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    wall = time.perf_counter() - w0
    t = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = units * args.steps / (float(t.item()) / 1e3)

    # ---- per-kernel-kind profile of one eager step + roofline of the dominant kernel ----
    roofline, breakdown = None, None
    with ops.profiler() as prof:  # every rank runs the step (it contains collectives); rank 0 reports
        step_device()
    barrier()
    att_roof = None
    if rank == 0:
        peaks, peak_src = _peaks()
        breakdown = prof.summary()
        roofline = dominant_roofline(ops, torch, V, peaks, peak_src, breakdown)
        att_roof = attention_roofline(ops, torch, V, peaks, peak_src)
    barrier()

    # ---- N > 1 only: the same GPUs running one independent scene each (no collective), for comparison ----
    scene_parallel = None
    if world > 1 and args.scene_parallel:
        try:
            for _ in range(2):
                model(dev_imgs, ts, CLASSES)
            barrier()
            e0.record()
            for _ in range(args.steps):
                model(dev_imgs, ts, CLASSES)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            scene_parallel = {"value": world * V * args.steps / (float(t.item()) / 1e3), "unit": UNIT, "scaling": "weak",
                              "note": "one independent 16-keyframe scene per GPU, eager launches, no data-path collective"}
        except Exception as e:  # noqa: BLE001
            scene_parallel = {"error": repr(e)}

    if rank == 0:
        cpu_base = None
        if world == 1 and args.cpu_baseline:
            cpu_base = cpu_baseline_leg(args)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": (f"{V}-keyframe 512x384 batch" if not keyframes else
                                    f"{keyframes} keyframes + {V - keyframes} render-only frames at 512x384 (memory-query path)") +
                                   f", {'v1 PixelShuffle' if args.variant == 'v1' else 'v2 InputMixer + LoftUp'} head, bf16, " +
                                   ("PanSt3R.forward()" if not keyframes else "PanSt3R.forward_inference_multi_ar()"),
                       "views": V, "keyframes": keyframes or V, "variant": args.variant, "classes": len(CLASSES),
                       "parallelism": f"views sharded x{world}",
                       "cuda_graph": graph is not None,
                       "l2": "per-step working set (1.7 GB bf16 weights + >2 GB activations) exceeds the 126 MB L2; no flush needed"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "wall_ms_per_step": 1e3 * wall / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline,
            "attention_roofline": att_roof,
            "total_views_per_s": V * args.steps / (ms_total / 1e3),
            "scene_parallel": scene_parallel,
            "kernel_breakdown_ms": breakdown,
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down: drop the captured graph (it references the NCCL communicator) and leave through os._exit once every
        # rank is here.  destroy_process_group() was observed to block forever after NCCL collectives had been captured
        # into a CUDA graph (profiles/r01_multi_gpu.md); process exit releases the communicator just as well.
        del graph
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def dominant_roofline(ops, torch, V, peaks, peak_src, breakdown):
    """Time the dominant kernel (by share of the step) alone with CUDA events at its in-step shape."""
    dom = max(breakdown, key=lambda k: breakdown[k]["ms"]) if breakdown else "attention"
    dev = "cuda"
    reps = 10
    if dom.startswith("attention"):
        B, Hh, Nq, Nk, hd = V, 12, 768, V * 768, 64  # render cross-attention over the keyframe memory
        q = torch.randn(B, Nq, Hh, hd, device=dev).bfloat16()
        k = torch.randn(1, Nk, Hh, hd, device=dev).bfloat16()
        v = torch.randn(1, Nk, Hh, hd, device=dev).bfloat16()
        fn = lambda: ops.attention(q, k, v)  # noqa: E731
        flops = 4.0 * B * Hh * Nq * Nk * hd
        name = f"attention_fwd_kernel<64> B{B} H{Hh} Nq{Nq} Nk{Nk} (decoder render cross-attention)"
    else:
        M, N, K = V * 768, 4096, 1024  # encoder / DINOv2 fc1
        a = torch.randn(M, K, device=dev).bfloat16()
        w = torch.randn(N, K, device=dev).bfloat16()
        o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fn = lambda: ops.gemm(a, w, out=o)  # noqa: E731
        flops = 2.0 * M * N * K
        name = f"gemm_bf16_tn_kernel M{M} N{N} K{K} (ViT-L fc1)"
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ach = flops / (ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops", 1590.0))
    # DRAM bytes per launch of this kernel at this shape from the committed `ncu --set full` capture
    # (profiles/r01_ncu_full_key_kernels_v2.md): dram__bytes_read.sum + dram__bytes_write.sum
    traffic = {"gemm": 34.19e6 + 48.33e6, "attention": 57.09e6 + 7.11e6}.get("attention" if dom.startswith("attention") else "gemm")
    if V != 16:
        traffic = None
    return {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": traffic, "peak_source": peak_src + ", burst figure (kernel timed alone)", "kernel_ms": ms,
            "share_of_step": breakdown[dom]["share"] if breakdown and dom in breakdown else None}


def attention_roofline(ops, torch, V, peaks, peak_src):
    """The north star's 'achieved fraction of the attention-GEMM roofline': QK^T + PV FLOPs of the render
    cross-attention (the largest attention call of the step) / its CUDA-event time / the bf16 tensor peak."""
    B, Hh, Nq, Nk, hd = V, 12, 768, V * 768, 64
    q = torch.randn(B, Nq, Hh, hd, device="cuda").bfloat16()
    k = torch.randn(1, Nk, Hh, hd, device="cuda").bfloat16()
    v = torch.randn(1, Nk, Hh, hd, device="cuda").bfloat16()
    for _ in range(3):
        ops.attention(q, k, v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.attention(q, k, v)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ach = 4.0 * B * Hh * Nq * Nk * hd / (ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops", 1590.0))
    return {"kernel": f"attention3_fwd_kernel B{B} H{Hh} Nq{Nq} Nk{Nk} hd{hd} (render cross-attention over keyframe memory)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "kernel_ms": ms,
            "bound": "MUFU ex2 (16/clk/SM) at head_dim 64: ceiling ~1150 TFLOP/s; ncu: XU pipe 59 %, tensor pipe 29 %",
            "peak_source": peak_src}


def cpu_baseline_leg(args):
    """Oracle port timed on the host cores on a bounded sample (reported baseline, not the target)."""
    import torch
    from oracle.panst3r import build_panst3r as build_oracle
    from oracle import weights as OW
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Vs = args.ref_views
    torch.manual_seed(1)
    model = build_oracle(args.variant)
    model.panoptic_decoder.text_encoder.class_embeddings = OW.synth_class_embeddings(CLASSES)
    imgs, ts = make_inputs(Vs, "cpu")
    with torch.no_grad():
        model(imgs, ts, CLASSES)  # warm-up
        t0 = time.perf_counter()
        n = 0
        while n < 2 or (time.perf_counter() - t0 < 12.0 and n < 4):
            model(imgs, ts, CLASSES)
            n += 1
        dt = (time.perf_counter() - t0) / n
    return {"value": Vs / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{Vs}-keyframe 512x384 scene, full-depth fp32 oracle forward, {n} timed runs after 1 warm-up"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--views", type=int, default=16)
    ap.add_argument("--variant", default="v1", choices=["v1", "v2"])
    ap.add_argument("--ref-views", type=int, default=2)
    ap.add_argument("--keyframes", type=int, default=0,
                    help="BASELINE config 3: this many keyframes, the remaining --views frames are render-only (e.g. --views 64 --keyframes 8)")
    ap.add_argument("--no-scene-parallel", dest="scene_parallel", action="store_false")
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
