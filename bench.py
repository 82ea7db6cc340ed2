#!/usr/bin/env python
"""bench.py — keyframe-views/sec of the PanSt3R forward hot path at 512x384 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference path on the host CPU cores

A step = one `PanSt3R.forward()` over one synthetic scene: 16 keyframes at 512x384 (W x H), v1 PixelShuffle head,
random-init weights of the reference architecture (BASELINE config 2), the reference's precision policy: bf16 tensor-core
math with fp32 accumulation for DINOv2 / encoder / decoder, fp32-grade (split-bf16) arithmetic for the panoptic head
(`--head-precision bf16` times the all-bf16 variant; the default run reports it next to the headline as `bf16_head`).
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
        self.index, self.rows, self.proc, self.t_mark = index, [], None, None

    def mark(self):
        """The timed region starts now (the sampler is started a little earlier, during the last warm-up replays of the same
        step, because nvidia-smi needs a few hundred ms to deliver its first line)."""
        self.t_mark = time.monotonic()

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
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

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
        timed = [r for t, r in self.rows if self.t_mark is None or t >= self.t_mark]
        window = "timed region"
        if not timed:  # region shorter than the sampling latency: the warm-up replays of the same step just before it
            timed, window = [r for _, r in self.rows], "last warm-up replays + timed region"
        for r in timed:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


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
def _timed(torch, dist, world, dev, fn, n):
    """n calls of fn between barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _capture(torch, fn):
    """Whole step in one CUDA graph (NCCL all-gathers and the side stream are capturable); None if capture fails."""
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        g.replay()
        torch.cuda.synchronize()
        return g, out
    except Exception as e:  # noqa: BLE001
        print(f"[bench] CUDA graph capture failed ({e!r}); timing eager launches", file=sys.stderr)
        torch.cuda.synchronize()
        return None, None


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
        model = build_panst3r(args.variant, head_precision=args.head_precision)
    init_weights_(model)
    g = torch.Generator().manual_seed(7)
    model.panoptic_decoder.text_encoder.class_embeddings = {c: torch.randn(768, generator=g) for c in CLASSES}
    if world > 1:
        from panst3r_b200.dist import ShardedPanSt3R, partition_views
        runner = ShardedPanSt3R(model, rank, world)
        v0, v1 = partition_views(V, world)[rank]
    else:
        runner = model
        v0, v1 = 0, V
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
    ops.flop_count.clear()
    out = step_device()
    launches_per_step = ops.launches - launches0
    flops_step = dict(ops.flop_count)
    barrier()

    graph, gout = _capture(torch, step_device) if args.graph else (None, None)

    def run_step():
        if graph is not None:
            graph.replay()
            return gout
        return step_device()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(2 if world == 1 else 12):  # same step, untimed (a fixed count: the sharded step holds collectives);
        run_step()                             # at N > 1 the steps are short and nvidia-smi needs ~0.3 s for its first line
    torch.cuda.synchronize()
    barrier()

    # ---- timed region: `value` (inputs resident in HBM) ----
    sampler.mark()
    ms_total = _timed(torch, dist, world, dev, run_step, args.steps)
    out = gout if graph is not None else out
    clocks = sampler.stop() if rank == 0 else None
    units = keyframes or V  # the metric counts keyframe views
    value = units * args.steps / (ms_total / 1e3)

    # ---- e2e: host buffers in, host buffers out, through the public forward() ----
    # Every rank moves only ITS views: the image slice it owns goes host -> device, its pointmaps / mask logits (and the
    # replicated class logits) come back.  Serving-loop pipelining: step i's H2D (copy-in stream) and step i-1's D2H
    # (copy-out stream) overlap the forward of step i; every step still lands its own results in pinned host memory
    # inside the timed region.
    res_dev = flat(out)
    h2d = host_imgs[:, v0:v1].numel() * host_imgs.element_size()
    d2h = sum(r.numel() * r.element_size() for r in res_dev)
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    in_stage = [torch.empty_like(dev_imgs[:, v0:v1]) for _ in range(2)]
    out_stage = [[torch.empty_like(r) for r in res_dev] for _ in range(2)]
    res_host2 = [[torch.empty(r.shape, dtype=r.dtype).pin_memory() for r in res_dev] for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_staged = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        for i in range(n):
            b_ = i & 1
            with torch.cuda.stream(s_in):           # host -> device of this step's (local) images
                in_stage[b_].copy_(host_imgs[:, v0:v1], non_blocking=True)
                ev_in[b_].record(s_in)
            main.wait_event(ev_in[b_])
            dev_imgs[:, v0:v1].copy_(in_stage[b_], non_blocking=True)
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
    ms_e2e = _timed(torch, dist, world, dev, lambda: e2e_loop(args.steps), 1)
    wall = time.perf_counter() - w0
    e2e_value = units * args.steps / (ms_e2e / 1e3)
    if world > 1:  # whole-job byte counts
        tb = torch.tensor([float(h2d), float(d2h)], device=dev)
        dist.all_reduce(tb)
        h2d_total, d2h_total = int(tb[0].item()), int(tb[1].item())
    else:
        h2d_total, d2h_total = h2d, d2h

    # ---- per-kernel breakdown of ONE timed-configuration step (CUPTI device durations via torch.profiler) ----
    # taken on ONE stream (DINOv2 not overlapped with the memory build): a kernel that shares the SMs with another stream's
    # kernels reports an inflated duration, and the sum would no longer compare with the step time
    # ... and without programmatic dependent launch: a dependent kernel that starts early waits for its predecessor
    # INSIDE its own measured duration
    mt_ = model.panoptic_decoder.mask_transformer
    ov, ova, pdl = model.overlap_dino, mt_.overlap_aux_masks, ops.set_pdl(False)
    model.overlap_dino = mt_.overlap_aux_masks = False
    try:
        breakdown = kernel_breakdown(torch, step_device, barrier)
    finally:
        model.overlap_dino, mt_.overlap_aux_masks = ov, ova
        ops.set_pdl(pdl)
    roofline = att_roof = gemm_class = None
    if rank == 0:
        peaks, peak_src = _peaks()
        roofline = dominant_roofline(ops, torch, V, peaks, peak_src, breakdown)
        att_roof = attention_roofline(ops, torch, V, peaks, peak_src)
        gemm_class = class_fractions(breakdown, flops_step, peaks)
    barrier()

    # ---- N = 1: the all-bf16 head variant, and the reference's own GPU path as denominator ----
    bf16_head = gpu_ref = None
    if world == 1 and not keyframes and args.head_precision == "fp32" and args.bf16_head:
        model.panoptic_decoder.precision = "bf16"
        for _ in range(3):
            step_device()
        g2, _ = _capture(torch, step_device) if args.graph else (None, None)
        fn2 = (lambda: g2.replay()) if g2 is not None else step_device
        for _ in range(2):
            fn2()
        ms2 = _timed(torch, dist, world, dev, fn2, args.steps)
        bf16_head = {"value": V * args.steps / (ms2 / 1e3), "unit": UNIT, "ms_per_step": ms2 / args.steps,
                     "note": "same step with PanopticDecoder.precision='bf16' (plain bf16 operands in the head, ~1e-2 parity)"}
        model.panoptic_decoder.precision = "fp32"
        del g2
    ids_only = None
    if world == 1 and not keyframes and args.ids_only:
        ids_only = ids_only_leg(args, torch, model, step_device, host_imgs, dev_imgs, ts, V, dev)
    if world == 1 and not keyframes and args.gpu_reference:
        gpu_ref = gpu_reference_leg(args, model, dev_imgs, ts, value)

    # ---- N > 1: BASELINE config 5 shape (8 keyframes per GPU, weak scaling) and scene-parallel replicas ----
    weak = scene_parallel = None
    if world > 1 and args.weak:
        weak = weak_scaling_leg(args, torch, dist, world, rank, local, dev, runner)
    if world > 1 and args.scene_parallel:
        try:
            for _ in range(2):
                model(dev_imgs, ts, CLASSES)
            ms3 = _timed(torch, dist, world, dev, lambda: model(dev_imgs, ts, CLASSES), args.steps)
            scene_parallel = {"value": world * V * args.steps / (ms3 / 1e3), "unit": UNIT, "scaling": "weak",
                              "note": "one independent 16-keyframe scene per GPU, eager launches, no data-path collective"}
        except Exception as e:  # noqa: BLE001
            scene_parallel = {"error": repr(e)}

    if rank == 0:
        cpu_base = None
        if world == 1 and args.cpu_baseline:
            cpu_base = cpu_baseline_leg(args)
        head_desc = "fp32-grade (split bf16) head" if args.head_precision == "fp32" else "bf16 head"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": (f"{V}-keyframe 512x384 batch" if not keyframes else
                                    f"{keyframes} keyframes + {V - keyframes} render-only frames at 512x384 (memory-query path)") +
                                   f", {'v1 PixelShuffle' if args.variant == 'v1' else 'v2 InputMixer + LoftUp'} head, bf16 trunk + " +
                                   head_desc + ", " +
                                   ("PanSt3R.forward()" if not keyframes else "PanSt3R.forward_inference_multi_ar()"),
                       "views": V, "keyframes": keyframes or V, "variant": args.variant, "classes": len(CLASSES),
                       "head_precision": args.head_precision,
                       "parallelism": f"views sharded x{world}",
                       "cuda_graph": graph is not None,
                       "l2": "per-step working set (1.7 GB bf16 weights + >2 GB activations) exceeds the 126 MB L2; no flush needed"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h_total,
                    "h2d_bytes_per_step_rank0": h2d, "wall_ms_per_step": 1e3 * wall / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline,
            "attention_roofline": att_roof,
            "class_fractions": gemm_class,
            "total_views_per_s": V * args.steps / (ms_total / 1e3),
            "bf16_head": bf16_head,
            "ids_only": ids_only,
            "gpu_reference": gpu_ref,
            "baseline_config5": weak,
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


def kernel_breakdown(torch, run_step, barrier):
    """Device time per kernel of ONE step (the timed step's launches, issued on a single stream) from CUPTI kernel
    records (torch.profiler): every kernel's own duration.  `_total_ms` compares with `ms_per_step` (the timed loop
    additionally overlaps DINOv2 with the memory build and replays a CUDA graph).  Never inside a timed region."""
    import re
    try:
        from torch.profiler import ProfilerActivity, profile
        run_step()
        barrier()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            run_step()
            torch.cuda.synchronize()
        barrier()
        agg = {}
        for ev in prof.key_averages():
            t_us = float(getattr(ev, "device_time_total", 0.0) or getattr(ev, "cuda_time_total", 0.0))
            if t_us <= 0:
                continue
            name = ev.key
            m = re.search(r"pst3r::(\w+)(<[^(]*>)?", name)
            if m:
                name = m.group(1) + (m.group(2) or "")
            elif "nccl" in name.lower():
                name = "nccl:" + name.split("(")[0][:48]
            elif name.lower().startswith(("memcpy", "memset")):
                name = name.split(" ")[0]
            else:
                name = "other:" + name.split("(")[0][:48]
            d = agg.setdefault(name, {"ms": 0.0, "calls": 0})
            d["ms"] += t_us / 1e3
            d["calls"] += int(ev.count)
        total = sum(d["ms"] for d in agg.values()) or 1.0
        for d in agg.values():
            d["share"] = round(d["ms"] / total, 4)
            d["ms"] = round(d["ms"], 4)
        out = dict(sorted(agg.items(), key=lambda kv: -kv[1]["ms"]))
        out["_total_ms"] = round(total, 3)
        out["_source"] = "CUPTI kernel durations (torch.profiler) of one step: single stream, eager launches, no PDL"
        return out
    except Exception as e:  # noqa: BLE001
        barrier()
        return {"_error": repr(e)}


def _load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the profiled kernels, parsed from the committed
    `ncu --set full` captures (tools/ncu_traffic.py writes profiles/r02_ncu_traffic.json from the raw CSV)."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def dominant_roofline(ops, torch, V, peaks, peak_src, breakdown):
    """The kernel with the largest share of the step (CUPTI breakdown), timed alone with CUDA events at its in-step shape
    AND with its in-step epilogue (bias + erf-GELU for the ViT-L fc1)."""
    names = [k for k in (breakdown or {}) if not k.startswith("_")]
    dom = names[0] if names else "attention5_fwd_kernel"
    dev = "cuda"
    reps = 20
    traffic_tab = _load_traffic()
    if dom.startswith("attention"):
        B, Hh, Nq, Nk, hd = V, 12, 768, V * 768, 64  # render cross-attention over the keyframe memory
        q = torch.randn(B, Nq, Hh, hd, device=dev).bfloat16()
        k = torch.randn(1, Nk, Hh, hd, device=dev).bfloat16()
        v = torch.randn(1, Nk, Hh, hd, device=dev).bfloat16()
        fn = lambda: ops.attention(q, k, v)  # noqa: E731
        flops = 4.0 * B * Hh * Nq * Nk * hd
        name = f"attention5_fwd_kernel B{B} H{Hh} Nq{Nq} Nk{Nk} hd64 (decoder render cross-attention)"
        tkey = "attention3_render"
    else:
        M, N, K = V * 768, 4096, 1024  # encoder / DINOv2 fc1 with its in-step epilogue
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
        bias = torch.randn(N, device=dev)
        o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fn = lambda: ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, out=o)  # noqa: E731
        flops = 2.0 * M * N * K
        name = f"gemm2_bf16_tn_kernel M{M} N{N} K{K} + bias + erf-GELU epilogue (ViT-L fc1, as in the step)"
        tkey = "gemm2_fc1_gelu"
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
    traffic = traffic_tab.get(tkey, {}).get("dram_bytes") if V == 16 else None
    return {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": traffic, "traffic_source": traffic_tab.get(tkey, {}).get("source"),
            "peak_source": peak_src + ", burst figure (kernel timed alone)", "kernel_ms": ms,
            "share_of_step": breakdown[dom]["share"] if breakdown and dom in breakdown else None,
            "dominant_by": "CUPTI share of the timed step"}


def class_fractions(breakdown, flops_step, peaks):
    """In-step fraction of the SUSTAINED bf16 peak per kernel class: algorithmic FLOPs the wrappers issued in one step
    (2MNK per GEMM incl. the split-precision terms, 4 Nq Nk hd per attention head) / CUPTI time of the class."""
    if not breakdown or "_error" in breakdown:
        return None
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    cls = {"gemm": ("gemm", "gemm2"), "attention": ("attention",)}
    out = {}
    for c, prefixes in cls.items():
        ms = sum(v["ms"] for k, v in breakdown.items() if not k.startswith("_") and k.startswith(prefixes) and "combine" not in k)
        fl = float(flops_step.get(c, 0.0))
        if ms > 0 and fl > 0:
            out[c] = {"tflop_per_step": fl / 1e12, "ms": round(ms, 3), "tflops": fl / (ms * 1e-3) / 1e12,
                      "frac_of_sustained_peak": fl / (ms * 1e-3) / 1e12 / peak}
    return out


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
    return {"kernel": f"attention5_fwd_kernel B{B} H{Hh} Nq{Nq} Nk{Nk} hd{hd} (render cross-attention over keyframe memory)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "kernel_ms": ms,
            "bound": "MUFU ex2 (16/clk/SM) at head_dim 64: ceiling ~1150 TFLOP/s",
            "peak_source": peak_src}


def ids_only_leg(args, torch, model, step_device, host_imgs, dev_imgs, ts, V, dev):
    """The same scene for a caller that wants panoptic ids, not mask logits (tools/demo_panst3r.py:232-242 runs
    `panoptic_inference_v2` on `pred_masks` right after the forward): `PanopticDecoder.lazy_masks` + the band-wise
    post-processing (panst3r_b200/postprocess.py, LazyMasks).  The forward (without the mask einsums) is one CUDA graph,
    the post-processing runs eagerly (its filtering rounds read two small counter arrays on the host).  Host images in,
    segment ids / confidences / pointmaps / class logits out, every step; next to it the same post-processing on the
    materialised `pred_masks` of the default step."""
    from panst3r_b200 import postprocess as pp
    pd = model.panoptic_decoder
    size = tuple(int(v) for v in ts[0][0])
    try:
        def timed(fn, n):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        n = max(3, min(args.steps, 10))
        host = {}
        # random-init class logits sit below the reference's 0.1 score threshold: nothing would survive and the
        # post-processing would have nothing to do.  Threshold = the median class score of this scene: half of the 200
        # queries go through the argmax rounds (the count is reported).
        pan0, _ = step_device()
        thr = float(pan0["pred_logits"][0].sigmoid().amax(-1).median())
        kept = int((pan0["pred_logits"][0].sigmoid().amax(-1) > thr).sum())
        del pan0

        def to_host(name, t):
            if name not in host:
                host[name] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            host[name].copy_(t, non_blocking=True)

        def make_step(lazy):
            pd.lazy_masks = lazy
            for _ in range(2):
                step_device()
            g, out = _capture(torch, step_device) if args.graph else (None, None)

            def step():
                dev_imgs.copy_(host_imgs, non_blocking=True)
                if g is not None:
                    g.replay()
                    pan, pm = out
                else:
                    pan, pm = step_device()
                res = pp.panoptic_inference_v2(pan["pred_logits"], pan["pred_masks"], size, cls_threshold=thr)[0]
                to_host("pan", res["pan"]); to_host("conf", res["conf"]); to_host("pm", pm); to_host("cls", pan["pred_logits"])
                return res
            return step

        lazy_step = make_step(True)
        ms_lazy = timed(lazy_step, n)
        res = lazy_step()
        torch.cuda.synchronize()
        d2h = sum(h.numel() * h.element_size() for h in host.values())
        mat_step = make_step(False)
        ms_mat = timed(mat_step, n)
        return {"value": V / (ms_lazy / 1e3), "unit": UNIT, "ms_per_step": ms_lazy, "segments": len(res["segments_info"]),
                "queries_above_class_threshold": kept,
                "h2d_bytes_per_step": host_imgs.numel() * host_imgs.element_size(), "d2h_bytes_per_step": d2h,
                "materialised": {"value": V / (ms_mat / 1e3), "ms_per_step": ms_mat,
                                 "note": "default forward (all seven mask einsums materialised) + the same post-processing"},
                "what": "PanSt3R forward with lazy masks (CUDA graph) + panoptic_inference_v2 on the GPU, host images in, "
                        "ids / confidences / pointmaps / class logits to pinned host memory; no (V, Q, H/2, W/2) tensor"}
    except Exception as e:  # noqa: BLE001  (a secondary leg must never take the headline line down)
        torch.cuda.synchronize()
        return {"error": repr(e)}
    finally:
        pd.lazy_masks = False


def gpu_reference_leg(args, model, dev_imgs, ts, our_value):
    """The reference's OWN single-GPU path on this box, as the denominator of the north star's '>= 8x the reference's
    single-GPU forward': the oracle modules (the reference's panoptic-head code restated + restated MUSt3R, same weights
    as the CUDA model) on `cuda`, torch's fused SDPA standing in for xformers (not installable here;
    gradio_panst3r.py:25), TF32 matmuls as the demo enables them (gradio_panst3r.py:19), bf16 autocast on DINOv2 /
    encoder / decoder and an fp32 head (panst3r.py:174, 204, 236-245).  CUDA events, 2 warm-up + 3 timed runs."""
    import torch
    from oracle import blocks as OB
    from oracle.panst3r import build_panst3r as build_oracle
    try:
        V = dev_imgs.shape[1]
        with torch.device("cuda"):
            ref = build_oracle(args.variant)
        ref.load_state_dict(model.state_dict(), strict=True)  # identical state-dict surface
        ref.panoptic_decoder.text_encoder.class_embeddings = {k: v.cuda() for k, v in
                                                              model.panoptic_decoder.text_encoder.class_embeddings.items()}
        OB.toggle_memory_efficient_attention(True)
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            with torch.no_grad():
                for _ in range(2):
                    ref(dev_imgs, ts, CLASSES, amp=True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 3
                e0.record()
                for _ in range(n):
                    ref(dev_imgs, ts, CLASSES, amp=True)
                e1.record()
                torch.cuda.synchronize()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32
            OB.toggle_memory_efficient_attention(False)
        ms = e0.elapsed_time(e1) / n
        val = V / (ms / 1e3)
        del ref
        torch.cuda.empty_cache()
        return {"value": val, "unit": UNIT, "ms_per_step": ms, "ours_over_reference": our_value / val,
                "what": "oracle modules on cuda: torch SDPA (xformers stand-in), TF32 matmul, bf16 autocast trunk + fp32 head, "
                        "eager PyTorch, same weights and inputs as our arm; 3 timed runs after 2 warm-ups",
                "kind": "port (upstream must3r/croco are not installable: oracle restatement)"}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def weak_scaling_leg(args, torch, dist, world, rank, local, dev, runner):
    """BASELINE config 5's shape at this N: 8 keyframes per GPU (64 keyframes on 8 GPUs), views sharded, NCCL token
    all-gather; the sequential memory build over all 8N keyframes is replicated."""
    Vw = 8 * world
    try:
        host, tsw = make_inputs(Vw, dev)
        imgs = host.to(dev)
        for _ in range(3):
            runner(imgs, tsw, CLASSES)
        g, _ = _capture(torch, lambda: runner(imgs, tsw, CLASSES)) if args.graph else (None, None)
        fn = (lambda: g.replay()) if g is not None else (lambda: runner(imgs, tsw, CLASSES))
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        fn()
        fn()
        sampler.mark()
        n = max(3, min(args.steps, 10))
        ms = _timed(torch, dist, world, dev, fn, n)
        clocks = sampler.stop() if rank == 0 else None
        del imgs, g
        return {"value": Vw * n / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / n, "steps": n, "scaling": "weak",
                "config": {"workload": f"{Vw}-keyframe 512x384 batch, 8 views per GPU, views sharded x{world}",
                           "views": Vw, "views_per_gpu": 8, "cuda_graph": args.graph},
                "clocks": clocks}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


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
    ap.add_argument("--head-precision", default="fp32", choices=["fp32", "bf16"],
                    help="panoptic head arithmetic: fp32 = the reference's policy (split-bf16 operands), bf16 = plain bf16")
    ap.add_argument("--no-bf16-head", dest="bf16_head", action="store_false", help="skip the secondary all-bf16-head timing")
    ap.add_argument("--no-ids-only", dest="ids_only", action="store_false", help="skip the lazy-masks + post-processing timing")
    ap.add_argument("--no-gpu-reference", dest="gpu_reference", action="store_false",
                    help="skip timing the reference's own PyTorch path on the GPU (N = 1)")
    ap.add_argument("--no-weak", dest="weak", action="store_false", help="N > 1: skip the 8-views-per-GPU (config 5) leg")
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
