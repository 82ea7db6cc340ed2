"""Probe: can DINOv2 and the sequential memory build share the GPU through two green contexts (disjoint SM partitions)?
Development tool; prints what works on this driver (eager two-stream run, CUDA-graph capture across the partitions)."""
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from panst3r_b200 import ops  # noqa: E402
from panst3r_b200.panst3r import DEC_DIM, ENC_DIM, build_panst3r  # noqa: E402
from cuda.bindings import driver as cu  # noqa: E402


def ck(r):
    if isinstance(r, tuple):
        err, rest = r[0], r[1:]
    else:
        err, rest = r, ()
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"driver error {err}")
    return rest[0] if len(rest) == 1 else rest


def make_partitions(dino_sms):
    ck(cu.cuInit(0))
    dev = ck(cu.cuDeviceGet(0))
    res = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
    print("device SMs:", res.sm.smCount, flush=True)
    groups, nb, remaining = ck(cu.cuDevSmResourceSplitByCount(1, res, 0, dino_sms))
    print("split:", nb, "group of", groups[0].sm.smCount, "SMs; remaining", remaining.sm.smCount, flush=True)
    out = []
    for r in (groups[0], remaining):
        desc = ck(cu.cuDevResourceGenerateDesc([r], 1))
        g = ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        s = ck(cu.cuGreenCtxStreamCreate(g, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
        out.append((g, s, r.sm.smCount))
    return out


def main():
    dino_sms = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    V = 16
    with torch.device("cuda"):
        m = build_panst3r("v1")
    bench.init_weights_(m)
    m.overlap_dino = False
    m.panoptic_decoder.mask_transformer.overlap_aux_masks = False
    imgs, ts = bench.make_inputs(V, "cuda")
    imgs = imgs.cuda()
    cat, rows, x, pos, _ = m._features(imgs, ts)
    torch.cuda.synchronize()
    (gA, sA, nA), (gB, sB, nB) = make_partitions(dino_sms)
    stA = torch.cuda.ExternalStream(int(sA))
    stB = torch.cuda.ExternalStream(int(sB))

    def dino():
        with ops.sm_budget(nA):
            m.forward_dino(imgs, ts, out=rows[:, ENC_DIM + DEC_DIM:])

    def membuild():
        with ops.sm_budget(nB):
            return m.build_memory(x, pos, ts)

    def wall(fn, n=3):
        fn()
        torch.cuda.synchronize()
        stA.synchronize(); stB.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        stA.synchronize(); stB.synchronize()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    def on(st, f):
        def g():
            with torch.cuda.stream(st):
                f()
        return g

    try:
        print(f"eager dino on partition A ({nA} SMs): {wall(on(stA, dino)):.2f} ms", flush=True)
        print(f"eager membuild on partition B ({nB} SMs): {wall(on(stB, membuild)):.2f} ms", flush=True)

        def both():
            with torch.cuda.stream(stA):
                dino()
            with torch.cuda.stream(stB):
                membuild()
        print(f"eager both: {wall(both):.2f} ms", flush=True)
    except Exception:
        traceback.print_exc()

    # ---- graph capture across the two partitions
    try:
        def body():
            cur = torch.cuda.current_stream()
            stA.wait_stream(cur)
            with torch.cuda.stream(stA):
                dino()
            membuild()
            cur.wait_stream(stA)
        with torch.cuda.stream(stB):
            body()
        stA.synchronize(); stB.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stB):
            body()
        for name, st in (("partition-B stream", stB), ("default-context stream", torch.cuda.Stream())):
            with torch.cuda.stream(st):
                g.replay()
                st.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(5):
                    g.replay()
                e1.record(st)
                st.synchronize()
                print(f"graph (dino || membuild) replayed on {name}: {e0.elapsed_time(e1) / 5:.2f} ms", flush=True)
    except Exception:
        traceback.print_exc()

    # ---- same graph without partitions, for reference (serial)
    try:
        def serial():
            m.forward_dino(imgs, ts, out=rows[:, ENC_DIM + DEC_DIM:])
            m.build_memory(x, pos, ts)
        serial()
        torch.cuda.synchronize()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            serial()
        g2.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g2.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"graph serial (dino ; membuild) full GPU: {e0.elapsed_time(e1) / 5:.2f} ms", flush=True)
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
