"""In-kernel timeline of attention3 (development tool): runs the render cross-attention once with PST3R_ATT_TRACE set
(CTA (0,0,0) records clock64() at its hand-over points) and prints, for the steady-state tiles, the mean interval between
consecutive events of the MMA thread and of one softmax thread per (sub-tile, half).

    python tools/attn_trace.py [out.md]
"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE = os.path.join(ROOT, "gpurun_out", "attn_trace.txt")
os.makedirs(os.path.dirname(TRACE), exist_ok=True)
os.environ["PST3R_ATT_TRACE"] = TRACE

import torch  # noqa: E402
from panst3r_b200 import ops  # noqa: E402

MMA_EV = {0: "k/v of the tile landed", 1: "p_full[0] seen", 2: "S_0(j+1), PV_0(j) issued", 3: "p_full[1] seen", 4: "S_1(j+1), PV_1(j) issued"}
SM_EV = {0: "s_full seen", 1: "max pass done", 2: "half-row exchange done", 3: "o_full(j-1) seen", 4: "O folded",
         5: "exp sweep issued", 6: "P stored, p_full arrive"}


MODES = {0: "full kernel", 1: "no max pass (no first read of the scores)", 2: "no read / fold of O", 3: "neither max pass nor O fold",
         4: "no exp sweep (no read of S, no exponentials, no store of P)", 7: "softmax warps only hand the barriers over (pure MMA rate)",
         8: "sweep = tcgen05.ld + tcgen05.st only, no arithmetic", 16: "sweep arithmetic on register values, no tcgen05.ld of the scores",
         19: "sweep arithmetic only: no tensor-memory reads at all"}


def run(q, k, v, mode):
    os.environ["PST3R_ATT_TRACE_MODE"] = str(mode)
    ops.attention(q, k, v)
    torch.cuda.synchronize()
    ev = defaultdict(dict)
    for line in open(TRACE):
        sl, e, j, t = line.split()
        ev[(int(sl), int(e))][int(j)] = int(t)
    return ev


def mean(xs):
    return sum(xs) / len(xs)


def main():
    B, H, Nq, Nk = 16, 12, 768, 12288
    r = lambda *s: torch.randn(*s, device="cuda").bfloat16()
    q, k, v = r(B, Nq, H, 64), r(1, Nk, H, 64), r(1, Nk, H, 64)
    lo, hi = 16, 80  # steady state
    rows = []
    for mode in (1, 2, 3, 4, 7, 8, 16, 19):
        ev = run(q, k, v, mode)
        per = mean([ev[(0, 1)][j + 1] - ev[(0, 1)][j] for j in range(lo, hi)])
        iss0 = mean([ev[(0, 2)][j] - ev[(0, 1)][j] for j in range(lo, hi)])
        iss1 = mean([ev[(0, 4)][j] - ev[(0, 3)][j] for j in range(lo, hi)])
        s_lat = mean([ev[(1, 0)][j + 1] - ev[(0, 1)][j] for j in range(lo, hi)])
        o_lat = mean([ev[(1, 3)][j + 1] - ev[(0, 1)][j] for j in range(lo, hi)])
        rows.append(f"| {mode} | {MODES[mode]} | {per:.0f} | {iss0:.0f} / {iss1:.0f} | {s_lat:.0f} | {o_lat:.0f} |")
    ev = run(q, k, v, 0)
    out = ["# attention3 in-kernel timeline (CTA (0,0,0), render cross-attention B16 H12 Nq768 Nk12288; clock64 cycles)", ""]
    t_first = min(min(d.values()) for d in ev.values())
    t_last = max(max(d.values()) for d in ev.values())
    n_tiles = max(max(d.keys()) for d in ev.values()) + 1
    out.append(f"traced span {t_last - t_first} clk over {n_tiles} key tiles = {(t_last - t_first) / n_tiles:.0f} clk per tile (two sub-tiles each)")
    per = [ev[(0, 1)][j + 1] - ev[(0, 1)][j] for j in range(lo, hi)]
    out.append(f"steady-state period (p_full[0] to p_full[0]): mean {sum(per) / len(per):.0f}, min {min(per)}, max {max(per)}")
    out += ["", "## MMA thread: mean clk from the previous event (tiles 16..79)", "", "| event | clk since previous | ", "|---|---|"]
    order = [(0, 0), (0, 1), (0, 2), (0, 3), (0, 4)]
    for a, b in zip(order, order[1:]):
        d = [ev[b][j] - ev[a][j] for j in range(lo, hi)]
        out.append(f"| {MMA_EV[b[1]]} | {sum(d) / len(d):.0f} |")
    d = [ev[(0, 0)][j + 1] - ev[(0, 4)][j] for j in range(lo, hi)]
    out.append(f"| next tile: {MMA_EV[0]} | {sum(d) / len(d):.0f} |")
    for sl in (1, 2, 3, 4):
        t, half = (sl - 1) % 2, (sl - 1) // 2
        out += ["", f"## softmax thread, sub-tile {t}, key half {half}: mean clk from the previous event", "", "| event | clk since previous |", "|---|---|"]
        seq = [0, 1, 2, 3, 4, 5, 6]
        for a, b in zip(seq, seq[1:]):
            d = [ev[(sl, b)][j] - ev[(sl, a)][j] for j in range(lo, hi)]
            out.append(f"| {SM_EV[b]} | {sum(d) / len(d):.0f} |")
        d = [ev[(sl, 0)][j + 1] - ev[(sl, 6)][j] for j in range(lo, hi)]
        out.append(f"| next tile: {SM_EV[0]} (wait for the scores) | {sum(d) / len(d):.0f} |")
    # phase of sub-tile 1 against sub-tile 0
    d = [ev[(2, 0)][j] - ev[(1, 0)][j] for j in range(lo, hi)]
    out += ["", f"sub-tile 1 sees its scores {sum(d) / len(d):.0f} clk after sub-tile 0 (mean; min {min(d)}, max {max(d)})"]
    # raw rows for a few tiles
    out += ["", "## raw (tile 40..43), relative to tile 40's first event", "", "```"]
    base = ev[(0, 0)][40]
    for j in range(40, 44):
        out.append(f"tile {j}: mma " + " ".join(str(ev[(0, e)][j] - base) for e in range(5)))
        for sl in (1, 2, 3, 4):
            out.append(f"   sm slot {sl}: " + " ".join(str(ev[(sl, e)][j] - base) for e in range(7)))
    out.append("```")
    out += ["", "## Timing experiments (traced kernel with parts of the softmax side removed; results are wrong, only the clock counts matter)", "",
            "| mode | what runs | period clk / key tile | MMA thread blocked issuing S(j+1)+PV(j), sub-tile 0 / 1 | p_full[0] seen -> S_0(j+1) visible | p_full[0] seen -> PV_0(j) visible |",
            "|---|---|---|---|---|---|"] + rows
    text = "\n".join(out) + "\n"
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)


if __name__ == "__main__":
    main()
