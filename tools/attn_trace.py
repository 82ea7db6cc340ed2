"""In-kernel timeline of the hd-64 attention kernel (development tool): runs the render cross-attention once with
PST3R_ATT_TRACE set (CTA (0,0,0) of the traced template instance records clock64() at its hand-over points) and prints, for the
steady-state tiles, the mean interval between consecutive events of the MMA threads and of one softmax thread per (sub-tile, half).

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

def run(q, k, v, mode=0):
    ops.attention(q, k, v)
    torch.cuda.synchronize()
    ev = defaultdict(dict)
    for line in open(TRACE):
        sl, e, j, t = line.split()
        ev[(int(sl), int(e))][int(j)] = int(t)
    return ev


def mean(xs):
    return sum(xs) / len(xs)


MMA5 = [(0, "K(j+1), V(j) landed"), (1, "s_free[0](j) seen, S_0(j+1) issued"), (2, "p_full[0](j) seen"), (3, "P V_0(j) issued"),
        (4, "s_free[1](j) seen, S_1(j+1) issued"), (5, "p_full[1](j) seen"), (6, "P V_1(j) issued")]
SM5 = [(0, "s_full seen"), (1, "maximum pass done (2 loads in flight)"), (2, "half-row exchange done"), (3, "o_full(j-1) seen"),
       (4, "sweep token (the other sub-tile has finished its sweep) seen"), (5, "exp sweep issued (4 x 16 columns)"),
       (6, "P stored (tcgen05.wait::st), p_full arrive")]


def main5(q, k, v, lo, hi):
    """attention5 (256 queries per CTA, alternating sweeps): event scheme of attention5.cuh"""
    ev = run(q, k, v, 0)
    out = ["# attention5 in-kernel timeline (CTA (0,0,0), render cross-attention B16 H12 Nq768 Nk12288; clock64 cycles)", ""]
    per = [ev[(0, 2)][j + 1] - ev[(0, 2)][j] for j in range(lo, hi)]
    out.append(f"steady-state period per key tile (both sub-tiles, 256 queries): mean {mean(per):.0f}, min {min(per)}, max {max(per)}")
    out += ["", "## MMA thread: mean clk from the previous event (tiles 16..79)", "", "| event | clk since previous |", "|---|---|"]
    for (a, _), (b, name) in zip(MMA5, MMA5[1:]):
        out.append(f"| {name} | {mean([ev[(0, b)][j] - ev[(0, a)][j] for j in range(lo, hi)]):.0f} |")
    out.append(f"| next tile: {MMA5[0][1]} | {mean([ev[(0, 0)][j + 1] - ev[(0, 6)][j] for j in range(lo, hi)]):.0f} |")
    for sl in (1, 2, 3, 4):
        t, half = (sl - 1) % 2, (sl - 1) // 2
        out += ["", f"## softmax thread, sub-tile {t}, key half {half}: mean clk from the previous event", "", "| event | clk since previous |", "|---|---|"]
        for (a, _), (b, name) in zip(SM5, SM5[1:]):
            out.append(f"| {name} | {mean([ev[(sl, b)][j] - ev[(sl, a)][j] for j in range(lo, hi)]):.0f} |")
        out.append(f"| next tile: {SM5[0][1]} | {mean([ev[(sl, 0)][j + 1] - ev[(sl, 6)][j] for j in range(lo, hi)]):.0f} |")
    out += ["", "## raw (tile 40..43), relative to tile 40's first MMA event", "", "```"]
    base = ev[(0, 0)][40]
    for j in range(40, 44):
        out.append(f"tile {j}: mma " + " ".join(str(ev[(0, e)][j] - base) for e, _ in MMA5))
        for sl in (1, 2, 3, 4):
            out.append(f"   sm slot {sl}: " + " ".join(str(ev[(sl, e)][j] - base) for e, _ in SM5))
    out.append("```")
    text = "\n".join(out) + "\n"
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)


def main():
    B, H, Nq, Nk = 16, 12, 768, 12288
    r = lambda *s: torch.randn(*s, device="cuda").bfloat16()
    q, k, v = r(B, Nq, H, 64), r(1, Nk, H, 64), r(1, Nk, H, 64)
    main5(q, k, v, 16, 80)  # steady-state tiles


if __name__ == "__main__":
    main()
