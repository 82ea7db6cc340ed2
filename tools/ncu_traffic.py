"""Parse an `ncu --page raw --csv` export of tools/profile_step.py kernels into profiles/r02_ncu_traffic.json
(dram__bytes_read.sum + dram__bytes_write.sum per launch, duration, pipe utilisation) — bench.py reads the JSON for
`roofline.traffic` instead of carrying constants.

    ncu -i gpurun_out/r02_key_kernels.ncu-rep --page raw --csv > gpurun_out/r02_key_kernels_raw.csv
    python tools/ncu_traffic.py gpurun_out/r02_key_kernels_raw.csv
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# launch order of tools/profile_step.py kernels -> key used by bench.py
ORDER = ["attention3_render", "attention3_encoder_self", "gemm2_fc1_gelu", "gemm_membuild_qkv", "gemm_mask_logits_tma",
         "layernorm_12288x1024", "gemm2_split_fc2_promote", "gemm_mask_logits_split_tma", "gemm_membuild_fc2_splitk",
         "gemm_membuild_fc2_unsplit"]


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return None


def main():
    src = sys.argv[1]
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, scale_from_unit=True):
        if name not in col:
            return None
        v = num(r[col[name]])
        if v is None:
            return None
        u = units[col[name]].lower()
        if scale_from_unit:
            v *= {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9,
                  "second": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
        return v

    out = {}
    kernels = [r for r in data if "pst3r::" in r[col["Kernel Name"]] and "combine" not in r[col["Kernel Name"]]
               and "convert" not in r[col["Kernel Name"]]]
    for key, r in zip(ORDER, kernels):
        rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
        out[key] = {
            "kernel": r[col["Kernel Name"]][:80], "grid": r[col["Grid Size"]], "duration_us": (get(r, "gpu__time_duration.sum") or 0) * 1e6,
            "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": None if rd is None or wr is None else rd + wr,
            "tensor_pipe_pct": get(r, "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", False)
            or get(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", False),
            "dram_gbs": None if rd is None or wr is None else (rd + wr) / 1e9 / max(get(r, "gpu__time_duration.sum") or 1e-9, 1e-9),
            "sm_throughput_pct": get(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed", False),
            "registers": get(r, "launch__registers_per_thread", False),
            "source": f"ncu --set full --clock-control none, tools/profile_step.py kernels ({os.path.basename(src)})",
        }
    dst = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
