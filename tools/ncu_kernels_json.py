#!/usr/bin/env python
"""Turns an `ncu --set full ... --page raw --csv` export into the per-kernel JSON committed under
profiles/: duration, DRAM bytes read + written per launch (the `traffic` of the roofline objects),
L2 hit rate, registers, occupancy, issue-slot utilisation and the dominant warp stalls.

    python tools/ncu_kernels_json.py gpurun_out/x_raw.csv [label] > profiles/x.json
"""
import csv
import json
import sys


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def to_bytes(value, unit):
    v = f(value)
    if v is None:
        return None
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(value, unit):
    v = f(value)
    if v is None:
        return None
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "nsecond": 1e-3}.get(unit, 1)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith(".ratio") and "not_issued" not in h and "selected" not in h.split("stalled_")[-1][:9]]
    out = {"source": sys.argv[1], "label": sys.argv[2] if len(sys.argv) > 2 else "", "kernels": []}

    def g(r, k):
        return r[idx[k]] if k in idx else ""

    for r in data:
        name = g(r, "Kernel Name").replace("void ", "").replace("mlsb::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        rd = to_bytes(g(r, "dram__bytes_read.sum"), units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(g(r, "dram__bytes_write.sum"), units[idx["dram__bytes_write.sum"]])
        st = sorted([(f(g(r, h)) or 0.0, h.split("issue_stalled_")[-1].replace("_per_issue_active.ratio", "")) for h in stall], reverse=True)[:4]
        out["kernels"].append({
            "kernel": name.split("(")[0][:60],
            "grid": g(r, "launch__grid_size"), "block": g(r, "launch__block_size"),
            "duration_us": to_us(g(r, "gpu__time_duration.sum"), units[idx["gpu__time_duration.sum"]]),
            "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": (rd or 0) + (wr or 0),
            "l2_hit_pct": f(g(r, "lts__t_sector_hit_rate.pct")),
            "registers": f(g(r, "launch__registers_per_thread")),
            "achieved_occupancy_pct": f(g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")),
            "issue_active_pct": f(g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")),
            "warp_inst_executed": f(g(r, "smsp__inst_executed.sum")),
            "pipe_fma_pct": f(g(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")),
            "pipe_alu_pct": f(g(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")),
            "top_stalls_per_issue": {n: round(v, 2) for v, n in st},
        })
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
