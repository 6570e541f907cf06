#!/usr/bin/env python
"""Hot spots from ncu exports.
  python tools/ncu_hot.py raw gpurun_out/x_raw.csv          per-kernel summary (time, DRAM bytes, stalls)
  python tools/ncu_hot.py src gpurun_out/x_src_K.csv [N]    top-N source lines by stall samples (view: CUDA source if present)
"""
import csv
import sys


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[0], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith(".ratio") and "not_issued" not in h]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__grid_size", "launch__registers_per_thread", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    for r in data:
        print("==", r[idx["Kernel Name"]][:70])
        print("   " + "  ".join("%s=%s%s" % (k.split(".")[0].replace("launch__", "").replace("__", "."), r[idx[k]], rows[1][idx[k]][:6])
                                for k in keys if k in idx))
        st = sorted([(f(r[idx[h]]), h) for h in stall], reverse=True)[:6]
        print("   stalls: " + ", ".join("%s %.1f" % (h.split("issue_stalled_")[-1].replace("_per_issue_active.ratio", ""), v) for v, h in st))


def src(path, top=25):
    rows = list(csv.reader(open(path)))
    # several tables may follow each other (one per kernel instance / view); take each
    i = 0
    while i < len(rows):
        r = rows[i]
        if "Source" in r and "Instructions Executed" in r:
            hdr = r
            si, ie, ws = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
            j = i + 1
            body = []
            while j < len(rows) and len(rows[j]) > max(si, ie, ws) and rows[j][si] != "Source":
                body.append(rows[j])
                j += 1
            ti, ts = sum(f(b[ie]) for b in body) or 1, sum(f(b[ws]) for b in body) or 1
            print("---- table at row %d: %d lines, %.0f inst, %.0f samples" % (i, len(body), ti, ts))
            for b in sorted(body, key=lambda b: -f(b[ws]))[:top]:
                print("%5.1f%% stall %5.1f%% inst  %s" % (100 * f(b[ws]) / ts, 100 * f(b[ie]) / ti, b[si].strip()[:120]))
            i = j
        else:
            i += 1


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2])
    else:
        src(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
