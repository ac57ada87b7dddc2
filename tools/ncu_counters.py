#!/usr/bin/env python3
"""Aggregates an ncu --csv launch list of ONE frame (tools/one_frame.py under ncu with the metric list below) into per-kernel
counters: what bench.py's `roofline` block quotes next to the §8d byte formula (DRAM and L2 bytes actually moved, issue
utilisation, active lanes per instruction).

    ncu --metrics $(python tools/ncu_counters.py --metrics) --clock-control none --csv --log-file raw.csv python tools/one_frame.py dragon 1024 1024 256
    python tools/ncu_counters.py raw.csv profiles/r02_counters_dragon_1024x1024x256.json "dragon 1024x1024x256"
"""
import collections
import csv
import json
import re
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1.0, "ns": 1.0, "usecond": 1e3, "us": 1e3,
              "msecond": 1e6, "ms": 1e6, "second": 1e9, "s": 1e9}


def kernel_key(name):
    """k_shade<0, 2> and k_shade<1, 5> are different kernels; k_trace<0,0,0> keeps its template arguments too."""
    m = re.search(r"(k_\w+)(<[^>(]*>)?", name)
    if not m:
        return name[:40]
    args = re.sub(r"\(\w+\)", "", m.group(2) or "").replace(" ", "")
    return m.group(1) + args


def main():
    if sys.argv[1:2] == ["--metrics"]:
        print(",".join(METRICS))
        return
    lines = open(sys.argv[1]).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    per_launch = collections.OrderedDict()  # launch id -> {kernel, metric: value}
    for r in csv.DictReader(lines[start:]):
        rec = per_launch.setdefault(r["ID"], {"kernel": kernel_key(r["Kernel Name"])})
        value = float(r["Metric Value"].replace(",", "")) * UNIT_SCALE.get(r.get("Metric Unit", ""), 1.0)
        rec[r["Metric Name"]] = value
    agg = collections.OrderedDict()
    for rec in per_launch.values():
        a = agg.setdefault(rec["kernel"], collections.defaultdict(float))
        t = rec.get("gpu__time_duration.sum", 0.0)
        a["launches"] += 1
        a["time_ns"] += t
        a["dram_read_bytes"] += rec.get("dram__bytes_read.sum", 0.0)
        a["dram_write_bytes"] += rec.get("dram__bytes_write.sum", 0.0)
        a["l2_bytes"] += rec.get("lts__t_bytes.sum", 0.0)
        a["warp_inst"] += rec.get("smsp__inst_executed.sum", 0.0)
        a["thread_inst"] += rec.get("smsp__thread_inst_executed.sum", 0.0)
        for pct, key in (("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                         ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
                         ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct")):
            a[key + "_x_time"] += rec.get(pct, 0.0) * t
    out = {"_what": sys.argv[3] if len(sys.argv) > 3 else "", "_how": "ncu --metrics " + ",".join(METRICS) + " --clock-control none over every kernel of ONE frame "
           "(tools/one_frame.py, no warm-up: cold-cache, serialised launches); time-weighted averages for the percentages",
           "_frame_time_ms": sum(a["time_ns"] for a in agg.values()) / 1e6, "kernels": {}}
    total = sum(a["time_ns"] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_ns"]):
        t = max(a["time_ns"], 1.0)
        rec = {"launches": int(a["launches"]), "time_ms": a["time_ns"] / 1e6, "share_of_frame": a["time_ns"] / max(total, 1.0),
               "dram_read_bytes": a["dram_read_bytes"], "dram_write_bytes": a["dram_write_bytes"], "dram_gbs": (a["dram_read_bytes"] + a["dram_write_bytes"]) / t,
               "l2_bytes": a["l2_bytes"], "l2_gbs": a["l2_bytes"] / t, "warp_inst": a["warp_inst"], "thread_inst": a["thread_inst"],
               "active_lanes": a["thread_inst"] / max(a["warp_inst"], 1.0), "issue_active": a["issue_active_pct_x_time"] / t / 100.0,
               "warps_active": a["warps_active_pct_x_time"] / t / 100.0, "l1_hit": a["l1_hit_pct_x_time"] / t / 100.0, "l2_hit": a["l2_hit_pct_x_time"] / t / 100.0}
        out["kernels"][k] = rec
    json.dump(out, open(sys.argv[2], "w"), indent=1)  # before the table: a reader that closes the pipe (| head) must not cost the file
    for k, rec in out["kernels"].items():
        print(f"{k:34s} x{rec['launches']:4d} {rec['time_ms']:9.3f} ms {100 * rec['share_of_frame']:5.1f}%  DRAM {rec['dram_gbs']:7.1f} GB/s  L2 {rec['l2_gbs']:8.1f} GB/s  "
              f"issue {100 * rec['issue_active']:5.1f}%  lanes {rec['active_lanes']:5.2f}  warps {100 * rec['warps_active']:5.1f}%  L1 hit {100 * rec['l1_hit']:5.1f}%  L2 hit {100 * rec['l2_hit']:5.1f}%")


if __name__ == "__main__":
    main()
