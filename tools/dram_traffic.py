#!/usr/bin/env python3
"""Aggregates an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of ONE frame
into per-kernel DRAM traffic per launch (the `roofline.traffic` figure of bench.py).  usage: dram_traffic.py in.csv out.json"""
import collections
import csv
import json
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in csv.DictReader(lines[start:]):
    m = re.search(r"(k_\w+)", r["Kernel Name"])
    agg[m.group(1) if m else r["Kernel Name"][:40]][r["Metric Name"]].append(float(r["Metric Value"].replace(",", "")))
out = {}
for k, v in agg.items():
    n = len(v["gpu__time_duration.sum"])
    rd, wr, t = sum(v["dram__bytes_read.sum"]), sum(v["dram__bytes_write.sum"]), sum(v["gpu__time_duration.sum"])
    print(f"{k:45s} launches {n:3d}  read {rd / 1e6:10.1f} MB  write {wr / 1e6:10.1f} MB  time {t / 1e6:9.3f} ms  -> {(rd + wr) / t:8.1f} GB/s")
    if k.startswith("k_"):
        out[k] = (rd + wr) / n
        out[k + "_detail"] = {"launches": n, "dram_read_bytes_total": rd, "dram_write_bytes_total": wr, "ncu_time_ms_total": t / 1e6}
out["_how"] = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over every kernel of ONE "
               "Dragon 1024x1024x256spp frame (tools/one_frame.py); value per kernel = (read+write bytes) / launches, i.e. per launch like roofline.achieved")
json.dump(out, open(sys.argv[2], "w"), indent=1)
