#!/usr/bin/env python3
"""SASS-level source page of one kernel of an .ncu-rep as a gzip'ed CSV (every column ncu exports; parsed with the csv module,
the SASS text contains commas).  usage: ncu_source.py report.ncu-rep 'regex:k_trace' out.csv.gz [launch-skip]"""
import csv
import gzip
import io
import subprocess
import sys

rep, kernel, out = sys.argv[1:4]
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel, "--launch-skip", skip, "--launch-count", "1"],
                      capture_output=True, text=True).stdout
lines = text.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"') or l.startswith('"#"') or '"Source"' in l)
rows = list(csv.reader(lines[start:]))
with gzip.open(out, "wt", newline="") as f:
    w = csv.writer(f)
    for line in lines[:start]:
        f.write("# " + line + "\n")
    w.writerows(rows)
print(f"{out}: {len(rows) - 1} SASS rows, columns: {rows[0]}")
