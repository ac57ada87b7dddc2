#!/usr/bin/env python3
"""Fills the @...@ placeholders of README.md's status table from profiles/r02_bench_n1.json (the bench line of the evidence run)."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")).read().strip().splitlines()[-1])
c = d["configs"]
fmt = lambda x: f"{x:,.0f}".replace(",", " ")
values = {
    "C2V": fmt(d["value"]), "C2MS": f"{d['ms_per_step']:.1f}", "C2E": fmt(d["e2e"]["value"]),
    "C1": fmt(c["C1"]["Msamples_s"]), "C3": fmt(c["C3"]["Msamples_s"]), "C4": fmt(c["C4"]["Msamples_s"]), "C5": fmt(c["C5"]["Msamples_s"]),
    "CPU": f"{d['cpu_baseline']['value']:.1f}", "CPUF": f"{d['cpu_baseline']['fast_build']['value']:.1f}",
    "CREATE": f"{d['config']['scene_create_s']:.2f}", "INIT": f"{d['config'].get('library_init_s', float('nan')):.2f}",
}
path = os.path.join(ROOT, "README.md")
text = open(path).read()
template = os.path.join(ROOT, "tools", "README.status.template")
if "@C2V@" in text:
    open(template, "w").write(text)          # keep the template so the numbers can be refreshed
else:
    text = open(template).read()
open(path, "w").write(re.sub(r"@(\w+)@", lambda m: values[m.group(1)], text))
print(values)
