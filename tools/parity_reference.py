#!/usr/bin/env python3
"""Renders the BASELINE scenes with the UNMODIFIED reference CPU renderer (oracle/_ref/libcsrt_ref_woop.so, i.e. `RayTracer --cpu`)
at the sizes of tools/parity_images.py and stores the linear frames (float16 .npy) under profiles/parity/.  Runs where
/root/reference was compiled; takes a while on 8 cores."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refcheck  # noqa: E402

CASES = json.load(open(os.path.join(ROOT, "profiles", "parity", "cases.json")))
ref = refcheck.ref_lib("woop")
for name, (w, h, spp) in CASES.items():
    out = os.path.join(ROOT, "profiles", "parity", f"{name}_reference_cpu.npy")
    if os.path.exists(out):
        continue
    t = time.time()
    frame, _, seconds = ref.render_pack(os.path.join(ROOT, "scenes", name + ".b200scene"), w, h, spp)
    np.save(out, frame.astype(np.float16))
    print(name, w, h, spp, f"render {seconds:.1f} s (wall {time.time() - t:.1f} s) mean {frame.mean():.5f}", flush=True)
