#!/usr/bin/env python3
"""Scratch: timing breakdown of the Dragon workload for different wavefront capacities."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
import torch  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "dragon"
w, h, spp = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (1024, 1024, 256)
caps = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [24, 26, 28]
scene = pkg.Scene(os.path.join(ROOT, "scenes", name + ".b200scene"))
frame = torch.zeros(h * w * 3, dtype=torch.float32, device="cuda")
for cap in caps:
    r = pkg.Renderer(scene, device=0, max_paths_in_flight=1 << cap, max_leaf_size=int(os.environ.get("LEAF", "0")))
    for _ in range(2):
        r.draw_device(frame, w, h, spp, seed=1)
    torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        r.draw_device(frame, w, h, spp, seed=1)
        torch.cuda.synchronize()
        ms.append(r.stats()["render_ms"])
    r.draw_device(frame, w, h, spp, seed=1, stats=pkg.STATS_TIMING)
    torch.cuda.synchronize()
    st = r.stats()
    print(json.dumps({"cap_log2": cap, "nodes": st["num_bvh_nodes"], "bvh_build_ms": round(st["bvh_build_ms"], 1), "bvh_gpu_ms": round(st["bvh_gpu_ms"], 2), "ms": ms, "Msamples_s": w * h * spp / min(ms) / 1e3, "launches": st["kernel_launches"],
                      **{k: round(st[k]["ms"], 2) for k in ("primary", "extend", "shadow", "shade", "other", "tail")}}), flush=True)
    r.close()
