#!/usr/bin/env python3
"""Renders exactly ONE frame of a workload on cuda:0 (no warm-up) — the process ncu wraps for per-kernel captures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
import torch  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "dragon"
w, h, spp = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (1024, 1024, 256)
scene = pkg.Scene(os.path.join(ROOT, "scenes", name + ".b200scene"))
frame = torch.zeros(h * w * 3, dtype=torch.float32, device="cuda")
r = pkg.Renderer(scene, device=0)
r.draw_device(frame, w, h, spp, seed=1)
torch.cuda.synchronize()
print("render_ms", r.stats()["render_ms"], "launches", r.stats()["kernel_launches"])
r.close()
