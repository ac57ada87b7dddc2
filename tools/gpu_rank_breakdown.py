#!/usr/bin/env python3
"""Scratch: what ONE rank of an N-way tile split spends per kernel class (emulated on one GPU), and the frame time without timing."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
import torch  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w, h, spp = 1024, 1024, 256
scene = pkg.Scene(os.path.join(ROOT, "scenes", "dragon.b200scene"))
r = pkg.Renderer(scene, device=0)
tiles = torch.zeros(pkg.tile_buffer_floats(w, h, world), dtype=torch.float32, device="cuda")
for _ in range(3):
    r.draw_tiles_device(tiles, 0, world, w, h, spp, seed=1)
torch.cuda.synchronize()
ms = []
for _ in range(5):
    r.draw_tiles_device(tiles, 0, world, w, h, spp, seed=1)
    torch.cuda.synchronize()
    ms.append(r.stats()["render_ms"])
r.draw_tiles_device(tiles, 0, world, w, h, spp, seed=1, stats=pkg.STATS_TIMING)
torch.cuda.synchronize()
st = r.stats()
print(json.dumps({"world": world, "arenas": os.environ.get("B200PT_ARENAS", "auto"), "ms": [round(x, 3) for x in ms], "timed_total_ms": round(st["render_ms"], 3),
                  "launches": st["kernel_launches"], "active_tiles": st["active_tiles"],
                  **{k: round(st[k]["ms"], 3) for k in ("primary", "extend", "shade", "other", "tail")}}))
