#!/usr/bin/env python3
"""Diagnostic (GPU box): where do the two tree layouts / the reference disagree on the probe rays of tests/test_gpu_traversal.py?
Prints every ray whose closest hit differs between B200PT_CREATE_BVH8, the default binary tree and TLAS::Intersect."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge
import refcheck
import test_gpu_traversal as tt

pkg = ge.load_package()
scene = sys.argv[1] if len(sys.argv) > 1 else "synthetic_dielectrics_conductor_cylinder"
path = os.path.join(ROOT, "tests", "golden", scene + ".b200scene") if scene.startswith("synthetic_") else os.path.join(ROOT, "scenes", scene + ".b200scene")
sc = pkg.Scene(path)
tracer = refcheck.RefTracer(refcheck.ref_lib("woop"), path)
table = tt.instance_table(sc.desc)
print("instances (first tri, count, analytic rank):", table)
rng = np.random.RandomState(3)
probe = tt.make_rays(rng.randn(4096, 3) * 1e3, rng.randn(4096, 3))
results = {}
for flags in (0, pkg.CREATE_BVH8):
    r = pkg.Renderer(sc, device=0, flags=flags)
    far = tt.make_rays(rng.randn(20000, 3) * 50.0, rng.randn(20000, 3))
    far[:, 3:6] = -far[:, 0:3] / np.linalg.norm(far[:, 0:3], axis=1, keepdims=True)
    t, prim, _ = r.debug_trace(far)
    pts = far[:, 0:3] + t[:, None] * far[:, 3:6]
    pts = pts[prim != 0xFFFFFFFF]
    lo, hi = (pts.min(axis=0), pts.max(axis=0)) if len(pts) > 16 else (np.full(3, -5.0), np.full(3, 5.0))
    rays = np.concatenate([tt.random_rays(lo, hi, 60000, rng), probe, far[:2000]])
    rays[::5, 7] = np.float32(0.5) * np.linalg.norm(hi - lo)
    ref = tracer.trace(rays)
    for other_flags in (0, pkg.CREATE_BVH8):
        r2 = pkg.Renderer(sc, device=0, flags=other_flags) if other_flags != flags else r
        for per_lane in (False, True):
            t, prim, uv = r2.debug_trace(rays, per_lane_loop=per_lane)
            inst, local = tt.ours_to_instance(prim, table)
            bad = np.flatnonzero((prim != 0xFFFFFFFF) & (ref["valid"] != 0) & (np.abs(t / np.where(ref["t"] != 0, ref["t"], 1) - 1) > 1e-5) & ((prim & 0x80000000) == 0))
            print(f"rays of pass flags={flags}; tree flags={other_flags} per_lane={per_lane}: {len(bad)} triangle hits with another distance than the reference")
            for k in bad[:12]:
                print("   ray", k, rays[k].tolist(), "ours t", t[k], "prim", hex(int(prim[k])), "inst/local", inst[k], local[k],
                      "| ref t", ref["t"][k], "inst/prim", ref["id_instance"][k], ref["id_primitive"][k])
        if r2 is not r:
            r2.close()
    r.close()
tracer.close()
