"""Exact-mode report: b200pt_debug_render_replay (the reference's loop shape + per-pixel LCG around the product's device
functions) against reference-made frames, per pixel.  Prints, per scene: share of pixels within 1e-4 / 1e-3 / 1e-2
relative, the frame rel-L2 and the mean ratio.  Usage: python tools/replay_report.py [--live W H SPP]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
GOLDEN = os.path.join(ROOT, "tests", "golden")
S = json.load(open(os.path.join(GOLDEN, "settings.json")))


def report(name, a, b):
    d = np.abs(a.astype(np.float64) - b).max(axis=2) / np.maximum(np.abs(b).max(axis=2), 1e-3)
    print(f"{name:48s} px<=1e-4 {np.mean(d <= 1e-4):.4f}  <=1e-3 {np.mean(d <= 1e-3):.4f}  <=1e-2 {np.mean(d <= 1e-2):.4f}  "
          f"rel_l2 {np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30):.2e}  mean ratio {a.mean() / max(b.mean(), 1e-30):.6f}", flush=True)


if "--fullsize" in sys.argv:
    # the BASELINE configs at their own resolution against the float16 frames of the reference build (tests/golden/make_fullsize_golden.py)
    import time
    for name, scene in (("mercury", "mercury"), ("dragon", "dragon"), ("matpreview", "matpreview"), ("volumetric-caustic", "volumetric-caustic")):
        g = np.load(os.path.join(GOLDEN, f"fullsize_{name}.npz"))
        w, h, spp = (int(v) for v in g["size"])
        r = pkg.Renderer(pkg.Scene(os.path.join(ROOT, "scenes", scene + ".b200scene")), device=0, max_paths_in_flight=1 << 20)
        t0 = time.time()
        a = r.render_replay(w, h, spp)
        dt = time.time() - t0
        b = g["frame"].astype(np.float32)
        d = np.abs(a.astype(np.float64) - b).max(axis=2) / np.maximum(np.abs(b).max(axis=2), 1e-3)
        print(f"fullsize {name} {w}x{h}x{spp} ({dt:.1f} s): px<=1e-3 {np.mean(d <= 1e-3):.4f}  <=2e-3 {np.mean(d <= 2e-3):.4f}  <=1e-2 {np.mean(d <= 1e-2):.4f}  "
              f"median {np.median(d):.2e}  rel_l2 {np.linalg.norm(a - b) / np.linalg.norm(b):.2e}  mean ratio {a.mean() / b.mean():.6f}", flush=True)
        # what float16 storage alone does to the frame
        q = a.astype(np.float16).astype(np.float32)
        print(f"    (float16 rounding of our own frame: rel_l2 {np.linalg.norm(a - q) / np.linalg.norm(a):.2e})", flush=True)
        r.close()
    sys.exit(0)
live = None
if "--live" in sys.argv:
    k = sys.argv.index("--live")
    live = tuple(int(v) for v in sys.argv[k + 1:k + 4])
    import refcheck
for scene, (w, h, spp) in sorted(S["exact"].items()):
    path = os.path.join(ROOT, "scenes", scene + ".b200scene")
    if not os.path.exists(path):
        print(scene, "pack missing")
        continue
    r = pkg.Renderer(pkg.Scene(path), device=0, max_paths_in_flight=1 << 20)
    try:
        report(scene, r.render_replay(w, h, spp), np.load(os.path.join(GOLDEN, f"exact_{scene}_woop.npy")))
        if live and scene in ("cornell-box", "volumetric-caustic", "mercury", "matpreview", "box", "lte-orb-silver"):
            expected, kind = refcheck.render_checker(path, *live)
            report(f"{scene} live {live} vs {kind}", r.render_replay(*live), expected)
    except pkg.MyException as e:
        print(scene, "refused:", e)
    r.close()
for scene, v in sorted(S["synthetic"].items()):
    w, h, spp = v["exact"]
    r = pkg.Renderer(pkg.Scene(os.path.join(GOLDEN, f"synthetic_{scene}.b200scene")), device=0, max_paths_in_flight=1 << 20)
    try:
        report("synthetic_" + scene, r.render_replay(w, h, spp), np.load(os.path.join(GOLDEN, f"exact_synthetic_{scene}_woop.npy")))
    except pkg.MyException as e:
        print("synthetic_" + scene, "refused:", e)
    r.close()
