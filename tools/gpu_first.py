#!/usr/bin/env python3
"""Scratch GPU check: render the config scenes on the GPU and compare with the reference CPU build."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge  # noqa: E402
import refcheck  # noqa: E402

pkg = ge.load_package()
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def save_png(name, frame):
    from PIL import Image
    Image.fromarray(pkg.linear_to_srgb8(frame)).save(os.path.join(OUT, name))


def compare(name, w, h, spp, with_ref=True):
    pack = os.path.join(ROOT, "scenes", name + ".b200scene")
    if not os.path.exists(pack):
        print(name, "pack missing")
        return
    t0 = time.time()
    scene = pkg.Scene(pack)
    r = pkg.Renderer(scene, device=0)
    t1 = time.time()
    frame = r.Draw(width=w, height=h, spp=spp, seed=7, stats=True)
    st = r.stats()
    frame2 = r.Draw(width=w, height=h, spp=spp, seed=7)
    st2 = r.stats()
    out = {"scene": name, "w": w, "h": h, "spp": spp, "create_s": t1 - t0, "render_ms_stats": st["render_ms"],
           "render_ms": st2["render_ms"], "Msamples_s": w * h * spp / st2["render_ms"] / 1e3,
           "closest_per_sample": st["closest_rays"] / (w * h * spp), "shadow_per_sample": st["shadow_rays"] / (w * h * spp),
           "nodes_per_sample": st["node_visits"] / (w * h * spp), "prims_per_sample": st["prim_tests"] / (w * h * spp),
           "launches": st2["kernel_launches"], "bvh_nodes": st["num_bvh_nodes"], "bvh_build_ms": st["bvh_build_ms"],
           "deterministic": bool(np.array_equal(frame, frame2)), "nan": int(np.isnan(frame).sum())}
    save_png(f"{name}_gpu.png", frame)
    np.save(os.path.join(OUT, f"{name}_gpu.npy"), frame)
    if with_ref:
        t2 = time.time()
        ref, b, rs = refcheck.ref_lib("woop").render_pack(pack, w, h, spp)
        out.update(refcheck.metrics(frame, ref))
        out["ref_render_s"] = rs
        out["ref_Msamples_s"] = w * h * spp / rs / 1e6
        save_png(f"{name}_ref.png", ref)
        np.save(os.path.join(OUT, f"{name}_ref.npy"), ref)
    print(json.dumps(out), flush=True)
    r.close()


if __name__ == "__main__":
    print("cpus", os.cpu_count())
    compare("cornell-box", 256, 256, 64)
    compare("dragon", 512, 512, 16)
    compare("volumetric-caustic", 256, 256, 64)
    compare("matpreview", 256, 256, 32)
    compare("mercury", 256, 256, 32)
    compare("dragon", 1024, 1024, 256, with_ref=False)
