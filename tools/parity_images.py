#!/usr/bin/env python3
"""Parity artefacts (BASELINE.json: "image diff committed").

  on the GPU box:   python tools/parity_images.py render     -> gpurun_out/parity/<scene>_b200_seed{1,2}.npy (float16)
  here afterwards:  python tools/parity_images.py report     -> profiles/parity/<scene>_{b200,reference_cpu,diff}.png + parity.json

The reference frames come from tools/parity_reference.py (the reference's own `--cpu` renderer, same scene / size / spp).
Metrics per scene, all relative to the reference frame B: per-pixel rel-L2 ||A-B||/||B||, the same after 8x8 box filtering, the
whole-image mean ratio, and the noise floor = rel-L2 between two GPU renders that differ only in the seed."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARITY = os.path.join(ROOT, "profiles", "parity")
CASES = json.load(open(os.path.join(PARITY, "cases.json")))


def render():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    out = os.path.join(ROOT, "gpurun_out", "parity")
    os.makedirs(out, exist_ok=True)
    for name, (w, h, spp) in CASES.items():
        r = pkg.Renderer(pkg.Scene(os.path.join(ROOT, "scenes", name + ".b200scene")), device=0)
        for seed in (1, 2):
            frame = r.Draw(width=w, height=h, spp=spp, seed=seed)
            np.save(os.path.join(out, f"{name}_b200_seed{seed}.npy"), frame.astype(np.float16))
        print(name, w, h, spp, f"{r.stats()['render_ms']:.1f} ms", flush=True)
        r.close()


def srgb8(f):
    f = np.clip(f.astype(np.float32), 0.0, 1.0)
    return (np.where(f <= 0.0031308, 12.92 * f, 1.055 * np.power(f, 1 / 2.4) - 0.055) * 255.0 + 0.5).astype(np.uint8)


def report():
    from PIL import Image
    src = os.path.join(ROOT, "gpurun_out", "parity")
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    box = lambda f: f[: f.shape[0] // 8 * 8, : f.shape[1] // 8 * 8].reshape(f.shape[0] // 8, 8, f.shape[1] // 8, 8, 3).mean(axis=(1, 3))
    summary = {}
    for name, (w, h, spp) in CASES.items():
        a = np.load(os.path.join(src, f"{name}_b200_seed1.npy")).astype(np.float64)
        a2 = np.load(os.path.join(src, f"{name}_b200_seed2.npy")).astype(np.float64)
        b = np.load(os.path.join(PARITY, f"{name}_reference_cpu.npy")).astype(np.float64)
        summary[name] = {"width": w, "height": h, "spp": spp, "mean_ratio": float(a.mean() / b.mean()),
                         "rel_l2_per_pixel": rel(a, b), "rel_l2_box8": rel(box(a), box(b)),
                         "noise_floor_per_pixel": rel(a2, a), "noise_floor_box8": rel(box(a2), box(a)),
                         "max_abs_diff_box8": float(np.abs(box(a) - box(b)).max())}
        Image.fromarray(srgb8(a)).save(os.path.join(PARITY, f"{name}_b200.png"))
        Image.fromarray(srgb8(b)).save(os.path.join(PARITY, f"{name}_reference_cpu.png"))
        Image.fromarray(srgb8(np.abs(a - b) * 8.0)).save(os.path.join(PARITY, f"{name}_absdiff_x8.png"))
        print(name, json.dumps(summary[name]))
    json.dump(summary, open(os.path.join(PARITY, "parity.json"), "w"), indent=1)


if __name__ == "__main__":
    {"render": render, "report": report}[sys.argv[1]]()
