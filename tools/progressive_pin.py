#!/usr/bin/env python3
"""Pins the CPU restatement of the preview path (oracle_render_progressive) against the reference's OWN code: the reference runs
Renderer::Draw(index_frame, frame, frame_srgb) only on its CUDA backend, so this needs a GPU box and oracle/_ref/libcsrt_ref_cuda.so.
Both walk the same per-pixel LCG streams (seed Tea<4>(pixel_offset, index_frame)); they differ only by device vs host float
contraction.  Also renders the same frames with b200pt_render_progressive_device.  Prints one JSON line."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge  # noqa: E402
import refcheck  # noqa: E402

pkg = ge.load_package()
import torch  # noqa: E402

name, w, h, frames = "cornell-box", 64, 64, 64
pack = os.path.join(ROOT, "scenes", name + ".b200scene")
L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcsrt_ref_cuda.so"))
L.ref_create_cuda.restype = ctypes.c_void_p
L.ref_create_cuda.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
L.ref_draw_progressive_cuda.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
L.ref_destroy_cuda.argtypes = [ctypes.c_void_p]
L.ref_last_error.restype = ctypes.c_char_p
scene = pkg.Scene(pack)
handle = L.ref_create_cuda(scene.desc, w, h, 1, None)
ref, ref_srgb = np.zeros((h, w, 3), np.float32), np.zeros((h, w, 3), np.float32)
if not handle or L.ref_draw_progressive_cuda(handle, frames, ref.ctypes.data, ref_srgb.ctypes.data) != 0:
    print(json.dumps({"error": L.ref_last_error().decode(errors="replace")}))
    sys.exit(0)
L.ref_destroy_cuda(handle)
port, port_srgb = refcheck.OracleLib().render_progressive(pack, w, h, frames)
r = pkg.Renderer(scene, device=0)
frame = torch.zeros(h * w * 3, dtype=torch.float32, device="cuda")
for k in range(frames):
    r.draw_progressive_device(frame, None, k, width=w, height=h, seed=5)
torch.cuda.synchronize()
ours = frame.cpu().numpy().reshape(h, w, 3)
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))
box = lambda f: f.reshape(h // 8, 8, w // 8, 8, 3).mean(axis=(1, 3))
print(json.dumps({"scene": name, "size": [w, h], "frames": frames,
                  "port_vs_reference": {"identical_pixels": float((port == ref).all(axis=2).mean()), "rel_l2": rel(port, ref), "srgb_rel_l2": rel(port_srgb, ref_srgb),
                                        "mean_ratio": float(port.mean() / ref.mean())},
                  "b200pt_vs_reference": {"rel_l2_box8": rel(box(ours), box(ref)), "mean_ratio": float(ours.mean() / ref.mean())}}))
