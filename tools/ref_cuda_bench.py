#!/usr/bin/env python3
"""On-box comparator: the reference's OWN CUDA backend (its one-thread-per-pixel megakernel, rebuilt for
sm_100a as oracle/_ref/libcsrt_ref_cuda.so) against this repo's wavefront path, same scene / size / spp, same B200.
Prints one JSON line per workload: both times, Msamples/s, and the parity metrics between the two frames."""
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


def run(name, w, h, spp, repeats=2):
    pack = os.path.join(ROOT, "scenes", name + ".b200scene")
    L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcsrt_ref_cuda.so"))
    L.ref_create_cuda.restype = ctypes.c_void_p
    L.ref_create_cuda.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    L.ref_draw_cuda.restype = ctypes.c_double
    L.ref_draw_cuda.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.ref_destroy_cuda.argtypes = [ctypes.c_void_p]
    L.ref_last_error.restype = ctypes.c_char_p
    scene = pkg.Scene(pack)
    build = ctypes.c_double()
    handle = L.ref_create_cuda(scene.desc, w, h, spp, ctypes.byref(build))
    if not handle:
        print(json.dumps({"scene": name, "error": L.ref_last_error().decode(errors="replace")}))
        return
    ref = np.zeros((h, w, 3), dtype=np.float32)
    times = []
    for _ in range(repeats):  # the first Draw also pages the managed scene in
        t = L.ref_draw_cuda(handle, ref.ctypes.data)
        if t < 0:
            print(json.dumps({"scene": name, "error": L.ref_last_error().decode(errors="replace")}))
            return
        times.append(t)
    L.ref_destroy_cuda(handle)
    r = pkg.Renderer(scene, device=0)
    ours = r.Draw(width=w, height=h, spp=spp, seed=1)
    ours = r.Draw(width=w, height=h, spp=spp, seed=1)
    ms = r.stats()["render_ms"]
    r.close()
    out = {"scene": name, "width": w, "height": h, "spp": spp, "reference_cuda_s": times, "reference_cuda_build_s": build.value,
           "reference_cuda_Msamples_s": w * h * spp / min(times) / 1e6, "b200pt_ms": ms, "b200pt_Msamples_s": w * h * spp / ms / 1e3,
           "speedup": min(times) * 1e3 / ms}
    out.update(refcheck.metrics(ours, ref))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    for args in (("cornell-box", 512, 512, 64), ("dragon", 1024, 1024, 256), ("matpreview", 1024, 1024, 64), ("volumetric-caustic", 512, 512, 256)):
        run(*args)
