#!/usr/bin/env python3
"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference build (oracle/_ref).

Run once in a container that has /root/reference (after `python -c "import __graft_entry__ as g; g.build()"`):

    python tests/golden/make_golden.py [--synthetic NAME ... | --scenes NAME ... | --kats]
        --synthetic: only (re)make the named synthetic scenes;  --scenes: only the named reference scenes (scenes/NAME.b200scene);
        --kats: only the known answers of leaf functions

Outputs (all produced by reference code, none by the oracle restatement or the CUDA path):
  kats.json                      known answers of reference leaf functions (Tea<4>, LCG, VdC, MisWeight, cos-hemisphere, LBVH)
  kulla_conty.npz                csrt::ComputeKullaConty tables (kulla_conty.cpp:62-80)
  exact_<scene>_<mt|woop>.npy    tiny frames (per-pixel LCG, deterministic): the C restatement must match them BIT FOR BIT
  converged_<scene>.npy          64x64 frames at high spp (Woop build): statistical target for the Philox-driven CUDA path
"""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refcheck  # noqa: E402

EXACT = {  # scene -> (w, h, spp)
    "cornell-box": (24, 24, 4), "dragon": (32, 32, 2), "mercury": (24, 24, 4), "matpreview": (24, 24, 2),
    "volumetric-caustic": (24, 24, 4),
    # the other scenes the reference ships (resources/scene/*)
    "box": (24, 24, 4), "classroom": (24, 24, 4), "dining-room": (24, 24, 4), "lte-orb-silver": (24, 24, 4),
    "lte-orb-rough-glass": (24, 24, 4), "material-testball": (24, 24, 4),
}
CONVERGED = {  # scene -> (w, h, spp)
    "cornell-box": (64, 64, 1024), "dragon": (64, 64, 1024), "mercury": (64, 64, 256), "matpreview": (64, 64, 512),
    "volumetric-caustic": (64, 64, 1024),
    "box": (64, 64, 1024), "classroom": (64, 64, 512), "dining-room": (64, 64, 1024), "lte-orb-silver": (64, 64, 512),
    "lte-orb-rough-glass": (64, 64, 1024), "material-testball": (64, 64, 512),
}


def main():
    woop, mt = refcheck.ref_lib("woop"), refcheck.ref_lib("mt")
    L = woop.lib
    only_synthetic = sys.argv[2:] if len(sys.argv) > 2 and sys.argv[1] == "--synthetic" else None
    if len(sys.argv) > 2 and sys.argv[1] == "--scenes":
        make_reference_frames(woop, mt, sys.argv[2:])
        settings_path = os.path.join(HERE, "settings.json")
        settings = json.load(open(settings_path))
        settings["exact"], settings["converged"] = EXACT, CONVERGED
        with open(settings_path, "w") as f:
            json.dump(settings, f, indent=1)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--kats":  # only the known answers of the leaf functions (no renders)
        make_kats(L)
        return
    if only_synthetic is None:
        make_reference_scene_goldens(woop, mt, L)
    make_synthetic_goldens(woop, mt, only_synthetic)


def make_reference_scene_goldens(woop, mt, L):
    make_kats(L)
    make_reference_frames(woop, mt)


def make_kats(L):
    kats = {"tea4": [[a, b, int(L.ref_tea4(a, b))] for a, b in ((0, 0), (3, 0), (3145725, 0), (12, 7), (4294967295, 1))]}
    seed = ctypes.c_uint32(L.ref_tea4(0, 0))
    kats["lcg_from_tea4_0_0"] = [[float(L.ref_random_float(ctypes.byref(seed))), int(seed.value)] for _ in range(8)]
    kats["van_der_corput2"] = [[i, float(L.ref_van_der_corput2(i))] for i in list(range(0, 17)) + [255, 256, 4097, 65535]]
    kats["van_der_corput3"] = [[i, float(L.ref_van_der_corput3(i))] for i in list(range(0, 17)) + [26, 27, 242, 243, 6560, 59049, 1000003]]
    kats["mis_weight"] = [[a, b, float(L.ref_mis_weight(a, b))] for a, b in ((0.3, 0.1), (1.0, 1.0), (0.01, 5.0), (7.5, 0.0))]
    L.ref_sample_hemis_cos.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
    hemis = []
    for xi in ((0.25, 0.5), (0.0, 0.0), (0.999, 0.123), (0.5, 0.75)):
        v = np.zeros(3, dtype=np.float32)
        pdf = ctypes.c_float()
        L.ref_sample_hemis_cos(xi[0], xi[1], v.ctypes.data, ctypes.byref(pdf))
        hemis.append([xi[0], xi[1], [float(x) for x in v], float(pdf.value)])
    kats["sample_hemis_cos"] = hemis
    # LBVH of 5 unit boxes at x = 0..4 with areas 1..5
    L.ref_build_bvh.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    L.ref_build_bvh.restype = ctypes.c_uint32
    rng = np.random.RandomState(5)
    bvh_cases = []
    for n, boxes in ((5, None), (37, "random")):
        if boxes is None:
            lo = np.stack([np.arange(n, dtype=np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)], axis=1)
        else:
            lo = rng.rand(n, 3).astype(np.float32) * 10
        bb = np.concatenate([lo, lo + 1.0], axis=1).astype(np.float32)
        areas = np.arange(1, n + 1, dtype=np.float32)
        nodes = np.zeros((2 * n, 4), dtype=np.uint32)
        node_area = np.zeros(2 * n, dtype=np.float32)
        count = L.ref_build_bvh(n, bb.ctypes.data, areas.ctypes.data, nodes.ctypes.data, node_area.ctypes.data, 2 * n)
        bvh_cases.append({"boxes": bb.tolist(), "areas": areas.tolist(), "nodes": nodes[:count].tolist(), "node_area": node_area[:count].tolist()})
    kats["lbvh"] = bvh_cases
    with open(os.path.join(HERE, "kats.json"), "w") as f:
        json.dump(kats, f, indent=1)

    brdf = np.zeros((128, 128), dtype=np.float32)
    albedo = np.zeros(128, dtype=np.float32)
    L.ref_kulla_conty.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.ref_kulla_conty(brdf.ctypes.data, albedo.ctypes.data)
    np.savez_compressed(os.path.join(HERE, "kulla_conty.npz"), brdf_avg=brdf, albedo_avg=albedo)


def make_reference_frames(woop, mt, only=None):
    for name, (w, h, spp) in EXACT.items():
        if only is not None and name not in only:
            continue
        pack = os.path.join(ROOT, "scenes", name + ".b200scene")
        for variant, ref in (("woop", woop), ("mt", mt)):
            frame, _, _ = ref.render_pack(pack, w, h, spp)
            np.save(os.path.join(HERE, f"exact_{name}_{variant}.npy"), frame)
            print("exact", name, variant, frame.mean())
    for name, (w, h, spp) in CONVERGED.items():
        if only is not None and name not in only:
            continue
        pack = os.path.join(ROOT, "scenes", name + ".b200scene")
        frame, _, seconds = woop.render_pack(pack, w, h, spp)
        np.save(os.path.join(HERE, f"converged_{name}.npy"), frame.astype(np.float32))
        print("converged", name, frame.mean(), f"{seconds:.1f}s")


def make_synthetic_goldens(woop, mt, only):
    """Synthetic scenes (tests/scene_builder.py): the branches the BASELINE scenes never reach."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    import scene_builder
    import __graft_entry__ as ge
    pkg = ge.load_package()
    settings_path = os.path.join(HERE, "settings.json")
    synthetic = json.load(open(settings_path)).get("synthetic", {}) if only is not None else {}
    for name, builder in scene_builder.synthetic_scenes().items():
        if only is not None and name not in only:
            continue
        desc = builder.desc()
        pack = os.path.join(HERE, f"synthetic_{name}.b200scene")
        assert pkg.lib().b200pt_scene_save(ctypes.byref(desc), pack.encode()) == 0
        for variant, ref in (("woop", woop), ("mt", mt)):
            frame, _, _ = ref.render_pack(pack, 32, 32, 4)
            np.save(os.path.join(HERE, f"exact_synthetic_{name}_{variant}.npy"), frame)
        frame, _, seconds = woop.render_pack(pack, 64, 64, 512)
        np.save(os.path.join(HERE, f"converged_synthetic_{name}.npy"), frame.astype(np.float32))
        synthetic[name] = {"exact": [32, 32, 4], "converged": [64, 64, 512]}
        print("synthetic", name, frame.mean(), f"{seconds:.1f}s")
    with open(settings_path, "w") as f:
        json.dump({"exact": EXACT, "converged": CONVERGED, "synthetic": synthetic}, f, indent=1)


if __name__ == "__main__":
    main()
