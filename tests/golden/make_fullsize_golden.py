#!/usr/bin/env python3
"""Full-size reference frames for the BASELINE configs (tests/golden/fullsize_<scene>.npz), made by the UNMODIFIED reference
build (oracle/_ref/libcsrt_ref_woop.so, csrt::Renderer::Draw on all host threads) at the configs' own resolution:

    C2 dragon             1024x1024 at 256 spp (the config's own spp)
    C3 matpreview         1024x1024 at  64 spp (of 512: u = s/spp makes frames spp-dependent, so the GPU test renders 64 too)
    C4 volumetric-caustic 1024x1024 at  64 spp (of 2048)
    C1 mercury             256x256  at  32 spp (the config's own spp)
    C5 dragon-1080p       1920x1080 at  16 spp (of 4096; bench.py's parity line of the scaling config)

stored as float16 (per-pixel rounding 5e-4 relative, far below the Monte Carlo noise the comparison has to tolerate).
Run where /root/reference's build exists: python tests/golden/make_fullsize_golden.py [scene ...]"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refcheck  # noqa: E402

# golden name -> (scene pack, width, height, spp)
FULLSIZE = {"dragon": ("dragon", 1024, 1024, 256), "matpreview": ("matpreview", 1024, 1024, 64), "volumetric-caustic": ("volumetric-caustic", 1024, 1024, 64),
            "mercury": ("mercury", 256, 256, 32), "dragon-1080p": ("dragon", 1920, 1080, 16)}

if __name__ == "__main__":
    ref = refcheck.ref_lib("woop")
    for name in (sys.argv[1:] or FULLSIZE):
        scene, w, h, spp = FULLSIZE[name]
        t0 = time.time()
        frame, _, seconds = ref.render_pack(os.path.join(ROOT, "scenes", scene + ".b200scene"), w, h, spp)
        np.savez_compressed(os.path.join(HERE, f"fullsize_{name}.npz"), frame=frame.astype(np.float16), size=np.array([w, h, spp]))
        print(name, w, h, spp, "mean", float(frame.mean()), f"render {seconds:.0f} s, total {time.time() - t0:.0f} s", flush=True)
